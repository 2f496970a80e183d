"""Coefficients of the odd polynomial atan(z) ~ z * P(z^2) on |z| <= tan(pi/8), used by the device
map evaluation (x360_common.cuh: fast_atan2).  Chebyshev interpolation of f(w) = atan(sqrt(w))/sqrt(w)
on w in [0, tan(pi/8)^2] at 50 digits, converted to the monomial basis; prints the max relative error
of the double-rounded polynomial evaluated in double arithmetic (Horner/FMA emulated with mpmath)."""
import mpmath as mp

mp.mp.dps = 60
T = mp.tan(mp.pi / 8)
W = T * T
DEG = 13          # P has degree DEG in w = z^2


def f(w):
    if w == 0:
        return mp.mpf(1)
    s = mp.sqrt(w)
    return mp.atan(s) / s


def cheb_poly_coeffs(deg):
    n = deg + 1
    nodes = [mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
    xs = [(x + 1) * W / 2 for x in nodes]
    fs = [f(x) for x in xs]
    c = [2 / mp.mpf(n) * sum(fs[k] * mp.cos(mp.pi * j * (2 * k + 1) / (2 * n)) for k in range(n)) for j in range(n)]
    c[0] /= 2
    # Chebyshev series in t = 2 w / W - 1  ->  monomials in w
    import sympy as sp
    w = sp.symbols("w")
    t = 2 * w / sp.Float(str(W), 60) - 1
    expr = sum(sp.Float(str(c[j]), 60) * sp.chebyshevt(j, t) for j in range(n))
    poly = sp.Poly(sp.expand(expr), w)
    return [mp.mpf(str(a)) for a in reversed(poly.all_coeffs())]   # a0 + a1 w + ...


if __name__ == "__main__":
    a = cheb_poly_coeffs(DEG)
    ad = [float(x) for x in a]
    worst = mp.mpf(0)
    for i in range(2001):
        z = T * i / 2000
        w = float(z) * float(z)
        p = ad[-1]
        for cco in reversed(ad[:-1]):
            p = float(mp.mpf(p) * mp.mpf(w) + mp.mpf(cco))      # one rounding per FMA
        approx = mp.mpf(float(z)) * mp.mpf(p)
        exact = mp.atan(mp.mpf(float(z)))
        if exact != 0:
            worst = max(worst, abs(approx - exact) / exact)
    print("degree", DEG, "max rel err", mp.nstr(worst, 5), "(double eps 1.1e-16)")
    for k, x in enumerate(ad):
        print(f"    {x.hex()},   // a{k} = {x!r}")

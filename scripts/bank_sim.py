"""Bank-conflict model of the tiled kernel's gather (LDS.64 of pair records) for a geometry.
Usage: python scripts/bank_sim.py  -- prints average wavefronts per LDS.64 for candidate layouts."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib
pkg = importlib.import_module("360extractor_b200")

def maps(H, W, d, fov, yaw, pitch):
    mx, my = pkg.GeometryProcessor.create_rectilinear_map(H, W, d, d, fov, yaw, pitch, 0) if hasattr(pkg.GeometryProcessor,'create_rectilinear_map_cpu') else (None,None)
    return mx, my

def host_map(H, W, d, fov, yaw, pitch):
    R = pkg.GeometryProcessor.get_rotation_matrix(yaw, pitch, 0.0)
    f = (0.5 * d) / np.tan(0.5 * np.radians(fov))
    x, y = np.meshgrid(np.arange(d), np.arange(d))
    xn = (x - d / 2) / f; yn = (y - d / 2) / f
    v = np.stack([xn, yn, np.ones_like(xn)], -1) @ R.T
    th = np.arctan2(v[..., 0], v[..., 2]); ph = np.arcsin(v[..., 1] / np.linalg.norm(v, axis=-1))
    return ((th / (2 * np.pi) + 0.5) * W).astype(np.float32), ((ph / np.pi + 0.5) * H).astype(np.float32)

def wavefronts(addr8):
    """addr8: byte addresses (16 lanes) of 8-byte loads -> wavefronts for that half-warp phase."""
    words = np.unique(addr8 // 8)
    banks = (words * 2) % 32          # first bank of the pair
    return np.bincount(banks // 2, minlength=16).max()

def simulate(mx, my, layout, tiles):
    d = mx.shape[0]
    sx = np.rint(mx * 32).astype(np.int64); sy = np.rint(my * 32).astype(np.int64)
    ix = sx >> 5; iy = sy >> 5
    tot = 0; n = 0
    for (ty, tx) in tiles:
        X = ix[ty*32:ty*32+32, tx*32:tx*32+32]; Y = iy[ty*32:ty*32+32, tx*32:tx*32+32]
        xmin, xmax, ymin, ymax = X.min(), X.max(), Y.min(), Y.max()
        if xmax - xmin > 200: continue
        x_org = xmin & ~3
        npx = xmax - x_org + 1
        delta = (3 * x_org) % 16
        bw = -(-(delta + 3 * (npx + 1)) // 48) * 48
        for r in range(32):
            for h in (0, 16):
                rx = X[r, h:h+16] - x_org; ry = Y[r, h:h+16] - ymin
                tot += wavefronts(layout(rx, ry, bw)); n += 1
    return tot / n

def lay_pad144(rx, ry, bw):   return ry * 3 * bw + (rx >> 4) * 144 + (rx & 15) * 8
def lay_contig(rx, ry, bw):   return ry * (8 * (bw // 3) + 0) + rx * 8
def lay_contig_p8(rx, ry, bw):  return ry * (8 * (bw // 3) + 8) + rx * 8
def lay_contig_p16(rx, ry, bw): return ry * (8 * (bw // 3) + 16) + rx * 8
def lay_pad144_p(rx, ry, bw, extra): return ry * (3 * bw + extra) + (rx >> 4) * 144 + (rx & 15) * 8

if __name__ == "__main__":
    H, W, d, fov = 2880, 5760, 1600, 90
    mx, my = host_map(H, W, d, fov, 45.0, 0.0)
    rng = np.random.default_rng(0)
    tiles = [(int(a), int(b)) for a, b in rng.integers(0, 50, (300, 2))]
    for name, lay in [("pad144 pitch 3bw", lay_pad144), ("contig pitch 8*px", lay_contig), ("contig +8", lay_contig_p8),
                      ("contig +16", lay_contig_p16)] + [(f"pad144 pitch 3bw+{e}", (lambda e: lambda rx, ry, bw: lay_pad144_p(rx, ry, bw, e))(e)) for e in (16, 32, 48, 64, 80)]:
        print(f"{name:24s} {simulate(mx, my, lay, tiles):.3f} wavefronts per half-warp phase (ideal 1)")

def sweep():
    H, W, d, fov = 2880, 5760, 1600, 90
    mx, my = host_map(H, W, d, fov, 45.0, 0.0)
    sx = np.rint(mx * 32).astype(np.int64); sy = np.rint(my * 32).astype(np.int64)
    ix = sx >> 5; iy = sy >> 5
    rng = np.random.default_rng(1)
    tiles = [(int(a), int(b)) for a, b in rng.integers(0, 50, (200, 2))]
    Ps = list(range(0, 128, 8))
    for name, rec in (("contig", lambda rx: rx * 8), ("pad144", lambda rx: (rx >> 4) * 144 + (rx & 15) * 8)):
        res = np.zeros((len(tiles), len(Ps))); slope = np.zeros(len(tiles))
        for t, (ty, tx) in enumerate(tiles):
            X = ix[ty*32:ty*32+32, tx*32:tx*32+32]; Y = iy[ty*32:ty*32+32, tx*32:tx*32+32]
            x_org = X.min() & ~3; ymin = Y.min()
            slope[t] = (Y[:, -1] - Y[:, 0]).mean()
            for p, P in enumerate(Ps):
                tot = 0
                for r in range(32):
                    for h in (0, 16):
                        tot += wavefronts((Y[r, h:h+16] - ymin) * (1024 + P) + rec(X[r, h:h+16] - x_org))
                res[t, p] = tot / 64
        print(name, "fixed P:", " ".join(f"{P}:{res[:, p].mean():.2f}" for p, P in enumerate(Ps)))
        print(name, "per-tile best:", res.min(1).mean())
        for Ppos, Pneg in ((8, 120), (16, 112), (24, 104), (32, 96), (120, 8), (112, 16), (104, 24), (96, 32)):
            sel = np.where(slope >= 0, res[:, Ps.index(Ppos)], res[:, Ps.index(Pneg)])
            print(f"   slope>=0 -> P={Ppos}, else P={Pneg}: {sel.mean():.3f}")

if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "sweep":
    sweep()

def wf32(addr4):
    words = np.unique(addr4 // 4)
    return np.bincount(words % 32, minlength=32).max()

def sweep_bgrx():
    H, W, d, fov = 2880, 5760, 1600, 90
    mx, my = host_map(H, W, d, fov, 45.0, 0.0)
    sx = np.rint(mx * 32).astype(np.int64); sy = np.rint(my * 32).astype(np.int64)
    ix = sx >> 5; iy = sy >> 5
    rng = np.random.default_rng(1)
    tiles = [(int(a), int(b)) for a, b in rng.integers(0, 50, (200, 2))]
    for P in range(0, 128, 4):
        tot = 0; n = 0
        for (ty, tx) in tiles:
            X = ix[ty*32:ty*32+32, tx*32:tx*32+32]; Y = iy[ty*32:ty*32+32, tx*32:tx*32+32]
            x_org = X.min() & ~3; ymin = Y.min()
            for r in range(32):
                a = (Y[r] - ymin) * (1024 + P) + (X[r] - x_org) * 4
                tot += wf32(a) + wf32(a + 4); n += 2
        print(f"BGRX P={P:3d}: {tot / n:.3f} wavefronts per LDS.32 (ideal 1)")

if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "bgrx":
    sweep_bgrx()

def sweep_raw():
    H, W, d, fov = 2880, 5760, 1600, 90
    mx, my = host_map(H, W, d, fov, 45.0, 0.0)
    sx = np.rint(mx * 32).astype(np.int64); sy = np.rint(my * 32).astype(np.int64)
    ix = sx >> 5; iy = sy >> 5
    rng = np.random.default_rng(1)
    tiles = [(int(a), int(b)) for a, b in rng.integers(0, 50, (200, 2))]
    for mode in ("min16", "P0", "P16", "P32", "P48", "P64", "P80", "P96", "P112"):
        tot = 0; n = 0; bytes_ = 0
        for (ty, tx) in tiles:
            X = ix[ty*32:ty*32+32, tx*32:tx*32+32]; Y = iy[ty*32:ty*32+32, tx*32:tx*32+32]
            xmin = X.min(); ymin = Y.min()
            gx0 = (3 * xmin) & ~15; delta = 3 * xmin - gx0
            need = delta + 3 * (X.max() - xmin + 2)
            bw = -(-need // 16) * 16
            if mode != "min16":
                P = int(mode[1:])
                while bw % 128 != P: bw += 16
            bytes_ += bw * (Y.max() - ymin + 2)
            for r in range(32):
                o = (Y[r] - ymin) * bw + delta + 3 * (X[r] - xmin)
                a = o & ~3
                for k in range(3):
                    tot += wf32(a + 4 * k); n += 1
        print(f"raw gather {mode:6s}: {tot / n:.3f} wavefronts per LDS.32; staged bytes/tile {bytes_ / len(tiles):.0f}")

if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "raw":
    sweep_raw()

def rules():
    H, W, d, fov = 2880, 5760, 1600, 90
    mx, my = host_map(H, W, d, fov, 45.0, 0.0)
    sx = np.rint(mx * 32).astype(np.int64); sy = np.rint(my * 32).astype(np.int64)
    ix = sx >> 5; iy = sy >> 5
    rng = np.random.default_rng(2)
    tiles = [(int(a), int(b)) for a, b in rng.integers(0, 50, (250, 2))]
    rec = lambda rx: (rx >> 4) * 144 + (rx & 15) * 8
    out = {k: 0.0 for k in ("P0", "rule c in -1,0,1", "rule c=+-1 always", "floor16 skew", "oracle c in -1,0,1")}
    n = 0
    for (ty, tx) in tiles:
        X = ix[ty*32:ty*32+32, tx*32:tx*32+32]; Y = iy[ty*32:ty*32+32, tx*32:tx*32+32]
        x_org = X.min() & ~3; ymin = Y.min()
        s = (X[:, -1] - X[:, 0]).mean() / 31.0
        slope = (Y[:, -1] - Y[:, 0]).mean()
        sg = 1 if slope > 0 else (-1 if slope < 0 else 0)
        def run(rowoff):
            tot = 0
            for r in range(32):
                for h in (0, 16):
                    ry = Y[r, h:h+16] - ymin
                    tot += wavefronts(ry * 1024 + rowoff(ry) + rec(X[r, h:h+16] - x_org) + 4096)
            return tot / 64
        res = {c: run(lambda ry, c=c: 8 * c * ry) for c in (-1, 0, 1)}
        out["P0"] += res[0]
        out["rule c in -1,0,1"] += res[0] if (s > 0.93 or sg == 0) else res[sg]
        out["rule c=+-1 always"] += res[sg]
        out["oracle c in -1,0,1"] += min(res.values())
        out["floor16 skew"] += res[0] if (s > 0.93 or sg == 0) else run(lambda ry: 16 * ((sg * ry) >> 1))
        n += 1
    for k, v in out.items(): print(f"{k:24s} {v / n:.3f}")

if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "rules":
    rules()

def planes():
    """Split planes: lo words [B0 B1 G0 G1] 4 B per source pixel (LDS.32), hi halfwords [R0 R1] 2 B (LDS.U16)."""
    H, W, d, fov = 2880, 5760, 1600, 90
    mx, my = host_map(H, W, d, fov, 45.0, 0.0)
    sx = np.rint(mx * 32).astype(np.int64); sy = np.rint(my * 32).astype(np.int64)
    ix = sx >> 5; iy = sy >> 5
    rng = np.random.default_rng(3)
    tiles = [(int(a), int(b)) for a, b in rng.integers(0, 50, (250, 2))]
    tot_lo = tot_hi = tot_64 = 0; n = 0
    for (ty, tx) in tiles:
        X = ix[ty*32:ty*32+32, tx*32:tx*32+32]; Y = iy[ty*32:ty*32+32, tx*32:tx*32+32]
        x_org = X.min() & ~3; ymin = Y.min()
        npx = X.max() - x_org + 1
        delta = (3 * x_org) % 16
        bw = -(-(delta + 3 * (npx + 1)) // 48) * 48
        gpr = bw // 12
        for r in range(32):
            rx = X[r] - x_org; ry = Y[r] - ymin + 1           # the lower record (the one loaded per step)
            tot_lo += wf32(ry * 16 * gpr + rx * 4)
            tot_hi += wf32(ry * 8 * gpr + rx * 2)
            tot_64 += wavefronts(ry[:16] * 3 * bw + (rx[:16] >> 4) * 144 + (rx[:16] & 15) * 8) + \
                      wavefronts(ry[16:] * 3 * bw + (rx[16:] >> 4) * 144 + (rx[16:] & 15) * 8)
            n += 1
    print(f"per warp row-load: pair records LDS.64 {tot_64 / n:.2f} wavefronts; planes LDS.32 {tot_lo / n:.2f} + LDS.U16 {tot_hi / n:.2f} = {(tot_lo + tot_hi) / n:.2f}")

if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "planes":
    planes()

def twin():
    """Twin-copy raw gather: the footprint is staged twice (byte shifts 0 and 2) so that every pixel pair's 6 bytes
    lie in two aligned words of one copy; no transform.  Wavefronts of the two LDS.32 per record."""
    H, W, d, fov = 2880, 5760, 1600, 90
    mx, my = host_map(H, W, d, fov, 45.0, 0.0)
    sx = np.rint(mx * 32).astype(np.int64); sy = np.rint(my * 32).astype(np.int64)
    ix = sx >> 5; iy = sy >> 5
    rng = np.random.default_rng(3)
    tiles = [(int(a), int(b)) for a, b in rng.integers(0, 50, (250, 2))]
    for gran in (16, 32, 64, 128):
        for cs_off in (0, 16, 32, 64):
            tot = 0; n = 0; nbytes = 0
            for (ty, tx) in tiles:
                X = ix[ty*32:ty*32+32, tx*32:tx*32+32]; Y = iy[ty*32:ty*32+32, tx*32:tx*32+32]
                x_org = X.min(); ymin = Y.min(); rows = Y.max() - ymin + 2
                gx0 = (3 * x_org) & ~15; delta = 3 * x_org - gx0
                need = delta + 3 * (X.max() - x_org + 2) + 2
                bw = -(-need // gran) * gran
                CS = -(-(rows * bw) // 128) * 128 + cs_off
                nbytes += 2 * rows * bw
                for r in range(32):
                    p = delta + 3 * (X[r] - x_org)
                    copy = (p & 3) >> 1
                    q = p + 2 * copy
                    a = copy * CS + (Y[r] - ymin + 1) * bw + (q & ~3)
                    tot += wf32(a) + wf32(a + 4); n += 1
            print(f"twin gran {gran:3d} copy offset {cs_off:3d}: {tot / n:.2f} wavefronts per record (2 LDS.32); TMA bytes/tile {nbytes / len(tiles):.0f}")

if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "twin":
    twin()

def planes_skew():
    """Split planes with a per-row skew of c words (lo) / c halfwords (hi): wavefronts of the per-step record load."""
    H, W, d, fov = 2880, 5760, 1600, 90
    mx, my = host_map(H, W, d, fov, 45.0, 0.0)
    sx = np.rint(mx * 32).astype(np.int64); sy = np.rint(my * 32).astype(np.int64)
    ix = sx >> 5; iy = sy >> 5
    rng = np.random.default_rng(3)
    tiles = [(int(a), int(b)) for a, b in rng.integers(0, 50, (250, 2))]
    res = {}
    for c in (-3, -2, -1, 0, 1, 2, 3):
        per_tile = []
        for (ty, tx) in tiles:
            X = ix[ty*32:ty*32+32, tx*32:tx*32+32]; Y = iy[ty*32:ty*32+32, tx*32:tx*32+32]
            x_org = X.min() & ~3; ymin = Y.min()
            npx = X.max() - x_org + 1
            delta = (3 * x_org) % 16
            bw = -(-(delta + 3 * (npx + 1)) // 48) * 48
            gpr = bw // 12
            lo = hi = 0
            for r in range(32):
                rx = X[r] - x_org; ry = Y[r] - ymin + 1
                lo += wf32(ry * (16 * gpr + 4 * c) + rx * 4)
                hi += wf32(ry * (8 * gpr + 2 * c) + rx * 2)
            slope = np.sign((Y[:, -1] - Y[:, 0]).mean())
            per_tile.append((lo / 32, hi / 32, slope))
        res[c] = per_tile
    for c, v in res.items():
        print(f"skew {c:+d}: lo {np.mean([a for a, _, _ in v]):.2f} hi {np.mean([b for _, b, _ in v]):.2f}")
    best = np.mean([min(res[c][i][0] + res[c][i][1] for c in res) for i in range(len(tiles))])
    print(f"oracle per-tile choice: {best:.2f}")
    for cp, cn in ((1, -1), (-1, 1), (2, -2), (-2, 2), (1, 0), (0, -1)):
        tot = np.mean([(res[cp][i][0] + res[cp][i][1]) if res[0][i][2] >= 0 else (res[cn][i][0] + res[cn][i][1]) for i in range(len(tiles))])
        print(f"rule slope>=0 -> {cp:+d}, else {cn:+d}: {tot:.2f}")

if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "planes_skew":
    planes_skew()

def planes_abs():
    """Split planes with the row pitch forced to a residue: lo pitch = t (mod 32 words), hi pitch = t (mod 64 halfwords)."""
    H, W, d, fov = 2880, 5760, 1600, 90
    mx, my = host_map(H, W, d, fov, 45.0, 0.0)
    sx = np.rint(mx * 32).astype(np.int64); sy = np.rint(my * 32).astype(np.int64)
    ix = sx >> 5; iy = sy >> 5
    rng = np.random.default_rng(3)
    tiles = [(int(a), int(b)) for a, b in rng.integers(0, 50, (250, 2))]
    res = {}
    for t in (None, 0, 1, -1, 2, -2, 3, -3, 4, -4, 8, -8):
        per_tile = []
        for (ty, tx) in tiles:
            X = ix[ty*32:ty*32+32, tx*32:tx*32+32]; Y = iy[ty*32:ty*32+32, tx*32:tx*32+32]
            x_org = X.min() & ~3; ymin = Y.min()
            npx = X.max() - x_org + 1
            delta = (3 * x_org) % 16
            bw = -(-(delta + 3 * (npx + 1)) // 48) * 48
            gpr = bw // 12
            plo, phi = 4 * gpr, 4 * gpr                    # words / halfwords per row
            if t is not None:
                plo += (t - plo) % 32
                phi += (t - phi) % 64
            lo = hi = 0
            for r in range(32):
                rx = X[r] - x_org; ry = Y[r] - ymin + 1
                lo += wf32((ry * plo + rx) * 4)
                hi += wf32((ry * phi + rx) * 2)
            slope = np.sign((Y[:, -1] - Y[:, 0]).mean())
            per_tile.append((lo / 32, hi / 32, slope))
        res[t] = per_tile
        print(f"pitch residue {t}: lo {np.mean([a for a, _, _ in per_tile]):.2f} hi {np.mean([b for _, b, _ in per_tile]):.2f}")
    for cp, cn in ((1, -1), (-1, 1), (2, -2), (-2, 2), (1, 0), (0, -1)):
        tot = np.mean([(res[cp][i][0] + res[cp][i][1]) if res[0][i][2] >= 0 else (res[cn][i][0] + res[cn][i][1]) for i in range(len(tiles))])
        print(f"rule slope>=0 -> {cp:+d}, else {cn:+d}: {tot:.2f}")
    print("oracle:", np.mean([min(res[t][i][0] + res[t][i][1] for t in res) for i in range(len(tiles))]))
    for cp, cn in ((1, -1), (2, -2), (4, -4), (8, -8)):
        lo = np.mean([res[cp][i][0] if res[0][i][2] >= 0 else res[cn][i][0] for i in range(len(tiles))])
        hi = np.mean([res[cp][i][1] if res[0][i][2] >= 0 else res[cn][i][1] for i in range(len(tiles))])
        print(f"rule {cp:+d}/{cn:+d}: lo {lo:.2f} hi {hi:.2f}; lo by rule + hi natural: {lo + np.mean([v[1] for v in res[None]]):.2f}")

if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "planes_abs":
    planes_abs()

/*
 * x360_oracle.c — CPU ORACLE. TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the arithmetic on the reference's equirect ->
 * rectilinear hot path.  Nothing in the product package may import, link or
 * execute this file: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and there only as the checker.
 *
 * What it restates (file:line relative to /root/reference):
 *   orc_focal, orc_make_map   src/core/geometry.py:122-192  (create_rectilinear_map)
 *   orc_remap_linear          src/core/processor.py:334 with cv2.INTER_LINEAR,
 *                             borderMode=cv2.BORDER_WRAP  (interp select :199-200)
 *   orc_remap_lanczos4        same call with cv2.INTER_LANCZOS4
 *   orc_blur_sums             src/utils/image_utils.py:6-27 (calculate_blur_score)
 *   orc_gauss13_taps, orc_unsharp   src/core/processor.py:379-381 (= src/ui/preview_widget.py:118-120):
 *                             cv2.GaussianBlur(img, (0, 0), 2.0) + cv2.addWeighted(img, 1 + s, blur, -s, 0)
 *
 * The pixel arithmetic itself lives in a third-party dependency that is NOT
 * under /root/reference: opencv-python (requirements.txt:2 ">=4.8.0", no
 * lockfile; 4.13.0 installed in this image) and numpy (requirements.txt:3).
 * This file restates OpenCV's published fixed-point remap scheme
 * (INTER_BITS=5, INTER_REMAP_COEF_BITS=15).
 *
 * Parity pinning: the reference's own tests hold no pixel-level golden vectors
 * for this path (tests/test_core.py:67-78 checks map shape/dtype only), so the
 * oracle is pinned against outputs of the reference itself run in the build
 * container: tests/golden/make_golden.py imports /root/reference/src and calls
 * cv2.remap / calculate_blur_score exactly as processor.py does; the committed
 * fixtures under tests/golden/ are checked against this file by
 * tests/test_oracle.py (bit-exact for both interpolations, scores to 1e-12 rel).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -shared; the omp pragmas are inert, libgomp is not in the image).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define ORC_INTER_BITS 5
#define ORC_TAB 32               /* INTER_TAB_SIZE */
#define ORC_COEF_BITS 15         /* INTER_REMAP_COEF_BITS */
#define ORC_SCALE 32768

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* geometry.py:141 — f = (0.5*dest_w) / tan(0.5*radians(fov_deg));
 * numpy's radians is x * (pi/180). */
double orc_focal(int dest_w, double fov_deg)
{
    double rad = fov_deg * (M_PI / 180.0);
    return (0.5 * (double)dest_w) / tan(0.5 * rad);
}

/* geometry.py:141-190.  R is the row-major fp64 3x3 from get_rotation_matrix
 * (geometry.py:80-119), built by the caller with numpy so that the residues
 * (sin(pi)=1.22e-16 ...) that decide the +-pi seam are the reference's own.
 * v_rot = [xn, yn, 1] @ R.T  (geometry.py:164). */
void orc_make_map(int src_h, int src_w, int dest_h, int dest_w, double fov_deg,
                  const double *R, float *map_x, float *map_y)
{
    const double f = orc_focal(dest_w, fov_deg);
    const double cx = dest_w / 2.0, cy = dest_h / 2.0;      /* geometry.py:142 */
    const double two_pi = 2.0 * M_PI;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < dest_h; ++y) {
        const double yn = ((double)y - cy) / f;               /* geometry.py:150 */
        for (int x = 0; x < dest_w; ++x) {
            const double xn = ((double)x - cx) / f;           /* geometry.py:149 */
            const double v0 = xn * R[0] + yn * R[1] + R[2];
            const double v1 = xn * R[3] + yn * R[4] + R[5];
            const double v2 = xn * R[6] + yn * R[7] + R[8];
            const double theta = atan2(v0, v2);               /* geometry.py:172 */
            const double r = sqrt(v0 * v0 + v1 * v1 + v2 * v2);
            const double phi = asin(v1 / r);                  /* geometry.py:176 */
            const double uf = (theta / two_pi + 0.5) * (double)src_w;   /* :183 */
            const double vf = (phi / M_PI + 0.5) * (double)src_h;       /* :186 */
            map_x[(size_t)y * dest_w + x] = (float)uf;        /* :189 */
            map_y[(size_t)y * dest_w + x] = (float)vf;        /* :190 */
        }
    }
}

/* cv2 BORDER_WRAP index (borderInterpolate): true modulo on both axes. */
static inline int orc_wrap(int p, int len)
{
    if ((unsigned)p < (unsigned)len) return p;
    p %= len;
    if (p < 0) p += len;
    return p;
}

static inline int orc_sat_s16(int v)
{
    return v < -32768 ? -32768 : (v > 32767 ? 32767 : v);
}

static inline uint8_t orc_sat_u8(int v)
{
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

/* Q5 quantisation of one float32 map coordinate: cvRound(m * 32) is
 * round-half-to-even of the fp32 product (exact: a power-of-two scale). */
static inline int orc_q5(float m)
{
    return (int)lrintf(m * (float)ORC_TAB);   /* default rounding mode = nearest-even */
}

/* SURVEY.md A-2.  src: uint8 [H][W][3]; dst: uint8 [dh][dw][3]. */
void orc_remap_linear(const uint8_t *src, int H, int W,
                      const float *map_x, const float *map_y,
                      int dh, int dw, uint8_t *dst)
{
#pragma omp parallel for schedule(static)
    for (int y = 0; y < dh; ++y) {
        for (int x = 0; x < dw; ++x) {
            const size_t i = (size_t)y * dw + x;
            const int sx = orc_q5(map_x[i]), sy = orc_q5(map_y[i]);
            const int fx = sx & (ORC_TAB - 1), fy = sy & (ORC_TAB - 1);
            const int ix = orc_sat_s16(sx >> ORC_INTER_BITS);
            const int iy = orc_sat_s16(sy >> ORC_INTER_BITS);
            const int x0 = orc_wrap(ix, W), x1 = orc_wrap(ix + 1, W);
            const int y0 = orc_wrap(iy, H), y1 = orc_wrap(iy + 1, H);
            /* 2-D Q15 weights: rint(f32(b*a)*32768) with a,b multiples of 1/32
             * is exactly (32-fy)(32-fx)*32 ... ; sum is always 32768. */
            const int w00 = (ORC_TAB - fy) * (ORC_TAB - fx) * 32;
            const int w01 = (ORC_TAB - fy) * fx * 32;
            const int w10 = fy * (ORC_TAB - fx) * 32;
            const int w11 = fy * fx * 32;
            const uint8_t *p00 = src + ((size_t)y0 * W + x0) * 3;
            const uint8_t *p01 = src + ((size_t)y0 * W + x1) * 3;
            const uint8_t *p10 = src + ((size_t)y1 * W + x0) * 3;
            const uint8_t *p11 = src + ((size_t)y1 * W + x1) * 3;
            for (int c = 0; c < 3; ++c) {
                const int acc = w00 * p00[c] + w01 * p01[c] + w10 * p10[c] + w11 * p11[c];
                dst[i * 3 + c] = orc_sat_u8((acc + (1 << (ORC_COEF_BITS - 1))) >> ORC_COEF_BITS);
            }
        }
    }
}

/* 1-D Lanczos-4 coefficients for t = k/32 (float32 results,
 * fp64 trig, float32 running sum, float32 reciprocal normalisation). */
static void orc_lanczos4_coeffs(float t, float coef[8])
{
    static const double s45 = 0.70710678118654752440;
    static const double cs[8][2] = {
        {1, 0}, {-s45, -s45}, {0, 1}, {s45, -s45}, {-1, 0}, {s45, s45}, {0, -1}, {-s45, s45}};
    if (t < FLT_EPSILON) {
        for (int i = 0; i < 8; ++i) coef[i] = 0.f;
        coef[3] = 1.f;
        return;
    }
    float sum = 0.f;
    const double y0 = -(t + 3) * M_PI * 0.25, s0 = sin(y0), c0 = cos(y0);
    for (int i = 0; i < 8; ++i) {
        const double y = -(t + 3 - i) * M_PI * 0.25;
        coef[i] = (float)((cs[i][0] * s0 + cs[i][1] * c0) / (y * y));
        sum += coef[i];
    }
    sum = 1.f / sum;
    for (int i = 0; i < 8; ++i) coef[i] *= sum;
}

/* The [32*32][8][8] int16 table, entry index = fy*32 + fx. */
void orc_lanczos4_table(int16_t *tab)
{
    float c1[ORC_TAB][8];
    for (int k = 0; k < ORC_TAB; ++k)
        orc_lanczos4_coeffs((float)k * (1.f / ORC_TAB), c1[k]);
    for (int fy = 0; fy < ORC_TAB; ++fy) {
        for (int fx = 0; fx < ORC_TAB; ++fx) {
            int16_t *it = tab + (size_t)(fy * ORC_TAB + fx) * 64;
            int isum = 0;
            for (int r = 0; r < 8; ++r) {
                for (int c = 0; c < 8; ++c) {
                    const float v = c1[fy][r] * c1[fx][c];
                    const int q = orc_sat_s16((int)lrintf(v * (float)ORC_SCALE));
                    it[r * 8 + c] = (int16_t)q;
                    isum += q;
                }
            }
            if (isum != ORC_SCALE) {
                const int diff = isum - ORC_SCALE;
                const int k0 = 4, k1 = 6;       /* ksize/2 .. ksize/2+1 inclusive */
                int Mk1 = k0, Mk2 = k0, mk1 = k0, mk2 = k0;
                for (int r = k0; r < k1; ++r) {
                    for (int c = k0; c < k1; ++c) {
                        if (it[r * 8 + c] < it[mk1 * 8 + mk2]) { mk1 = r; mk2 = c; }
                        else if (it[r * 8 + c] > it[Mk1 * 8 + Mk2]) { Mk1 = r; Mk2 = c; }
                    }
                }
                if (diff < 0) it[Mk1 * 8 + Mk2] = (int16_t)(it[Mk1 * 8 + Mk2] - diff);
                else          it[mk1 * 8 + mk2] = (int16_t)(it[mk1 * 8 + mk2] - diff);
            }
        }
    }
}

void orc_remap_lanczos4(const uint8_t *src, int H, int W,
                        const float *map_x, const float *map_y,
                        int dh, int dw, uint8_t *dst)
{
    int16_t *tab = (int16_t *)malloc(sizeof(int16_t) * ORC_TAB * ORC_TAB * 64);
    orc_lanczos4_table(tab);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < dh; ++y) {
        int xs[8], ys[8];
        for (int x = 0; x < dw; ++x) {
            const size_t i = (size_t)y * dw + x;
            const int sx = orc_q5(map_x[i]), sy = orc_q5(map_y[i]);
            const int fx = sx & (ORC_TAB - 1), fy = sy & (ORC_TAB - 1);
            const int ix = orc_sat_s16(sx >> ORC_INTER_BITS) - 3;
            const int iy = orc_sat_s16(sy >> ORC_INTER_BITS) - 3;
            const int16_t *it = tab + (size_t)(fy * ORC_TAB + fx) * 64;
            for (int k = 0; k < 8; ++k) { xs[k] = orc_wrap(ix + k, W); ys[k] = orc_wrap(iy + k, H); }
            int acc[3] = {0, 0, 0};
            for (int r = 0; r < 8; ++r) {
                const uint8_t *row = src + (size_t)ys[r] * W * 3;
                for (int c = 0; c < 8; ++c) {
                    const int w = it[r * 8 + c];
                    const uint8_t *p = row + (size_t)xs[c] * 3;
                    acc[0] += w * p[0]; acc[1] += w * p[1]; acc[2] += w * p[2];
                }
            }
            for (int c = 0; c < 3; ++c)
                dst[i * 3 + c] = orc_sat_u8((acc[c] + (1 << (ORC_COEF_BITS - 1))) >> ORC_COEF_BITS);
        }
    }
    free(tab);
}

/* SURVEY.md A-4 / image_utils.py:22-27.  ch = 3 (BGR -> gray via the Q15
 * 3735/19235/9798 weights, cv2.COLOR_BGR2GRAY) or 1 (already gray).
 * L = 4-neighbour Laplacian with BORDER_REFLECT_101; returns sum L, sum L^2. */
static inline int orc_reflect101(int p, int len)
{
    if (len == 1) return 0;
    if (p < 0) p = -p;
    if (p >= len) p = 2 * len - 2 - p;
    return p;
}

void orc_blur_sums(const uint8_t *img, int h, int w, int ch, int64_t *sum, int64_t *sum2)
{
    uint8_t *g = (uint8_t *)malloc((size_t)h * w);
    if (ch == 3) {
        for (size_t i = 0; i < (size_t)h * w; ++i) {
            const uint8_t *p = img + i * 3;
            g[i] = (uint8_t)((3735 * p[0] + 19235 * p[1] + 9798 * p[2] + 16384) >> 15);
        }
    } else {
        memcpy(g, img, (size_t)h * w);
    }
    int64_t s = 0, s2 = 0;
#pragma omp parallel for schedule(static) reduction(+ : s, s2)
    for (int y = 0; y < h; ++y) {
        const uint8_t *r0 = g + (size_t)orc_reflect101(y - 1, h) * w;
        const uint8_t *r1 = g + (size_t)y * w;
        const uint8_t *r2 = g + (size_t)orc_reflect101(y + 1, h) * w;
        for (int x = 0; x < w; ++x) {
            const int xl = orc_reflect101(x - 1, w), xr = orc_reflect101(x + 1, w);
            const int L = r0[x] + r2[x] + r1[xl] + r1[xr] - 4 * r1[x];
            s += L;
            s2 += (int64_t)L * L;
        }
    }
    free(g);
    *sum = s;
    *sum2 = s2;
}

/* population variance (ndarray.var, ddof=0) from the two exact integer sums */
double orc_score_from_sums(int64_t sum, int64_t sum2, int64_t n)
{
    if (n <= 0) return 0.0;
    const double mean = (double)sum / (double)n;
    return (double)sum2 / (double)n - mean * mean;
}

/* The batched entry the CUDA library replaces (processor.py:327-342 body for
 * all frames x views): frames [F][H][W][3], R [V][9], out [F][V][d][d][3],
 * sums nullable [F][V].  interp: 0 = linear, 1 = lanczos4. */
void orc_extract(const uint8_t *frames, int F, int H, int W, const double *R, int V,
                 int d, double fov_deg, int interp, uint8_t *out,
                 int64_t *sum_lap, int64_t *sum_lap2)
{
    float *mx = (float *)malloc(sizeof(float) * (size_t)d * d);
    float *my = (float *)malloc(sizeof(float) * (size_t)d * d);
    const size_t fsz = (size_t)H * W * 3, vsz = (size_t)d * d * 3;
    for (int v = 0; v < V; ++v) {
        orc_make_map(H, W, d, d, fov_deg, R + 9 * v, mx, my);
        for (int f = 0; f < F; ++f) {
            uint8_t *o = out + ((size_t)f * V + v) * vsz;
            if (interp == 1) orc_remap_lanczos4(frames + f * fsz, H, W, mx, my, d, d, o);
            else             orc_remap_linear(frames + f * fsz, H, W, mx, my, d, d, o);
            if (sum_lap && sum_lap2)
                orc_blur_sums(o, d, d, 3, sum_lap + (size_t)f * V + v, sum_lap2 + (size_t)f * V + v);
        }
    }
    free(mx);
    free(my);
}

/* ---- unsharp mask --------------------------------------------------
 * processor.py:379-381:  gaussian = cv2.GaussianBlur(rect_img, (0, 0), 2.0)
 *                        rect_img = cv2.addWeighted(rect_img, 1.0 + s, gaussian, -s, 0)
 * OpenCV's 8-bit GaussianBlur is fixed point (smooth.dispatch.cpp, "GaussianBlurFixedPoint"):
 *   ksize = cvRound(sigma * 6 + 1) | 1 = 13 for sigma = 2; the double kernel exp(-x^2 / (2 sigma^2)) / sum
 *   is quantised to 8 fractional bits with error diffusion from the outside in, the centre tap takes
 *   what is left of 256 (getGaussianKernelFixedPoint_ED) -> {1,2,7,16,31,45,52,45,31,16,7,2,1};
 *   rows then columns without intermediate rounding, BORDER_REFLECT_101 (repeated for tiny images):
 *   blur = (sum_j sum_i k_j k_i src + 2^15) >> 16.
 * addWeighted on uint8 runs in float32: dst = sat_u8(rint(fma(src, f32(1 + s), f32(blur * f32(-s)))))
 * (round half to even).  Both verified against cv2 4.13.0: 0 differing bytes, every (src, blur) byte
 * pair x 10 strengths, images from 1x1 to 1023x1023 (tests/golden/make_golden.py, tests/test_oracle.py). */
void orc_gauss13_taps(double sigma, int *taps /* [13] */)
{
    const int n = 13;
    double k[13], sum = 0.0;
    for (int i = 0; i < n; ++i) {
        const double x = i - (n - 1) * 0.5;
        k[i] = exp(-0.5 / (sigma * sigma) * x * x);
        sum += k[i];
    }
    double err = 0.0;
    int acc = 0;
    for (int i = 0; i < n / 2; ++i) {
        const double adj = k[i] / sum * 256.0 + err;
        const int v = (int)nearbyint(adj);
        err = adj - v;
        taps[i] = taps[n - 1 - i] = v;
        acc += 2 * v;
    }
    taps[n / 2] = 256 - acc;
}

static inline int orc_reflect101_any(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;
    return p;
}

/* cv2.addWeighted(a, 1 + s, g, -s, 0) on n bytes */
void orc_add_weighted(const uint8_t *a, const uint8_t *g, size_t n, double strength, uint8_t *out)
{
    const float alpha = (float)(1.0 + strength), beta = (float)(-strength);
    for (size_t i = 0; i < n; ++i) {
        const float t = (float)g[i] * beta;                        /* rounded to float32 */
        const float q = nearbyintf(fmaf((float)a[i], alpha, t));   /* half to even */
        out[i] = (uint8_t)(q < 0.f ? 0 : (q > 255.f ? 255 : (int)q));
    }
}

/* img, out: uint8 [h][w][ch] (out != img) */
void orc_unsharp(const uint8_t *img, int h, int w, int ch, double strength, uint8_t *out)
{
    int k[13];
    orc_gauss13_taps(2.0, k);
    const float alpha = (float)(1.0 + strength), beta = (float)(-strength);
    const size_t pitch = (size_t)w * ch;
    int32_t *H = (int32_t *)malloc(sizeof(int32_t) * (size_t)h * pitch);
    int *xi = (int *)malloc(sizeof(int) * (size_t)w * 13), *yi = (int *)malloc(sizeof(int) * (size_t)h * 13);
    for (int x = 0; x < w; ++x) for (int i = 0; i < 13; ++i) xi[x * 13 + i] = orc_reflect101_any(x + i - 6, w);
    for (int y = 0; y < h; ++y) for (int j = 0; j < 13; ++j) yi[y * 13 + j] = orc_reflect101_any(y + j - 6, h);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
            for (int c = 0; c < ch; ++c) {
                int32_t a = 0;
                for (int i = 0; i < 13; ++i) a += k[i] * img[y * pitch + (size_t)xi[x * 13 + i] * ch + c];
                H[y * pitch + (size_t)x * ch + c] = a;
            }
    for (int y = 0; y < h; ++y)
        for (size_t b = 0; b < pitch; ++b) {
            int32_t v = 0;
            for (int j = 0; j < 13; ++j) v += k[j] * H[(size_t)yi[y * 13 + j] * pitch + b];
            const int g = (v + 32768) >> 16;
            const float t = (float)g * beta;                       /* rounded to float32 */
            const float r = fmaf((float)img[y * pitch + b], alpha, t);
            const float q = nearbyintf(r);                         /* half to even */
            out[y * pitch + b] = (uint8_t)(q < 0.f ? 0 : (q > 255.f ? 255 : (int)q));
        }
    free(H); free(xi); free(yi);
}

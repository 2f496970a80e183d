// x360_common.cuh — shared device helpers for the sm_100a extraction kernels.
//
// Arithmetic contract (what every kernel in this directory reproduces):
//   map     src/core/geometry.py:141-190 of the reference, evaluated in fp64 per
//           output pixel from the caller's fp64 rotation matrix, rounded to float32
//           (geometry.py:189-190) and quantised to Q5 exactly like cv2.remap does
//           for CV_32FC1 maps (cvRound(m * 32), round-half-even).
//   gather  cv2.remap(..., INTER_LINEAR | INTER_LANCZOS4, BORDER_WRAP) as called at
//           src/core/processor.py:334: Q15 fixed-point weights, (sum + 2^14) >> 15,
//           wrap-around in x AND y.
//   score   src/utils/image_utils.py:22-27: Q15 BGR->gray, 4-neighbour Laplacian with
//           reflect-101 borders, exact integer sum(L) and sum(L^2).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace x360 {

constexpr int kTile = 32;          // output tile is kTile x kTile pixels
constexpr int kThreads = 256;      // 8 warps; warp w owns tile rows 4w .. 4w+3
constexpr int kSlots = 4;          // interior pixels per thread
constexpr double kPi = 3.14159265358979323846;
constexpr double kTwoPi = 2.0 * 3.14159265358979323846;

// Per-job geometry, uniform over a launch.
struct Geom {
    int W, H;            // source frame
    int ow, oh;          // output view
    int tiles_x, tiles_y;
    double f, cx, cy;    // geometry.py:141-142
};

struct ExtractArgs {
    Geom g;
    const uint8_t *frames;      // [F][H][W][3]
    uint8_t *out;               // [F][V][oh][ow][3]
    long long *sum_lap;         // [F][V] or nullptr
    long long *sum_lap2;        // [F][V] or nullptr
    const double *R;            // [V][9] device
    const int16_t *lz_tab;      // [1024][8][8] device (Lanczos4 only)
    int F, V;
    int n_tiles;                // V * tiles_y * tiles_x
    int aligned;                // 1 if frames/out pointers and pitches allow the vector paths
    // staged path (x360_staged.cuh): capacity of the shared-memory footprint buffers
    int stage_ok;               // 1 if frames base / row pitch are 16-B aligned (bulk-copy requirement)
    int rows_max, px_max;       // footprint rows / source pixels per row that fit
    int pp;                     // pair-record row pitch in pixels (px_max + 2: 16-B rows, bank skew)
    int raw_pitch;              // raw row pitch in bytes (multiple of 16)
    unsigned *tile_modes;       // optional [3] counters: staged / direct-aligned / direct-bytewise tiles
};

__device__ __forceinline__ int wrap_idx(int p, int len)
{
    if ((unsigned)p < (unsigned)len) return p;
    p %= len;
    return p < 0 ? p + len : p;
}

__device__ __forceinline__ int sat_s16(int v) { return max(-32768, min(32767, v)); }

// ---- fp64 atan2 for the map ------------------------------------------------------------------
// The map needs atan2 / asin to within a few ulp ("map exactness"); CUDA's libm
// versions cost ~200 instructions each.  This one: one division (argument reduced to
// |z| <= tan(pi/8) by folding (mn - mx)/(mn + mx) into the same quotient), an odd polynomial
// z + z w Q(w), w = z^2 (coefficients from scripts/atan_coeffs.py: Chebyshev fit at 50 digits, max
// relative error 6.5e-17 in double evaluation), quadrant assembly with hi + lo constants.
__constant__ double kAtanQ[13] = {
    -0x1.a71378d0839b7p-7, 0x1.e3ded60e0b4b5p-6, -0x1.4bdeae2516d21p-5, 0x1.816326bd079a1p-5, -0x1.ae847c59c5c84p-5,
    0x1.e1d20955386eap-5, -0x1.111085dfdddd8p-4, 0x1.3b13aa8a180dbp-4, -0x1.745d170db805ep-4, 0x1.c71c71c5eae8cp-4,
    -0x1.249249249055bp-3, 0x1.9999999999964p-3, -0x1.5555555555555p-2};

__device__ __forceinline__ double fast_atan2(double y, double x)
{
    const double ax = fabs(x), ay = fabs(y);
    const double mx = fmax(ax, ay), mn = fmin(ax, ay);
    const bool red = mn > 0x1.a827999fcef32p-2 * mx;                 // tan(pi/8)
    const double num = red ? mn - mx : mn;
    const double den = red ? mn + mx : mx;
    double z = __ddiv_rn(num, den);
    if (mx == 0.0) z = 0.0;                                          // atan2(+-0, +-0)
    const double w = z * z;
    double q = kAtanQ[0];
#pragma unroll
    for (int i = 1; i < 13; ++i) q = fma(q, w, kAtanQ[i]);           // constant-bank operands: no register or uniform moves
    double a = fma(z, q * w, z);                                     // atan(z), |z| <= tan(pi/8)
    if (red) a = 0x1.921fb54442d18p-1 + (a + 0x1.1a62633145c07p-55); // pi/4 + atan(z)
    if (ay > ax) a = 0x1.921fb54442d18p+0 - (a - 0x1.1a62633145c07p-54);   // pi/2 - a
    if (signbit(x)) a = 0x1.921fb54442d18p+1 - (a - 0x1.1a62633145c07p-53); // pi - a
    return copysign(a, y);
}

// geometry.py:149-190 for one output pixel -> fp64 map_x (unrounded) and the float32 map_y.
// The ray v = [xn, yn, 1] @ R.T keeps every rounding of the reference's sequence (no contraction): where a pixel looks straight at a
// pole its components are rounding residues and decide the result.  From the ray on, the only freedom against numpy is the last ulp
// or two of the angles: atan2 / asin (asin(v1 / |v|) is evaluated as atan2(v1, hypot(v0, v2))) and the divisions by 2 pi and pi,
// which are multiplications by the rounded reciprocals here (a division costs ~13 fp64 instructions) — all absorbed by the float32
// cast + Q5 rounding except with probability ~1e-10 per pixel ("map exactness").
__device__ __forceinline__ double map_xn(const Geom &g, int x) { return __ddiv_rn((double)x - g.cx, g.f); }
__device__ __forceinline__ double map_yn(const Geom &g, int y) { return __ddiv_rn((double)y - g.cy, g.f); }
__device__ __forceinline__ void map_ray_n(const double *__restrict__ R, double xn, double yn, double &v0, double &v1, double &v2)
{
    // [xn, yn, 1] @ R.T  (geometry.py:164)
    v0 = __dadd_rn(__fma_rn(yn, R[1], __dmul_rn(xn, R[0])), R[2]);
    v1 = __dadd_rn(__fma_rn(yn, R[4], __dmul_rn(xn, R[3])), R[5]);
    v2 = __dadd_rn(__fma_rn(yn, R[7], __dmul_rn(xn, R[6])), R[8]);
}
__device__ __forceinline__ void map_ray(const double *__restrict__ R, const Geom &g, int x, int y, double &v0, double &v1, double &v2)
{
    map_ray_n(R, map_xn(g, x), map_yn(g, y), v0, v1, v2);
}
constexpr double kInvTwoPi = 0.15915494309189535;       // fl(1 / 2 pi)
constexpr double kInvPi = 0.3183098861837907;           // fl(1 / pi)
// fp64 map_x before the float32 cast (geometry.py:172,183)
__device__ __forceinline__ double map_u_of(const Geom &g, double v0, double v2)
{
    return __dmul_rn(__dadd_rn(__dmul_rn(fast_atan2(v0, v2), kInvTwoPi), 0.5), (double)g.W);
}
// |(v0, v2)|: the horizontal length of the ray (geometry.py:175)
__device__ __forceinline__ double map_rxz(double v0, double v2) { return sqrt(__dadd_rn(__dmul_rn(v0, v0), __dmul_rn(v2, v2))); }
// float32 map_y (geometry.py:175-176,186,190) from v1 and the horizontal length
__device__ __forceinline__ float map_y_from(const Geom &g, double v1, double rxz)
{
    const double phi = fast_atan2(v1, rxz);
    return __double2float_rn(__dmul_rn(__dadd_rn(__dmul_rn(phi, kInvPi), 0.5), (double)g.H));
}
__device__ __forceinline__ float map_y_of(const Geom &g, double v0, double v1, double v2) { return map_y_from(g, v1, map_rxz(v0, v2)); }
__device__ __forceinline__ void map_f64(const double *__restrict__ R, const Geom &g, int x, int y, double &uf, float &my)
{
    double v0, v1, v2;
    map_ray(R, g, x, y, v0, v1, v2);
    uf = map_u_of(g, v0, v2);
    my = map_y_of(g, v0, v1, v2);
}

__device__ __forceinline__ void map_f32(const double *__restrict__ R, const Geom &g, int x, int y,
                                        float &mx, float &my)
{
    double uf;
    map_f64(R, g, x, y, uf, my);
    mx = __double2float_rn(uf);
}

// float32 map value -> Q5 exactly like cv2.remap: cvRound(m * 32), round-half-even
__device__ __forceinline__ int q5_of(float m) { return __float2int_rn(__fmul_rn(m, 32.0f)); }

__device__ __forceinline__ void map_q5(const double *__restrict__ R, const Geom &g, int x, int y,
                                       int &sx, int &sy)
{
    float mx, my;
    map_f32(R, g, x, y, mx, my);
    sx = q5_of(mx);
    sy = q5_of(my);
}

// Q15 BGR->gray of a packed pixel word [B, G, R, x]: (3735 B + 19235 G + 9798 R + 2^14) >> 15
__device__ __forceinline__ uint32_t gray_of(uint32_t bgrx)
{
    uint32_t acc = 16384u;
    asm("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(acc) : "r"(3735u | (19235u << 16)), "r"(bgrx));
    asm("dp2a.hi.u32.u32 %0, %1, %2, %0;" : "+r"(acc) : "r"(9798u), "r"(bgrx));
    return acc >> 15;
}

__device__ __forceinline__ uint32_t dp4a_uu(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// unsigned bytes (a) x signed bytes (b)
__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

__device__ __forceinline__ uint32_t dp2a_lo_uu(uint32_t a16x2, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a16x2), "r"(b), "r"(c));
    return d;
}

// signed 16-bit pair (a) x unsigned byte pair (low / high half of b)
__device__ __forceinline__ int dp2a_lo_su(uint32_t a16x2, uint32_t b, int c)
{
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a16x2), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_hi_su(uint32_t a16x2, uint32_t b, int c)
{
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a16x2), "r"(b), "r"(c));
    return d;
}

__device__ __forceinline__ uint32_t ldg_u32(const uint8_t *p)
{
    return __ldg(reinterpret_cast<const unsigned int *>(p));
}

// Decode (view, tile_y, tile_x) of a linear tile index; tiles are view-major, row-major.
__device__ __forceinline__ void tile_coords(const Geom &g, int t, int &v, int &x0, int &y0)
{
    const int per_view = g.tiles_x * g.tiles_y;
    v = t / per_view;
    const int r = t - v * per_view;
    const int ty = r / g.tiles_x;
    x0 = (r - ty * g.tiles_x) * kTile;
    y0 = ty * kTile;
}

// ---- shared-memory layout of one output tile (+ its gray halo) -------------------
// Gray tile with a 1-px halo: rows -1..32 as compact 32-B rows (row 0 = top halo), the left and
// right halo columns kept apart so interior rows stay word-aligned and conflict-free.
struct GrayTile {
    uint8_t g[kTile + 2][kTile];
    uint8_t hl[kTile];
    uint8_t hr[kTile];
};

struct TileSmem {
    uint8_t out[2][kTile][kTile * 3];        // double-buffered BGR tile, row = 96 B
    GrayTile gt[2];
    long long sums[2][2];                    // per-buffer (sum L, sum L^2)
    int special;                             // tile has a wrapped tap -> generic gather
};

// gray value at tile-local (ly, lx) with ly, lx in [-1, kTile]
__device__ __forceinline__ int gray_at(const GrayTile &t, int ly, int lx)
{
    if (lx < 0) return t.hl[ly];
    if (lx >= kTile) return t.hr[ly];
    return t.g[ly + 1][lx];
}

// Halo pixel of thread `tid` (< 128): 0-31 top, 32-63 bottom, 64-95 left, 96-127 right.
__device__ __forceinline__ void halo_coords(int tid, int x0, int y0, int &hx, int &hy)
{
    const int k = tid & 31;
    switch (tid >> 5) {
        case 0: hx = x0 + k; hy = y0 - 1; break;
        case 1: hx = x0 + k; hy = y0 + kTile; break;
        case 2: hx = x0 - 1; hy = y0 + k; break;
        default: hx = x0 + kTile; hy = y0 + k; break;
    }
}

__device__ __forceinline__ void halo_store(GrayTile &t, int tid, uint8_t g)
{
    const int k = tid & 31;
    switch (tid >> 5) {
        case 0: t.g[0][k] = g; break;
        case 1: t.g[kTile + 1][k] = g; break;
        case 2: t.hl[k] = g; break;
        default: t.hr[k] = g; break;
    }
}

// Store the finished BGR tile from shared memory: 128-bit rows when the geometry allows
// (full tile, 16-B aligned pitch), byte-granular otherwise.
__device__ __forceinline__ void store_tile(const TileSmem &s, int b, uint8_t *__restrict__ out_view,
                                           const Geom &g, int x0, int y0, bool vec_ok, int tid)
{
    const size_t pitch = (size_t)g.ow * 3;
    if (vec_ok) {
        for (int i = tid; i < kTile * 6; i += kThreads) {
            const int row = i / 6, ch = i - row * 6;
            const uint4 v = *reinterpret_cast<const uint4 *>(&s.out[b][row][ch * 16]);
            *reinterpret_cast<uint4 *>(out_view + (size_t)(y0 + row) * pitch + (size_t)x0 * 3 + ch * 16) = v;
        }
    } else {
        const int wbytes = min(kTile, g.ow - x0) * 3;
        const int rows = min(kTile, g.oh - y0);
        for (int i = tid; i < kTile * kTile * 3; i += kThreads) {
            const int row = i / (kTile * 3), bb = i - row * (kTile * 3);
            if (row < rows && bb < wbytes)
                out_view[(size_t)(y0 + row) * pitch + (size_t)x0 * 3 + bb] = s.out[b][row][bb];
        }
    }
}

// Laplacian sums of the tile in buffer b.  Returns this thread's partial (sum L, sum L^2).
// full: the tile is entirely inside the image (fast 4-pixels-per-thread dp4a path).
__device__ __forceinline__ void laplacian_partial(const GrayTile &t, const Geom &g, int x0, int y0,
                                                  bool full, int tid, int &pl, unsigned &pl2)
{
    pl = 0;
    pl2 = 0;
    if (full) {
        const int ly = tid >> 3, wx = tid & 7;           // one 4-pixel word per thread
        const uint32_t C = *reinterpret_cast<const uint32_t *>(&t.g[ly + 1][wx * 4]);
        uint32_t U = *reinterpret_cast<const uint32_t *>(&t.g[ly][wx * 4]);
        uint32_t D = *reinterpret_cast<const uint32_t *>(&t.g[ly + 2][wx * 4]);
        uint32_t l = (wx > 0) ? t.g[ly + 1][wx * 4 - 1] : t.hl[ly];
        uint32_t r = (wx < 7) ? t.g[ly + 1][wx * 4 + 4] : t.hr[ly];
        // reflect-101 at the image border: g[-1] = g[1], g[n] = g[n-2]
        if (y0 + ly == 0) U = D;
        if (y0 + ly == g.oh - 1) D = U;
        if (x0 + wx * 4 == 0) l = (C >> 8) & 0xffu;
        if (x0 + wx * 4 + 3 == g.ow - 1) r = (C >> 16) & 0xffu;
        const uint32_t Lw = __byte_perm(C, l, 0x2104);   // [l, c0, c1, c2]
        const uint32_t Rw = __byte_perm(C, r, 0x4321);   // [c1, c2, c3, r]
        // signed byte weights, little-endian packs
        constexpr uint32_t kA = 0x0001fc01u;             // [ 1, -4,  1,  0]
        constexpr uint32_t kB = 0x01fc0100u;             // [ 0,  1, -4,  1]
        const int L0 = dp4a_us(D, 0x00000001u, dp4a_us(U, 0x00000001u, dp4a_us(Lw, kA, 0)));
        const int L1 = dp4a_us(D, 0x00000100u, dp4a_us(U, 0x00000100u, dp4a_us(Lw, kB, 0)));
        const int L2 = dp4a_us(D, 0x00010000u, dp4a_us(U, 0x00010000u, dp4a_us(Rw, kA, 0)));
        const int L3 = dp4a_us(D, 0x01000000u, dp4a_us(U, 0x01000000u, dp4a_us(Rw, kB, 0)));
        pl = L0 + L1 + L2 + L3;
        pl2 = (unsigned)(L0 * L0) + (unsigned)(L1 * L1) + (unsigned)(L2 * L2) + (unsigned)(L3 * L3);
    } else {
        for (int i = tid; i < kTile * kTile; i += kThreads) {
            const int ly = i >> 5, lx = i & 31;
            const int x = x0 + lx, y = y0 + ly;
            if (x >= g.ow || y >= g.oh) continue;
            // reflect-101 in image coordinates, then back to tile-local
            const int yu = (y == 0) ? ((g.oh > 1) ? 1 : 0) : y - 1;
            const int yd = (y == g.oh - 1) ? ((g.oh > 1) ? g.oh - 2 : 0) : y + 1;
            const int xl = (x == 0) ? ((g.ow > 1) ? 1 : 0) : x - 1;
            const int xr = (x == g.ow - 1) ? ((g.ow > 1) ? g.ow - 2 : 0) : x + 1;
            const int c = gray_at(t, ly, lx);
            const int L = gray_at(t, yu - y0, lx) + gray_at(t, yd - y0, lx) +
                          gray_at(t, ly, xl - x0) + gray_at(t, ly, xr - x0) - 4 * c;
            pl += L;
            pl2 += (unsigned)(L * L);
        }
    }
}

}  // namespace x360

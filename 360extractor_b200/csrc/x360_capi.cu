// x360_capi.cu — host side of libx360: the C-ABI declared in include/x360.h.
//
// One translation unit: kernels (x360_*.cuh) + plan management + the pipelined host path.
// Built by 360extractor_b200/build.py with
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
// There is no CPU fallback anywhere in this file: every compute entry needs a CUDA device.
#include "../../include/x360.h"

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <algorithm>
#include <string>
#include <vector>

#include "x360_common.cuh"
#include "x360_extract.cuh"
#include "x360_jpeg.cuh"
#include "x360_jpeg_tables.hpp"
#include "x360_lanczos.cuh"
#include "x360_linear.cuh"
#include "x360_misc.cuh"
#include "x360_staged.cuh"
#include "x360_tables.hpp"
#include "x360_tiled.cuh"
#include "x360_unsharp.cuh"

using namespace x360;

namespace {

constexpr int kRingSlots = 64, kRingWords = 16;

thread_local std::string g_err;
std::atomic<long long> g_launches{0};

int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                 \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess)                                                                   \
            return fail(X360_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// Tuning switches for experiments exist only in libraries built with -DX360_LAB; the default build reads no
// environment variable (plan variants are chosen by the plan, or by x360_params.flags).
#ifdef X360_LAB
inline const char *lab_env(const char *name) { return getenv(name); }
#else
inline const char *lab_env(const char *) { return nullptr; }
#endif
inline int lab_int(const char *name, int dflt) { const char *e = lab_env(name); return e ? atoi(e) : dflt; }

// Dynamic shared memory the tiled kernels may be given (choose_capacity never exceeds it).  The limit is a property of
// the kernel FUNCTION, shared by every plan of the process, so it is always raised to this one cap — never to a plan's
// own size, which would lower it under another live plan.
constexpr int kTiledSmemCap = 110 * 1024;

// Device scratch of the one-shot entries (x360_remap_maps, x360_blur_sums, x360_make_maps: INTEGRATION.md option A, called once
// per view by a host that keeps the reference's loop): grow-only buffers kept per thread and device instead of a cudaMalloc /
// cudaFree pair per call (each costs a device synchronisation).
struct ScratchSlot { void *p = nullptr; size_t cap = 0; int device = -1; };
void *scratch_get(int slot, int device, size_t bytes)
{
    thread_local ScratchSlot slots[6];
    ScratchSlot &s = slots[slot];
    if (s.p && (s.device != device || s.cap < bytes)) {
        cudaSetDevice(s.device);
        cudaFree(s.p);
        cudaSetDevice(device);
        s.p = nullptr;
        s.cap = 0;
    }
    if (!s.p) {
        const size_t want = bytes + bytes / 4 + 4096;
        if (cudaMalloc(&s.p, want) != cudaSuccess) { cudaGetLastError(); s.p = nullptr; return nullptr; }
        s.cap = want;
        s.device = device;
    }
    return s.p;
}

// geometry.py:141 — f = (0.5*dest_w) / tan(0.5*radians(fov_deg)); radians(x) = x*(pi/180)
double focal_of(int out_w, double fov_deg)
{
    const double rad = fov_deg * (3.14159265358979323846 / 180.0);
    return (0.5 * (double)out_w) / std::tan(0.5 * rad);
}

Geom make_geom(int W, int H, int ow, int oh, double fov_deg)
{
    Geom g;
    g.W = W; g.H = H; g.ow = ow; g.oh = oh;
    g.tiles_x = (ow + kTile - 1) / kTile;
    g.tiles_y = (oh + kTile - 1) / kTile;
    g.f = focal_of(ow, fov_deg);
    g.cx = ow / 2.0;
    g.cy = oh / 2.0;
    return g;
}

int check_dims(int W, int H, int ow, int oh)
{
    if (W < 2 || H < 2 || W > 32767 || H > 32767)
        return fail(X360_E_INVALID, "source size %dx%d outside [2, 32767] (cv2.remap's int16 coordinate range)", W, H);
    if (ow < 1 || oh < 1 || ow > 32768 || oh > 32768) return fail(X360_E_INVALID, "bad output size %dx%d", ow, oh);
    if ((size_t)W * H * 3 > 0xffffff00ull) return fail(X360_E_INVALID, "frame larger than 4 GiB");
    return X360_OK;
}

#ifdef X360_LAB     // the round-1 comparison kernels (x360_params.path 1 / 2)
typedef void (*extract_fn)(const ExtractArgs);
typedef void (*staged_fn)(const ExtractArgs, const TmapSet);

// The direct extraction kernel for (interpolation, scores); Lanczos4 has the direct path only.
#ifndef X360_LZ_MINB
#define X360_LZ_MINB 2
#endif
extract_fn pick_direct(int interp, bool scores)
{
    if (interp == X360_INTERP_LANCZOS4)
        return scores ? extract_direct<LanczosSampler, true, X360_LZ_MINB> : extract_direct<LanczosSampler, false, X360_LZ_MINB>;
    return scores ? extract_direct<LinearSampler, true, 3> : extract_direct<LinearSampler, false, 4>;
}

// MINB (CTAs per SM the register allocation aims for) is a tuning knob: X360_STAGED_MINB=3 trades
// occupancy for instruction-level parallelism inside the gather.
staged_fn pick_staged(bool scores)
{
    static const int minb = lab_int("X360_STAGED_MINB", 4);
    if (minb == 3) return scores ? extract_linear_staged<true, 3> : extract_linear_staged<false, 3>;
    return scores ? extract_linear_staged<true, 4> : extract_linear_staged<false, 4>;
}
#endif

// cuTensorMapEncodeTiled through the runtime's driver entry point (libcuda is not linked).
typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

encode_tiled_fn tensor_map_encoder()
{
    static encode_tiled_fn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (encode_tiled_fn)p;
    }
    return fn;
}

#ifdef X360_LAB
// uint8 [F][H][3W] frame tensor, boxes {128 + 32 k bytes, kBoxRows rows, 1 frame}
bool encode_frame_maps(const uint8_t *frames, int F, int H, int W, TmapSet *out)
{
    encode_tiled_fn enc = tensor_map_encoder();
    if (!enc) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)W * 3, (cuuint64_t)H, (cuuint64_t)F};
    const cuuint64_t strides[2] = {(cuuint64_t)W * 3, (cuuint64_t)W * 3 * H};
    const cuuint32_t estr[3] = {1, 1, 1};
    for (int k = 0; k < kNumBoxWidths; ++k) {
        const cuuint32_t box[3] = {(cuuint32_t)(128 + 32 * k), (cuuint32_t)kBoxRows, 1};
        if (enc(&out->m[k], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void *)frames, dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
    }
    return true;
}
#endif

int round_up(int v, int m) { return (v + m - 1) / m * m; }

// ---- tiled path (x360_tiled.cuh) ---------------------------------------------------------------
typedef void (*tiled_fn)(const TiledArgs, const TiledMaps);

// Tile variants: 32x32 = 4 warps x 8 rows per thread; 32x16 = 4 warps x 4 rows.  (2 warps x 8 rows, 64-thread
// CTAs, was measured too: equal on cfg2, 20-25 % slower where footprints are large — half the warps per SM.)
int tiled_warps(bool, int) { return 4; }

// What a plan may be told instead of choosing by itself: x360_params.flags (include/x360.h), plus — in -DX360_LAB
// builds only — the environment switches the round-1 experiments used.
struct PlanOpts {
    bool planes = false;        // bilinear on the pair-record planes instead of the raw gather
    int score_mode = 0;         // 0 = the plan's default, 1 = fused into the extraction kernel, 2 = split pass over the views
    int force_th = 0;           // 16 / 32: output tile height
    bool no_wide = false;       // no wide / seam boxes
    bool full_upload = false;   // host entries upload whole frames
};
PlanOpts plan_opts(int flags)
{
    PlanOpts o;
    o.planes = (flags & X360_FLAG_PLANES) != 0 || lab_int("X360_NO_RAW", 0) != 0;
    o.score_mode = (flags & X360_FLAG_SPLIT_SCORE) ? 2 : ((flags & X360_FLAG_FUSED_SCORE) ? 1 : 0);
    if (lab_int("X360_SPLIT_SCORE", 0)) o.score_mode = 2;
    if (lab_int("X360_RAW_SCORES", 0)) o.score_mode = 1;
    o.force_th = (flags & X360_FLAG_TILE_H16) ? 16 : ((flags & X360_FLAG_TILE_H32) ? 32 : lab_int("X360_TILED_TH", 0));
    o.no_wide = (flags & X360_FLAG_NO_WIDE) != 0 || lab_int("X360_NO_WIDE", 0) != 0;
    o.full_upload = (flags & X360_FLAG_FULL_UPLOAD) != 0;
    return o;
}
// Score policy.  By default the blur score is a second, streaming pass over the finished views (blur_sums_march_kernel: warp
// shuffles for the neighbours, block reduction, one atomic pair per CTA): measured on cfg4 (B200) 2.40 ms per 32 frames against
// 2.87 ms with the score fused into the extraction kernel — the 1-px halo of every 32 x 16 tile costs more than re-reading
// 3 B / pixel.  X360_FLAG_FUSED_SCORE selects the fused form (on the raw gather; with X360_FLAG_PLANES on the planes).
bool score_is_split(const PlanOpts &o) { return o.score_mode != 1; }
// The raw gather (no transform, x360_tiled.cuh RAW) serves every bilinear launch except the ones forced onto the
// planes (Lanczos4 has its own record layout).
bool tiled_raw(bool, bool lanczos, const PlanOpts &o) { return !lanczos && !o.planes; }

#ifndef X360_RAW_MINB
#define X360_RAW_MINB 8       // register allocation target of the 32x32 column-only raw gather: 64 registers, no spills (7 CTAs/SM fit the shared memory; measured +1 % over the 72-register build)
#endif
tiled_fn pick_tiled(bool scores, bool lanczos, int th, bool col, bool raw)
{
    if (raw) {                          // bilinear: selectors per thread (column-only map_x) or per pixel
        if (scores) {
            if (th == 16) return col ? extract_tiled<true, false, 4, 4, 6, true, true> : extract_tiled<true, false, 4, 4, 6, true, false>;
            return col ? extract_tiled<true, false, 8, 4, 5, true, true> : extract_tiled<true, false, 8, 4, 5, true, false>;
        }
        if (th == 16) return col ? extract_tiled<false, false, 4, 4, 7, true, true> : extract_tiled<false, false, 4, 4, 7, true, false>;
        return col ? extract_tiled<false, false, 8, 4, X360_RAW_MINB, true, true> : extract_tiled<false, false, 8, 4, X360_MINB8, true, false>;
    }
    if (col && !scores && !lanczos && th == 32) return extract_tiled<false, false, 8, 4, 7>;
    if (th == 16) {
        if (lanczos) return scores ? extract_tiled<true, true, 4, 4> : extract_tiled<false, true, 4, 4>;
        return scores ? extract_tiled<true, false, 4, 4> : extract_tiled<false, false, 4, 4>;
    }
    if (lanczos) return scores ? extract_tiled<true, true, 8, 4> : extract_tiled<false, true, 8, 4>;
    return scores ? extract_tiled<true, false, 8, 4> : extract_tiled<false, false, 8, 4>;
}

// load boxes {48 (w + 1) bytes, rows_max - 8 (2 - h) rows, 1 frame} of uint8 [F][H][3W] (raw gather: 64 (w + 1));
// wide boxes {tiled_wide_bw(c) bytes, wide_rows[c]} through a uint32 view; seam boxes over the window tensor (x360_tiled.cuh);
// store box {96, tile height, 1} of uint8 [F*V][oh][3 ow].  *wide_ok / *seam_ok report what could be encoded.
bool encode_tiled_maps(const uint8_t *frames, int F, int H, int W, int rows_max, bool with_in, uint8_t *out, int FV, int oh,
                       int ow, int tile_h, bool with_out, TiledMaps *tm, bool raw, bool col_only, const int *wide_rows, bool *wide_ok, bool *seam_ok)
{
    encode_tiled_fn enc = tensor_map_encoder();
    if (!enc) return false;
    const cuuint32_t estr[3] = {1, 1, 1};
    auto put = [&](CUtensorMap *m, CUtensorMapDataType dt, void *base, const cuuint64_t *dims, const cuuint64_t *strides, cuuint32_t bx,
                   cuuint32_t by, CUtensorMapL2promotion l2) {
        const cuuint32_t box[3] = {bx, by, 1};
        return enc(m, dt, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, l2,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    };
    if (with_in) {
        const cuuint64_t dims[3] = {(cuuint64_t)W * 3, (cuuint64_t)H, (cuuint64_t)F};
        const cuuint64_t strides[2] = {(cuuint64_t)W * 3, (cuuint64_t)W * 3 * H};
        for (int k = 0; k < kTNumBox; ++k) {
            const int rows = rows_max - kTBoxRows * (kTNumBoxH - 1 - k / kTNumBoxW);
            const int wk = k % kTNumBoxW;
            if (!put(&tm->m[k], CU_TENSOR_MAP_DATA_TYPE_UINT8, (void *)frames, dims, strides, (cuuint32_t)tiled_box_bw(raw, col_only, raw ? std::min(wk, 3) : wk),
                     (cuuint32_t)(rows < 8 ? 8 : rows), CU_TENSOR_MAP_L2_PROMOTION_L2_128B))
                return false;
        }
        // wide classes: uint32 elements (a box dimension holds at most 256 elements)
        if (*wide_ok) {
            const cuuint64_t d4[3] = {(cuuint64_t)W * 3 / 4, (cuuint64_t)H, (cuuint64_t)F};
            for (int c = 0; c < kTNumWide && *wide_ok; ++c)
                if (wide_rows[c] >= 1 &&
                    !put(&tm->m[kTMapWide0 + c], CU_TENSOR_MAP_DATA_TYPE_UINT32, (void *)frames, d4, strides, (cuuint32_t)(tiled_wide_bw(c) / 4),
                         (cuuint32_t)wide_rows[c], CU_TENSOR_MAP_L2_PROMOTION_L2_128B))
                    *wide_ok = false;
        }
        // seam window: row r = the last kTSeamHalf bytes of global row r and the first kTSeamHalf bytes of row r + 1
        if (*seam_ok) {
            const cuuint64_t n_rows = (cuuint64_t)F * H - 1;
            const cuuint64_t d1[3] = {(cuuint64_t)2 * kTSeamHalf, n_rows, 1}, d4[3] = {(cuuint64_t)2 * kTSeamHalf / 4, n_rows, 1};
            const cuuint64_t st[2] = {(cuuint64_t)W * 3, (cuuint64_t)W * 3 * n_rows};
            void *base = (void *)(frames + (size_t)W * 3 - kTSeamHalf);
            for (int w = 0; w < 4 && *seam_ok; ++w)
                if (!put(&tm->m[kTMapSeam0 + w], CU_TENSOR_MAP_DATA_TYPE_UINT8, base, d1, st, (cuuint32_t)tiled_box_bw(true, col_only, w), (cuuint32_t)(rows_max < 8 ? 8 : rows_max),
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B))
                    *seam_ok = false;
            for (int c = 0; c < kTNumWide && *seam_ok && *wide_ok; ++c)
                if (wide_rows[c] >= 1 &&
                    !put(&tm->m[kTMapSeam0 + 4 + c], CU_TENSOR_MAP_DATA_TYPE_UINT32, base, d4, st, (cuuint32_t)(tiled_wide_bw(c) / 4), (cuuint32_t)wide_rows[c],
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B))
                    *seam_ok = false;
        }
    } else {
        *wide_ok = false;
        *seam_ok = false;
    }
    if (with_out) {
        const cuuint64_t dims[3] = {(cuuint64_t)ow * 3, (cuuint64_t)oh, (cuuint64_t)FV};
        const cuuint64_t strides[2] = {(cuuint64_t)ow * 3, (cuuint64_t)ow * 3 * oh};
        if (!put(&tm->m[kTMapOut], CU_TENSOR_MAP_DATA_TYPE_UINT8, (void *)out, dims, strides, (cuuint32_t)(kTile * 3), (cuuint32_t)tile_h, CU_TENSOR_MAP_L2_PROMOTION_NONE))
            return false;
        // transposed tiles of the 16-row variant: tile_h pixels wide, kTile rows
        if (tile_h == 16 && oh >= kTile &&
            !put(&tm->m[kTMapOutT], CU_TENSOR_MAP_DATA_TYPE_UINT8, (void *)out, dims, strides, (cuuint32_t)(tile_h * 3), (cuuint32_t)kTile, CU_TENSOR_MAP_L2_PROMOTION_NONE))
            return false;
    }
    return true;
}

// box heights of the wide classes for a raw buffer of rows_max x bw_max bytes
void wide_class_rows(int rows_max, int bw_max, int *wide_rows)
{
    const int raw = rows_max * bw_max;                            // (not rounded up: the pair-record area is sized from this product)
    for (int c = 0; c < kTNumWide; ++c) {
        int r = raw / tiled_wide_bw(c);
        wide_rows[c] = r > 256 ? 256 : (r >= 6 ? r : 0);          // a class with fewer than 6 rows serves nothing
    }
}

// Rb = Ry(delta) Ra ?  (the reference's Ry, geometry.py:105-109) -> delta in radians.
// Ring views and the four horizontal cube faces differ only by such a yaw: map_y is identical and
// map_x is shifted by delta / 2pi * W.
bool yaw_delta(const double *Ra, const double *Rb, double *delta)
{
    double M[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            M[i][j] = Rb[3 * i] * Ra[3 * j] + Rb[3 * i + 1] * Ra[3 * j + 1] + Rb[3 * i + 2] * Ra[3 * j + 2];
    const double tol = 1e-12;
    if (std::fabs(M[1][1] - 1.0) > tol || std::fabs(M[0][1]) > tol || std::fabs(M[1][0]) > tol ||
        std::fabs(M[1][2]) > tol || std::fabs(M[2][1]) > tol || std::fabs(M[0][0] - M[2][2]) > tol ||
        std::fabs(M[0][2] + M[2][0]) > tol)
        return false;
    *delta = std::atan2(M[0][2], M[0][0]);
    return true;
}

struct HostUnits {
    std::vector<TiledUnit> units;      // largest first
    std::vector<int> view;
    std::vector<double> shift;
};

// Does an output pixel of this view look (within 1e-3 px) straight at a pole?  There v0 and v2 are rounding residues
// (~1e-16) and map_x = atan2(v0, v2) is decided by each view's own residues (geometry.py:158-172), so the identity
// "map_x of a yaw-shifted view = map_x of the first view + shift" does not hold bit for bit: such views keep a unit of
// their own and evaluate their own map.  (The pole direction in camera coordinates is +-(R10, R11, R12).)
bool has_pole_pixel(const double *Rv, const Geom &g)
{
    if (std::fabs(Rv[5]) < 1e-12) return false;                           // poles on the horizon of the image plane
    const double px = g.cx + g.f * (Rv[3] / Rv[5]), py = g.cy + g.f * (Rv[4] / Rv[5]);
    const double rx = std::nearbyint(px), ry = std::nearbyint(py);
    return std::fabs(px - rx) < 1e-3 && std::fabs(py - ry) < 1e-3 && rx >= 0.0 && rx < (double)g.ow && ry >= 0.0 && ry < (double)g.oh;
}

HostUnits build_units(const double *R, int V, int W, const Geom &g, bool share, int max_views = kTMaxUnitViews)
{
    std::vector<std::vector<int>> members;
    std::vector<std::vector<double>> shifts;
    for (int v = 0; v < V; ++v) {
        bool placed = false;
        const bool alone = has_pole_pixel(R + 9 * v, g);
        for (size_t u = 0; share && !alone && u < members.size() && !placed; ++u) {
            if (has_pole_pixel(R + 9 * members[u][0], g)) continue;
            double d;
            if ((int)members[u].size() < max_views && yaw_delta(R + 9 * members[u][0], R + 9 * v, &d)) {
                double sh = d / (2.0 * 3.14159265358979323846) * (double)W;
                if (sh < 0.0) sh += (double)W;
                if (sh >= (double)W) sh = 0.0;
                members[u].push_back(v);
                shifts[u].push_back(sh);
                placed = true;
            }
        }
        if (!placed) {
            members.push_back({v});
            shifts.push_back({0.0});
        }
    }
    std::vector<size_t> order(members.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](size_t x, size_t y) { return members[x].size() > members[y].size(); });
    HostUnits h;
    for (size_t o : order) {
        TiledUnit t;
        t.first = (int)h.view.size();
        t.n = (int)members[o].size();
        h.units.push_back(t);
        h.view.insert(h.view.end(), members[o].begin(), members[o].end());
        h.shift.insert(h.shift.end(), shifts[o].begin(), shifts[o].end());
    }
    return h;
}

// 1 / D when every multi-view unit's shifts are 0, D, 2D, ... (source pixels) with one D for all of them and n D <= W; else 0.
double even_yaw_spacing(const HostUnits &h, int W)
{
    double D = 0.0;
    for (const TiledUnit &t : h.units) {
        if (t.n < 2) continue;
        const double d = h.shift[(size_t)t.first + 1];
        if (d <= 0.0 || (D > 0.0 && std::fabs(d - D) > 1e-6 * W)) return 0.0;
        D = d;
        if (t.n * d > (double)W * (1.0 + 1e-9)) return 0.0;
        for (int i = 0; i < t.n; ++i)
            if (std::fabs(h.shift[(size_t)t.first + i] - i * d) > 1e-6 * W) return 0.0;
    }
    return D > 0.0 ? 1.0 / D : 0.0;
}

// Host estimate of the per-tile source footprints (9 sample points of each tile, + the 1-px halo the fused
// score gathers when `scores`) -> the tile
// height (32 or 16 output rows) and staging capacity (footprint rows, raw bytes per row) that maximise
//   (throughput factor of the CTAs/SM the shared-memory size allows) x (factor of the tile height)
//   / (cost of the tiles that fit + 3 x the tiles that do not: those go through the direct gathers).
// Seam-wrapping and pole tiles never fit whatever the capacity.
// map_x independent of the output row for every view?  v0 = xn R00 + yn R01 + R02 and v2 = xn R20 + yn R21 + R22 lose
// yn exactly when R01 = R21 = 0 (yaw-only rotations: ring layouts at pitch 0, the horizontal cube faces).  The one
// case where the SIGN of a zero product could still matter (v0 = +-0 with v2 < 0: theta = +-pi) is excluded.
bool column_only_map_x(const double *R, int V)
{
    for (int v = 0; v < V; ++v) {
        const double *r = R + 9 * v;
        if (r[1] != 0.0 || r[7] != 0.0) return false;
        if (r[2] == 0.0 && r[8] < 0.0) return false;
    }
    return V > 0;
}

// Transposed tiles (x360_tiled.cuh, TiledArgs::tr_tab): which plans may have them, and which 32 x 32 output blocks take them.
bool transposable_plan(const Geom &g, bool col_only, bool raw, bool scores, bool lanczos)
{
    return raw && !col_only && !scores && !lanczos && g.ow == g.oh && g.ow % kTile == 0 && !lab_int("X360_NO_TR", 0);
}
// the source ROW moves faster along output x than along output y at the block's centre: lanes along x would read a steep line
bool block_transposed(const Geom &g, const double *Rv, int bx, int by)
{
    auto src_row = [&](double x, double y) {
        const double xn = (x - g.cx) / g.f, yn = (y - g.cy) / g.f;
        const double v0 = xn * Rv[0] + yn * Rv[1] + Rv[2], v1 = xn * Rv[3] + yn * Rv[4] + Rv[5], v2 = xn * Rv[6] + yn * Rv[7] + Rv[8];
        return (std::asin(v1 / std::sqrt(v0 * v0 + v1 * v1 + v2 * v2)) / 3.14159265358979323846 + 0.5) * g.H;
    };
    const double xc = bx * kTile + 0.5 * (kTile - 1), yc = by * kTile + 0.5 * (kTile - 1), d = 8.0;
    return std::fabs(src_row(xc + d, yc) - src_row(xc - d, yc)) > std::fabs(src_row(xc, yc + d) - src_row(xc, yc - d));
}

void choose_capacity(const Geom &g, const double *R, int V, bool lanczos, bool scores, const PlanOpts &opts, int *rows_out, int *bw_out, int *tile_h_out,
                     long long *unfit_out = nullptr)
{
    long long best_unfit = 0;
    const bool col = column_only_map_x(R, V);
    // measured on cfg2 (B200): 4 CTAs/SM 0.92, 5 CTAs 1.0, 6 CTAs 1.10; fewer CTAs hide less latency
    static const double occ_factor[9] = {0.0, 0.35, 0.6, 0.8, 0.92, 1.0, 1.10, 1.14, 1.17};
    double best = -1.0;
    int best_rows = 40, best_bw = 144, best_th = 32;
    const int m = lanczos ? 7 : 0;                                        // 8x8 taps: 3 before, 4 after
    const bool raw = tiled_raw(scores, lanczos, opts);
    const bool can_tr = transposable_plan(g, col, raw, scores, lanczos);
    struct Foot { int rows, need; bool seam; };
    for (int th = 32; th >= 16; th -= 16) {
        if (opts.force_th && opts.force_th != th) continue;
        std::vector<Foot> feet;
        size_t total = 0;
        const int tiles_y = (g.oh + th - 1) / th;
        for (int v = 0; v < V; ++v) {
            const double *Rv = R + 9 * v;
            for (int ty = 0; ty < tiles_y; ++ty)
                for (int tx = 0; tx < g.tiles_x; ++tx) {
                    ++total;
                    // the tile's rectangle of the view: 32 x th at (32 tx, th ty), or — a transposed block — th x 32
                    int rx = tx * kTile, ry = ty * th, rw = kTile, rh = th;
                    if (th == 16 && can_tr && block_transposed(g, Rv, tx, ty >> 1)) { rx = tx * kTile + (ty & 1) * th; ry = (ty >> 1) * kTile; rw = th; rh = kTile; }
                    int xs[9], ymin = INT_MAX, ymax = INT_MIN;
                    for (int j = 0; j < 3; ++j)
                        for (int i = 0; i < 3; ++i) {
                            const int hl = scores ? 1 : 0;                        // halo
                            const double x = std::min<double>(std::max(0, rx - hl + i * (rw - 1 + 2 * hl) / 2), g.ow - 1);
                            const double y = std::min<double>(std::max(0, ry - hl + j * (rh - 1 + 2 * hl) / 2), g.oh - 1);
                            const double xn = (x - g.cx) / g.f, yn = (y - g.cy) / g.f;
                            const double v0 = xn * Rv[0] + yn * Rv[1] + Rv[2], v1 = xn * Rv[3] + yn * Rv[4] + Rv[5],
                                         v2 = xn * Rv[6] + yn * Rv[7] + Rv[8];
                            const double u = (std::atan2(v0, v2) / (2.0 * 3.14159265358979323846) + 0.5) * g.W;
                            const double w = (std::asin(v1 / std::sqrt(v0 * v0 + v1 * v1 + v2 * v2)) / 3.14159265358979323846 + 0.5) * g.H;
                            xs[3 * j + i] = (int)std::floor(u);
                            const int iy = (int)std::floor(w);
                            ymin = std::min(ymin, iy); ymax = std::max(ymax, iy);
                        }
                    int xmin = *std::min_element(xs, xs + 9), xmax = *std::max_element(xs, xs + 9);
                    bool seam = false;
                    if (xmax - xmin > g.W / 2 || xmax + 1 + (lanczos ? 3 : 0) > g.W - 1) {   // across the seam: unwrapped columns (raw gather only)
                        if (!raw || opts.no_wide) continue;
                        for (int &x : xs) x += (x < g.W / 2) ? g.W : 0;
                        xmin = *std::min_element(xs, xs + 9); xmax = *std::max_element(xs, xs + 9);
                        seam = true;
                    }
                    if (ymin - (lanczos ? 3 : 0) < (seam ? 1 : 0) || ymax + 1 + (lanczos ? 4 : 0) > g.H - 1 - (seam ? 1 : 0)) continue;   // touches a pole row
                    Foot ft;
                    ft.rows = ymax - ymin + 3 + m + (seam ? 1 : 0);
                    const int npx = xmax - (xmin & ~3) + 2 + m;
                    ft.need = raw ? 15 + 3 * (xmax - xmin + 3) + (seam ? 8 : 0) : 12 + 3 * (npx + 1);   // worst alignment
                    ft.seam = seam;
                    feet.push_back(ft);
                }
        }
        // 16-row tiles pay the per-frame overheads over half the pixels and reuse fewer records (cfg2: 0.89 at equal CTAs/SM)
        const double th_factor = th == 32 ? 1.0 : 0.93;
        for (int wc = 1; wc < (raw ? 4 : kTNumBoxW); ++wc)
            for (int r = 24; r <= 96; r += kTBoxRows) {
                const int bw = tiled_box_bw(raw, col, wc);
                const size_t smem = tiled_smem_bytes(r, bw, scores, lanczos, th, col, raw) + 1024;
                if (smem > (size_t)kTiledSmemCap) continue;
                int occ = (int)((227 * 1024) / smem);
                occ = occ > 8 ? 8 : occ;
                int wr[kTNumWide];
                wide_class_rows(r, bw, wr);
                size_t fit = 0, fit_wide = 0;
                for (const Foot &ft : feet) {
                    if (ft.rows <= r && tiled_box_bw(raw, col, std::max(0, tiled_box_class(raw, col, ft.need))) <= bw) { ++fit; continue; }
                    if (opts.no_wide) continue;
                    for (int c = 0; c < kTNumWide; ++c)
                        if (ft.need <= tiled_wide_bw(c) && ft.rows <= wr[c]) { ++fit_wide; break; }
                }
                // cost per tile: 1 in a regular box, 3 otherwise.  Wide boxes and the second launch take most of the rest off the
                // direct path, but measured (cfg1, cfg4) they do not pay for a larger capacity: a cube plan runs 11 % faster
                // with 32-row buffers than with 40-row ones at the same 7 CTAs/SM, so the capacity is chosen for the regular tiles.
                // A group of general views (pole faces) is another matter: it is latency-bound whatever the capacity (measured:
                // the same 0.36 ms for cfg1's pole faces from 27 KB to 58 KB of shared memory), and a wide box is far cheaper
                // than gathering the tile from global memory, so there the tiles a wide class takes count as covered too.
                const double n = (double)(total ? total : 1);
                const double covered = (double)fit + (col ? 0.0 : (double)fit_wide);
                const double cost = (covered + 3.0 * ((double)total - covered)) / n;
                const double score = th_factor * occ_factor[occ] / cost;
                if (score > best * 1.0001) { best = score; best_rows = r; best_bw = bw; best_th = th; best_unfit = (long long)(total - fit - fit_wide); }
            }
    }
    *rows_out = best_rows;
    *bw_out = best_bw;
    *tile_h_out = best_th;
    if (unfit_out) *unfit_out = best_unfit;
}

}  // namespace

struct x360_jpeg;

// tensor maps of the last (frames, batch, out) buffer triple per variant: re-encoded only when a pointer changes
struct MapCache {
    const void *frames = nullptr; const void *out = nullptr; int n_frames = 0;
    int stage_ok = 0, store_ok = 0; bool wide_ok = false, seam_ok = false;
    TiledMaps tms;
};

// A plan's views are served by one or two GROUPS, each with its own kernel variant, tile height, staging capacity and
// launch: the views whose map_x does not depend on the output row (yaw-only rotations: every view of a ring at pitch 0, the
// four horizontal faces of a cube) take the column-only raw gather (32 x 32 tiles, per-thread selectors, 7 CTAs/SM); the
// others (pole faces, inclined or fibonacci views) the general one.  A cube used to run all six faces through the general
// variant because two of them need it.
struct TiledGroup {
    int n_views = 0;
    bool col_u0 = false;                 // map_x depends on the output column only (yaw-only rotations)
    int t_bw[2] = {144, 144}, t_rows[2] = {40, 40}, t_th[2] = {32, 32};   // staging capacity, tile height (32 or 16), without / with fused scores
    bool t_raw[2] = {false, false};      // the launch takes the raw-gather variant (no transform)
    size_t t_smem[2] = {0, 0};           // dynamic shared memory without / with fused scores
    int grid_cap[2] = {0, 0};            // SMs x resident CTAs (occupancy query at creation)
    int wide_rows[2][kTNumWide] = {};
    MapCache maps[2], maps2[2];
    // second launch (x360_tiled.cuh, TiledArgs::mode): larger staging buffers for the (tile, view) pairs the first launch defers
    bool two_launch[2] = {false, false};
    int t2_rows[2] = {0, 0}, t2_bw[2] = {0, 0}, wide_rows2[2][kTNumWide] = {}, grid_cap2[2] = {0, 0};
    size_t t2_smem[2] = {0, 0};
    // view-unit tables at four granularities (at most 8, 4, 2, 1 views per unit, GLOBAL view indices): a launch takes the
    // coarsest one that still yields enough work items without cutting the frame batch into chunks
    int n_units[4] = {0, 0, 0, 0};
    TiledUnit *d_units[4] = {};
    int *d_unit_view[4] = {};
    double *d_unit_shift[4] = {};
    std::vector<int> h_ids;              // the group's plan view indices, ascending (= the order of its one-view units, level 3)
    uint8_t *d_tr_tab = nullptr;         // TiledArgs::tr_tab: [plan views][32 x 32 blocks], or nullptr
    long long n_tr_blocks = 0;
    double rot_inv_pitch[4] = {0, 0, 0, 0};   // TiledArgs::rot_inv_pitch of each level (0: views not evenly spaced, or a general group)
};

struct x360_plan {
    x360_params p;
    Geom g;
    int sm_count = 0;
    double *d_R = nullptr;
    int16_t *d_lz = nullptr;
    float *d_lz_k1 = nullptr;            // Lanczos4: the table's factors (x360_tables.hpp lanczos4_factors)
    uint32_t *d_lz_fix = nullptr;
    int path = X360_PATH_DIRECT;        // resolved path of the extraction kernel
    int px_max = 0, rows_max = 0, pp = 0, raw_pitch = 0;
    size_t smem = 0;                     // dynamic shared memory of the staged kernel
    // Per-launch device scratch: a ring of kRingSlots x kRingWords words — [0..4] tile-mode counters, [5..8] the work-item
    // counters of up to four kernel launches of one extract call (two groups x two launches) — so calls in flight on different
    // streams never share a counter.
    unsigned *d_modes = nullptr;
    int ring_next = 0, ring_last = 0;
    PlanOpts opts;
    TiledGroup grp[2];
    int n_groups = 0;
    std::vector<int> view_group;         // group of each view
    static constexpr int kDeferBufs = 4;
    unsigned *d_defer[kDeferBufs] = {};  // [2 defer_cap] entries + [1] count each
    cudaEvent_t defer_ev[kDeferBufs] = {};
    bool defer_used[kDeferBufs] = {};
    int defer_cap = 0, defer_next = 0;
    bool split_score = false;            // scores by a separate pass over the finished views instead of the fused reduction
    cudaStream_t s_side = nullptr;       // the second group's launches run beside the first group's (forked from / joined to the caller's stream)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // pipelined host path
    int chunk = 0;
    cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[2] = {}, ev_k[2] = {}, ev_out[2] = {};
    // Single-frame chunks of plans whose views share no map (fibonacci, inclined layouts): the views are extracted in view_parts
    // launches over consecutive view ranges, and each range is downloaded while the next one is computed
#ifndef X360_VIEW_PARTS
#define X360_VIEW_PARTS 8             // cfg3 (20 views, one frame) end to end: 2 ranges 6.97 ms, 4: 6.40, 8: 6.06
#endif
    static constexpr int kViewParts = X360_VIEW_PARTS;
    int view_parts = 1;
    cudaEvent_t ev_part[2][kViewParts] = {};
    uint8_t *d_frames[2] = {}, *d_views[2] = {}, *d_sharp[2] = {};
    bool sharpen = false;                // x360_extract_host returns unsharp-masked views (scores stay those of the plain views)
    double sharpen_strength = 0.5;
    long long *d_sums[2] = {};
    bool pipe_ready = false;
    // encoded host path (x360_extract_host_jpeg): the chunk's views are JPEG-encoded on the device, only the files come back
    x360_jpeg *jpeg = nullptr;
    uint8_t *d_jpeg[2] = {};
    long long *d_joff[2] = {}, *h_joff[2] = {};
    long long *h_sums[2] = {};           // filtered host path: the chunk's sum pairs, pinned
    int up_row0 = 0, up_rows = 0;         // the band of source rows the host entries upload (x360_plan_upload_rows)
    long long *h_sum_stage = nullptr;    // x360_extract_host: the call's sum pairs, pinned (a copy into the caller's pageable arrays
    size_t h_sum_stage_cap = 0;          // would block the host until the chunk's download ends, and the next upload behind it)
    long long jpeg_cap = 0;
    cudaStream_t s_off = nullptr;
    cudaEvent_t ev_off[2] = {};
};

static int tiled_occupancy(const x360_plan *pl, const TiledGroup &G, int sc, size_t smem)
{
    int occ = 0;
    const bool lz = pl->p.interp == X360_INTERP_LANCZOS4;
    const cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pick_tiled(sc != 0, lz, G.t_th[sc], G.col_u0, G.t_raw[sc]),
                                                                        32 * tiled_warps(sc != 0, G.t_th[sc]), smem);
    if (e != cudaSuccess || occ < 1) { cudaGetLastError(); occ = 1; }
    return occ;
}

static int plan_grid(const x360_plan *pl, bool scores)
{
    if (pl->path == X360_PATH_TILED) return pl->grp[0].grid_cap[scores && !pl->split_score ? 1 : 0];   // capped by the item count at launch
#ifdef X360_LAB
    int occ = 0;
    cudaError_t e;
    if (pl->path == X360_PATH_STAGED)
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pick_staged(scores), kThreads, pl->smem);
    else
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pick_direct(pl->p.interp, scores), kThreads, 0);
    if (e != cudaSuccess || occ < 1) occ = 1;
    const int n_tiles = pl->p.n_views * pl->g.tiles_x * pl->g.tiles_y;
    const int cap = pl->sm_count * occ;
    return n_tiles < cap ? n_tiles : cap;
#else
    return 0;
#endif
}

// Tensor maps of a (frames, batch size, out) triple at one staging capacity: encoded when the triple changes and kept
// (the pipelined host path alternates between two triples; a device-resident caller usually has one).
static const MapCache *tiled_maps(x360_plan *pl, const TiledGroup &G, MapCache &mc, const uint8_t *frames, int n_frames, uint8_t *out_views, int sc,
                                  int t_rows, const int *wide_rows)
{
    if (mc.frames == frames && mc.out == out_views && mc.n_frames == n_frames) return &mc;
    const Geom &g = pl->g;
    const int t_th = G.t_th[sc];
    mc.frames = nullptr;                               // invalid until the encode below has succeeded
    mc.stage_ok = ((uintptr_t)frames % 16 == 0) && (((size_t)g.W * 3) % 16 == 0) && g.H >= t_rows && g.W * 3 >= 256;
    mc.store_ok = ((uintptr_t)out_views % 16 == 0) && (((size_t)g.ow * 3) % 16 == 0) && g.ow >= kTile && g.oh >= t_th;
    if (lab_int("X360_NO_TMA_STORE", 0)) mc.store_ok = 0;
    mc.wide_ok = mc.stage_ok && !pl->opts.no_wide;
    const bool seam_wanted = mc.wide_ok;
    if (lab_int("X360_NO_WIDE_CLASSES", 0)) mc.wide_ok = false;
    mc.seam_ok = seam_wanted && G.t_raw[sc] && g.W * 3 >= 2 * kTSeamHalf && (long long)n_frames * g.H >= 2 && !lab_int("X360_NO_SEAM", 0);
    memset(&mc.tms, 0, sizeof mc.tms);
    if ((mc.stage_ok || mc.store_ok) &&
        !encode_tiled_maps(frames, n_frames, g.H, g.W, t_rows, mc.stage_ok != 0, out_views, n_frames * pl->p.n_views, g.oh, g.ow, t_th, mc.store_ok != 0,
                           &mc.tms, G.t_raw[sc], G.col_u0, wide_rows, &mc.wide_ok, &mc.seam_ok)) {
        mc.stage_ok = 0;                               // every tile goes direct, byte stores
        mc.store_ok = 0;
        mc.wide_ok = mc.seam_ok = false;
    }
    mc.frames = frames; mc.out = out_views; mc.n_frames = n_frames;
    return &mc;
}

static int launch_score_pass(x360_plan *pl, const uint8_t *views, int n, int64_t *sum_lap, int64_t *sum_lap2, cudaStream_t st);

// One group's kernel launch(es) of an extract call.  `scratch`: this call's slot of the ring (already zeroed);
// `gi`: group index (selects the work-item counters inside the slot).
static int launch_group(x360_plan *pl, TiledGroup &G, int gi, unsigned *scratch, const uint8_t *frames, int n_frames, uint8_t *out_views,
                        int64_t *sum_lap, int64_t *sum_lap2, cudaStream_t st, int v_lo = 0, int v_hi = INT_MAX)
{
    const bool scores = sum_lap != nullptr;
    const Geom &g = pl->g;
    TiledArgs a;
    memset(&a, 0, sizeof a);
    a.g = g;
    a.frames = frames;
    a.out = out_views;
    a.sum_lap = (long long *)sum_lap;
    a.sum_lap2 = (long long *)sum_lap2;
    a.R = pl->d_R;
    a.lz_tab = pl->d_lz;
    a.lz_k1 = pl->d_lz_k1;
    a.lz_fix = pl->d_lz_fix;
    a.counter = scratch + 5 + 2 * gi;
    a.tile_modes = scratch;
    a.F = n_frames;
    a.V = pl->p.n_views;
    const int sc = scores ? 1 : 0, t_rows = G.t_rows[sc], t_th = G.t_th[sc];
    a.rows_max = t_rows;
    a.bw_max = G.t_bw[sc];
    a.tiles_y = (g.oh + t_th - 1) / t_th;
    a.col_u0 = G.col_u0 ? 1 : 0;
    a.aligned = ((uintptr_t)frames % 4 == 0) && (g.W % 4 == 0);
    for (int c = 0; c < kTNumWide; ++c) a.wide_rows[c] = G.wide_rows[sc][c];

    // Work items = units x tiles x frame chunks.  The map of a tile is evaluated once per item and the per-view setup once per
    // (item, view), so coarse units and whole batches amortise best — but the persistent grid needs several items per CTA to
    // balance, and a small job (cfg1: 1024 tiles per view) cannot have both.  Pick the unit granularity (8 / 4 / 2 / 1 views)
    // and the chunk count that minimise, in warp instructions per pixel and frame,
    //     (walk + map / (views per unit x frames per chunk) + setup / frames per chunk) x (1 + k / items per CTA).
    // Constants fitted to sweeps on a B200 (profiles/r02_chunking_sweep.txt): yaw-only groups walk 30, map 360, setup 30, k 1;
    // general groups (pole faces: direct tiles, two record loads per pixel) walk 100, map 400, setup 60, k 2.4.
    const int cap = G.grid_cap[sc];
    const long long tiles = (long long)g.tiles_x * a.tiles_y;
    const double c_walk = G.col_u0 ? 30.0 : 100.0, c_map = G.col_u0 ? 360.0 : 400.0, c_setup = G.col_u0 ? 30.0 : 60.0, c_k = G.col_u0 ? 1.0 : 2.4;
    int level = 0, n_chunks = 1;
    double best_cost = 1e300;
    for (int lv = 0; lv < 4; ++lv) {
        if (lv > 0 && G.n_units[lv] == G.n_units[lv - 1]) continue;
        const double vpu = (double)G.n_views / (double)G.n_units[lv];
        for (int ch = 1; ch <= n_frames; ++ch) {
            const int fg = (n_frames + ch - 1) / ch;
            if (ch > 1 && (n_frames + ch - 2) / (ch - 1) == fg) continue;          // same chunk length as with one chunk fewer
            const long long items = (long long)G.n_units[lv] * tiles * ((n_frames + fg - 1) / fg);
            const double cost = (c_walk + c_map / (vpu * fg) + c_setup / fg) * (1.0 + c_k * (double)cap / (double)items);
            if (cost < best_cost * 0.999) { best_cost = cost; level = lv; n_chunks = (n_frames + fg - 1) / fg; }
        }
    }
    if (const char *e = lab_env("X360_UNIT_LEVEL")) level = atoi(e) >= 0 && atoi(e) <= 3 ? atoi(e) : level;
    // a view range [v_lo, v_hi) of the plan: the one-view units of this group's views in it (level 3 lists them in view order)
    int u_begin = 0, n_units = G.n_units[level];
    if (v_lo > 0 || v_hi < pl->p.n_views) {
        level = 3;
        u_begin = (int)(std::lower_bound(G.h_ids.begin(), G.h_ids.end(), v_lo) - G.h_ids.begin());
        n_units = (int)(std::lower_bound(G.h_ids.begin(), G.h_ids.end(), v_hi) - G.h_ids.begin()) - u_begin;
        n_chunks = 1;
        if (n_units <= 0) return X360_OK;
    }
    a.units = G.d_units[level] + u_begin;
    a.unit_view = G.d_unit_view[level];
    a.unit_shift = G.d_unit_shift[level];
    a.n_units = n_units;
    a.rot_inv_pitch = lab_int("X360_NO_ROT", 0) ? 0.0 : G.rot_inv_pitch[level];
    a.tr_tab = (!scores && t_th == 16 && G.t_raw[sc]) ? G.d_tr_tab : nullptr;
    int FG = (n_frames + n_chunks - 1) / n_chunks;
    if (const char *e = lab_env("X360_FG")) FG = atoi(e) > 0 ? atoi(e) : FG;
    if (const char *e = G.col_u0 ? nullptr : lab_env("X360_FG_GENERAL")) FG = atoi(e) > 0 ? atoi(e) : FG;
    if (FG > n_frames) FG = n_frames;
    a.FG = FG;
    a.n_chunks = (n_frames + FG - 1) / FG;
    const long long n_items = (long long)n_units * tiles * a.n_chunks;
    if (n_items > 0x7fffffff) return fail(X360_E_INVALID, "too many work items (%lld)", n_items);
    a.n_items = (int)n_items;

    const bool two = G.two_launch[sc] && pl->d_defer[0] != nullptr;
    const MapCache *mc = tiled_maps(pl, G, G.maps[sc], frames, n_frames, out_views, sc, t_rows, G.wide_rows[sc]);
    const MapCache *mc2 = two ? tiled_maps(pl, G, G.maps2[sc], frames, n_frames, out_views, sc, G.t2_rows[sc], G.wide_rows2[sc]) : nullptr;
    a.stage_ok = mc->stage_ok;
    a.store_ok = mc->store_ok;
    a.wide_ok = mc->wide_ok ? 1 : 0;
    a.seam_ok = mc->seam_ok ? 1 : 0;
    int db = -1;
    if (two && mc->stage_ok && mc2->stage_ok) {
        // the deferred-pair list of this launch pair: one of kDeferBufs buffers; a buffer still in use by an earlier launch
        // pair (another stream) is waited for on this stream
        db = pl->defer_next;
        pl->defer_next = (db + 1) % x360_plan::kDeferBufs;
        if (pl->defer_used[db]) CU(cudaStreamWaitEvent(st, pl->defer_ev[db], 0));
        a.mode = 1;
        a.defer_cap = pl->defer_cap;
        // frame sub-ranges per deferred pair: 1 (the map and per-view setup of a pair are paid once per item: cutting the 32-frame
        // batch in 4 measured 8 % slower on cfg1 and cfg4)
        a.defer_sub = lab_int("X360_DEFER_SUB", 1);
        a.defer_list = pl->d_defer[db];
        a.defer_count = pl->d_defer[db] + 2 * (size_t)pl->defer_cap;
        CU(cudaMemsetAsync(a.defer_count, 0, sizeof(unsigned), st));
    }
    const int grid = (int)std::min<long long>(n_items, cap);
    const tiled_fn fn = pick_tiled(scores, pl->p.interp == X360_INTERP_LANCZOS4, t_th, G.col_u0, G.t_raw[sc]);
    fn<<<grid, 32 * tiled_warps(scores, t_th), G.t_smem[sc], st>>>(a, mc->tms);
    g_launches.fetch_add(1);
    CU(cudaGetLastError());
    if (db >= 0) {
        a.mode = 2;
        a.counter = scratch + 6 + 2 * gi;
        a.rows_max = G.t2_rows[sc];
        a.bw_max = G.t2_bw[sc];
        for (int c = 0; c < kTNumWide; ++c) a.wide_rows[c] = G.wide_rows2[sc][c];
        a.stage_ok = mc2->stage_ok;
        a.store_ok = mc2->store_ok;
        a.wide_ok = mc2->wide_ok ? 1 : 0;
        a.seam_ok = mc2->seam_ok ? 1 : 0;
        fn<<<G.grid_cap2[sc], 32 * tiled_warps(scores, t_th), G.t2_smem[sc], st>>>(a, mc2->tms);
        g_launches.fetch_add(1);
        CU(cudaGetLastError());
        CU(cudaEventRecord(pl->defer_ev[db], st));
        pl->defer_used[db] = true;
    }
    return X360_OK;
}

static int launch_tiled(x360_plan *pl, const uint8_t *frames, int n_frames, uint8_t *out_views, int64_t *sum_lap,
                        int64_t *sum_lap2, cudaStream_t st, int v_lo = 0, int v_hi = INT_MAX)
{
    if (sum_lap && pl->split_score) {
        // extraction without the fused score, then one streaming pass over the finished views
        if (int rc = launch_tiled(pl, frames, n_frames, out_views, nullptr, nullptr, st)) return rc;
        return launch_score_pass(pl, out_views, n_frames * pl->p.n_views, sum_lap, sum_lap2, st);
    }
    // this call's slot of the scratch ring (work counters + tile-mode counters), zeroed on the launch stream
    const int slot = pl->ring_next;
    pl->ring_next = (slot + 1) % kRingSlots;
    pl->ring_last = slot;
    unsigned *scratch = pl->d_modes + (size_t)slot * kRingWords;
    CU(cudaMemsetAsync(scratch, 0, sizeof(unsigned) * kRingWords, st));
    if (sum_lap) {
        CU(cudaMemsetAsync(sum_lap, 0, sizeof(int64_t) * (size_t)n_frames * pl->p.n_views, st));
        CU(cudaMemsetAsync(sum_lap2, 0, sizeof(int64_t) * (size_t)n_frames * pl->p.n_views, st));
    }
    if (pl->n_groups < 2 || lab_int("X360_SERIAL_GROUPS", 0))  {
        for (int gi = 0; gi < pl->n_groups; ++gi)
            if (int rc = launch_group(pl, pl->grp[gi], gi, scratch, frames, n_frames, out_views, sum_lap, sum_lap2, st, v_lo, v_hi)) return rc;
        return X360_OK;
    }
    // Two groups: the second one (pole faces: the longer items) goes to a side stream forked from the caller's, so that each
    // persistent grid fills the SMs the other one leaves idle at its tail; joined before returning to the caller's stream.
    if (!pl->s_side) {
        CU(cudaStreamCreateWithFlags(&pl->s_side, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&pl->ev_fork, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&pl->ev_join, cudaEventDisableTiming));
    }
    CU(cudaEventRecord(pl->ev_fork, st));
    CU(cudaStreamWaitEvent(pl->s_side, pl->ev_fork, 0));
    int rc = launch_group(pl, pl->grp[1], 1, scratch, frames, n_frames, out_views, sum_lap, sum_lap2, pl->s_side, v_lo, v_hi);
    if (rc == X360_OK) rc = launch_group(pl, pl->grp[0], 0, scratch, frames, n_frames, out_views, sum_lap, sum_lap2, st, v_lo, v_hi);
    const std::string first_error = rc != X360_OK ? g_err : std::string();
    const cudaError_t e1 = cudaEventRecord(pl->ev_join, pl->s_side), e2 = cudaStreamWaitEvent(st, pl->ev_join, 0);   // always re-join
    if (rc != X360_OK) { g_err = first_error; return rc; }
    if (e1 != cudaSuccess || e2 != cudaSuccess) return fail(X360_E_CUDA, "joining the group streams failed: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
    return X360_OK;
}

// The split score: one streaming pass over finished views (blur_sums_views_kernel, x360_misc.cuh)
static int launch_score_pass(x360_plan *pl, const uint8_t *views, int n, int64_t *sum_lap, int64_t *sum_lap2, cudaStream_t st)
{
    const Geom &g = pl->g;
    CU(cudaMemsetAsync(sum_lap, 0, sizeof(int64_t) * (size_t)n, st));
    CU(cudaMemsetAsync(sum_lap2, 0, sizeof(int64_t) * (size_t)n, st));
    const int fast = ((uintptr_t)views % 4 == 0) && (g.ow % 4 == 0);
    const bool march = fast && g.ow >= 8 && g.oh >= 2 && !lab_int("X360_SCORE_TILES", 0);
    for (int n0 = 0; n0 < n; n0 += 32768) {
        const int m = std::min(32768, n - n0);
        const uint8_t *v0 = views + (size_t)n0 * g.oh * g.ow * 3;
        if (march)
            blur_sums_march_kernel<<<dim3((unsigned)((g.ow + kScStride - 1) / kScStride), (unsigned)((g.oh + kScWarps * kScRows - 1) / (kScWarps * kScRows)), (unsigned)m),
                                     32 * kScWarps, 0, st>>>(v0, g.oh, g.ow, (long long *)sum_lap + n0, (long long *)sum_lap2 + n0);
        else
            blur_sums_views_kernel<<<dim3((unsigned)((g.ow + kBsW - 1) / kBsW), (unsigned)((g.oh + kBsH - 1) / kBsH), (unsigned)m), 256, 0, st>>>(
                v0, g.oh, g.ow, fast, (long long *)sum_lap + n0, (long long *)sum_lap2 + n0);
        g_launches.fetch_add(1);
    }
    CU(cudaGetLastError());
    return X360_OK;
}

extern "C" {

const char *x360_last_error(void) { return g_err.c_str(); }
const char *x360_version(void) { return "x360 0.1.0 (sm_100a; abi 1)"; }
int64_t x360_launch_count(void) { return (int64_t)g_launches.load(); }

int x360_lanczos4_table(int16_t *out)
{
    if (!out) return fail(X360_E_INVALID, "null table pointer");
    const std::vector<int16_t> t = lanczos4_q15_table();
    memcpy(out, t.data(), t.size() * sizeof(int16_t));
    return X360_OK;
}

int x360_lanczos4_factors(float *k1d, uint32_t *fix)
{
    if (!k1d || !fix) return fail(X360_E_INVALID, "null pointer");
    std::vector<float> k;
    std::vector<uint32_t> f;
    lanczos4_factors(k, f);
    memcpy(k1d, k.data(), k.size() * sizeof(float));
    memcpy(fix, f.data(), f.size() * sizeof(uint32_t));
    return X360_OK;
}

double x360_focal(int32_t out_w, double fov_deg) { return focal_of(out_w, fov_deg); }

int x360_plan_preview(const x360_params *params, int32_t *n_units, int32_t *unit_of_view, double *shift_px,
                      int32_t *rows_max, int32_t *bw_max, int32_t *tile_h)
{
    if (!params || !params->R || params->n_views < 1) return fail(X360_E_INVALID, "null argument");
    if (params->struct_size != (int32_t)sizeof(x360_params)) return fail(X360_E_INVALID, "x360_params.struct_size mismatch");
    if (int rc = check_dims(params->src_w, params->src_h, params->out_w, params->out_h)) return rc;
    if (!(params->fov_deg > 0.0 && params->fov_deg < 180.0)) return fail(X360_E_INVALID, "fov_deg %g outside (0, 180)", params->fov_deg);
    const HostUnits h = build_units(params->R, params->n_views, params->src_w,
                                    make_geom(params->src_w, params->src_h, params->out_w, params->out_h, params->fov_deg), true);
    if (n_units) *n_units = (int32_t)h.units.size();
    for (size_t u = 0; u < h.units.size(); ++u)
        for (int i = 0; i < h.units[u].n; ++i) {
            const int e = h.units[u].first + i;
            if (unit_of_view) unit_of_view[h.view[e]] = (int32_t)u;
            if (shift_px) shift_px[h.view[e]] = h.shift[e];
        }
    // capacity / tile height of the group that serves view 0 (x360_plan_preview_views has them per view)
    std::vector<double> Rg;
    const bool col0 = column_only_map_x(params->R, 1);
    for (int v = 0; v < params->n_views; ++v)
        if (column_only_map_x(params->R + 9 * v, 1) == col0) Rg.insert(Rg.end(), params->R + 9 * v, params->R + 9 * v + 9);
    for (int sc = 0; sc < 2; ++sc) {
        int rows = 0, bw = 0, th = 0;
        choose_capacity(make_geom(params->src_w, params->src_h, params->out_w, params->out_h, params->fov_deg), Rg.data(),
                        (int)(Rg.size() / 9), params->interp == X360_INTERP_LANCZOS4, sc != 0, plan_opts(params->flags), &rows, &bw, &th);
        if (rows_max) rows_max[sc] = rows;
        if (bw_max) bw_max[sc] = bw;
        if (tile_h) tile_h[sc] = th;
    }
    return X360_OK;
}

int x360_plan_preview_views(const x360_params *params, int32_t *group_of_view, int32_t *tile_h_of_view)
{
    if (!params || !params->R || params->n_views < 1) return fail(X360_E_INVALID, "null argument");
    if (params->struct_size != (int32_t)sizeof(x360_params)) return fail(X360_E_INVALID, "x360_params.struct_size mismatch");
    if (int rc = check_dims(params->src_w, params->src_h, params->out_w, params->out_h)) return rc;
    if (!(params->fov_deg > 0.0 && params->fov_deg < 180.0)) return fail(X360_E_INVALID, "fov_deg %g outside (0, 180)", params->fov_deg);
    const Geom g = make_geom(params->src_w, params->src_h, params->out_w, params->out_h, params->fov_deg);
    int next_group = 0;
    for (int cls = 0; cls < 2; ++cls) {
        std::vector<double> Rg;
        std::vector<int> ids;
        for (int v = 0; v < params->n_views; ++v)
            if (column_only_map_x(params->R + 9 * v, 1) == (cls == 0)) { ids.push_back(v); Rg.insert(Rg.end(), params->R + 9 * v, params->R + 9 * v + 9); }
        if (ids.empty()) continue;
        int th[2] = {32, 32};
        for (int sc = 0; sc < 2; ++sc) {
            int rows = 0, bw = 0;
            choose_capacity(g, Rg.data(), (int)ids.size(), params->interp == X360_INTERP_LANCZOS4, sc != 0, plan_opts(params->flags), &rows, &bw, &th[sc]);
        }
        for (int v : ids) {
            if (group_of_view) group_of_view[v] = next_group;
            if (tile_h_of_view) { tile_h_of_view[2 * v] = th[0]; tile_h_of_view[2 * v + 1] = th[1]; }
        }
        ++next_group;
    }
    return X360_OK;
}

int x360_plan_create(const x360_params *params, x360_plan **out_plan)
{
    if (!params || !out_plan) return fail(X360_E_INVALID, "null argument");
    if (params->struct_size != (int32_t)sizeof(x360_params))
        return fail(X360_E_INVALID, "x360_params.struct_size %d != %zu (ABI mismatch)", params->struct_size, sizeof(x360_params));
    *out_plan = nullptr;
    if (int rc = check_dims(params->src_w, params->src_h, params->out_w, params->out_h)) return rc;
    if (params->n_views < 1 || !params->R) return fail(X360_E_INVALID, "need at least one view and its rotation matrix");
    if (!(params->fov_deg > 0.0 && params->fov_deg < 180.0)) return fail(X360_E_INVALID, "fov_deg %g outside (0, 180)", params->fov_deg);
    if (params->interp != X360_INTERP_LINEAR && params->interp != X360_INTERP_LANCZOS4)
        return fail(X360_E_INVALID, "unknown interp %d", params->interp);
    if (params->path < X360_PATH_AUTO || params->path > X360_PATH_TILED) return fail(X360_E_INVALID, "unknown path %d", params->path);
#ifdef X360_LAB
    if (params->path == X360_PATH_STAGED && params->interp == X360_INTERP_LANCZOS4)
        return fail(X360_E_UNSUPPORTED, "the staged path is built for bilinear only; Lanczos4 uses the tiled or the direct path");
#else
    if (params->path == X360_PATH_DIRECT || params->path == X360_PATH_STAGED)
        return fail(X360_E_UNSUPPORTED, "path %d is a comparison kernel, compiled only into -DX360_LAB builds of libx360", params->path);
#endif
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev < 1)
        return fail(X360_E_CUDA, "no CUDA device available (libx360 has no CPU fallback)");
    if (params->device < 0 || params->device >= n_dev) return fail(X360_E_INVALID, "device %d out of range [0, %d)", params->device, n_dev);
    CU(cudaSetDevice(params->device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, params->device));
    if (prop.major != 10)
        return fail(X360_E_UNSUPPORTED, "device %d is sm_%d%d; libx360 is built for sm_100a only", params->device, prop.major, prop.minor);

    x360_plan *pl = new (std::nothrow) x360_plan();
    if (!pl) return fail(X360_E_NOMEM, "out of host memory");
    pl->p = *params;
    pl->p.R = nullptr;
    pl->g = make_geom(params->src_w, params->src_h, params->out_w, params->out_h, params->fov_deg);
    pl->sm_count = prop.multiProcessorCount;
    pl->chunk = params->max_batch > 0 ? params->max_batch : 2;
    pl->path = params->path == X360_PATH_DIRECT ? X360_PATH_DIRECT
               : (params->path == X360_PATH_STAGED ? X360_PATH_STAGED : X360_PATH_TILED);
    if (pl->path == X360_PATH_TILED) {
        const bool lz = params->interp == X360_INTERP_LANCZOS4;
        pl->opts = plan_opts(params->flags);
        pl->split_score = score_is_split(pl->opts);
        // groups: yaw-only views (column-only map_x) / the others
        std::vector<int> members[2];
        const bool no_col = lab_int("X360_NO_COL_U0", 0) != 0, one_group = lab_int("X360_ONE_GROUP", 0) != 0;
        const bool all_col = column_only_map_x(params->R, params->n_views) && !no_col;
        for (int v = 0; v < params->n_views; ++v) {
            const bool col = one_group ? all_col : (column_only_map_x(params->R + 9 * v, 1) && !no_col);
            members[col ? 0 : 1].push_back(v);
        }
        pl->view_group.assign((size_t)params->n_views, 0);
        for (int cls = 0; cls < 2; ++cls) {
            if (members[cls].empty()) continue;
            TiledGroup &G = pl->grp[pl->n_groups];
            for (int v : members[cls]) pl->view_group[(size_t)v] = pl->n_groups;
            ++pl->n_groups;
            G.n_views = (int)members[cls].size();
            G.col_u0 = cls == 0;
            std::vector<double> Rg((size_t)G.n_views * 9);
            for (int i = 0; i < G.n_views; ++i) memcpy(&Rg[(size_t)i * 9], params->R + 9 * members[cls][(size_t)i], 9 * sizeof(double));
            // staging capacity and tile height from a host survey of the group's tile footprints, without / with the score halo
            for (int sc = 0; sc < 2; ++sc) {
                int rows = 40, bw = 144, th = 32;
                long long unfit = 0;
                choose_capacity(pl->g, Rg.data(), G.n_views, lz, sc != 0, pl->opts, &rows, &bw, &th, &unfit);
                const bool raw = tiled_raw(sc != 0, lz, pl->opts);
                G.t_raw[sc] = raw;
                const bool lab_cap = !(lab_int("X360_TILED_GENERAL_ONLY", 0) && G.col_u0);
                if (const char *e = lab_cap ? lab_env("X360_TILED_BW") : nullptr) {
                    for (int wc = 0; wc < (raw ? 4 : kTNumBoxW); ++wc)
                        if (tiled_box_bw(raw, G.col_u0, wc) == atoi(e)) bw = atoi(e);
                }
                if (const char *e = lab_cap ? lab_env("X360_TILED_ROWS") : nullptr) rows = round_up(atoi(e) < 24 ? 24 : atoi(e), kTBoxRows);
                while (rows > 24 && tiled_smem_bytes(rows, bw, sc != 0, lz, th, G.col_u0, raw) > (size_t)kTiledSmemCap) rows -= kTBoxRows;
                G.t_th[sc] = th;
                G.t_bw[sc] = bw;
                G.t_rows[sc] = rows;
                wide_class_rows(rows, bw, G.wide_rows[sc]);
                G.t_smem[sc] = tiled_smem_bytes(rows, bw, sc != 0, lz, th, G.col_u0, raw);
                // the limit belongs to the kernel function, not to this plan: always the common cap (see kTiledSmemCap)
                const cudaError_t ea = cudaFuncSetAttribute(pick_tiled(sc != 0, lz, th, G.col_u0, raw), cudaFuncAttributeMaxDynamicSharedMemorySize, kTiledSmemCap);
                if (ea != cudaSuccess) {
                    delete pl;
                    return fail(X360_E_CUDA, "cannot reserve %d B of shared memory: %s", kTiledSmemCap, cudaGetErrorString(ea));
                }
                G.grid_cap[sc] = pl->sm_count * tiled_occupancy(pl, G, sc, G.t_smem[sc]);
                // A second launch with 16 KB staging buffers for the footprints beyond this capacity (pole faces), where the
                // survey finds enough of them to fill that launch's grid several times over: measured neutral on cfg4 (~5000
                // pairs, direct share 12 % -> 2 %), 10 % slower on cfg1 (~1100 pairs: two half-empty waves), so small jobs keep
                // the direct gathers for those tiles.
                // Lanczos4 (single-frame jobs, 8 x 8 taps: a direct tile costs 64 global loads per pixel): 80-row buffers, used from
                // one wave of such pairs on — cfg3 21.5 -> 23.0 Gpix/s, direct share 11.5 % -> 3 %.
                const int rows2 = lab_int("X360_T2_ROWS", lz ? 80 : 64), bw2 = lab_int("X360_T2_BW", raw ? tiled_box_bw(true, G.col_u0, 3) : 240);
                if (!(raw && G.col_u0) && !pl->opts.no_wide && rows2 * bw2 > rows * bw && !lab_int("X360_ONE_LAUNCH", 0)) {
                    G.t2_rows[sc] = rows2;
                    G.t2_bw[sc] = bw2;
                    wide_class_rows(rows2, bw2, G.wide_rows2[sc]);
                    G.t2_smem[sc] = tiled_smem_bytes(rows2, bw2, sc != 0, lz, th, G.col_u0, raw);
                    if (G.t2_smem[sc] <= (size_t)kTiledSmemCap) {
                        G.grid_cap2[sc] = pl->sm_count * tiled_occupancy(pl, G, sc, G.t2_smem[sc]);
                        G.two_launch[sc] = unfit >= (long long)lab_int("X360_DEFER_MIN_WAVES", lz ? 1 : 4) * G.grid_cap2[sc];
                    }
                }
            }
        }
    }
#ifdef X360_LAB
    if (pl->path == X360_PATH_STAGED) {
        // Footprint capacity: a 34-px tile side (32 + halo) times the source/output scale at the view
        // centre, with head-room for shear and latitude stretch; anything larger goes direct.
        const double scale = ((double)params->src_w / (2.0 * 3.14159265358979323846)) / pl->g.f;
        double factor = 1.4;
        if (const char *e = lab_env("X360_STAGE_FACTOR")) factor = atof(e) > 0.5 ? atof(e) : factor;
        int cap = round_up((int)std::ceil((kTile + 2) * scale * factor) + 8, 4);
        cap = cap < 16 ? 16 : (cap > 128 ? 128 : cap);
        for (;; cap -= 4) {                                   // shrink until two CTAs fit one SM
            pl->px_max = cap; pl->rows_max = cap; pl->pp = cap + 2; pl->raw_pitch = kMaxBoxBytes;
            pl->smem = staged_smem_bytes(pl->rows_max, pl->pp);
            if (pl->smem <= 110 * 1024 || cap <= 16) break;
        }
        for (int sc = 0; sc < 2; ++sc) {
            const cudaError_t ea = cudaFuncSetAttribute(pick_staged(sc != 0), cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
            if (ea != cudaSuccess) {
                delete pl;
                return fail(X360_E_CUDA, "cannot reserve %zu B of shared memory: %s", pl->smem, cudaGetErrorString(ea));
            }
        }
    }
#endif
    cudaError_t e = cudaMalloc(&pl->d_R, sizeof(double) * 9 * params->n_views);
    if (e == cudaSuccess) e = cudaMalloc(&pl->d_modes, sizeof(unsigned) * kRingSlots * kRingWords);
    if (e == cudaSuccess) e = cudaMemset(pl->d_modes, 0, sizeof(unsigned) * kRingSlots * kRingWords);
    bool any_two = false;
    for (int gi = 0; gi < pl->n_groups; ++gi) any_two = any_two || pl->grp[gi].two_launch[0] || pl->grp[gi].two_launch[1];
    if (e == cudaSuccess && any_two) {
        const long long pairs = (long long)params->n_views * pl->g.tiles_x * ((pl->g.oh + 15) / 16) * 8;      // (tile, view, frame chunk) triples, generously
        pl->defer_cap = (int)std::min<long long>(pairs, 1 << 18);
        for (int i = 0; i < x360_plan::kDeferBufs && e == cudaSuccess; ++i) {
            e = cudaMalloc(&pl->d_defer[i], sizeof(unsigned) * (2 * (size_t)pl->defer_cap + 1));
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&pl->defer_ev[i], cudaEventDisableTiming);
        }
    }
    if (e == cudaSuccess && pl->path == X360_PATH_TILED) {
        const bool share = !lab_int("X360_NO_SHARE", 0);
        for (int gi = 0; gi < pl->n_groups && e == cudaSuccess; ++gi) {
            TiledGroup &G = pl->grp[gi];
            std::vector<int> ids;
            for (int v = 0; v < params->n_views; ++v)
                if (pl->view_group[(size_t)v] == gi) ids.push_back(v);
            G.h_ids = ids;
            std::vector<double> Rg(ids.size() * 9);
            for (size_t i = 0; i < ids.size(); ++i) memcpy(&Rg[i * 9], params->R + 9 * ids[i], 9 * sizeof(double));
            // transposed-tile table of the group's views (TiledArgs::tr_tab), where the no-score launch can take them
            if (G.t_th[0] == 16 && transposable_plan(pl->g, G.col_u0, G.t_raw[0], false, params->interp == X360_INTERP_LANCZOS4)) {
                const int bxn = pl->g.tiles_x, byn = pl->g.oh / kTile;
                std::vector<uint8_t> tab((size_t)params->n_views * bxn * byn, 0);
                size_t n_tr = 0;
                for (int v : ids)
                    for (int by = 0; by < byn; ++by)
                        for (int bx = 0; bx < bxn; ++bx)
                            n_tr += (tab[((size_t)v * byn + by) * bxn + bx] = block_transposed(pl->g, params->R + 9 * v, bx, by) ? 1 : 0);
                if (n_tr) {
                    G.n_tr_blocks = (long long)n_tr;
                    e = cudaMalloc(&G.d_tr_tab, tab.size());
                    if (e == cudaSuccess) e = cudaMemcpy(G.d_tr_tab, tab.data(), tab.size(), cudaMemcpyHostToDevice);
                }
            }
            for (int lv = 0; lv < 4 && e == cudaSuccess; ++lv) {
                HostUnits h = build_units(Rg.data(), (int)ids.size(), params->src_w, pl->g, share, kTMaxUnitViews >> lv);
                for (int &v : h.view) v = ids[(size_t)v];                       // group-local -> plan view index
                G.n_units[lv] = (int)h.units.size();
                G.rot_inv_pitch[lv] = G.col_u0 ? even_yaw_spacing(h, params->src_w) : 0.0;
                e = cudaMalloc(&G.d_units[lv], sizeof(TiledUnit) * h.units.size());
                if (e == cudaSuccess) e = cudaMalloc(&G.d_unit_view[lv], sizeof(int) * h.view.size());
                if (e == cudaSuccess) e = cudaMalloc(&G.d_unit_shift[lv], sizeof(double) * h.shift.size());
                if (e == cudaSuccess) e = cudaMemcpy(G.d_units[lv], h.units.data(), sizeof(TiledUnit) * h.units.size(), cudaMemcpyHostToDevice);
                if (e == cudaSuccess) e = cudaMemcpy(G.d_unit_view[lv], h.view.data(), sizeof(int) * h.view.size(), cudaMemcpyHostToDevice);
                if (e == cudaSuccess) e = cudaMemcpy(G.d_unit_shift[lv], h.shift.data(), sizeof(double) * h.shift.size(), cudaMemcpyHostToDevice);
            }
        }
    }
    if (e == cudaSuccess) e = cudaMemcpy(pl->d_R, params->R, sizeof(double) * 9 * params->n_views, cudaMemcpyHostToDevice);
    // the band of source rows the views can read: min / max of floor(map_y) over every output pixel (the kernel's own map),
    // widened by the taps and two rows of margin; a band that wraps (a view sees a pole) or covers most of the frame: all rows
    pl->up_row0 = 0;
    pl->up_rows = params->src_h;
    if (e == cudaSuccess && !pl->opts.full_upload) {
        const int init[2] = {INT_MAX, INT_MIN};
        int lohi[2] = {0, 0};
        int *d_lohi = reinterpret_cast<int *>(pl->d_modes);                 // (scratch, zeroed again below)
        e = cudaMemcpy(d_lohi, init, sizeof init, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) {
            source_row_range_kernel<<<dim3((unsigned)((pl->g.ow + 31) / 32), (unsigned)((pl->g.oh + 7) / 8), (unsigned)params->n_views), 256>>>(pl->g, pl->d_R, d_lohi);
            g_launches.fetch_add(1);
            e = cudaMemcpy(lohi, d_lohi, sizeof lohi, cudaMemcpyDeviceToHost);
        }
        if (e == cudaSuccess) e = cudaMemset(pl->d_modes, 0, sizeof(unsigned) * kRingSlots * kRingWords);
        if (e == cudaSuccess) {
            const bool lz = params->interp == X360_INTERP_LANCZOS4;
            const int t_lo = lz ? 3 : 0, t_hi = lz ? 4 : 1, margin = 2;
            if (lohi[0] - t_lo >= 0 && lohi[1] + t_hi <= params->src_h - 1) {
                const int r0 = std::max(0, lohi[0] - t_lo - margin), r1 = std::min(params->src_h - 1, lohi[1] + t_hi + margin);
                if ((long long)(r1 - r0 + 1) * 10 < (long long)params->src_h * 9) { pl->up_row0 = r0; pl->up_rows = r1 - r0 + 1; }
            }
        }
    }
    if (e == cudaSuccess && params->interp == X360_INTERP_LANCZOS4) {
        const std::vector<int16_t> t = lanczos4_q15_table();
        e = cudaMalloc(&pl->d_lz, t.size() * sizeof(int16_t));
        if (e == cudaSuccess) e = cudaMemcpy(pl->d_lz, t.data(), t.size() * sizeof(int16_t), cudaMemcpyHostToDevice);
        std::vector<float> k1d;
        std::vector<uint32_t> fix;
        lanczos4_factors(k1d, fix);
        if (e == cudaSuccess) e = cudaMalloc(&pl->d_lz_k1, k1d.size() * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&pl->d_lz_fix, fix.size() * sizeof(uint32_t));
        if (e == cudaSuccess) e = cudaMemcpy(pl->d_lz_k1, k1d.data(), k1d.size() * sizeof(float), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(pl->d_lz_fix, fix.data(), fix.size() * sizeof(uint32_t), cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess && pl->path == X360_PATH_TILED && params->n_views >= 4) {
        bool shared_maps = false;                                           // some unit holds several views: one launch amortises their map
        for (int gi = 0; gi < pl->n_groups; ++gi) shared_maps = shared_maps || pl->grp[gi].n_units[0] != pl->grp[gi].n_views;
        if (!shared_maps) pl->view_parts = std::min<int>(x360_plan::kViewParts, params->n_views);
    }
    if (e != cudaSuccess) {
        x360_plan_destroy(pl);
        return fail(X360_E_CUDA, "plan allocation failed: %s", cudaGetErrorString(e));
    }
    *out_plan = pl;
    return X360_OK;
}

int x360_plan_destroy(x360_plan *pl)
{
    if (!pl) return X360_OK;
    cudaSetDevice(pl->p.device);
    cudaFree(pl->d_R);
    cudaFree(pl->d_lz);
    cudaFree(pl->d_lz_k1);
    cudaFree(pl->d_lz_fix);
    cudaFree(pl->d_modes);
    for (int i = 0; i < x360_plan::kDeferBufs; ++i) {
        cudaFree(pl->d_defer[i]);
        if (pl->defer_ev[i]) cudaEventDestroy(pl->defer_ev[i]);
    }
    for (int gi = 0; gi < 2; ++gi)
        for (int lv = 0; lv < 4; ++lv) {
            cudaFree(pl->grp[gi].d_units[lv]);
            cudaFree(pl->grp[gi].d_unit_view[lv]);
            cudaFree(pl->grp[gi].d_unit_shift[lv]);
        }
    for (int gi = 0; gi < 2; ++gi) cudaFree(pl->grp[gi].d_tr_tab);
    for (int i = 0; i < 2; ++i) {
        cudaFree(pl->d_frames[i]);
        cudaFree(pl->d_views[i]);
        cudaFree(pl->d_sharp[i]);
        cudaFree(pl->d_sums[i]);
        if (pl->ev_in[i]) cudaEventDestroy(pl->ev_in[i]);
        if (pl->ev_k[i]) cudaEventDestroy(pl->ev_k[i]);
        if (pl->ev_out[i]) cudaEventDestroy(pl->ev_out[i]);
        for (int k = 0; k < x360_plan::kViewParts; ++k)
            if (pl->ev_part[i][k]) cudaEventDestroy(pl->ev_part[i][k]);
    }
    if (pl->jpeg) x360_jpeg_destroy(pl->jpeg);
    for (int i = 0; i < 2; ++i) {
        cudaFree(pl->d_jpeg[i]);
        cudaFree(pl->d_joff[i]);
        if (pl->h_joff[i]) cudaFreeHost(pl->h_joff[i]);
        if (pl->h_sums[i]) cudaFreeHost(pl->h_sums[i]);
        if (i == 0 && pl->h_sum_stage) cudaFreeHost(pl->h_sum_stage);
        if (pl->ev_off[i]) cudaEventDestroy(pl->ev_off[i]);
    }
    if (pl->s_off) cudaStreamDestroy(pl->s_off);
    if (pl->s_side) cudaStreamDestroy(pl->s_side);
    if (pl->ev_fork) cudaEventDestroy(pl->ev_fork);
    if (pl->ev_join) cudaEventDestroy(pl->ev_join);
    if (pl->s_in) cudaStreamDestroy(pl->s_in);
    if (pl->s_k) cudaStreamDestroy(pl->s_k);
    if (pl->s_out) cudaStreamDestroy(pl->s_out);
    delete pl;
    return X360_OK;
}

int x360_column_only_map_x(const double *R, int32_t n_views)
{
    if (!R || n_views < 1) return 0;
    return column_only_map_x(R, n_views) ? 1 : 0;
}

int x360_plan_set_sharpen(x360_plan *pl, int32_t enabled, double strength)
{
    if (!pl) return fail(X360_E_INVALID, "null plan");
    if (!(strength == strength) || std::fabs(strength) > 1e6) return fail(X360_E_INVALID, "bad strength %g", strength);
    pl->sharpen = enabled != 0;
    pl->sharpen_strength = strength;
    return X360_OK;
}

int x360_plan_info(const x360_plan *pl, int32_t *path, int32_t *grid, int32_t *block, int32_t *smem_bytes)
{
    if (!pl) return fail(X360_E_INVALID, "null plan");
    if (path) *path = pl->path;
    if (grid) *grid = plan_grid(pl, false);
    if (block) *block = pl->path == X360_PATH_TILED ? 32 * tiled_warps(false, pl->grp[0].t_th[0]) : kThreads;
    if (smem_bytes) *smem_bytes = pl->path == X360_PATH_TILED ? (int32_t)pl->grp[0].t_smem[0]
                                  : (pl->path == X360_PATH_STAGED ? (int32_t)pl->smem : (int32_t)sizeof(TileSmem));
    return X360_OK;
}

int x360_plan_upload_rows(const x360_plan *pl, int32_t *first_row, int32_t *n_rows)
{
    if (!pl) return fail(X360_E_INVALID, "null argument");
    if (first_row) *first_row = pl->up_row0;
    if (n_rows) *n_rows = pl->up_rows;
    return X360_OK;
}

int x360_plan_transposed_blocks(const x360_plan *pl, int64_t *n_blocks)
{
    if (!pl || !n_blocks) return fail(X360_E_INVALID, "null argument");
    *n_blocks = 0;
    for (int gi = 0; gi < pl->n_groups; ++gi) *n_blocks += pl->grp[gi].n_tr_blocks;
    return X360_OK;
}

int x360_plan_tile_modes(x360_plan *pl, uint32_t *out5)
{
    if (!pl || !out5) return fail(X360_E_INVALID, "null argument");
    CU(cudaSetDevice(pl->p.device));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(out5, pl->d_modes + (size_t)pl->ring_last * kRingWords, sizeof(unsigned) * 5, cudaMemcpyDeviceToHost));
    return X360_OK;
}

int x360_extract_device(x360_plan *pl, const uint8_t *frames, int32_t n_frames, uint8_t *out_views,
                        int64_t *sum_lap, int64_t *sum_lap2, void *cuda_stream)
{
    if (!pl || !frames || !out_views) return fail(X360_E_INVALID, "null argument");
    if (n_frames < 0) return fail(X360_E_INVALID, "negative frame count");
    if ((sum_lap == nullptr) != (sum_lap2 == nullptr)) return fail(X360_E_INVALID, "sum_lap and sum_lap2 must both be given or both be NULL");
    if (n_frames == 0) return X360_OK;
    CU(cudaSetDevice(pl->p.device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const bool scores = sum_lap != nullptr;
    if (pl->path == X360_PATH_TILED) return launch_tiled(pl, frames, n_frames, out_views, sum_lap, sum_lap2, st);
#ifdef X360_LAB
    ExtractArgs a;
    a.g = pl->g;
    a.frames = frames;
    a.out = out_views;
    a.sum_lap = (long long *)sum_lap;
    a.sum_lap2 = (long long *)sum_lap2;
    a.R = pl->d_R;
    a.lz_tab = pl->d_lz;
    a.F = n_frames;
    a.V = pl->p.n_views;
    a.n_tiles = pl->p.n_views * pl->g.tiles_x * pl->g.tiles_y;
    a.aligned = ((uintptr_t)frames % 4 == 0) && (pl->g.W % 4 == 0) && ((uintptr_t)out_views % 16 == 0) &&
                (((size_t)pl->g.ow * 3) % 16 == 0);
    a.stage_ok = ((uintptr_t)frames % 16 == 0) && (((size_t)pl->g.W * 3) % 16 == 0) && pl->g.H >= kBoxRows &&
                 pl->g.W * 3 >= kMaxBoxBytes;
    a.rows_max = pl->rows_max; a.px_max = pl->px_max; a.pp = pl->pp; a.raw_pitch = pl->raw_pitch;
    a.tile_modes = pl->d_modes;
    pl->ring_last = 0;
    CU(cudaMemsetAsync(pl->d_modes, 0, sizeof(unsigned) * kRingWords, st));
    if (scores) {
        CU(cudaMemsetAsync(sum_lap, 0, sizeof(int64_t) * (size_t)n_frames * a.V, st));
        CU(cudaMemsetAsync(sum_lap2, 0, sizeof(int64_t) * (size_t)n_frames * a.V, st));
    }
    const int grid = plan_grid(pl, scores);
    if (pl->path == X360_PATH_STAGED) {
        TmapSet tms;
        memset(&tms, 0, sizeof tms);
        if (a.stage_ok && !encode_frame_maps(frames, n_frames, pl->g.H, pl->g.W, &tms)) a.stage_ok = 0;   // every tile goes direct
        pick_staged(scores)<<<grid, kThreads, pl->smem, st>>>(a, tms);
    } else {
        pick_direct(pl->p.interp, scores)<<<grid, kThreads, 0, st>>>(a);
    }
    g_launches.fetch_add(1);
    CU(cudaGetLastError());
    return X360_OK;
#else
    (void)scores;
    return fail(X360_E_UNSUPPORTED, "plan path %d is not compiled into this build", pl->path);
#endif
}

static int unsharp_launch(const uint8_t *d_in, long long n, int h, int w, double strength, uint8_t *d_out, cudaStream_t st);

// Streams, events and the two chunk buffers of the pipelined host path.  Idempotent: a call that failed half-way leaves
// what it created in the plan, the next call creates only what is still missing (x360_plan_destroy frees everything).
static int pipe_setup(x360_plan *pl)
{
    if (pl->pipe_ready && !(pl->sharpen && !pl->d_sharp[1])) return X360_OK;
    const size_t fb = (size_t)pl->g.H * pl->g.W * 3, vb = (size_t)pl->g.oh * pl->g.ow * 3 * pl->p.n_views;
    if (!pl->s_in) CU(cudaStreamCreateWithFlags(&pl->s_in, cudaStreamNonBlocking));
    if (!pl->s_k) CU(cudaStreamCreateWithFlags(&pl->s_k, cudaStreamNonBlocking));
    if (!pl->s_out) CU(cudaStreamCreateWithFlags(&pl->s_out, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        if (!pl->ev_in[i]) CU(cudaEventCreateWithFlags(&pl->ev_in[i], cudaEventDisableTiming));
        if (!pl->ev_k[i]) CU(cudaEventCreateWithFlags(&pl->ev_k[i], cudaEventDisableTiming));
        if (!pl->ev_out[i]) CU(cudaEventCreateWithFlags(&pl->ev_out[i], cudaEventDisableTiming));
        for (int k = 0; k < x360_plan::kViewParts; ++k)
            if (!pl->ev_part[i][k]) CU(cudaEventCreateWithFlags(&pl->ev_part[i][k], cudaEventDisableTiming));
        if (!pl->d_frames[i]) CU(cudaMalloc(&pl->d_frames[i], fb * pl->chunk));
        if (!pl->d_views[i]) CU(cudaMalloc(&pl->d_views[i], vb * pl->chunk));
        if (pl->sharpen && !pl->d_sharp[i]) CU(cudaMalloc(&pl->d_sharp[i], vb * pl->chunk));
        if (!pl->d_sums[i]) CU(cudaMalloc(&pl->d_sums[i], sizeof(long long) * 2 * pl->chunk * pl->p.n_views));
    }
    pl->pipe_ready = true;
    return X360_OK;
}

// n frames from the caller's host memory into a chunk buffer: whole frames, or only the rows the views can read
static cudaError_t upload_frames(const x360_plan *pl, uint8_t *dst, const uint8_t *src, int n, cudaStream_t st)
{
    const size_t pitch = (size_t)pl->g.W * 3, fb = pitch * pl->g.H;
    if (pl->up_rows >= pl->g.H) return cudaMemcpyAsync(dst, src, fb * n, cudaMemcpyHostToDevice, st);
    const size_t off = (size_t)pl->up_row0 * pitch, bytes = (size_t)pl->up_rows * pitch;
    for (int i = 0; i < n; ++i) {
        const cudaError_t e = cudaMemcpyAsync(dst + (size_t)i * fb + off, src + (size_t)i * fb + off, bytes, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// one chunk of the pipeline (see x360_extract_host); any failure is reported to the caller, which drains the streams
static int pipe_chunk(x360_plan *pl, int c, const uint8_t *frames, int f0, int n, uint8_t *out_views, int64_t *sum_lap, int64_t *sum_lap2)
{
    const int b = c & 1, V = pl->p.n_views;
    const size_t fb = (size_t)pl->g.H * pl->g.W * 3, vb = (size_t)pl->g.oh * pl->g.ow * 3 * V;
    const bool scores = sum_lap != nullptr;
    // upload: the input buffer is free once the kernel of chunk c-2 has run
    if (c >= 2) CU(cudaStreamWaitEvent(pl->s_in, pl->ev_k[b], 0));
    CU(upload_frames(pl, pl->d_frames[b], frames + (size_t)f0 * fb, n, pl->s_in));
    CU(cudaEventRecord(pl->ev_in[b], pl->s_in));
    // compute: needs the upload, and the download of chunk c-2 out of this output buffer
    CU(cudaStreamWaitEvent(pl->s_k, pl->ev_in[b], 0));
    if (c >= 2) CU(cudaStreamWaitEvent(pl->s_k, pl->ev_out[b], 0));
    long long *ds = scores ? pl->d_sums[b] : nullptr;
    long long *ds2 = scores ? pl->d_sums[b] + (size_t)pl->chunk * V : nullptr;
    if (n == 1 && pl->view_parts > 1 && !scores && !pl->sharpen && pl->path == X360_PATH_TILED) {
        // a single frame of views that share no map: consecutive view ranges, each downloaded while the next is computed
        const size_t view_bytes = (size_t)pl->g.oh * pl->g.ow * 3;
        for (int k = 0; k < pl->view_parts; ++k) {
            const int v_lo = (int)((long long)V * k / pl->view_parts), v_hi = (int)((long long)V * (k + 1) / pl->view_parts);
            if (int rc = launch_tiled(pl, pl->d_frames[b], 1, pl->d_views[b], nullptr, nullptr, pl->s_k, v_lo, v_hi)) return rc;
            CU(cudaEventRecord(pl->ev_part[b][k], pl->s_k));
            CU(cudaStreamWaitEvent(pl->s_out, pl->ev_part[b][k], 0));
            CU(cudaMemcpyAsync(out_views + (size_t)f0 * vb + (size_t)v_lo * view_bytes, pl->d_views[b] + (size_t)v_lo * view_bytes,
                               (size_t)(v_hi - v_lo) * view_bytes, cudaMemcpyDeviceToHost, pl->s_out));
        }
        CU(cudaEventRecord(pl->ev_k[b], pl->s_k));
        CU(cudaEventRecord(pl->ev_out[b], pl->s_out));
        return X360_OK;
    }
    if (int rc = x360_extract_device(pl, pl->d_frames[b], n, pl->d_views[b], (int64_t *)ds, (int64_t *)ds2, pl->s_k)) return rc;
    const uint8_t *d_result = pl->d_views[b];
    if (pl->sharpen) {                          // processor.py:379-381 on the chunk's views, still on the device
        if (int rc = unsharp_launch(pl->d_views[b], (long long)n * V, pl->g.oh, pl->g.ow, pl->sharpen_strength, pl->d_sharp[b], pl->s_k)) return rc;
        d_result = pl->d_sharp[b];
    }
    CU(cudaEventRecord(pl->ev_k[b], pl->s_k));
    // download
    CU(cudaStreamWaitEvent(pl->s_out, pl->ev_k[b], 0));
    CU(cudaMemcpyAsync(out_views + (size_t)f0 * vb, d_result, vb * n, cudaMemcpyDeviceToHost, pl->s_out));
    if (scores) {
        CU(cudaMemcpyAsync(sum_lap + (size_t)f0 * V, ds, sizeof(long long) * n * V, cudaMemcpyDeviceToHost, pl->s_out));
        CU(cudaMemcpyAsync(sum_lap2 + (size_t)f0 * V, ds2, sizeof(long long) * n * V, cudaMemcpyDeviceToHost, pl->s_out));
    }
    CU(cudaEventRecord(pl->ev_out[b], pl->s_out));
    return X360_OK;
}

int x360_extract_host(x360_plan *pl, const uint8_t *frames, int32_t n_frames, uint8_t *out_views,
                      int64_t *sum_lap, int64_t *sum_lap2)
{
    if (!pl || !frames || !out_views) return fail(X360_E_INVALID, "null argument");
    if (n_frames < 0) return fail(X360_E_INVALID, "negative frame count");
    if ((sum_lap == nullptr) != (sum_lap2 == nullptr)) return fail(X360_E_INVALID, "sum_lap and sum_lap2 must both be given or both be NULL");
    if (n_frames == 0) return X360_OK;
    CU(cudaSetDevice(pl->p.device));
    if (int rc = pipe_setup(pl)) return rc;
    // the sums travel through pinned memory and reach the caller's arrays after the drain
    const size_t n_sums = (size_t)n_frames * pl->p.n_views;
    int64_t *st_lap = nullptr, *st_lap2 = nullptr;
    if (sum_lap) {
        if (pl->h_sum_stage_cap < 2 * n_sums) {
            if (pl->h_sum_stage) cudaFreeHost(pl->h_sum_stage);
            pl->h_sum_stage = nullptr; pl->h_sum_stage_cap = 0;
            CU(cudaHostAlloc(&pl->h_sum_stage, sizeof(long long) * 2 * n_sums, cudaHostAllocDefault));
            pl->h_sum_stage_cap = 2 * n_sums;
        }
        st_lap = (int64_t *)pl->h_sum_stage;
        st_lap2 = st_lap + n_sums;
    }
    // Chunks of pl->chunk frames, with a half chunk first and last where the batch is long enough: the pipeline's fill (the
    // first upload, nothing to overlap it with) and drain (the last download) are then half as long.
    int rc = X360_OK, c = 0;
    const int half = pl->chunk / 2 > 0 ? pl->chunk / 2 : 1;
    const bool ramp = pl->chunk >= 2 && n_frames >= 2 * pl->chunk;
    for (int f0 = 0; f0 < n_frames && rc == X360_OK; ++c) {
        int n = n_frames - f0 < pl->chunk ? n_frames - f0 : pl->chunk;
        if (ramp && (f0 == 0 || n_frames - f0 <= pl->chunk) && n > half) n = half;
        rc = pipe_chunk(pl, c, frames, f0, n, out_views, st_lap, st_lap2);
        f0 += n;
    }
    // Drain all three streams on EVERY exit: copies still in flight target the caller's buffers, which it may free as
    // soon as this function has returned (with an error as much as with success).
    const std::string first_error = rc != X360_OK ? g_err : std::string();
    const cudaError_t e_out = cudaStreamSynchronize(pl->s_out), e_k = cudaStreamSynchronize(pl->s_k), e_in = cudaStreamSynchronize(pl->s_in);
    if (rc != X360_OK) { g_err = first_error; return rc; }
    for (cudaError_t e : {e_out, e_k, e_in})
        if (e != cudaSuccess) return fail(X360_E_CUDA, "pipelined extraction failed: %s", cudaGetErrorString(e));
    if (sum_lap) {
        memcpy(sum_lap, st_lap, sizeof(int64_t) * n_sums);
        memcpy(sum_lap2, st_lap2, sizeof(int64_t) * n_sums);
    }
    return X360_OK;
}

int x360_extract(x360_plan *pl, const uint8_t *frames, int32_t n_frames, uint8_t *out_views,
                 int64_t *sum_lap, int64_t *sum_lap2, void *cuda_stream)
{
    if (!pl || !frames) return fail(X360_E_INVALID, "null argument");
    CU(cudaSetDevice(pl->p.device));
    cudaPointerAttributes at;
    const cudaError_t e = cudaPointerGetAttributes(&at, frames);
    if (e != cudaSuccess) { cudaGetLastError(); }
    const bool on_device = (e == cudaSuccess) && (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged);
    return on_device ? x360_extract_device(pl, frames, n_frames, out_views, sum_lap, sum_lap2, cuda_stream)
                     : x360_extract_host(pl, frames, n_frames, out_views, sum_lap, sum_lap2);
}

int x360_make_maps(x360_plan *pl, int32_t view, float *map_x, float *map_y)
{
    if (!pl || !map_x || !map_y) return fail(X360_E_INVALID, "null argument");
    if (view < 0 || view >= pl->p.n_views) return fail(X360_E_INVALID, "view %d out of range", view);
    CU(cudaSetDevice(pl->p.device));
    const size_t n = (size_t)pl->g.ow * pl->g.oh;
    float *d = (float *)scratch_get(0, pl->p.device, sizeof(float) * 2 * n);
    if (!d) return fail(X360_E_NOMEM, "out of device memory for the maps");
    const dim3 grid((pl->g.ow + 31) / 32, (pl->g.oh + 7) / 8);
    make_maps_kernel<<<grid, 256>>>(pl->g, pl->d_R + 9 * view, d, d + n);
    g_launches.fetch_add(1);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpy(map_x, d, sizeof(float) * n, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(map_y, d + n, sizeof(float) * n, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return fail(X360_E_CUDA, "make_maps failed: %s", cudaGetErrorString(e));
    return X360_OK;
}

int x360_remap_maps(int32_t device, const uint8_t *src, int32_t src_h, int32_t src_w, const float *map_x,
                    const float *map_y, int32_t out_h, int32_t out_w, int32_t interp, uint8_t *dst)
{
    if (!src || !map_x || !map_y || !dst) return fail(X360_E_INVALID, "null argument");
    if (int rc = check_dims(src_w, src_h, out_w, out_h)) return rc;
    if (interp != X360_INTERP_LINEAR && interp != X360_INTERP_LANCZOS4) return fail(X360_E_INVALID, "unknown interp %d", interp);
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev < 1)
        return fail(X360_E_CUDA, "no CUDA device available (libx360 has no CPU fallback)");
    CU(cudaSetDevice(device));
    const Geom g = make_geom(src_w, src_h, out_w, out_h, 90.0);
    const size_t sb = (size_t)src_h * src_w * 3, n = (size_t)out_h * out_w;
    uint8_t *d_src = (uint8_t *)scratch_get(1, device, sb), *d_dst = (uint8_t *)scratch_get(2, device, n * 3);
    float *d_map = (float *)scratch_get(3, device, sizeof(float) * 2 * n);
    int16_t *d_tab = nullptr;
    if (!d_src || !d_dst || !d_map) return fail(X360_E_NOMEM, "out of device memory");
    cudaError_t e = cudaMemcpy(d_src, src, sb, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_map, map_x, sizeof(float) * n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_map + n, map_y, sizeof(float) * n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && interp == X360_INTERP_LANCZOS4) {
        d_tab = (int16_t *)scratch_get(4, device, 1024 * 64 * sizeof(int16_t));
        if (!d_tab) return fail(X360_E_NOMEM, "out of device memory");
        const std::vector<int16_t> t = lanczos4_q15_table();
        e = cudaMemcpy(d_tab, t.data(), t.size() * sizeof(int16_t), cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess) {
        const dim3 grid((out_w + 31) / 32, (out_h + 7) / 8);
        if (interp == X360_INTERP_LANCZOS4)
            remap_maps_kernel<LanczosSampler><<<grid, 256>>>(g, d_src, d_map, d_map + n, d_tab, d_dst);
        else
            remap_maps_kernel<LinearSampler><<<grid, 256>>>(g, d_src, d_map, d_map + n, d_tab, d_dst);
        g_launches.fetch_add(1);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(dst, d_dst, n * 3, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return fail(X360_E_CUDA, "remap_maps failed: %s", cudaGetErrorString(e));
    return X360_OK;
}

int x360_blur_sums(int32_t device, const uint8_t *image, int32_t h, int32_t w, int32_t channels,
                   int64_t *sum_lap, int64_t *sum_lap2)
{
    if (!image || !sum_lap || !sum_lap2) return fail(X360_E_INVALID, "null argument");
    if (h < 1 || w < 1 || (channels != 1 && channels != 3)) return fail(X360_E_INVALID, "bad image %dx%dx%d", h, w, channels);
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev < 1)
        return fail(X360_E_CUDA, "no CUDA device available (libx360 has no CPU fallback)");
    CU(cudaSetDevice(device));
    const size_t n = (size_t)h * w;
    uint8_t *d_img = (uint8_t *)scratch_get(1, device, n * channels), *d_gray = channels == 3 ? (uint8_t *)scratch_get(2, device, n) : nullptr;
    long long *d_s = (long long *)scratch_get(5, device, sizeof(long long) * 2);
    if (!d_img || !d_s || (channels == 3 && !d_gray)) return fail(X360_E_NOMEM, "out of device memory");
    cudaError_t e = cudaMemset(d_s, 0, sizeof(long long) * 2);
    if (e == cudaSuccess) e = cudaMemcpy(d_img, image, n * channels, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && channels == 3) {
        gray_kernel<<<(unsigned)((n + 255) / 256), 256>>>(d_img, d_gray, n);
        g_launches.fetch_add(1);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) {
        size_t blocks = (n + 255) / 256;
        if (blocks > 148 * 8) blocks = 148 * 8;
        laplacian_sums_kernel<<<(unsigned)blocks, 256>>>(channels == 3 ? d_gray : d_img, h, w, d_s, d_s + 1);
        g_launches.fetch_add(1);
        e = cudaGetLastError();
    }
    long long hs[2] = {0, 0};
    if (e == cudaSuccess) e = cudaMemcpy(hs, d_s, sizeof hs, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return fail(X360_E_CUDA, "blur_sums failed: %s", cudaGetErrorString(e));
    *sum_lap = hs[0];
    *sum_lap2 = hs[1];
    return X360_OK;
}

// ---- unsharp mask: processor.py:379-381 --------------------------------------------------------------
static int unsharp_launch(const uint8_t *d_in, long long n, int h, int w, double strength, uint8_t *d_out, cudaStream_t st)
{
    const float alpha = (float)(1.0 + strength), beta = (float)(-strength);       // addWeighted casts its scalars to float
    const size_t pitch = (size_t)w * 3;
    bool tiled = (pitch % 16 == 0) && pitch >= (size_t)kUsBoxW && h >= kUsRows && ((uintptr_t)d_in % 16 == 0) && ((uintptr_t)d_out % 16 == 0);
    if (lab_int("X360_UNSHARP_GENERIC", 0)) tiled = false;
    UnsharpMaps tms;
    memset(&tms, 0, sizeof tms);
    if (tiled) {
        encode_tiled_fn enc = tensor_map_encoder();
        const cuuint32_t estr[3] = {1, 1, 1};
        const cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)h, (cuuint64_t)n};
        const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * h};
        const cuuint32_t box_in[3] = {(cuuint32_t)kUsBoxW, (cuuint32_t)kUsRows, 1}, box_out[3] = {(cuuint32_t)(kUsTW * 3), (cuuint32_t)kUsTH, 1};
        tiled = enc &&
                enc(&tms.in, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void *)d_in, dims, strides, box_in, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS &&
                enc(&tms.out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void *)d_out, dims, strides, box_out, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    }
    if (tiled) {
        static bool attr_set = false;
        if (!attr_set) {
            CU(cudaFuncSetAttribute(unsharp_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kUsSmem));
            attr_set = true;
        }
        for (long long n0 = 0; n0 < n; n0 += 65535) {
            const dim3 grid((unsigned)((w + kUsTW - 1) / kUsTW), (unsigned)((h + kUsTH - 1) / kUsTH), (unsigned)std::min<long long>(65535, n - n0));
            unsharp_tiled<<<grid, kUsThreads, kUsSmem, st>>>(tms, h, w, (int)n0, alpha, beta);
            g_launches.fetch_add(1);
        }
    } else {
        const size_t img = (size_t)h * pitch;
        for (long long n0 = 0; n0 < n; n0 += 65535) {
            const dim3 grid((unsigned)((img + 255) / 256), (unsigned)std::min<long long>(65535, n - n0));
            unsharp_generic<<<grid, 256, 0, st>>>(d_in + (size_t)n0 * img, d_out + (size_t)n0 * img, h, w, alpha, beta);
            g_launches.fetch_add(1);
        }
    }
    CU(cudaGetLastError());
    return X360_OK;
}

static int unsharp_check(int32_t device, const void *in, int64_t n, int32_t h, int32_t w, double strength, const void *out)
{
    if (!in || !out) return fail(X360_E_INVALID, "null argument");
    if (in == out) return fail(X360_E_INVALID, "unsharp cannot run in place (every output needs a 13 x 13 input window)");
    if (n < 0 || h < 1 || w < 1 || h > 65535 || w > 65535) return fail(X360_E_INVALID, "bad image batch %lld x %d x %d", (long long)n, h, w);
    if (!(strength == strength) || std::fabs(strength) > 1e6) return fail(X360_E_INVALID, "bad strength %g", strength);
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev < 1)
        return fail(X360_E_CUDA, "no CUDA device available (libx360 has no CPU fallback)");
    if (device < 0 || device >= n_dev) return fail(X360_E_INVALID, "device %d out of range [0, %d)", device, n_dev);
    CU(cudaSetDevice(device));
    return X360_OK;
}

int x360_unsharp_device(int32_t device, const uint8_t *images, int64_t n_images, int32_t h, int32_t w, double strength,
                        uint8_t *out, void *cuda_stream)
{
    if (int rc = unsharp_check(device, images, n_images, h, w, strength, out)) return rc;
    if (n_images == 0) return X360_OK;
    return unsharp_launch(images, n_images, h, w, strength, out, (cudaStream_t)cuda_stream);
}

int x360_unsharp_host(int32_t device, const uint8_t *images, int64_t n_images, int32_t h, int32_t w, double strength, uint8_t *out)
{
    if (int rc = unsharp_check(device, images, n_images, h, w, strength, out)) return rc;
    if (n_images == 0) return X360_OK;
    const size_t img = (size_t)h * w * 3;
    long long chunk = (long long)std::max<size_t>(1, ((size_t)512 << 20) / img);       // <= 512 MB per device buffer
    if (chunk > n_images) chunk = n_images;
    uint8_t *d_in = nullptr, *d_out = nullptr;
    cudaError_t e = cudaMalloc(&d_in, img * chunk);
    if (e == cudaSuccess) e = cudaMalloc(&d_out, img * chunk);
    int rc = X360_OK;
    for (long long n0 = 0; e == cudaSuccess && rc == X360_OK && n0 < n_images; n0 += chunk) {
        const long long m = std::min(chunk, (long long)n_images - n0);
        e = cudaMemcpy(d_in, images + (size_t)n0 * img, img * m, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) rc = unsharp_launch(d_in, m, h, w, strength, d_out, nullptr);
        if (e == cudaSuccess && rc == X360_OK) e = cudaMemcpy(out + (size_t)n0 * img, d_out, img * m, cudaMemcpyDeviceToHost);
    }
    cudaFree(d_in);
    cudaFree(d_out);
    if (e != cudaSuccess) return fail(X360_E_CUDA, "unsharp failed: %s", cudaGetErrorString(e));
    return rc;
}

int x360_views_to_nchw(int32_t device, const uint8_t *views, int64_t n_views, int32_t h, int32_t w, const int32_t *select,
                       int32_t n_select, int32_t to_rgb, float *out, void *cuda_stream)
{
    if (!views || !out) return fail(X360_E_INVALID, "null argument");
    if (h < 1 || w < 1 || n_views < 0 || n_select < 0) return fail(X360_E_INVALID, "bad shape %lld x %d x %d, %d selected", (long long)n_views, h, w, n_select);
    const long long total = select ? n_select : n_views;
    for (int i = 0; select && i < n_select; ++i)
        if (select[i] < 0 || select[i] >= n_views) return fail(X360_E_INVALID, "select[%d] = %d outside [0, %lld)", i, select[i], (long long)n_views);
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev < 1)
        return fail(X360_E_CUDA, "no CUDA device available (libx360 has no CPU fallback)");
    CU(cudaSetDevice(device));
    if (total == 0) return X360_OK;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const size_t hw = (size_t)h * w, view_bytes = hw * 3;
    const bool fast = hw % 4 == 0 && (uintptr_t)views % 4 == 0 && (uintptr_t)out % 16 == 0;
    for (long long n0 = 0; n0 < total; n0 += 64) {
        NchwSel sel;
        const int m = (int)std::min<long long>(64, total - n0);
        for (int i = 0; i < m; ++i) sel.idx[i] = select ? select[n0 + i] : (int)(n0 + i);
        float *o = out + (size_t)n0 * 3 * hw;
        if (fast) {
            const int groups = (int)(hw / 4);
            views_to_nchw_kernel<<<dim3((unsigned)((groups + 255) / 256), (unsigned)m), 256, 0, st>>>(views, view_bytes, groups, sel, o, to_rgb);
        } else {
            views_to_nchw_generic<<<dim3((unsigned)((hw + 255) / 256), (unsigned)m), 256, 0, st>>>(views, view_bytes, hw, sel, o, to_rgb);
        }
        g_launches.fetch_add(1);
        CU(cudaGetLastError());
    }
    return X360_OK;
}

// ---- baseline JPEG encode of view batches on the device (x360_jpeg.cuh) ---------------------------------------------
struct x360_jpeg {
    int device = 0, h = 0, w = 0, max_images = 0, quality = 95;
    JpegGeom g;
    JpegTables tab;
    std::vector<uint8_t> header;
    int max_segs = 0;
    unsigned long long cap_words = 0;
    // device scratch
    int16_t *d_coef = nullptr;
    uint32_t *d_bits = nullptr, *d_stream = nullptr, *d_segff = nullptr;
    unsigned long long *d_view_bits = nullptr, *d_word_base = nullptr, *d_file_bytes = nullptr;
    uint8_t *d_header = nullptr;
    int *d_overflow = nullptr;
    // host-path staging
    uint8_t *d_images = nullptr, *d_out = nullptr;
    long long *d_offsets = nullptr;
};

static int jpeg_launch(x360_jpeg *e, const uint8_t *images, int n, uint8_t *out, long long out_capacity, long long *offsets, cudaStream_t st)
{
    const JpegGeom &g = e->g;
    CU(cudaMemsetAsync(e->d_overflow, 0, sizeof(int), st));
    jpeg_blocks_kernel<<<dim3((unsigned)((g.mw + kJbMcus - 1) / kJbMcus), (unsigned)g.mh, (unsigned)n), kJbThreads, 0, st>>>(images, g, e->tab, e->d_coef);
    if (g.w % 16 || g.h % 16) jpeg_dummy_dc_kernel<<<dim3((unsigned)((g.n_mcu + 127) / 128), (unsigned)n), 128, 0, st>>>(g, e->d_coef);
    jpeg_bits_kernel<<<dim3((unsigned)((g.n_blocks + 127) / 128), (unsigned)n), 128, 0, st>>>(e->d_coef, g, e->d_bits);
    jpeg_scan_kernel<<<(unsigned)n, 1024, 0, st>>>(e->d_bits, g.n_blocks, e->d_view_bits);
    jpeg_bases_kernel<<<1, 32, 0, st>>>(e->d_view_bits, n, e->d_word_base, e->cap_words, e->d_overflow);
    jpeg_zero_kernel<<<148 * 4, 256, 0, st>>>(e->d_stream, e->d_word_base, n, e->d_overflow);
    jpeg_emit_kernel<<<dim3((unsigned)((g.n_blocks + 127) / 128), (unsigned)n), 128, 0, st>>>(e->d_coef, g, e->d_bits, e->d_word_base, e->d_stream, e->d_overflow);
    jpeg_ffcount_kernel<<<dim3((unsigned)e->max_segs, (unsigned)n), 256, 0, st>>>(e->d_stream, e->d_word_base, e->d_view_bits, e->d_segff, e->max_segs, e->d_overflow);
    jpeg_ffscan_kernel<<<(unsigned)n, 256, 0, st>>>(e->d_view_bits, e->d_segff, e->max_segs, (int)e->header.size(), e->d_file_bytes, e->d_overflow);
    jpeg_offsets_kernel<<<1, 32, 0, st>>>(e->d_file_bytes, n, offsets, out_capacity, e->d_overflow);
    jpeg_pack_kernel<<<dim3((unsigned)e->max_segs, (unsigned)n), 256, 0, st>>>(e->d_stream, e->d_word_base, e->d_view_bits, e->d_segff, e->max_segs,
                                                                             e->d_header, (int)e->header.size(), offsets, out, e->d_overflow);
    g_launches.fetch_add(10);
    CU(cudaGetLastError());
    return X360_OK;
}

int x360_jpeg_create(int32_t device, int32_t h, int32_t w, int32_t max_images, int32_t quality, x360_jpeg **out)
{
    if (!out) return fail(X360_E_INVALID, "null argument");
    *out = nullptr;
    if (h < 1 || w < 1 || h > 65535 || w > 65535) return fail(X360_E_INVALID, "bad image size %d x %d (JPEG dimensions are 16-bit)", h, w);
    if (max_images < 1) return fail(X360_E_INVALID, "max_images must be positive");
    if (quality < 1 || quality > 100) return fail(X360_E_INVALID, "quality %d outside [1, 100]", quality);
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev < 1)
        return fail(X360_E_CUDA, "no CUDA device available (libx360 has no CPU fallback)");
    if (device < 0 || device >= n_dev) return fail(X360_E_INVALID, "device %d out of range [0, %d)", device, n_dev);
    CU(cudaSetDevice(device));
    x360_jpeg *e = new (std::nothrow) x360_jpeg();
    if (!e) return fail(X360_E_NOMEM, "out of host memory");
    e->device = device; e->h = h; e->w = w; e->max_images = max_images; e->quality = quality;
    JpegGeom &g = e->g;
    g.h = h; g.w = w;
    g.mw = (w + 15) / 16; g.mh = (h + 15) / 16;
    g.bw_y = (w + 7) / 8; g.bh_y = (h + 7) / 8;
    g.crows = (h + 1) / 2;
    g.n_mcu = g.mw * g.mh; g.n_blocks = 6 * g.n_mcu;
    jpeg_make_tables(quality, &e->tab);
    e->header = jpeg_make_header(h, w, quality);
    // Stream capacity: the raw size of the batch (3 bytes per pixel) + slack.  A quality-95 file of camera-like content is
    // 5-15 % of that and of uniform noise 40 %; a batch that needs more is refused (X360_E_NOMEM), not truncated.
    const size_t raw = (size_t)h * w * 3;
    e->max_segs = (int)((raw + raw / 2 + 4096 + kJpegSeg - 1) / kJpegSeg);          // per view: 1.5 x its raw size
    e->cap_words = (((unsigned long long)raw * (unsigned long long)max_images) / 4ull + 8ull * (unsigned long long)max_images + 1024ull + 3ull) & ~3ull;
    JpegHuff hf;
    jpeg_make_huff(&hf);
    cudaError_t c = cudaMemcpyToSymbol(c_jpeg_huff, &hf, sizeof hf);
    const size_t nb = (size_t)g.n_blocks * max_images;
    if (c == cudaSuccess) c = cudaMalloc(&e->d_coef, nb * 64 * sizeof(int16_t));
    if (c == cudaSuccess) c = cudaMalloc(&e->d_bits, nb * sizeof(uint32_t));
    if (c == cudaSuccess) c = cudaMalloc(&e->d_stream, (size_t)e->cap_words * sizeof(uint32_t));
    if (c == cudaSuccess) c = cudaMalloc(&e->d_segff, (size_t)e->max_segs * max_images * sizeof(uint32_t));
    if (c == cudaSuccess) c = cudaMalloc(&e->d_view_bits, sizeof(unsigned long long) * (size_t)max_images);
    if (c == cudaSuccess) c = cudaMalloc(&e->d_word_base, sizeof(unsigned long long) * ((size_t)max_images + 1));
    if (c == cudaSuccess) c = cudaMalloc(&e->d_file_bytes, sizeof(unsigned long long) * (size_t)max_images);
    if (c == cudaSuccess) c = cudaMalloc(&e->d_header, e->header.size());
    if (c == cudaSuccess) c = cudaMalloc(&e->d_overflow, sizeof(int));
    if (c == cudaSuccess) c = cudaMemcpy(e->d_header, e->header.data(), e->header.size(), cudaMemcpyHostToDevice);
    if (c != cudaSuccess) {
        x360_jpeg_destroy(e);
        return fail(X360_E_CUDA, "jpeg encoder allocation failed: %s", cudaGetErrorString(c));
    }
    *out = e;
    return X360_OK;
}

int x360_jpeg_destroy(x360_jpeg *e)
{
    if (!e) return X360_OK;
    cudaSetDevice(e->device);
    cudaFree(e->d_coef); cudaFree(e->d_bits); cudaFree(e->d_stream); cudaFree(e->d_segff); cudaFree(e->d_view_bits);
    cudaFree(e->d_word_base); cudaFree(e->d_file_bytes); cudaFree(e->d_header); cudaFree(e->d_overflow);
    cudaFree(e->d_images); cudaFree(e->d_out); cudaFree(e->d_offsets);
    delete e;
    return X360_OK;
}

int64_t x360_jpeg_max_bytes(const x360_jpeg *e, int32_t n_images)
{
    if (!e || n_images < 0) return 0;
    return (int64_t)n_images * ((int64_t)e->h * e->w * 3 + (int64_t)e->header.size() + 64);
}

int x360_jpeg_encode_device(x360_jpeg *e, const uint8_t *images, int32_t n_images, uint8_t *out, int64_t out_capacity, int64_t *offsets,
                            void *cuda_stream)
{
    if (!e || !images || !out || !offsets) return fail(X360_E_INVALID, "null argument");
    if (n_images < 0 || n_images > e->max_images) return fail(X360_E_INVALID, "n_images %d outside [0, %d]", n_images, e->max_images);
    if (out_capacity < 0) return fail(X360_E_INVALID, "negative capacity");
    CU(cudaSetDevice(e->device));
    if (n_images == 0) { CU(cudaMemsetAsync(offsets, 0, sizeof(int64_t), (cudaStream_t)cuda_stream)); return X360_OK; }
    return jpeg_launch(e, images, n_images, out, (long long)out_capacity, (long long *)offsets, (cudaStream_t)cuda_stream);
}

int x360_jpeg_status(x360_jpeg *e, void *cuda_stream)
{
    if (!e) return fail(X360_E_INVALID, "null argument");
    CU(cudaSetDevice(e->device));
    int ov = 0;
    CU(cudaStreamSynchronize((cudaStream_t)cuda_stream));
    CU(cudaMemcpy(&ov, e->d_overflow, sizeof ov, cudaMemcpyDeviceToHost));
    if (ov) return fail(X360_E_NOMEM, ov == 2 ? "jpeg output larger than the capacity given" : "jpeg stream larger than the encoder's scratch (3 bytes per pixel of the batch)");
    return X360_OK;
}

int x360_jpeg_encode_host(x360_jpeg *e, const uint8_t *images, int64_t n_images, uint8_t *out, int64_t out_capacity, int64_t *offsets)
{
    if (!e || !images || !out || !offsets) return fail(X360_E_INVALID, "null argument");
    if (n_images < 0 || out_capacity < 0) return fail(X360_E_INVALID, "negative count or capacity");
    CU(cudaSetDevice(e->device));
    const size_t img = (size_t)e->h * e->w * 3;
    const long long dev_cap = (long long)x360_jpeg_max_bytes(e, e->max_images);
    if (!e->d_images) CU(cudaMalloc(&e->d_images, img * e->max_images));
    if (!e->d_out) CU(cudaMalloc(&e->d_out, (size_t)dev_cap));
    if (!e->d_offsets) CU(cudaMalloc(&e->d_offsets, sizeof(long long) * ((size_t)e->max_images + 1)));
    std::vector<long long> off((size_t)e->max_images + 1);
    long long total = 0;
    offsets[0] = 0;
    for (int64_t n0 = 0; n0 < n_images; n0 += e->max_images) {
        const int m = (int)std::min<int64_t>(e->max_images, n_images - n0);
        CU(cudaMemcpy(e->d_images, images + (size_t)n0 * img, img * m, cudaMemcpyHostToDevice));
        if (int rc = jpeg_launch(e, e->d_images, m, e->d_out, dev_cap, e->d_offsets, nullptr)) return rc;
        if (int rc = x360_jpeg_status(e, nullptr)) return rc;
        CU(cudaMemcpy(off.data(), e->d_offsets, sizeof(long long) * ((size_t)m + 1), cudaMemcpyDeviceToHost));
        if (total + off[(size_t)m] > out_capacity)
            return fail(X360_E_NOMEM, "jpeg output needs more than the %lld bytes given (at image %lld)", (long long)out_capacity, (long long)n0);
        CU(cudaMemcpy(out + total, e->d_out, (size_t)off[(size_t)m], cudaMemcpyDeviceToHost));
        for (int i = 0; i < m; ++i) offsets[n0 + i + 1] = total + off[(size_t)i + 1];
        total += off[(size_t)m];
    }
    return X360_OK;
}

int x360_jpeg_coefficients_host(x360_jpeg *e, const uint8_t *images, int32_t n_images, int16_t *coef)
{
    if (!e || !images || !coef) return fail(X360_E_INVALID, "null argument");
    if (n_images < 0 || n_images > e->max_images) return fail(X360_E_INVALID, "n_images %d outside [0, %d]", n_images, e->max_images);
    if (n_images == 0) return X360_OK;
    CU(cudaSetDevice(e->device));
    const size_t img = (size_t)e->h * e->w * 3;
    if (!e->d_images) CU(cudaMalloc(&e->d_images, img * e->max_images));
    CU(cudaMemcpy(e->d_images, images, img * n_images, cudaMemcpyHostToDevice));
    const JpegGeom &g = e->g;
    jpeg_blocks_kernel<<<dim3((unsigned)((g.mw + kJbMcus - 1) / kJbMcus), (unsigned)g.mh, (unsigned)n_images), kJbThreads>>>(e->d_images, g, e->tab, e->d_coef);
    jpeg_dummy_dc_kernel<<<dim3((unsigned)((g.n_mcu + 127) / 128), (unsigned)n_images), 128>>>(g, e->d_coef);
    g_launches.fetch_add(2);
    CU(cudaGetLastError());
    CU(cudaMemcpy(coef, e->d_coef, (size_t)g.n_blocks * n_images * 64 * sizeof(int16_t), cudaMemcpyDeviceToHost));
    return X360_OK;
}

int x360_jpeg_header(int32_t h, int32_t w, int32_t quality, uint8_t *out, int32_t capacity)
{
    if (h < 1 || w < 1 || h > 65535 || w > 65535 || quality < 1 || quality > 100) return fail(X360_E_INVALID, "bad jpeg header arguments");
    const std::vector<uint8_t> hd = jpeg_make_header(h, w, quality);
    if (out && capacity >= (int32_t)hd.size()) memcpy(out, hd.data(), hd.size());
    return (int)hd.size();
}

// The pipelined host path with the save step's encode on the device: frames in (HOST), JPEG files out (HOST).
int x360_extract_host_jpeg(x360_plan *pl, const uint8_t *frames, int32_t n_frames, int32_t quality, uint8_t *out, int64_t out_capacity,
                           int64_t *offsets, int64_t *sum_lap, int64_t *sum_lap2)
{
    if (!pl || !frames || !out || !offsets) return fail(X360_E_INVALID, "null argument");
    if (n_frames < 0 || out_capacity < 0) return fail(X360_E_INVALID, "negative frame count or capacity");
    if ((sum_lap == nullptr) != (sum_lap2 == nullptr)) return fail(X360_E_INVALID, "sum_lap and sum_lap2 must both be given or both be NULL");
    offsets[0] = 0;
    if (n_frames == 0) return X360_OK;
    CU(cudaSetDevice(pl->p.device));
    if (int rc = pipe_setup(pl)) return rc;
    const int V = pl->p.n_views, per = pl->chunk * V;
    if (pl->jpeg && pl->jpeg->quality != quality) { x360_jpeg_destroy(pl->jpeg); pl->jpeg = nullptr; }
    if (!pl->jpeg)
        if (int rc = x360_jpeg_create(pl->p.device, pl->g.oh, pl->g.ow, per, quality, &pl->jpeg)) return rc;
    pl->jpeg_cap = (long long)x360_jpeg_max_bytes(pl->jpeg, per);
    if (!pl->s_off) CU(cudaStreamCreateWithFlags(&pl->s_off, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        if (!pl->d_jpeg[i]) CU(cudaMalloc(&pl->d_jpeg[i], (size_t)pl->jpeg_cap));
        if (!pl->d_joff[i]) CU(cudaMalloc(&pl->d_joff[i], sizeof(long long) * ((size_t)per + 2)));
        if (!pl->h_joff[i]) CU(cudaHostAlloc(&pl->h_joff[i], sizeof(long long) * ((size_t)per + 2), cudaHostAllocDefault));
        if (!pl->ev_off[i]) CU(cudaEventCreateWithFlags(&pl->ev_off[i], cudaEventDisableTiming));
    }
    const size_t fb = (size_t)pl->g.H * pl->g.W * 3;
    const bool scores = sum_lap != nullptr;
    const size_t n_sums = (size_t)n_frames * V;
    int64_t *st_lap = nullptr, *st_lap2 = nullptr;
    if (scores) {
        if (pl->h_sum_stage_cap < 2 * n_sums) {
            if (pl->h_sum_stage) cudaFreeHost(pl->h_sum_stage);
            pl->h_sum_stage = nullptr; pl->h_sum_stage_cap = 0;
            CU(cudaHostAlloc(&pl->h_sum_stage, sizeof(long long) * 2 * n_sums, cudaHostAllocDefault));
            pl->h_sum_stage_cap = 2 * n_sums;
        }
        st_lap = (int64_t *)pl->h_sum_stage;
        st_lap2 = st_lap + n_sums;
    }
    long long base = 0;
    int rc = X360_OK;
    // finish chunk c: its offsets are on the host -> the size of its files is known -> download exactly those bytes
    auto finish = [&](int c) -> int {
        const int b = c & 1, f0 = c * pl->chunk, n = std::min(pl->chunk, n_frames - f0), m = n * V;
        CU(cudaEventSynchronize(pl->ev_off[b]));
        const long long *off = pl->h_joff[b];
        if (off[m + 1] != 0) return fail(X360_E_NOMEM, "jpeg scratch too small for chunk %d (code %lld)", c, off[m + 1]);
        if (base + off[m] > out_capacity) return fail(X360_E_NOMEM, "jpeg output needs more than the %lld bytes given (chunk %d)", (long long)out_capacity, c);
        CU(cudaMemcpyAsync(out + base, pl->d_jpeg[b], (size_t)off[m], cudaMemcpyDeviceToHost, pl->s_out));
        CU(cudaEventRecord(pl->ev_out[b], pl->s_out));
        for (int i = 0; i < m; ++i) offsets[(size_t)f0 * V + i + 1] = base + off[i + 1];
        base += off[m];
        return X360_OK;
    };
    const int n_chunks = (n_frames + pl->chunk - 1) / pl->chunk;
    for (int c = 0; c < n_chunks && rc == X360_OK; ++c) {
        const int b = c & 1, f0 = c * pl->chunk, n = std::min(pl->chunk, n_frames - f0), m = n * V;
        auto enqueue = [&]() -> int {
            if (c >= 2) CU(cudaStreamWaitEvent(pl->s_in, pl->ev_k[b], 0));
            CU(upload_frames(pl, pl->d_frames[b], frames + (size_t)f0 * fb, n, pl->s_in));
            CU(cudaEventRecord(pl->ev_in[b], pl->s_in));
            CU(cudaStreamWaitEvent(pl->s_k, pl->ev_in[b], 0));
            if (c >= 2) CU(cudaStreamWaitEvent(pl->s_k, pl->ev_out[b], 0));       // the files of chunk c-2 have left d_jpeg[b]
            long long *ds = scores ? pl->d_sums[b] : nullptr;
            long long *ds2 = scores ? pl->d_sums[b] + (size_t)pl->chunk * V : nullptr;
            if (int r = x360_extract_device(pl, pl->d_frames[b], n, pl->d_views[b], (int64_t *)ds, (int64_t *)ds2, pl->s_k)) return r;
            const uint8_t *d_result = pl->d_views[b];
            if (pl->sharpen) {
                if (int r = unsharp_launch(pl->d_views[b], (long long)m, pl->g.oh, pl->g.ow, pl->sharpen_strength, pl->d_sharp[b], pl->s_k)) return r;
                d_result = pl->d_sharp[b];
            }
            if (int r = jpeg_launch(pl->jpeg, d_result, m, pl->d_jpeg[b], pl->jpeg_cap, pl->d_joff[b], pl->s_k)) return r;
            CU(cudaMemsetAsync(pl->d_joff[b] + m + 1, 0, sizeof(long long), pl->s_k));
            CU(cudaMemcpyAsync(pl->d_joff[b] + m + 1, pl->jpeg->d_overflow, sizeof(int), cudaMemcpyDeviceToDevice, pl->s_k));
            CU(cudaEventRecord(pl->ev_k[b], pl->s_k));
            CU(cudaStreamWaitEvent(pl->s_off, pl->ev_k[b], 0));
            CU(cudaMemcpyAsync(pl->h_joff[b], pl->d_joff[b], sizeof(long long) * ((size_t)m + 2), cudaMemcpyDeviceToHost, pl->s_off));
            if (scores) {                                                       // (pinned staging: see x360_extract_host)
                CU(cudaMemcpyAsync(st_lap + (size_t)f0 * V, ds, sizeof(long long) * m, cudaMemcpyDeviceToHost, pl->s_off));
                CU(cudaMemcpyAsync(st_lap2 + (size_t)f0 * V, ds2, sizeof(long long) * m, cudaMemcpyDeviceToHost, pl->s_off));
            }
            CU(cudaEventRecord(pl->ev_off[b], pl->s_off));
            return X360_OK;
        };
        rc = enqueue();
        if (rc == X360_OK && c >= 1) rc = finish(c - 1);
    }
    if (rc == X360_OK) rc = finish(n_chunks - 1);
    const std::string first_error = rc != X360_OK ? g_err : std::string();
    const cudaError_t e_out = cudaStreamSynchronize(pl->s_out), e_off = cudaStreamSynchronize(pl->s_off), e_k = cudaStreamSynchronize(pl->s_k),
                      e_in = cudaStreamSynchronize(pl->s_in);
    if (rc != X360_OK) { g_err = first_error; return rc; }
    for (cudaError_t e : {e_out, e_off, e_k, e_in})
        if (e != cudaSuccess) return fail(X360_E_CUDA, "pipelined extraction failed: %s", cudaGetErrorString(e));
    if (scores) {
        memcpy(sum_lap, st_lap, sizeof(int64_t) * n_sums);
        memcpy(sum_lap2, st_lap2, sizeof(int64_t) * n_sums);
    }
    return X360_OK;
}

// ---- the blur accept / reject machine (processor.py:341-376) and the filtered download --------------------------------
void x360_blur_state_init(x360_blur_state *st, int32_t enabled, int32_t smart, double threshold)
{
    if (!st) return;
    memset(st, 0, sizeof *st);
    st->enabled = enabled != 0;
    st->smart = smart != 0;
    st->threshold = threshold;
}

int x360_blur_accept(x360_blur_state *st, double score)
{
    if (!st || !st->enabled) return 1;                         // processor.py:341 — no score, view always kept
    bool blurry = false;
    if (st->smart) {
        if (score < st->threshold) blurry = true;              // 1. floor
        else if (st->hist_len > 0) {                           // 2. adaptive check against the rolling mean (sum in deque order)
            double sum = 0.0;
            for (int i = 0; i < st->hist_len; ++i) sum += st->history[(st->hist_head + i) % 10];
            if (score < (sum / (double)st->hist_len) * 0.6) blurry = true;
        }
        if (blurry) {                                          // 3. safety override
            if (++st->consecutive_skips > 5) {
                blurry = false;
                st->consecutive_skips = 0;
                st->forced++;
            }
        }
        if (!blurry) {                                         // 4. accepted scores feed the history (deque(maxlen=10))
            st->consecutive_skips = 0;
            if (st->hist_len < 10) st->history[(st->hist_head + st->hist_len++) % 10] = score;
            else { st->history[st->hist_head] = score; st->hist_head = (st->hist_head + 1) % 10; }
        }
    } else {
        blurry = score < st->threshold;                        // standard mode, processor.py:369-371
    }
    if (blurry) st->skipped++;
    return blurry ? 0 : 1;
}

int x360_blur_replay(x360_blur_state *st, const double *scores, int64_t n, uint8_t *keep)
{
    if (!st || !keep || (n > 0 && !scores && st->enabled)) return fail(X360_E_INVALID, "null argument");
    if (n < 0) return fail(X360_E_INVALID, "negative count");
    for (int64_t i = 0; i < n; ++i) keep[i] = (uint8_t)x360_blur_accept(st, st->enabled ? scores[i] : 0.0);
    return X360_OK;
}

int x360_extract_host_filtered(x360_plan *pl, const uint8_t *frames, int32_t n_frames, x360_blur_state *state, uint8_t *out_views,
                               int64_t capacity_views, uint8_t *keep, double *scores, int64_t *n_kept)
{
    if (!pl || !frames || !out_views || !state || !keep || !n_kept) return fail(X360_E_INVALID, "null argument");
    if (n_frames < 0 || capacity_views < 0) return fail(X360_E_INVALID, "negative frame count or capacity");
    *n_kept = 0;
    if (n_frames == 0) return X360_OK;
    CU(cudaSetDevice(pl->p.device));
    if (int rc = pipe_setup(pl)) return rc;
    const int V = pl->p.n_views, per = pl->chunk * V;
    const bool scoring = state->enabled != 0;
    if (!pl->s_off) CU(cudaStreamCreateWithFlags(&pl->s_off, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        if (!pl->h_sums[i]) CU(cudaHostAlloc(&pl->h_sums[i], sizeof(long long) * 2 * (size_t)per, cudaHostAllocDefault));
        if (!pl->ev_off[i]) CU(cudaEventCreateWithFlags(&pl->ev_off[i], cudaEventDisableTiming));
    }
    const size_t fb = (size_t)pl->g.H * pl->g.W * 3, v1 = (size_t)pl->g.oh * pl->g.ow * 3;
    const double N = (double)pl->g.oh * (double)pl->g.ow;
    long long kept = 0;
    int rc = X360_OK;
    const uint8_t *d_result[2] = {nullptr, nullptr};
    // finish chunk c: its sums are on the host -> replay the machine in (frame, view) order -> one download per kept view
    auto finish = [&](int c) -> int {
        const int b = c & 1, f0 = c * pl->chunk, n = std::min(pl->chunk, n_frames - f0), m = n * V;
        if (scoring) CU(cudaEventSynchronize(pl->ev_off[b]));
        CU(cudaStreamWaitEvent(pl->s_out, pl->ev_k[b], 0));              // the chunk's kernels (a no-op after the host wait above)
        const long long *hs = pl->h_sums[b], *hs2 = pl->h_sums[b] + per;
        for (int i = 0; i < m; ++i) {
            double sc = 0.0;
            if (scoring) {                                     // ndarray.var: s2/N - (s/N)^2 in fp64 (image_utils.py:27)
                const double mean = (double)hs[i] / N;
                sc = (double)hs2[i] / N - mean * mean;
            }
            if (scores) scores[(size_t)f0 * V + i] = sc;
            const int k = x360_blur_accept(state, sc);
            keep[(size_t)f0 * V + i] = (uint8_t)k;
            if (!k) continue;
            if (kept >= capacity_views) return fail(X360_E_NOMEM, "more kept views than the %lld the output holds", (long long)capacity_views);
            CU(cudaMemcpyAsync(out_views + (size_t)kept * v1, d_result[b] + (size_t)i * v1, v1, cudaMemcpyDeviceToHost, pl->s_out));
            ++kept;
        }
        CU(cudaEventRecord(pl->ev_out[b], pl->s_out));
        return X360_OK;
    };
    const int n_chunks = (n_frames + pl->chunk - 1) / pl->chunk;
    for (int c = 0; c < n_chunks && rc == X360_OK; ++c) {
        const int b = c & 1, f0 = c * pl->chunk, n = std::min(pl->chunk, n_frames - f0), m = n * V;
        auto enqueue = [&]() -> int {
            if (c >= 2) CU(cudaStreamWaitEvent(pl->s_in, pl->ev_k[b], 0));
            CU(upload_frames(pl, pl->d_frames[b], frames + (size_t)f0 * fb, n, pl->s_in));
            CU(cudaEventRecord(pl->ev_in[b], pl->s_in));
            CU(cudaStreamWaitEvent(pl->s_k, pl->ev_in[b], 0));
            if (c >= 2) CU(cudaStreamWaitEvent(pl->s_k, pl->ev_out[b], 0));       // the kept views of chunk c-2 have left this buffer
            long long *ds = scoring ? pl->d_sums[b] : nullptr;
            long long *ds2 = scoring ? pl->d_sums[b] + (size_t)pl->chunk * V : nullptr;
            if (int r = x360_extract_device(pl, pl->d_frames[b], n, pl->d_views[b], (int64_t *)ds, (int64_t *)ds2, pl->s_k)) return r;
            d_result[b] = pl->d_views[b];
            if (pl->sharpen) {
                if (int r = unsharp_launch(pl->d_views[b], (long long)m, pl->g.oh, pl->g.ow, pl->sharpen_strength, pl->d_sharp[b], pl->s_k)) return r;
                d_result[b] = pl->d_sharp[b];
            }
            CU(cudaEventRecord(pl->ev_k[b], pl->s_k));
            if (scoring) {
                CU(cudaStreamWaitEvent(pl->s_off, pl->ev_k[b], 0));
                CU(cudaMemcpyAsync(pl->h_sums[b], ds, sizeof(long long) * m, cudaMemcpyDeviceToHost, pl->s_off));
                CU(cudaMemcpyAsync(pl->h_sums[b] + per, ds2, sizeof(long long) * m, cudaMemcpyDeviceToHost, pl->s_off));
                CU(cudaEventRecord(pl->ev_off[b], pl->s_off));
            }
            return X360_OK;
        };
        rc = enqueue();
        if (rc == X360_OK && c >= 1) rc = finish(c - 1);
    }
    if (rc == X360_OK) rc = finish(n_chunks - 1);
    const std::string first_error = rc != X360_OK ? g_err : std::string();
    const cudaError_t e_out = cudaStreamSynchronize(pl->s_out), e_off = cudaStreamSynchronize(pl->s_off), e_k = cudaStreamSynchronize(pl->s_k),
                      e_in = cudaStreamSynchronize(pl->s_in);
    *n_kept = kept;
    if (rc != X360_OK) { g_err = first_error; return rc; }
    for (cudaError_t e : {e_out, e_off, e_k, e_in})
        if (e != cudaSuccess) return fail(X360_E_CUDA, "pipelined extraction failed: %s", cudaGetErrorString(e));
    return X360_OK;
}

int x360_host_alloc(size_t bytes, void **out_ptr)
{
    if (!out_ptr) return fail(X360_E_INVALID, "null argument");
    CU(cudaHostAlloc(out_ptr, bytes, cudaHostAllocDefault));
    return X360_OK;
}

int x360_host_free(void *ptr)
{
    if (ptr) CU(cudaFreeHost(ptr));
    return X360_OK;
}

}  // extern "C"

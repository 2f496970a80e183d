// x360_tiled.cuh — the extraction kernel: TMA in, pair records in shared memory, dp2a gather, TMA out.
//
// Contract: src/core/processor.py:327-342 for a frame batch — cv2.remap(INTER_LINEAR | INTER_LANCZOS4,
// BORDER_WRAP) and, optionally, calculate_blur_score — with the map of src/core/geometry.py:141-190
// evaluated on the fly and never stored in HBM.  extract_tiled<SCORES, LZ, TP, NW, MINB, RAW, COLSEL>:
//
//   work item   = (view unit, 32 x (NW TP) output tile, frame chunk), handed out by a global atomic
//                 counter.  A view unit is a set of views whose rotations differ only by a yaw (ring
//                 layouts, the four horizontal cube faces): their maps are one fp64 map plus a per-view
//                 shift of map_x, so the atan2 evaluations are done once per item and shared by all
//                 views and frames of the item.
//   per view    : Q5 coordinates from the shared map (+ shift, in fp64 before the float32 rounding),
//                 bounding box of the footprints, staging decision, per-pixel gather state.
//   per frame   : 1. the footprint arrives in one of two `raw` buffers as ONE TMA box (UTMALDG.3D) on
//                    an mbarrier, two frames ahead;
//                 2. transform: thread i turns the i-th 4-pixel group of the box (12 B of packed BGR)
//                    into the "pair records" of its pixels — lo = [B(x) B(x+1) G(x) G(x+1)],
//                    hi = [R(x) R(x+1)] — bilinear: two planes (STS.128 + STS.64, conflict-free),
//                    Lanczos4: contiguous 8-byte records;
//                 3. gather, bilinear: per output pixel the upper and lower record (LDS.32 + LDS.U16
//                    each), 6 dp2a (weights (32-fy|fy)(32-fx|fx)*64 as u16 pairs, bias 2^15: the
//                    result (V + 512) >> 10 lands in byte 2) + 2 PRMT; a thread walks its output
//                    column and keeps the records in registers, so most pixels load ONE record;
//                    Lanczos4: 8 source rows x 4 LDS.64 x 3 dp2a with cv2's int16 weight pairs, the
//                    per-pixel weight rows transposed through shared memory;
//                 4. pack: one SHFL + one PRMT turn four lanes' pixels into three BGR words, STS.32
//                    into the packed tile, which leaves by ONE TMA store (UTMASTG) per tile and frame;
//                 5. SCORES: gray tile + 1-px halo, dp4a Laplacian, integer sums (image_utils.py:22-27).
//   RAW (bilinear plans without the fused score, the default there): steps 2-3 collapse — no transform, no
//                 record planes; the gather reads the packed BGR bytes of the TMA box itself (2-3 LDS.32 + LOP3
//                 + 2 PRMT rebuild the pair record in registers, step_raw), three raw buffers (the box of frame
//                 f + 3 flies during frame f), two packed tiles alternating per frame, ONE barrier per frame.
//
// Exactness of the u16 weights: w*64 <= 65535 except w00 = 1024 (fx = fy = 0, all other weights
// 0), which is clamped to 65535: (65535 s + 32768) >> 16 = s for every byte s.  Everything else
// is the same integer expression as cv2.remap's (sum w s + 2^14) >> 15.
//
// Tiles whose footprint wraps the seam, touches a pole, exceeds the staging capacity or whose
// geometry defeats the TMA alignment rules fall back per (tile, view), inside the same launch, to
// the aligned-word / byte gathers of LinearSampler / LanczosSampler.
#pragma once
#include <climits>
#include <type_traits>
#include <cuda.h>

#include "x360_common.cuh"
#include "x360_lanczos.cuh"
#include "x360_linear.cuh"
#include "x360_staged.cuh"      // mbarrier / TMA load helpers

namespace x360 {

#ifndef X360_MINB8
#define X360_MINB8 6          // CTAs per SM the 32x32 bilinear variant's register allocation aims for
#endif
constexpr int kTWMax = 4;             // warps per CTA: NW = 4 or 2, lanes along x; warp w owns tile rows TP w .. TP w + TP - 1
// TP = output pixels (rows) per thread: 8 (32x32 tiles) or 4 (32x16 tiles, for geometries whose
// footprints are too large to keep enough CTAs per SM with 32 rows)
constexpr int kTBoxRows = 8;          // granularity of the footprint row capacity
constexpr int kTNumBoxW = 5;          // load box widths 48, 96, 144, 192, 240 bytes (16, 32, ... source pixels)
constexpr int kTNumBoxH = 3;          // load box heights rows_max - 16, rows_max - 8, rows_max
constexpr int kTNumBox = kTNumBoxW * kTNumBoxH;
// Width (= shared-memory row pitch) of regular box class w.  Planes: 48 (w + 1).  Raw gather, column-only views: 64 (w + 1) —
// a warp's 32 lanes read one source row, any pitch does.  Raw gather, general views (pole faces, inclined views): 64 (w + 1) - 16
// = 48, 112, 176, 240: a warp's lanes follow a line of any slope through the footprint, and with a pitch of 128 or 256 bytes
// every source row starts in bank 0 (ncu, cfg4's pole faces: 57 % of the shared-memory wavefronts were bank-conflict replays).
// A TMA box is a multiple of 16 bytes wide, so pitch / 4 mod 32 cannot be odd; = 4 mod 8 is the best there is (rows r, r + 8
// alias).  w < 4 for the raw gather.
#ifndef X360_RAW_SKEW
#define X360_RAW_SKEW 16
#endif
__host__ __device__ constexpr int tiled_box_bw(bool raw, bool col_only, int w) { return raw ? 64 * (w + 1) - (col_only ? 0 : X360_RAW_SKEW) : 48 * (w + 1); }
__host__ __device__ constexpr int tiled_box_class(bool raw, bool col_only, int need)       // smallest class that holds `need` bytes per row
{
    return raw ? (need + (col_only ? 0 : X360_RAW_SKEW) + 63) / 64 - 1 : (need + 47) / 48 - 1;
}
constexpr int kTMaxUnitViews = 8;     // views sharing one map evaluation
#ifndef X360_RAW_NB
#define X360_RAW_NB 3
#endif
constexpr int kTRawBufs = X360_RAW_NB;   // raw buffers of the raw-gather variant (the TMA box of frame f + NB - 1 flies during frame f)
// ... of the column-only plans.  The general variant (pole faces, inclined views) is latency-bound at 5 CTAs per SM (ncu: issue
// slots 44 % busy, long + short scoreboard stalls): it takes two larger buffers instead of three — fewer tiles on the direct
// gathers at the same shared memory per CTA.
#ifndef X360_RAW_NB_GENERAL
#define X360_RAW_NB_GENERAL 2
#endif
__host__ __device__ constexpr int tiled_raw_bufs(bool col_only) { return col_only ? kTRawBufs : X360_RAW_NB_GENERAL; }

// Wide boxes: tiles whose footprint is wider than the regular capacity but still fits the raw buffer as a flatter box
// (near the poles a 32-px tile spans 32 / r rad of longitude at r px from the pole; the footprint is a few rows of
// many pixels).  Width classes in bytes (multiples of 192 = lcm(48, 64): valid row pitches for both gathers), loaded
// through a uint32 view of the frame (a TMA box has at most 256 elements per dimension); the class height is whatever
// the raw buffer holds, TiledArgs::wide_rows.
constexpr int kTNumWide = 4;
__host__ __device__ constexpr int tiled_wide_bw(int c) { return 384 + 192 * c; }     // 384, 576, 768, 960
// Seam boxes (raw gather): a tile whose footprint wraps the +-pi seam reads columns [xa, W) and [0, xb] of each source
// row.  Frames are dense, so the bytes after column W-1 of row y are columns 0.. of row y+1: one box over a "window
// tensor" — base = frames + 3W - kTSeamHalf, rows of 2 kTSeamHalf bytes, row pitch 3W — brings row y's tail and row
// y+1's head side by side.  Regular widths 64 k (k <= 4) at the full box height, then the wide classes.
constexpr int kTSeamHalf = 1024;
constexpr int kTNumSeam = 4 + kTNumWide;
constexpr int kTMapWide0 = kTNumBox, kTMapSeam0 = kTMapWide0 + kTNumWide, kTMapOut = kTMapSeam0 + kTNumSeam, kTMapOutT = kTMapOut + 1, kTNumMaps = kTMapOutT + 1;

struct TiledMaps {
    // [0, 15)  uint8 [F][H][3W], box {48 (w + 1), rows_max - 8 (2 - h), 1} at [h * 5 + w]; raw gather: {64 (w + 1), ..}, w < 4
    // [15, 19) wide: uint32 [F][H][3W/4], box {tiled_wide_bw(c) / 4, wide_rows[c], 1}
    // [19, 27) seam: window tensor [1][F H - 1][2 kTSeamHalf] (uint8: widths 64 (w + 1) x rows_max; uint32: the wide classes)
    // [27]     out: uint8 [F*V][oh][3 ow], box {96, tile height, 1}
    // [28]     out, transposed tiles (16-row variant): box {48, 32, 1}
    CUtensorMap m[kTNumMaps];
};

// One view unit: `n` views starting at `first` in the (view id, shift) tables share a map.
struct TiledUnit {
    int first, n;
};

struct TiledArgs {
    Geom g;
    const uint8_t *frames;            // [F][H][W][3]
    uint8_t *out;                     // [F][V][oh][ow][3]
    long long *sum_lap, *sum_lap2;    // [F][V] or nullptr
    const double *R;                  // [V][9]
    const int16_t *lz_tab;            // [1024][8][8] Lanczos4 weights (Lanczos4 plans only; the direct gathers read it)
    const float *lz_k1;               // [32][8] cv2's 1-D Lanczos-4 coefficients (float32)
    const uint32_t *lz_fix;           // [1024] per (fy, fx): the table's sum-to-2^15 fix-up — tap (4 + (p >> 1), 4 + (p & 1)), p = bits 0-1; delta = bits 16-31 (signed)
    const TiledUnit *units;           // [n_units], largest units first
    const int *unit_view;             // view index per table entry
    const double *unit_shift;         // map_x shift (source pixels, in [0, W)) per table entry
    unsigned *counter;                // work-item counter (zeroed before the launch)
    unsigned *tile_modes;             // [3] or nullptr
    int n_units, n_items;
    int F, V, FG, n_chunks;           // frames, views, frames per chunk, chunks
    int aligned;                      // direct gathers may use aligned words
    int stage_ok;                     // TMA loads possible (frame base and pitch 16-B aligned)
    int store_ok;                     // TMA stores possible (out base and pitch 16-B aligned)
    int rows_max, bw_max;             // staging capacity: footprint rows, raw bytes per footprint row (48 k)
    int tiles_y;                      // tile rows per view (tile height 32 or 16)
    int col_u0;                       // map_x of every view is independent of the output row (R01 = R21 = 0: yaw-only rotations)
    int wide_ok, seam_ok;             // wide / seam tensor maps are encoded
    // Two-launch mode: the first launch (mode 1) runs at the capacity that keeps the most CTAs per SM and, instead of gathering
    // a (tile, view) whose footprint does not fit from global memory, appends {item, view-in-unit} to defer_list; the second
    // launch (mode 2: larger buffers, fewer CTAs per SM) takes its items from that list, each cut into defer_sub frame ranges.
    // mode 0: one launch, such tiles go straight to the direct gathers.
    int mode, defer_cap, defer_sub;
    unsigned *defer_list, *defer_count;
    // Column-only units whose views are spaced by a constant yaw (a ring): 1 / (that spacing in source pixels), else 0.  A tile
    // then starts its walk over the unit's views at the view for which its footprint lies in source sector k = 0, 1, ... in
    // step with every other tile: overlapping views read the same source bytes at the same time (L2 hits instead of a
    // second trip to DRAM).
    double rot_inv_pitch;
    // Transposed tiles (raw gather of general views, 16-row tiles, square views of whole 32 x 32 blocks): tr_tab[view][block] != 0
    // -> the two tiles of that 32 x 32 output block are its left and right halves (16 wide, 32 tall) instead of its upper and
    // lower ones, and the lanes run down the block's rows.  Set where the source row changes faster along output x than along
    // output y (the left / right quadrants of a pole face): there a warp's lanes would read a near-vertical line of the
    // footprint (bank conflicts) and the walk down a column would never reuse a record.  Same arithmetic per pixel.
    const uint8_t *tr_tab;
    int wide_rows[kTNumWide];         // box height of each wide class (raw buffer bytes / class width, at most 256)
};

struct TiledCtrl {
    int wbb[kTWMax][4];                  // per-warp footprint bounding boxes (16-B aligned rows)
    unsigned long long mbar[3];       // one per raw buffer
    int wsum[kTWMax][2];                 // per-warp (sum L, sum L^2) of the current frame's tile (exact in 32 bits)
    int item;
    int pad0;
    // the TMA box of the current (tile, view), written and read by thread 0 only: coordinates of frame f are
    // {cx, cy0 + f cyf, f czf}; `bytes` = box bytes, `map` = index into TiledMaps::m
    int box_cx, box_cy0, box_cyf, box_czf, box_bytes, box_map;
    int seam_ea, seam_rows;           // seam boxes: byte column of a row's end inside the box row; box rows in use
    int def_vi, def_sub, def_took;    // second launch: view-in-unit and frame sub-range of the item; first launch: the pair was deferred
};

// Dynamic shared memory, fixed offsets first so every hot address is base + constant:
//   ctrl | packed BGR tile | Q5 map_y | gray tile | fp64 map_x (tile or one row) | raw x2 (TMA boxes) | pair records
//
// raw:   `rows` footprint rows of bw = 48 m bytes, dense, so that 4-pixel groups (12 B) tile it without
//        a seam between rows: group i of the flat sequence starts at delta + 12 i.
// pairs: the "pair record" of source pixel x is lo = [B(x) B(x+1) G(x) G(x+1)], hi = [R(x) R(x+1)].
//        Bilinear keeps them in two planes — lo words at 16 i (group i = 4 pixels), hi halfwords at
//        8 i — so the transform stores are contiguous (conflict-free STS.128 / STS.64) and the gather
//        is one LDS.32 + one LDS.U16 per record: 32 lanes span <= 128 B of the lo plane and <= 64 B of
//        the hi plane, which costs ~3.4 wavefronts per warp where 8-byte records (LDS.64) cost ~4.4
//        (scripts/bank_sim.py planes).  Row r, pixel x: lo at r * 4 bw/3 + 4 x, hi at r * 2 bw/3 + 2 x.
//        Lanczos4 keeps contiguous 8-byte records (4 LDS.64 per source row), row pitch 8 bw/3.
constexpr uint32_t kOffCtl = 0;
constexpr uint32_t kOffTile = 256;
static_assert(sizeof(TiledCtrl) <= 256, "ctrl block");

// gray tile (TH rows + 1-px halo): rows -1..TH as compact 32-B rows, left / right halo columns apart
template <int TH> struct TGray {
    uint8_t g[TH + 2][kTile];
    uint8_t hl[TH];
    uint8_t hr[TH];
};

__host__ __device__ constexpr uint32_t tiled_off_sy(int th) { return kOffTile + (uint32_t)th * kTile * 3; }          // Q5 map_y (< 2^20): 2 byte planes + 1 nibble plane
__host__ __device__ constexpr uint32_t tiled_off_gray(int th) { return tiled_off_sy(th) + (uint32_t)th * kTile * 5 / 2; }   // 20 bits per value
// fp64 map_x: a TH x 32 tile, or ONE row of 32 where map_x does not depend on the output row (column-only units)
__host__ __device__ constexpr uint32_t tiled_gray_bytes(int th) { return ((uint32_t)((th + 2) * kTile + 2 * th) + 127u) & ~127u; }
__host__ __device__ constexpr uint32_t tiled_off_u0(bool scores, int th)
{
    return scores ? tiled_off_gray(th) + tiled_gray_bytes(th) : tiled_off_gray(th);
}
__host__ __device__ constexpr uint32_t tiled_u0_bytes(int th, bool col_only) { return col_only ? kTile * 8u : (uint32_t)th * kTile * 8u; }   // column-only: one row (every warp writes the same values)
__host__ __device__ constexpr uint32_t tiled_off_raw(bool scores, int th, bool col_only) { return tiled_off_u0(scores, th) + tiled_u0_bytes(th, col_only); }
static_assert(tiled_off_gray(32) % 128 == 0 && tiled_off_gray(16) % 128 == 0, "raw buffers must be 128-byte aligned");

// one raw buffer: rows_max x bw_max, rounded up to the TMA destination alignment
__host__ __device__ inline uint32_t tiled_raw_bytes(int rows_max, int bw_max) { return ((uint32_t)(rows_max * bw_max) + 127u) & ~127u; }
// bilinear: lo plane 4 B + hi plane 2 B per source pixel (bw / 3 pixels per row); Lanczos4: 8-byte records
__host__ __device__ inline uint32_t tiled_lo_bytes(int rows_max, int bw_max) { return (uint32_t)(rows_max * (bw_max / 3) * 4); }
__host__ __device__ inline uint32_t tiled_pairs_bytes(int rows_max, int bw_max, bool lanczos)
{
    return ((uint32_t)(rows_max * (bw_max / 3) * (lanczos ? 8 : 6)) + 16u + 15u) & ~15u;
}
__host__ __device__ inline size_t tiled_smem_bytes(int rows_max, int bw_max, bool scores, bool lanczos, int th, bool col_only, bool raw = false)
{
    // RAW: ctrl | packed tile | map_y | [gray] | map_x | raw x3 | second packed tile | [second gray tile | second per-warp sums]
    // (+ with the fused score: second gray tile, second set of per-warp sums — a frame's barrier is shared by both copies)
    if (raw) return (size_t)tiled_off_raw(scores, th, col_only) + tiled_raw_bufs(col_only) * tiled_raw_bytes(rows_max, bw_max) + (size_t)th * kTile * 3 +
                    (scores ? tiled_gray_bytes(th) + 128u : 0u);
    // ctrl | packed tile | map_y | [gray] | map_x | raw x2 | pair records
    return (size_t)tiled_off_raw(scores, th, col_only) + 2 * tiled_raw_bytes(rows_max, bw_max) + tiled_pairs_bytes(rows_max, bw_max, lanczos);
}

// Halo pixel of thread `tid`: 0-31 top, 32-63 bottom, then TH left, TH right.  False if tid has none.
template <int TH> __device__ __forceinline__ bool halo_coords_t(int tid, int x0, int y0, int &hx, int &hy)
{
    if (tid < 32) { hx = x0 + tid; hy = y0 - 1; }
    else if (tid < 64) { hx = x0 + tid - 32; hy = y0 + TH; }
    else if (tid < 64 + TH) { hx = x0 - 1; hy = y0 + tid - 64; }
    else if (tid < 64 + 2 * TH) { hx = x0 + kTile; hy = y0 + tid - 64 - TH; }
    else return false;
    return true;
}
template <int TH> __device__ __forceinline__ void halo_store_t(TGray<TH> &t, int tid, uint8_t v)
{
    if (tid < 32) t.g[0][tid] = v;
    else if (tid < 64) t.g[TH + 1][tid - 32] = v;
    else if (tid < 64 + TH) t.hl[tid - 64] = v;
    else t.hr[tid - 64 - TH] = v;
}
template <int TH> __device__ __forceinline__ int gray_at_t(const TGray<TH> &t, int ly, int lx)
{
    if (lx < 0) return t.hl[ly];
    if (lx >= kTile) return t.hr[ly];
    return t.g[ly + 1][lx];
}
// Laplacian sums of the tile (image_utils.py:27, reflect-101 borders): partial (sum L, sum L^2) of virtual
// thread vt in [0, 8 TH): one 4-pixel word each on full tiles (dp4a), a strided loop on partial tiles.
template <int TH> __device__ __forceinline__ void laplacian_partial_t(const TGray<TH> &t, const Geom &g, int x0, int y0, bool full,
                                                                      int vt, int &pl, unsigned &pl2)
{
    pl = 0;
    pl2 = 0;
    if (full) {
        const int ly = vt >> 3, wx = vt & 7;
        const uint32_t C = *reinterpret_cast<const uint32_t *>(&t.g[ly + 1][wx * 4]);
        uint32_t U = *reinterpret_cast<const uint32_t *>(&t.g[ly][wx * 4]);
        uint32_t D = *reinterpret_cast<const uint32_t *>(&t.g[ly + 2][wx * 4]);
        uint32_t l = (wx > 0) ? t.g[ly + 1][wx * 4 - 1] : t.hl[ly];
        uint32_t r = (wx < 7) ? t.g[ly + 1][wx * 4 + 4] : t.hr[ly];
        if (y0 + ly == 0) U = D;                                   // reflect-101: g[-1] = g[1], g[n] = g[n-2]
        if (y0 + ly == g.oh - 1) D = U;
        if (x0 + wx * 4 == 0) l = (C >> 8) & 0xffu;
        if (x0 + wx * 4 + 3 == g.ow - 1) r = (C >> 16) & 0xffu;
        const uint32_t Lw = __byte_perm(C, l, 0x2104);             // [l, c0, c1, c2]
        const uint32_t Rw = __byte_perm(C, r, 0x4321);             // [c1, c2, c3, r]
        constexpr uint32_t kA = 0x0001fc01u;                       // [ 1, -4,  1,  0]
        constexpr uint32_t kB = 0x01fc0100u;                       // [ 0,  1, -4,  1]
        const int L0 = dp4a_us(D, 0x00000001u, dp4a_us(U, 0x00000001u, dp4a_us(Lw, kA, 0)));
        const int L1 = dp4a_us(D, 0x00000100u, dp4a_us(U, 0x00000100u, dp4a_us(Lw, kB, 0)));
        const int L2 = dp4a_us(D, 0x00010000u, dp4a_us(U, 0x00010000u, dp4a_us(Rw, kA, 0)));
        const int L3 = dp4a_us(D, 0x01000000u, dp4a_us(U, 0x01000000u, dp4a_us(Rw, kB, 0)));
        pl = L0 + L1 + L2 + L3;
        pl2 = (unsigned)(L0 * L0) + (unsigned)(L1 * L1) + (unsigned)(L2 * L2) + (unsigned)(L3 * L3);
    } else {
        for (int i = vt; i < TH * kTile; i += 8 * TH) {
            const int ly = i >> 5, lx = i & 31;
            const int x = x0 + lx, y = y0 + ly;
            if (x >= g.ow || y >= g.oh) continue;
            const int yu = (y == 0) ? ((g.oh > 1) ? 1 : 0) : y - 1;
            const int yd = (y == g.oh - 1) ? ((g.oh > 1) ? g.oh - 2 : 0) : y + 1;
            const int xl = (x == 0) ? ((g.ow > 1) ? 1 : 0) : x - 1;
            const int xr = (x == g.ow - 1) ? ((g.ow > 1) ? g.ow - 2 : 0) : x + 1;
            const int c = gray_at_t(t, ly, lx);
            const int L = gray_at_t(t, yu - y0, lx) + gray_at_t(t, yd - y0, lx) + gray_at_t(t, ly, xl - x0) + gray_at_t(t, ly, xr - x0) - 4 * c;
            pl += L;
            pl2 += (unsigned)(L * L);
        }
    }
}

__device__ __forceinline__ uint32_t dp2a_lo_u(uint32_t a16x2, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a16x2), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t dp2a_hi_u(uint32_t a16x2, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a16x2), "r"(b), "r"(c));
    return d;
}

// 6 dp2a + 2 PRMT on the two pair records (upper / lower source row): the packed pixel [B, G, R, 0]
__device__ __forceinline__ uint32_t combine_pairs(uint2 r0, uint2 r1, uint32_t wA, uint32_t wB)
{
    const uint32_t vb = dp2a_lo_u(wB, r1.x, dp2a_lo_u(wA, r0.x, 32768u));
    const uint32_t vg = dp2a_hi_u(wB, r1.x, dp2a_hi_u(wA, r0.x, 32768u));
    const uint32_t vr = dp2a_lo_u(wB, r1.y, dp2a_lo_u(wA, r0.y, 32768u));
    return __byte_perm(__byte_perm(vb, vg, 0x0062), vr, 0x7610);
}

// One step down an output column: (X, Y) are the upper / lower pair records of the previous pixel
// (.x = lo word, .y = hi halfword).  Mode bits of pixel K in pm: bit 2K = the lower record becomes the
// upper one and a new lower record is loaded; bit 2K+1 = the upper record is loaded too.
// Predicated-off loads cost no shared-memory bandwidth, which is what bounds this kernel.
template <int K, bool A = true> __device__ __forceinline__ void step_records(uint2 &X, uint2 &Y, uint32_t lo_upper, uint32_t lo_lower,
                                                                            uint32_t hi_upper, uint32_t hi_lower, uint32_t pm)
{
    if (!A) {                                            // no lane of the warp reloads the upper record
        asm volatile("{\n\t.reg .pred pB;\n\t.reg .b32 t;\n\t"
                     "and.b32 t, %6, %7;\n\tsetp.ne.u32 pB, t, 0;\n\t"
                     "@pB mov.b32 %0, %2;\n\t@pB mov.b32 %1, %3;\n\t"
                     "@pB ld.shared.u32 %2, [%4];\n\t@pB ld.shared.u16 %3, [%5];\n\t}"
                     : "+r"(X.x), "+r"(X.y), "+r"(Y.x), "+r"(Y.y)
                     : "r"(lo_lower), "r"(hi_lower), "r"(pm), "n"(1u << (2 * K)));
        return;
    }
    asm volatile("{\n\t.reg .pred pA, pB;\n\t.reg .b32 t;\n\t"
                 "and.b32 t, %8, %9;\n\tsetp.ne.u32 pA, t, 0;\n\t"
                 "and.b32 t, %8, %10;\n\tsetp.ne.u32 pB, t, 0;\n\t"
                 "@pB mov.b32 %0, %2;\n\t@pB mov.b32 %1, %3;\n\t"
                 "@pB ld.shared.u32 %2, [%5];\n\t@pB ld.shared.u16 %3, [%7];\n\t"         // u16 -> b32 register: zero-extended
                 "@pA ld.shared.u32 %0, [%4];\n\t@pA ld.shared.u16 %1, [%6];\n\t}"
                 : "+r"(X.x), "+r"(X.y), "+r"(Y.x), "+r"(Y.y)
                 : "r"(lo_upper), "r"(lo_lower), "r"(hi_upper), "r"(hi_lower), "r"(pm), "n"(2u << (2 * K)), "n"(1u << (2 * K)));
}

// The same walking UP the column (pixel K after pixel K+1): bit 2K = the upper record becomes the
// lower one and a new upper record is loaded; bit 2K+1 = the lower record is loaded too.
template <int K, bool A = true> __device__ __forceinline__ void step_records_up(uint2 &X, uint2 &Y, uint32_t lo_upper, uint32_t lo_lower,
                                                                               uint32_t hi_upper, uint32_t hi_lower, uint32_t pm)
{
    if (!A) {
        asm volatile("{\n\t.reg .pred pB;\n\t.reg .b32 t;\n\t"
                     "and.b32 t, %6, %7;\n\tsetp.ne.u32 pB, t, 0;\n\t"
                     "@pB mov.b32 %2, %0;\n\t@pB mov.b32 %3, %1;\n\t"
                     "@pB ld.shared.u32 %0, [%4];\n\t@pB ld.shared.u16 %1, [%5];\n\t}"
                     : "+r"(X.x), "+r"(X.y), "+r"(Y.x), "+r"(Y.y)
                     : "r"(lo_upper), "r"(hi_upper), "r"(pm), "n"(1u << (2 * K)));
        return;
    }
    asm volatile("{\n\t.reg .pred pA, pB;\n\t.reg .b32 t;\n\t"
                 "and.b32 t, %8, %9;\n\tsetp.ne.u32 pA, t, 0;\n\t"
                 "and.b32 t, %8, %10;\n\tsetp.ne.u32 pB, t, 0;\n\t"
                 "@pB mov.b32 %2, %0;\n\t@pB mov.b32 %3, %1;\n\t"
                 "@pB ld.shared.u32 %0, [%4];\n\t@pB ld.shared.u16 %1, [%6];\n\t"
                 "@pA ld.shared.u32 %2, [%5];\n\t@pA ld.shared.u16 %3, [%7];\n\t}"
                 : "+r"(X.x), "+r"(X.y), "+r"(Y.x), "+r"(Y.y)
                 : "r"(lo_upper), "r"(lo_lower), "r"(hi_upper), "r"(hi_lower), "r"(pm), "n"(2u << (2 * K)), "n"(1u << (2 * K)));
}

// RAW variant (no transform): the gather reads the packed BGR bytes of the TMA box directly.  The record of source
// pixels (x, x+1) is the 6 bytes [B G R B G R] starting at byte a = o & 3 of the aligned word at `addr` (o = byte
// offset of pixel x in its footprint row): words t0, t1 and, where a = 3, t2.  Two PRMT rebuild the pair record —
// lo = [B(x) B(x+1) G(x) G(x+1)] from (t0, t1), selector 0x4130 + 0x1111 a; hi = [R(x) R(x+1) . .] from (t1, a = 3 ? t2 : t0),
// selector bytes (a + 6) & 7, (a + 1) & 7 — so the arithmetic is combine_pairs' as is.  m3 = ~0 where a = 3, else 0.
// One step of a column walk: bit 2K of pm = P <- Q and Q <- the record at aq; bit 2K+1 = P <- the record at ap.
// Walking down, P / Q are the upper / lower record; walking up, the lower / upper one.  The words land in the record's
// own registers; the third word is loaded unpredicated (a register written only under a predicate would be kept
// alive from frame to frame), then reused under the predicate.
template <int K, bool A> __device__ __forceinline__ void step_raw(uint2 &P, uint2 &Q, uint32_t ap, uint32_t aq, uint32_t sel_lo,
                                                                  uint32_t sel_hi, uint32_t pm, uint32_t m3)
{
    if (!A) {
        asm volatile("{\n\t.reg .pred pB;\n\t.reg .b32 t, t2;\n\t"
                     "and.b32 t, %7, %9;\n\tsetp.ne.u32 pB, t, 0;\n\t"
                     "@pB mov.b32 %0, %2;\n\t@pB mov.b32 %1, %3;\n\t"
                     "@pB ld.shared.u32 %2, [%4];\n\t@pB ld.shared.u32 %3, [%4+4];\n\tld.shared.u32 t2, [%4+8];\n\t"
                     "lop3.b32 t2, t2, %2, %8, 0xE4;\n\t"
                     "@pB prmt.b32 %2, %2, %3, %5;\n\t@pB prmt.b32 %3, %3, t2, %6;\n\t}"
                     : "+r"(P.x), "+r"(P.y), "+r"(Q.x), "+r"(Q.y)
                     : "r"(aq), "r"(sel_lo), "r"(sel_hi), "r"(pm), "r"(m3), "n"(1u << (2 * K)));
        return;
    }
    asm volatile("{\n\t.reg .pred pA, pB;\n\t.reg .b32 t, t2;\n\t"
                 "and.b32 t, %7, %9;\n\tsetp.ne.u32 pB, t, 0;\n\t"
                 "and.b32 t, %7, %10;\n\tsetp.ne.u32 pA, t, 0;\n\t"
                 "@pB mov.b32 %0, %2;\n\t@pB mov.b32 %1, %3;\n\t"
                 "@pB ld.shared.u32 %2, [%4];\n\t@pB ld.shared.u32 %3, [%4+4];\n\tld.shared.u32 t2, [%4+8];\n\t"
                 "lop3.b32 t2, t2, %2, %8, 0xE4;\n\t"
                 "@pB prmt.b32 %2, %2, %3, %5;\n\t@pB prmt.b32 %3, %3, t2, %6;\n\t"
                 "@pA ld.shared.u32 %0, [%11];\n\t@pA ld.shared.u32 %1, [%11+4];\n\t@pA ld.shared.u32 t2, [%11+8];\n\t"
                 "lop3.b32 t2, t2, %0, %8, 0xE4;\n\t"
                 "@pA prmt.b32 %0, %0, %1, %5;\n\t@pA prmt.b32 %1, %1, t2, %6;\n\t}"
                 : "+r"(P.x), "+r"(P.y), "+r"(Q.x), "+r"(Q.y)
                 : "r"(aq), "r"(sel_lo), "r"(sel_hi), "r"(pm), "r"(m3), "n"(1u << (2 * K)), "n"(2u << (2 * K)), "r"(ap));
}

// pair-record words of source pixel J of a 4-pixel group from its packed BGR words w[0..3]
template <int J> __device__ __forceinline__ uint32_t pair_lo(const uint32_t (&w)[4])
{
    constexpr int n = 3 * J, w0 = n >> 2, r = n - 4 * w0;           // bytes n, n+3, n+1, n+4
    constexpr uint32_t sel = (uint32_t)r | ((uint32_t)(r + 3) << 4) | ((uint32_t)(r + 1) << 8) | ((uint32_t)(r + 4) << 12);
    return __byte_perm(w[w0], w[w0 + 1], sel);
}
template <int J> __device__ __forceinline__ uint32_t pair_hi(const uint32_t (&w)[4])
{
    constexpr int n = 3 * J + 2, w0 = n >> 2, r = n - 4 * w0;       // bytes n, n+3
    constexpr uint32_t sel = (uint32_t)r | ((uint32_t)(r + 3) << 4);
    return __byte_perm(w[w0], w[(w0 + 1) < 4 ? w0 + 1 : 3], sel);
}

// 12 (+3) raw bytes at `src` -> the 4 pair records of one group, 32 B at dst
__device__ __forceinline__ void transform_group(const uint8_t *src, uint8_t *dst)
{
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) w[i] = *reinterpret_cast<const uint32_t *>(src + 4 * i);
    *reinterpret_cast<uint4 *>(dst) = make_uint4(pair_lo<0>(w), pair_hi<0>(w), pair_lo<1>(w), pair_hi<1>(w));
    *reinterpret_cast<uint4 *>(dst + 16) = make_uint4(pair_lo<2>(w), pair_hi<2>(w), pair_lo<3>(w), pair_hi<3>(w));
}

// Lanczos4 from pair records: 8 source rows x 4 records (taps c = 2j, 2j+1) x 3 channels = 96 dp2a with
// the int16 weight pairs of cv2's table row (w0|w1<<16, ..., w6|w7<<16); contiguous record layout (8 B per
// source pixel).  dst = sat_u8((sum + 2^14) >> 15) exactly as LanczosSampler::finish.
//
// Every pixel has its own (fy, fx) row of the weight table (128 B).  Round 1 fetched it (8 coalesced loads per warp, a
// transposition through 18 KB of shared memory, 8 LDS.128 per pixel: 64 of the kernel's 204 shared-memory wavefronts per 32
// pixels, with the LSU pipe 90 % busy and the issue slots 26 %).  Round 2 COMPUTES the 64 weights in registers from what the
// table is made of (x360_tables.hpp): it[r][c] = sat_s16(rint(f32(k[fy][r] * k[fx][c]) * 2^15)), plus the residue the table's
// sum-to-2^15 fix-up puts on one of the taps (4..5, 4..5).  FFMA(k[fy][r], k[fx][c] * 2^15, 1.5 * 2^23) rounds the exact
// product to the nearest even integer in the low mantissa bits — the low 16 bits ARE the int16 weight, and a PRMT packs two of
// them.  (OpenCV rounds twice — the product to float32, then to the integer; for all 65 536 entries of this table the single
// rounding gives the same integer: tests/test_host_api.py checks it exhaustively, so one FFMA replaces FMUL + FFMA.)  The only
// product that saturates is 1.0 * 1.0 (fy = fx = 0, tap (3, 3)): clamped explicitly.
__device__ __forceinline__ uint32_t lz_weight_bits(float cy, float cx_q15) { return __float_as_uint(fmaf(cy, cx_q15, 12582912.0f)); }

__device__ __forceinline__ uint32_t lanczos_pairs(const uint8_t *dsm, uint32_t a0, uint32_t ppitch, const float *__restrict__ k1,
                                                  const uint32_t *__restrict__ fix, uint32_t idx)
{
    const uint32_t fx = idx & 31u, fy = idx >> 5;
    const float4 xa = __ldg(reinterpret_cast<const float4 *>(k1 + 8 * fx)), xb = __ldg(reinterpret_cast<const float4 *>(k1 + 8 * fx) + 1);
    const float4 ya = __ldg(reinterpret_cast<const float4 *>(k1 + 8 * fy)), yb = __ldg(reinterpret_cast<const float4 *>(k1 + 8 * fy) + 1);
    const float cx[8] = {xa.x * 32768.0f, xa.y * 32768.0f, xa.z * 32768.0f, xa.w * 32768.0f,       // (exact: a power of two)
                         xb.x * 32768.0f, xb.y * 32768.0f, xb.z * 32768.0f, xb.w * 32768.0f};
    const float cy[8] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z, yb.w};
    const uint32_t fw = __ldg(fix + idx);
    const uint32_t fpos = fw & 3u, fdelta = (uint32_t)((int)fw >> 16);                 // residue as a two's-complement word
    int ab = 0, ag = 0, ar = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r, a0 += ppitch) {
        uint32_t m[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) m[c] = lz_weight_bits(cy[r], cx[c]);
        if (r == 3) m[3] = (uint32_t)min((int)(m[3] - 0x4B400000u), 32767);           // sat_s16: 1.0 * 1.0 * 2^15 = 32768
        if (r == 4 || r == 5) {                                                        // the fix-up lands on one of the taps (4..5, 4..5)
            m[4] += (fpos == (uint32_t)((r - 4) * 2)) ? fdelta : 0u;
            m[5] += (fpos == (uint32_t)((r - 4) * 2 + 1)) ? fdelta : 0u;
        }
        const uint32_t w0 = __byte_perm(m[0], m[1], 0x5410), w1 = __byte_perm(m[2], m[3], 0x5410);
        const uint32_t w2 = __byte_perm(m[4], m[5], 0x5410), w3 = __byte_perm(m[6], m[7], 0x5410);
        const uint2 c0 = *reinterpret_cast<const uint2 *>(dsm + a0), c1 = *reinterpret_cast<const uint2 *>(dsm + a0 + 16);
        const uint2 c2 = *reinterpret_cast<const uint2 *>(dsm + a0 + 32), c3 = *reinterpret_cast<const uint2 *>(dsm + a0 + 48);
        ab = dp2a_lo_su(w3, c3.x, dp2a_lo_su(w2, c2.x, dp2a_lo_su(w1, c1.x, dp2a_lo_su(w0, c0.x, ab))));
        ag = dp2a_hi_su(w3, c3.x, dp2a_hi_su(w2, c2.x, dp2a_hi_su(w1, c1.x, dp2a_hi_su(w0, c0.x, ag))));
        ar = dp2a_lo_su(w3, c3.y, dp2a_lo_su(w2, c2.y, dp2a_lo_su(w1, c1.y, dp2a_lo_su(w0, c0.y, ar))));
    }
    return LanczosSampler::finish(ab, ag, ar);
}

// 12 (+3) raw bytes at `src` -> the lo words (16 B at dst_lo) and hi halfwords (8 B at dst_hi) of one group
__device__ __forceinline__ void transform_group_planes(const uint8_t *src, uint8_t *dst_lo, uint8_t *dst_hi)
{
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) w[i] = *reinterpret_cast<const uint32_t *>(src + 4 * i);
    *reinterpret_cast<uint4 *>(dst_lo) = make_uint4(pair_lo<0>(w), pair_lo<1>(w), pair_lo<2>(w), pair_lo<3>(w));
    // R(j) = byte 3j + 2: [R0 R1 | R1 R2] = bytes 2, 5, 5, 8; [R2 R3 | R3 R4] = bytes 8, 11, 11, 14
    const uint32_t h0 = __byte_perm(__byte_perm(w[0], w[1], 0x0552), w[2], 0x4210);
    const uint32_t h1 = __byte_perm(w[2], w[3], 0x6330);
    *reinterpret_cast<uint2 *>(dst_hi) = make_uint2(h0, h1);
}

__device__ __forceinline__ void tma_store_tile(const CUtensorMap *tm, uint32_t src_smem, int x, int y, int z)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                 ::"l"(tm), "r"(x), "r"(y), "r"(z), "r"(src_smem) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// mbarrier / TMA helpers on 32-bit shared-window addresses kept in registers (the pointer forms make the
// compiler rebuild the window base per use)
__device__ __forceinline__ void mbar_wait_s(uint32_t bar_s, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "W_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@!p bra W_%=;\n\t}"
                 ::"r"(bar_s), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_issue_box_s(uint32_t dst_s, const CUtensorMap *tm, int x, int y, int z, uint32_t bar_s, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst_s), "l"(tm), "r"(x), "r"(y), "r"(z), "r"(bar_s) : "memory");
}

// Occupancy targets (CTAs per SM) the register allocation aims for, by variant.
__host__ __device__ constexpr int tiled_min_ctas(bool scores, bool lz, int tp, int nw)
{
    return nw == 2 ? (lz ? 8 : 10) : (lz ? (tp == 4 ? 5 : 4) : (tp == 4 ? (scores ? 6 : 7) : (scores ? 4 : X360_MINB8)));
}

// MINB: the CTAs/SM the register allocation aims for.  The default table, or 7 for 32 x 32 bilinear tiles of
// yaw-only units: their map_x row leaves 30 KB of shared memory per CTA, and 72 registers then buy a 7th CTA.
//
// RAW: no transform and no pair-record planes — the gather reads the TMA box itself (step_raw), the footprint rows
// are 64 k bytes wide, the packed output tile is double-buffered (second copy after the raw buffers) and a frame
// costs ONE block barrier.  Bilinear, no fused score.  COLSEL: column-only map_x — every pixel of a thread's column
// has the same byte alignment in its source row, so the two PRMT selectors and the a = 3 mask are per-thread
// constants; otherwise they travel per pixel in one register (selector lo | selector hi << 16 | (a = 3) << 31).
template <bool SCORES, bool LZ, int TP, int NW, int MINB = tiled_min_ctas(SCORES, LZ, TP, NW), bool RAW = false, bool COLSEL = true>
__global__ void __launch_bounds__(32 * NW, MINB)
extract_tiled(const __grid_constant__ TiledArgs a, const __grid_constant__ TiledMaps tms)
{
    static_assert(!RAW || !LZ, "the raw gather is bilinear");
    extern __shared__ __align__(128) uint8_t dsm[];
    const Geom &g = a.g;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t pitch = (uint32_t)g.W * 3u;
    const size_t frame_bytes = (size_t)g.H * pitch;
    const size_t view_bytes = (size_t)g.oh * g.ow * 3;
    const size_t opitch = (size_t)g.ow * 3;

    static_assert(TP == 8 || TP == 4, "the gather is written out for 8 or 4 rows per thread");
    static_assert(NW == 4 || (NW == 2 && !SCORES), "the halo of the fused score needs 64 + 2 TH threads");
    constexpr int kTP = TP;
    constexpr int NB = RAW ? tiled_raw_bufs(COLSEL) : 2;                            // raw buffers
    constexpr int kTW = NW, kTT = 32 * NW;                                          // warps / threads per CTA
    constexpr int TH = kTW * TP;                                                    // tile height
    TiledCtrl &ctl = *reinterpret_cast<TiledCtrl *>(dsm + kOffCtl);
    double *const u0s = reinterpret_cast<double *>(dsm + tiled_off_u0(SCORES, TH));
    uint8_t *const sy_s = dsm + tiled_off_sy(TH);                                   // 20-bit values: 2 byte planes + 1 nibble plane of TH x 32
    const bool col_u0 = a.col_u0 != 0;
    uint32_t kOffRaw = tiled_off_raw(SCORES, TH, col_u0);                           // (a register: two layouts)
    asm volatile("" : "+r"(kOffRaw));
    uint32_t raw_bytes = tiled_raw_bytes(a.rows_max, a.bw_max);
    asm volatile("" : "+r"(raw_bytes));                                             // keep in a register (cheaper than rematerialising)
    const uint32_t off_pairs = kOffRaw + (uint32_t)NB * raw_bytes;                  // dsm-relative (RAW: the second packed tile)
    const uint32_t off_hi = off_pairs + tiled_lo_bytes(a.rows_max, a.bw_max);          // bilinear: hi plane after the lo plane
    const uint32_t off_gray2 = off_pairs + (uint32_t)TH * kTile * 3u;                  // RAW + SCORES: second gray tile, then second per-warp sums
    const uint32_t off_wsum2 = off_gray2 + tiled_gray_bytes(TH);
    const uint32_t dsm_s = smem_u32(dsm);
    const uint32_t raw_s = dsm_s + kOffRaw, otile_s = dsm_s + kOffTile;
    const uint32_t mbar_s = dsm_s + kOffCtl + (uint32_t)offsetof(TiledCtrl, mbar);

    if (tid == 0) { mbar_init(&ctl.mbar[0], 1); mbar_init(&ctl.mbar[1], 1); mbar_init(&ctl.mbar[2], 1); }
    uint32_t phase = 0;               // bit b: parity of raw buffer b's mbarrier
    uint32_t rbn = 0;                 // raw buffer of the next frame (it mod NB; RAW: of the next STAGED frame)
    uint32_t rpar = 0;                // RAW: mbarrier parity of the current round through the raw buffers
    uint32_t it = 0;                  // frames processed by this CTA: selects the raw buffer

    // frame-invariant: where this lane's packed word of row k goes (+ k * 96), and how it is built
    const int q = lane & 3;
    uint32_t pack_sel = q == 0 ? 0x4210u : (q == 1 ? 0x5421u : 0x6542u);
    uint32_t my_word = (uint32_t)(warp * kTP) * (kTile * 3) + (uint32_t)((lane >> 2) * 3 + q) * 4u;
    asm volatile("" : "+r"(pack_sel), "+r"(my_word));      // keep both in registers: rematerialising them costs more than the gather
    // RAW: shared-window address of this lane's word in the current packed tile, flipped between the two tiles per frame
    uint32_t ot = dsm_s + kOffTile + my_word, ot_toggle = RAW ? (ot ^ (dsm_s + off_pairs + my_word)) : 0u;
    asm volatile("" : "+r"(ot), "+r"(ot_toggle));
    const int tiles_per_view = g.tiles_x * a.tiles_y;
    const double Wd = (double)g.W;

    for (;;) {
        __syncthreads();                                   // previous item's shared state is dead
        // DEFER: the two-launch mode exists in every variant but the raw gather of column-only plans (yaw-only views never
        // come near a pole; their 64-register build has no room for it)
        constexpr bool DEFER = !(RAW && COLSEL);
        if (tid == 0) {
            int it0 = (int)atomicAdd(a.counter, 1u);
            if (DEFER && a.mode == 2) {                        // items of the second launch: the pairs the first one deferred
                const unsigned n_def = min(*a.defer_count, (unsigned)a.defer_cap);
                if ((unsigned)it0 >= n_def * (unsigned)a.defer_sub) it0 = INT_MAX;
                else {
                    const int e = it0 / a.defer_sub;
                    ctl.def_sub = it0 - e * a.defer_sub;
                    ctl.def_vi = (int)a.defer_list[2 * e + 1];
                    it0 = (int)a.defer_list[2 * e];
                }
            }
            ctl.item = it0;
        }
        __syncthreads();
        // (everything below comes from shared memory at CTA-uniform addresses: the compiler keeps item, tile, frame range and
        // what derives from them — buffer indices, row pitches — in uniform registers)
        const int item = ctl.item;
        if (item >= a.n_items) break;
        const int only_vi = (DEFER && a.mode == 2) ? ctl.def_vi : -1, sub = (DEFER && a.mode == 2) ? ctl.def_sub : 0;
        const int per_unit = tiles_per_view * a.n_chunks;      // every unit has the same number of items
        const int u = item / per_unit, local = item - u * per_unit;
        TiledUnit unit = a.units[u];
        if (only_vi >= 0) { unit.first += only_vi; unit.n = 1; }   // second launch: the one deferred view, as a unit of its own
        const int chunk = local / tiles_per_view, tile = local - chunk * tiles_per_view;
        const int ty = tile / g.tiles_x;
        int x0 = (tile - ty * g.tiles_x) * kTile, y0 = ty * TH;
        // transposed tile: from here on (x0, y0), the lane and the row index k are coordinates of the TRANSPOSED view — lane along
        // the real y axis, rows along the real x axis; only the map's arguments and the store know
        constexpr bool TRANSPOSABLE = RAW && !COLSEL && !SCORES && TP == 4;
        bool tr = false;
        if constexpr (TRANSPOSABLE) {
            if (a.tr_tab) tr = a.tr_tab[(size_t)a.unit_view[a.units[u].first] * ((size_t)g.tiles_x * (a.tiles_y >> 1)) + (size_t)(ty >> 1) * g.tiles_x + x0 / kTile] != 0;
            if (tr) { const int bx = x0; x0 = (ty >> 1) * kTile; y0 = bx + (ty & 1) * TH; }
        }
        int f_begin = chunk * a.FG, f_end = min(a.F, f_begin + a.FG);
        if (DEFER && a.mode == 2) {                            // this item's share of the chunk's frames
            const int len = (f_end - f_begin + a.defer_sub - 1) / a.defer_sub;
            f_begin += sub * len;
            f_end = min(f_end, f_begin + len);
            if (f_begin >= f_end) continue;
        }
        const bool full = (x0 + kTile <= g.ow) && (y0 + TH <= g.oh);

        // ---- 1. the unit's map for this tile, to shared memory: fp64 map_x, Q5 map_y -------------
        double hu0 = 0.0, u_first = 0.0;
        int hsy = 0;
        bool hvalid = false;
        {
            const double *R = a.R + 9 * a.unit_view[unit.first];
            const int x = min(x0 + lane, g.ow - 1);
            const double xn = tr ? map_yn(g, x) : map_xn(g, x);          // transposed: the lane's coordinate is the real y
            // yn of this warp's rows: lane k divides for row k, the loop broadcasts (a division per pixel otherwise)
            const int yrow = min(y0 + warp * kTP + min(lane, kTP - 1), g.oh - 1);
            const double yn_lane = tr ? map_xn(g, yrow) : map_yn(g, yrow);
            double rxz = 0.0;
            int sy_hi = 0;
            // (Lanczos4 plans are typically single-frame jobs — nothing amortises the map — and run 12 warps per SM: two pixels
            // per iteration give the dependent fp64 chains of atan2 some instruction-level parallelism)
            constexpr int kMapUnroll = LZ ? 2 : 1;
#pragma unroll kMapUnroll
            for (int k = 0; k < kTP; ++k) {
                double v0, v1, v2;
                const double yn_k = __shfl_sync(0xffffffffu, yn_lane, k);
                if (TRANSPOSABLE && tr) map_ray_n(R, yn_k, xn, v0, v1, v2);           // (real xn, real yn): the same expression, bit for bit
                else map_ray_n(R, xn, yn_k, v0, v1, v2);
                // yaw-only rotations: v0 and v2 — hence map_x and the ray's horizontal length — do not depend on the row:
                // one atan2 and one square root per column and warp
                if (!col_u0) { u0s[k * kTT + tid] = map_u_of(g, v0, v2); rxz = map_rxz(v0, v2); }
                else if (k == 0) { u_first = map_u_of(g, v0, v2); u0s[lane] = u_first; rxz = map_rxz(v0, v2); }
                const int syq = q5_of(map_y_from(g, v1, rxz));             // 0 .. 32 H < 2^20
                sy_s[k * kTT + tid] = (uint8_t)syq;
                sy_s[TH * kTile + k * kTT + tid] = (uint8_t)(syq >> 8);
                if (k & 1) sy_s[2 * TH * kTile + (k >> 1) * kTT + tid] = (uint8_t)(sy_hi | ((syq >> 16) << 4));   // bits 16-19 of rows k-1, k
                sy_hi = syq >> 16;
            }
            if (SCORES) {
                int hx = -1, hy = -1;
                hvalid = halo_coords_t<TH>(tid, x0, y0, hx, hy) && hx >= 0 && hx < g.ow && hy >= 0 && hy < g.oh;
                if (hvalid) {
                    float my;
                    map_f64(R, g, hx, hy, hu0, my);
                    hsy = q5_of(my);
                }
            }
        }

        int rot = 0;                                             // uniform: every warp holds the same column values
        if (col_u0 && a.rot_inv_pitch > 0.0 && unit.n > 1) rot = (int)(__shfl_sync(0xffffffffu, u_first, kTile / 2) * a.rot_inv_pitch) % unit.n;
        for (int vk = 0; vk < unit.n; ++vk) {
            const int vi = vk >= rot ? vk - rot : vk - rot + unit.n;
            const int v = a.unit_view[unit.first + vi];
            const double shift = (DEFER && a.mode == 2) ? 0.0 : a.unit_shift[unit.first + vi];   // second launch: the view's own map (unit of one)

            // ---- 2. this view's Q5 coordinates and the footprint bounding box ----------------
            constexpr int NS = SCORES ? kTP + 1 : kTP;
            int sxs[kTP + 1], sys[kTP + 1];
#pragma unroll
            for (int k = 0; k < kTP; ++k) {
                double uu = u0s[col_u0 ? lane : k * kTT + tid] + shift;
                if (uu > Wd) uu -= Wd;
                sxs[k] = q5_of(__double2float_rn(uu));
                sys[k] = (int)sy_s[k * kTT + tid] | ((int)sy_s[TH * kTile + k * kTT + tid] << 8) |
                         ((((int)sy_s[2 * TH * kTile + (k >> 1) * kTT + tid] >> (4 * (k & 1))) & 15) << 16);
            }
            sxs[kTP] = sxs[0];
            sys[kTP] = sys[0];
            if (SCORES && hvalid) {
                double uu = hu0 + shift;
                if (uu > Wd) uu -= Wd;
                sxs[kTP] = q5_of(__double2float_rn(uu));
                sys[kTP] = hsy;
            }
            int xmin = INT_MAX, xmax = INT_MIN, ymin = INT_MAX, ymax = INT_MIN;
#pragma unroll
            for (int k = 0; k < NS; ++k) {
                const int ix = sxs[k] >> 5, iy = sys[k] >> 5;
                xmin = min(xmin, ix); xmax = max(xmax, ix);
                ymin = min(ymin, iy); ymax = max(ymax, iy);
            }
            xmin = __reduce_min_sync(0xffffffffu, xmin); xmax = __reduce_max_sync(0xffffffffu, xmax);
            ymin = __reduce_min_sync(0xffffffffu, ymin); ymax = __reduce_max_sync(0xffffffffu, ymax);
            if (lane == 0) *reinterpret_cast<int4 *>(ctl.wbb[warp]) = make_int4(xmin, xmax, ymin, ymax);
            __syncthreads();
#pragma unroll
            for (int w = 0; w < kTW; ++w) {
                const int4 b = *reinterpret_cast<const int4 *>(ctl.wbb[w]);
                xmin = min(xmin, b.x); xmax = max(xmax, b.y); ymin = min(ymin, b.z); ymax = max(ymax, b.w);
            }

            // footprint of the taps: bilinear (ix, iy) .. (+1, +1); Lanczos4 (ix-3, iy-3) .. (ix+4, iy+4).
            // Pair records for x in [x_org, xmax + MX1 - 1], source rows [y_org, ymax + MY1].
            constexpr int MX0 = LZ ? 3 : 0, MX1 = LZ ? 4 : 1, MY0 = LZ ? 3 : 0, MY1 = LZ ? 4 : 1;
            // RAW: a footprint across the +-pi seam (columns near W-1 and near 0, or a tap at column W) is re-measured in
            // unwrapped columns — those left of W/2 continue past W — and staged through the seam window (kTSeamHalf)
            bool seam = false;
            if (RAW && a.seam_ok && (xmax - xmin > (g.W >> 1) || xmax + MX1 > g.W - 1)) {      // uniform over the CTA
                int lo = INT_MAX, hi = INT_MIN;
#pragma unroll
                for (int k = 0; k < NS; ++k) {
                    int ix = sxs[k] >> 5;
                    ix += (ix < (g.W >> 1)) ? g.W : 0;
                    lo = min(lo, ix); hi = max(hi, ix);
                }
                lo = __reduce_min_sync(0xffffffffu, lo); hi = __reduce_max_sync(0xffffffffu, hi);
                __syncthreads();                                         // every thread has read the first exchange
                if (lane == 0) { ctl.wbb[warp][0] = lo; ctl.wbb[warp][1] = hi; }
                __syncthreads();
#pragma unroll
                for (int w = 0; w < kTW; ++w) { lo = min(lo, ctl.wbb[w][0]); hi = max(hi, ctl.wbb[w][1]); }
                xmin = lo; xmax = hi;
                seam = true;
            }
            const int x_org = RAW ? xmin : ((xmin - MX0) & ~3);
            const int y_org = ymin - MY0 - (seam ? 1 : 0);               // seam: box row j holds row y_org + j up to column W-1, then row y_org + j + 1 from column 0
            const int npx = xmax + MX1 - x_org, rows = ymax + MY1 + 1 - y_org;
            const int gx0 = (3 * x_org) & ~15;                           // TMA: 16-byte aligned byte column
            const int delta = 3 * x_org - gx0;                           // 0, 4, 8 or 12 (RAW: 0 .. 15)
            const int need = delta + 3 * (npx + 1) + (seam ? 8 : 0);     // bytes of a row the records depend on (seam: + the 8-byte straddle strip at the row's end)
            const int e_a = delta + 3 * (g.W - x_org);                   // seam: where a box row passes column W-1 of its source row
            // box class: a regular box {64 / 48 (wsel + 1) bytes x one of three heights}, else a wide class (kTNumWide)
            int tmi = -1;
            uint32_t bw = (uint32_t)tiled_box_bw(RAW, COLSEL, 0), box_rows = 8u;
            {
                const int wsel = max(0, tiled_box_class(RAW, COLSEL, need));
                const uint32_t bwr = (uint32_t)tiled_box_bw(RAW, COLSEL, wsel);
                if (rows <= a.rows_max && bwr <= (uint32_t)a.bw_max) {
                    // one TMA box per frame: the smallest of the three box heights that covers the footprint (seam: the full height)
                    const int hsel = seam ? kTNumBoxH - 1 : max(0, kTNumBoxH - 1 - (a.rows_max - rows) / kTBoxRows);
                    bw = bwr;
                    box_rows = (uint32_t)(a.rows_max - kTBoxRows * (kTNumBoxH - 1 - hsel));
                    tmi = seam ? kTMapSeam0 + wsel : hsel * kTNumBoxW + wsel;
                } else if (a.wide_ok) {
#pragma unroll
                    for (int c = kTNumWide - 1; c >= 0; --c)
                        if (need <= tiled_wide_bw(c) && rows <= a.wide_rows[c]) {
                            bw = (uint32_t)tiled_wide_bw(c); box_rows = (uint32_t)a.wide_rows[c];
                            tmi = (seam ? kTMapSeam0 + 4 : kTMapWide0) + c;
                        }
                }
            }
            // (coordinates outside [0, W-2] x [0, H-2] mean a wrapped tap: seam or pole)
            bool staged = a.stage_ok && tmi >= 0 && ymin - MY0 >= 0 && ymax + MY1 <= g.H - 1;
            if (seam) staged = staged && y_org >= 0 && y_org + (int)box_rows <= g.H - 1 &&                  // the last box row borrows the row after it
                               3 * g.W - gx0 <= kTSeamHalf && gx0 + (int)bw - 3 * g.W <= kTSeamHalf;        // inside the seam window
            else staged = staged && xmin - MX0 >= 0 && xmax + MX1 <= g.W - 1;
            const bool wide = staged && tmi >= kTMapWide0 && !(seam && tmi < kTMapSeam0 + 4);
            bool generic = !a.aligned;
            bool deferred = false;
            if (DEFER && !staged && a.mode == 1) {                       // first launch: leave it to the second one (larger buffers)
                if (tid == 0) {
                    const unsigned idx = atomicAdd(a.defer_count, 1u);
                    const bool took = idx < (unsigned)a.defer_cap;        // (a full list: gather it here after all)
                    if (took) {
                        a.defer_list[2 * idx] = (unsigned)item;
                        a.defer_list[2 * idx + 1] = (unsigned)vi;
                    }
                    ctl.def_took = took ? 1 : 0;
                }
                __syncthreads();
                deferred = ctl.def_took != 0;                            // (rewritten only after this view's closing barrier)
            }
            if (!staged) {                                               // uniform over the CTA
                bool special = false;
#pragma unroll
                for (int k = 0; k < NS; ++k) {
                    if constexpr (LZ) {
                        (void)LanczosSampler::make_px(sxs[k], sys[k], g, nullptr, special);
                    } else {
                        const int wx0 = wrap_idx(sat_s16(sxs[k] >> 5), g.W), wy0 = wrap_idx(sat_s16(sys[k] >> 5), g.H);
                        const uint32_t off = (uint32_t)(wy0 * g.W + wx0) * 3u;
                        special |= (wx0 == g.W - 1) | (wy0 == g.H - 1) | ((off & ~3u) + pitch + 12u > (uint32_t)g.H * pitch);
                    }
                }
                generic = (__syncthreads_or(special) != 0) || !a.aligned;
            }
            if (a.tile_modes && tid == 0 && f_begin == 0 && !deferred) atomicAdd(&a.tile_modes[staged ? (seam ? 4 : (wide ? 3 : 0)) : (generic ? 2 : 1)], 1u);

            // ---- 3. per-pixel gather state: staged {upper record, lower record, w0x, w1x}, direct {off, wx, wy} --
            // (Lanczos4: contiguous records, 8 B per source pixel, state {record of tap (0, 0), table row offset})
            const uint32_t ppitch = (bw / 3u) * (LZ ? 8u : 4u);                  // Lanczos4 record rows / bilinear lo-plane rows
            uint32_t pa[kTP + 1], pa2[kTP + 1], ph[kTP + 1], ph2[kTP + 1], pb[kTP + 1], pc[kTP + 1];   // pa/pa2: lo plane, ph/ph2: hi plane
            pa[kTP] = pa2[kTP] = ph[kTP] = ph2[kTP] = pb[kTP] = pc[kTP] = 0;
            uint32_t psel[kTP + 1];                                              // RAW without COLSEL: per-pixel selectors
#pragma unroll
            for (int k = 0; k <= kTP; ++k) psel[k] = 0;
#pragma unroll
            for (int k = 0; k < NS; ++k) {
                if constexpr (LZ) {
                    if (staged) {
                        const uint32_t rx = (uint32_t)((sxs[k] >> 5) - 3 - x_org), ry = (uint32_t)((sys[k] >> 5) - 3 - y_org);
                        pa[k] = off_pairs + ry * ppitch + rx * 8u;                       // dsm-relative
                        pb[k] = (uint32_t)((sys[k] & 31) * 32 + (sxs[k] & 31));              // (fy, fx): index of the weight table row
                    } else {
                        bool dummy = false;
                        const LanczosSampler::Px d = LanczosSampler::make_px(sxs[k], sys[k], g, nullptr, dummy);
                        pa[k] = d.off; pb[k] = d.info;
                    }
                    pa2[k] = 0; ph[k] = 0; ph2[k] = 0; pc[k] = 0;
                } else if (staged) {
                    const uint32_t fx = sxs[k] & 31, fy = sys[k] & 31;
                    const uint32_t rx = (uint32_t)((sxs[k] >> 5) - x_org), ry = (uint32_t)((sys[k] >> 5) - y_org);
                    if constexpr (RAW) {
                        uint32_t o = (uint32_t)delta + 3u * rx, ryb = ry;
                        if (seam) {                                                      // columns in unwrapped coordinates
                            const int ixu = (sxs[k] >> 5) + (((sxs[k] >> 5) < (g.W >> 1)) ? g.W : 0);
                            o = (uint32_t)delta + 3u * (uint32_t)(ixu - x_org);
                            if (ixu >= g.W) ryb = ry - 1u;                               // past the seam: the head of the next source row sits one box row up
                            else if (ixu == g.W - 1) o = bw - 8u;                        // straddles it: the strip [pixel W-1 | pixel 0] built per frame
                        }
                        pa[k] = raw_s + ryb * bw + (o & ~3u);                           // upper record in raw buffer 0; lower: + bw
                        pa2[k] = 0; ph[k] = 0; ph2[k] = 0;
                        if (!COLSEL || k == kTP) {
                            const uint32_t al_k = o & 3u;
                            psel[k] = (0x4130u + 0x1111u * al_k) | ((((al_k + 6u) & 7u) | (((al_k + 1u) & 7u) << 4)) << 16) | (al_k == 3u ? 0x80000000u : 0u);
                        }
                    } else {
                    pa[k] = dsm_s + off_pairs + ry * ppitch + rx * 4u;                  // shared-window addresses
                    pa2[k] = pa[k] + ppitch;
                    ph[k] = dsm_s + off_hi + ry * (ppitch >> 1) + rx * 2u;
                    ph2[k] = ph[k] + (ppitch >> 1);
                    }
                    const uint32_t ax0 = 32u - fx, ay0 = (32u - fy) * 64u, ay1 = fy * 64u;
                    pb[k] = min(ay0 * ax0, 65535u) | ((ay0 * fx) << 16);
                    pc[k] = (ay1 * ax0) | ((ay1 * fx) << 16);
                } else {
                    bool dummy = false;
                    const LinearSampler::Px d = LinearSampler::make_px(sxs[k], sys[k], g, nullptr, dummy);
                    pa[k] = d.off; pa2[k] = 0; ph[k] = 0; ph2[k] = 0; pb[k] = d.wx; pc[k] = d.wy;
                }
            }

            // vertical reuse: a thread walks down one output column, so consecutive pixels usually share
            // a pair record (same source column; source row the same or one further down).  Two bits
            // per pixel: 0 keep both records, 1 lower record becomes the upper one (load one), 2 load both.
            // Two chains start from the middle rows and walk apart (interleaved for ILP): rows HC-1 .. 0 upwards
            // (modes relative to the pixel below) and rows HC .. kTP-1 downwards (relative to the pixel above).
            constexpr int HC = kTP / 2;
            uint32_t pm = 3u << (2 * (HC - 1));
#pragma unroll
            for (int k = 0; k < kTP; ++k) {
                if (LZ || k == HC - 1) continue;
                const int o = k < HC ? k + 1 : k - 1;                 // the pixel this one is derived from
                const bool same = (sxs[k] >> 5) == (sxs[o] >> 5);
                const int adv = k < HC ? (sys[o] >> 5) - (sys[k] >> 5) : (sys[k] >> 5) - (sys[o] >> 5);
                pm |= ((same && adv == 0) ? 0u : ((same && adv == 1) ? 1u : 3u)) << (2 * k);
            }

            // does any lane of this warp ever reload the upper record after its first pixel?  (Never where the
            // source is not minified along y and the source column is constant: most ring / cube tiles.)
            const bool any_a = __any_sync(0xffffffffu, ((pm & ~(3u << (2 * (HC - 1)))) & 0xaaaaaaaau) != 0u);
            // RAW: byte alignment of this thread's source column (the same for its 8 pixels: column-only map_x)
            uint32_t al = 0u;
            if (RAW) {
                const int ix0 = (sxs[0] >> 5) + ((seam && (sxs[0] >> 5) < (g.W >> 1)) ? g.W : 0);
                al = (seam && ix0 == g.W - 1) ? 0u : (((uint32_t)delta + 3u * (uint32_t)(ix0 - x_org)) & 3u);   // (the strip is word-aligned)
            }
            uint32_t sel_lo = 0x4130u + 0x1111u * al, sel_hi = ((al + 6u) & 7u) | (((al + 1u) & 7u) << 4);
            uint32_t m3 = al == 3u ? 0xffffffffu : 0u;
            asm volatile("" : "+r"(sel_lo), "+r"(sel_hi), "+r"(m3));

            // transform: the flat sequence of 4-pixel groups tid, tid + 128, ... of rows * (bw / 12)
            const int tgroups = (staged && !RAW) ? rows * (int)(bw / 12u) : 0;
            int tn = tgroups > tid ? (tgroups - tid + kTT - 1) / kTT : 0;        // this thread's groups per frame
            asm volatile("" : "+r"(tn));
            uint32_t tsrc0 = kOffRaw + (uint32_t)delta + 12u * tid;                 // in raw buffer 0
            uint32_t tdst0 = off_pairs + (LZ ? 32u : 16u) * tid;                   // records / lo plane; hi plane: off_hi + 8 i
            uint32_t thi0 = off_hi + 8u * tid;
            asm volatile("" : "+r"(tsrc0), "+r"(tdst0), "+r"(thi0));     // hoisted out of the frame loop for good
            auto issue_box = [&](int f, uint32_t b) {                    // thread 0: frame f -> raw buffer b (box parameters in ctl, its own)
                tma_issue_box_s(raw_s + b * raw_bytes, &tms.m[ctl.box_map], ctl.box_cx, ctl.box_cy0 + f * ctl.box_cyf, f * ctl.box_czf,
                                mbar_s + 8u * b, (uint32_t)ctl.box_bytes);
            };
            if (staged && tid == 0) {                                    // NB frames in flight
                const int cxb = seam ? gx0 - (3 * g.W - kTSeamHalf) : gx0;           // byte column in the frame tensor / the seam window
                const bool u32 = tmi >= kTMapWide0 && !(seam && tmi < kTMapSeam0 + 4);  // wide classes: uint32 elements
                ctl.box_cx = u32 ? cxb >> 2 : cxb;
                ctl.box_cy0 = y_org; ctl.box_cyf = seam ? g.H : 0; ctl.box_czf = seam ? 0 : 1;
                ctl.box_bytes = (int)(box_rows * bw); ctl.box_map = tmi;
                ctl.seam_ea = e_a; ctl.seam_rows = rows;
                uint32_t b = rbn;
#pragma unroll
                for (int j = 0; j < NB; ++j, b = (b + 1u == (uint32_t)NB) ? 0u : b + 1u)
                    if (f_begin + j < f_end) issue_box(f_begin + j, b);
            }

            // ---- 4. walk the frame chunk ---------------------------------------------------------
            const int f_lo = deferred ? f_end : f_begin;                 // a deferred pair: no frames here
            for (int f = f_lo; f < f_end; ++f, ++it) {
                const uint32_t rb = rbn;
                if (!RAW || staged) rbn = (rbn + 1u == (uint32_t)NB) ? 0u : rbn + 1u;
                if (staged) {
                    if constexpr (RAW) {                                  // buffers are used in cyclic order: one parity per round
                        mbar_wait_s(mbar_s + 8u * rb, rpar);
                        if (rbn == 0u) rpar ^= 1u;
                        if (seam) {                                       // uniform over the CTA
                            // the straddle strip: box row j gets [pixel W-1 of its source row | pixel 0 of the same row, which
                            // arrived at the tail of box row j-1] in its last 8 bytes (the box is at least 8 bytes wider than the footprint)
                            const uint32_t ea = (uint32_t)ctl.seam_ea, base = raw_s + rb * raw_bytes;
                            for (int j = tid + 1; j < ctl.seam_rows; j += kTT) {
                                const uint32_t row = base + (uint32_t)j * bw;
#pragma unroll
                                for (uint32_t i = 0; i < 3u; ++i) {
                                    uint32_t p0, p1;
                                    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(p0) : "r"(row + ea - 3u + i));
                                    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(p1) : "r"(row - bw + ea + i));
                                    asm volatile("st.shared.u8 [%0], %1;" ::"r"(row + bw - 8u + i), "r"(p0) : "memory");
                                    asm volatile("st.shared.u8 [%0], %1;" ::"r"(row + bw - 5u + i), "r"(p1) : "memory");
                                }
                            }
                            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the next box will overwrite these bytes
                            __syncthreads();
                        }
                    } else {
                    mbar_wait_s(mbar_s + 8u * rb, (phase >> rb) & 1u);
                    phase ^= 1u << rb;
                    }
                    uint32_t src = tsrc0 + rb * raw_bytes, dst = tdst0;
                    if constexpr (RAW) {
                        (void)src; (void)dst;
                    } else if constexpr (LZ) {
#pragma unroll 2
                        for (int i = tid; i < tgroups; i += kTT, src += 12u * kTT, dst += 32u * kTT)
                            transform_group(dsm + src, dsm + dst);
                    } else {
                        uint32_t dsth = thi0;
                        if (tn & 1) {
                            transform_group_planes(dsm + src, dsm + dst, dsm + dsth);
                            src += 12u * kTT; dst += 16u * kTT; dsth += 8u * kTT;
                        }
#pragma unroll 1
                        for (int n = tn >> 1; n > 0; --n, src += 24u * kTT, dst += 32u * kTT, dsth += 16u * kTT) {
                            transform_group_planes(dsm + src, dsm + dst, dsm + dsth);
                            transform_group_planes(dsm + src + 12u * kTT, dsm + dst + 16u * kTT, dsm + dsth + 8u * kTT);
                        }
                    }
                }
                // the TMA store of the previous frame must have finished reading the packed tile
                if constexpr (!RAW) {
                    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    __syncthreads();                  // pairs ready, raw consumed, tile buffer free
                }
                // RAW: gray tile and per-warp sums alternate between two copies like the packed tile (one barrier per frame)
                TGray<TH> &gt = *reinterpret_cast<TGray<TH> *>(dsm + ((RAW && (it & 1u)) ? off_gray2 : tiled_off_gray(TH)));
                int (*const wsum)[2] = (RAW && (it & 1u)) ? reinterpret_cast<int (*)[2]>(dsm + off_wsum2) : ctl.wsum;
                int (*const wsum_prev)[2] = (RAW && !(it & 1u)) ? reinterpret_cast<int (*)[2]>(dsm + off_wsum2) : ctl.wsum;
                auto flush_prev = [&]() {                     // every warp's partial of frame f-1 is in (a barrier since)
                    long long s1 = 0, s2 = 0;
#pragma unroll
                    for (int w = 0; w < kTW; ++w) { s1 += wsum_prev[w][0]; s2 += (unsigned)wsum_prev[w][1]; }
                    atomicAdd((unsigned long long *)&a.sum_lap[(size_t)(f - 1) * a.V + v], (unsigned long long)s1);
                    atomicAdd((unsigned long long *)&a.sum_lap2[(size_t)(f - 1) * a.V + v], (unsigned long long)s2);
                };
                if (SCORES && !RAW && tid == 0 && f > f_begin) flush_prev();
                const uint32_t tile_off = (RAW && (it & 1u)) ? off_pairs : kOffTile;   // RAW: two packed tiles, alternating
                uint32_t obase = tile_off + my_word;
                if (!(RAW && staged)) asm volatile("" : "+r"(obase));
                uint32_t *const orow = reinterpret_cast<uint32_t *>(dsm + obase);
                if (staged) {
                    if (!RAW && f + 2 < f_end && tid == kTT - 32) issue_box(f + 2, rb);  // this frame's raw rows are consumed
                  if constexpr (RAW) {
                    uint2 X = make_uint2(0u, 0u), Y = make_uint2(0u, 0u), X2 = make_uint2(0u, 0u), Y2 = make_uint2(0u, 0u);
                    const uint32_t up = rb * raw_bytes, lw = up + bw;        // this frame's buffer: upper / lower record of pixel k at pa[k] + up / + lw
                    uint32_t tp1 = 0u, tp2 = 0u;                              // transposed tiles: pixels 1 and 2 of this thread's four
                    auto emit = [&](int k, const uint2 &U, const uint2 &V) {
                        const uint32_t pxl = combine_pairs(U, V, pb[k], pc[k]);
                        if constexpr (TRANSPOSABLE) {
                            if (tr) {                                         // uniform.  The thread's pixels are 12 consecutive bytes of tile row `lane` (rows of 48)
                                const uint32_t tw = ot - my_word + (uint32_t)(lane * (TH * 3) + warp * (kTP * 3));
                                if (k == 1) tp1 = pxl;                        // (walk order: 1, 2, 0, 3)
                                if (k == 2) { tp2 = pxl; asm volatile("st.shared.u32 [%0], %1;" ::"r"(tw + 4u), "r"(__byte_perm(tp1, pxl, 0x5421)) : "memory"); }
                                if (k == 0) asm volatile("st.shared.u32 [%0], %1;" ::"r"(tw), "r"(__byte_perm(pxl, tp1, 0x4210)) : "memory");
                                if (k == 3) asm volatile("st.shared.u32 [%0], %1;" ::"r"(tw + 8u), "r"(__byte_perm(tp2, pxl, 0x6542)) : "memory");
                                return;
                            }
                        }
                        const uint32_t nb = __shfl_down_sync(0xffffffffu, pxl, 1);
                        if (q != 3) asm volatile("st.shared.u32 [%0], %1;" ::"r"(ot + (uint32_t)(k * kTile * 3)), "r"(__byte_perm(pxl, nb, pack_sel)) : "memory");
                        if (SCORES) gt.g[warp * kTP + k + 1][lane] = (uint8_t)gray_of(pxl);
                    };
                    auto st = [&](auto kc, auto ac, uint2 &P, uint2 &Q, bool up_walk) {
                        constexpr int K = decltype(kc)::value;
                        constexpr bool A = decltype(ac)::value;
                        const uint32_t slo = COLSEL ? sel_lo : psel[K];                       // (PRMT reads bits 0-15 of its selector)
                        const uint32_t shi = COLSEL ? sel_hi : (psel[K] >> 16);
                        const uint32_t msk = COLSEL ? m3 : (uint32_t)((int)psel[K] >> 31);
                        if (up_walk) step_raw<K, A>(P, Q, pa[K] + lw, pa[K] + up, slo, shi, pm, msk);        // upwards: roles swapped
                        else step_raw<K, A>(P, Q, pa[K] + up, pa[K] + lw, slo, shi, pm, msk);
                    };
                    auto walk = [&](auto reload) {
                        using KA = decltype(reload);
                        st(std::integral_constant<int, HC - 1>{}, std::true_type{}, X, Y, false);          // both records (mode 3)
                        X2 = X; Y2 = Y;
                        st(std::integral_constant<int, HC>{}, KA{}, X2, Y2, false);
                        emit(HC - 1, X, Y); emit(HC, X2, Y2);
                        st(std::integral_constant<int, HC - 2>{}, KA{}, Y, X, true);
                        st(std::integral_constant<int, HC + 1>{}, KA{}, X2, Y2, false);
                        emit(HC - 2, X, Y); emit(HC + 1, X2, Y2);
                        if constexpr (kTP == 8) {
                            st(std::integral_constant<int, 1>{}, KA{}, Y, X, true);
                            st(std::integral_constant<int, 6>{}, KA{}, X2, Y2, false);
                            emit(1, X, Y); emit(6, X2, Y2);
                            st(std::integral_constant<int, 0>{}, KA{}, Y, X, true);
                            st(std::integral_constant<int, 7>{}, KA{}, X2, Y2, false);
                            emit(0, X, Y); emit(7, X2, Y2);
                        }
                    };
                    if (any_a) walk(std::true_type{}); else walk(std::false_type{});
                    if (SCORES && hvalid) {                                  // the halo pixel: both records, its own selectors
                        const uint32_t slo = psel[kTP], shi = psel[kTP] >> 16;
                        const bool a3 = (int)psel[kTP] < 0;
                        const uint32_t *u = reinterpret_cast<const uint32_t *>(dsm + (pa[kTP] + up - dsm_s));
                        const uint32_t *l = reinterpret_cast<const uint32_t *>(dsm + (pa[kTP] + lw - dsm_s));
                        const uint32_t u0 = u[0], u1 = u[1], u2 = a3 ? u[2] : u0, l0 = l[0], l1 = l[1], l2 = a3 ? l[2] : l0;
                        const uint2 h0 = make_uint2(__byte_perm(u0, u1, slo), __byte_perm(u1, u2, shi));
                        const uint2 h1 = make_uint2(__byte_perm(l0, l1, slo), __byte_perm(l1, l2, shi));
                        halo_store_t<TH>(gt, tid, (uint8_t)gray_of(combine_pairs(h0, h1, pb[kTP], pc[kTP])));
                    }
                  } else if constexpr (LZ) {
#pragma unroll
                    for (int k = 0; k < kTP; ++k) {
                        const uint32_t pxl = lanczos_pairs(dsm, pa[k], ppitch, a.lz_k1, a.lz_fix, pb[k]);
                        const uint32_t nb = __shfl_down_sync(0xffffffffu, pxl, 1);
                        if (q != 3) orow[k * (kTile * 3 / 4)] = __byte_perm(pxl, nb, pack_sel);
                        if (SCORES) gt.g[warp * kTP + k + 1][lane] = (uint8_t)gray_of(pxl);
                    }
                    if (SCORES && hvalid) halo_store_t<TH>(gt, tid, (uint8_t)gray_of(lanczos_pairs(dsm, pa[kTP], ppitch, a.lz_k1, a.lz_fix, pb[kTP])));
                  } else {
                    uint2 X = make_uint2(0u, 0u), Y = make_uint2(0u, 0u);          // upper / lower pair record, carried down the column
                    uint2 X2 = make_uint2(0u, 0u), Y2 = make_uint2(0u, 0u);        // second chain (rows 4-7)
                    auto emit = [&](int k, const uint2 &U, const uint2 &V) {
                        const uint32_t pxl = combine_pairs(U, V, pb[k], pc[k]);
                        const uint32_t nb = __shfl_down_sync(0xffffffffu, pxl, 1);
                        if (q != 3) orow[k * (kTile * 3 / 4)] = __byte_perm(pxl, nb, pack_sel);
                        if (SCORES) gt.g[warp * kTP + k + 1][lane] = (uint8_t)gray_of(pxl);
                    };
                    auto walk = [&](auto reload) {
                        constexpr bool A = decltype(reload)::value;
                        step_records<HC - 1>(X, Y, pa[HC - 1], pa2[HC - 1], ph[HC - 1], ph2[HC - 1], pm);          // both records (mode 3)
                        X2 = X; Y2 = Y;
                        step_records<HC, A>(X2, Y2, pa[HC], pa2[HC], ph[HC], ph2[HC], pm);
                        emit(HC - 1, X, Y); emit(HC, X2, Y2);
                        step_records_up<HC - 2, A>(X, Y, pa[HC - 2], pa2[HC - 2], ph[HC - 2], ph2[HC - 2], pm);
                        step_records<HC + 1, A>(X2, Y2, pa[HC + 1], pa2[HC + 1], ph[HC + 1], ph2[HC + 1], pm);
                        emit(HC - 2, X, Y); emit(HC + 1, X2, Y2);
                        if constexpr (kTP == 8) {
                            step_records_up<1, A>(X, Y, pa[1], pa2[1], ph[1], ph2[1], pm); step_records<6, A>(X2, Y2, pa[6], pa2[6], ph[6], ph2[6], pm);
                            emit(1, X, Y); emit(6, X2, Y2);
                            step_records_up<0, A>(X, Y, pa[0], pa2[0], ph[0], ph2[0], pm); step_records<7, A>(X2, Y2, pa[7], pa2[7], ph[7], ph2[7], pm);
                            emit(0, X, Y); emit(7, X2, Y2);
                        }
                    };
                    if (any_a) walk(std::true_type{}); else walk(std::false_type{});
                    if (SCORES && hvalid) {
                        const uint2 h0 = make_uint2(*reinterpret_cast<const uint32_t *>(dsm + (pa[kTP] - dsm_s)),
                                                    *reinterpret_cast<const uint16_t *>(dsm + (ph[kTP] - dsm_s)));
                        const uint2 h1 = make_uint2(*reinterpret_cast<const uint32_t *>(dsm + (pa2[kTP] - dsm_s)),
                                                    *reinterpret_cast<const uint16_t *>(dsm + (ph2[kTP] - dsm_s)));
                        halo_store_t<TH>(gt, tid, (uint8_t)gray_of(combine_pairs(h0, h1, pb[kTP], pc[kTP])));
                    }
                  }
                } else {
                    const uint8_t *fb = a.frames + (size_t)f * frame_bytes;
                    auto direct = [&](int k) -> uint32_t {
                        if constexpr (LZ) {
                            const LanczosSampler::Px d = {pa[k], pb[k]};
                            return generic ? LanczosSampler::gather_generic(fb, d, g, a.lz_tab) : LanczosSampler::gather_fast(fb, d, g, a.lz_tab);
                        } else {
                            const LinearSampler::Px d = {pa[k], pb[k], pc[k]};
                            return generic ? LinearSampler::gather_generic(fb, d, g, nullptr) : LinearSampler::gather_fast(fb, d, g, nullptr);
                        }
                    };
                    bool plain = true;
                    if constexpr (TRANSPOSABLE) {
                        if (tr) {                                             // (see emit above)
                            uint32_t *const trow = reinterpret_cast<uint32_t *>(dsm + tile_off + (uint32_t)(lane * (TH * 3) + warp * (kTP * 3)));
                            const uint32_t p0 = direct(0), p1 = direct(1), p2 = direct(2), p3 = direct(3);
                            trow[0] = __byte_perm(p0, p1, 0x4210);
                            trow[1] = __byte_perm(p1, p2, 0x5421);
                            trow[2] = __byte_perm(p2, p3, 0x6542);
                            plain = false;
                        }
                    }
#pragma unroll
                    for (int k = 0; k < kTP; ++k) {
                        if (!plain) break;
                        const uint32_t pxl = direct(k);
                        const uint32_t nb = __shfl_down_sync(0xffffffffu, pxl, 1);
                        if (q != 3) orow[k * (kTile * 3 / 4)] = __byte_perm(pxl, nb, pack_sel);
                        if (SCORES) gt.g[warp * kTP + k + 1][lane] = (uint8_t)gray_of(pxl);
                    }
                    if (SCORES && hvalid) halo_store_t<TH>(gt, tid, (uint8_t)gray_of(direct(kTP)));
                }
                // RAW: the store of the previous frame has read the OTHER tile buffer (the next frame's) before anyone gets there
                // (every thread executes the wait — only thread 0 has bulk groups — which saves the branch around it)
                if (RAW) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                if (a.store_ok) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncthreads();                      // tile (+ gray halo) complete; pairs / raw rows free for the next frame
                if constexpr (RAW) ot ^= ot_toggle;
                if (SCORES && RAW && tid == 0 && f > f_begin) flush_prev();

                if (a.store_ok) {
                    if (tid == 0) {
                        if (RAW && staged && f + NB < f_end) issue_box(f + NB, rb);
                        if (TRANSPOSABLE && tr) tma_store_tile(&tms.m[kTMapOutT], dsm_s + tile_off, y0 * 3, x0, f * a.V + v);   // real (x, y) = (y0, x0)
                        else tma_store_tile(&tms.m[kTMapOut], RAW ? dsm_s + tile_off : otile_s, x0 * 3, y0, f * a.V + v);
                    }
                } else {
                    if (RAW && staged && f + NB < f_end && tid == 0) issue_box(f + NB, rb);
                    uint8_t *out_view = a.out + ((size_t)f * a.V + v) * view_bytes;
                    if (TRANSPOSABLE && tr) {                                 // whole blocks only: kTile rows of TH pixels at real (y0, x0)
                        for (int i = tid; i < TH * kTile * 3; i += kTT) {
                            const int row = i / (TH * 3), bb = i - row * (TH * 3);
                            out_view[(size_t)(x0 + row) * opitch + (size_t)y0 * 3 + bb] = dsm[tile_off + i];
                        }
                    } else {
                    const int wbytes = min(kTile, g.ow - x0) * 3, hrows = min(TH, g.oh - y0);
                    for (int i = tid; i < TH * kTile * 3; i += kTT) {
                        const int row = i / (kTile * 3), bb = i - row * (kTile * 3);
                        if (row < hrows && bb < wbytes) out_view[(size_t)(y0 + row) * opitch + (size_t)x0 * 3 + bb] = dsm[tile_off + i];
                    }
                    }
                }
                if (SCORES) {
                    int pl = 0; unsigned pl2 = 0;
#pragma unroll
                    for (int h = 0; h < 8 * TH / kTT; ++h) {
                        int l1; unsigned l2;
                        laplacian_partial_t<TH>(gt, g, x0, y0, full, tid + h * kTT, l1, l2);
                        pl += l1; pl2 += l2;
                    }
                    pl = __reduce_add_sync(0xffffffffu, pl);
                    pl2 = __reduce_add_sync(0xffffffffu, pl2);
                    if (lane == 0) {
                        wsum[warp][0] = pl;                   // |sum L| <= 1020 * 256, sum L^2 <= 1020^2 * 256 < 2^32
                        wsum[warp][1] = (int)pl2;
                    }
                }
            }
            __syncthreads();                           // every warp's partial of the last frame is in
            if (SCORES && tid == 0 && f_lo < f_end) {
                long long s1 = 0, s2 = 0;
                int (*const wsum_last)[2] = (RAW && !(it & 1u)) ? reinterpret_cast<int (*)[2]>(dsm + off_wsum2) : ctl.wsum;   // `it` is past the last frame
#pragma unroll
                for (int w = 0; w < kTW; ++w) { s1 += wsum_last[w][0]; s2 += (unsigned)wsum_last[w][1]; }
                atomicAdd((unsigned long long *)&a.sum_lap[(size_t)(f_end - 1) * a.V + v], (unsigned long long)s1);
                atomicAdd((unsigned long long *)&a.sum_lap2[(size_t)(f_end - 1) * a.V + v], (unsigned long long)s2);
            }
        }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

}  // namespace x360

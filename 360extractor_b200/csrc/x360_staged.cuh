// x360_staged.cuh — bilinear extraction with the tile's source footprint staged in shared memory by TMA.
//
// Same contract as extract_direct<LinearSampler> (src/core/processor.py:327-342 for a frame batch,
// map of src/core/geometry.py:141-190 on the fly), different data movement:
//
//   per tile (once):  fp64 map -> Q5 coordinates; block-wide bounding box of the 2x2 footprints;
//                     every per-frame address (transform, gather, store) is hoisted here.
//   per frame:
//     1. the footprint arrives in `raw` as a few 3-D TMA boxes (cp.async.bulk.tensor -> UTMALDG:
//        BW bytes x 8 rows x 1 frame of the uint8 [F][H][3W] frame tensor, BW in {128..256},
//        starting at a 16-byte aligned byte column — measured: an unaligned inner coordinate faults),
//        issued by ONE thread and completing on an mbarrier; the boxes of frame f+1 are issued as
//        soon as frame f's rows have been consumed, so they fly during the gather and the store;
//     2. a transform pass rewrites the packed BGR rows into "pair" records, 8 B per source pixel x:
//            [B(x) B(x+1) G(x) G(x+1)] [R(x) R(x+1) . .]
//        (4 pixels per thread: 4 LDS.32 + 8 PRMT + 2 STS.128), which removes every alignment
//        case from the per-output-pixel work;
//     3. each output pixel is 2 LDS.64 (upper / lower source row) + 6 dp4a (horizontal pass,
//        weights 2(32-fx), 2fx) + 6 IMAD (vertical pass, weights 32(32-fy), 32fy, bias 2^15):
//        64 V + 2^15 puts dst = (V + 512) >> 10 in byte 2 of each accumulator;
//     4. the finished tile is packed BGR in shared memory and leaves as 128-bit row stores
//        (24 lanes of each warp store the 4 rows that warp produced); gray + Laplacian sums as in
//        the direct kernel.
//
// Tiles whose footprint does not fit the staging buffers (poles), wraps around the seam, or whose
// geometry defeats the TMA alignment rules fall back, tile by tile inside the same launch, to the
// aligned-word / byte gathers of LinearSampler.
#pragma once
#include <climits>
#include <cuda.h>

#include "x360_common.cuh"
#include "x360_linear.cuh"

namespace x360 {

constexpr int kBoxRows = 8;          // rows per TMA box
constexpr int kNumBoxWidths = 5;     // box widths 128, 160, 192, 224, 256 bytes
constexpr int kMaxBoxBytes = 256;

struct TmapSet {
    CUtensorMap m[kNumBoxWidths];    // uint8 [F][H][3W], box {128 + 32 k, kBoxRows, 1}
};

// ---- mbarrier / TMA PTX ----------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}

// one box {BW, kBoxRows, 1} at byte column x, row y, frame z -> dense [kBoxRows][BW] in shared memory
__device__ __forceinline__ void tma_load_box(uint32_t dst_smem, const CUtensorMap *tm, int x, int y, int z, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst_smem), "l"(tm), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
                 : "memory");
}

struct StageCtrl {
    unsigned long long mbar;
    long long sums[2];
    int bb[4];          // min ix, max ix, min iy, max iy over the tile's pixels (+ halo)
    int special;        // a wrapped / frame-end footprint exists (direct mode must go byte-wise)
    int pad;
};

// dynamic shared memory: raw | pairs | packed BGR tile | gray tile | ctrl
__host__ __device__ inline size_t staged_raw_bytes(int rows_max) { return (size_t)((rows_max + kBoxRows - 1) / kBoxRows) * kBoxRows * kMaxBoxBytes; }
__host__ __device__ inline size_t staged_smem_bytes(int rows_max, int pp)
{
    return staged_raw_bytes(rows_max) + (size_t)rows_max * pp * 8 + kTile * kTile * 3 + sizeof(GrayTile) + sizeof(StageCtrl);
}

// per-pixel state, one register triple reused by both modes
struct PxState {
    uint32_t a, b, c;   // staged: pair address (shared window), wx2 = 2(32-fx) | 2fx << 8, wy0s = 32 (32-fy)
                        // direct: LinearSampler::Px {off, wx, wy}
};

__device__ __forceinline__ uint2 lds64(uint32_t addr)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}

__device__ __forceinline__ uint32_t sample_staged(const PxState &p, uint32_t row_bytes)
{
    const uint2 r0 = lds64(p.a);
    const uint2 r1 = lds64(p.a + row_bytes);
    const uint32_t wlo = p.b, whi = p.b << 16;
    const uint32_t hb0 = dp4a_uu(r0.x, wlo, 0), hg0 = dp4a_uu(r0.x, whi, 0), hr0 = dp4a_uu(r0.y, wlo, 0);
    const uint32_t hb1 = dp4a_uu(r1.x, wlo, 0), hg1 = dp4a_uu(r1.x, whi, 0), hr1 = dp4a_uu(r1.y, wlo, 0);
    const uint32_t wy0 = p.c, wy1 = 1024u - p.c;
    const uint32_t vb = hb0 * wy0 + hb1 * wy1 + 32768u;
    const uint32_t vg = hg0 * wy0 + hg1 * wy1 + 32768u;
    const uint32_t vr = hr0 * wy0 + hr1 * wy1 + 32768u;
    return __byte_perm(__byte_perm(vb, vg, 0x0062), vr, 0x7610);       // [vb.b2, vg.b2, vr.b2, 0]
}

template <bool SCORES, int MINB>
__global__ void __launch_bounds__(kThreads, MINB)
extract_linear_staged(const __grid_constant__ ExtractArgs a, const __grid_constant__ TmapSet tms)
{
    extern __shared__ __align__(128) uint8_t dsm[];
    const Geom &g = a.g;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t pitch = (uint32_t)g.W * 3u;
    const size_t frame_bytes = (size_t)g.H * pitch;
    const size_t view_bytes = (size_t)g.oh * g.ow * 3;
    const size_t opitch = (size_t)g.ow * 3;
    const size_t out_frame_stride = view_bytes * a.V;
    const uint32_t pair_row = (uint32_t)a.pp * 8u;

    uint8_t *raw = dsm;
    uint8_t *pairs = raw + staged_raw_bytes(a.rows_max);
    uint8_t *otile = pairs + (size_t)a.rows_max * pair_row;            // packed BGR [kTile][96]
    GrayTile &gt = *reinterpret_cast<GrayTile *>(otile + kTile * kTile * 3);
    StageCtrl &ctl = *reinterpret_cast<StageCtrl *>(reinterpret_cast<uint8_t *>(&gt) + sizeof(GrayTile));
    const uint32_t raw_s = smem_u32(raw), pairs_s = smem_u32(pairs);

    if (tid == 0) mbar_init(&ctl.mbar, 1);
    uint32_t phase = 0;
    __syncthreads();

    // frame-invariant addresses of this thread's share of the tile
    uint8_t *const my_otile = otile + (warp * kSlots) * (kTile * 3) + lane * 3;                 // slot k: + k * 96
    const int st_row = lane / 6, st_chunk = lane - st_row * 6;                                  // store: 24 lanes, 16 B each
    const uint8_t *const st_src = otile + (warp * kSlots + st_row) * (kTile * 3) + st_chunk * 16;

    for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
        int v, x0, y0;
        tile_coords(g, t, v, x0, y0);
        const double *R = a.R + 9 * v;
        const bool full = (x0 + kTile <= g.ow) && (y0 + kTile <= g.oh);
        const bool vec_store = full && a.aligned;

        // ---- 1. map of this tile (once for the whole batch) + footprint bounding box -------
        constexpr int NS = SCORES ? kSlots + 1 : kSlots;     // slot kSlots = this thread's halo pixel
        int sxs[kSlots + 1], sys[kSlots + 1];
        bool hvalid = false;
        {
            const int x = min(x0 + lane, g.ow - 1);
#pragma unroll
            for (int k = 0; k < kSlots; ++k) map_q5(R, g, x, min(y0 + warp * kSlots + k, g.oh - 1), sxs[k], sys[k]);
            sxs[kSlots] = sxs[0];
            sys[kSlots] = sys[0];
            if (SCORES && tid < 128) {
                int hx, hy;
                halo_coords(tid, x0, y0, hx, hy);
                hvalid = (hx >= 0 && hx < g.ow && hy >= 0 && hy < g.oh);
                if (hvalid) map_q5(R, g, hx, hy, sxs[kSlots], sys[kSlots]);
            }
        }
        int xmin = INT_MAX, xmax = INT_MIN, ymin = INT_MAX, ymax = INT_MIN;
        bool special = false;
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            const int ix = sat_s16(sxs[k] >> 5), iy = sat_s16(sys[k] >> 5);
            xmin = min(xmin, ix); xmax = max(xmax, ix);
            ymin = min(ymin, iy); ymax = max(ymax, iy);
            const int wx0 = wrap_idx(ix, g.W), wy0 = wrap_idx(iy, g.H);
            const uint32_t off = (uint32_t)(wy0 * g.W + wx0) * 3u;
            special |= (wx0 == g.W - 1) | (wy0 == g.H - 1) | ((off & ~3u) + pitch + 12u > (uint32_t)g.H * pitch);
        }
        if (tid == 0) {
            ctl.bb[0] = INT_MAX; ctl.bb[1] = INT_MIN; ctl.bb[2] = INT_MAX; ctl.bb[3] = INT_MIN;
            ctl.special = 0; ctl.sums[0] = 0; ctl.sums[1] = 0;
        }
        __syncthreads();
        xmin = __reduce_min_sync(0xffffffffu, xmin); xmax = __reduce_max_sync(0xffffffffu, xmax);
        ymin = __reduce_min_sync(0xffffffffu, ymin); ymax = __reduce_max_sync(0xffffffffu, ymax);
        special = __any_sync(0xffffffffu, special);
        if (lane == 0) {
            atomicMin(&ctl.bb[0], xmin); atomicMax(&ctl.bb[1], xmax);
            atomicMin(&ctl.bb[2], ymin); atomicMax(&ctl.bb[3], ymax);
            if (special) ctl.special = 1;
        }
        __syncthreads();
        xmin = ctl.bb[0]; xmax = ctl.bb[1]; ymin = ctl.bb[2]; ymax = ctl.bb[3];
        const bool generic = (ctl.special != 0) || !a.aligned;

        // footprint: pair records for x in [x_org, xmax], source rows [ymin, ymax + 1]
        const int x_org = xmin & ~3, y_org = ymin;
        const int npx = xmax - x_org + 1, rows = ymax + 2 - ymin;
        const int ngroups = (npx + 3) >> 2;
        const int gx0 = (3 * x_org) & ~15;                               // TMA needs a 16-byte aligned byte column
        const int delta = 3 * x_org - gx0;                               // 0, 4, 8 or 12
        const int need = delta + 3 * npx + 3;                            // bytes gx0 .. 3 xmax + 5
        const int wsel = max(0, (need + 31) / 32 - 4);                   // box width 128 + 32 wsel
        const uint32_t bw = 128u + 32u * wsel;
        const int nboxes = (rows + kBoxRows - 1) / kBoxRows;
        const bool staged = a.stage_ok && xmin >= 0 && ymin >= 0 && xmax <= g.W - 2 && ymax <= g.H - 2 &&
                            npx <= a.px_max && rows <= a.rows_max && wsel < kNumBoxWidths;

        // ---- 2. per-pixel gather state and hoisted addresses ----------------------------------
        PxState px[kSlots + 1];
        px[kSlots].a = px[kSlots].b = px[kSlots].c = 0;
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            if (staged) {
                const int fx = sxs[k] & 31, fy = sys[k] & 31;
                px[k].a = pairs_s + (uint32_t)((sys[k] >> 5) - y_org) * pair_row + (uint32_t)((sxs[k] >> 5) - x_org) * 8u;
                px[k].b = (uint32_t)(2 * (32 - fx)) | ((uint32_t)(2 * fx) << 8);
                px[k].c = (uint32_t)(32 * (32 - fy));
            } else {
                bool dummy = false;
                const LinearSampler::Px d = LinearSampler::make_px(sxs[k], sys[k], g, nullptr, dummy);
                px[k].a = d.off; px[k].b = d.wx; px[k].c = d.wy;
            }
        }
        if (a.tile_modes && tid == 0) atomicAdd(&a.tile_modes[staged ? 0 : (generic ? 2 : 1)], 1u);

        // transform: lanes run along the 4-pixel groups of a row, 16 wide when that covers the row
        const int gw = (ngroups <= 16) ? 16 : 32;
        const int gxi = tid & (gw - 1), tr0 = tid / gw, trstep = kThreads / gw;
        const bool tr_active = staged && gxi < ngroups;
        const uint32_t tr_src0 = raw_s + (uint32_t)tr0 * bw + (uint32_t)delta + 12u * gxi, tr_src_step = (uint32_t)trstep * bw;
        const uint32_t tr_dst0 = pairs_s + (uint32_t)tr0 * pair_row + 32u * gxi, tr_dst_step = (uint32_t)trstep * pair_row;
        // store: this lane's 16-byte chunk of one of the warp's 4 rows, frame 0
        uint8_t *st_dst = a.out + (size_t)v * view_bytes + (size_t)(y0 + warp * kSlots + st_row) * opitch + (size_t)x0 * 3 + st_chunk * 16;
        const uint8_t *fb = a.frames;
        const CUtensorMap *tm = &tms.m[staged ? wsel : 0];

        auto issue_boxes = [&](int f) {              // one thread
            mbar_arrive_expect_tx(&ctl.mbar, (uint32_t)nboxes * kBoxRows * bw);
            for (int b = 0; b < nboxes; ++b)
                tma_load_box(raw_s + (uint32_t)b * kBoxRows * bw, tm, gx0, y_org + b * kBoxRows, f, &ctl.mbar);
        };
        if (staged && tid == 0) issue_boxes(0);

        // ---- 3. walk the frame batch ---------------------------------------------------------
        for (int f = 0; f < a.F; ++f, fb += frame_bytes, st_dst += out_frame_stride) {
            if (staged) {
                mbar_wait(&ctl.mbar, phase);
                phase ^= 1u;
                if (tr_active) {
                    uint32_t src = tr_src0, dst = tr_dst0;
                    for (int r = tr0; r < rows; r += trstep, src += tr_src_step, dst += tr_dst_step) {
                        uint32_t wa, wb, wc, wd;
                        asm volatile("ld.shared.u32 %0, [%4];\n\tld.shared.u32 %1, [%4+4];\n\tld.shared.u32 %2, [%4+8];\n\tld.shared.u32 %3, [%4+12];"
                                     : "=&r"(wa), "=&r"(wb), "=&r"(wc), "=&r"(wd) : "r"(src));
                        const uint32_t o0 = __byte_perm(wa, wb, 0x4130), o1 = __byte_perm(wa, wb, 0x0052);
                        const uint32_t o2 = __byte_perm(wa, wb, 0x7463), o3 = __byte_perm(wb, wc, 0x0041);
                        const uint32_t o4 = __byte_perm(wb, wc, 0x6352), o5 = __byte_perm(wc, wc, 0x0030);
                        const uint32_t o6 = __byte_perm(wc, wd, 0x5241), o7 = __byte_perm(wc, wd, 0x0063);
                        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};\n\tst.shared.v4.u32 [%0+16], {%5, %6, %7, %8};"
                                     ::"r"(dst), "r"(o0), "r"(o1), "r"(o2), "r"(o3), "r"(o4), "r"(o5), "r"(o6), "r"(o7) : "memory");
                    }
                }
            }
            __syncthreads();                      // pairs ready, raw consumed, previous store phase done
            if (SCORES && tid == 0 && f > 0) {    // every warp's partial of frame f-1 is in: flush it
                atomicAdd((unsigned long long *)&a.sum_lap[(size_t)(f - 1) * a.V + v], (unsigned long long)ctl.sums[0]);
                atomicAdd((unsigned long long *)&a.sum_lap2[(size_t)(f - 1) * a.V + v], (unsigned long long)ctl.sums[1]);
                ctl.sums[0] = 0; ctl.sums[1] = 0;
            }
            if (staged) {
                // the issuing thread rotates over the warps so the TMA issue is nobody's critical path
                if (f + 1 < a.F && tid == ((f & 7) << 5)) issue_boxes(f + 1);
#pragma unroll
                for (int k = 0; k < kSlots; ++k) {
                    const uint32_t pxl = sample_staged(px[k], pair_row);
                    uint8_t *o = my_otile + k * (kTile * 3);
                    o[0] = (uint8_t)pxl; o[1] = (uint8_t)(pxl >> 8); o[2] = (uint8_t)(pxl >> 16);
                    if (SCORES) gt.g[warp * kSlots + k + 1][lane] = (uint8_t)gray_of(pxl);
                }
                if (SCORES && hvalid) halo_store(gt, tid, (uint8_t)gray_of(sample_staged(px[kSlots], pair_row)));
            } else {
#pragma unroll 1
                for (int k = 0; k < kSlots; ++k) {
                    const LinearSampler::Px d = {px[k].a, px[k].b, px[k].c};
                    const uint32_t pxl = generic ? LinearSampler::gather_generic(fb, d, g, nullptr)
                                                 : LinearSampler::gather_fast(fb, d, g, nullptr);
                    uint8_t *o = my_otile + k * (kTile * 3);
                    o[0] = (uint8_t)pxl; o[1] = (uint8_t)(pxl >> 8); o[2] = (uint8_t)(pxl >> 16);
                    if (SCORES) gt.g[warp * kSlots + k + 1][lane] = (uint8_t)gray_of(pxl);
                }
                if (SCORES && hvalid) {
                    const LinearSampler::Px d = {px[kSlots].a, px[kSlots].b, px[kSlots].c};
                    halo_store(gt, tid, (uint8_t)gray_of(generic ? LinearSampler::gather_generic(fb, d, g, nullptr)
                                                                  : LinearSampler::gather_fast(fb, d, g, nullptr)));
                }
            }
            // Tile (+ gray halo) complete; also: every warp is done reading `pairs`, which the next
            // frame's transform overwrites.
            __syncthreads();

            // ---- store: packed BGR rows, 128-bit ------------------------------------------------
            if (vec_store) {
                if (lane < 24) *reinterpret_cast<uint4 *>(st_dst) = *reinterpret_cast<const uint4 *>(st_src);
            } else {
                const int wbytes = min(kTile, g.ow - x0) * 3, hrows = min(kTile, g.oh - y0);
                uint8_t *out_view = a.out + (size_t)f * out_frame_stride + (size_t)v * view_bytes;
                for (int i = tid; i < kTile * kTile * 3; i += kThreads) {
                    const int row = i / (kTile * 3), bb = i - row * (kTile * 3);
                    if (row < hrows && bb < wbytes) out_view[(size_t)(y0 + row) * opitch + (size_t)x0 * 3 + bb] = otile[i];
                }
            }
            if (SCORES) {
                int pl; unsigned pl2;
                laplacian_partial(gt, g, x0, y0, full, tid, pl, pl2);
                pl = __reduce_add_sync(0xffffffffu, pl);
                pl2 = __reduce_add_sync(0xffffffffu, pl2);
                if (lane == 0) {
                    atomicAdd((unsigned long long *)&ctl.sums[0], (unsigned long long)(long long)pl);
                    atomicAdd((unsigned long long *)&ctl.sums[1], (unsigned long long)pl2);
                }
            }
        }
        __syncthreads();                           // the tile's shared state is reused by the next tile
        if (SCORES && tid == 0) {                  // last frame's sums
            atomicAdd((unsigned long long *)&a.sum_lap[(size_t)(a.F - 1) * a.V + v], (unsigned long long)ctl.sums[0]);
            atomicAdd((unsigned long long *)&a.sum_lap2[(size_t)(a.F - 1) * a.V + v], (unsigned long long)ctl.sums[1]);
        }
    }
}

}  // namespace x360

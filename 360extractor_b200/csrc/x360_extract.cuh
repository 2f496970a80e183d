// x360_extract.cuh — the batched extraction kernel (direct path): map + gather (+ blur score).
//
// Replaces, for a whole frame batch in one launch, the body of the reference's per-view loop
//   src/core/processor.py:327-342
//       rect_img = cv2.remap(frame, map_x, map_y, interp_flag, borderMode=cv2.BORDER_WRAP)
//       score    = ImageUtils.calculate_blur_score(rect_img)
// with the map of src/core/geometry.py:141-190 computed on the fly per output tile.
//
// Work decomposition: one CTA owns one 32x32 output tile of one view at a time and walks ALL
// frames of the batch with it, so the fp64 map of the tile is evaluated once and kept in
// registers as the sampler's per-pixel state — it is never stored in HBM.  Lanes run along x so
// one warp-level load touches one or two source lines.  The finished tile goes through shared
// memory: BGR bytes leave as 128-bit row stores, gray bytes (with a 1-px halo recomputed by
// warps 0-3) feed the fused Laplacian sums, reduced per warp (redux) -> per CTA (shared atomics)
// -> one global int64 atomic pair per tile and frame.  Integer sums: order-independent, exact.
#pragma once
#include "x360_common.cuh"

namespace x360 {

template <class S, bool SCORES, int MINB>
__global__ void __launch_bounds__(kThreads, MINB)
extract_direct(const __grid_constant__ ExtractArgs a)
{
    __shared__ __align__(16) TileSmem s;
    const Geom &g = a.g;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t frame_bytes = (size_t)g.H * g.W * 3;
    const size_t view_bytes = (size_t)g.oh * g.ow * 3;

    for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
        int v, x0, y0;
        tile_coords(g, t, v, x0, y0);
        const double *R = a.R + 9 * v;
        const bool full = (x0 + kTile <= g.ow) && (y0 + kTile <= g.oh);
        const bool vec_store = full && a.aligned;

        // ---- map of this tile, once for the whole batch --------------------------------
        typename S::Px px[kSlots];
        typename S::Px hpx = {};
        bool special = false, hvalid = false;
        const int x = min(x0 + lane, g.ow - 1);          // clamp: out-of-image lanes stay in bounds
#pragma unroll
        for (int k = 0; k < kSlots; ++k) {
            const int y = min(y0 + warp * kSlots + k, g.oh - 1);
            int sx, sy;
            map_q5(R, g, x, y, sx, sy);
            px[k] = S::make_px(sx, sy, g, a.lz_tab, special);
        }
        if (SCORES && tid < 128) {
            int hx, hy;
            halo_coords(tid, x0, y0, hx, hy);
            hvalid = (hx >= 0 && hx < g.ow && hy >= 0 && hy < g.oh);
            if (hvalid) {
                int sx, sy;
                map_q5(R, g, hx, hy, sx, sy);
                hpx = S::make_px(sx, sy, g, a.lz_tab, special);
            }
        }
        if (tid == 0) { s.special = 0; s.sums[0][0] = s.sums[0][1] = s.sums[1][0] = s.sums[1][1] = 0; }
        __syncthreads();
        if (special || !a.aligned) s.special = 1;         // benign race: every writer stores 1
        __syncthreads();
        const bool generic = (s.special != 0);

        // ---- walk the frame batch ------------------------------------------------------
        for (int f = 0; f < a.F; ++f) {
            const int b = f & 1;
            const uint8_t *fb = a.frames + (size_t)f * frame_bytes;
            if (!generic) {
#pragma unroll
                for (int k = 0; k < kSlots; ++k) {
                    const uint32_t pxl = S::gather_fast(fb, px[k], g, a.lz_tab);
                    uint8_t *o = &s.out[b][warp * kSlots + k][lane * 3];
                    o[0] = (uint8_t)pxl; o[1] = (uint8_t)(pxl >> 8); o[2] = (uint8_t)(pxl >> 16);
                    if (SCORES) s.gt[b].g[warp * kSlots + k + 1][lane] = (uint8_t)gray_of(pxl);
                }
                if (SCORES && hvalid) halo_store(s.gt[b], tid, (uint8_t)gray_of(S::gather_fast(fb, hpx, g, a.lz_tab)));
            } else {
#pragma unroll 1
                for (int k = 0; k < kSlots; ++k) {
                    const uint32_t pxl = S::gather_generic(fb, px[k], g, a.lz_tab);
                    uint8_t *o = &s.out[b][warp * kSlots + k][lane * 3];
                    o[0] = (uint8_t)pxl; o[1] = (uint8_t)(pxl >> 8); o[2] = (uint8_t)(pxl >> 16);
                    if (SCORES) s.gt[b].g[warp * kSlots + k + 1][lane] = (uint8_t)gray_of(pxl);
                }
                if (SCORES && hvalid) halo_store(s.gt[b], tid, (uint8_t)gray_of(S::gather_generic(fb, hpx, g, a.lz_tab)));
            }
            __syncthreads();
            // Everything below reads buffer b only; the next iteration writes buffer b^1, whose
            // readers all passed this barrier one iteration ago.
            uint8_t *out_view = a.out + ((size_t)f * a.V + v) * view_bytes;
            store_tile(s, b, out_view, g, x0, y0, vec_store, tid);
            if (SCORES) {
                if (tid == 0 && f > 0) {                  // flush the previous frame's tile sums
                    atomicAdd((unsigned long long *)&a.sum_lap[(size_t)(f - 1) * a.V + v], (unsigned long long)s.sums[b ^ 1][0]);
                    atomicAdd((unsigned long long *)&a.sum_lap2[(size_t)(f - 1) * a.V + v], (unsigned long long)s.sums[b ^ 1][1]);
                    s.sums[b ^ 1][0] = 0; s.sums[b ^ 1][1] = 0;
                }
                int pl; unsigned pl2;
                laplacian_partial(s.gt[b], g, x0, y0, full, tid, pl, pl2);
                pl = __reduce_add_sync(0xffffffffu, pl);
                pl2 = __reduce_add_sync(0xffffffffu, pl2);
                if (lane == 0) {
                    atomicAdd((unsigned long long *)&s.sums[b][0], (unsigned long long)(long long)pl);
                    atomicAdd((unsigned long long *)&s.sums[b][1], (unsigned long long)pl2);
                }
            }
        }
        __syncthreads();
        if (SCORES && tid == 0) {                          // last frame's sums
            const int b = (a.F - 1) & 1;
            atomicAdd((unsigned long long *)&a.sum_lap[(size_t)(a.F - 1) * a.V + v], (unsigned long long)s.sums[b][0]);
            atomicAdd((unsigned long long *)&a.sum_lap2[(size_t)(a.F - 1) * a.V + v], (unsigned long long)s.sums[b][1]);
        }
        __syncthreads();
    }
}

}  // namespace x360

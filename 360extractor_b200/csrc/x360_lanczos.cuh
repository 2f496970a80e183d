// x360_lanczos.cuh — Lanczos4 sampler: cv2.remap(..., INTER_LANCZOS4, BORDER_WRAP) in fixed point.
//
// Reference call site: src/core/processor.py:334 with interp_flag = cv2.INTER_LANCZOS4
// (src/core/processor.py:199-200, interpolation_mode == 'lanczos').  8x8 taps at
// (iy-3 .. iy+4, ix-3 .. ix+4), wrap-around in both axes, int16 Q15 weights from the
// [32*32][8][8] table (built on the host, x360_tables.hpp), dst = sat_u8((sum + 2^14) >> 15).
#pragma once
#include "x360_common.cuh"

namespace x360 {

struct LanczosSampler {
    struct Px {
        uint32_t off;    // byte offset of tap (y0, x0) = wrapped (iy-3, ix-3); low 2 bits = alignment
        uint32_t info;   // fy*32+fx | xwrap << 16 | ywrap << 17
    };

    static __device__ __forceinline__ Px make_px(int sx, int sy, const Geom &g, const int16_t *, bool &special)
    {
        const int fx = sx & 31, fy = sy & 31;
        const int x0 = wrap_idx(sat_s16(sx >> 5) - 3, g.W);
        const int y0 = wrap_idx(sat_s16(sy >> 5) - 3, g.H);
        const bool xw = (x0 + 7 >= g.W), yw = (y0 + 7 >= g.H);
        Px p;
        p.off = (uint32_t)(y0 * g.W + x0) * 3u;
        const uint32_t pitch = (uint32_t)g.W * 3u;
        // the fast path reads 28 bytes per row from the aligned base, 8 rows
        special |= (xw | yw) | ((p.off & ~3u) + 7u * pitch + 28u > (uint32_t)g.H * pitch);
        p.info = (uint32_t)(fy * 32 + fx) | ((uint32_t)xw << 16) | ((uint32_t)yw << 17);
        return p;
    }

    static __device__ __forceinline__ uint32_t finish(int ab, int ag, int ar)
    {
        const int b = min(255, max(0, (ab + 16384) >> 15));
        const int gg = min(255, max(0, (ag + 16384) >> 15));
        const int r = min(255, max(0, (ar + 16384) >> 15));
        return (uint32_t)b | ((uint32_t)gg << 8) | ((uint32_t)r << 16);
    }

    // Fast path: per source row, 7 aligned words -> 24 normalised bytes
    //   n0..n5 = [B0 G0 R0 B1 | G1 R1 B2 G2 | R2 B3 G3 R3 | B4 G4 R4 B5 | G5 R5 B6 G6 | R6 B7 G7 R7]
    // -> per channel two byte quads [s0..s3], [s4..s7] -> four dp2a (int16 pair x byte pair) each.
    static __device__ __forceinline__ uint32_t gather_fast(const uint8_t *__restrict__ fb, const Px &p,
                                                           const Geom &g, const int16_t *__restrict__ tab)
    {
        const uint32_t pitch = (uint32_t)g.W * 3u;
        const uint8_t *row = fb + (p.off & ~3u);
        const uint4 *wrow = reinterpret_cast<const uint4 *>(tab + (size_t)(p.info & 1023u) * 64);
        const uint32_t sh = p.off << 3;
        int ab = 0, ag = 0, ar = 0;
#pragma unroll
        for (int r = 0; r < 8; ++r, row += pitch) {
            const uint32_t a0 = ldg_u32(row), a1 = ldg_u32(row + 4), a2 = ldg_u32(row + 8), a3 = ldg_u32(row + 12);
            const uint32_t a4 = ldg_u32(row + 16), a5 = ldg_u32(row + 20), a6 = ldg_u32(row + 24);
            const uint4 w = __ldg(wrow + r);        // (w0|w1<<16, w2|w3<<16, w4|w5<<16, w6|w7<<16)
            const uint32_t n0 = __funnelshift_r(a0, a1, sh), n1 = __funnelshift_r(a1, a2, sh);
            const uint32_t n2 = __funnelshift_r(a2, a3, sh), n3 = __funnelshift_r(a3, a4, sh);
            const uint32_t n4 = __funnelshift_r(a4, a5, sh), n5 = __funnelshift_r(a5, a6, sh);
            // B: bytes 0,3,6,9 | 12,15,18,21
            const uint32_t bl = __byte_perm(__byte_perm(n0, n1, 0x0630), n2, 0x5210);
            const uint32_t bh = __byte_perm(__byte_perm(n3, n4, 0x0630), n5, 0x5210);
            // G: bytes 1,4,7,10 | 13,16,19,22
            const uint32_t gl = __byte_perm(__byte_perm(n0, n1, 0x0741), n2, 0x6210);
            const uint32_t gh = __byte_perm(__byte_perm(n3, n4, 0x0741), n5, 0x6210);
            // R: bytes 2,5,8,11 | 14,17,20,23
            const uint32_t rl = __byte_perm(__byte_perm(n0, n1, 0x0052), n2, 0x7410);
            const uint32_t rh = __byte_perm(__byte_perm(n3, n4, 0x0052), n5, 0x7410);
            ab = dp2a_hi_su(w.w, bh, dp2a_lo_su(w.z, bh, dp2a_hi_su(w.y, bl, dp2a_lo_su(w.x, bl, ab))));
            ag = dp2a_hi_su(w.w, gh, dp2a_lo_su(w.z, gh, dp2a_hi_su(w.y, gl, dp2a_lo_su(w.x, gl, ag))));
            ar = dp2a_hi_su(w.w, rh, dp2a_lo_su(w.z, rh, dp2a_hi_su(w.y, rl, dp2a_lo_su(w.x, rl, ar))));
        }
        return finish(ab, ag, ar);
    }

    // Generic path: every tap wrapped individually.
    static __device__ __forceinline__ uint32_t gather_generic(const uint8_t *__restrict__ fb, const Px &p,
                                                              const Geom &g, const int16_t *__restrict__ tab)
    {
        const uint32_t pitch = (uint32_t)g.W * 3u;
        const int y0 = (int)(p.off / pitch);
        const int x0 = (int)((p.off - (uint32_t)y0 * pitch) / 3u);
        const int16_t *w = tab + (size_t)(p.info & 1023u) * 64;
        int ab = 0, ag = 0, ar = 0;
        for (int r = 0; r < 8; ++r) {
            const int y = (y0 + r) % g.H;
            const uint8_t *row = fb + (size_t)y * pitch;
            for (int c = 0; c < 8; ++c) {
                const int x = (x0 + c) % g.W;
                const int wt = __ldg(w + r * 8 + c);
                const uint8_t *q = row + x * 3;
                ab += wt * (int)__ldg(q);
                ag += wt * (int)__ldg(q + 1);
                ar += wt * (int)__ldg(q + 2);
            }
        }
        return finish(ab, ag, ar);
    }
};

}  // namespace x360

// x360_linear.cuh — bilinear sampler: cv2.remap(..., INTER_LINEAR, BORDER_WRAP) in fixed point.
//
// Reference call site: src/core/processor.py:334 with interp_flag = cv2.INTER_LINEAR
// (src/core/processor.py:199-200).  Q5 sub-pixel position, Q15 weights:
//   w_rc = (32-fy | fy)(32-fx | fx) * 32,  dst = (sum w s + 2^14) >> 15 = (V + 512) >> 10
// with V = sum (32-fy|fy)(32-fx|fx) s <= 255 * 1024.
#pragma once
#include "x360_common.cuh"

namespace x360 {

struct LinearSampler {
    // frame-invariant per-pixel gather state for the 2x2 footprint
    struct Px {
        uint32_t off;   // byte offset of tap (y0, x0) in the frame; low 2 bits = byte alignment
        uint32_t wx;    // (32-fx) | fx << 8
        uint32_t wy;    // (32-fy) | fy << 8 | xwrap << 16 | ywrap << 17
    };

    // `special` marks a footprint the aligned-word gather must not touch: a wrapped tap, or a
    // footprint so close to the end of the frame that the 12-byte row read would run past it.
    static __device__ __forceinline__ Px make_px(int sx, int sy, const Geom &g, const int16_t *, bool &special)
    {
        const int fx = sx & 31, fy = sy & 31;
        const int x0 = wrap_idx(sat_s16(sx >> 5), g.W);
        const int y0 = wrap_idx(sat_s16(sy >> 5), g.H);
        const bool xw = (x0 == g.W - 1), yw = (y0 == g.H - 1);
        Px p;
        p.off = (uint32_t)(y0 * g.W + x0) * 3u;
        const uint32_t pitch = (uint32_t)g.W * 3u;
        special |= (xw | yw) | ((p.off & ~3u) + pitch + 12u > (uint32_t)g.H * pitch);
        p.wx = (uint32_t)(32 - fx) | ((uint32_t)fx << 8);
        p.wy = (uint32_t)(32 - fy) | ((uint32_t)fy << 8) | ((uint32_t)xw << 16) | ((uint32_t)yw << 17);
        return p;
    }

    // Fixed-point combine of the 2x2 taps.  lo = [B0 G0 R0 B1], hi = [G1 R1 . .] for the upper (0)
    // and lower (1) source row.  Returns the packed pixel [B, G, R, 0].
    static __device__ __forceinline__ uint32_t combine(uint32_t lo0, uint32_t hi0, uint32_t lo1, uint32_t hi1,
                                                       uint32_t wx, uint32_t wy)
    {
        // horizontal pass per row and channel: dp4a([s_c0, s_c1, junk, junk], [32-fx, fx, 0, 0])
        const uint32_t hb0 = dp4a_uu(__byte_perm(lo0, hi0, 0x0030), wx, 0);
        const uint32_t hg0 = dp4a_uu(__byte_perm(lo0, hi0, 0x0041), wx, 0);
        const uint32_t hr0 = dp4a_uu(__byte_perm(lo0, hi0, 0x0052), wx, 0);
        const uint32_t hb1 = dp4a_uu(__byte_perm(lo1, hi1, 0x0030), wx, 0);
        const uint32_t hg1 = dp4a_uu(__byte_perm(lo1, hi1, 0x0041), wx, 0);
        const uint32_t hr1 = dp4a_uu(__byte_perm(lo1, hi1, 0x0052), wx, 0);
        // vertical pass: dp2a([h_row0, h_row1] as 2x16, [32-fy, fy]) + 512
        const uint32_t vb = dp2a_lo_uu(__byte_perm(hb0, hb1, 0x5410), wy, 512u);
        const uint32_t vg = dp2a_lo_uu(__byte_perm(hg0, hg1, 0x5410), wy, 512u);
        const uint32_t vr = dp2a_lo_uu(__byte_perm(hr0, hr1, 0x5410), wy, 512u);
        return (vb >> 10) | ((vg >> 10) << 8) | ((vr >> 10) << 16);
    }

    // Aligned-word gather (no wrap inside the footprint, pitch % 4 == 0, frame base 4-B aligned).
    static __device__ __forceinline__ uint32_t gather_fast(const uint8_t *__restrict__ fb, const Px &p,
                                                           const Geom &g, const int16_t *)
    {
        const uint32_t pitch = (uint32_t)g.W * 3u;
        const uint8_t *r0 = fb + (p.off & ~3u);
        const uint8_t *r1 = r0 + pitch;
        const uint32_t a0 = ldg_u32(r0), a1 = ldg_u32(r0 + 4), a2 = ldg_u32(r0 + 8);
        const uint32_t b0 = ldg_u32(r1), b1 = ldg_u32(r1 + 4), b2 = ldg_u32(r1 + 8);
        const uint32_t sh = p.off << 3;     // funnel shift uses the low 5 bits: (off & 3) * 8
        return combine(__funnelshift_r(a0, a1, sh), __funnelshift_r(a1, a2, sh),
                       __funnelshift_r(b0, b1, sh), __funnelshift_r(b1, b2, sh), p.wx, p.wy);
    }

    // Byte-granular gather with wrap-around in both axes (seam / pole tiles, odd widths).
    static __device__ __forceinline__ uint32_t gather_generic(const uint8_t *__restrict__ fb, const Px &p,
                                                              const Geom &g, const int16_t *)
    {
        const uint32_t pitch = (uint32_t)g.W * 3u;
        const uint32_t o00 = p.off;
        const uint32_t o01 = (p.wy & (1u << 16)) ? o00 - (uint32_t)(g.W - 1) * 3u : o00 + 3u;
        const uint32_t dy = (p.wy & (1u << 17)) ? (uint32_t)0 - (uint32_t)(g.H - 1) * pitch : pitch;
        const uint32_t o10 = o00 + dy, o11 = o01 + dy;
        const uint32_t lo0 = (uint32_t)__ldg(fb + o00) | ((uint32_t)__ldg(fb + o00 + 1) << 8) |
                             ((uint32_t)__ldg(fb + o00 + 2) << 16) | ((uint32_t)__ldg(fb + o01) << 24);
        const uint32_t hi0 = (uint32_t)__ldg(fb + o01 + 1) | ((uint32_t)__ldg(fb + o01 + 2) << 8);
        const uint32_t lo1 = (uint32_t)__ldg(fb + o10) | ((uint32_t)__ldg(fb + o10 + 1) << 8) |
                             ((uint32_t)__ldg(fb + o10 + 2) << 16) | ((uint32_t)__ldg(fb + o11) << 24);
        const uint32_t hi1 = (uint32_t)__ldg(fb + o11 + 1) | ((uint32_t)__ldg(fb + o11 + 2) << 8);
        return combine(lo0, hi0, lo1, hi1, p.wx & 0xffffu, p.wy);
    }
};

}  // namespace x360

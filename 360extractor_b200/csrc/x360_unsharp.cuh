// x360_unsharp.cuh — the sharpening step that follows the blur filter on accepted views:
//   gaussian = cv2.GaussianBlur(rect_img, (0, 0), 2.0)
//   rect_img = cv2.addWeighted(rect_img, 1.0 + s, gaussian, -s, 0)        (src/core/processor.py:379-381,
//                                                                          src/ui/preview_widget.py:118-120)
// OpenCV's 8-bit Gaussian is fixed point: 13 taps {1,2,7,16,31,45,52,45,31,16,7,2,1} / 256 (the double kernel
// quantised with error diffusion), rows then columns with no rounding in between, BORDER_REFLECT_101,
// blur = (sum_j sum_i k_j k_i src + 2^15) >> 16.  addWeighted runs in float32:
// dst = sat_u8(rint(fma(src, f32(1 + s), f32(blur * f32(-s))))).   (oracle: orc_unsharp, 0 differing bytes vs cv2)
//
// unsharp_tiled: one CTA per 64 x 32 output tile of one image.
//   1. the tile + 6-pixel apron arrives as ONE TMA box (256 B x 44 rows, zero fill outside the image);
//   2. the packed BGR bytes are split into three byte planes (PRMT; border tiles: per byte, with the
//      reflect-101 source index), so that the taps of a row are consecutive bytes;
//   3. rows: 4 dp4a per output byte on constant weight words; two image rows at a time, written as
//      (row 2m | row 2m+1 << 16) words;
//   4. columns: the 13 taps of TWO output rows are 7 of those words: 7 dp2a.lo + 7 dp2a.hi on constant byte
//      weights; then the float32 addWeighted and the packed output tile in shared memory;
//   5. the tile leaves by one TMA store (clipped at the image edge).
// unsharp_generic: one thread per output byte, 13 x 13 taps with repeated reflection — for images the TMA
// path cannot take (row pitch not a multiple of 16 B, tiny or unaligned images).
#pragma once
#include <cuda.h>

#include "x360_common.cuh"
#include "x360_staged.cuh"      // mbarrier / TMA helpers

namespace x360 {

constexpr int kUsTW = 64, kUsTH = 32, kUsR = 6;              // tile, kernel radius
constexpr int kUsRows = kUsTH + 2 * kUsR;                    // 44 region rows
constexpr int kUsPx = kUsTW + 2 * kUsR;                      // 76 region pixels per row
constexpr int kUsBoxW = 256;                                 // box bytes per row: starts 32 B left of the tile
constexpr int kUsLead = 32 - 3 * kUsR;                       // 14: byte column of region pixel 0 in the box
constexpr int kUsPlanePitch = 80;                            // bytes per plane row (76 used)
constexpr int kUsThreads = 256;

__host__ __device__ constexpr int us_tap(int i)
{
    return i == 0 || i == 12 ? 1 : i == 1 || i == 11 ? 2 : i == 2 || i == 10 ? 7 : i == 3 || i == 9 ? 16
         : i == 4 || i == 8 ? 31 : i == 5 || i == 7 ? 45 : i == 6 ? 52 : 0;
}
// row pass: weight word w (bytes 4w .. 4w+3 of the 16-byte window) of output j in 0..3: tap i = 4w + b - j
__host__ __device__ constexpr uint32_t us_hweight(int j, int w)
{
    uint32_t v = 0;
    for (int b = 0; b < 4; ++b) {
        const int i = 4 * w + b - j;
        if (i >= 0 && i <= 12) v |= (uint32_t)us_tap(i) << (8 * b);
    }
    return v;
}
// column pass: byte weights of word i (rows 2i, 2i+1 of the window): (k[2i], k[2i+1]) for the even output row,
// (k[2i-1], k[2i]) for the odd one
__host__ __device__ constexpr uint32_t us_vweight(int i)
{
    return (uint32_t)us_tap(2 * i) | ((uint32_t)us_tap(2 * i + 1) << 8) | ((uint32_t)us_tap(2 * i - 1) << 16) |
           ((uint32_t)us_tap(2 * i) << 24);
}

struct UnsharpMaps {
    CUtensorMap in;      // uint8 [n][h][3w], box {256, 44, 1}
    CUtensorMap out;     // uint8 [n][h][3w], box {192, 32, 1}
};

// dynamic shared memory: box / packed output tile | planes | row-pass words | mbarrier
constexpr uint32_t kUsOffBox = 0;
constexpr uint32_t kUsOffPlanes = kUsOffBox + kUsRows * kUsBoxW;                       // 11264
constexpr uint32_t kUsOffH = kUsOffPlanes + 3 * kUsRows * kUsPlanePitch;               // + 10560
constexpr uint32_t kUsOffBar = kUsOffH + 3 * (kUsRows / 2) * kUsTW * 4;                // + 16896
constexpr uint32_t kUsSmem = kUsOffBar + 16;

__device__ __forceinline__ int us_reflect(int p, int len)        // BORDER_REFLECT_101, any distance
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;
    return p;
}

__device__ __forceinline__ uint32_t us_dp2a_lo(uint32_t a16x2, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a16x2), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t us_dp2a_hi(uint32_t a16x2, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a16x2), "r"(b), "r"(c));
    return d;
}

// cv2.addWeighted(src, 1 + s, blur, -s, 0) for one byte
__device__ __forceinline__ uint32_t us_blend(uint32_t src, uint32_t blur, float alpha, float beta)
{
    const float t = __fmul_rn((float)blur, beta);
    uint32_t q;
    asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(q) : "f"(__fmaf_rn((float)src, alpha, t)));      // half to even, clamped to 0 .. 255
    return q;
}

__global__ void __launch_bounds__(kUsThreads, 5)
unsharp_tiled(const __grid_constant__ UnsharpMaps tms, int h, int w, int n0, float alpha, float beta)
{
    extern __shared__ __align__(128) uint8_t dsm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * kUsTW, y0 = blockIdx.y * kUsTH, n = n0 + blockIdx.z;
    uint8_t *const box = dsm + kUsOffBox;
    uint8_t *const planes = dsm + kUsOffPlanes;
    uint32_t *const hw = reinterpret_cast<uint32_t *>(dsm + kUsOffH);
    unsigned long long *const bar = reinterpret_cast<unsigned long long *>(dsm + kUsOffBar);

    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_arrive_expect_tx(bar, kUsRows * kUsBoxW);
        tma_load_box(smem_u32(box), &tms.in, 3 * x0 - 32, y0 - kUsR, n, bar);
    }
    __syncthreads();
    mbar_wait(bar, 0);

    // ---- 2. byte planes: planes[c][r][xp] = pixel (x0 - 6 + xp, y0 - 6 + r), channel c ----------------
    // (outside the image the box holds zeros; border tiles patch those pixels afterwards)
    for (int t = tid; t < kUsRows * (kUsPx / 4); t += kUsThreads) {
        const int r = t / (kUsPx / 4), g = t - r * (kUsPx / 4);
        const uint32_t *src = reinterpret_cast<const uint32_t *>(box + r * kUsBoxW + (kUsLead - 2) + 12 * g);   // bytes 2 .. 13 of 16
        const uint32_t w0 = src[0], w1 = src[1], w2 = src[2], w3 = src[3];
        const uint32_t b = __byte_perm(__byte_perm(w0, w1, 0x0052), w2, 0x7410);                 // bytes 2, 5, 8, 11
        const uint32_t g_ = __byte_perm(__byte_perm(w0, w1, 0x0063), __byte_perm(w2, w3, 0x0041), 0x5410);   // 3, 6, 9, 12
        const uint32_t r_ = __byte_perm(__byte_perm(w1, w2, 0x0630), w3, 0x5210);               // 4, 7, 10, 13
        uint8_t *dst = planes + r * kUsPlanePitch + 4 * g;
        *reinterpret_cast<uint32_t *>(dst) = b;
        *reinterpret_cast<uint32_t *>(dst + kUsRows * kUsPlanePitch) = g_;
        *reinterpret_cast<uint32_t *>(dst + 2 * kUsRows * kUsPlanePitch) = r_;
    }
    const bool border = x0 < kUsR || y0 < kUsR || x0 + kUsTW + kUsR > w || y0 + kUsTH + kUsR > h;
    if (border) {                                                // uniform over the CTA
        __syncthreads();
        // reflect-101: every out-of-image pixel a valid output needs mirrors a pixel inside the box (h, w >= 7 on this path)
        for (int t = tid; t < kUsRows * kUsPx; t += kUsThreads) {
            const int r = t / kUsPx, xp = t - r * kUsPx;
            const int gy = y0 - kUsR + r, gx = x0 - kUsR + xp;
            if (gy >= 0 && gy < h && gx >= 0 && gx < w) continue;
            const int sy = us_reflect(gy, h) - (y0 - kUsR), sx = us_reflect(gx, w) - (x0 - kUsR);
            const bool ok = sy >= 0 && sy < kUsRows && sx >= -4 && 3 * sx + kUsLead + 2 < kUsBoxW;
            const uint8_t *src = box + sy * kUsBoxW + kUsLead + 3 * sx;
#pragma unroll
            for (int c = 0; c < 3; ++c) planes[(c * kUsRows + r) * kUsPlanePitch + xp] = ok ? src[c] : (uint8_t)0;
        }
    }
    __syncthreads();

    // the 16 row-pass weight words, pinned in registers (as immediates each use costs a uniform move)
    uint32_t hk[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            hk[j][i] = us_hweight(j, i);
            asm volatile("" : "+r"(hk[j][i]));
        }
    // ---- 3. rows: hw[c][m][xo] = (H[2m][xo] | H[2m+1][xo] << 16),  H[r][xo] = sum_i k_i planes[c][r][xo + i] ----
    for (int t = tid; t < 3 * (kUsRows / 2) * (kUsTW / 4); t += kUsThreads) {
        const int g = t & 15, cm = t >> 4;                       // cm = c * 22 + m
        const int c = cm / (kUsRows / 2), m = cm - c * (kUsRows / 2);
        const uint32_t *p0 = reinterpret_cast<const uint32_t *>(planes + (c * kUsRows + 2 * m) * kUsPlanePitch) + g;
        const uint32_t *p1 = p0 + kUsPlanePitch / 4;
        uint32_t a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { a[i] = p0[i]; b[i] = p1[i]; }
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t s0 = 0, s1 = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                s0 = __dp4a(a[i], hk[j][i], s0);
                s1 = __dp4a(b[i], hk[j][i], s1);
            }
            o[j] = s0 | (s1 << 16);                              // both <= 255 * 256
        }
        *reinterpret_cast<uint4 *>(hw + (size_t)cm * kUsTW + 4 * g) = make_uint4(o[0], o[1], o[2], o[3]);
    }
    __syncthreads();                                             // the box is dead: its space takes the output tile

    // ---- 4. columns + addWeighted: task (c, 32-pixel segment, 8 output rows) per warp -----------------
    uint8_t *const otile = box;                                  // [32][192]
    uint32_t vk[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        vk[i] = us_vweight(i);
        asm volatile("" : "+r"(vk[i]));
    }
    for (int task = warp; task < 3 * 2 * 4; task += kUsThreads / 32) {
        const int c = task >> 3, seg = (task >> 2) & 1, rb = task & 3;
        const int xo = 32 * seg + lane;
        const uint32_t *col = hw + (size_t)(c * (kUsRows / 2) + 4 * rb) * kUsTW + xo;
        uint32_t wv[10];
#pragma unroll
        for (int i = 0; i < 10; ++i) wv[i] = col[i * kUsTW];
        const uint8_t *ctr = planes + (c * kUsRows + 8 * rb + kUsR) * kUsPlanePitch + xo + kUsR;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            uint32_t v0 = 32768u, v1 = 32768u;
#pragma unroll
            for (int i = 0; i < 7; ++i) {
                v0 = us_dp2a_lo(wv[t + i], vk[i], v0);
                v1 = us_dp2a_hi(wv[t + i], vk[i], v1);
            }
            const int y = 8 * rb + 2 * t;
            otile[y * (kUsTW * 3) + 3 * xo + c] = (uint8_t)us_blend(ctr[(2 * t) * kUsPlanePitch], v0 >> 16, alpha, beta);
            otile[(y + 1) * (kUsTW * 3) + 3 * xo + c] = (uint8_t)us_blend(ctr[(2 * t + 1) * kUsPlanePitch], v1 >> 16, alpha, beta);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                     ::"l"(&tms.out), "r"(3 * x0), "r"(y0), "r"(n), "r"(smem_u32(otile)) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

// Any size, any alignment: one thread per output byte of image blockIdx.y.
__global__ void __launch_bounds__(256)
unsharp_generic(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, int h, int w, float alpha, float beta)
{
    const size_t pitch = (size_t)w * 3, img = (size_t)h * pitch;
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= img) return;
    const uint8_t *src = in + (size_t)blockIdx.y * img;
    const int y = (int)(i / pitch), bx = (int)(i - (size_t)y * pitch), x = bx / 3, c = bx - 3 * x;
    uint32_t v = 32768u;
    for (int j = 0; j < 13; ++j) {
        const uint8_t *row = src + (size_t)us_reflect(y + j - kUsR, h) * pitch + c;
        uint32_t hsum = 0;
        for (int k = 0; k < 13; ++k) hsum += (uint32_t)us_tap(k) * row[3 * us_reflect(x + k - kUsR, w)];
        v += (uint32_t)us_tap(j) * hsum;
    }
    out[(size_t)blockIdx.y * img + i] = (uint8_t)us_blend(src[i], v >> 16, alpha, beta);
}

}  // namespace x360

// x360_misc.cuh — the small entry points around the hot path: materialised maps,
// explicit-map remap, stand-alone blur sums, the device-resident hand-off to the masker.  These exist so that every reference call
// signature on the path stays callable on the GPU; none of them is on the batch hot path.
#pragma once
#include "x360_common.cuh"
#include "x360_lanczos.cuh"
#include "x360_linear.cuh"

namespace x360 {

// GeometryProcessor.create_rectilinear_map (src/core/geometry.py:122-192), one view.
__global__ void __launch_bounds__(256)
make_maps_kernel(const Geom g, const double *__restrict__ R, float *__restrict__ map_x, float *__restrict__ map_y)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= g.ow || y >= g.oh) return;
    float mx, my;
    map_f32(R, g, x, y, mx, my);
    map_x[(size_t)y * g.ow + x] = mx;
    map_y[(size_t)y * g.ow + x] = my;
}

// min / max of floor(map_y) over every output pixel of every view (plan creation: the host entries upload only the source rows
// the views can read, x360_plan_upload_rows).  lohi[0] starts at INT_MAX, lohi[1] at INT_MIN.
__global__ void __launch_bounds__(256)
source_row_range_kernel(const Geom g, const double *__restrict__ R, int *__restrict__ lohi)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    int lo = INT_MAX, hi = INT_MIN;
    if (x < g.ow && y < g.oh) {
        double uf;
        float my;
        map_f64(R + 9 * blockIdx.z, g, x, y, uf, my);
        lo = hi = q5_of(my) >> 5;
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if ((threadIdx.x & 31) == 0) { atomicMin(&lohi[0], lo); atomicMax(&lohi[1], hi); }
}

// cv2.remap(frame, map_x, map_y, interp, borderMode=BORDER_WRAP) with caller-held maps
// (src/core/processor.py:334, src/core/analyzer.py:63).
template <class S>
__global__ void __launch_bounds__(256)
remap_maps_kernel(const Geom g, const uint8_t *__restrict__ src, const float *__restrict__ map_x,
                  const float *__restrict__ map_y, const int16_t *__restrict__ tab, uint8_t *__restrict__ dst)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= g.ow || y >= g.oh) return;
    const size_t i = (size_t)y * g.ow + x;
    bool special = false;
    const typename S::Px p = S::make_px(q5_of(map_x[i]), q5_of(map_y[i]), g, tab, special);
    const uint32_t pxl = S::gather_generic(src, p, g, tab);
    dst[i * 3 + 0] = (uint8_t)pxl;
    dst[i * 3 + 1] = (uint8_t)(pxl >> 8);
    dst[i * 3 + 2] = (uint8_t)(pxl >> 16);
}

// cv2.cvtColor(image, COLOR_BGR2GRAY) (src/utils/image_utils.py:22-23)
__global__ void __launch_bounds__(256)
gray_kernel(const uint8_t *__restrict__ bgr, uint8_t *__restrict__ gray, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t p = (uint32_t)bgr[i * 3] | ((uint32_t)bgr[i * 3 + 1] << 8) | ((uint32_t)bgr[i * 3 + 2] << 16);
    gray[i] = (uint8_t)gray_of(p);
}

__device__ __forceinline__ int reflect101(int p, int len)
{
    if (len == 1) return 0;
    if (p < 0) p = -p;
    if (p >= len) p = 2 * len - 2 - p;
    return p;
}

// cv2.Laplacian(gray, CV_64F) + ndarray.var() sums (src/utils/image_utils.py:27)
__global__ void __launch_bounds__(256)
laplacian_sums_kernel(const uint8_t *__restrict__ gray, int h, int w, long long *sum, long long *sum2)
{
    long long pl = 0, pl2 = 0;
    const size_t n = (size_t)h * w;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int y = (int)(i / w), x = (int)(i - (size_t)y * w);
        const uint8_t *r1 = gray + (size_t)y * w;
        const int L = (int)gray[(size_t)reflect101(y - 1, h) * w + x] + (int)gray[(size_t)reflect101(y + 1, h) * w + x] +
                      (int)r1[reflect101(x - 1, w)] + (int)r1[reflect101(x + 1, w)] - 4 * (int)r1[x];
        pl += L;
        pl2 += (long long)L * L;
    }
    for (int o = 16; o > 0; o >>= 1) {
        pl += __shfl_down_sync(0xffffffffu, pl, o);
        pl2 += __shfl_down_sync(0xffffffffu, pl2, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd((unsigned long long *)sum, (unsigned long long)pl);
        atomicAdd((unsigned long long *)sum2, (unsigned long long)pl2);
    }
}

// Laplacian sums of a whole batch of views already in device memory: uint8 [n][h][w][3] -> (sum L, sum L^2)[n]
// (ImageUtils.calculate_blur_score, src/utils/image_utils.py:22-27, for every view of a launch).  The split
// alternative to the score fused into the extraction kernel: one more pass over the output (3 B / pixel), no halo
// gathers.  One CTA per 128 x 32 tile: Q15 gray of the tile and its 1-px ring (reflect-101 at the image borders,
// applied while loading) in shared memory, then the same dp4a Laplacian as the fused path.
constexpr int kBsW = 128, kBsH = 32, kBsPitch = 136;       // gray tile: column c at byte 4 + c (c = -1 .. 128)

__device__ __forceinline__ int reflect101_clamped(int p, int len)
{
    if (p < 0) p = -p;
    if (p >= len) p = 2 * len - 2 - p;
    return min(max(p, 0), len - 1);
}

__global__ void __launch_bounds__(256)
blur_sums_views_kernel(const uint8_t *__restrict__ views, int h, int w, int fast, long long *__restrict__ sum, long long *__restrict__ sum2)
{
    __shared__ __align__(16) uint8_t g[kBsH + 2][kBsPitch];
    __shared__ int red[8][2];
    const int tid = threadIdx.x, x0 = blockIdx.x * kBsW, y0 = blockIdx.y * kBsH;
    const uint8_t *img = views + (size_t)blockIdx.z * h * w * 3;
    // main part: (kBsH + 2) rows x 32 four-pixel groups, 5 per thread; all loads first (the pass is latency-bound otherwise)
    constexpr int kIt = ((kBsH + 2) * 32 + 255) / 256;
    uint32_t wv[kIt][3];
    bool quick[kIt];
#pragma unroll
    for (int k = 0; k < kIt; ++k) {
        const int i = tid + 256 * k, r = i >> 5, x = x0 + 4 * (i & 31);
        quick[k] = fast && r < kBsH + 2 && x + 3 < w;
        if (quick[k]) {
            const uint32_t *src = reinterpret_cast<const uint32_t *>(img + ((size_t)reflect101_clamped(y0 - 1 + r, h) * w + x) * 3);
            wv[k][0] = __ldg(src); wv[k][1] = __ldg(src + 1); wv[k][2] = __ldg(src + 2);   // B0 G0 R0 B1 | G1 R1 B2 G2 | R2 B3 G3 R3
        }
    }
    if (tid < 2 * (kBsH + 2)) {                                   // the ring's left / right columns
        const int r = tid >> 1, side = tid & 1;
        const uint8_t *px = img + ((size_t)reflect101_clamped(y0 - 1 + r, h) * w + reflect101_clamped(side ? x0 + kBsW : x0 - 1, w)) * 3;
        g[r][side ? 4 + kBsW : 3] = (uint8_t)gray_of((uint32_t)px[0] | ((uint32_t)px[1] << 8) | ((uint32_t)px[2] << 16));
    }
#pragma unroll
    for (int k = 0; k < kIt; ++k) {
        const int i = tid + 256 * k, r = i >> 5, gq = i & 31, x = x0 + 4 * gq;
        if (r >= kBsH + 2) continue;
        uint32_t gw;
        if (quick[k]) {
            gw = gray_of(wv[k][0]) | (gray_of(__byte_perm(wv[k][0], wv[k][1], 0x0543)) << 8) |
                 (gray_of(__byte_perm(wv[k][1], wv[k][2], 0x0432)) << 16) | (gray_of(wv[k][2] >> 8) << 24);
        } else {                                                  // image edge or unaligned rows: pixel by pixel, reflected
            const uint8_t *row = img + (size_t)reflect101_clamped(y0 - 1 + r, h) * w * 3;
            gw = 0;
            for (int j = 0; j < 4; ++j) {
                const uint8_t *px = row + (size_t)reflect101_clamped(x + j, w) * 3;
                gw |= gray_of((uint32_t)px[0] | ((uint32_t)px[1] << 8) | ((uint32_t)px[2] << 16)) << (8 * j);
            }
        }
        *reinterpret_cast<uint32_t *>(&g[r][4 + 4 * gq]) = gw;
    }
    __syncthreads();
    int pl = 0;
    unsigned pl2 = 0;
    const int gq = tid & 31, x = x0 + 4 * gq;
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
        const int r = (tid >> 5) * 4 + rr;
        if (y0 + r >= h || x >= w) continue;
        const uint32_t C = *reinterpret_cast<const uint32_t *>(&g[r + 1][4 + 4 * gq]);
        const uint32_t U = *reinterpret_cast<const uint32_t *>(&g[r][4 + 4 * gq]);
        const uint32_t D = *reinterpret_cast<const uint32_t *>(&g[r + 2][4 + 4 * gq]);
        const uint32_t l = g[r + 1][3 + 4 * gq], rg = g[r + 1][8 + 4 * gq];
        const uint32_t Lw = __byte_perm(C, l, 0x2104);             // [l, c0, c1, c2]
        const uint32_t Rw = __byte_perm(C, rg, 0x4321);            // [c1, c2, c3, r]
        constexpr uint32_t kA = 0x0001fc01u;                       // [ 1, -4,  1,  0]
        constexpr uint32_t kB = 0x01fc0100u;                       // [ 0,  1, -4,  1]
        int L[4];
        L[0] = dp4a_us(D, 0x00000001u, dp4a_us(U, 0x00000001u, dp4a_us(Lw, kA, 0)));
        L[1] = dp4a_us(D, 0x00000100u, dp4a_us(U, 0x00000100u, dp4a_us(Lw, kB, 0)));
        L[2] = dp4a_us(D, 0x00010000u, dp4a_us(U, 0x00010000u, dp4a_us(Rw, kA, 0)));
        L[3] = dp4a_us(D, 0x01000000u, dp4a_us(U, 0x01000000u, dp4a_us(Rw, kB, 0)));
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (x + j < w) { pl += L[j]; pl2 += (unsigned)(L[j] * L[j]); }
    }
    pl = __reduce_add_sync(0xffffffffu, pl);
    pl2 = __reduce_add_sync(0xffffffffu, pl2);                    // <= 32 x 16 x 1020^2 < 2^32
    if ((tid & 31) == 0) { red[tid >> 5][0] = pl; red[tid >> 5][1] = (int)pl2; }
    __syncthreads();
    if (tid == 0) {
        long long s1 = 0, s2 = 0;
        for (int k = 0; k < 8; ++k) { s1 += red[k][0]; s2 += (unsigned)red[k][1]; }
        atomicAdd((unsigned long long *)&sum[blockIdx.z], (unsigned long long)s1);
        atomicAdd((unsigned long long *)&sum2[blockIdx.z], (unsigned long long)s2);
    }
}

// The same sums by MARCHING down the view (the split score pass of plans that run it, launch_score_pass): a warp owns a
// 128-pixel-wide column strip (lane = 4 pixels = 12 packed bytes = 3 coalesced LDG.32) and walks kScRows rows down it with
// the gray words of rows y-1, y, y+1 in registers, so every pixel is loaded and converted once and nothing goes through
// shared memory: left / right neighbours are the adjacent lanes' words (2 SHFL per 4 pixels).  Strips overlap by one lane on
// either side (stride 120 pixels): lanes 0 and 31 only provide their neighbours' halo.  image_utils.py:22-27, reflect-101.
// Needs w % 4 == 0, a 4-byte aligned base, w >= 8 and h >= 2 (anything else takes blur_sums_views_kernel).
constexpr int kScRows = 32, kScStride = 120, kScWarps = 8;

// Four packed BGR pixels (B0 G0 R0 B1 | G1 R1 B2 G2 | R2 B3 G3 R3) -> their four gray bytes.  The Q15 weights are doubled so that
// the gray value ((3735 B + 19235 G + 9798 R + 2^14) >> 15, see gray_of) is byte 2 of the accumulator, and placed per pixel
// position so that no pixel has to be realigned first: 8 DP2A + 3 PRMT.
__device__ __forceinline__ uint32_t gray4_of(uint32_t w0, uint32_t w1, uint32_t w2)
{
    constexpr uint32_t kB = 2 * 3735u, kG = 2 * 19235u, kR = 2 * 9798u;
    uint32_t a0 = 32768u, a1 = 32768u, a2 = 32768u, a3 = 32768u;
    asm("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(a0) : "r"(kB | (kG << 16)), "r"(w0));
    asm("dp2a.hi.u32.u32 %0, %1, %2, %0;" : "+r"(a0) : "r"(kR), "r"(w0));
    asm("dp2a.hi.u32.u32 %0, %1, %2, %0;" : "+r"(a1) : "r"(kB << 16), "r"(w0));
    asm("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(a1) : "r"(kG | (kR << 16)), "r"(w1));
    asm("dp2a.hi.u32.u32 %0, %1, %2, %0;" : "+r"(a2) : "r"(kB | (kG << 16)), "r"(w1));
    asm("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(a2) : "r"(kR), "r"(w2));
    asm("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(a3) : "r"(kB << 16), "r"(w2));
    asm("dp2a.hi.u32.u32 %0, %1, %2, %0;" : "+r"(a3) : "r"(kG | (kR << 16)), "r"(w2));
    return __byte_perm(__byte_perm(__byte_perm(a0, a1, 0x0062), a2, 0x0610), a3, 0x6210);
}

// One row of a lane's four pixels: C = this row's gray word, U / D the rows above / below; the left / right neighbours are
// bytes of the adjacent lanes' words, picked by per-lane PRMT selectors (which also encode the reflect-101 image edges).
__device__ __forceinline__ void march_row(uint32_t U, uint32_t C, uint32_t D, uint32_t sel_l, uint32_t sel_r, int &pl, unsigned &pl2)
{
    const uint32_t up = __shfl_up_sync(0xffffffffu, C, 1), dn = __shfl_down_sync(0xffffffffu, C, 1);
    const uint32_t Lw = __byte_perm(C, up, sel_l);                // [l, c0, c1, c2]
    const uint32_t Rw = __byte_perm(C, dn, sel_r);                // [c1, c2, c3, r]
    constexpr uint32_t kA = 0x0001fc01u;                          // [ 1, -4,  1,  0]
    constexpr uint32_t kB = 0x01fc0100u;                          // [ 0,  1, -4,  1]
    const int L0 = dp4a_us(D, 0x00000001u, dp4a_us(U, 0x00000001u, dp4a_us(Lw, kA, 0)));
    const int L1 = dp4a_us(D, 0x00000100u, dp4a_us(U, 0x00000100u, dp4a_us(Lw, kB, 0)));
    const int L2 = dp4a_us(D, 0x00010000u, dp4a_us(U, 0x00010000u, dp4a_us(Rw, kA, 0)));
    const int L3 = dp4a_us(D, 0x01000000u, dp4a_us(U, 0x01000000u, dp4a_us(Rw, kB, 0)));
    pl += L0 + L1 + L2 + L3;
    pl2 += (unsigned)(L0 * L0) + (unsigned)(L1 * L1) + (unsigned)(L2 * L2) + (unsigned)(L3 * L3);
}

__global__ void __launch_bounds__(32 * kScWarps)
blur_sums_march_kernel(const uint8_t *__restrict__ views, int h, int w, long long *__restrict__ sum, long long *__restrict__ sum2)
{
    __shared__ int red[kScWarps][2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x = (int)blockIdx.x * kScStride - 4 + 4 * lane;                 // this lane's four pixels x .. x+3
    const int r0 = ((int)blockIdx.y * kScWarps + warp) * kScRows, r1 = min(r0 + kScRows, h);
    const bool live = x >= 0 && x < w;
    const bool counts = live && lane > 0 && lane < 31;                        // lanes 0 / 31: halo providers
    // lanes outside the image walk column 0: their sums are dropped at the end
    const uint8_t *col = views + (size_t)blockIdx.z * h * w * 3 + (live ? (size_t)x * 3 : 0);
    const unsigned pitch = (unsigned)w * 3u;                                  // bytes per row (32-bit: one IMAD.WIDE per row address)
    const uint32_t sel_l = x == 0 ? 0x2101u : 0x2107u;                        // g[-1] = g[1]; else byte 3 of the lane on the left
    const uint32_t sel_r = x + 3 == w - 1 ? 0x2321u : 0x4321u;                // g[w] = g[w-2]; else byte 0 of the lane on the right
    int pl = 0;
    unsigned pl2 = 0;
    if (r0 < h) {
        auto load_row = [&](int y, uint32_t (&r)[3]) {
            const uint32_t *src = reinterpret_cast<const uint32_t *>(col + (size_t)((unsigned)min(y, h - 1) * (unsigned long long)pitch));
            r[0] = __ldg(src); r[1] = __ldg(src + 1); r[2] = __ldg(src + 2);
        };
        // rows are loaded four at a time, one group ahead of the group being summed: 2 x 4 x 12 bytes per lane in flight
        uint32_t ra[3], rb[3], ga[4][3], gb[4][3];
        load_row(r0 == 0 ? 1 : r0 - 1, ra);
        load_row(r0, rb);
#pragma unroll
        for (int j = 0; j < 4; ++j) load_row(r0 + 1 + j, ga[j]);
        uint32_t U = gray4_of(ra[0], ra[1], ra[2]), C = gray4_of(rb[0], rb[1], rb[2]);
        auto four_rows = [&](uint32_t (&below)[4][3]) {                       // rows whose row below is inside the image
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t D = gray4_of(below[j][0], below[j][1], below[j][2]);
                march_row(U, C, D, sel_l, sel_r, pl, pl2);
                U = C;
                C = D;
            }
        };
        // whole groups first (all four rows inside this warp's range and above the image's last row), two per trip so that
        // the two row buffers swap roles without register moves
        int y0 = r0, groups = max(0, min(r1, h - 1) - r0) / 4;
#pragma unroll 1
        for (; groups >= 2; groups -= 2, y0 += 8) {
#pragma unroll
            for (int j = 0; j < 4; ++j) load_row(y0 + 5 + j, gb[j]);
            four_rows(ga);
            if (y0 + 8 < r1) {
#pragma unroll
                for (int j = 0; j < 4; ++j) load_row(y0 + 9 + j, ga[j]);
            }
            four_rows(gb);
        }
        if (groups) {
            if (y0 + 4 < r1) {
#pragma unroll
                for (int j = 0; j < 4; ++j) load_row(y0 + 5 + j, gb[j]);
            }
            four_rows(ga);
#pragma unroll
            for (int j = 0; j < 4; ++j) { ga[j][0] = gb[j][0]; ga[j][1] = gb[j][1]; ga[j][2] = gb[j][2]; }
            y0 += 4;
        }
        // what is left (fewer than four rows, or the image's last group): row by row, with the reflect-101 bottom edge
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int y = y0 + j;
            if (y < r1) {                                                     // uniform over the warp
                const uint32_t D = (y + 1 < h) ? gray4_of(ga[j][0], ga[j][1], ga[j][2]) : U;        // g[h] = g[h-2]
                march_row(U, C, D, sel_l, sel_r, pl, pl2);
                U = C;
                C = D;
            }
        }
    }
    if (!counts) { pl = 0; pl2 = 0u; }
    pl = __reduce_add_sync(0xffffffffu, pl);
    pl2 = __reduce_add_sync(0xffffffffu, pl2);                                // <= 30 lanes x 32 rows x 4 x 1020^2 < 2^32
    if (lane == 0) { red[warp][0] = pl; red[warp][1] = (int)pl2; }
    __syncthreads();
    if (tid == 0) {
        long long s1 = 0, s2 = 0;
#pragma unroll
        for (int k = 0; k < kScWarps; ++k) { s1 += red[k][0]; s2 += (unsigned)red[k][1]; }
        atomicAdd((unsigned long long *)&sum[blockIdx.z], (unsigned long long)s1);
        atomicAdd((unsigned long long *)&sum2[blockIdx.z], (unsigned long long)s2);
    }
}

// hand the views of the selected faces to the masker without leaving the device.
// src/core/processor.py:392-420 passes the BGR views (all of them, or the faces named in 'ai_mask_cameras') to
// AIService.process_batch (src/core/ai_model.py:125-167), whose model turns each image into a normalised
// CHW float tensor (BGR -> RGB, HWC -> CHW, / 255).  Here: uint8 [.][h][w][3] views in device memory ->
// float32 [n_sel][3][h][w]; value = float(v) / 255 correctly rounded (what numpy's float32 division gives).
struct NchwSel {
    int idx[64];                       // view slots (flat [frame * V + view]) of this launch
};

__device__ __forceinline__ float u8_to_unit(uint32_t v) { return __fdiv_rn((float)v, 255.0f); }

// one thread per 4 pixels: 12 packed bytes in (3 LDG.32), one float4 to each of the three planes
__global__ void __launch_bounds__(256)
views_to_nchw_kernel(const uint8_t *__restrict__ views, size_t view_bytes, int groups, const NchwSel sel, float *__restrict__ out,
                     int to_rgb)
{
    const int g = blockIdx.x * 256 + threadIdx.x;
    if (g >= groups) return;
    const uint32_t *src = reinterpret_cast<const uint32_t *>(views + (size_t)sel.idx[blockIdx.y] * view_bytes) + 3 * (size_t)g;
    const uint32_t w0 = __ldg(src), w1 = __ldg(src + 1), w2 = __ldg(src + 2);       // B0 G0 R0 B1 | G1 R1 B2 G2 | R2 B3 G3 R3
    const float4 cb = make_float4(u8_to_unit(w0 & 255u), u8_to_unit(w0 >> 24), u8_to_unit((w1 >> 16) & 255u), u8_to_unit((w2 >> 8) & 255u));
    const float4 cg = make_float4(u8_to_unit((w0 >> 8) & 255u), u8_to_unit(w1 & 255u), u8_to_unit(w1 >> 24), u8_to_unit((w2 >> 16) & 255u));
    const float4 cr = make_float4(u8_to_unit((w0 >> 16) & 255u), u8_to_unit((w1 >> 8) & 255u), u8_to_unit(w2 & 255u), u8_to_unit(w2 >> 24));
    const size_t hw = (size_t)groups * 4;
    float4 *o = reinterpret_cast<float4 *>(out + (size_t)blockIdx.y * 3 * hw) + g;
    o[0] = to_rgb ? cr : cb;
    o[hw / 4] = cg;
    o[2 * (hw / 4)] = to_rgb ? cb : cr;
}

// any size / alignment: one thread per pixel
__global__ void __launch_bounds__(256)
views_to_nchw_generic(const uint8_t *__restrict__ views, size_t view_bytes, size_t hw, const NchwSel sel, float *__restrict__ out, int to_rgb)
{
    const size_t p = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (p >= hw) return;
    const uint8_t *px = views + (size_t)sel.idx[blockIdx.y] * view_bytes + 3 * p;
    float *o = out + (size_t)blockIdx.y * 3 * hw + p;
    o[0] = u8_to_unit(px[to_rgb ? 2 : 0]);
    o[hw] = u8_to_unit(px[1]);
    o[2 * hw] = u8_to_unit(px[to_rgb ? 0 : 2]);
}

}  // namespace x360

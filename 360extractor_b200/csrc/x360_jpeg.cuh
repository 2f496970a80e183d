// x360_jpeg.cuh — baseline JPEG encode of a view batch on the device (the codec step after the path).
//
// Replaces, for every emitted view of a frame batch, the save step of the reference
//     cv2.imwrite(path, rect_img, [cv2.IMWRITE_JPEG_QUALITY, quality])      src/utils/file_manager.py:21-32
//     (save_params: src/core/processor.py:121-123; quality default 95: src/core/settings_manager.py:24)
// so that only the compressed bytes cross PCIe.  The output is BYTE-IDENTICAL to what OpenCV's bundled libjpeg-turbo
// writes for the same pixels (the test suite pins it to cv2.imencode and to a plain-C restatement of libjpeg-turbo):
//   YCbCr 4:2:0, 8-bit baseline, the integer "islow" forward DCT, IJG-scaled Annex-K quantisation tables with
//   libjpeg-turbo's reciprocal-multiplication quantiser, Annex-K Huffman tables, JFIF 1.01 header, no restart markers.
//
// Kernels (one thread per 8 x 8 block where the work is sequential by nature; the encode runs beside a PCIe-bound
// pipeline — at 45 GB/s a 1600 x 1600 view is 170 us of D2H, its encode a few us of these kernels):
//   jpeg_blocks_kernel   pixels -> colour conversion, h2v2 chroma downsampling, FDCT, quantisation; int16 [mcu][6][64]
//                        in zigzag order (Y00 Y01 Y10 Y11 Cb Cr per 16 x 16 MCU), dummy luma blocks beyond the image
//   jpeg_bits_kernel     entropy-coded size of every block in bits (DC difference chains included)
//   jpeg_scan_kernel     per view: exclusive scan of the block sizes -> bit offset of every block, bits per view
//   jpeg_bases_kernel    across views: word offsets of the views' unstuffed streams; byte offsets of the final files
//   jpeg_emit_kernel     Huffman codes of every block OR-ed into the (zeroed) MSB-first word stream at its bit offset
//   jpeg_ffcount / _ffscan_kernel   0xFF bytes per 4096-byte segment of the stream; per view scanned -> stuffed segment offsets, file size
//   jpeg_pack_kernel     header + stream with 0xFF 0x00 stuffing + the 1-bit padding of the last byte + EOI -> final bytes
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace x360 {

constexpr int kJpegMaxHeader = 640;               // SOI .. SOS of a 3-component baseline file with 2 + 4 tables: 623 bytes
constexpr int kJpegSeg = 4096;                    // bytes per stuffing segment (one CTA of the packers)

struct JpegTables {                                // per call, by value (kernel parameter space)
    uint32_t recip[128];                           // [0,64): luma, [64,128): chroma — natural (row-major) coefficient order
    uint16_t corr[128];
    uint8_t shift[128];                            // right shift after the 16 bits of the reciprocal
    uint8_t q1[128];                               // 1 where the divisor is 1 (libjpeg-turbo's special case: no division)
};

struct JpegHuff {                                  // Annex-K tables as code / size per symbol: [dc lum, ac lum, dc chr, ac chr]
    uint16_t code[4][256];
    uint8_t size[4][256];
};
__constant__ JpegHuff c_jpeg_huff;
__constant__ uint8_t c_jpeg_zigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                          41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                          30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct JpegGeom {
    int h, w;                                      // view size
    int mw, mh;                                    // MCUs across / down (16 x 16 pixels)
    int bw_y, bh_y;                                // real luma blocks across / down: ceil(w / 8), ceil(h / 8)
    int crows;                                     // real downsampled chroma rows: ceil(h / 2)
    int n_mcu, n_blocks;                           // per view: mw * mh, 6 * n_mcu
};

// jfdctint.c (JDCT_ISLOW): CONST_BITS 13, PASS1_BITS 2.  One 1-D pass over 8 values with stride s.
__device__ __forceinline__ int jdescale(int x, int n) { return (x + (1 << (n - 1))) >> n; }
template <int PASS> __device__ __forceinline__ void fdct8(int *p, const int s)
{
    const int t0 = p[0] + p[7 * s], t7 = p[0] - p[7 * s], t1 = p[s] + p[6 * s], t6 = p[s] - p[6 * s];
    const int t2 = p[2 * s] + p[5 * s], t5 = p[2 * s] - p[5 * s], t3 = p[3 * s] + p[4 * s], t4 = p[3 * s] - p[4 * s];
    const int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
    constexpr int sh = PASS == 0 ? 13 - 2 : 13 + 2;
    if (PASS == 0) {
        p[0] = (t10 + t11) << 2;
        p[4 * s] = (t10 - t11) << 2;
    } else {
        p[0] = jdescale(t10 + t11, 2);
        p[4 * s] = jdescale(t10 - t11, 2);
    }
    int z1 = (t12 + t13) * 4433;
    p[2 * s] = jdescale(z1 + t13 * 6270, sh);
    p[6 * s] = jdescale(z1 + t12 * (-15137), sh);
    z1 = t4 + t7;
    int z2 = t5 + t6, z3 = t4 + t6, z4 = t5 + t7;
    const int z5 = (z3 + z4) * 9633;
    const int m4 = t4 * 2446, m5 = t5 * 16819, m6 = t6 * 25172, m7 = t7 * 12299;
    z1 *= -7373; z2 *= -20995; z3 *= -16069; z4 *= -3196;
    z3 += z5; z4 += z5;
    p[7 * s] = jdescale(m4 + z1 + z3, sh);
    p[5 * s] = jdescale(m5 + z2 + z4, sh);
    p[3 * s] = jdescale(m6 + z2 + z3, sh);
    p[s] = jdescale(m7 + z1 + z4, sh);
}

// jccolor.c rgb_ycc_convert (SCALEBITS 16) on a BGR pixel
__device__ __forceinline__ int jpeg_y(const uint8_t *p) { return (19595 * p[2] + 38470 * p[1] + 7471 * p[0] + 32768) >> 16; }
__device__ __forceinline__ int jpeg_cb(const uint8_t *p) { return (-11059 * p[2] - 21709 * p[1] + 32768 * p[0] + (128 << 16) + 32767) >> 16; }
__device__ __forceinline__ int jpeg_cr(const uint8_t *p) { return (32768 * p[2] - 27439 * p[1] - 5329 * p[0] + (128 << 16) + 32767) >> 16; }

// jcdctmgr.c quantize (reciprocal form); v = DCT output (scaled by 8), i = natural index (+64 for chroma)
__device__ __forceinline__ int jpeg_quant(int v, const JpegTables &t, int i)
{
    const uint32_t a = (uint32_t)(v < 0 ? -v : v);
    uint32_t q = t.q1[i] ? a : (((a + t.corr[i]) * t.recip[i]) >> (t.shift[i] + 16));
    q &= 0xffffu;
    const int r = (int)(int16_t)q;
    return v < 0 ? -r : r;
}

// pixels -> coefficients.  A CTA of 256 threads owns one strip of an MCU row: 16 pixel rows x 512 columns = 32 MCUs.
//   phase 1  thread = (row pair, 16-pixel column group): two rows of 48 bytes (3 x LDG.128 each where the view's rows are
//            16-byte aligned; clamped byte loads at the right / bottom edge), colour conversion, the h2v2 sums -> Y [16][512],
//            Cb / Cr [8][256] in shared memory (every pixel is read and converted once)
//   phase 2  thread = block (32 MCUs x 6 = 192 of the 256 threads): FDCT and quantisation out of shared memory, 128 bytes of
//            zigzag-ordered int16 out (8 x STG.128)
constexpr int kJbMcus = 32, kJbThreads = 256;

__device__ __forceinline__ int jpeg_y3(int b, int g, int r) { return (19595 * r + 38470 * g + 7471 * b + 32768) >> 16; }
__device__ __forceinline__ int jpeg_cb3(int b, int g, int r) { return (-11059 * r - 21709 * g + 32768 * b + (128 << 16) + 32767) >> 16; }
__device__ __forceinline__ int jpeg_cr3(int b, int g, int r) { return (32768 * r - 27439 * g - 5329 * b + (128 << 16) + 32767) >> 16; }

__global__ void __launch_bounds__(kJbThreads)
jpeg_blocks_kernel(const uint8_t *__restrict__ views, const JpegGeom g, const __grid_constant__ JpegTables tab, int16_t *__restrict__ coef)
{
    __shared__ __align__(16) uint8_t ys[16][kJbMcus * 16];
    __shared__ __align__(16) uint8_t cs[2][8][kJbMcus * 8];
    const int tid = threadIdx.x, my = blockIdx.y, mx0 = blockIdx.x * kJbMcus;
    const uint8_t *img = views + (size_t)blockIdx.z * g.h * g.w * 3;
    const size_t pitch = (size_t)g.w * 3;
    {
        const int rp = tid >> 5, cg = tid & 31;                              // row pair 0..7, column group 0..31
        const int x0 = mx0 * 16 + cg * 16, y0 = my * 16 + 2 * rp;
        if (x0 < g.mw * 16) {                                                // (a strip may end before 32 MCUs)
            const bool fast = ((reinterpret_cast<uintptr_t>(img) | pitch) & 15) == 0 && x0 + 16 <= g.w && y0 + 1 <= g.h - 1;
            if (fast) {
                uint32_t w[2][12];
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const uint4 *src = reinterpret_cast<const uint4 *>(img + (size_t)(y0 + r) * pitch + (size_t)x0 * 3);
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const uint4 v = __ldg(src + i);
                        w[r][4 * i] = v.x; w[r][4 * i + 1] = v.y; w[r][4 * i + 2] = v.z; w[r][4 * i + 3] = v.w;
                    }
                }
                auto chan = [&](int r, int byte) -> int { return (int)((w[r][byte >> 2] >> (8 * (byte & 3))) & 0xffu); };
                uint32_t yw[2][4] = {};
                uint32_t cbw[2] = {}, crw[2] = {};
#pragma unroll
                for (int j = 0; j < 8; ++j) {                                // pixel pairs 2j, 2j + 1 of both rows
                    int cb = 0, cr = 0;
#pragma unroll
                    for (int r = 0; r < 2; ++r)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int px = 2 * j + e, B = chan(r, 3 * px), G = chan(r, 3 * px + 1), R = chan(r, 3 * px + 2);
                            yw[r][px >> 2] |= (uint32_t)jpeg_y3(B, G, R) << (8 * (px & 3));
                            cb += jpeg_cb3(B, G, R);
                            cr += jpeg_cr3(B, G, R);
                        }
                    const int bias = (j & 1) ? 2 : 1;                        // (x0 / 2 is even: the bias alternates from 1)
                    cbw[j >> 2] |= (uint32_t)((cb + bias) >> 2) << (8 * (j & 3));
                    crw[j >> 2] |= (uint32_t)((cr + bias) >> 2) << (8 * (j & 3));
                }
#pragma unroll
                for (int r = 0; r < 2; ++r)
                    *reinterpret_cast<uint4 *>(&ys[2 * rp + r][cg * 16]) = make_uint4(yw[r][0], yw[r][1], yw[r][2], yw[r][3]);
                *reinterpret_cast<uint2 *>(&cs[0][rp][cg * 8]) = make_uint2(cbw[0], cbw[1]);
                *reinterpret_cast<uint2 *>(&cs[1][rp][cg * 8]) = make_uint2(crw[0], crw[1]);
            } else {
                // edge expansion (jcprepct.c / jcsample.c): luma repeats the last row / column; chroma replicates columns before
                // the downsampling, and below the image the DOWNSAMPLED rows repeat (an odd height repeats its last row once)
                for (int r = 0; r < 2; ++r) {
                    const uint8_t *row = img + (size_t)min(y0 + r, g.h - 1) * pitch;
                    for (int j = 0; j < 16; ++j) {
                        const uint8_t *p = row + (size_t)min(x0 + j, g.w - 1) * 3;
                        ys[2 * rp + r][cg * 16 + j] = (uint8_t)jpeg_y3(p[0], p[1], p[2]);
                    }
                }
                const int cy = min(my * 8 + rp, g.crows - 1);
                const uint8_t *r0 = img + (size_t)(2 * cy) * pitch, *r1 = img + (size_t)min(2 * cy + 1, g.h - 1) * pitch;
                for (int j = 0; j < 8; ++j) {
                    const size_t xa = (size_t)min(x0 + 2 * j, g.w - 1) * 3, xb = (size_t)min(x0 + 2 * j + 1, g.w - 1) * 3;
                    const int bias = (j & 1) ? 2 : 1;
                    const int cb = jpeg_cb3(r0[xa], r0[xa + 1], r0[xa + 2]) + jpeg_cb3(r0[xb], r0[xb + 1], r0[xb + 2]) +
                                   jpeg_cb3(r1[xa], r1[xa + 1], r1[xa + 2]) + jpeg_cb3(r1[xb], r1[xb + 1], r1[xb + 2]);
                    const int cr = jpeg_cr3(r0[xa], r0[xa + 1], r0[xa + 2]) + jpeg_cr3(r0[xb], r0[xb + 1], r0[xb + 2]) +
                                   jpeg_cr3(r1[xa], r1[xa + 1], r1[xa + 2]) + jpeg_cr3(r1[xb], r1[xb + 1], r1[xb + 2]);
                    cs[0][rp][cg * 8 + j] = (uint8_t)((cb + bias) >> 2);
                    cs[1][rp][cg * 8 + j] = (uint8_t)((cr + bias) >> 2);
                }
            }
        }
    }
    __syncthreads();
    if (tid >= kJbMcus * 6) return;
    const int ml = tid / 6, b = tid - 6 * ml, mx = mx0 + ml;
    if (mx >= g.mw) return;
    int16_t *out = coef + (((size_t)blockIdx.z * g.n_mcu + (size_t)my * g.mw + mx) * 6 + b) * 64;
    int d[64];
    if (b < 4) {
        const int bx = mx * 2 + (b & 1), by = my * 2 + (b >> 1);
        if (by >= g.bh_y || bx >= g.bw_y) {
            // dummy block (jccoefct.c): zeros here; its DC is copied from its predecessor by jpeg_dummy_dc_kernel
#pragma unroll
            for (int k = 0; k < 64; k += 8) *reinterpret_cast<uint4 *>(out + k) = make_uint4(0u, 0u, 0u, 0u);
            return;
        }
#pragma unroll
        for (int y = 0; y < 8; ++y) {
            const uint2 v = *reinterpret_cast<const uint2 *>(&ys[(b >> 1) * 8 + y][ml * 16 + (b & 1) * 8]);
#pragma unroll
            for (int x = 0; x < 8; ++x) d[8 * y + x] = (int)(((x < 4 ? v.x : v.y) >> (8 * (x & 3))) & 0xffu) - 128;
        }
    } else {
#pragma unroll
        for (int y = 0; y < 8; ++y) {
            const uint2 v = *reinterpret_cast<const uint2 *>(&cs[b - 4][y][ml * 8]);
#pragma unroll
            for (int x = 0; x < 8; ++x) d[8 * y + x] = (int)(((x < 4 ? v.x : v.y) >> (8 * (x & 3))) & 0xffu) - 128;
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) fdct8<0>(d + 8 * i, 1);
#pragma unroll
    for (int i = 0; i < 8; ++i) fdct8<1>(d + i, 8);
    const int tb = b < 4 ? 0 : 64;
    // quantise in natural order (all indices compile-time: d[] stays in registers), then store in zigzag order
    int16_t q[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) q[i] = (int16_t)jpeg_quant(d[i], tab, tb + i);
    constexpr uint8_t zz[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
#pragma unroll
    for (int k = 0; k < 64; k += 8) {
        uint32_t wv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) wv[i] = (uint16_t)q[zz[k + 2 * i]] | ((uint32_t)(uint16_t)q[zz[k + 2 * i + 1]] << 16);
        *reinterpret_cast<uint4 *>(out + k) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
    }
}

// Dummy luma blocks take the DC of the block before them in their MCU: the block to the left at the right edge, block
// [first of the row] - 1 for a whole dummy row at the bottom (jccoefct.c compress_data).  One thread per MCU; only sizes that are
// not multiples of 16 have any.
__global__ void jpeg_dummy_dc_kernel(const JpegGeom g, int16_t *__restrict__ coef)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= g.n_mcu) return;
    const int my = m / g.mw, mx = m - my * g.mw;
    int16_t *c = coef + ((size_t)blockIdx.y * g.n_mcu + m) * 6 * 64;
    for (int b = 1; b < 4; ++b) {                                            // in block order: a copy may chain through an earlier dummy
        const int bx = mx * 2 + (b & 1), by = my * 2 + (b >> 1);
        if (by >= g.bh_y) c[b * 64] = c[((b & ~1) - 1) * 64];
        else if (bx >= g.bw_y) c[b * 64] = c[(b - 1) * 64];
    }
}

__device__ __forceinline__ int jpeg_nbits(int v) { return 32 - __clz(v < 0 ? -v : v); }

// DC predictor of block bi (jchuff.c last_dc_val per component): the previous block of the same component in scan order
__device__ __forceinline__ int jpeg_pred(const int16_t *view_coef, int bi)
{
    const int m = bi / 6, b = bi - 6 * m;
    if (b >= 4) return m > 0 ? view_coef[((size_t)(m - 1) * 6 + b) * 64] : 0;
    if (b > 0) return view_coef[((size_t)bi - 1) * 64];
    return m > 0 ? view_coef[((size_t)(m - 1) * 6 + 3) * 64] : 0;
}

// The 64 coefficients of a block as 32 packed words (8 x LDG.128); coefficient k is a compile-time index in the loops below.
__device__ __forceinline__ void jpeg_load_block(const int16_t *__restrict__ c, uint32_t (&w)[32])
{
    const uint4 *c4 = reinterpret_cast<const uint4 *>(c);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint4 v = __ldg(c4 + i);
        w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
    }
}
__device__ __forceinline__ int jpeg_coef(const uint32_t (&w)[32], int k) { return (int)(int16_t)(w[k >> 1] >> (16 * (k & 1))); }

// Entropy-coded size of every block, in bits (jchuff.c encode_one_block without the writes).  One thread per block; the code
// sizes come from shared memory (a constant-bank read with a different index per lane is serialised).
__global__ void __launch_bounds__(128)
jpeg_bits_kernel(const int16_t *__restrict__ coef, const JpegGeom g, uint32_t *__restrict__ bits)
{
    __shared__ uint8_t s_size[4][256];
    for (int i = threadIdx.x; i < 1024; i += 128) s_size[i >> 8][i & 255] = c_jpeg_huff.size[i >> 8][i & 255];
    __syncthreads();
    const int bi = blockIdx.x * blockDim.x + threadIdx.x;
    if (bi >= g.n_blocks) return;
    const int16_t *vc = coef + (size_t)blockIdx.y * g.n_blocks * 64;
    uint32_t w[32];
    jpeg_load_block(vc + (size_t)bi * 64, w);
    const int b = bi % 6, tdc = b < 4 ? 0 : 2, tac = tdc + 1;
    int nb = jpeg_nbits(jpeg_coef(w, 0) - jpeg_pred(vc, bi));
    uint32_t n = s_size[tdc][nb] + nb;
    const uint32_t zrl = s_size[tac][0xf0];
    int run = 0;
#pragma unroll
    for (int k = 1; k < 64; ++k) {
        const int v = jpeg_coef(w, k);
        if (v == 0) { ++run; continue; }
        n += (uint32_t)(run >> 4) * zrl;
        nb = jpeg_nbits(v);
        n += s_size[tac][((run & 15) << 4) | nb] + nb;
        run = 0;
    }
    if (run > 0) n += s_size[tac][0];
    bits[(size_t)blockIdx.y * g.n_blocks + bi] = n;
}

// Per view (one CTA of 1024 threads, four items per thread and trip): exclusive scan of the block sizes in place; view_bits[v] = total.
__global__ void __launch_bounds__(1024)
jpeg_scan_kernel(uint32_t *__restrict__ bits, int n_items, unsigned long long *__restrict__ totals)
{
    __shared__ uint32_t warp_sum[32];
    __shared__ uint32_t carry;
    uint32_t *v = bits + (size_t)blockIdx.x * n_items;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n_items; base += 4096) {
        const int i = base + 4 * tid;
        uint32_t x[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) x[j] = i + j < n_items ? v[i + j] : 0u;
        const uint32_t mine = x[0] + x[1] + x[2] + x[3];
        uint32_t inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += y;
        }
        if (lane == 31) warp_sum[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = warp_sum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            warp_sum[lane] = w;                                             // inclusive over warps
        }
        __syncthreads();
        uint32_t before = carry + (warp > 0 ? warp_sum[warp - 1] : 0u) + inc - mine;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (i + j < n_items) v[i + j] = before;
            before += x[j];
        }
        __syncthreads();
        if (tid == 1023) carry = before;
        __syncthreads();
    }
    if (tid == 0) totals[blockIdx.x] = carry;
}

// Across views (one thread: a few thousand views at most): word offset of each view's unstuffed stream, a multiple of four
// words (the packers read it 16 bytes at a time).
__global__ void jpeg_bases_kernel(const unsigned long long *__restrict__ view_bits, int n, unsigned long long *__restrict__ word_base,
                                  unsigned long long cap_words, int *__restrict__ overflow)
{
    if (threadIdx.x || blockIdx.x) return;
    unsigned long long w = 0;
    for (int v = 0; v < n; ++v) {
        word_base[v] = w;
        w += (((view_bits[v] + 31ull) / 32ull + 1ull) + 3ull) & ~3ull;       // (+1: the emitter may touch the word after the last bit)
    }
    word_base[n] = w;
    if (w > cap_words) *overflow = 1;
}

// Zero the words the batch's streams will occupy (word_base[n] of them: a fraction of the scratch's capacity)
__global__ void __launch_bounds__(256)
jpeg_zero_kernel(uint32_t *__restrict__ stream, const unsigned long long *__restrict__ word_base, int n, const int *__restrict__ overflow)
{
    if (*overflow) return;
    const unsigned long long quads = word_base[n] >> 2;
    uint4 *s4 = reinterpret_cast<uint4 *>(stream);
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < quads; i += (unsigned long long)gridDim.x * blockDim.x)
        s4[i] = make_uint4(0u, 0u, 0u, 0u);
}

// Huffman codes of every block into the zeroed MSB-first word stream (jchuff.c encode_one_block).  One thread per block: the
// codes are assembled in a register, one stream word at a time.  A word that lies entirely inside the block's bit range is
// stored; the block's first and last words, shared with its neighbours, are OR-ed in atomically.
__global__ void __launch_bounds__(128)
jpeg_emit_kernel(const int16_t *__restrict__ coef, const JpegGeom g, const uint32_t *__restrict__ bit_off,
                 const unsigned long long *__restrict__ word_base, uint32_t *__restrict__ stream, const int *__restrict__ overflow)
{
    __shared__ uint32_t s_code[4][256];                                      // code | size << 16
    for (int i = threadIdx.x; i < 1024; i += 128) s_code[i >> 8][i & 255] = (uint32_t)c_jpeg_huff.code[i >> 8][i & 255] | ((uint32_t)c_jpeg_huff.size[i >> 8][i & 255] << 16);
    __syncthreads();
    const int bi = blockIdx.x * blockDim.x + threadIdx.x;
    if (bi >= g.n_blocks || *overflow) return;
    const int16_t *vc = coef + (size_t)blockIdx.y * g.n_blocks * 64;
    uint32_t w[32];
    jpeg_load_block(vc + (size_t)bi * 64, w);
    const int b = bi % 6, tdc = b < 4 ? 0 : 2, tac = tdc + 1;
    const uint32_t pos = bit_off[(size_t)blockIdx.y * g.n_blocks + bi];
    uint32_t *wp = stream + word_base[blockIdx.y] + (pos >> 5);
    int fill = (int)(pos & 31u);                                             // bits of the current word already taken
    uint32_t acc = 0u;
    bool shared_word = fill != 0;                                            // the current word began in an earlier block
    auto put = [&](uint32_t val, int len) {                                  // len <= 26 bits, MSB first
        const int room = 32 - fill;
        if (len < room) { acc |= val << (room - len); fill += len; return; }
        acc |= val >> (len - room);
        if (shared_word) atomicOr(wp, acc); else *wp = acc;
        shared_word = false;
        ++wp;
        fill = len - room;
        acc = fill ? val << (32 - fill) : 0u;
    };
    auto put_sym = [&](int table, int sym, int v, int nb) {                  // Huffman code of `sym`, then the nb value bits of v
        const uint32_t e = s_code[table][sym];
        const uint32_t vb = (uint32_t)(v < 0 ? v - 1 : v) & ((1u << nb) - 1u);
        put(((e & 0xffffu) << nb) | vb, (int)(e >> 16) + nb);
    };
    {
        const int diff = jpeg_coef(w, 0) - jpeg_pred(vc, bi);
        const int nb = jpeg_nbits(diff);
        put_sym(tdc, nb, diff, nb);
    }
    const uint32_t zrl = s_code[tac][0xf0];
    int run = 0;
#pragma unroll
    for (int k = 1; k < 64; ++k) {
        const int v = jpeg_coef(w, k);
        if (v == 0) { ++run; continue; }
        while (run > 15) { put(zrl & 0xffffu, (int)(zrl >> 16)); run -= 16; }
        const int nb = jpeg_nbits(v);
        put_sym(tac, (run << 4) | nb, v, nb);
        run = 0;
    }
    if (run > 0) { const uint32_t e = s_code[tac][0]; put(e & 0xffffu, (int)(e >> 16)); }
    if (fill) atomicOr(wp, acc);                                             // the last word: the next block continues it
}

// ---- byte stuffing and the final files.  The stream of a view is cut into segments of kJpegSeg = 4096 bytes, one CTA of 256
// threads each, 16 bytes (one LDG.128) per thread.
__device__ __forceinline__ uint32_t jpeg_ff_count(uint32_t w) { return (uint32_t)__popc(__vcmpeq4(w, 0xffffffffu)) >> 3; }

// 16 stream bytes from byte k0 (a multiple of 16) as four MSB-first words; the last byte of the stream padded with ones
// (flush_bits); everything at and after n_bytes is zero (the scratch is zeroed up to the next view's base).
__device__ __forceinline__ void jpeg_load16(const uint32_t *__restrict__ s, unsigned long long k0, unsigned long long n_bytes, unsigned long long bits,
                                            uint32_t (&w)[4])
{
    w[0] = w[1] = w[2] = w[3] = 0u;
    if (k0 >= n_bytes) return;
    const uint4 v = *reinterpret_cast<const uint4 *>(s + (k0 >> 2));
    w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    if (k0 + 16 >= n_bytes && (bits & 7ull)) {
        const int last = (int)(n_bytes - 1ull - k0);                         // 0 .. 15
        const uint32_t pad = (0xffu >> (int)(bits & 7ull)) << (24 - 8 * (last & 3));
        if ((last >> 2) == 0) w[0] |= pad; else if ((last >> 2) == 1) w[1] |= pad; else if ((last >> 2) == 2) w[2] |= pad; else w[3] |= pad;
    }
}

// 0xFF bytes per segment: seg_ff[v][seg] (a count, turned into an offset by jpeg_ffscan_kernel)
__global__ void __launch_bounds__(256)
jpeg_ffcount_kernel(const uint32_t *__restrict__ stream, const unsigned long long *__restrict__ word_base, const unsigned long long *__restrict__ view_bits,
                    uint32_t *__restrict__ seg_ff, int max_segs, const int *__restrict__ overflow)
{
    if (*overflow == 1) return;
    __shared__ uint32_t warp_sum[8];
    const int v = blockIdx.y, seg = blockIdx.x, tid = threadIdx.x;
    const unsigned long long bits = view_bits[v], n_bytes = (bits + 7ull) / 8ull;
    if ((unsigned long long)seg * kJpegSeg >= n_bytes) return;               // (uniform over the CTA)
    uint32_t w[4];
    jpeg_load16(stream + word_base[v], (unsigned long long)seg * kJpegSeg + 16ull * tid, n_bytes, bits, w);
    uint32_t c = jpeg_ff_count(w[0]) + jpeg_ff_count(w[1]) + jpeg_ff_count(w[2]) + jpeg_ff_count(w[3]);
    c = __reduce_add_sync(0xffffffffu, c);
    if ((tid & 31) == 0) warp_sum[tid >> 5] = c;
    __syncthreads();
    if (tid == 0) {
        uint32_t t = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += warp_sum[i];
        seg_ff[(size_t)v * max_segs + seg] = t;
    }
}

// Per view (one CTA of 256 threads): exclusive scan of the segments' 0xFF counts in place -> stuffed offset of every segment;
// file_bytes[v] = header + stream + stuffing + EOI.
__global__ void __launch_bounds__(256)
jpeg_ffscan_kernel(const unsigned long long *__restrict__ view_bits, uint32_t *__restrict__ seg_ff, int max_segs, int header_bytes,
                   unsigned long long *__restrict__ file_bytes, int *__restrict__ overflow)
{
    if (*overflow == 1) { if (threadIdx.x == 0) file_bytes[blockIdx.x] = 0; return; }
    __shared__ uint32_t warp_sum[8];
    __shared__ uint32_t carry;
    const int v = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned long long bits = view_bits[v], n_bytes = (bits + 7ull) / 8ull;
    const unsigned long long n_segs_ll = (n_bytes + kJpegSeg - 1) / kJpegSeg;
    uint32_t *out = seg_ff + (size_t)v * max_segs;
    if (n_segs_ll > (unsigned long long)max_segs) {                          // (uniform over the CTA) a view beyond 1.5 x its raw size
        if (tid == 0) { *overflow = 3; file_bytes[v] = 0; }
        return;
    }
    const int n_segs = (int)n_segs_ll;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n_segs; base += 256) {
        const int i = base + tid;
        const uint32_t x = i < n_segs ? out[i] : 0u;
        uint32_t inc = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += y;
        }
        if (lane == 31) warp_sum[warp] = inc;
        __syncthreads();
        uint32_t before = carry + inc - x;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (k < warp) before += warp_sum[k];
        if (i < n_segs) out[i] = before;
        __syncthreads();
        if (tid == 255) carry = before + x;
        __syncthreads();
    }
    if (tid == 0) file_bytes[v] = (unsigned long long)header_bytes + n_bytes + carry + 2ull;
}

// Across views: byte offset of every file in the output (offsets[n] = total); overflow if beyond the capacity.
__global__ void jpeg_offsets_kernel(const unsigned long long *__restrict__ file_bytes, int n, long long *__restrict__ offsets, long long cap_bytes,
                                    int *__restrict__ overflow)
{
    if (threadIdx.x || blockIdx.x) return;
    long long o = 0;
    for (int v = 0; v < n; ++v) { offsets[v] = o; o += (long long)file_bytes[v]; }
    offsets[n] = o;
    if (o > cap_bytes) *overflow = 2;
}

// Final bytes of a view: header | stream with 0xFF 0x00 stuffing | EOI.  One CTA per segment: a block scan of the threads' 0xFF
// counts places every thread's 16 (+ stuffing) bytes; segment 0 also writes the header, the last one EOI.
__global__ void __launch_bounds__(256)
jpeg_pack_kernel(const uint32_t *__restrict__ stream, const unsigned long long *__restrict__ word_base, const unsigned long long *__restrict__ view_bits,
                 const uint32_t *__restrict__ seg_ff, int max_segs, const uint8_t *__restrict__ header, int header_bytes,
                 const long long *__restrict__ offsets, uint8_t *__restrict__ out, const int *__restrict__ overflow)
{
    if (*overflow) return;
    __shared__ uint32_t warp_sum[8];
    const int v = blockIdx.y, seg = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned long long bits = view_bits[v], n_bytes = (bits + 7ull) / 8ull;
    const unsigned long long seg0 = (unsigned long long)seg * kJpegSeg;
    if (seg0 >= n_bytes && seg > 0) return;                                  // (uniform over the CTA)
    uint8_t *o = out + offsets[v];
    if (seg == 0) {
        for (int k = tid; k < header_bytes; k += 256) o[k] = header[k];
        if (n_bytes == 0) {                                                  // (no real image gets here)
            if (tid == 0) { o[header_bytes] = 0xff; o[header_bytes + 1] = 0xd9; }
            return;
        }
    }
    const unsigned long long k0 = seg0 + 16ull * tid;
    uint32_t w[4];
    jpeg_load16(stream + word_base[v], k0, n_bytes, bits, w);
    const uint32_t mine = jpeg_ff_count(w[0]) + jpeg_ff_count(w[1]) + jpeg_ff_count(w[2]) + jpeg_ff_count(w[3]);
    uint32_t inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += y;
    }
    if (lane == 31) warp_sum[warp] = inc;
    __syncthreads();
    uint32_t before = inc - mine;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (k < warp) before += warp_sum[k];
    uint8_t *p = o + header_bytes + k0 + seg_ff[(size_t)v * max_segs + seg] + before;
    if (k0 < n_bytes) {
        const int n_here = (int)min(16ull, n_bytes - k0);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (j < n_here) {
                const uint32_t byte = (w[j >> 2] >> (24 - 8 * (j & 3))) & 0xffu;
                *p++ = (uint8_t)byte;
                if (byte == 0xffu) *p++ = 0;
            }
        }
        if (k0 + 16ull >= n_bytes) { p[0] = 0xff; p[1] = 0xd9; }             // the thread holding the last byte: EOI
    }
}

}  // namespace x360

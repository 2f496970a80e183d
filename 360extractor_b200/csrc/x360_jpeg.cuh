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
//   jpeg_ffscan_kernel   per view: 0xFF bytes per 128-byte segment of the stream, scanned -> stuffed segment offsets, file size
//   jpeg_pack_kernel     header + stream with 0xFF 0x00 stuffing + the 1-bit padding of the last byte + EOI -> final bytes
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace x360 {

constexpr int kJpegMaxHeader = 640;               // SOI .. SOS of a 3-component baseline file with 2 + 4 tables: 623 bytes
constexpr int kJpegSeg = 128;                     // bytes per stuffing segment

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

// One thread per block: grid.x over the blocks of a view (6 per MCU), grid.y = view.
__global__ void __launch_bounds__(128)
jpeg_blocks_kernel(const uint8_t *__restrict__ views, const JpegGeom g, const __grid_constant__ JpegTables tab, int16_t *__restrict__ coef)
{
    const int bi = blockIdx.x * blockDim.x + threadIdx.x;
    if (bi >= g.n_blocks) return;
    const int m = bi / 6, b = bi - 6 * m;
    const int my = m / g.mw, mx = m - my * g.mw;
    const uint8_t *img = views + (size_t)blockIdx.y * g.h * g.w * 3;
    int16_t *out = coef + ((size_t)blockIdx.y * g.n_blocks + bi) * 64;
    int d[64];
    if (b < 4) {
        const int bx = mx * 2 + (b & 1), by = my * 2 + (b >> 1);
        if (by >= g.bh_y || bx >= g.bw_y) {
            // dummy block (jccoefct.c): zeros here; its DC is copied from its predecessor by jpeg_dummy_dc_kernel
#pragma unroll
            for (int k = 0; k < 64; k += 4) *reinterpret_cast<uint2 *>(out + k) = make_uint2(0u, 0u);
            return;
        }
#pragma unroll
        for (int y = 0; y < 8; ++y) {
            const int sy = min(by * 8 + y, g.h - 1);                         // edge expansion: the last row / column repeats
            const uint8_t *row = img + (size_t)sy * g.w * 3;
#pragma unroll
            for (int x = 0; x < 8; ++x) d[8 * y + x] = jpeg_y(row + (size_t)min(bx * 8 + x, g.w - 1) * 3) - 128;
        }
    } else {
        // h2v2 downsampling (jcsample.c): (a + b + c + d + bias) >> 2, bias 1, 2, 1, 2 along a row; columns are replicated before
        // the downsampling; below the image the DOWNSAMPLED rows repeat (jcprepct.c), an odd height repeating its last row once
#pragma unroll 1
        for (int y = 0; y < 8; ++y) {
            const int cy = min(my * 8 + y, g.crows - 1);
            const uint8_t *r0 = img + (size_t)(2 * cy) * g.w * 3, *r1 = img + (size_t)min(2 * cy + 1, g.h - 1) * g.w * 3;
#pragma unroll
            for (int x = 0; x < 8; ++x) {
                const size_t x0 = (size_t)min(mx * 16 + 2 * x, g.w - 1) * 3, x1 = (size_t)min(mx * 16 + 2 * x + 1, g.w - 1) * 3;
                const int bias = (x & 1) ? 2 : 1;
                const int s = b == 4 ? jpeg_cb(r0 + x0) + jpeg_cb(r0 + x1) + jpeg_cb(r1 + x0) + jpeg_cb(r1 + x1)
                                     : jpeg_cr(r0 + x0) + jpeg_cr(r0 + x1) + jpeg_cr(r1 + x0) + jpeg_cr(r1 + x1);
                d[8 * y + x] = ((s + bias) >> 2) - 128;
            }
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
    for (int k = 0; k < 64; k += 4) {
        const uint32_t lo = (uint16_t)q[zz[k]] | ((uint32_t)(uint16_t)q[zz[k + 1]] << 16);
        const uint32_t hi = (uint16_t)q[zz[k + 2]] | ((uint32_t)(uint16_t)q[zz[k + 3]] << 16);
        *reinterpret_cast<uint2 *>(out + k) = make_uint2(lo, hi);
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

// Entropy-coded size of every block, in bits (jchuff.c encode_one_block without the writes)
__global__ void __launch_bounds__(128)
jpeg_bits_kernel(const int16_t *__restrict__ coef, const JpegGeom g, uint32_t *__restrict__ bits)
{
    const int bi = blockIdx.x * blockDim.x + threadIdx.x;
    if (bi >= g.n_blocks) return;
    const int16_t *vc = coef + (size_t)blockIdx.y * g.n_blocks * 64, *c = vc + (size_t)bi * 64;
    const int b = bi % 6, tdc = b < 4 ? 0 : 2, tac = tdc + 1;
    int nb = jpeg_nbits(c[0] - jpeg_pred(vc, bi));
    uint32_t n = c_jpeg_huff.size[tdc][nb] + nb;
    int run = 0;
#pragma unroll 4
    for (int k = 1; k < 64; ++k) {
        const int v = c[k];
        if (v == 0) { ++run; continue; }
        n += (uint32_t)(run >> 4) * c_jpeg_huff.size[tac][0xf0];
        nb = jpeg_nbits(v);
        n += c_jpeg_huff.size[tac][((run & 15) << 4) | nb] + nb;
        run = 0;
    }
    if (run > 0) n += c_jpeg_huff.size[tac][0];
    bits[(size_t)blockIdx.y * g.n_blocks + bi] = n;
}

// Per view (one CTA of 1024 threads): exclusive scan of the block sizes in place; view_bits[v] = total.
__global__ void __launch_bounds__(1024)
jpeg_scan_kernel(uint32_t *__restrict__ bits, int n_items, unsigned long long *__restrict__ totals)
{
    __shared__ uint32_t warp_sum[32];
    __shared__ uint32_t carry;
    uint32_t *v = bits + (size_t)blockIdx.x * n_items;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n_items; base += 1024) {
        const int i = base + tid;
        const uint32_t x = i < n_items ? v[i] : 0u;
        uint32_t inc = x;
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
        const uint32_t before = carry + (warp > 0 ? warp_sum[warp - 1] : 0u) + inc - x;
        if (i < n_items) v[i] = before;
        __syncthreads();
        if (tid == 1023) carry = before + x;
        __syncthreads();
    }
    if (tid == 0) totals[blockIdx.x] = carry;
}

// Across views (one thread: a few thousand views at most): word offset of each view's unstuffed stream.
__global__ void jpeg_bases_kernel(const unsigned long long *__restrict__ view_bits, int n, unsigned long long *__restrict__ word_base,
                                  unsigned long long cap_words, int *__restrict__ overflow)
{
    if (threadIdx.x || blockIdx.x) return;
    unsigned long long w = 0;
    for (int v = 0; v < n; ++v) {
        word_base[v] = w;
        w += (view_bits[v] + 31ull) / 32ull + 1ull;                          // (+1: the emitter may touch the word after the last bit)
    }
    word_base[n] = w;
    if (w > cap_words) *overflow = 1;
}

// OR `len` bits (the low bits of `val`, MSB first) into the word stream at bit position `pos`
__device__ __forceinline__ void jpeg_put(uint32_t *__restrict__ stream, unsigned long long pos, uint64_t val, int len)
{
    // up to 59 bits: at most three 32-bit words
    unsigned long long w = pos >> 5;
    int off = (int)(pos & 31);                                               // bits already used in word w
    while (len > 0) {
        const int room = 32 - off, take = len < room ? len : room;
        const uint32_t chunk = (uint32_t)((val >> (len - take)) & ((take == 32) ? 0xffffffffull : ((1ull << take) - 1ull)));
        atomicOr(stream + w, chunk << (room - take));
        len -= take;
        off = 0;
        ++w;
    }
}

// Huffman codes of every block into the zeroed MSB-first word stream (jchuff.c encode_one_block)
__global__ void __launch_bounds__(128)
jpeg_emit_kernel(const int16_t *__restrict__ coef, const JpegGeom g, const uint32_t *__restrict__ bit_off,
                 const unsigned long long *__restrict__ word_base, uint32_t *__restrict__ stream, const int *__restrict__ overflow)
{
    const int bi = blockIdx.x * blockDim.x + threadIdx.x;
    if (bi >= g.n_blocks || *overflow) return;
    const int16_t *vc = coef + (size_t)blockIdx.y * g.n_blocks * 64, *c = vc + (size_t)bi * 64;
    const int b = bi % 6, tdc = b < 4 ? 0 : 2, tac = tdc + 1;
    uint32_t *s = stream + word_base[blockIdx.y];
    unsigned long long pos = bit_off[(size_t)blockIdx.y * g.n_blocks + bi];
    const int diff = c[0] - jpeg_pred(vc, bi);
    int nb = jpeg_nbits(diff);
    {
        const int sz = c_jpeg_huff.size[tdc][nb];
        const uint64_t bits = ((uint64_t)c_jpeg_huff.code[tdc][nb] << nb) | (uint64_t)((uint32_t)(diff < 0 ? diff - 1 : diff) & ((1u << nb) - 1u));
        jpeg_put(s, pos, bits, sz + nb);
        pos += sz + nb;
    }
    int run = 0;
    for (int k = 1; k < 64; ++k) {
        const int v = c[k];
        if (v == 0) { ++run; continue; }
        while (run > 15) {
            jpeg_put(s, pos, c_jpeg_huff.code[tac][0xf0], c_jpeg_huff.size[tac][0xf0]);
            pos += c_jpeg_huff.size[tac][0xf0];
            run -= 16;
        }
        nb = jpeg_nbits(v);
        const int sym = (run << 4) | nb, sz = c_jpeg_huff.size[tac][sym];
        const uint64_t bits = ((uint64_t)c_jpeg_huff.code[tac][sym] << nb) | (uint64_t)((uint32_t)(v < 0 ? v - 1 : v) & ((1u << nb) - 1u));
        jpeg_put(s, pos, bits, sz + nb);
        pos += sz + nb;
        run = 0;
    }
    if (run > 0) jpeg_put(s, pos, c_jpeg_huff.code[tac][0], c_jpeg_huff.size[tac][0]);
}

__device__ __forceinline__ uint32_t jpeg_stream_byte(const uint32_t *s, unsigned long long i, unsigned long long n_bytes, unsigned long long bits)
{
    uint32_t b = (s[i >> 2] >> (24 - 8 * (int)(i & 3))) & 0xffu;
    if (i == n_bytes - 1 && (bits & 7ull)) b |= 0xffu >> (int)(bits & 7ull);  // flush_bits: the last byte is padded with ones
    return b;
}

// Per view (one CTA): 0xFF bytes per 128-byte segment of the stream, exclusive scan in place -> stuffed offset of every segment;
// file_bytes[v] = header + stream + stuffing + EOI.
__global__ void __launch_bounds__(1024)
jpeg_ffscan_kernel(const uint32_t *__restrict__ stream, const unsigned long long *__restrict__ word_base, const unsigned long long *__restrict__ view_bits,
                   uint32_t *__restrict__ seg_ff, int max_segs, int header_bytes, unsigned long long *__restrict__ file_bytes, int *__restrict__ overflow)
{
    if (*overflow == 1) { if (threadIdx.x == 0) file_bytes[blockIdx.x] = 0; return; }
    __shared__ uint32_t warp_sum[32];
    __shared__ uint32_t carry;
    const int v = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t *s = stream + word_base[v];
    const unsigned long long bits = view_bits[v], n_bytes = (bits + 7ull) / 8ull;
    const int n_segs = (int)((n_bytes + kJpegSeg - 1) / kJpegSeg);
    uint32_t *out = seg_ff + (size_t)v * max_segs;
    if (n_segs > max_segs) {                                                 // (uniform over the CTA) a view beyond 1.5 x its raw size
        if (tid == 0) { *overflow = 3; file_bytes[v] = 0; }
        return;
    }
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n_segs; base += 1024) {
        const int i = base + tid;
        uint32_t x = 0;
        if (i < n_segs) {
            const unsigned long long b0 = (unsigned long long)i * kJpegSeg, b1 = min(b0 + (unsigned long long)kJpegSeg, n_bytes);
            for (unsigned long long k = b0; k < b1; ++k) x += jpeg_stream_byte(s, k, n_bytes, bits) == 0xffu;
        }
        uint32_t inc = x;
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
            warp_sum[lane] = w;
        }
        __syncthreads();
        const uint32_t before = carry + (warp > 0 ? warp_sum[warp - 1] : 0u) + inc - x;
        if (i < n_segs) out[i] = before;
        __syncthreads();
        if (tid == 1023) carry = before + x;
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

// Final bytes of a view: header | stream with 0xFF 0x00 stuffing | EOI.  One thread per segment (+ segment -1: the header, and the
// last one also writes EOI).
__global__ void __launch_bounds__(128)
jpeg_pack_kernel(const uint32_t *__restrict__ stream, const unsigned long long *__restrict__ word_base, const unsigned long long *__restrict__ view_bits,
                 const uint32_t *__restrict__ seg_ff, int max_segs, const uint8_t *__restrict__ header, int header_bytes,
                 const long long *__restrict__ offsets, uint8_t *__restrict__ out, const int *__restrict__ overflow)
{
    if (*overflow) return;
    const int v = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x - 1;   // segment index; -1 = header
    const unsigned long long bits = view_bits[v], n_bytes = (bits + 7ull) / 8ull;
    const int n_segs = (int)((n_bytes + kJpegSeg - 1) / kJpegSeg);
    uint8_t *o = out + offsets[v];
    if (i < 0) {
        for (int k = 0; k < header_bytes; ++k) o[k] = header[k];
        if (n_segs == 0) { o[header_bytes] = 0xff; o[header_bytes + 1] = 0xd9; }
        return;
    }
    if (i >= n_segs) return;
    const uint32_t *s = stream + word_base[v];
    const unsigned long long b0 = (unsigned long long)i * kJpegSeg, b1 = min(b0 + (unsigned long long)kJpegSeg, n_bytes);
    uint8_t *p = o + header_bytes + b0 + seg_ff[(size_t)v * max_segs + i];
    for (unsigned long long k = b0; k < b1; ++k) {
        const uint32_t b = jpeg_stream_byte(s, k, n_bytes, bits);
        *p++ = (uint8_t)b;
        if (b == 0xffu) *p++ = 0;
    }
    if (i == n_segs - 1) { p[0] = 0xff; p[1] = 0xd9; }
}

}  // namespace x360

// x360_jpeg_tables.hpp — host-side constants of the baseline JPEG the reference's save step writes (cv2.imwrite with
// IMWRITE_JPEG_QUALITY: src/utils/file_manager.py:21-32, src/core/processor.py:121-123): ITU-T T.81 Annex K tables, the IJG
// quality scaling (jcparam.c), libjpeg-turbo's reciprocal quantiser constants (jcdctmgr.c compute_reciprocal) and the file
// header (jcmarker.c).  The kernels are in x360_jpeg.cuh.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

#include "x360_jpeg.cuh"

namespace x360 {

namespace jpegk {
const uint8_t std_luma_q[64] = {16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,  14, 13, 16, 24, 40,  57,
                                69, 56, 14, 17, 22,  29,  51,  87,  80, 62, 18, 22, 37,  56,  68,  109, 103, 77, 24, 35, 55,  64,
                                81, 104, 113, 92, 49, 64,  78,  87,  103, 121, 120, 101, 72,  92,  95,  98,  112, 100, 103, 99};
const uint8_t std_chroma_q[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99,
                                  99, 99, 47, 66, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                                  99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};
const uint8_t zigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                            41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                            30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
const uint8_t dc_lum_bits[17] = {0, 0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
const uint8_t dc_chr_bits[17] = {0, 0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
const uint8_t dc_val[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
const uint8_t ac_lum_bits[17] = {0, 0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
const uint8_t ac_lum_val[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1,
    0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26,
    0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56,
    0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85,
    0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa,
    0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6,
    0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9,
    0xfa};
const uint8_t ac_chr_bits[17] = {0, 0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
const uint8_t ac_chr_val[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42,
    0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19,
    0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55,
    0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83,
    0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8,
    0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4,
    0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9,
    0xfa};
}  // namespace jpegk

// jcparam.c: jpeg_quality_scaling + jpeg_add_quant_table(force_baseline); natural order
inline void jpeg_qtables(int quality, uint8_t *luma, uint8_t *chroma)
{
    quality = quality <= 0 ? 1 : (quality > 100 ? 100 : quality);
    const int scale = quality < 50 ? 5000 / quality : 200 - quality * 2;
    for (int i = 0; i < 64; ++i) {
        const long l = ((long)jpegk::std_luma_q[i] * scale + 50L) / 100L, c = ((long)jpegk::std_chroma_q[i] * scale + 50L) / 100L;
        luma[i] = (uint8_t)(l <= 0 ? 1 : (l > 255 ? 255 : l));
        chroma[i] = (uint8_t)(c <= 0 ? 1 : (c > 255 ? 255 : c));
    }
}

// jcdctmgr.c compute_reciprocal for the divisor 8 q (the islow DCT leaves a factor 8)
inline void jpeg_make_tables(int quality, JpegTables *t)
{
    uint8_t q[128];
    jpeg_qtables(quality, q, q + 64);
    for (int i = 0; i < 128; ++i) {
        const unsigned d = 8u * q[i];
        t->q1[i] = d == 1;
        int b = 0;
        for (unsigned v = d; v; v >>= 1) ++b;
        b -= 1;
        int r = 16 + b;
        uint32_t fq = (uint32_t)(((uint64_t)1 << r) / d);
        const uint32_t fr = (uint32_t)(((uint64_t)1 << r) % d);
        uint32_t c = d / 2;
        if (fr == 0) { fq >>= 1; r--; }
        else if (fr <= d / 2U) c++;
        else fq++;
        t->recip[i] = fq;
        t->corr[i] = (uint16_t)c;
        t->shift[i] = (uint8_t)(r - 16);
    }
}

// jchuff.c jpeg_make_c_derived_tbl
inline void jpeg_make_huff(JpegHuff *h)
{
    memset(h, 0, sizeof *h);
    const uint8_t *bits[4] = {jpegk::dc_lum_bits, jpegk::ac_lum_bits, jpegk::dc_chr_bits, jpegk::ac_chr_bits};
    const uint8_t *vals[4] = {jpegk::dc_val, jpegk::ac_lum_val, jpegk::dc_val, jpegk::ac_chr_val};
    for (int t = 0; t < 4; ++t) {
        unsigned code = 0;
        int p = 0;
        for (int l = 1; l <= 16; ++l) {
            for (int i = 0; i < bits[t][l]; ++i, ++p) {
                h->code[t][vals[t][p]] = (uint16_t)code++;
                h->size[t][vals[t][p]] = (uint8_t)l;
            }
            code <<= 1;
        }
    }
}

// jcmarker.c: SOI, APP0 (JFIF 1.01, density 1:1), DQT x 2, SOF0 (4:2:0), DHT x 4, SOS
inline std::vector<uint8_t> jpeg_make_header(int h, int w, int quality)
{
    uint8_t q[128];
    jpeg_qtables(quality, q, q + 64);
    std::vector<uint8_t> o = {0xff, 0xd8, 0xff, 0xe0, 0, 16, 'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0};
    for (int t = 0; t < 2; ++t) {
        o.insert(o.end(), {0xff, 0xdb, 0, 67, (uint8_t)t});
        for (int k = 0; k < 64; ++k) o.push_back(q[64 * t + jpegk::zigzag[k]]);
    }
    o.insert(o.end(), {0xff, 0xc0, 0, 17, 8, (uint8_t)(h >> 8), (uint8_t)h, (uint8_t)(w >> 8), (uint8_t)w, 3, 1, 0x22, 0, 2, 0x11, 1, 3, 0x11, 1});
    const uint8_t *bits[4] = {jpegk::dc_lum_bits, jpegk::ac_lum_bits, jpegk::dc_chr_bits, jpegk::ac_chr_bits};
    const uint8_t *vals[4] = {jpegk::dc_val, jpegk::ac_lum_val, jpegk::dc_val, jpegk::ac_chr_val};
    const uint8_t ids[4] = {0x00, 0x10, 0x01, 0x11};
    for (int t = 0; t < 4; ++t) {
        int cnt = 0;
        for (int l = 1; l <= 16; ++l) cnt += bits[t][l];
        o.insert(o.end(), {0xff, 0xc4, (uint8_t)((19 + cnt) >> 8), (uint8_t)(19 + cnt), ids[t]});
        o.insert(o.end(), bits[t] + 1, bits[t] + 17);
        o.insert(o.end(), vals[t], vals[t] + cnt);
    }
    o.insert(o.end(), {0xff, 0xda, 0, 12, 3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0});
    return o;
}

}  // namespace x360

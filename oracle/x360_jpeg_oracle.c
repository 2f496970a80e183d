/*
 * x360_jpeg_oracle.c — CPU ORACLE, TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * Restates, in plain C, the baseline JPEG encoder that the reference's save step runs:
 *     cv2.imwrite(path, rect_img, [cv2.IMWRITE_JPEG_QUALITY, quality])       src/utils/file_manager.py:21-32,
 *     save_params at src/core/processor.py:121-123 (quality default 95, settings_manager.py:24)
 * The arithmetic lives in a third-party dependency that is absent from /root/reference: OpenCV's bundled
 * libjpeg-turbo (opencv-python >= 4.8.0 in requirements.txt, no lockfile; 4.13.0 with libjpeg-turbo 3.1.2 installed
 * here).  What OpenCV asks of it: JCS_RGB input (BGR swapped), jpeg_set_defaults + jpeg_set_quality(q, TRUE), i.e.
 *   YCbCr 4:2:0 (luma 2x2, chroma 1x1 sampling factors), 8-bit baseline sequential DCT, JDCT_ISLOW,
 *   Annex-K quantisation tables scaled by the IJG quality formula and clamped to 1..255,
 *   the Annex-K "typical" Huffman tables (optimize_coding off), no restart markers, JFIF 1.01 header, density 1:1.
 * Published algorithm restated below (IJG / libjpeg-turbo: jccolor.c, jcsample.c, jfdctint.c, jcdctmgr.c, jchuff.c,
 * jcmarker.c).  PINNED: tests/test_oracle.py compares orc_jpeg_encode byte for byte with cv2.imencode of this container
 * on seeded images (odd sizes, noise, smooth, flat, saturated), and tests/golden/jpeg_cases.json holds cv2's hashes.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- tables (ITU-T T.81 Annex K) ---------------------------------------------------------------------------- */
static const uint8_t kStdLumaQ[64] = {
    16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
    18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
static const uint8_t kStdChromaQ[64] = {
    17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};
static const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                    41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                    30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
static const uint8_t kDcLumBits[17] = {0, 0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
static const uint8_t kDcLumVal[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
static const uint8_t kDcChrBits[17] = {0, 0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
static const uint8_t kDcChrVal[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
static const uint8_t kAcLumBits[17] = {0, 0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
static const uint8_t kAcLumVal[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1,
    0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26,
    0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56,
    0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85,
    0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa,
    0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6,
    0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9,
    0xfa};
static const uint8_t kAcChrBits[17] = {0, 0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
static const uint8_t kAcChrVal[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42,
    0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19,
    0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55,
    0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83,
    0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8,
    0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4,
    0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9,
    0xfa};

/* jcparam.c jpeg_quality_scaling + jpeg_add_quant_table(force_baseline = TRUE): natural (row-major) order */
void orc_jpeg_qtables(int quality, uint8_t *luma64, uint8_t *chroma64)
{
    if (quality <= 0) quality = 1;
    if (quality > 100) quality = 100;
    const int scale = quality < 50 ? 5000 / quality : 200 - quality * 2;
    for (int i = 0; i < 64; ++i) {
        long l = ((long)kStdLumaQ[i] * scale + 50L) / 100L, c = ((long)kStdChromaQ[i] * scale + 50L) / 100L;
        luma64[i] = (uint8_t)(l <= 0 ? 1 : (l > 255 ? 255 : l));
        chroma64[i] = (uint8_t)(c <= 0 ? 1 : (c > 255 ? 255 : c));
    }
}

/* jchuff.c jpeg_make_c_derived_tbl: code and size per symbol */
static void derive(const uint8_t *bits, const uint8_t *vals, uint16_t *code, uint8_t *size)
{
    memset(code, 0, 256 * sizeof(uint16_t));
    memset(size, 0, 256);
    unsigned c = 0;
    int p = 0;
    for (int l = 1; l <= 16; ++l) {
        for (int i = 0; i < bits[l]; ++i, ++p) {
            code[vals[p]] = (uint16_t)c++;
            size[vals[p]] = (uint8_t)l;
        }
        c <<= 1;
    }
}

/* The four derived tables, for the CUDA side's constant memory and its tests: [dc lum, ac lum, dc chr, ac chr][256] */
void orc_jpeg_huffman(uint16_t *code4x256, uint8_t *size4x256)
{
    derive(kDcLumBits, kDcLumVal, code4x256, size4x256);
    derive(kAcLumBits, kAcLumVal, code4x256 + 256, size4x256 + 256);
    derive(kDcChrBits, kDcChrVal, code4x256 + 512, size4x256 + 512);
    derive(kAcChrBits, kAcChrVal, code4x256 + 768, size4x256 + 768);
}

/* jfdctint.c (JDCT_ISLOW), CONST_BITS 13, PASS1_BITS 2; in place on 64 ints (samples - 128); output scaled by 8 */
#define DESCALE(x, n) (((x) + (1 << ((n)-1))) >> (n))
static void fdct_islow(int *d)
{
    enum { C0298 = 2446, C0390 = 3196, C0541 = 4433, C0765 = 6270, C0899 = 7373, C1175 = 9633, C1501 = 12299, C1847 = 15137, C1961 = 16069,
           C2053 = 16819, C2562 = 20995, C3072 = 25172 };
    for (int pass = 0; pass < 2; ++pass) {
        for (int i = 0; i < 8; ++i) {
            int *p = pass == 0 ? d + 8 * i : d + i;
            const int s = pass == 0 ? 1 : 8;
            const int t0 = p[0] + p[7 * s], t7 = p[0] - p[7 * s], t1 = p[s] + p[6 * s], t6 = p[s] - p[6 * s];
            const int t2 = p[2 * s] + p[5 * s], t5 = p[2 * s] - p[5 * s], t3 = p[3 * s] + p[4 * s], t4 = p[3 * s] - p[4 * s];
            const int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
            const int sh = pass == 0 ? 13 - 2 : 13 + 2;
            if (pass == 0) {
                p[0] = (t10 + t11) << 2;
                p[4 * s] = (t10 - t11) << 2;
            } else {
                p[0] = DESCALE(t10 + t11, 2);
                p[4 * s] = DESCALE(t10 - t11, 2);
            }
            int z1 = (t12 + t13) * C0541;
            p[2 * s] = DESCALE(z1 + t13 * C0765, sh);
            p[6 * s] = DESCALE(z1 + t12 * (-C1847), sh);
            z1 = t4 + t7;
            int z2 = t5 + t6, z3 = t4 + t6, z4 = t5 + t7;
            const int z5 = (z3 + z4) * C1175;
            const int m4 = t4 * C0298, m5 = t5 * C2053, m6 = t6 * C3072, m7 = t7 * C1501;
            z1 *= -C0899; z2 *= -C2562; z3 *= -C1961; z4 *= -C0390;
            z3 += z5; z4 += z5;
            p[7 * s] = DESCALE(m4 + z1 + z3, sh);
            p[5 * s] = DESCALE(m5 + z2 + z4, sh);
            p[3 * s] = DESCALE(m6 + z2 + z3, sh);
            p[s] = DESCALE(m7 + z1 + z4, sh);
        }
    }
}

/* jcdctmgr.c compute_reciprocal + quantize (libjpeg-turbo: division by multiplication with a 16-bit reciprocal, built
 * to equal (|x| + d/2) / d for every 16-bit |x|); divisor = 8 q (the islow DCT leaves a factor 8) */
typedef struct { uint32_t recip, corr; int shift; } Recip;
static Recip make_recip(unsigned divisor)
{
    Recip r;
    if (divisor == 1) { r.recip = 1; r.corr = 0; r.shift = -16; return r; }
    int b = 0;
    for (unsigned v = divisor; v; v >>= 1) ++b;       /* flss */
    b -= 1;
    int rr = 16 + b;
    uint32_t fq = (uint32_t)(((uint64_t)1 << rr) / divisor), fr = (uint32_t)(((uint64_t)1 << rr) % divisor);
    uint32_t c = divisor / 2;
    if (fr == 0) { fq >>= 1; rr--; }
    else if (fr <= divisor / 2U) c++;
    else fq++;
    r.recip = fq; r.corr = c; r.shift = rr - 16;
    return r;
}
static int quantize_one(int v, const Recip *r)
{
    uint32_t t = (uint32_t)(v < 0 ? -v : v);
    uint32_t q = ((uint32_t)(t + r->corr) * r->recip) >> (r->shift + 16);      /* UDCTELEM2 = 32-bit unsigned */
    q &= 0xffffu;                                     /* DCTELEM is 16 bits wide */
    return v < 0 ? -(int)(int16_t)q : (int)(int16_t)q;
}

/* ---- bit writer with byte stuffing (jchuff.c emit_bits / flush_bits) --------------------------------------- */
typedef struct { uint8_t *p; size_t n, cap; uint64_t acc; int bits; } BitW;
static void put_byte(BitW *w, uint8_t b)
{
    if (w->n < w->cap) w->p[w->n] = b;
    w->n++;
}
static void put_bits(BitW *w, unsigned code, int size)
{
    if (size == 0) return;
    w->acc = (w->acc << size) | (code & ((1u << size) - 1u));
    w->bits += size;
    while (w->bits >= 8) {
        const uint8_t b = (uint8_t)(w->acc >> (w->bits - 8));
        put_byte(w, b);
        if (b == 0xff) put_byte(w, 0);
        w->bits -= 8;
    }
}
static int nbits_of(int v)
{
    int n = 0;
    for (unsigned a = (unsigned)(v < 0 ? -v : v); a; a >>= 1) ++n;
    return n;
}

/*
 * Quantised coefficients of an image in the order and layout the CUDA side produces: per MCU (16 x 16 pixels, row-major
 * over the MCU grid) 6 blocks Y00 Y01 Y10 Y11 Cb Cr of 64 int16 in ZIGZAG order.  bgr: uint8 [h][w][3].
 * coef: int16 [mcus][6][64] (may be NULL).  Returns the number of MCUs.
 */
long orc_jpeg_coefficients(const uint8_t *bgr, int h, int w, int quality, int16_t *coef)
{
    const int mw = (w + 15) / 16, mh = (h + 15) / 16;
    if (!coef) return (long)mw * mh;
    uint8_t ql[64], qc[64];
    orc_jpeg_qtables(quality, ql, qc);
    Recip rl[64], rc[64];
    for (int i = 0; i < 64; ++i) { rl[i] = make_recip(8u * ql[i]); rc[i] = make_recip(8u * qc[i]); }
    for (int my = 0; my < mh; ++my)
        for (int mx = 0; mx < mw; ++mx) {
            int Y[16][16];
            for (int y = 0; y < 16; ++y)
                for (int x = 0; x < 16; ++x) {
                    /* edge expansion (jcprepct.c expand_bottom_edge, jcsample.c expand_right_edge): replicate the last row / column */
                    const int sy = my * 16 + y < h ? my * 16 + y : h - 1, sx = mx * 16 + x < w ? mx * 16 + x : w - 1;
                    const uint8_t *p = bgr + ((size_t)sy * w + sx) * 3;
                    /* jccolor.c rgb_ycc_convert, SCALEBITS 16 */
                    Y[y][x] = (19595 * p[2] + 38470 * p[1] + 7471 * p[0] + 32768) >> 16;
                }
/* chroma of the pixel at MCU-local (row, col); rows may lie above this MCU row (bottom padding), columns are clamped */
#define PIX(yy, xx) (bgr + ((size_t)(my * 16 + (yy)) * w + (mx * 16 + (xx) < w ? mx * 16 + (xx) : w - 1)) * 3)
#define CbF(yy, xx) ((-11059 * PIX(yy, xx)[2] - 21709 * PIX(yy, xx)[1] + 32768 * PIX(yy, xx)[0] + (128 << 16) + 32767) >> 16)
#define CrF(yy, xx) ((32768 * PIX(yy, xx)[2] - 27439 * PIX(yy, xx)[1] - 5329 * PIX(yy, xx)[0] + (128 << 16) + 32767) >> 16)
            int blk[6][64];
            for (int b = 0; b < 4; ++b)
                for (int y = 0; y < 8; ++y)
                    for (int x = 0; x < 8; ++x) blk[b][8 * y + x] = Y[(b >> 1) * 8 + y][(b & 1) * 8 + x] - 128;
            /* jcprepct.c pre_process_data: at the bottom of the image the full-resolution rows are padded only to a whole row
             * group (2 rows: an odd height repeats its last row once); below that the DOWNSAMPLED rows are padded by repeating
             * the last downsampled row.  (Horizontally the input columns are replicated before the downsampling, as above.) */
            const int crows = (h + 1) / 2;
            for (int y = 0; y < 8; ++y) {
                const int cy = my * 8 + y < crows ? my * 8 + y : crows - 1;     /* downsampled row this block row shows */
                const int y0 = 2 * cy - my * 16, y1 = (2 * cy + 1 < h ? 2 * cy + 1 : h - 1) - my * 16;   /* its two source rows, MCU-local */
                for (int x = 0; x < 8; ++x) {
                    const int bias = (x & 1) ? 2 : 1;          /* jcsample.c h2v2_downsample: bias 1, 2, 1, 2, ... along a row */
                    blk[4][8 * y + x] = ((CbF(y0, 2 * x) + CbF(y0, 2 * x + 1) + CbF(y1, 2 * x) + CbF(y1, 2 * x + 1) + bias) >> 2) - 128;
                    blk[5][8 * y + x] = ((CrF(y0, 2 * x) + CrF(y0, 2 * x + 1) + CrF(y1, 2 * x) + CrF(y1, 2 * x + 1) + bias) >> 2) - 128;
                }
            }
            int16_t *out = coef + ((size_t)my * mw + mx) * 6 * 64;
            /* jccoefct.c compress_data: luma blocks that lie wholly beyond the image's last block column / row (sizes that are not
             * multiples of 16) are DUMMY blocks — all AC zero, DC copied from the block before them in the MCU: the block to the
             * left at the right edge; block [first of the row] - 1 for a whole dummy row at the bottom */
            const int bw_y = (w + 7) / 8, bh_y = (h + 7) / 8;
            for (int b = 0; b < 6; ++b) {
                const int bx = mx * 2 + (b & 1), by = my * 2 + (b >> 1);
                if (b < 4 && (by >= bh_y || bx >= bw_y)) {
                    memset(out + b * 64, 0, 64 * sizeof(int16_t));
                    out[b * 64] = by >= bh_y ? out[((b & ~1) - 1) * 64] : out[(b - 1) * 64];
                    continue;
                }
                fdct_islow(blk[b]);
                const Recip *r = b < 4 ? rl : rc;
                for (int k = 0; k < 64; ++k) out[b * 64 + k] = (int16_t)quantize_one(blk[b][kZigzag[k]], &r[kZigzag[k]]);
            }
        }
    return (long)mw * mh;
}

static size_t put_marker_segment(uint8_t *dst, size_t n, size_t cap, const uint8_t *src, size_t len)
{
    for (size_t i = 0; i < len; ++i, ++n)
        if (n < cap) dst[n] = src[i];
    return n;
}

/* The header cv2 / libjpeg-turbo writes (jcmarker.c): SOI, APP0 JFIF 1.01, DQT x2, SOF0, DHT x4, SOS.  Returns its length. */
size_t orc_jpeg_header(int h, int w, int quality, uint8_t *dst, size_t cap)
{
    uint8_t ql[64], qc[64], seg[512];
    orc_jpeg_qtables(quality, ql, qc);
    size_t n = 0, m;
    const uint8_t soi_app0[] = {0xff, 0xd8, 0xff, 0xe0, 0, 16, 'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0};
    n = put_marker_segment(dst, n, cap, soi_app0, sizeof soi_app0);
    for (int t = 0; t < 2; ++t) {
        m = 0;
        seg[m++] = 0xff; seg[m++] = 0xdb; seg[m++] = 0; seg[m++] = 67; seg[m++] = (uint8_t)t;
        for (int k = 0; k < 64; ++k) seg[m++] = (t ? qc : ql)[kZigzag[k]];
        n = put_marker_segment(dst, n, cap, seg, m);
    }
    const uint8_t sof[] = {0xff, 0xc0, 0, 17, 8, (uint8_t)(h >> 8), (uint8_t)h, (uint8_t)(w >> 8), (uint8_t)w, 3, 1, 0x22, 0, 2, 0x11, 1, 3, 0x11, 1};
    n = put_marker_segment(dst, n, cap, sof, sizeof sof);
    const uint8_t *bits[4] = {kDcLumBits, kAcLumBits, kDcChrBits, kAcChrBits};
    const uint8_t *vals[4] = {kDcLumVal, kAcLumVal, kDcChrVal, kAcChrVal};
    const uint8_t ids[4] = {0x00, 0x10, 0x01, 0x11};
    for (int t = 0; t < 4; ++t) {
        int cnt = 0;
        for (int l = 1; l <= 16; ++l) cnt += bits[t][l];
        m = 0;
        seg[m++] = 0xff; seg[m++] = 0xc4; seg[m++] = (uint8_t)((19 + cnt) >> 8); seg[m++] = (uint8_t)(19 + cnt); seg[m++] = ids[t];
        for (int l = 1; l <= 16; ++l) seg[m++] = bits[t][l];
        for (int i = 0; i < cnt; ++i) seg[m++] = vals[t][i];
        n = put_marker_segment(dst, n, cap, seg, m);
    }
    const uint8_t sos[] = {0xff, 0xda, 0, 12, 3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0};
    n = put_marker_segment(dst, n, cap, sos, sizeof sos);
    return n;
}

/* Whole file: header + entropy-coded scan + EOI.  Returns the length (which may exceed cap: nothing is written past cap). */
size_t orc_jpeg_encode(const uint8_t *bgr, int h, int w, int quality, uint8_t *dst, size_t cap)
{
    const long mcus = orc_jpeg_coefficients(bgr, h, w, quality, NULL);
    int16_t *coef = (int16_t *)malloc((size_t)mcus * 6 * 64 * sizeof(int16_t));
    if (!coef) return 0;
    orc_jpeg_coefficients(bgr, h, w, quality, coef);
    uint16_t code[4][256];
    uint8_t size[4][256];
    orc_jpeg_huffman(&code[0][0], &size[0][0]);
    BitW bw = {dst, orc_jpeg_header(h, w, quality, dst, cap), cap, 0, 0};
    int last_dc[3] = {0, 0, 0};
    for (long m = 0; m < mcus; ++m)
        for (int b = 0; b < 6; ++b) {
            const int16_t *c = coef + ((size_t)m * 6 + b) * 64;
            const int comp = b < 4 ? 0 : b - 3, tdc = b < 4 ? 0 : 2, tac = tdc + 1;
            /* jchuff.c encode_one_block */
            int diff = c[0] - last_dc[comp];
            last_dc[comp] = c[0];
            int nb = nbits_of(diff);
            put_bits(&bw, code[tdc][nb], size[tdc][nb]);
            if (nb) put_bits(&bw, (unsigned)(diff < 0 ? diff - 1 : diff), nb);
            int run = 0;
            for (int k = 1; k < 64; ++k) {
                const int v = c[k];
                if (v == 0) { ++run; continue; }
                while (run > 15) { put_bits(&bw, code[tac][0xf0], size[tac][0xf0]); run -= 16; }
                nb = nbits_of(v);
                const int sym = (run << 4) | nb;
                put_bits(&bw, code[tac][sym], size[tac][sym]);
                put_bits(&bw, (unsigned)(v < 0 ? v - 1 : v), nb);
                run = 0;
            }
            if (run > 0) put_bits(&bw, code[tac][0], size[tac][0]);
        }
    put_bits(&bw, 0x7f, 7);                           /* flush_bits: pad the last byte with ones */
    bw.bits = 0;
    put_byte(&bw, 0xff);
    put_byte(&bw, 0xd9);
    free(coef);
    return bw.n;
}

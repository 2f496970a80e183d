// x360_tables.hpp — host-side constant tables owned by a plan.
//
// lanczos4_q15_table: the int16 [32*32][8][8] weight table cv2.remap uses for INTER_LANCZOS4
// (third-party OpenCV, reached from src/core/processor.py:334).  Published scheme: 1-D
// Lanczos-4 coefficients for t = k/32 evaluated with fp64 trig and stored as float32,
// normalised by a float32 reciprocal of their float32 sum; 2-D entries are the float32 products
// scaled by 2^15 and rounded half-to-even; entries that do not sum to 2^15 get the residue
// added to the smallest / subtracted from the largest of the four centre taps (4..5, 4..5).
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <vector>

namespace x360 {

inline void lanczos4_kernel_1d(float t, float k[8])
{
    if (t < FLT_EPSILON) {
        for (int i = 0; i < 8; ++i) k[i] = (i == 3) ? 1.f : 0.f;
        return;
    }
    const double pi = 3.14159265358979323846, h = 0.70710678118654752440;
    // (sin, cos) mixing factors of the angle-addition expansion at quarter-pi steps
    const double mix[8][2] = {{1, 0}, {-h, -h}, {0, 1}, {h, -h}, {-1, 0}, {h, h}, {0, -1}, {-h, h}};
    const double base = -(t + 3) * pi * 0.25;
    const double sb = std::sin(base), cb = std::cos(base);
    float total = 0.f;
    for (int i = 0; i < 8; ++i) {
        const double ang = -(t + 3 - i) * pi * 0.25;
        k[i] = (float)((mix[i][0] * sb + mix[i][1] * cb) / (ang * ang));
        total += k[i];
    }
    const float inv = 1.f / total;
    for (int i = 0; i < 8; ++i) k[i] *= inv;
}

inline std::vector<int16_t> lanczos4_q15_table()
{
    float k1[32][8];
    for (int q = 0; q < 32; ++q) lanczos4_kernel_1d((float)q * (1.f / 32), k1[q]);
    std::vector<int16_t> tab(32 * 32 * 64);
    for (int fy = 0; fy < 32; ++fy)
        for (int fx = 0; fx < 32; ++fx) {
            int16_t *e = &tab[(size_t)(fy * 32 + fx) * 64];
            int total = 0;
            for (int r = 0; r < 8; ++r)
                for (int c = 0; c < 8; ++c) {
                    const float p = k1[fy][r] * k1[fx][c];
                    long q = std::lrintf(p * 32768.f);          // nearest-even in the default FP mode
                    q = q < -32768 ? -32768 : (q > 32767 ? 32767 : q);
                    e[r * 8 + c] = (int16_t)q;
                    total += (int)q;
                }
            const int residue = total - 32768;
            if (residue != 0) {
                int lo = 4 * 8 + 4, hi = 4 * 8 + 4;             // scan the 2x2 block (4..5, 4..5)
                for (int r = 4; r <= 5; ++r)
                    for (int c = 4; c <= 5; ++c) {
                        const int at = r * 8 + c;
                        if (e[at] < e[lo]) lo = at;
                        else if (e[at] > e[hi]) hi = at;
                    }
                const int at = residue < 0 ? hi : lo;
                e[at] = (int16_t)(e[at] - residue);
            }
        }
    return tab;
}

// The same table as its factors: the 32 x 8 float32 1-D coefficients and, per (fy, fx), the one tap the sum-to-2^15 fix-up
// changed — position p in bits 0-1 (tap (4 + (p >> 1), 4 + (p & 1))), residue in bits 16-31 (signed).  The kernel computes the
// 64 weights from these (x360_tiled.cuh, lanczos_pairs); tests rebuild the table from them and compare.
inline void lanczos4_factors(std::vector<float> &k1d, std::vector<uint32_t> &fix)
{
    k1d.assign(32 * 8, 0.f);
    for (int q = 0; q < 32; ++q) lanczos4_kernel_1d((float)q * (1.f / 32), &k1d[(size_t)q * 8]);
    const std::vector<int16_t> tab = lanczos4_q15_table();
    fix.assign(1024, 0u);
    for (int fy = 0; fy < 32; ++fy)
        for (int fx = 0; fx < 32; ++fx) {
            const int16_t *e = &tab[(size_t)(fy * 32 + fx) * 64];
            for (int pos = 0; pos < 4; ++pos) {
                const int r = 4 + (pos >> 1), c = 4 + (pos & 1);
                long q = std::lrintf(k1d[(size_t)fy * 8 + r] * k1d[(size_t)fx * 8 + c] * 32768.f);
                q = q < -32768 ? -32768 : (q > 32767 ? 32767 : q);
                const int delta = (int)e[r * 8 + c] - (int)q;
                if (delta != 0) fix[(size_t)(fy * 32 + fx)] = (uint32_t)pos | ((uint32_t)(uint16_t)(int16_t)delta << 16);
            }
        }
}

}  // namespace x360

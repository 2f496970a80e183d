// x360_misc.cuh — the small entry points around the hot path: materialised maps,
// explicit-map remap, stand-alone blur sums.  These exist so that every reference call
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

}  // namespace x360

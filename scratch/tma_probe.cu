// Probe: 3-D TMA box load of a uint8 [F][H][3W] tensor, descriptor passed as __grid_constant__.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
struct TmapSet { CUtensorMap m[5]; };
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ TmapSet tms, int wsel, int x, int y, int z, uint8_t *out, int dyn_off)
{
    extern __shared__ __align__(128) uint8_t dsm[];
    __shared__ unsigned long long bar;
    uint8_t *raw = dsm + dyn_off;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t bw = 128 + 32 * wsel;
    if (threadIdx.x == 32) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(8 * bw) : "memory");
        const CUtensorMap *tm = &tms.m[wsel];
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(smem_u32(raw)), "l"(tm), "r"(x), "r"(y), "r"(z), "r"(smem_u32(&bar)) : "memory");
    }
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    } while (!done);
    for (int i = threadIdx.x; i < 8 * bw; i += blockDim.x) out[i] = raw[i];
}
int main(int argc, char **argv)
{
    const int W = 192, H = 96, F = 2;
    std::vector<uint8_t> h((size_t)F * H * W * 3);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)(i * 7 + (i >> 8));
    uint8_t *d, *dout;
    cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    cudaMalloc(&dout, 8 * 256);
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    auto enc = (CUresult(*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                            const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill))p;
    TmapSet tms; memset(&tms, 0, sizeof tms);
    const cuuint64_t dims[3] = {(cuuint64_t)W * 3, H, F};
    const cuuint64_t strides[2] = {(cuuint64_t)W * 3, (cuuint64_t)W * 3 * H};
    const cuuint32_t estr[3] = {1, 1, 1};
    for (int k = 0; k < 5; ++k) {
        const cuuint32_t box[3] = {(cuuint32_t)(128 + 32 * k), 8, 1};
        CUresult r = enc(&tms.m[k], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode k=%d -> %d\n", k, (int)r);
    }
    const int cases[][5] = {{0, 0, 0, 0, 0}, {0, 16, 3, 1, 0}, {1, 0, 0, 0, 0}, {1, 32, 3, 1, 2048}, {4, 400, 90, 1, 1024}, {2, 48, 5, 0, 1024}, {0, 12, 3, 0, 0}};
    for (auto &c : cases) {
        cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
        probe<<<1, 64, 65536>>>(tms, c[0], c[1], c[2], c[3], dout, c[4]);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<uint8_t> o(8 * 256);
        cudaMemcpy(o.data(), dout, o.size(), cudaMemcpyDeviceToHost);
        const int bw = 128 + 32 * c[0];
        int bad = 0;
        for (int r = 0; r < 8; ++r) for (int b = 0; b < bw; ++b) {
            const int yy = c[2] + r, xx = c[1] + b;
            const uint8_t want = (yy < H && xx < W * 3) ? h[((size_t)c[3] * H + yy) * W * 3 + xx] : 0;
            bad += (o[r * bw + b] != want);
        }
        printf("case wsel=%d x=%d y=%d z=%d off=%d: %s, mismatches %d\n", c[0], c[1], c[2], c[3], c[4], cudaGetErrorString(e), bad);
        if (e != cudaSuccess) { break; }
    }
    return 0;
}

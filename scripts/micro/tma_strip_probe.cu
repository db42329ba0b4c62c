// Probe: TMA tile-mode loads of an image strip with a box wider than the tensor row and negative start coordinates (zero fill).
// usage: tma_strip_probe <rank 2|3> <boxW> <x0> <y0>    prints the strip's checksum vs the CPU's, or the CUDA error
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int RANK>
__global__ void k(const __grid_constant__ CUtensorMap tm, float* out, int n, int x0, int y0, int img, uint32_t bytes) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(bytes) : "memory");
        if (RANK == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(s32(smem)),
                         "l"(reinterpret_cast<uint64_t>(&tm)), "r"(s32(&bar)), "r"(x0), "r"(y0), "r"(img) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(s32(smem)),
                         "l"(reinterpret_cast<uint64_t>(&tm)), "r"(s32(&bar)), "r"(x0), "r"(y0) : "memory");
    }
    uint32_t done = 0;
    long spins = 0;
    while (!done && spins < 20000000) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(s32(&bar)), "r"(0u) : "memory");
        ++spins;
    }
    if (!done && threadIdx.x == 0) out[n] = -1.0f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = reinterpret_cast<float*>(smem)[i];
}
int main(int argc, char** argv) {
    const int rank = atoi(argv[1]), boxW = atoi(argv[2]), x0 = atoi(argv[3]), y0 = atoi(argv[4]);
    const int W = 120, H = 212, N = 3, boxH = 7, img = 1;
    std::vector<float> h((size_t)N * H * W);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 9973) + 1.0f;
    float *d, *out;
    cudaMalloc(&d, h.size() * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    const int n = boxW * boxH;
    cudaMalloc(&out, (n + 1) * 4);
    cudaMemset(out, 0, (n + 1) * 4);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    CUtensorMap tm;
    const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4};
    const cuuint32_t box[3] = {(cuuint32_t)boxW, (cuuint32_t)boxH, 1}, estr[3] = {1, 1, 1};
    CUresult r = ((EncodeTiledFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, rank == 3 ? d : d + (size_t)img * H * W, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("rank %d boxW %d x0 %d y0 %d: encode failed %d\n", rank, boxW, x0, y0, (int)r); return 0; }
    if (rank == 3) k<3><<<1, 128, n * 4 + 1024>>>(tm, out, n, x0, y0, img, (uint32_t)n * 4);
    else k<2><<<1, 128, n * 4 + 1024>>>(tm, out, n, x0, y0, img, (uint32_t)n * 4);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("rank %d boxW %d x0 %d y0 %d: %s\n", rank, boxW, x0, y0, cudaGetErrorString(e)); return 0; }
    std::vector<float> o(n + 1);
    cudaMemcpy(o.data(), out, (n + 1) * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int y = 0; y < boxH; ++y)
        for (int x = 0; x < boxW; ++x) {
            const int gx = x0 + x, gy = y0 + y;
            const float want = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? h[((size_t)img * H + gy) * W + gx] : 0.0f;
            if (o[y * boxW + x] != want) ++bad;
        }
    printf("rank %d boxW %d x0 %d y0 %d: timeout=%d wrong=%d of %d\n", rank, boxW, x0, y0, o[n] < 0, bad, n);
    return 0;
}

// umma_probe.cu — single-CTA tcgen05.mma (kind::tf32) probe: D[128,N] = A[128,K] * B[N,K]^T with both operands K-major in the
// canonical no-swizzle ("interleave") shared-memory layout, accumulators in TMEM, read back with tcgen05.ld.  Validates the
// descriptor encoding used by agx_mlp.cu's tcgen05 forward path against a CPU reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/micro/umma_probe scripts/micro/umma_probe.cu && scripts/micro/umma_probe
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

constexpr int M = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// canonical K-major no-swizzle layout, in floats: [K/4 chunks][rows/8 groups][8 rows][4 floats]
__host__ __device__ inline int canon(int r, int k, int rows) { return (((k >> 2) * (rows >> 3) + (r >> 3)) * 8 + (r & 7)) * 4 + (k & 3); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // version = 1 (Blackwell)
    return d;                 // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
}

template <int N, int K>
__global__ void __launch_bounds__(128) probe(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int swap_lbo_sbo) {
    extern __shared__ __align__(128) float smem[];
    float* sA = smem;            // M*K floats
    float* sB = smem + M * K;    // N*K floats
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < M * K; i += 128) sA[canon(i / K, i % K, M)] = A[i];
    for (int i = tid; i < N * K; i += 128) sB[canon(i / K, i % K, N)] = B[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes → visible to the tensor core (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t lboA = (M / 8) * 128, lboB = (N / 8) * 128, sbo = 128;
        for (int ks = 0; ks < K / 8; ++ks) {
            const uint32_t aaddr = smem_u32(sA) + ks * 2 * lboA, baddr = smem_u32(sB) + ks * 2 * lboB;
            const uint64_t da = swap_lbo_sbo ? make_desc(aaddr, sbo, lboA) : make_desc(aaddr, lboA, sbo);
            const uint64_t db = swap_lbo_sbo ? make_desc(baddr, sbo, lboB) : make_desc(baddr, lboB, sbo);
            const uint32_t acc = ks > 0 ? 1u : 0u;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // wait for the MMAs (phase 0)
    asm volatile(
        "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(&bar)), "r"(0u)
        : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // each warp reads its 32 TMEM lanes (rows 32w .. 32w+31), 8 columns at a time
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; ++j) D[tid * N + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

template <int N, int K>
int run(int swap) {
    float *hA = (float*)malloc(M * K * 4), *hB = (float*)malloc(N * K * 4), *hD = (float*)malloc(M * N * 4);
    for (int i = 0; i < M * K; ++i) hA[i] = (float)((i * 37 % 17) - 8) / 8.0f;     // exactly representable in tf32
    for (int i = 0; i < N * K; ++i) hB[i] = (float)((i * 53 % 13) - 6) / 4.0f;
    float *dA, *dB, *dD;
    cudaMalloc(&dA, M * K * 4); cudaMalloc(&dB, N * K * 4); cudaMalloc(&dD, M * N * 4);
    cudaMemcpy(dA, hA, M * K * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, N * K * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, M * N * 4);
    const int smem = (M * K + N * K) * 4;
    cudaFuncSetAttribute(probe<N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<N, K><<<1, 128, smem>>>(dA, dB, dD, swap);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d K=%d swap=%d: CUDA error %s\n", N, K, swap, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(hD, dD, M * N * 4, cudaMemcpyDeviceToHost);
    double worst = 0; int bad = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < K; ++k) ref += (double)hA[m * K + k] * hB[n * K + k];
            double err = fabs(ref - hD[m * N + n]);
            if (err > worst) worst = err;
            if (err > 1e-3) ++bad;
        }
    printf("N=%d K=%d swap=%d: worst abs err %.3e, %d / %d wrong; D[0,0..3] = %g %g %g %g\n", N, K, swap, worst, bad, M * N, hD[0], hD[1], hD[2], hD[3]);
    return bad != 0;
}

int main() {
    int rc = 0;
    rc |= run<64, 32>(0);
    if (rc) rc = run<64, 32>(1);
    run<128, 64>(0);
    run<64, 128>(0);
    run<16, 64>(0);
    return 0;
}

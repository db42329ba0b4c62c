// umma_layout_probe.cu — which shared-memory word does tcgen05.mma (kind::tf32, no swizzle) read for element (n, k) of its B
// operand (or (m, k) of A), as a function of the descriptor's LBO / SBO and the instruction descriptor's major bit?
// Method: ONE instruction (K = 8).  The probed operand's region holds a code of its own word index (two passes: index % 128
// and index / 128, both exact in tf32); the other operand is one-hot (row m selects k = m % 8), so D[m][n] = B_hw[n][m % 8]
// (or D[m][n] = A_hw[m][n % 8] when probing A).  The decoded index table is printed as a fit
//      word(n, k) = c_n4 * (n / 4) + c_n1 * (n % 4) + c_k8.. etc.   — simply listed for the first few n and k.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/micro/umma_layout_probe scripts/micro/umma_layout_probe.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
           ((uint64_t)1 << 46);
}
constexpr int kWords = 12288;  // 48 KB probed region

// probe_a = 0: probe B (N x 8), A one-hot K-major.  probe_a = 1: probe A (128 x 8), B one-hot K-major (N = 8: D[m][n] = A_hw[m][n]).
__global__ void __launch_bounds__(128) probe(float* __restrict__ D, int N, int mn_major, uint32_t lbo, uint32_t sbo, int pass, int probe_a) {
    extern __shared__ __align__(1024) float smem[];
    float* region = smem;               // the probed operand reads from here
    float* onehot = smem + kWords;      // 128 x 8 (or N x 8) K-major one-hot operand: [k/4][rows][k%4]
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < kWords; i += 128) region[i] = (float)(pass == 0 ? (i % 128) : (i / 128));
    const int oh_rows = probe_a ? N : 128;
    for (int i = tid; i < oh_rows * 8; i += 128) {
        const int r = i / 8, k = i % 8;
        onehot[((k >> 2) * oh_rows + r) * 4 + (k & 3)] = (k == (r % 8)) ? 1.0f : 0.0f;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    if (tid == 0) {
        const int M = 128;
        uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        uint64_t da, db;
        if (probe_a) {
            idesc |= (uint32_t)mn_major << 15;
            da = make_desc(smem_u32(region), lbo, sbo);
            db = make_desc(smem_u32(onehot), (uint32_t)N * 16, 128);
        } else {
            idesc |= (uint32_t)mn_major << 16;
            da = make_desc(smem_u32(onehot), 128 * 16, 128);
            db = make_desc(smem_u32(region), lbo, sbo);
        }
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem),
                     "l"(da), "l"(db), "r"(idesc), "r"(0u)
                     : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(&bar)), "r"(0u)
                 : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
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

void run(int N, int mn_major, uint32_t lbo, uint32_t sbo, int probe_a) {
    float* dD;
    cudaMalloc(&dD, 128 * N * 4);
    float* h[2];
    const int smem = (kWords + 128 * 8) * 4;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int pass = 0; pass < 2; ++pass) {
        h[pass] = (float*)malloc(128 * N * 4);
        probe<<<1, 128, smem>>>(dD, N, mn_major, lbo, sbo, pass, probe_a);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); exit(1); }
        cudaMemcpy(h[pass], dD, 128 * N * 4, cudaMemcpyDeviceToHost);
    }
    printf("probe %s  N=%d  %s-major  LBO=%u SBO=%u : word index read for (row, k)\n", probe_a ? "A" : "B", N, mn_major ? "MN" : "K", lbo, sbo);
    const int rows = probe_a ? 128 : N;
    const int show[] = {0, 1, 2, 3, 4, 5, 7, 8, 9, 12, 15, 16, 31, 32, 63, 64, 127};
    for (int s = 0; s < (int)(sizeof(show) / sizeof(int)); ++s) {
        const int r = show[s];
        if (r >= rows) continue;
        printf("   row %3d:", r);
        for (int k = 0; k < 8; ++k) {
            // probing B: D[m][n] = B_hw[n][m % 8] -> take m = k;  probing A: D[m][n] = A_hw[m][n % 8] -> take n = k
            const int idx = probe_a ? (r * N + k) : (k * N + r);
            const int word = (int)lrintf(h[1][idx]) * 128 + (int)lrintf(h[0][idx]);
            printf(" %6d", word);
        }
        printf("\n");
    }
}

int main() {
    run(64, 0, 64 * 16, 128, 0);      // reference: K-major B, rows = 64: word = ((k/4)*64 + n)*4 + k%4
    run(64, 1, 128, 2048, 0);         // MN-major B: LBO = 128 (8 k's), SBO = 2048 (4 n's)
    run(64, 1, 2048, 128, 0);         // swapped
    run(64, 1, 128, 512, 0);
    run(64, 1, 512, 128, 0);
    run(64, 1, 1024, 256, 0);
    run(16, 1, 128, 2048, 1);         // MN-major A
    run(16, 1, 2048, 128, 1);
    return 0;
}

// umma_probe2.cu — tcgen05.mma kind::tf32 probe for the operand forms the fused PPO-update kernel needs beyond umma_probe.cu:
//   * MN-major ("transposed") A and B operands in the no-swizzle canonical layout, read from the SAME bytes a K-major tile
//     occupies (tile X[r][c] stored as [c/4][r][c%4]: K-major with M = r, K = c;  MN-major with MN = c, K = r);
//   * M = 64 accumulators: which TMEM lanes hold which rows;
//   * N = 80 / 144 (multiples of 16 that are not powers of two).
// D[M,N] = A[M,K] * B[N,K]^T; the whole 128-lane x N-column TMEM block is dumped and matched against a CPU reference row by row.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/micro/umma_probe2 scripts/micro/umma_probe2.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// tile X[rows][cols] stored as [cols/4][rows][cols%4] floats (the layout of agx_mlp.cu's canon())
__host__ __device__ inline int tile_off(int r, int c, int rows) { return ((c >> 2) * rows + r) * 4 + (c & 3); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// A logical [M x K], B logical [N x K] (row-major in global memory).
// a_mn = 0: A tile stored with rows = M, cols = K (K-major operand).  a_mn = 1: A^T tile stored with rows = K, cols = M, i.e. the
// bytes of a [K x M] activation tile, consumed as an MN-major operand.  Same for B.
__global__ void __launch_bounds__(128) probe(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int M, int N,
                                              int K, int a_mn, int b_mn) {
    extern __shared__ __align__(128) float smem[];
    float* sA = smem;
    float* sB = smem + M * K;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < M * K; i += 128) {
        const int m = i / K, k = i % K;
        sA[a_mn ? tile_off(k, m, K) : tile_off(m, k, M)] = A[i];
    }
    for (int i = tid; i < N * K; i += 128) {
        const int n = i / K, k = i % K;
        sB[b_mn ? tile_off(k, n, K) : tile_off(n, k, N)] = B[i];
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
    // clear the TMEM block first (tcgen05.st of zeros) so untouched lanes read as a sentinel
    {
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
        for (int c0 = 0; c0 < 256; c0 += 8) {
            const uint32_t z = 0x7fc00000u;  // NaN sentinel
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr + c0), "r"(z) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
                               ((uint32_t)(M >> 4) << 24);
        for (int ks = 0; ks < K / 8; ++ks) {
            uint64_t da, db;
            if (a_mn) {  // [m/4][k][m%4]: MN groups of 4 are K*16 B apart (SBO), 8 k's are 128 B apart (start-address advance / LBO)
                da = make_desc(smem_u32(sA) + ks * 128, 128, (uint32_t)K * 16);
            } else {     // [k/4][m][k%4]: 8-row groups 128 B apart (SBO), the two 16-B K-chunks of an instruction M*16 B apart (LBO)
                da = make_desc(smem_u32(sA) + ks * 2 * M * 16, (uint32_t)M * 16, 128);
            }
            if (b_mn) db = make_desc(smem_u32(sB) + ks * 128, 128, (uint32_t)K * 16);
            else db = make_desc(smem_u32(sB) + ks * 2 * N * 16, (uint32_t)N * 16, 128);
            const uint32_t acc = ks > 0 ? 1u : 0u;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    asm volatile(
        "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(&bar)), "r"(0u)
        : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; ++j) D[tid * N + c0 + j] = __uint_as_float(v[j]);  // D[lane][col]
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

int run(int M, int N, int K, int a_mn, int b_mn) {
    float *hA = (float*)malloc(M * K * 4), *hB = (float*)malloc(N * K * 4), *hD = (float*)malloc(128 * N * 4);
    for (int i = 0; i < M * K; ++i) hA[i] = (float)(((i * 37 + (i / K) * 11) % 17) - 8) / 8.0f;  // exactly representable in tf32, rows distinct
    for (int i = 0; i < N * K; ++i) hB[i] = (float)(((i * 53 + (i / K) * 7) % 13) - 6) / 4.0f;
    float *dA, *dB, *dD;
    cudaMalloc(&dA, M * K * 4); cudaMalloc(&dB, N * K * 4); cudaMalloc(&dD, 128 * N * 4);
    cudaMemcpy(dA, hA, M * K * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, N * K * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, 128 * N * 4);
    const int smem = (M * K + N * K) * 4;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<<<1, 128, smem>>>(dA, dB, dD, M, N, K, a_mn, b_mn);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("M=%d N=%d K=%d a_mn=%d b_mn=%d: CUDA error %s\n", M, N, K, a_mn, b_mn, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(hD, dD, 128 * N * 4, cudaMemcpyDeviceToHost);
    double* ref = (double*)malloc(sizeof(double) * M * N);
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)hA[m * K + k] * hB[n * K + k];
            ref[m * N + n] = s;
        }
    // for every logical row find the TMEM lane that holds it
    int found = 0, identity = 1, map[128];
    for (int m = 0; m < M; ++m) {
        map[m] = -1;
        for (int l = 0; l < 128 && map[m] < 0; ++l) {
            int ok = 1;
            for (int n = 0; n < N && ok; ++n) ok = fabs(ref[m * N + n] - hD[l * N + n]) < 1e-3;
            if (ok) map[m] = l;
        }
        if (map[m] >= 0) ++found;
        if (map[m] != m) identity = 0;
    }
    int touched = 0;
    for (int l = 0; l < 128; ++l) if (hD[l * N] == hD[l * N]) ++touched;  // not the NaN sentinel
    printf("M=%d N=%d K=%d a_mn=%d b_mn=%d: %d / %d rows found, identity lane map: %s, lanes written: %d\n", M, N, K, a_mn, b_mn, found, M,
           identity ? "yes" : "no", touched);
    if (!identity && found == M) {
        printf("   row->lane:");
        for (int m = 0; m < M; ++m) printf(" %d", map[m]);
        printf("\n");
    }
    if (found != M) {
        printf("   D[lane 0][0..7] = ");
        for (int j = 0; j < 8; ++j) printf("%g ", hD[j]);
        printf("| ref[0][0..7] = ");
        for (int j = 0; j < 8; ++j) printf("%g ", ref[j]);
        printf("\n");
    }
    return found != M;
}

int main() {
    run(128, 64, 32, 0, 0);    // baseline (same as umma_probe)
    run(128, 64, 128, 1, 0);   // MN-major A
    run(128, 64, 128, 0, 1);   // MN-major B
    run(128, 64, 128, 1, 1);   // both: the weight-gradient form dW = dZ^T . X with K = 128 batch rows
    run(128, 80, 128, 1, 1);   // N = 80 (64 + the ones block for the bias gradient)
    run(128, 144, 128, 1, 1);  // N = 144
    run(128, 128, 64, 0, 1);   // dH2 = dZ3 . W3 : A K-major (K = 64), B = W3 [64 x 128] tile read MN-major (N = 128)
    run(128, 64, 16, 0, 1);    // dH3 = dout . Wh : K = 16
    run(64, 64, 32, 0, 0);     // M = 64 lane map
    run(64, 144, 128, 1, 1);   // M = 64 weight gradient of layer 3
    run(64, 32, 128, 1, 1);    // M = 64 weight gradient of layer 1
    run(64, 16, 128, 1, 1);    // heads
    return 0;
}

// umma_rate_probe.cu — issue rate of tcgen05.mma (cta_group::1, M = 128) by shape: N = 16..256, kind::tf32 (K = 8 per instruction) and
// kind::f16 with bf16 operands (K = 16), operands K-major in shared memory without swizzle / SWIZZLE_64B / SWIZZLE_128B.  One elected
// lane issues R instructions back to back into the same accumulator (or alternating between two), commits, and the warp waits;
// cycles = clock64 around issue + wait.  Operand VALUES are irrelevant (zeros).  Prints cycles per MMA and MAC/clk/SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/umma_rate_probe scripts/micro/umma_rate_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}
// mode: 0 no swizzle (canonical [kchunk][row][16 B]), 1 SWIZZLE_64B, 2 SWIZZLE_128B ; kind: 0 tf32, 1 bf16
__global__ void __launch_bounds__(128) probe(int N, int mode, int kind, int R, int two_acc, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    if (warp == 0) {
        const uint32_t fmt = kind == 0 ? 2u : 1u;  // a/b format: tf32 = 2, bf16 = 1
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a0 = (s32(smem) + 1023u) & ~1023u, b0 = a0 + 16 * 1024;
        uint64_t da, db;
        if (mode == 0) {
            const uint32_t lboA = 128 * 16, lboB = (uint32_t)N * 16;
            da = (uint64_t)((a0 >> 4) & 0x3FFF) | ((uint64_t)((lboA >> 4) & 0x3FFF) << 16) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
            db = (uint64_t)((b0 >> 4) & 0x3FFF) | ((uint64_t)((lboB >> 4) & 0x3FFF) << 16) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
        } else {
            const uint32_t swb = mode == 1 ? 64u : 128u;
            const uint64_t hi = (uint64_t)(((8u * swb) >> 4) & 0x3FFF) << 32 | ((uint64_t)1 << 46) | ((uint64_t)(mode == 1 ? 4u : 2u) << 61);
            da = hi | ((a0 >> 4) & 0x3FFF) | ((uint64_t)1 << 16);
            db = hi | ((b0 >> 4) & 0x3FFF) | ((uint64_t)1 << 16);
        }
        const long long t0 = clock64();
        if (elect_one()) {
            for (int r = 0; r < R; ++r) {
                const uint32_t d = tmem + ((two_acc && (r & 1)) ? 256u : 0u);
                if (kind == 0)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
        }
        __syncwarp();
        uint32_t done = 0;
        while (!done) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(s32(&bar)), "r"(0u) : "memory");
        const long long t1 = clock64();
        if (tid == 0) out[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}
int main() {
    long long* out;
    cudaMallocManaged(&out, 148 * 8);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int R = 2048;
    const char* mname[3] = {"noswz", "sw64", "sw128"};
    for (int grid : {1, 148})
        for (int kind : {0, 1})
            for (int mode : {0, 1, 2})
                for (int N : {16, 32, 64, 128, 256})
                    for (int two : {0, 1}) {
                        if (grid == 148 && (two || mode == 0)) continue;
                        probe<<<grid, 128, 64 * 1024>>>(N, mode, kind, R, two, out);
                        cudaError_t e = cudaDeviceSynchronize();
                        if (e != cudaSuccess) { printf("grid %d kind %d mode %d N %d: %s\n", grid, kind, mode, N, cudaGetErrorString(e)); return 0; }
                        long long mx = 0;
                        for (int i = 0; i < grid; ++i) mx = out[i] > mx ? out[i] : mx;
                        const double cyc = (double)mx / R, macs = 128.0 * N * (kind == 0 ? 8 : 16);
                        printf("grid %3d %s %-5s N %3d acc %d : %6.1f clk/MMA  %7.0f MAC/clk/SM\n", grid, kind == 0 ? "tf32" : "bf16", mname[mode], N, two + 1, cyc, macs / cyc);
                    }
    return 0;
}

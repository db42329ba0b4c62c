// agx_tc.cuh — tcgen05 / TMEM building blocks shared by the tensor-core MLP kernels (agx_mlp.cu forward, agx_mlp_train.cu backward
// and weight gradients): canonical no-swizzle K-major operand layout, shared-memory descriptors, single-thread MMA issue, commit /
// wait on an mbarrier, TMEM loads.  Descriptor encoding validated against a CPU GEMM by scripts/micro/umma_probe.cu; MN-major
// ("transposed") tf32 operands are NOT available in this layout (scripts/micro/umma_layout_probe.cu: the tensor core reads zeros;
// CUTLASS: "for mn-major tf32 operands, SW128_32B is the only available smem layout"), so every operand here is K-major.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace {
__device__ inline float tc_elu(float x) { return x > 0.0f ? x : __expf(x) - 1.0f; }
__device__ inline float tc_tf32r(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
namespace tc {
constexpr int kM = 128, kH1 = 64, kH2 = 128, kH3 = 64;
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__host__ __device__ inline int canon(int r, int k, int rows) { return ((k >> 2) * rows + r) * 4 + (k & 3); }  // rows % 8 == 0
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
           ((uint64_t)1 << 46);  // version 1 (Blackwell), no swizzle
}
// D[128, N] (+)= A[128, K] * B[N, K]^T, issued by ONE thread
__device__ __forceinline__ void gemm(uint32_t a_base, uint32_t b_base, int N, int K, uint32_t tmem_d, bool accumulate_first = false) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
    const uint32_t lboA = (kM / 8) * 128, lboB = (uint32_t)(N / 8) * 128;
    for (int ks = 0; ks < K / 8; ++ks) {
        const uint64_t da = smem_desc(a_base + ks * 2 * lboA, lboA, 128), db = smem_desc(b_base + ks * 2 * lboB, lboB, 128);
        const uint32_t acc = (ks > 0 || accumulate_first) ? 1u : 0u;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
                     "l"(da), "l"(db), "r"(idesc), "r"(acc)
                     : "memory");
    }
}
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(s32(bar)),
                 "r"(parity)
                 : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// make this thread's generic-proxy shared-memory writes visible to the tensor core, and order its TMEM reads before the barrier;
// the barrier is the 128-thread named barrier of this thread's tile group (id 1 or 2), or the whole CTA (id 0)
__device__ __forceinline__ void publish_and_sync(int bar_id, int nthreads) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nthreads) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
}  // namespace tc
}  // namespace

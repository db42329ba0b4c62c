// agx_mlp_train.cu — backward pass of the actor-critic MLP on the 5th-generation tensor cores (tcgen05 + TMEM), for the shipped
// 64-128-64 network (reference lib/network/mlp.py:4-39 + the mu / value heads, a2c_continuous_logstd_model.py:159-168; the
// gradients torch autograd computes inside calc_gradients, lib/agent/a2c_continuous.py:299-369).
//
// Every intermediate lives in HBM FEATURE-MAJOR, blocked by 128-row tile — [B/128][width][128], written by agx_mlp_forward_train and by
// the kernel below (a tile's planes are one contiguous region: with plain [width][B] planes every 16-row stage of the weight-gradient
// kernel touched 544 different DRAM pages for 64 bytes each and it ran at 2.1 TB/s) —
// because a tf32 tcgen05 operand has to be K-major (agx_tc.cuh) and the weight gradient contracts over the BATCH axis:
//   dW_l[out, in] = sum_b dZ_l[b, out] * a_{l-1}[b, in]   →   A = dZ_l^T (M = out, K = batch rows), B = a_{l-1}^T (N = in, K = batch rows),
// i.e. both operands are rows of those planes, 128 contiguous bytes per feature and 32-row stage.
//
// agx_mlp_backward_tc_kernel — activation-gradient chain, one 128-row tile per 128-thread group (two groups per CTA hide each
//   other's MMA / barrier latency, as in the forward kernel): dout → dH3 = dout·W_head → dZ3 = dH3∘elu'(h3) → dH2 = dZ3·W3 → dZ2 →
//   dH1 = dZ2·W2 → dZ1.  Each product is one accumulation chain of tcgen05.mma (M = 128 rows, N = layer width) with the transposed
//   weights staged K-major in shared memory once per CTA; thread r owns row r = TMEM lane r: tcgen05.ld, multiply by elu'(h) (h read
//   back from its plane, coalesced), store the dZ plane element (coalesced) and the next A operand (16-byte chunks).
// agx_mlp_wgrad_tc_kernel — all four weight gradients AND bias gradients of a batch slab in one persistent CTA per SM: 16-row stages
//   stream through a four-buffer cp.async pipeline straight into the canonical operand layout, one thread issues 8 MMAs per
//   stage (4 layers x 2 K-steps) into four TMEM accumulators that live for the whole slab; the bias gradient is the extra output
//   column produced by a constant ones row appended to each B operand (plane `in_dim` of the normalised input is all ones).
//   Per-CTA partials go through the same deterministic reduction as the mma.sync path (agx_mlp.cu).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "agx.h"
#include "agx_ppo_math.cuh"
#include "agx_tc.cuh"

int agx_internal_fail(int code, const char* msg);
extern "C" int agx_internal_tmap_tiled(void* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                                       int swizzle_bytes);  // agx_conv_tma.cu
int g_wgrad_tma = 1;  // agx_set_option("mlp_wgrad_tma", 0 | 1)
extern "C" int agx_internal_wgrad_reduce(const AgxMlpParams* p, const AgxMlpGrads* g, const float* w_partials, int n_cta_w,
                                         const float* b_partials, int n_cta_b, void* stream);

namespace {

constexpr int kOutPad = 16, kMaxW = 128, kBiasSlots = 3 * kMaxW + kOutPad;
extern __shared__ __align__(128) float t_smem[];

__device__ __forceinline__ float elu_grad_from_out(float h) { return h > 0.0f ? 1.0f : h + 1.0f; }

// ---- activation-gradient chain ------------------------------------------------------------------------------------------------
namespace tcb {
using namespace tc;
// two tile groups per CTA, 256 threads per group: thread (row, half) — the two warps of a TMEM lane quarter split every layer's
// columns (agx_mlp.cu's forward does the same: both kernels are bound by the per-tile epilogue chain, not by the MMAs)
constexpr int kGroup = 2 * kM, kThreads = 2 * kGroup;
constexpr int kW3T = kH2 * kH3, kW2T = kH1 * kH2, kWhT = kH3 * kOutPad;        // floats
constexpr int kGbuf = kM * kH2, kDbuf = kM * kOutPad;                            // per group
constexpr size_t kSmemBytes = sizeof(float) * (size_t)(kW3T + kW2T + kWhT + 2 * (kGbuf + kDbuf));

// dZ = dH ∘ elu'(h) for this thread's row: TMEM cols [col0, col0 + W) x the h plane → dZ plane (+ next A operand when a_next)
template <int W>
__device__ __forceinline__ void grad_epilogue(uint32_t tmem_row, int col0, const float* __restrict__ h_col, float* __restrict__ dz_col,
                                              int64_t B, float* a_next, int r, int half) {
#pragma unroll 1
    for (int c0 = half * (W / 2); c0 < (half + 1) * (W / 2); c0 += 16) {
        float h[16], v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) h[i] = __ldg(h_col + (int64_t)(c0 + i) * B);  // issued before the TMEM load completes
        tmem_ld16(tmem_row + (uint32_t)(col0 + c0), v);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] *= elu_grad_from_out(h[i]);
#pragma unroll
        for (int i = 0; i < 16; ++i) dz_col[(int64_t)(c0 + i) * B] = v[i];
        if (a_next) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(a_next + canon(r, c0 + i, kM)) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
    }
}

constexpr int kLossPartial = 16, kLossGridMax = 296;  // workspace layout of agx_ppo.cu: [grid][16] partials, then the ticket
// LOSS: the PPO loss of this thread's row computed in place of reading grad_mu / grad_value (agx.h AgxLossIO)
template <bool LOSS>
__global__ void __launch_bounds__(kThreads, 1)
agx_mlp_backward_tc_kernel(const __grid_constant__ AgxMlpParams P, int64_t B, const float* __restrict__ grad_mu,
                           const float* __restrict__ grad_value, const float* __restrict__ h1t, const float* __restrict__ h2t,
                           const float* __restrict__ h3t, float* __restrict__ dz1t, float* __restrict__ dz2t, float* __restrict__ dz3t,
                           float* __restrict__ doutt, const __grid_constant__ AgxPpoHyper hp, const __grid_constant__ AgxLossIO lio) {
    float* w3T = t_smem;              // B operand of dH2 = dZ3·W3: [n = 128 (layer-2 feature)][k = 64 (layer-3 feature)] canonical
    float* w2T = w3T + kW3T;          // dH1 = dZ2·W2: [n = 64][k = 128]
    float* whT = w2T + kW2T;          // dH3 = dout·W_head: [n = 64][k = 16]
    float* act = whT + kWhT;
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_base;
    const int tid_all = threadIdx.x, warp_all = tid_all >> 5, group = tid_all / kGroup, tig = tid_all % kGroup, tid = tig & (kM - 1), half = tig >> 7,
              A = P.actions_num;
    float* Gbuf = act + group * (kGbuf + kDbuf);   // dZ3 [128 x 64], then dZ2 [128 x 128]
    float* Dbuf = Gbuf + kGbuf;                    // dout [128 x 16]
    uint64_t* bar = &bars[group];
    // transposed weights, TF32-rounded, K-major canonical.  Global reads are coalesced (torch rows), the strided shared-memory
    // stores are a one-off per CTA.
    // All of a thread's loads are issued before its first store (ncu of the un-batched loops: 45 % of the kernel's stall samples sat on
    // their 16 + 16 + 2 serialised L2 round trips).
    {
        constexpr int k3 = kH3 * kH2 / kThreads, k2 = kH2 * kH1 / kThreads, kh = kOutPad * kH3 / kThreads;
        static_assert(kH3 * kH2 % kThreads == 0 && kH2 * kH1 % kThreads == 0 && kOutPad * kH3 % kThreads == 0, "staging loops assume exact division");
        float v3[k3], v2[k2], vh[kh];
#pragma unroll
        for (int u = 0; u < k3; ++u) v3[u] = __ldg(P.w3 + tid_all + u * kThreads);
#pragma unroll
        for (int u = 0; u < k2; ++u) v2[u] = __ldg(P.w2 + tid_all + u * kThreads);
#pragma unroll
        for (int u = 0; u < kh; ++u) {
            const int i = tid_all + u * kThreads, k = i / kH3, n = i - k * kH3;
            vh[u] = k < A ? __ldg(P.w_mu + k * kH3 + n) : (k == A ? __ldg(P.w_value + n) : 0.0f);
        }
#pragma unroll
        for (int u = 0; u < k3; ++u) { const int i = tid_all + u * kThreads, k = i / kH2, n = i - k * kH2; w3T[canon(n, k, kH2)] = tc_tf32r(v3[u]); }
#pragma unroll
        for (int u = 0; u < k2; ++u) { const int i = tid_all + u * kThreads, k = i / kH1, n = i - k * kH1; w2T[canon(n, k, kH1)] = tc_tf32r(v2[u]); }
#pragma unroll
        for (int u = 0; u < kh; ++u) { const int i = tid_all + u * kThreads, k = i / kH3, n = i - k * kH3; whT[canon(n, k, kH3)] = tc_tf32r(vh[u]); }
    }
    if (tid_all == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp_all == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    publish_and_sync(0, kThreads);
    const uint32_t tmem = tmem_base + (uint32_t)(group * 256);
    const uint32_t tmem_row = tmem + ((uint32_t)((warp_all & 3) * 32) << 16);
    const int gbar = 1 + group;
    uint32_t phase = 0;
    const int64_t n_tiles = B / kM;  // B % 128 == 0 (checked on the host)
    float lacc[kLossPartial];  // LOSS: this thread's sums of a_loss, c_loss, entropy, b_loss, kl, g_logstd[0..a)
#pragma unroll
    for (int i = 0; i < kLossPartial; ++i) lacc[i] = 0.0f;
    float lls[agx::kMaxAct];
    const float inv_b = 1.0f / (float)B;
    if (LOSS) {
#pragma unroll
        for (int i = 0; i < agx::kMaxAct; ++i) lls[i] = i < A ? lio.logstd[i] : 0.0f;
    }
    for (int64_t tile = (int64_t)blockIdx.x * 2 + group; tile < n_tiles; tile += (int64_t)gridDim.x * 2) {
        const int64_t row = tile * kM + tid;
        // planes are blocked by 128-row tile (agx_mlp.cu): element (row, c) of a W-wide tensor at (tile * W + c) * 128 + row % 128
        const float* h3p = h3t + tile * kH3 * kM + tid;
        const float* h2p = h2t + tile * kH2 * kM + tid;
        const float* h1p = h1t + tile * kH1 * kM + tid;
        float* dz3p = dz3t + tile * kH3 * kM + tid;
        float* dz2p = dz2t + tile * kH2 * kM + tid;
        float* dz1p = dz1t + tile * kH1 * kM + tid;
        if (half == 0) {  // dout row = [d loss / d mu (A) | d loss / d value | 0 ...] → its plane and the first A operand (K = 16)
            float d[kOutPad];
#pragma unroll
            for (int c = 0; c < kOutPad; ++c) d[c] = 0.0f;
            if (LOSS) {
                float m[agx::kMaxAct], ac[agx::kMaxAct], om[agx::kMaxAct], os[agx::kMaxAct];
#pragma unroll
                for (int i = 0; i < agx::kMaxAct; ++i) {
                    const bool on = i < A;
                    m[i] = on ? lio.mu[row * A + i] : 0.0f; ac[i] = on ? lio.actions[row * A + i] : 0.0f;
                    om[i] = on ? lio.old_mu[row * A + i] : 0.0f; os[i] = on ? lio.old_sigma[row * A + i] : 1.0f;
                }
                agx::PpoSampleOut o;
                agx::ppo_sample(hp, A, m, lls, lio.value[row], ac, lio.old_neglogp[row], lio.adv[row], lio.returns[row], om, os, o);
#pragma unroll
                for (int i = 0; i < agx::kMaxAct; ++i) {
                    if (i < A) {
                        d[i] = o.g_mu[i] * inv_b;
                        lio.old_mu[row * A + i] = m[i];  // PPODataset.update_mu_sigma
                        lio.old_sigma[row * A + i] = expf(lls[i]);
                        lacc[5 + i] += o.g_logstd[i];
                    }
                }
                if (A == 4) d[4] = o.g_value * inv_b; else d[5] = o.g_value * inv_b;
                lacc[0] += o.a_loss; lacc[1] += o.c_loss; lacc[2] += o.entropy; lacc[3] += o.b_loss; lacc[4] += o.kl;
            } else {
                const float gv = grad_value[row];
#pragma unroll
                for (int c = 0; c < 6; ++c) d[c] = c < A ? grad_mu[row * A + c] : (c == A ? gv : 0.0f);
            }
#pragma unroll
            for (int c = 0; c < kOutPad; ++c) doutt[(tile * kOutPad + c) * kM + tid] = d[c];
#pragma unroll
            for (int c = 0; c < kOutPad; c += 4) *reinterpret_cast<float4*>(Dbuf + canon(tid, c, kM)) = make_float4(d[c], d[c + 1], d[c + 2], d[c + 3]);
        }
        publish_and_sync(gbar, kGroup);
        if (tig == 0) { gemm(s32(Dbuf), s32(whT), kH3, kOutPad, tmem + 0); commit(bar); }
        wait(bar, phase); phase ^= 1;
        grad_epilogue<kH3>(tmem_row, 0, h3p, dz3p, kM, Gbuf, tid, half);
        publish_and_sync(gbar, kGroup);
        if (tig == 0) { gemm(s32(Gbuf), s32(w3T), kH2, kH3, tmem + 64); commit(bar); }
        wait(bar, phase); phase ^= 1;
        grad_epilogue<kH2>(tmem_row, 64, h2p, dz2p, kM, Gbuf, tid, half);  // dZ3 is dead: the MMA that read it has completed
        publish_and_sync(gbar, kGroup);
        if (tig == 0) { gemm(s32(Gbuf), s32(w2T), kH1, kH2, tmem + 192); commit(bar); }
        wait(bar, phase); phase ^= 1;
        grad_epilogue<kH1>(tmem_row, 192, h1p, dz1p, kM, nullptr, tid, half);
        // the next tile's first MMA writes TMEM columns [0, 64): every thread of the group has finished reading them long ago
        // (two barriers back); its Dbuf / Gbuf stores are ordered behind this tile's last MMA by the wait above
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp_all == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    if (LOSS) {  // loss statistics: warp sums → CTA partial → the CTA drawing the last ticket adds the partials in CTA order (deterministic)
        __shared__ float s_part[kThreads / 32][kLossPartial];
        __shared__ bool s_last;
        const int lane = tid_all & 31;
#pragma unroll
        for (int i = 0; i < kLossPartial; ++i) {
            float v = lacc[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) s_part[warp_all][i] = v;
        }
        __syncthreads();
        float* partials = lio.workspace;
        unsigned int* ticket = reinterpret_cast<unsigned int*>(lio.workspace + (int64_t)kLossGridMax * kLossPartial);
        if (tid_all < 5 + A) {
            float v = 0.0f;
            for (int w = 0; w < kThreads / 32; ++w) v += s_part[w][tid_all];
            partials[(int64_t)blockIdx.x * kLossPartial + tid_all] = v;
        }
        __threadfence();
        __syncthreads();
        if (tid_all == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
        __syncthreads();
        if (s_last) {
            __threadfence();
            for (int i = warp_all; i < 5 + A; i += kThreads / 32) {
                float v = 0.0f;
                for (unsigned int c = lane; c < gridDim.x; c += 32) v += __ldcg(partials + (int64_t)c * kLossPartial + i);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) {
                    if (i < 5) lio.stats[i] = v * inv_b;
                    else lio.grad_logstd[i - 5] = v * inv_b - hp.entropy_coef;  // d(-coef * mean entropy)/d logstd_i = -coef
                }
            }
            if (tid_all == 0) *ticket = 0;
        }
    }
}
}  // namespace tcb

// ---- weight + bias gradients --------------------------------------------------------------------------------------------------
namespace tcw {
using namespace tc;
constexpr int kThreads = 256;
constexpr int kStage = 16;            // batch rows (= K) per pipeline stage
constexpr int kBufs = 4;              // stage buffers: three stages of copies in flight under the MMAs of the fourth (with two 32-row
                                      // buffers every stage paid a full DRAM round trip: 34.7 us per 32 768-row minibatch, ncu)
constexpr int kKc = kStage / 4;       // 16-byte K-chunks per operand row and stage
constexpr int kN2 = kH1 + 16, kN3 = kH2 + 16;   // B operands of layers 2 / 3 carry a 16-row block whose first row is all ones (bias column)
// TMEM columns of the four accumulators (M = 128 lanes each; only the first `out` lanes are meaningful)
constexpr int kC1 = 0, kC2 = 96, kC3 = kC2 + kN2, kCh = kC3 + kN3;   // 0 (up to 96 input columns) | 96 | 176 | 320 (+16) <= 512

template <int IN_PAD>
struct Layout {  // float offsets inside one stage buffer; every tile is [kKc][rows][4]
    static constexpr int a1 = 0, a2 = a1 + kKc * kM * 4, a3 = a2 + kKc * kM * 4, ah = a3 + kKc * kM * 4;
    static constexpr int b1 = ah + kKc * kM * 4, b2 = b1 + kKc * IN_PAD * 4, b3 = b2 + kKc * kN2 * 4, bh = b3 + kKc * kN3 * 4;
    static constexpr int floats = bh + kKc * kOutPad * 4;
};
__device__ __forceinline__ void cp16(float* dst, const float* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(dst)), "l"(src) : "memory");
}
// one operand tile of a stage: R feature rows x 32 batch rows from a feature-major plane into [kc][RA][16 B].  A warp covers
// 16 features x 2 K-chunks: 32-byte global sectors fully used, shared-memory stores at most 2-way conflicted.
template <int R, int RA>
__device__ __forceinline__ void load_tile(float* dst, const float* __restrict__ plane, int64_t B, int64_t r0) {
    const float* src = plane + (r0 >> 7) * (R * kM) + (r0 & (kM - 1));  // the 128-row tile's contiguous [R][128] block (R = width of the tensor)
    for (int j = threadIdx.x; j < R * kKc; j += kThreads) {
        const int kc = ((j >> 5) & (kKc / 2 - 1)) * 2 + (j & 1), m = (j / (16 * kKc)) * 16 + ((j >> 1) & 15);
        cp16(dst + (kc * RA + m) * 4, src + m * kM + kc * 4);
    }
}

template <int IN_PAD>
__global__ void __launch_bounds__(kThreads, 1)
agx_mlp_wgrad_tc_kernel(const __grid_constant__ AgxMlpParams P, int64_t B, const float* __restrict__ xt, const float* __restrict__ h1t,
                        const float* __restrict__ h2t, const float* __restrict__ h3t, const float* __restrict__ dz1t,
                        const float* __restrict__ dz2t, const float* __restrict__ dz3t, const float* __restrict__ doutt,
                        float* __restrict__ w_partials, float* __restrict__ b_partials, int partial_floats) {
    using L = Layout<IN_PAD>;
    __shared__ __align__(8) uint64_t bars[kBufs];
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // constant parts of the stage buffers: zero rows that pad M to 128, and the ones rows behind the bias columns
    for (int i = tid; i < kBufs * L::floats / 4; i += kThreads) reinterpret_cast<float4*>(t_smem)[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    __syncthreads();
    for (int i = tid; i < kBufs * kStage; i += kThreads) {
        float* buf = t_smem + (i / kStage) * L::floats;
        const int k = i % kStage;
        buf[L::ah + canon(kH3, k, kM)] = 1.0f;    // A of the heads: row 64 = ones → lane 64 of its accumulator = sum_b dout[b, :]
        buf[L::b2 + canon(kH1, k, kN2)] = 1.0f;   // B of layer 2: row 64 = ones → column 64 = sum_b dZ2[b, :]
        buf[L::b3 + canon(kH2, k, kN3)] = 1.0f;   // B of layer 3: row 128 = ones
    }
    if (tid == 0) {
        for (int b = 0; b < kBufs; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[b])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    publish_and_sync(0, kThreads);
    const uint32_t tmem = tmem_base;
    const int64_t n_stages_total = B / kStage;  // B % 128 == 0
    const int64_t s_begin = (int64_t)blockIdx.x * n_stages_total / gridDim.x, s_end = (int64_t)(blockIdx.x + 1) * n_stages_total / gridDim.x;
    const int S = (int)(s_end - s_begin);

    auto issue = [&](int s) {
        float* buf = t_smem + (s % kBufs) * L::floats;
        const int64_t r0 = (s_begin + s) * kStage;
        load_tile<kH1, kM>(buf + L::a1, dz1t, B, r0);
        load_tile<kH2, kM>(buf + L::a2, dz2t, B, r0);
        load_tile<kH3, kM>(buf + L::a3, dz3t, B, r0);
        load_tile<kH3, kM>(buf + L::ah, h3t, B, r0);
        load_tile<IN_PAD, IN_PAD>(buf + L::b1, xt, B, r0);
        load_tile<kH1, kN2>(buf + L::b2, h1t, B, r0);
        load_tile<kH2, kN3>(buf + L::b3, h2t, B, r0);
        load_tile<kOutPad, kOutPad>(buf + L::bh, doutt, B, r0);
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    for (int s = 0; s < kBufs - 1 && s < S; ++s) issue(s);
    uint32_t phase[kBufs];
#pragma unroll
    for (int b = 0; b < kBufs; ++b) phase[b] = 0u;
    for (int s = 0; s < S; ++s) {
        const int pending = (S - s - 1) < (kBufs - 2) ? (S - s - 1) : (kBufs - 2);  // younger copy groups that may still be in flight
        if (pending >= 2) asm volatile("cp.async.wait_group 2;" ::: "memory");
        else if (pending == 1) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        publish_and_sync(0, kThreads);  // every thread's copies of stage s have landed and are visible to the tensor core
        const int b = s % kBufs;
        if (tid == 0) {
            const uint32_t base = s32(t_smem + b * L::floats);
            const bool acc = s > 0;
            gemm(base + 4 * L::a1, base + 4 * L::b1, IN_PAD, kStage, tmem + kC1, acc);
            gemm(base + 4 * L::a2, base + 4 * L::b2, kN2, kStage, tmem + kC2, acc);
            gemm(base + 4 * L::a3, base + 4 * L::b3, kN3, kStage, tmem + kC3, acc);
            gemm(base + 4 * L::ah, base + 4 * L::bh, kOutPad, kStage, tmem + kCh, acc);
            commit(&bars[b]);
        }
        if (s + kBufs - 1 < S) {  // stage s + kBufs - 1 goes into the buffer of stage s - 1: its MMAs (committed one iteration ago) must be done
            if (s > 0) {
                const int pb = (s - 1) % kBufs;
#pragma unroll
                for (int q = 0; q < kBufs; ++q)
                    if (q == pb) { wait(&bars[q], phase[q]); phase[q] ^= 1u; }
            }
            issue(s + kBufs - 1);
        }
    }
    if (S > 0) {  // the last commit covers every earlier MMA of the issuing thread; earlier unwaited commits only advance their barriers
        // Stages S-4 .. S-1 were never waited for and sit on four different barriers; every earlier commit of barrier lb was, so the
        // barrier is at most one completion behind and the parity of its last commit is (number of its commits - 1) & 1.
        const int lb = (S - 1) % kBufs;
        const int commits = (S - 1 - lb) / kBufs + 1;
        wait(&bars[lb], (uint32_t)((commits - 1) & 1));
    }
    // ---- epilogue: accumulators → this CTA's partials (layout of agx_mlp.cu: dense [out x in_pad] blocks back to back; bias slots b1|b2|b3|heads)
    float* wp = w_partials + (int64_t)blockIdx.x * partial_floats;
    float* bp = b_partials + (int64_t)blockIdx.x * kBiasSlots;
    const int q = warp & 3, half = warp >> 2, m = 32 * q + lane;
    const uint32_t trow = tmem + ((uint32_t)(32 * q) << 16);
    const int in_dim = P.in_dim;
    constexpr int base2 = kH1 * IN_PAD, base3 = base2 + kH2 * kH1, baseh = base3 + kH3 * kH2;
    constexpr int n1 = IN_PAD / 16, n2 = kN2 / 16, n3 = kN3 / 16;
    for (int ci = half; ci < n1 + n2 + n3 + 1; ci += 2) {
        float v[16];
        if (S == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = 0.0f;
        }
        if (ci < n1) {
            const int c0 = ci * 16;
            if (S > 0) tmem_ld16(trow + (uint32_t)(kC1 + c0), v);
            if (m < kH1) {
#pragma unroll
                for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(wp + m * IN_PAD + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
#pragma unroll
                for (int i = 0; i < 16; ++i) if (c0 + i == in_dim) bp[0 * kMaxW + m] = v[i];
            }
        } else if (ci < n1 + n2) {
            const int c0 = (ci - n1) * 16;
            if (S > 0) tmem_ld16(trow + (uint32_t)(kC2 + c0), v);
            if (c0 < kH1) {
#pragma unroll
                for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(wp + base2 + m * kH1 + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            } else {
                bp[1 * kMaxW + m] = v[0];
            }
        } else if (ci < n1 + n2 + n3) {
            const int c0 = (ci - n1 - n2) * 16;
            if (S > 0) tmem_ld16(trow + (uint32_t)(kC3 + c0), v);
            if (m < kH3) {
                if (c0 < kH2) {
#pragma unroll
                    for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(wp + base3 + m * kH2 + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                } else {
                    bp[2 * kMaxW + m] = v[0];
                }
            }
        } else {
            if (S > 0) tmem_ld16(trow + (uint32_t)kCh, v);
            if (m < kH3) {  // accumulator = dW_head^T: lane = layer-3 feature, column = head output
#pragma unroll
                for (int o = 0; o < kOutPad; ++o) wp[baseh + o * kH3 + m] = v[o];
            } else if (m == kH3) {
#pragma unroll
                for (int o = 0; o < kOutPad; ++o) bp[3 * kMaxW + o] = v[o];
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}
}  // namespace tcw

// ---- weight + bias gradients, operands by TMA ---------------------------------------------------------------------------------------
// Same GEMMs, accumulators, ones rows and epilogue as tcw above; what changes is how the operands reach shared memory.  A feature-major
// plane blocked by 128-row tile, [tile][R][128], IS a K-major operand [R rows][K = batch]: a 3-D tensor map (128, R, tiles) with a box of
// (32, R, 1) and SWIZZLE_128B drops the stage's [R][32] slice into the layout tcgen05 reads — ONE cp.async.bulk.tensor per operand and
// stage (8 per 32 batch rows) instead of 71 680 16-byte cp.async per stage pair (the gather kernel was bound by the request rate of those:
// 31 us per 32 768-row minibatch at 2.1-2.4 TB/s).  Warp 0 = producer, warp 1 = MMA issuer (both converged, elect.sync), two stages of
// ~100 KB; all eight warps run the epilogue.
namespace tcw2 {
using namespace tc;
constexpr int kThreads = 256, kStageRows = 32, kStages = 2;
constexpr int kN2 = kH1 + 16, kN3 = kH2 + 16;
constexpr int kC1 = 0, kC2 = 96, kC3 = kC2 + kN2, kCh = kC3 + kN3;  // TMEM columns, as tcw
template <int IN_PAD>
struct Layout {  // byte offsets inside one stage; every operand is [rows][128 B] (SWIZZLE_128B), 1024-byte aligned
    static constexpr uint32_t a1 = 0, a2 = a1 + kM * 128, a3 = a2 + kM * 128, ah = a3 + kM * 128;
    static constexpr uint32_t b1 = ah + kM * 128, b2 = b1 + ((IN_PAD * 128 + 1023) & ~1023), b3 = b2 + ((kN2 * 128 + 1023) & ~1023), bh = b3 + ((kN3 * 128 + 1023) & ~1023);
    static constexpr uint32_t bytes = bh + ((kOutPad * 128 + 1023) & ~1023);
    static constexpr uint32_t tx = (uint32_t)(kH1 + kH2 + kH3 + kH3 + IN_PAD + kH1 + kH2 + kOutPad) * 128u;  // bytes the eight boxes of a stage bring
};
struct Maps { CUtensorMap dz1, dz2, dz3, h3, x, h1, h2, dout; };

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_wait_plain(uint64_t* b, uint32_t parity) {
    uint32_t done = 0;
    while (!done) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(s32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma3(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(tm)), "r"(s32(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
// D[128, N] (+)= A[128, 32] * B[N, 32]^T, both operands [rows][128 B] SWIZZLE_128B: 4 K steps of 8 (start address + 32 bytes each)
__device__ __forceinline__ void gemm_sw128(uint32_t a, uint32_t b, int N, uint32_t d, bool acc) {
    constexpr uint32_t kHi = ((1024u >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
    const uint32_t al = (a >> 4) | 0x10000u, bl = (b >> 4) | 0x10000u;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        const uint32_t on = (ks > 0 || acc) ? 1u : 0u;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
                     "l"(((uint64_t)kHi << 32) | (al + 2u * ks)), "l"(((uint64_t)kHi << 32) | (bl + 2u * ks)), "r"(idesc), "r"(on)
                     : "memory");
    }
}

extern __shared__ uint8_t w_smem[];

template <int IN_PAD>
__global__ void __launch_bounds__(kThreads, 1)
agx_mlp_wgrad_tma_kernel(const __grid_constant__ Maps M, const __grid_constant__ AgxMlpParams P, int64_t B, float* __restrict__ w_partials,
                         float* __restrict__ b_partials, int partial_floats) {
    using L = Layout<IN_PAD>;
    __shared__ __align__(8) uint64_t full[kStages], empty[kStages], done;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t smem0 = (s32(w_smem) + 1023u) & ~1023u;
    uint8_t* gen0 = w_smem + (smem0 - s32(w_smem));
    // constant parts of the stage buffers: zero rows that pad the operands, and the ones rows behind the bias columns
    for (uint32_t i = tid; i < kStages * L::bytes / 16; i += kThreads) reinterpret_cast<float4*>(gen0)[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    __syncthreads();
    for (int i = tid; i < kStages * 3 * 32; i += kThreads) {
        const int st = i / 96, which = (i / 32) % 3, k = i & 31;
        const uint32_t off = which == 0 ? L::ah + kH3 * 128 : (which == 1 ? L::b2 + kH1 * 128 : L::b3 + kH2 * 128);  // row 64 of A_heads / row 64 of B2 / row 128 of B3
        reinterpret_cast<float*>(gen0 + st * L::bytes + off)[k] = 1.0f;  // a row of 32 ones reads the same under any swizzle
    }
    if (tid == 0) {
        for (int b = 0; b < kStages; ++b) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[b])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&empty[b])));
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&done)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    publish_and_sync(0, kThreads);
    const uint32_t tmem = tmem_base;
    const int64_t n_stages_total = B / kStageRows;  // B % 128 == 0
    const int64_t s_begin = (int64_t)blockIdx.x * n_stages_total / gridDim.x, s_end = (int64_t)(blockIdx.x + 1) * n_stages_total / gridDim.x;
    const int S = (int)(s_end - s_begin);
    if (warp == 0) {  // ---- producer
        for (int s = 0; s < S; ++s) {
            const int b = s % kStages;
            mbar_wait_plain(&empty[b], (uint32_t)(((s / kStages) & 1) ^ 1));
            if (elect_one()) {
                const int64_t r0 = (s_begin + s) * kStageRows;
                const int tile = (int)(r0 >> 7), k0 = (int)(r0 & (kM - 1));
                const uint32_t base = smem0 + (uint32_t)b * L::bytes;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[b])), "r"(L::tx) : "memory");
                tma3(base + L::a1, &M.dz1, &full[b], k0, 0, tile);
                tma3(base + L::a2, &M.dz2, &full[b], k0, 0, tile);
                tma3(base + L::a3, &M.dz3, &full[b], k0, 0, tile);
                tma3(base + L::ah, &M.h3, &full[b], k0, 0, tile);
                tma3(base + L::b1, &M.x, &full[b], k0, 0, tile);
                tma3(base + L::b2, &M.h1, &full[b], k0, 0, tile);
                tma3(base + L::b3, &M.h2, &full[b], k0, 0, tile);
                tma3(base + L::bh, &M.dout, &full[b], k0, 0, tile);
            }
            __syncwarp();
        }
    } else if (warp == 1) {  // ---- MMA issuer
        for (int s = 0; s < S; ++s) {
            const int b = s % kStages;
            mbar_wait_plain(&full[b], (uint32_t)((s / kStages) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                const uint32_t base = smem0 + (uint32_t)b * L::bytes;
                const bool acc = s > 0;
                gemm_sw128(base + L::a1, base + L::b1, IN_PAD, tmem + kC1, acc);
                gemm_sw128(base + L::a2, base + L::b2, kN2, tmem + kC2, acc);
                gemm_sw128(base + L::a3, base + L::b3, kN3, tmem + kC3, acc);
                gemm_sw128(base + L::ah, base + L::bh, kOutPad, tmem + kCh, acc);
                commit(&empty[b]);
                if (s + 1 == S) commit(&done);
            }
            __syncwarp();
        }
    }
    __syncwarp();
    if (S > 0) { mbar_wait_plain(&done, 0u); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
    // ---- epilogue: accumulators → this CTA's partials (layout of agx_mlp.cu: dense [out x in_pad] blocks back to back; bias slots b1|b2|b3|heads)
    float* wp = w_partials + (int64_t)blockIdx.x * partial_floats;
    float* bp = b_partials + (int64_t)blockIdx.x * kBiasSlots;
    const int q = warp & 3, half = warp >> 2, m = 32 * q + lane;
    const uint32_t trow = tmem + ((uint32_t)(32 * q) << 16);
    const int in_dim = P.in_dim;
    constexpr int base2 = kH1 * IN_PAD, base3 = base2 + kH2 * kH1, baseh = base3 + kH3 * kH2;
    constexpr int n1 = IN_PAD / 16, n2 = kN2 / 16, n3 = kN3 / 16;
    for (int ci = half; ci < n1 + n2 + n3 + 1; ci += 2) {
        float v[16];
        if (S == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = 0.0f;
        }
        if (ci < n1) {
            const int c0 = ci * 16;
            if (S > 0) tmem_ld16(trow + (uint32_t)(kC1 + c0), v);
            if (m < kH1) {
#pragma unroll
                for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(wp + m * IN_PAD + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
#pragma unroll
                for (int i = 0; i < 16; ++i) if (c0 + i == in_dim) bp[0 * kMaxW + m] = v[i];
            }
        } else if (ci < n1 + n2) {
            const int c0 = (ci - n1) * 16;
            if (S > 0) tmem_ld16(trow + (uint32_t)(kC2 + c0), v);
            if (c0 < kH1) {
#pragma unroll
                for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(wp + base2 + m * kH1 + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            } else {
                bp[1 * kMaxW + m] = v[0];
            }
        } else if (ci < n1 + n2 + n3) {
            const int c0 = (ci - n1 - n2) * 16;
            if (S > 0) tmem_ld16(trow + (uint32_t)(kC3 + c0), v);
            if (m < kH3) {
                if (c0 < kH2) {
#pragma unroll
                    for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(wp + base3 + m * kH2 + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                } else {
                    bp[2 * kMaxW + m] = v[0];
                }
            }
        } else {
            if (S > 0) tmem_ld16(trow + (uint32_t)kCh, v);
            if (m < kH3) {  // accumulator = dW_head^T: lane = layer-3 feature, column = head output
#pragma unroll
                for (int o = 0; o < kOutPad; ++o) wp[baseh + o * kH3 + m] = v[o];
            } else if (m == kH3) {
#pragma unroll
                for (int o = 0; o < kOutPad; ++o) bp[3 * kMaxW + o] = v[o];
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}
}  // namespace tcw2

bool train_ok(const AgxMlpParams* p) {
    return p && p->h1 == tc::kH1 && p->h2 == tc::kH2 && p->h3 == tc::kH3 && (p->in_pad == 32 || p->in_pad == 48 || p->in_pad == 64 || p->in_pad == 96) &&
           p->in_dim > 0 && p->in_dim < p->in_pad && (p->actions_num == 4 || p->actions_num == 5) && p->w1 && p->w2 && p->w3 && p->w_mu && p->w_value;
}

}  // namespace

extern "C" {

int agx_mlp_train_supported(const AgxMlpParams* p) { return train_ok(p) ? 1 : 0; }

static int backward_train_impl(const AgxPpoHyper* hp, const AgxLossIO* lio, const AgxMlpParams* p, const AgxMlpGrads* g, int64_t b, const float* grad_mu,
                               const float* grad_value, const float* xt, const float* h1t, const float* h2t, const float* h3t, float* dz1t, float* dz2t,
                               float* dz3t, float* doutt, float* workspace, void* stream) {
    if (!train_ok(p) || !g || b <= 0 || (b % tc::kM) != 0 || (!lio && (!grad_mu || !grad_value)) || !xt || !h1t || !h2t || !h3t || !dz1t || !dz2t || !dz3t ||
        !doutt || !workspace || !g->gw1 || !g->gb1 || !g->gw2 || !g->gb2 || !g->gw3 || !g->gb3 || !g->gw_mu || !g->gb_mu || !g->gw_value || !g->gb_value)
        return agx_internal_fail(AGX_ERR_ARG, "agx_mlp_backward_train: bad argument (64-128-64 network, in_pad in {32,48,64,96} > in_dim, batch % 128 == 0)");
    if (lio && (!hp || !lio->mu || !lio->logstd || !lio->value || !lio->actions || !lio->old_neglogp || !lio->adv || !lio->returns || !lio->old_mu ||
                !lio->old_sigma || !lio->grad_logstd || !lio->stats || !lio->workspace || lio->a != p->actions_num))
        return agx_internal_fail(AGX_ERR_ARG, "agx_ppo_loss_backward_train: bad loss block (a must equal the network's actions_num)");
    const uintptr_t al = (uintptr_t)xt | (uintptr_t)h1t | (uintptr_t)h2t | (uintptr_t)h3t | (uintptr_t)dz1t | (uintptr_t)dz2t | (uintptr_t)dz3t |
                         (uintptr_t)doutt | (uintptr_t)workspace;
    if (al & 15u) return agx_internal_fail(AGX_ERR_ALIGN, "agx_mlp_backward_train: buffers must be 16-byte aligned");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    constexpr int kGrid = 148;
    const int64_t tiles = b / tc::kM, pairs = (tiles + 1) / 2;
    AgxPpoHyper hz;
    AgxLossIO lz;
    memset(&hz, 0, sizeof(hz));
    memset(&lz, 0, sizeof(lz));
    if (lio) {
        cudaFuncSetAttribute(tcb::agx_mlp_backward_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tcb::kSmemBytes);
        tcb::agx_mlp_backward_tc_kernel<true><<<(unsigned)(pairs < kGrid ? pairs : kGrid), tcb::kThreads, tcb::kSmemBytes, st>>>(
            *p, b, nullptr, nullptr, h1t, h2t, h3t, dz1t, dz2t, dz3t, doutt, *hp, *lio);
    } else {
        cudaFuncSetAttribute(tcb::agx_mlp_backward_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tcb::kSmemBytes);
        tcb::agx_mlp_backward_tc_kernel<false><<<(unsigned)(pairs < kGrid ? pairs : kGrid), tcb::kThreads, tcb::kSmemBytes, st>>>(
            *p, b, grad_mu, grad_value, h1t, h2t, h3t, dz1t, dz2t, dz3t, doutt, hz, lz);
    }
    const int pf = p->h1 * p->in_pad + p->h2 * p->h1 + p->h3 * p->h2 + kOutPad * p->h3;
    const int64_t stages = b / tcw::kStage;
    const unsigned gw = (unsigned)(stages < kGrid ? stages : kGrid);
    float* w_partials = workspace;
    float* b_partials = workspace + (int64_t)kGrid * pf;
#define AGX_WGRAD_TC(PAD)                                                                                                         \
    do {                                                                                                                          \
        constexpr int kSm = tcw::kBufs * tcw::Layout<PAD>::floats * (int)sizeof(float);                                                    \
        cudaFuncSetAttribute(tcw::agx_mlp_wgrad_tc_kernel<PAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSm);                \
        tcw::agx_mlp_wgrad_tc_kernel<PAD><<<gw, tcw::kThreads, kSm, st>>>(*p, b, xt, h1t, h2t, h3t, dz1t, dz2t, dz3t, doutt, w_partials, \
                                                                          b_partials, pf);                                        \
    } while (0)
    // operands by TMA (tcw2) when the tensor maps can be built; the cp.async gather kernel (tcw) otherwise or on agx_set_option("mlp_wgrad_tma", 0)
    tcw2::Maps maps;
    bool tma = g_wgrad_tma != 0;
    if (tma) {
        const uint64_t ntile = (uint64_t)tiles;
        auto mk = [&](CUtensorMap* m, const float* plane, int R) {
            const uint64_t dims[3] = {(uint64_t)tc::kM, (uint64_t)R, ntile}, strides[2] = {(uint64_t)tc::kM * 4, (uint64_t)R * tc::kM * 4};
            const uint32_t box[3] = {(uint32_t)tcw2::kStageRows, (uint32_t)R, 1};
            return agx_internal_tmap_tiled(m, plane, 3, dims, strides, box, 128) == 1;
        };
        tma = mk(&maps.dz1, dz1t, tc::kH1) && mk(&maps.dz2, dz2t, tc::kH2) && mk(&maps.dz3, dz3t, tc::kH3) && mk(&maps.h3, h3t, tc::kH3) &&
              mk(&maps.x, xt, p->in_pad) && mk(&maps.h1, h1t, tc::kH1) && mk(&maps.h2, h2t, tc::kH2) && mk(&maps.dout, doutt, kOutPad);
    }
    unsigned gused = gw;
    if (tma) {
        const int64_t st2 = b / tcw2::kStageRows;
        gused = (unsigned)(st2 < kGrid ? st2 : kGrid);
#define AGX_WGRAD_TMA(PAD)                                                                                                        \
    do {                                                                                                                          \
        constexpr int kSm = tcw2::kStages * (int)tcw2::Layout<PAD>::bytes + 1024;                                                  \
        cudaFuncSetAttribute(tcw2::agx_mlp_wgrad_tma_kernel<PAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSm);               \
        tcw2::agx_mlp_wgrad_tma_kernel<PAD><<<gused, tcw2::kThreads, kSm, st>>>(maps, *p, b, w_partials, b_partials, pf);          \
    } while (0)
        if (p->in_pad == 32) AGX_WGRAD_TMA(32); else if (p->in_pad == 48) AGX_WGRAD_TMA(48); else if (p->in_pad == 64) AGX_WGRAD_TMA(64); else AGX_WGRAD_TMA(96);
#undef AGX_WGRAD_TMA
    } else {
        if (p->in_pad == 32) AGX_WGRAD_TC(32); else if (p->in_pad == 48) AGX_WGRAD_TC(48); else if (p->in_pad == 64) AGX_WGRAD_TC(64); else AGX_WGRAD_TC(96);
    }
#undef AGX_WGRAD_TC
    if (cudaGetLastError() != cudaSuccess) return agx_internal_fail(AGX_ERR_CUDA, "agx_mlp_backward_train: launch failed");
    return agx_internal_wgrad_reduce(p, g, w_partials, (int)gused, b_partials, (int)gused, stream);
}

int agx_mlp_backward_train(const AgxMlpParams* p, const AgxMlpGrads* g, int64_t b, const float* grad_mu, const float* grad_value,
                           const float* xt, const float* h1t, const float* h2t, const float* h3t, float* dz1t, float* dz2t, float* dz3t,
                           float* doutt, float* workspace, void* stream) {
    return backward_train_impl(nullptr, nullptr, p, g, b, grad_mu, grad_value, xt, h1t, h2t, h3t, dz1t, dz2t, dz3t, doutt, workspace, stream);
}

int agx_sizeof_loss_io(void) { return (int)sizeof(AgxLossIO); }

int agx_ppo_loss_backward_train(const AgxPpoHyper* hp, const AgxLossIO* lio, const AgxMlpParams* p, const AgxMlpGrads* g, int64_t b,
                                const float* xt, const float* h1t, const float* h2t, const float* h3t, float* dz1t, float* dz2t, float* dz3t,
                                float* doutt, float* workspace, void* stream) {
    if (!lio || !hp) return agx_internal_fail(AGX_ERR_ARG, "agx_ppo_loss_backward_train: null loss block");
    return backward_train_impl(hp, lio, p, g, b, nullptr, nullptr, xt, h1t, h2t, h3t, dz1t, dz2t, dz3t, doutt, workspace, stream);
}

}  // extern "C"

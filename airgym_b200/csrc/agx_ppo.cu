// agx_ppo.cu — PPO update kernels for sm_100a behind the C ABI (include/agx.h, row a13).
// All three are bandwidth-/latency-trivial next to the env step; the point of fusing them is to remove the hundreds
// of tiny torch launches and every host sync (.item()) from the update loop so that it can live in one CUDA graph.
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "agx.h"
#include "agx_comm.cuh"
#include "agx_ppo_math.cuh"

int agx_internal_fail(int code, const char* msg);  // agx_step.cu: sets the thread-local text behind agx_error_string()

namespace {

int fail_ppo(int code, const char* msg) { return agx_internal_fail(code, msg); }

constexpr int kGaeBlock = 128;
constexpr int kLossBlock = 256;
constexpr int kLossGridMax = 296;   // 2 CTAs per SM on 148 SMs
constexpr int kPartial = 16;        // floats per CTA partial: 5 stats + 5 logstd grads (+pad)

// ---- GAE: one thread per env, rows staged through shared memory so global traffic stays coalesced -----------
__global__ void __launch_bounds__(kGaeBlock)
agx_gae_kernel(int64_t n, int h, float gamma, float tau, const float* __restrict__ rewards,
               const float* __restrict__ values, const uint8_t* __restrict__ dones,
               const float* __restrict__ last_values, const uint8_t* __restrict__ last_dones,
               float* __restrict__ adv, float* __restrict__ ret) {
    extern __shared__ float smem[];
    float* s_r = smem;                          // [kGaeBlock, h+1] padded rows (odd stride when h is even)
    const int stride = h | 1;
    float* s_v = s_r + kGaeBlock * stride;
    float* s_a = s_v + kGaeBlock * stride;
    float* s_t = s_a + kGaeBlock * stride;
    uint8_t* s_d = reinterpret_cast<uint8_t*>(s_t + kGaeBlock * stride);  // [kGaeBlock, h]
    const int64_t tile0 = (int64_t)blockIdx.x * kGaeBlock;
    const int tile_n = (int)((n - tile0) < kGaeBlock ? (n - tile0) : kGaeBlock);
    const int tid = threadIdx.x;
    for (int i = tid; i < tile_n * h; i += kGaeBlock) {
        const int r = i / h, c = i - r * h;
        s_r[r * stride + c] = rewards[tile0 * h + i];
        s_v[r * stride + c] = values[tile0 * h + i];
        s_d[r * h + c] = dones[tile0 * h + i];
    }
    __syncthreads();
    if (tid < tile_n) {
        agx::gae_row(h, gamma, tau, s_r + tid * stride, s_v + tid * stride, s_d + tid * h, last_values[tile0 + tid],
                     (float)last_dones[tile0 + tid], s_a + tid * stride, s_t + tid * stride);
    }
    __syncthreads();
    for (int i = tid; i < tile_n * h; i += kGaeBlock) {
        const int r = i / h, c = i - r * h;
        adv[tile0 * h + i] = s_a[r * stride + c];
        ret[tile0 * h + i] = s_t[r * stride + c];
    }
}

// ---- fused PPO loss forward + backward ---------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int A>
__global__ void __launch_bounds__(kLossBlock)
agx_ppo_loss_kernel(const __grid_constant__ AgxPpoHyper hp, int64_t b, const float* __restrict__ mu,
                    const float* __restrict__ logstd, const float* __restrict__ value,
                    const float* __restrict__ actions, const float* __restrict__ old_neglogp,
                    const float* __restrict__ adv, const float* __restrict__ returns, float* old_mu,
                    float* old_sigma, float* __restrict__ grad_mu, float* __restrict__ grad_value,
                    float* __restrict__ grad_logstd, float* __restrict__ stats, float* workspace) {
    __shared__ float s_part[kLossBlock / 32][kPartial];
    __shared__ bool s_last;
    float ls[agx::kMaxAct], sig[agx::kMaxAct];
#pragma unroll
    for (int i = 0; i < agx::kMaxAct; ++i) { ls[i] = i < A ? logstd[i] : 0.0f; sig[i] = i < A ? expf(ls[i]) : 0.0f; }
    float acc[kPartial];
#pragma unroll
    for (int i = 0; i < kPartial; ++i) acc[i] = 0.0f;
    const float inv_b = 1.0f / (float)b;
    const bool vec4 = ((reinterpret_cast<uintptr_t>(mu) | reinterpret_cast<uintptr_t>(actions) | reinterpret_cast<uintptr_t>(old_mu) |
                        reinterpret_cast<uintptr_t>(old_sigma) | reinterpret_cast<uintptr_t>(grad_mu)) & 15) == 0;
    for (int64_t s = (int64_t)blockIdx.x * kLossBlock + threadIdx.x; s < b; s += (int64_t)gridDim.x * kLossBlock) {
        float m[agx::kMaxAct], ac[agx::kMaxAct], om[agx::kMaxAct], os[agx::kMaxAct];
#pragma unroll
        for (int i = A; i < agx::kMaxAct; ++i) { m[i] = 0; ac[i] = 0; om[i] = 0; os[i] = 1; }
        if (A == 4 && vec4) {  // one 16-B access per row and array
            const float4 a = reinterpret_cast<const float4*>(mu)[s], c = reinterpret_cast<const float4*>(actions)[s];
            const float4 d = reinterpret_cast<const float4*>(old_mu)[s], f = reinterpret_cast<const float4*>(old_sigma)[s];
            m[0] = a.x; m[1] = a.y; m[2] = a.z; m[3] = a.w; ac[0] = c.x; ac[1] = c.y; ac[2] = c.z; ac[3] = c.w;
            om[0] = d.x; om[1] = d.y; om[2] = d.z; om[3] = d.w; os[0] = f.x; os[1] = f.y; os[2] = f.z; os[3] = f.w;
        } else {
#pragma unroll
            for (int i = 0; i < A; ++i) { m[i] = mu[s * A + i]; ac[i] = actions[s * A + i]; om[i] = old_mu[s * A + i]; os[i] = old_sigma[s * A + i]; }
        }
        agx::PpoSampleOut o;
        agx::ppo_sample(hp, A, m, ls, value[s], ac, old_neglogp[s], adv[s], returns[s], om, os, o);
#pragma unroll
        for (int i = 0; i < A; ++i) acc[5 + i] += o.g_logstd[i];
        if (A == 4 && vec4) {
            reinterpret_cast<float4*>(grad_mu)[s] = make_float4(o.g_mu[0] * inv_b, o.g_mu[1] * inv_b, o.g_mu[2] * inv_b, o.g_mu[3] * inv_b);
            reinterpret_cast<float4*>(old_mu)[s] = make_float4(m[0], m[1], m[2], m[3]);  // PPODataset.update_mu_sigma
            reinterpret_cast<float4*>(old_sigma)[s] = make_float4(sig[0], sig[1], sig[2], sig[3]);
        } else {
#pragma unroll
            for (int i = 0; i < A; ++i) {
                grad_mu[s * A + i] = o.g_mu[i] * inv_b;
                old_mu[s * A + i] = m[i];          // PPODataset.update_mu_sigma
                old_sigma[s * A + i] = sig[i];
            }
        }
        grad_value[s] = o.g_value * inv_b;
        acc[0] += o.a_loss; acc[1] += o.c_loss; acc[2] += o.entropy; acc[3] += o.b_loss; acc[4] += o.kl;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 5 + A; ++i) {
        const float v = warp_sum(acc[i]);
        if (lane == 0) s_part[warp][i] = v;
    }
    __syncthreads();
    float* partials = workspace;                                   // [gridDim.x, kPartial]
    unsigned int* ticket = reinterpret_cast<unsigned int*>(workspace + (int64_t)kLossGridMax * kPartial);
    if (threadIdx.x < 5 + A) {
        float v = 0.0f;
        for (int w = 0; w < kLossBlock / 32; ++w) v += s_part[w][threadIdx.x];
        partials[(int64_t)blockIdx.x * kPartial + threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (s_last) {  // fixed-order final reduction → bitwise reproducible
        __threadfence();
        for (int i = warp; i < 5 + A; i += kLossBlock / 32) {  // one warp per statistic: strided partial sums, then a fixed xor tree
            float v = 0.0f;
            for (unsigned int c = lane; c < gridDim.x; c += 32) v += __ldcg(partials + (int64_t)c * kPartial + i);
            v = warp_sum(v);
            if (lane == 0) {
                if (i < 5) stats[i] = v * inv_b;
                else grad_logstd[i - 5] = v * inv_b - hp.entropy_coef;  // d(-coef * mean entropy)/d logstd_i = -coef
            }
        }
        if (threadIdx.x == 0) *ticket = 0;
    }
}

// ---- fused grad-scale + clip_grad_norm_ + Adam + adaptive LR (one thread-block cluster) ---------------------------
// One SM's load/store path bounds a single-CTA version (4 arrays in, 3 out through one SM: 17 us for 47 k parameters), so the
// update runs as ONE cluster of 8 CTAs: each CTA reduces the squared norm of its interleaved slice, the 8 partial sums are
// exchanged through distributed shared memory and added in rank order by every thread (deterministic), then each CTA updates
// its slice.  lr / step are read before the closing cluster barrier and written after it by rank 0.
constexpr int kAdamBlock = agxc::kBlock;
constexpr int kAdamCluster = agxc::kCluster;
// COMM = true: the multi-GPU variant — the kernel first all-reduces `g` [n + n_extra] (flat gradients ‖ loss statistics incl. the KL,
// reference a2c_base.py:293-309 + a2c_continuous.py:112-123) across the ranks through NVLink peer memory (agx_comm.cuh: push into
// every peer's slot, flag, wait, add the W slots in rank order), writes the sums back into `g` and carries straight on with
// the norm / clip / Adam / learning-rate rule on them: ONE launch per minibatch instead of NCCL all-reduce + Adam, and bitwise
// identical parameters on every rank.
template <bool COMM>
__global__ void __cluster_dims__(kAdamCluster, 1, 1) __launch_bounds__(kAdamBlock)
agx_adam_kernel(const __grid_constant__ AgxPpoHyper hp, const __grid_constant__ AgxComm comm, int64_t n, int64_t n_extra,
                float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, float* lr_dev,
                long long* step_dev, const float* kl_dev, float grad_scale, float* norm_out) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ float s_red[kAdamBlock / 32];
    __shared__ float s_part;
    const int64_t first = (int64_t)cluster.block_rank() * kAdamBlock + threadIdx.x, stride = (int64_t)kAdamCluster * kAdamBlock;
    float ss = 0.0f;
    unsigned long long seq = 0;
    if (COMM) {
        seq = agxc::push<float>(comm, g, n + n_extra);
        // (element i of g is pushed and later overwritten with the sum by the same thread: no barrier needed in between)
        unsigned long long t0 = 0ull;
        for (int64_t i = first; i < n + n_extra; i += stride) {
            const float sum = agxc::reduce_elem<float>(comm, seq, i, t0);
            g[i] = sum;  // re-read below by this same thread
            if (i < n) { const float x = sum * grad_scale; ss += x * x; }
        }
    } else {
        for (int64_t i = first; i < n; i += stride) { const float x = g[i] * grad_scale; ss += x * x; }
    }
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = threadIdx.x < kAdamBlock / 32 ? s_red[threadIdx.x] : 0.0f;
        t = warp_sum(t);
        if (threadIdx.x == 0) s_part = t;
    }
    cluster.sync();
    float total = 0.0f;
#pragma unroll
    for (int r = 0; r < kAdamCluster; ++r) total += *cluster.map_shared_rank(&s_part, r);
    const float norm = sqrtf(total);
    float clip = 1.0f;
    if (hp.grad_norm > 0.0f) { clip = hp.grad_norm / (norm + 1e-6f); clip = clip > 1.0f ? 1.0f : clip; }  // clip_grad_norm_
    const long long t = step_dev[0] + 1;
    const float lr = lr_dev[0];
    // the KL rides in the extra slots of `g`: with COMM it was written by OTHER threads of this cluster above — the cluster
    // barrier ordered those stores; read it through L2
    const float kl = (hp.adaptive_lr && kl_dev) ? __ldcg(kl_dev) : 0.0f;
    const float bc1 = 1.0f - powf(hp.beta1, (float)t), bc2 = 1.0f - powf(hp.beta2, (float)t);
    const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
#pragma unroll 2
    for (int64_t i = first; i < n; i += stride) {
        float gi = g[i] * grad_scale * clip;
        const float pi = p[i];
        if (hp.weight_decay != 0.0f) gi += hp.weight_decay * pi;
        const float mi = hp.beta1 * m[i] + (1.0f - hp.beta1) * gi;
        const float vi = hp.beta2 * v[i] + (1.0f - hp.beta2) * gi * gi;
        m[i] = mi; v[i] = vi;
        p[i] = pi - step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + hp.eps);
    }
    cluster.sync();  // every CTA has read lr / step (and the call counter) and every remote s_part read has completed
    if (cluster.block_rank() == 0 && threadIdx.x == 0) {
        step_dev[0] = t;
        if (norm_out) *norm_out = norm;
        if (hp.adaptive_lr && kl_dev) lr_dev[0] = agx::adaptive_lr(lr, kl * grad_scale, hp.kl_threshold);
        if (COMM) *static_cast<volatile unsigned long long*>(comm.region[comm.rank]) = seq + 1ull;
    }
}

// ---- RunningMeanStd update in two launches (lib/core/running_mean_std.py:45-60) ------------------------------------------------------
// (1) float64 column sums / sums of squares of x [n, k]: per-CTA partials, the CTA drawing the last ticket adds them in CTA order
// (deterministic); (2) after an optional all-reduce of those sums: batch mean, UNBIASED batch variance, parallel-variance merge.
constexpr int kSumsBlock = 256, kSumsGridMax = 296, kSumsMaxK = 128;
__global__ void __launch_bounds__(kSumsBlock)
agx_col_sums_kernel(const float* __restrict__ x, int64_t n, int k, int64_t ld, int cpr, double* __restrict__ sums, double* __restrict__ ws) {
    __shared__ double s_a[kSumsBlock], s_b[kSumsBlock];
    __shared__ bool s_last;
    const int c = threadIdx.x % cpr, rg = threadIdx.x / cpr, rgs = kSumsBlock / cpr;
    double a = 0.0, b = 0.0;
    if (c < k) {  // four rows' loads in flight per thread, added in row order (the sums are bit-identical to the one-load-per-iteration loop, whose
                  // ~14 serialised L2 round trips were most of this kernel's 17 us)
        const int64_t stride = (int64_t)gridDim.x * rgs;
        for (int64_t r = (int64_t)blockIdx.x * rgs + rg; r < n; r += 4 * stride) {
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = (r + u * stride < n) ? __ldg(x + (r + u * stride) * ld + c) : 0.0f;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (r + u * stride < n) { const double w = (double)v[u]; a += w; b += w * w; }
        }
    }
    s_a[threadIdx.x] = a; s_b[threadIdx.x] = b;
    __syncthreads();
    double* part = ws + 8 + (int64_t)blockIdx.x * 2 * kSumsMaxK;
    if (threadIdx.x < cpr && threadIdx.x < k) {
        double ta = 0.0, tb = 0.0;
        for (int g = 0; g < rgs; ++g) { ta += s_a[g * cpr + threadIdx.x]; tb += s_b[g * cpr + threadIdx.x]; }
        part[threadIdx.x] = ta; part[kSumsMaxK + threadIdx.x] = tb;
    }
    __threadfence();
    __syncthreads();
    unsigned int* ticket = reinterpret_cast<unsigned int*>(ws);
    if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (s_last) {  // partials of all CTAs: row group rg adds CTAs rg, rg + rgs, ... (independent loads), then the groups in order → deterministic
        __threadfence();
        double ta = 0.0, tb = 0.0;
        if (c < k)
            for (unsigned int g = rg; g < gridDim.x; g += rgs) { ta += __ldcg(ws + 8 + (int64_t)g * 2 * kSumsMaxK + c); tb += __ldcg(ws + 8 + (int64_t)g * 2 * kSumsMaxK + kSumsMaxK + c); }
        __syncthreads();
        s_a[threadIdx.x] = ta; s_b[threadIdx.x] = tb;
        __syncthreads();
        if (threadIdx.x < cpr && threadIdx.x < k) {
            double fa = 0.0, fb = 0.0;
            for (int g = 0; g < rgs; ++g) { fa += s_a[g * cpr + threadIdx.x]; fb += s_b[g * cpr + threadIdx.x]; }
            sums[threadIdx.x] = fa; sums[k + threadIdx.x] = fb;
        }
        if (threadIdx.x == 0) *ticket = 0;
    }
}
__global__ void agx_rms_merge_kernel(const double* __restrict__ sums, int k, double n_total, double* mean, double* var, double* count) {
    const int c = threadIdx.x;
    const double cnt = count[0];
    __syncthreads();  // every thread has read the count before thread 0 rewrites it
    if (c < k) {
        const double bm = sums[c] / n_total, bv = (sums[k + c] - n_total * bm * bm) / (n_total - 1.0);
        const double delta = bm - mean[c], tot = cnt + n_total;
        const double m2 = var[c] * cnt + bv * n_total + delta * delta * cnt * n_total / tot;
        mean[c] = mean[c] + delta * n_total / tot;
        var[c] = m2 / tot;
    }
    if (c == 0) count[0] = cnt + n_total;
}

// ---- after env.step of a rollout (agx.h AgxPostIO) ------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) agx_rollout_post_kernel(const __grid_constant__ AgxPostIO io, int64_t n) {
    __shared__ double s_part[8][4];
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < n; e += (int64_t)gridDim.x * 256) {
        const float r = io.reward[e];
        float sh = (r + io.shift) * io.scale;                       // DefaultRewardsShaper (tr_helpers.py:16-42)
        sh = fminf(fmaxf(sh, io.min_val), io.max_val);
        if (io.bootstrap && io.timeout[e]) sh += io.gamma * io.values[e * io.ld_values];  // a2c_base.py:675-676
        io.rewards_out[e * io.ld_rewards] = sh;
        const float cr = io.cur_reward[e] + r, cs = io.cur_shaped[e] + sh, cl = io.cur_length[e] + 1.0f;
        const bool d = io.reset_u8[e] != 0;
        if (d) { acc[0] += cr; acc[1] += cs; acc[2] += cl; acc[3] += 1.0; }
        io.cur_reward[e] = d ? 0.0f : cr;
        io.cur_shaped[e] = d ? 0.0f : cs;
        io.cur_length[e] = d ? 0.0f : cl;
        io.dones_state[e] = d ? 1 : 0;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
        if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5][k] = acc[k];
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += s_part[w][threadIdx.x];
        if (t != 0.0) atomicAdd(io.ep_stats + threadIdx.x, t);
    }
}

}  // namespace

extern "C" {

int64_t agx_col_sums_workspace_doubles(void) { return 8 + (int64_t)kSumsGridMax * 2 * kSumsMaxK; }

int agx_col_sums(const float* x, int64_t n, int k, int64_t ld, double* sums, double* workspace, void* stream) {
    if (!x || !sums || !workspace || n <= 0 || k <= 0 || k > kSumsMaxK || ld < k) return fail_ppo(AGX_ERR_ARG, "agx_col_sums: bad argument (k <= 128)");
    const int cpr = k <= 32 ? 32 : (k <= 64 ? 64 : 128), rgs = kSumsBlock / cpr;
    int64_t grid = (n + rgs * 32 - 1) / (rgs * 32);
    if (grid > 148) grid = 148;
    agx_col_sums_kernel<<<(unsigned)grid, kSumsBlock, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, n, k, ld, cpr, sums, workspace);
    return cudaGetLastError() == cudaSuccess ? AGX_OK : fail_ppo(AGX_ERR_CUDA, "agx_col_sums: launch failed");
}

int agx_rms_merge(const double* sums, int k, double n_total, double* mean, double* var, double* count, void* stream) {
    if (!sums || !mean || !var || !count || k <= 0 || k > kSumsMaxK || n_total < 2.0) return fail_ppo(AGX_ERR_ARG, "agx_rms_merge: bad argument");
    agx_rms_merge_kernel<<<1, kSumsMaxK, 0, reinterpret_cast<cudaStream_t>(stream)>>>(sums, k, n_total, mean, var, count);
    return cudaGetLastError() == cudaSuccess ? AGX_OK : fail_ppo(AGX_ERR_CUDA, "agx_rms_merge: launch failed");
}

int agx_sizeof_post_io(void) { return (int)sizeof(AgxPostIO); }

int agx_rollout_post(const AgxPostIO* io, int64_t n, void* stream) {
    if (!io || n <= 0 || !io->reward || !io->reset_u8 || !io->rewards_out || !io->cur_reward || !io->cur_shaped || !io->cur_length ||
        !io->dones_state || !io->ep_stats || (io->bootstrap && (!io->timeout || !io->values)))
        return fail_ppo(AGX_ERR_ARG, "agx_rollout_post: bad argument");
    int64_t grid = (n + 255) / 256;
    if (grid > 592) grid = 592;
    agx_rollout_post_kernel<<<(unsigned)grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*io, n);
    return cudaGetLastError() == cudaSuccess ? AGX_OK : fail_ppo(AGX_ERR_CUDA, "agx_rollout_post: launch failed");
}


int agx_gae(int64_t n, int h, float gamma, float tau, const float* rewards, const float* values, const uint8_t* dones,
            const float* last_values, const uint8_t* last_dones, float* adv, float* returns, void* stream) {
    if (n < 0 || h <= 0 || !rewards || !values || !dones || !last_values || !last_dones || !adv || !returns)
        return fail_ppo(AGX_ERR_ARG, "agx_gae: bad argument");
    if (n == 0) return AGX_OK;
    const size_t smem = (size_t)kGaeBlock * (4 * (h | 1) * sizeof(float) + h);
    if (smem > 200 * 1024) return fail_ppo(AGX_ERR_UNSUPPORTED, "agx_gae: horizon too long for the shared-memory tile");
    if (smem > 48 * 1024) cudaFuncSetAttribute(agx_gae_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const unsigned grid = (unsigned)((n + kGaeBlock - 1) / kGaeBlock);
    agx_gae_kernel<<<grid, kGaeBlock, smem, reinterpret_cast<cudaStream_t>(stream)>>>(n, h, gamma, tau, rewards, values, dones,
                                                                                     last_values, last_dones, adv, returns);
    return cudaGetLastError() == cudaSuccess ? AGX_OK : fail_ppo(AGX_ERR_CUDA, "agx_gae: launch failed");
}

int64_t agx_ppo_workspace_floats(void) { return (int64_t)kLossGridMax * kPartial + 4; }

int agx_ppo_loss(const AgxPpoHyper* hp, int64_t b, int a, const float* mu, const float* logstd, const float* value,
                 const float* actions, const float* old_neglogp, const float* adv, const float* returns, float* old_mu,
                 float* old_sigma, float* grad_mu, float* grad_value, float* grad_logstd, float* stats, float* workspace,
                 void* stream) {
    if (!hp || b <= 0 || !mu || !logstd || !value || !actions || !old_neglogp || !adv || !returns || !old_mu || !old_sigma ||
        !grad_mu || !grad_value || !grad_logstd || !stats || !workspace)
        return fail_ppo(AGX_ERR_ARG, "agx_ppo_loss: bad argument");
    if (a != 4 && a != 5) return fail_ppo(AGX_ERR_UNSUPPORTED, "agx_ppo_loss: actions_num must be 4 or 5");
    int64_t grid = (b + kLossBlock - 1) / kLossBlock;
    if (grid > kLossGridMax) grid = kLossGridMax;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (a == 4)
        agx_ppo_loss_kernel<4><<<(unsigned)grid, kLossBlock, 0, st>>>(*hp, b, mu, logstd, value, actions, old_neglogp, adv, returns,
                                                                       old_mu, old_sigma, grad_mu, grad_value, grad_logstd, stats, workspace);
    else
        agx_ppo_loss_kernel<5><<<(unsigned)grid, kLossBlock, 0, st>>>(*hp, b, mu, logstd, value, actions, old_neglogp, adv, returns,
                                                                       old_mu, old_sigma, grad_mu, grad_value, grad_logstd, stats, workspace);
    return cudaGetLastError() == cudaSuccess ? AGX_OK : fail_ppo(AGX_ERR_CUDA, "agx_ppo_loss: launch failed");
}

int agx_adam_step(const AgxPpoHyper* hp, int64_t n_params, float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                  float* lr_dev, int64_t* step_dev, const float* kl_dev, float grad_scale, float* grad_norm_out, void* stream) {
    if (!hp || n_params <= 0 || !params || !grads || !exp_avg || !exp_avg_sq || !lr_dev || !step_dev)
        return fail_ppo(AGX_ERR_ARG, "agx_adam_step: bad argument");
    AgxComm none;
    memset(&none, 0, sizeof(none));
    agx_adam_kernel<false><<<kAdamCluster, kAdamBlock, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        *hp, none, n_params, 0, params, const_cast<float*>(grads), exp_avg, exp_avg_sq, lr_dev, reinterpret_cast<long long*>(step_dev), kl_dev,
        grad_scale, grad_norm_out);
    return cudaGetLastError() == cudaSuccess ? AGX_OK : fail_ppo(AGX_ERR_CUDA, "agx_adam_step: launch failed");
}

int agx_adam_step_allreduce(const AgxPpoHyper* hp, const AgxComm* comm, int64_t n_params, int64_t n_extra, float* params, float* grads,
                            float* exp_avg, float* exp_avg_sq, float* lr_dev, int64_t* step_dev, const float* kl_dev, float grad_scale,
                            float* grad_norm_out, void* stream) {
    if (!hp || !comm || n_params <= 0 || n_extra < 0 || !params || !grads || !exp_avg || !exp_avg_sq || !lr_dev || !step_dev)
        return fail_ppo(AGX_ERR_ARG, "agx_adam_step_allreduce: bad argument");
    if (comm->world < 1 || comm->world > AGX_COMM_MAX_RANKS || comm->rank < 0 || comm->rank >= comm->world || (comm->slot_bytes & 255) ||
        (n_params + n_extra) * 4 > comm->slot_bytes)
        return fail_ppo(AGX_ERR_ARG, "agx_adam_step_allreduce: bad communicator or message larger than its slot");
    for (int i = 0; i < comm->world; ++i)
        if (!comm->region[i]) return fail_ppo(AGX_ERR_ARG, "agx_adam_step_allreduce: unmapped peer region");
    agx_adam_kernel<true><<<kAdamCluster, kAdamBlock, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        *hp, *comm, n_params, n_extra, params, grads, exp_avg, exp_avg_sq, lr_dev, reinterpret_cast<long long*>(step_dev), kl_dev, grad_scale,
        grad_norm_out);
    return cudaGetLastError() == cudaSuccess ? AGX_OK : fail_ppo(AGX_ERR_CUDA, "agx_adam_step_allreduce: launch failed");
}

}  // extern "C"

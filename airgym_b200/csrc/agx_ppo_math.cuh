// agx_ppo_math.cuh — per-sample arithmetic of the PPO update (reference lib/agent/a2c_continuous.py:299-369,
// lib/core/common_losses.py:10-48, lib/core/torch_ext.py:27-36, lib/model/a2c_continuous_logstd_model.py:195-198),
// written once as __host__ __device__ so tests/hostsim can run it without a GPU.
#pragma once
#include <math.h>
#include <stdint.h>
#include "agx.h"

#if defined(__CUDACC__)
#define AGX_HD __host__ __device__ __forceinline__
#else
#define AGX_HD inline
#endif

namespace agx {

constexpr int kMaxAct = AGX_MAX_ACTIONS;

struct PpoSampleOut {
    float a_loss, c_loss, entropy, b_loss, kl;  // per-sample terms (means are taken by the caller)
    float g_mu[kMaxAct];      // d(loss)/d(mu_i)   * B   (caller scales by 1/B)
    float g_value;            // d(loss)/d(value)  * B
    float g_logstd[kMaxAct];  // d(loss)/d(logstd_i) * B  (entropy term excluded; it is sample independent)
};

// One sample of calc_gradients: clipped-surrogate actor loss, squared-error critic loss (clip_value False),
// soft bound loss on |mu| > 1.1, entropy of N(mu, sigma), and the KL(new || old) diagnostic that drives the LR.
AGX_HD void ppo_sample(const AgxPpoHyper& H, int A, const float* mu, const float* logstd, float value,
                       const float* action, float old_neglogp, float adv, float ret, const float* old_mu,
                       const float* old_sigma, PpoSampleOut& o) {
    const float kHalfLog2Pi = 0.91893853320467274178f;
    float nlp = 0.0f, sum_logstd = 0.0f, ent = 0.0f, b = 0.0f, kl = 0.0f;
    float z[kMaxAct], inv_sigma[kMaxAct];
#pragma unroll
    for (int i = 0; i < kMaxAct; ++i) {
        if (i < A) {
            const float sigma = expf(logstd[i]);
            inv_sigma[i] = 1.0f / sigma;
            z[i] = (action[i] - mu[i]) * inv_sigma[i];
            nlp += z[i] * z[i];
            sum_logstd += logstd[i];
            ent += 0.5f + kHalfLog2Pi + logstd[i];
            const float hi = mu[i] - 1.1f, lo = mu[i] + 1.1f;
            const float bh = hi > 0.0f ? hi : 0.0f, bl = lo < 0.0f ? lo : 0.0f;
            b += bh * bh + bl * bl;
            o.g_mu[i] = H.bounds_loss_coef * 2.0f * (bh + bl);
            const float dmu = old_mu[i] - mu[i];
            kl += logf(old_sigma[i] * inv_sigma[i] + 1e-5f) +
                  (sigma * sigma + dmu * dmu) / (2.0f * (old_sigma[i] * old_sigma[i] + 1e-5f)) - 0.5f;
        }
    }
    nlp = 0.5f * nlp + kHalfLog2Pi * (float)A + sum_logstd;
    const float ratio = expf(old_neglogp - nlp);
    const float lo = 1.0f - H.e_clip, hi = 1.0f + H.e_clip;
    const float clipped = ratio < lo ? lo : (ratio > hi ? hi : ratio);
    const float s1 = -adv * ratio, s2 = -adv * clipped;
    o.a_loss = s1 > s2 ? s1 : s2;
    // d a_loss / d nlp: inside the clip range both surrogates coincide (torch.max splits the tie, the sum is the same)
    const bool inside = (ratio >= lo) && (ratio <= hi);
    const float g_nlp = (inside || s1 > s2) ? adv * ratio : 0.0f;
    const float dv = ret - value;
    o.c_loss = dv * dv;
    o.g_value = 0.5f * H.critic_coef * (-2.0f * dv);
    o.entropy = ent;
    o.b_loss = b;
    o.kl = kl;
#pragma unroll
    for (int i = 0; i < kMaxAct; ++i) {
        if (i < A) {
            o.g_mu[i] += g_nlp * (-z[i] * inv_sigma[i]);
            o.g_logstd[i] = g_nlp * (1.0f - z[i] * z[i]);
        } else {
            o.g_mu[i] = 0.0f;
            o.g_logstd[i] = 0.0f;
        }
    }
}

// GAE over one env's horizon (lib/agent/a2c_base.py:463-478), env-major rows [H]:
// dones[t] is the done flag stored BEFORE step t (a2c_base.py:663), last_done the flag after the last step.
AGX_HD void gae_row(int Hn, float gamma, float tau, const float* rewards, const float* values, const uint8_t* dones,
                    float last_value, float last_done, float* adv, float* ret) {
    float lastgaelam = 0.0f;
    for (int t = Hn - 1; t >= 0; --t) {
        const float nonterm = (t == Hn - 1) ? 1.0f - last_done : 1.0f - (float)dones[t + 1];
        const float nextv = (t == Hn - 1) ? last_value : values[t + 1];
        const float delta = rewards[t] + gamma * nextv * nonterm - values[t];
        lastgaelam = delta + gamma * tau * nonterm * lastgaelam;
        adv[t] = lastgaelam;
        ret[t] = lastgaelam + values[t];
    }
}

// AdaptiveScheduler.update (lib/core/schedulers.py:19-32)
AGX_HD float adaptive_lr(float lr, float kl, float kl_threshold) {
    if (kl > 2.0f * kl_threshold) lr = fmaxf(lr / 1.5f, 1e-6f);
    if (kl < 0.5f * kl_threshold) lr = fminf(lr * 1.5f, 1e-2f);
    return lr;
}

}  // namespace agx

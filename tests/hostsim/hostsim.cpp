// hostsim.cpp — TEST INFRASTRUCTURE.  Compiles airgym_b200/csrc/agx_math.cuh (the exact text the sm_100a kernel
// instantiates per thread) with g++ and loops it over envs on the CPU, so the kernel's arithmetic can be
// debugged against oracle/ in a container without a GPU.  Never loaded by the product.
#include <stdint.h>
#include <vector>
#include "agx.h"
#include "agx_math.cuh"

using namespace agx;

template <int TASK, int MODE>
static void run(const AgxParams& P, int64_t n, const AgxStepIO& io) {
    constexpr int A = (MODE == AGX_CTL_ATTI) ? 5 : 4;
    const int K = P.ctrl_state_dim;
    for (int64_t env = 0; env < n; ++env) {
        EnvRegs e;
        for (int i = 0; i < 13; ++i) e.s[i] = io.state[env * 13 + i];
        for (int i = 0; i < 5; ++i) { e.a[i] = 0; e.pa[i] = 0; }
        for (int i = 0; i < A; ++i) { e.a[i] = io.action[env * A + i]; e.pa[i] = io.prev_action[env * A + i]; }
        for (int k = 0; k < AGX_CTRL_STATE_MAX; ++k) e.cs[k] = (k < K) ? io.ctrl_state[(int64_t)k * n + env] : 0.0f;
        e.progress = io.progress[env];
        e.pending = io.reset[env] != 0;
        for (int k = 0; k < AGX_AUX_MAX; ++k) e.aux[k] = io.aux ? io.aux[env * AGX_AUX_MAX + k] : 0.0f;
        e.reset = 0;
        if (io.phase == AGX_PHASE_TASK) {  // the TASK half reads the actions the PHYSICS half shaped, and has no pending reset
            for (int i = 0; i < A; ++i) e.a[i] = io.actions_out[env * A + i];
            e.pending = 0;
        }
        SceneRef sc;
        sc.assets_row = io.assets ? io.assets + env * (int64_t)AGX_ASSET_ROW : nullptr;
        sc.trees = io.trees;
        RandSrc rnd;
        rnd.reset_row = io.rand_reset ? io.rand_reset + env * (int64_t)(2 * P.reset_draws) : nullptr;
        rnd.noise_row = io.rand_noise ? io.rand_noise + env * (int64_t)AGX_NOISE_DRAWS : nullptr;
        const uint64_t genv = (uint64_t)(io.env_offset + env);
        rnd.ph.k0 = (uint32_t)io.seed; rnd.ph.k1 = (uint32_t)(io.seed >> 32);
        rnd.ph.env_lo = (uint32_t)genv; rnd.ph.env_hi = (uint32_t)(genv >> 32);
        rnd.ph.step_lo = (uint32_t)io.step; rnd.ph.step_hi = (uint32_t)(io.step >> 32);
        float z[AGX_NOISE_DRAWS];
        scaled_noise(P, rnd, z);
        float obs_row[48];
        env_step<TASK, MODE>(P, rnd, z, e, sc, obs_row, io.phase);
        const bool phys = io.phase != AGX_PHASE_TASK, task = io.phase != AGX_PHASE_PHYSICS;
        for (int i = 0; i < 13; ++i) io.state[env * 13 + i] = e.s[i];
        for (int i = 0; i < A; ++i) { if (phys) io.actions_out[env * A + i] = e.a[i]; io.prev_action[env * A + i] = e.pa[i]; }
        if (phys) {
            if ((P.flags & AGX_FLAG_MUTATE_ACTIONS) && (MODE == AGX_CTL_RATE || MODE == AGX_CTL_ATTI))
                io.action[env * A + (A - 1)] = e.a_last_remap;
            for (int k = 0; k < K; ++k) io.ctrl_state[(int64_t)k * n + env] = e.cs[k];
            if (io.cmd) for (int i = 0; i < 4; ++i) io.cmd[env * 4 + i] = e.cmd[i];
        }
        if (io.aux) for (int k = 0; k < AGX_AUX_MAX; ++k) io.aux[env * AGX_AUX_MAX + k] = e.aux[k];
        io.progress[env] = e.progress;
        if (task) {
            for (int i = 0; i < P.num_obs; ++i) io.obs[env * P.num_obs + i] = obs_row[i];
            io.reset[env] = e.reset;
            io.timeout[env] = (uint8_t)e.timeout;
            io.reward[env] = e.rew;
            const int nt = (TASK == AGX_TASK_PLANNING) ? 11 : 9;
            if (io.reward_terms) for (int k = 0; k < nt; ++k) io.reward_terms[(int64_t)k * n + env] = e.terms[k];
        }
    }
}

template <int TASK>
static int by_mode(const AgxParams& P, int64_t n, const AgxStepIO& io) {
    switch (P.ctl_mode) {
        case AGX_CTL_POS: run<TASK, AGX_CTL_POS>(P, n, io); return 0;
        case AGX_CTL_VEL: run<TASK, AGX_CTL_VEL>(P, n, io); return 0;
        case AGX_CTL_ATTI:
            if constexpr (TASK == AGX_TASK_AVOID || TASK == AGX_TASK_PLANNING) return -4;
            else { run<TASK, AGX_CTL_ATTI>(P, n, io); return 0; }
        case AGX_CTL_RATE: run<TASK, AGX_CTL_RATE>(P, n, io); return 0;
        case AGX_CTL_PROP: run<TASK, AGX_CTL_PROP>(P, n, io); return 0;
    }
    return -1;
}

extern "C" int hostsim_step(const AgxParams* p, int64_t n, const AgxStepIO* io) {
    if (p->task == AGX_TASK_HOVERING) return by_mode<AGX_TASK_HOVERING>(*p, n, *io);
    if (p->task == AGX_TASK_TRACKING) return by_mode<AGX_TASK_TRACKING>(*p, n, *io);
    if (p->task == AGX_TASK_BALLOON) return by_mode<AGX_TASK_BALLOON>(*p, n, *io);
    if (p->task == AGX_TASK_AVOID) return by_mode<AGX_TASK_AVOID>(*p, n, *io);
    if (p->task == AGX_TASK_PLANNING) return by_mode<AGX_TASK_PLANNING>(*p, n, *io);
    return -4;
}

// Host twin of agx_render_kernel (airgym_b200/csrc/agx_render.cu): same per-pixel functions, explicit noise only.
extern "C" int hostsim_render(const AgxParams* p, int64_t n, const AgxRenderIO* io) {
    const int W = AGX_CAM_W, H = AGX_CAM_H;
    std::vector<float> img(W * H);
    for (int64_t env = 0; env < n; ++env) {
        const Camera cam = make_camera(io->state + env * 13);
        const float* aux = io->aux + env * AGX_AUX_MAX;
        const V3 obj = v3(aux[0], aux[1], aux[2]);
        std::vector<Capsule> caps;
        std::vector<int> cu0, cu1;
        if (p->task == AGX_TASK_PLANNING) {
            const float* row = io->assets + env * (int64_t)AGX_ASSET_ROW;
            for (int j = 1; j < AGX_NUM_ASSETS; ++j) {
                const Capsule k = place_tree(io->trees + (j - 1) * 8, row[j], row[AGX_NUM_ASSETS + j], row[2 * AGX_NUM_ASSETS + j],
                                             row[3 * AGX_NUM_ASSETS + j]);
                int u0, u1;
                capsule_columns(cam, k, &u0, &u1);
                if (u0 <= u1) { caps.push_back(k); cu0.push_back(u0); cu1.push_back(u1); }
            }
        }
        float m = 0.0f;
        for (int u = 0; u < W; ++u)
            for (int v = 0; v < H; ++v) {
                const V3 d = pixel_dir(cam, u, v);
                float t = hit_ground(cam.o, d);
                if (p->task == AGX_TASK_PLANNING) {
                    for (size_t c = 0; c < caps.size(); ++c)
                        if (u >= cu0[c] && u <= cu1[c]) t = fminf(t, hit_capsule(cam.o, d, caps[c]));
                    t = fminf(t, hit_sphere(cam.o, d, obj, kBallRadius));
                } else {
                    t = fminf(t, hit_box(cam.o, d, obj, kCubeHalf));
                }
                img[u * H + v] = normalize_depth(t);
                m = fmaxf(m, img[u * H + v]);
            }
        for (int pass = 0; pass < 2; ++pass) {
            const float* ex = (pass == 0 ? io->rand_add : io->rand_mul) + env * (int64_t)W * H;
            float pm = 0.0f;
            for (int i = 0; i < W * H; ++i) {
                float t = pass == 0 ? img[i] + ex[i] : img[i] * ex[i];
                t = t < 0.0f ? 0.0f : t;
                t = t > m ? m : t;
                img[i] = t;
                pm = fmaxf(pm, t);
            }
            m = pm;
        }
        const float* kk = io->rand_kern + env * 25;
        float mn = kInf;
        for (int u = 0; u < W; ++u)
            for (int v = 0; v < H; ++v) {
                float acc = 0.0f;
                for (int i = 0; i < 5; ++i)
                    for (int j = 0; j < 5; ++j) {
                        const int uu = u + i - 2, vv = v + j - 2;
                        if (uu >= 0 && uu < W && vv >= 0 && vv < H) acc = fmaf(kk[i * 5 + j], img[uu * H + vv], acc);
                    }
                io->image[env * (int64_t)W * H + u * H + v] = acc;
                mn = fminf(mn, acc);
            }
        if (p->task == AGX_TASK_PLANNING) io->aux[env * AGX_AUX_MAX + 7] = mn;
    }
    return 0;
}

extern "C" int hostsim_philox_fill(float* out, int64_t n, int width, int stream_id, uint64_t seed, uint64_t step,
                                   int64_t env_offset) {
    for (int64_t env = 0; env < n; ++env) {
        PhiloxCtx ph;
        const uint64_t genv = (uint64_t)(env_offset + env);
        ph.k0 = (uint32_t)seed; ph.k1 = (uint32_t)(seed >> 32);
        ph.env_lo = (uint32_t)genv; ph.env_hi = (uint32_t)(genv >> 32);
        ph.step_lo = (uint32_t)step; ph.step_hi = (uint32_t)(step >> 32);
        float v[20];
        if (stream_id == 2) philox_normals(ph, width, v);
        else philox_uniforms(ph, (uint32_t)stream_id, width, v);
        for (int i = 0; i < width; ++i) out[env * width + i] = v[i];
    }
    return 0;
}

// ---- PPO math (agx_ppo_math.cuh) on the host ----------------------------------------------------------------------
#include "agx_ppo_math.cuh"

extern "C" int hostsim_gae(int64_t n, int h, float gamma, float tau, const float* rewards, const float* values,
                           const uint8_t* dones, const float* last_values, const uint8_t* last_dones, float* adv, float* ret) {
    for (int64_t e = 0; e < n; ++e)
        gae_row(h, gamma, tau, rewards + e * h, values + e * h, dones + e * h, last_values[e], (float)last_dones[e],
                adv + e * h, ret + e * h);
    return 0;
}

extern "C" int hostsim_ppo_loss(const AgxPpoHyper* hp, int64_t b, int a, const float* mu, const float* logstd,
                                const float* value, const float* actions, const float* old_neglogp, const float* adv,
                                const float* returns, float* old_mu, float* old_sigma, float* grad_mu, float* grad_value,
                                float* grad_logstd, float* stats) {
    double acc[16] = {0};
    for (int64_t s = 0; s < b; ++s) {
        float m[kMaxAct] = {0}, ac[kMaxAct] = {0}, om[kMaxAct] = {0}, os[kMaxAct] = {1, 1, 1, 1, 1};
        for (int i = 0; i < a; ++i) { m[i] = mu[s * a + i]; ac[i] = actions[s * a + i]; om[i] = old_mu[s * a + i]; os[i] = old_sigma[s * a + i]; }
        PpoSampleOut o;
        ppo_sample(*hp, a, m, logstd, value[s], ac, old_neglogp[s], adv[s], returns[s], om, os, o);
        for (int i = 0; i < a; ++i) {
            grad_mu[s * a + i] = o.g_mu[i] / (float)b;
            old_mu[s * a + i] = m[i];
            old_sigma[s * a + i] = expf(logstd[i]);
            acc[5 + i] += o.g_logstd[i];
        }
        grad_value[s] = o.g_value / (float)b;
        acc[0] += o.a_loss; acc[1] += o.c_loss; acc[2] += o.entropy; acc[3] += o.b_loss; acc[4] += o.kl;
    }
    for (int i = 0; i < 5; ++i) stats[i] = (float)(acc[i] / (double)b);
    for (int i = 0; i < a; ++i) grad_logstd[i] = (float)(acc[5 + i] / (double)b) - hp->entropy_coef;
    return 0;
}

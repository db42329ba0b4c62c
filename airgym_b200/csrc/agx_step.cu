// agx_step.cu — the C ABI declared in include/agx.h for the env path (the kernel template lives in
// agx_step_kernel.cuh and is instantiated per task in agx_step_<task>.cu).
//
// One thread = one env.  A CTA owns a tile of BLOCK consecutive envs:
//   * the [BLOCK,13] root-state rows (52-B rows, not 16-B aligned individually) are one contiguous,
//     16-B aligned span, fetched with a single TMA bulk copy (cp.async.bulk → SASS UBLKCP) into shared
//     memory behind an mbarrier; each thread then reads its row at stride 13 words (bank-conflict free),
//   * everything else a thread needs is already coalesced (float4 action rows, SoA controller planes,
//     int64 progress/reset),
//   * results go back the same way: the state tile and the [BLOCK,18] observation tile are written to
//     shared memory and leave with one bulk store each; 48/16-wide observation rows use a padded
//     shared layout and a coalesced float4 copy-out instead (dense rows would be 16-way bank conflicted).
// A partial last tile (or a caller buffer that breaks the 16-B rule) takes the cooperative-copy path.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "agx.h"
#include "agx_math.cuh"
#include "agx_step_kernel.cuh"

namespace {
thread_local char g_err[512] = "";
}  // namespace

namespace agxk {
int g_block = 128;
int g_use_bulk = 1;
int g_pdl = -1;  // -1 auto: noise-first PDL (mode 3) for grids of at most ~one wave, where the kernel boundary dominates
                 // (65 536 envs: 10.3 -> 7.9 us/step); off for multi-wave grids, where early CTAs only steal slots (4 M envs: 357 -> 402 us)
int g_sm_count = 0;
int g_balanced = 0;  // equal tiles on a multiple of the SM count for single-wave grids (agx_set_option("balanced", 0|1)): measured 7.28 vs 7.19 us/step at 65 536 envs, i.e. no gain — per-SM load balance is not the limiter — so off by default
int sm_count() {
    if (g_sm_count == 0) {
        int dev = 0, sms = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess) g_sm_count = sms;
        if (g_sm_count <= 0) g_sm_count = 148;
    }
    return g_sm_count;
}

int fail(int code, const char* fmt, const char* detail) {
    snprintf(g_err, sizeof(g_err), fmt, detail);
    return code;
}

// one translation unit per task (agx_step_<task>.cu)
extern template int agx_dispatch_task<AGX_TASK_HOVERING>(const AgxParams&, int64_t, const AgxStepIO&, cudaStream_t);
extern template int agx_dispatch_task<AGX_TASK_TRACKING>(const AgxParams&, int64_t, const AgxStepIO&, cudaStream_t);
extern template int agx_dispatch_task<AGX_TASK_BALLOON>(const AgxParams&, int64_t, const AgxStepIO&, cudaStream_t);
extern template int agx_dispatch_task<AGX_TASK_AVOID>(const AgxParams&, int64_t, const AgxStepIO&, cudaStream_t);
extern template int agx_dispatch_task<AGX_TASK_PLANNING>(const AgxParams&, int64_t, const AgxStepIO&, cudaStream_t);
extern template int agx_observe_task<AGX_TASK_HOVERING>(const AgxParams&, int64_t, const AgxStepIO&, cudaStream_t, int);
extern template int agx_observe_task<AGX_TASK_TRACKING>(const AgxParams&, int64_t, const AgxStepIO&, cudaStream_t, int);
extern template int agx_observe_task<AGX_TASK_BALLOON>(const AgxParams&, int64_t, const AgxStepIO&, cudaStream_t, int);
extern template int agx_observe_task<AGX_TASK_AVOID>(const AgxParams&, int64_t, const AgxStepIO&, cudaStream_t, int);
extern template int agx_observe_task<AGX_TASK_PLANNING>(const AgxParams&, int64_t, const AgxStepIO&, cudaStream_t, int);
}  // namespace agxk

using namespace agxk;

int agx_internal_fail(int code, const char* msg) { return fail(code, "%s", msg); }
extern "C" int agx_internal_mlp_option(const char* key, int value);  // agx_mlp.cu
extern "C" int agx_internal_conv_option(const char* key, int value);  // agx_conv_tma.cu

namespace {

__global__ void agx_philox_fill_kernel(float* out, int64_t n, int width, int stream_id, uint64_t seed,
                                       uint64_t step, int64_t env_offset) {
    using namespace agx;
    const int64_t env = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= n) return;
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


bool misaligned(const void* p) { return p && (reinterpret_cast<uintptr_t>(p) & 15u) != 0; }

}  // namespace

// ---- C ABI ---------------------------------------------------------------------------------------------
extern "C" {

int agx_version(void) { return AGX_VERSION; }
const char* agx_error_string(void) { return g_err; }
int agx_sizeof_params(void) { return (int)sizeof(AgxParams); }
int agx_sizeof_step_io(void) { return (int)sizeof(AgxStepIO); }
int agx_sizeof_render_io(void) { return (int)sizeof(AgxRenderIO); }

int agx_set_option(const char* key, int value) {
    if (!key) return fail(AGX_ERR_ARG, "agx_set_option: null key%s");
    if (!strcmp(key, "block")) {
        if (value != 64 && value != 128) return fail(AGX_ERR_ARG, "agx_set_option: block must be 64|128%s");
        g_block = value;
        return AGX_OK;
    }
    if (!strcmp(key, "use_bulk")) { g_use_bulk = value ? 1 : 0; return AGX_OK; }
    if (!strcmp(key, "balanced")) { g_balanced = value ? 1 : 0; return AGX_OK; }
    if (!strcmp(key, "pdl")) {
        if (value < -1 || value > 3) return fail(AGX_ERR_ARG, "agx_set_option: pdl must be -1 (auto) or 0..3%s");
        g_pdl = value;
        return AGX_OK;
    }
    int r = agx_internal_mlp_option(key, value);
    if (r == 0) r = agx_internal_conv_option(key, value);
    if (r == 1) return AGX_OK;
    if (r < 0) return fail(AGX_ERR_ARG, "agx_set_option: bad value for '%s'", key);
    return fail(AGX_ERR_ARG, "agx_set_option: unknown key '%s'", key);
}

int agx_params_default(AgxParams* p, int task, int ctl_mode) {
    if (!p) return fail(AGX_ERR_ARG, "agx_params_default: null params%s");
    if (task < AGX_TASK_HOVERING || task > AGX_TASK_PLANNING) return fail(AGX_ERR_ARG, "agx_params_default: unknown task%s");
    const bool image_task = (task == AGX_TASK_AVOID || task == AGX_TASK_PLANNING);
    if (image_task && ctl_mode == AGX_CTL_ATTI)
        return fail(AGX_ERR_UNSUPPORTED, "agx_params_default: avoid/planning have no atti mode (obs[12:16] holds the 4 actions, avoid.py:226)%s");
    if (ctl_mode < AGX_CTL_POS || ctl_mode > AGX_CTL_PROP) return fail(AGX_ERR_ARG, "agx_params_default: bad ctl_mode%s");
    memset(p, 0, sizeof(*p));
    const double pi = 3.14159265358979323846;
    p->task = task;
    p->ctl_mode = ctl_mode;
    p->num_actions = (ctl_mode == AGX_CTL_ATTI) ? 5 : 4;
    p->num_obs = (task == AGX_TASK_TRACKING) ? 48 : (image_task ? 16 : 18);
    p->integrator = AGX_INT_RK4;
    p->flags = AGX_FLAG_MUTATE_ACTIONS;  // (collision flag for the Customized family is added below)
    p->dt = 0.01f;
    const double episode_s = (task == AGX_TASK_TRACKING) ? 36.0 : (task == AGX_TASK_BALLOON ? 8.0 : (task == AGX_TASK_AVOID ? 6.0 : (task == AGX_TASK_PLANNING ? 16.0 : 24.0)));  // *_config.py episode_length_s
    p->max_episode_length = (int)(episode_s / 0.01);
    p->ctrl_state_dim = (ctl_mode == AGX_CTL_PROP) ? 0 : ((ctl_mode == AGX_CTL_RATE || ctl_mode == AGX_CTL_ATTI) ? 6 : 12);
    p->reset_draws = (task == AGX_TASK_BALLOON) ? 15 : (task == AGX_TASK_AVOID ? 11 : (task == AGX_TASK_PLANNING ? AGX_PLANNING_DRAWS : 12));
    p->collision_radius = 0.2f;
    if (task == AGX_TASK_BALLOON || task == AGX_TASK_AVOID) p->flags |= AGX_FLAG_RESET_ON_COLLISION;  // balloon_config.py:19, avoid_config.py:19 (planning: False)
    p->gravity = 9.81f;
    const double m_base = 0.585, m_prop = 0.004, arm = 0.05374, hz = 0.024;
    p->mass = (float)(m_base + 4.0 * m_prop);
    p->inertia[0] = (float)(0.04 + 4.0 * (1e-6 + m_prop * (arm * arm + hz * hz)));
    p->inertia[1] = p->inertia[0];
    p->inertia[2] = (float)(0.04 + 4.0 * (1e-6 + m_prop * (2.0 * arm * arm)));
    p->arm = (float)arm;
    p->k_thrust = 9.59f;
    p->k_torque = 0.2f;
    p->max_lin_vel = 100.0f;
    p->max_ang_vel = 100.0f;
    const float lim_pos = (task == AGX_TASK_TRACKING) ? 6.0f : 3.0f;
    float lo[5] = {0, 0, 0, 0, 0}, hi[5] = {0, 0, 0, 0, 0};
    switch (ctl_mode) {
        case AGX_CTL_POS: for (int i = 0; i < 3; ++i) { lo[i] = -lim_pos; hi[i] = lim_pos; } lo[3] = -6; hi[3] = 6; break;
        case AGX_CTL_VEL: for (int i = 0; i < 4; ++i) { lo[i] = -6; hi[i] = 6; } break;
        case AGX_CTL_ATTI: for (int i = 0; i < 4; ++i) { lo[i] = -1; hi[i] = 1; } lo[4] = 0; hi[4] = 1; break;
        case AGX_CTL_RATE: {  // Customized family: +-1 rad/s (customized.py:109-113), Hovering/Tracking +-6
            const float r = (task == AGX_TASK_BALLOON || task == AGX_TASK_AVOID || task == AGX_TASK_PLANNING) ? 1.0f : 6.0f;
            for (int i = 0; i < 3; ++i) { lo[i] = -r; hi[i] = r; }
            lo[3] = 0; hi[3] = 1;
            break;
        }
        case AGX_CTL_PROP: for (int i = 0; i < 4; ++i) { lo[i] = 0; hi[i] = 1; } break;
    }
    for (int i = 0; i < 5; ++i) { p->act_lo[i] = lo[i]; p->act_hi[i] = hi[i]; }
    const float rp[3] = {0.15f, 0.15f, 0.2f}, ri[3] = {0.2f, 0.2f, 0.1f}, rd[3] = {0.003f, 0.003f, 0.0f};
    const float ap[3] = {6.5f, 6.5f, 2.8f};
    const double arl[3] = {220.0, 220.0, 200.0};
    const float vp[3] = {1.8f, 1.8f, 4.0f}, vi[3] = {0.4f, 0.4f, 2.0f}, vd[3] = {0.2f, 0.2f, 0.0f};
    const float vil[3] = {1.0f, 1.0f, 2.0f}, pp[3] = {0.95f, 0.95f, 1.0f}, vsl[3] = {6.0f, 6.0f, 6.0f};
    for (int i = 0; i < 3; ++i) {
        p->rate_p[i] = rp[i]; p->rate_i[i] = ri[i]; p->rate_d[i] = rd[i];
        p->att_p[i] = ap[i]; p->att_rate_lim[i] = (float)(arl[i] * pi / 180.0);
        p->vel_p[i] = vp[i]; p->vel_i[i] = vi[i]; p->vel_d[i] = vd[i]; p->vel_int_lim[i] = vil[i];
        p->pos_p[i] = pp[i]; p->vel_sp_lim[i] = vsl[i];
    }
    p->rate_int_lim = 0.3f;
    p->rate_i_fade = (float)(400.0 * pi / 180.0);
    p->att_yaw_w = 0.4f;
    p->hover_thrust = (float)((m_base + 4.0 * m_prop) * 9.81 / (4.0 * 9.59));
    p->tilt_max_tan = 1.0f;
    p->thr_min = 0.0f;
    p->thr_max = 1.0f;
    const float tgt[18] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 18; ++i) p->target[i] = tgt[i];
    if (task == AGX_TASK_AVOID) p->target[11] = 1.0f;  // avoid_config.py:11: hover target at z = 1
    p->target_yaw = 0.0f;  // atan2(-0, 1)
    p->noise_sigma[0] = 1e-3f; p->noise_sigma[1] = 5e-3f; p->noise_sigma[2] = 2e-2f; p->noise_sigma[3] = 4e-1f;
    return AGX_OK;
}

int agx_step(const AgxParams* p, int64_t n, const AgxStepIO* io, void* stream) {
    if (!p || !io) return fail(AGX_ERR_ARG, "agx_step: null params/io%s");
    if (n < 0) return fail(AGX_ERR_ARG, "agx_step: n < 0%s");
    if (!io->state || !io->action || !io->actions_out || !io->prev_action || !io->progress || !io->reset ||
        !io->timeout || !io->obs || !io->reward)
        return fail(AGX_ERR_ARG, "agx_step: a required buffer is null%s");
    if (p->ctrl_state_dim > 0 && !io->ctrl_state) return fail(AGX_ERR_ARG, "agx_step: ctrl_state is null%s");
    const void* ptrs[] = {io->state, io->action, io->actions_out, io->prev_action, io->ctrl_state, io->progress,
                          io->reset, io->obs, io->reward, io->cmd, io->reward_terms, io->rand_reset, io->rand_noise};
    for (const void* q : ptrs)
        if (misaligned(q)) return fail(AGX_ERR_ALIGN, "agx_step: buffer not 16-byte aligned%s");
    const int want_actions = (p->ctl_mode == AGX_CTL_ATTI) ? 5 : 4;
    if (p->num_actions != want_actions) return fail(AGX_ERR_ARG, "agx_step: num_actions does not match ctl_mode%s");
    {
        const int want = (p->task == AGX_TASK_BALLOON) ? 15 : (p->task == AGX_TASK_AVOID ? 11 : (p->task == AGX_TASK_PLANNING ? AGX_PLANNING_DRAWS : 12));
        if (p->reset_draws != want) return fail(AGX_ERR_ARG, "agx_step: reset_draws does not match the task%s");
    }
    const bool image_task = (p->task == AGX_TASK_AVOID || p->task == AGX_TASK_PLANNING);
    if (io->phase < AGX_PHASE_FUSED || io->phase > AGX_PHASE_TASK || (!image_task && io->phase != AGX_PHASE_FUSED))
        return fail(AGX_ERR_ARG, "agx_step: bad phase (only avoid/planning split the step)%s");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch (p->task) {
        case AGX_TASK_HOVERING:
            if (p->num_obs != 18) return fail(AGX_ERR_ARG, "agx_step: hovering needs num_obs=18%s");
            return agx_dispatch_task<AGX_TASK_HOVERING>(*p, n, *io, st);
        case AGX_TASK_TRACKING:
            if (p->num_obs != 48) return fail(AGX_ERR_ARG, "agx_step: tracking needs num_obs=48%s");
            return agx_dispatch_task<AGX_TASK_TRACKING>(*p, n, *io, st);
        case AGX_TASK_BALLOON:
            if (p->num_obs != 18) return fail(AGX_ERR_ARG, "agx_step: balloon needs num_obs=18%s");
            if (!io->aux || misaligned(io->aux)) return fail(AGX_ERR_ARG, "agx_step: balloon needs a 16-byte aligned aux buffer%s");
            return agx_dispatch_task<AGX_TASK_BALLOON>(*p, n, *io, st);
        case AGX_TASK_AVOID:
            if (p->num_obs != 16) return fail(AGX_ERR_ARG, "agx_step: avoid needs num_obs=16%s");
            if (!io->aux || misaligned(io->aux)) return fail(AGX_ERR_ARG, "agx_step: avoid needs a 16-byte aligned aux buffer%s");
            return agx_dispatch_task<AGX_TASK_AVOID>(*p, n, *io, st);
        case AGX_TASK_PLANNING:
            if (p->num_obs != 16) return fail(AGX_ERR_ARG, "agx_step: planning needs num_obs=16%s");
            if (!io->aux || misaligned(io->aux)) return fail(AGX_ERR_ARG, "agx_step: planning needs a 16-byte aligned aux buffer%s");
            if (!io->assets || misaligned(io->assets) || !io->trees) return fail(AGX_ERR_ARG, "agx_step: planning needs the assets and trees buffers%s");
            return agx_dispatch_task<AGX_TASK_PLANNING>(*p, n, *io, st);
        default: return fail(AGX_ERR_UNSUPPORTED, "agx_step: unknown task%s");
    }
}

int agx_observe(const AgxParams* p, int64_t n, const AgxStepIO* io, int what, void* stream) {
    if (!p || !io || n < 0 || what < 1 || what > 3) return fail(AGX_ERR_ARG, "agx_observe: bad argument (what = 1 observations | 2 reward | 3 both)%s");
    if (!io->state || !io->actions_out || !io->prev_action || !io->progress || !io->reset || !io->obs || !io->reward)
        return fail(AGX_ERR_ARG, "agx_observe: a required buffer is null%s");
    const void* ptrs[] = {io->state, io->actions_out, io->prev_action, io->progress, io->reset, io->obs, io->reward, io->cmd, io->reward_terms, io->rand_noise, io->aux};
    for (const void* q : ptrs)
        if (misaligned(q)) return fail(AGX_ERR_ALIGN, "agx_observe: buffer not 16-byte aligned%s");
    const bool needs_aux = (p->task == AGX_TASK_BALLOON || p->task == AGX_TASK_AVOID || p->task == AGX_TASK_PLANNING);
    if (needs_aux && !io->aux) return fail(AGX_ERR_ARG, "agx_observe: this task needs the aux buffer%s");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch (p->task) {
        case AGX_TASK_HOVERING: return agx_observe_task<AGX_TASK_HOVERING>(*p, n, *io, st, what);
        case AGX_TASK_TRACKING: return agx_observe_task<AGX_TASK_TRACKING>(*p, n, *io, st, what);
        case AGX_TASK_BALLOON: return agx_observe_task<AGX_TASK_BALLOON>(*p, n, *io, st, what);
        case AGX_TASK_AVOID: return agx_observe_task<AGX_TASK_AVOID>(*p, n, *io, st, what);
        case AGX_TASK_PLANNING:
            if (!io->assets || !io->trees) return fail(AGX_ERR_ARG, "agx_observe: planning needs the assets and trees buffers%s");
            return agx_observe_task<AGX_TASK_PLANNING>(*p, n, *io, st, what);
        default: return fail(AGX_ERR_UNSUPPORTED, "agx_observe: unknown task%s");
    }
}

int agx_reset_idx(const AgxParams* p, int64_t n, int64_t m, const int64_t* env_ids, float* state,
                  float* prev_action, float* ctrl_state, int64_t* progress, int64_t* reset, float* aux, float* assets,
                  const float* rand, uint64_t seed, uint64_t step, int64_t env_offset, void* stream) {
    if (!p || !env_ids || !state || !prev_action || !progress || !reset) return fail(AGX_ERR_ARG, "agx_reset_idx: null argument%s");
    if (n < 0 || m < 0) return fail(AGX_ERR_ARG, "agx_reset_idx: negative size%s");
    if (m == 0) return AGX_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const unsigned grid = (unsigned)((m + 127) / 128);
    if (p->task == AGX_TASK_HOVERING)
        agx_reset_idx_kernel<AGX_TASK_HOVERING><<<grid, 128, 0, st>>>(*p, n, m, env_ids, state, prev_action, ctrl_state,
                                                                      progress, reset, aux, assets, rand, seed, step, env_offset);
    else if (p->task == AGX_TASK_TRACKING)
        agx_reset_idx_kernel<AGX_TASK_TRACKING><<<grid, 128, 0, st>>>(*p, n, m, env_ids, state, prev_action, ctrl_state,
                                                                      progress, reset, aux, assets, rand, seed, step, env_offset);
    else if (p->task == AGX_TASK_BALLOON) {
        if (!aux) return fail(AGX_ERR_ARG, "agx_reset_idx: balloon needs aux%s");
        agx_reset_idx_kernel<AGX_TASK_BALLOON><<<grid, 128, 0, st>>>(*p, n, m, env_ids, state, prev_action, ctrl_state,
                                                                     progress, reset, aux, assets, rand, seed, step, env_offset);
    } else if (p->task == AGX_TASK_AVOID) {
        if (!aux) return fail(AGX_ERR_ARG, "agx_reset_idx: avoid needs aux%s");
        agx_reset_idx_kernel<AGX_TASK_AVOID><<<grid, 128, 0, st>>>(*p, n, m, env_ids, state, prev_action, ctrl_state,
                                                                   progress, reset, aux, assets, rand, seed, step, env_offset);
    } else if (p->task == AGX_TASK_PLANNING) {
        if (!aux || !assets) return fail(AGX_ERR_ARG, "agx_reset_idx: planning needs aux and assets%s");
        agx_reset_idx_kernel<AGX_TASK_PLANNING><<<grid, 128, 0, st>>>(*p, n, m, env_ids, state, prev_action, ctrl_state,
                                                                      progress, reset, aux, assets, rand, seed, step, env_offset);
    } else
        return fail(AGX_ERR_UNSUPPORTED, "agx_reset_idx: unknown task%s");
    const cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return fail(AGX_ERR_CUDA, "agx_reset_idx launch: %s", cudaGetErrorString(err));
    return AGX_OK;
}

int agx_philox_fill(float* out, int64_t n, int width, int stream_id, uint64_t seed, uint64_t step,
                    int64_t env_offset, void* stream) {
    if (!out || n < 0 || width < 1 || width > 20 || stream_id < 0 || stream_id > 3)
        return fail(AGX_ERR_ARG, "agx_philox_fill: bad argument%s");
    if (n == 0) return AGX_OK;
    const unsigned grid = (unsigned)((n + 127) / 128);
    agx_philox_fill_kernel<<<grid, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(out, n, width, stream_id, seed,
                                                                                   step, env_offset);
    const cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return fail(AGX_ERR_CUDA, "agx_philox_fill launch: %s", cudaGetErrorString(err));
    return AGX_OK;
}

}  // extern "C"

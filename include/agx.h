/*
 * agx.h — C ABI of libagx.so, the B200 (sm_100a) batched quadrotor env-step library.
 *
 * This is the drop-in boundary for the hot path of emNavi/AirGym: every entry
 * point replaces a stretch of the reference's Python/torch/IsaacGym/rlPx4Controller step and is
 * what a ctypes binding on the reference side would call.
 *
 * Conventions
 *   - plain C: POD structs, raw device pointers, sizes; no C++/torch types.
 *   - all pointers in AgxStepIO are DEVICE pointers owned by the caller (torch tensors);
 *     `stream` is a cudaStream_t passed as void*; all work is stream-ordered, nothing syncs.
 *   - return value: 0 = AGX_OK, negative = error (agx_error_string() explains); no exceptions.
 *   - state rows are the reference's root-state rows: [px py pz | qx qy qz qw | vx vy vz | wx wy wz]
 *     f32, world frame, xyzw quaternion (reference airgym/envs/base/hovering.py:70-77).
 */
#ifndef AGX_H
#define AGX_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AGX_VERSION 200

enum AgxError {
    AGX_OK = 0,
    AGX_ERR_ARG = -1,      /* bad argument (null pointer, n < 0, unknown task/mode ...) */
    AGX_ERR_ALIGN = -2,    /* a buffer violates the documented alignment */
    AGX_ERR_CUDA = -3,     /* CUDA launch/runtime error; agx_error_string() holds the text */
    AGX_ERR_UNSUPPORTED = -4
};

/* task ids: reference airgym/envs/__init__.py:5-62 (names of the registered tasks) */
enum AgxTask {
    AGX_TASK_HOVERING = 0, /* airgym/envs/base/hovering.py */
    AGX_TASK_TRACKING = 1, /* airgym/envs/task/tracking.py */
    AGX_TASK_BALLOON = 2,  /* airgym/envs/task/balloon.py */
    AGX_TASK_AVOID = 3,    /* airgym/envs/task/avoid.py */
    AGX_TASK_PLANNING = 4  /* airgym/envs/task/planning.py */
};

/* control modes: reference --ctl_mode {pos,vel,atti,rate,prop} = PY/LV/CTA/CTBR/SRT
 * (airgym/utils/helpers.py:103, hovering.py:93-123) */
enum AgxCtlMode {
    AGX_CTL_POS = 0,  /* PY:   position + yaw        (A=4) */
    AGX_CTL_VEL = 1,  /* LV:   linear velocity + yaw (A=4) */
    AGX_CTL_ATTI = 2, /* CTA:  quaternion wxyz + collective thrust (A=5) */
    AGX_CTL_RATE = 3, /* CTBR: body rates + collective thrust (A=4) */
    AGX_CTL_PROP = 4  /* SRT:  single-rotor thrusts (A=4) */
};

enum AgxIntegrator { AGX_INT_RK4 = 0, AGX_INT_EULER = 1 };

enum AgxFlags {
    AGX_FLAG_MUTATE_ACTIONS = 1,   /* write the 0.5+0.5a remap of the last action column back into
                                      `action` (reference quirk Q4, hovering.py:212-215) */
    AGX_FLAG_CTRL_RESET = 2,       /* zero controller integrators on episode reset (reference never
                                      does: the rlPx4Controller objects are not told about resets) */
    AGX_FLAG_NO_NOISE = 4,         /* skip observation noise (debug/KATs) */
    AGX_FLAG_RESET_ON_COLLISION = 8 /* cfg.env.reset_on_collision (customized.py:328-330) */
};

#define AGX_MAX_ACTIONS 5
#define AGX_CTRL_STATE_MAX 12
#define AGX_RESET_DRAWS_MAX 16
#define AGX_NOISE_DRAWS 18
#define AGX_AUX_MAX 8           /* floats of per-env task state in AgxStepIO.aux */
#define AGX_REWARD_TERMS 12     /* planes of AgxStepIO.reward_terms (9 used by hovering/tracking/balloon/avoid, 11 by planning) */
#define AGX_NUM_TREES 40        /* `thin` group assets of the planning task (planning_config.py:74-78) */
#define AGX_NUM_ASSETS 41       /* goal ball + trees, in the reference's asset order (asset_manager.py:79-152) */
#define AGX_ASSET_ROW 164       /* floats per env in AgxStepIO.assets: x[41] | y[41] | cos(yaw)[41] | sin(yaw)[41] */
#define AGX_PLANNING_DRAWS 124  /* uniforms of one Planning.reset_idx: asset x[41] | y[41] | yaw[41] | goal y */
#define AGX_CAM_W 212           /* depth camera (avoid_config.py:55-68): width, height, horizontal fov 87 deg, far plane 5 m, */
#define AGX_CAM_H 120           /* mounted at body (0.15, 0, 0.1); images are stored width-major [W,H] like the reference's */
                                /* full_camera_array (customized.py:144,402) */

/* agx_step phases for the depth-camera tasks: the reference renders between gym.simulate and compute_reward
 * (customized.py:318-325), so on a render step the fused step is split around agx_render_depth */
enum AgxPhase {
    AGX_PHASE_FUSED = 0,   /* everything (all tasks; avoid/planning on the 3 of 4 steps without a render) */
    AGX_PHASE_PHYSICS = 1, /* pre-step reset, action shaping, controller, rigid body, object flight, contacts, progress += 1 */
    AGX_PHASE_TASK = 2     /* observations, reward, termination, end-of-step reset, time-outs */
};

/* Everything the fused step needs that is not per-env data.  Mirrors the reference's nested cfg
 * classes (hovering_config.py:8-69), URDF constants (assets/robots/X152b/model.urdf) and the
 * literals in hovering.py; controller gains are builder-defined. */
typedef struct AgxParams {
    int32_t task;              /* AgxTask */
    int32_t ctl_mode;          /* AgxCtlMode */
    int32_t num_actions;       /* 5 for atti else 4 (hovering.py:46) */
    int32_t num_obs;           /* 18 hovering/balloon, 48 tracking, 16 avoid/planning (atti is not available there: the
                                  reference writes its [N,A] actions into obs[12:16], avoid.py:226) */
    int32_t integrator;        /* AgxIntegrator */
    int32_t flags;             /* AgxFlags bit set */
    int32_t max_episode_length;/* int(episode_length_s / dt) (hovering.py:48) */
    int32_t ctrl_state_dim;    /* floats of controller state per env: 0 prop, 6 rate/atti, 12 vel/pos */
    int32_t reset_draws;       /* uniforms consumed by one reset_idx of this task (12 hovering/tracking) */
    int32_t _pad0;

    float dt;                  /* 0.01 (hovering_config.py:29) */
    float gravity;             /* 9.81, along -z (hovering_config.py:31) */
    float mass;                /* 0.585 + 4*0.004 (model.urdf:19,36) */
    float inertia[3];          /* composite diag inertia about base_link */
    float arm;                 /* 0.05374: |x|=|y| of the rotor joints (model.urdf:86-105) */
    float k_thrust;            /* 9.59 N per unit cmd per rotor (hovering.py:256) */
    float k_torque;            /* 0.2 N m per unit cmd (hovering.py:270) */
    float max_lin_vel;         /* 100 (assets/__init__.py:34-35) */
    float max_ang_vel;         /* 100 */

    float act_lo[AGX_MAX_ACTIONS]; /* action limits (hovering.py:93-123, tracking.py:95-123) */
    float act_hi[AGX_MAX_ACTIONS];

    /* PX4-aligned cascade gains (builder-defined, PX4 defaults) */
    float rate_p[3], rate_i[3], rate_d[3];
    float rate_int_lim;        /* 0.3 */
    float rate_i_fade;         /* rad(400 deg/s): integrator fade-out scale */
    float att_p[3];
    float att_yaw_w;           /* 0.4 */
    float att_rate_lim[3];     /* rad(220,220,200 deg/s) */
    float vel_p[3], vel_i[3], vel_d[3];
    float vel_int_lim[3];
    float pos_p[3];
    float vel_sp_lim[3];       /* clamp of the velocity set-point out of the position loop */
    float hover_thrust;        /* m g / (4 k_thrust) */
    float tilt_max_tan;        /* tan(45 deg) */
    float thr_min, thr_max;

    float target[18];          /* cfg.env.target_state (hovering_config.py:12) */
    float target_yaw;          /* third intrinsic-XYZ euler angle of target[0:9] = atan2(-t01, t00) */
    float noise_sigma[4];      /* 1e-3, 5e-3, 2e-2, 4e-1 (hovering.py:350-353) */
    float collision_radius;    /* 0.2: the drone's collision sphere (robots/X152b/model.urdf:13-18) */
    float _pad1;
} AgxParams;

/* Per-call buffer table.  N = number of envs, A = num_actions, K = ctrl_state_dim, D = reset_draws.
 * Alignment: every non-null pointer must be 16-byte aligned (torch allocations are 512-B aligned). */
typedef struct AgxStepIO {
    float*   state;        /* [N,13] in/out: root states (hovering.py:73)                      */
    float*   action;       /* [N,A]  in (out for the last column when AGX_FLAG_MUTATE_ACTIONS) */
    float*   actions_out;  /* [N,A]  out: shaped+clamped actions = reference self.actions      */
    float*   prev_action;  /* [N,A]  in/out: reference self.pre_actions                        */
    float*   ctrl_state;   /* [K,N]  in/out: controller integrators (SoA planes); NULL if K=0  */
    int64_t* progress;     /* [N]    in/out: reference progress_buf (int64, hovering.py:164)   */
    int64_t* reset;        /* [N]    in: pending resets, out: new resets (reference reset_buf) */
    uint8_t* timeout;      /* [N]    out: reference time_out_buf (bool)                        */
    float*   obs;          /* [N,num_obs] out: reference obs_buf                               */
    float*   reward;       /* [N]    out: reference rew_buf                                    */
    float*   cmd;          /* [N,4]  out: reference cmd_thrusts, or NULL                       */
    float*   reward_terms; /* [AGX_REWARD_TERMS,N] out: reference item_reward_info planes, or NULL */
    float*   aux;          /* [N,AGX_AUX_MAX] task state in/out, or NULL for hovering/tracking.  Balloon: ball xyz (balloon_states
                              positions), previous drone xyz (pre_root_positions), collision flag (collisions), pad */
    const float* rand_reset; /* [N,2,D] U[0,1) draws for the pre-/post-step reset, or NULL → Philox */
    const float* rand_noise; /* [N,18]  N(0,1) draws for the observation noise, or NULL → Philox  */
    uint64_t seed;         /* Philox key   (used when a rand_* pointer is NULL) */
    uint64_t step;         /* Philox counter word: the caller's global step index */
    uint64_t* step_dev;    /* optional DEVICE counter [2] = {step, ticket}: when non-NULL the kernel uses
                              step_dev[0] instead of `step` and the last CTA to retire increments it, so a
                              captured CUDA graph can be replayed without re-baking the step index */
    int64_t  env_offset;   /* global id of env 0 of this shard (partition-invariant RNG, §8e) */
    float*   assets;       /* [N,AGX_ASSET_ROW] planning only: world placement of the goal ball + 40 trees (in/out: rewritten on reset) */
    const float* trees;    /* [AGX_NUM_TREES,8] planning only: cylinder table, asset frame: centre xyz, unit axis xyz, radius, half length */
    int32_t  phase;        /* AgxPhase (avoid/planning); must be 0 for the other tasks */
    int32_t  _pad;
    uint8_t* reset_u8;     /* [N] out, optional (NULL = skip): the new reset flags once more as bytes, so that a caller shipping
                              results to the HOST can place them behind obs and reward in one block and read all three back
                              with a single copy (reset_buf itself stays int64 as the reference API has it, base_task.py:75) */
} AgxStepIO;

/* Depth camera + post-processing of one render step (replaces IsaacGym's camera sensor and Customized.dump_images,
 * customized.py:386-435): analytic ray cast of the ground plane, the thrown cube (avoid) or the 40 tree cylinders + goal ball
 * (planning) → planar depth clipped at 4.5 m / 4.5 → + N(0,0.1) clamped to [0, max] → x N(1,0.3) clamped to [0, max] →
 * 5x5 correlation with a random kernel (randint(0,256)/256, zero padding).  Run between the PHYSICS and TASK phases. */
typedef struct AgxRenderIO {
    const float* state;    /* [N,13] root states after the physics phase */
    float*       aux;      /* [N,AGX_AUX_MAX] task state: cube / goal position in; column 7 out = min over the image (Planning esdf_dist, planning.py:162-163) */
    const float* assets;   /* [N,AGX_ASSET_ROW] planning, else NULL */
    const float* trees;    /* [AGX_NUM_TREES,8] planning, else NULL */
    float*       image;    /* [N,AGX_CAM_W,AGX_CAM_H] out: reference full_camera_array[:,0] */
    const float* rand_add; /* [N,W,H] the additive noise samples, or NULL → Philox stream 4 */
    const float* rand_mul; /* [N,W,H] the multiplicative noise samples (mean 1), or NULL → Philox stream 4 */
    const float* rand_kern;/* [N,25]  the blur kernel, or NULL → Philox stream 5 */
    uint64_t seed, step;
    int64_t  env_offset;
    const uint64_t* step_dev; /* optional DEVICE step counter (AgxStepIO.step_dev): when non-NULL its value replaces `step` as the Philox
                                 counter word, so a captured CUDA graph draws fresh image noise on every replay */
} AgxRenderIO;

/* library info */
int         agx_version(void);
const char* agx_error_string(void);   /* thread-local text of the last error */
int         agx_sizeof_params(void);  /* sizeof(AgxParams): lets a binding check its struct mirror */
int         agx_sizeof_step_io(void);

/* Tuning knobs (process-wide): "block" = 64|128 threads per CTA, "use_bulk" = 0|1 (TMA bulk-copy staging vs
 * cooperative copies), "pdl" = programmatic dependent launch of agx_step: -1 auto (default: mode 3 for grids of at most
 * about one wave, else off), 0 off, 1 trigger at entry / wait at entry, 2 trigger before the stores, 3 noise-first — the
 * step's state-independent prologue (Philox counter, observation noise, L2 prefetch hints for its inputs) runs while the
 * previous kernel of the stream is still executing; every global read or write of caller data happens after
 * griddepcontrol.wait, so stream order of all memory effects is unchanged.
 * "mlp_forward" = 0 TF32 mma.sync kernel | 1 tcgen05 + TMEM kernel for inference calls only | 2 tcgen05 always (default);
 * "mlp_wgrad_staged" = 0|1 (cp.async-staged weight-gradient kernel). */
int agx_set_option(const char* key, int value);

/* Fill `p` with the defaults of (task, ctl_mode): replaces the reference's cfg classes + the
 * limits set in Hovering.__init__ (hovering.py:93-123) / Tracking.__init__ (tracking.py:95-123). */
int agx_params_default(AgxParams* p, int task, int ctl_mode);

/* One fused env step over n envs: replaces Hovering.step (hovering.py:286-308) =
 * pre_physics_step (:203-281, incl. the rlPx4Controller FFI :217-250) + gym.simulate (:290) +
 * compute_observations (:337-358) + compute_reward (:360-459) + reset_idx (:310-335) + time-outs (:304);
 * for AGX_TASK_TRACKING additionally tracking.py:159-296. */
int agx_step(const AgxParams* p, int64_t n, const AgxStepIO* io, void* stream);

/* compute_observations() / compute_reward() as stand-alone calls (hovering.py:337-358 / :360-459 and the task overrides): the TASK
 * phase of the step kernel over the CURRENT buffers without stepping — no reset, no progress / time-out / state / task-state write,
 * the Philox step counter is read but not advanced.  what = 1: obs (fresh observation noise, or io->rand_noise);
 * what = 2: reward, reset (overwritten, as compute_reward does), reset_u8, reward_terms, prev_action <- actions_out
 * (`self.pre_actions = self.actions.clone()`) and, for the Customized family, the task state `aux`
 * (`pre_root_positions = root_positions.clone()`); 3: both.  Reads actions_out (the shaped actions of the last step) and cmd. */
int agx_observe(const AgxParams* p, int64_t n, const AgxStepIO* io, int what, void* stream);

int agx_render_depth(const AgxParams* p, int64_t n, const AgxRenderIO* io, void* stream);
int agx_sizeof_render_io(void);

/* reset_idx(env_ids) as a standalone call (hovering.py:310-335, tracking.py:159-192):
 * env_ids [m] int64 device pointer; rand [m,D] U[0,1) draws in env_ids order or NULL → Philox. */
int agx_reset_idx(const AgxParams* p, int64_t n, int64_t m, const int64_t* env_ids,
                  float* state, float* prev_action, float* ctrl_state, int64_t* progress,
                  int64_t* reset, float* aux, float* assets, const float* rand, uint64_t seed, uint64_t step,
                  int64_t env_offset, void* stream);

/* Fill out[n, width] with the library's Philox4x32-10 stream `stream_id` (0 pre-reset uniforms,
 * 1 post-reset uniforms, 2 noise normals) exactly as agx_step would draw it — lets tests and
 * the oracle consume identical numbers. */
int agx_philox_fill(float* out, int64_t n, int width, int stream_id, uint64_t seed, uint64_t step,
                    int64_t env_offset, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * PPO update kernels.  Reference: lib/agent/a2c_base.py:463-478 (GAE), a2c_continuous.py:299-369
 * (calc_gradients), lib/core/common_losses.py:10-48, lib/core/torch_ext.py:27-36 (policy_kl), lib/core/schedulers.py:19-32,
 * a2c_base.py:293-316 (grad clip + optimizer step; torch.optim.Adam, a2c_continuous.py:401).
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct AgxPpoHyper {
    float e_clip;            /* 0.2  (ppo_hovering.yaml:52) */
    float critic_coef;       /* 2    (:58) — loss uses 0.5 * c_loss * critic_coef */
    float entropy_coef;      /* 0    (:50) */
    float bounds_loss_coef;  /* 1e-4 (:61) */
    float kl_threshold;      /* 0.008 (:45) adaptive LR */
    float grad_norm;         /* 1.5  (:48) clip_grad_norm_ max norm; <= 0 disables clipping */
    float beta1, beta2, eps; /* Adam 0.9 / 0.999 / 1e-8 (a2c_continuous.py:401) */
    float weight_decay;      /* 0 */
    int32_t adaptive_lr;     /* 1: lr_schedule adaptive ("legacy": after every minibatch, a2c_continuous.py:112-118) */
    int32_t _pad;
} AgxPpoHyper;

#define AGX_PPO_STATS 8  /* a_loss, c_loss, entropy, b_loss, kl, (3 spare) — means over the minibatch */

/* GAE + returns over env-major rollout rows.  rewards/values/adv/returns [n,h] f32, dones [n,h] u8 = the done flag
 * stored BEFORE each step (a2c_base.py:663), last_values [n], last_dones [n] u8 (flags after the last step). */
int agx_gae(int64_t n, int h, float gamma, float tau, const float* rewards, const float* values,
            const uint8_t* dones, const float* last_values, const uint8_t* last_dones, float* adv, float* returns,
            void* stream);

/* floats of scratch agx_ppo_loss needs (per-CTA partial sums + a ticket); the caller zero-fills it ONCE at allocation */
int64_t agx_ppo_workspace_floats(void);

/* One minibatch of calc_gradients, forward AND backward of everything after the network heads:
 * in:  mu [b,a], logstd [a], value [b] (normalised), actions [b,a], old_neglogp [b], adv [b], returns [b] (normalised),
 * i/o: old_mu/old_sigma [b,a] — read for the KL, then overwritten with the current mu/sigma
 *      (PPODataset.update_mu_sigma, lib/core/datasets.py:20-24),
 * out: grad_mu [b,a], grad_value [b] = d(mean loss)/d(.), grad_logstd [a], stats [AGX_PPO_STATS] (means).
 * Deterministic: per-CTA partials are combined in a fixed order by the last CTA. */
int agx_ppo_loss(const AgxPpoHyper* hp, int64_t b, int a, const float* mu, const float* logstd, const float* value,
                 const float* actions, const float* old_neglogp, const float* adv, const float* returns,
                 float* old_mu, float* old_sigma, float* grad_mu, float* grad_value, float* grad_logstd,
                 float* stats, float* workspace, void* stream);

/* Fused trancate_gradients_and_step (a2c_base.py:293-316) + adaptive LR (schedulers.py:19-32) on flat buffers:
 * grads are scaled by grad_scale (1/world after an all-reduce SUM), clipped to hp->grad_norm (torch clip_grad_norm_),
 * Adam-stepped with the learning rate held in lr_dev[0]; then lr_dev[0] is updated from kl_dev[0]*grad_scale for the
 * NEXT minibatch.  step_dev[0] is Adam's step count.  One CTA; no host sync anywhere. */
int agx_adam_step(const AgxPpoHyper* hp, int64_t n_params, float* params, const float* grads, float* exp_avg,
                  float* exp_avg_sq, float* lr_dev, int64_t* step_dev, const float* kl_dev, float grad_scale,
                  float* grad_norm_out, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Multi-GPU: the env path has no collective; the PPO update has one small all-reduce per minibatch
 * ([flat gradients ‖ loss statistics incl. the KL], reference lib/agent/a2c_base.py:293-309 + a2c_continuous.py:112-123) and a
 * few per epoch (advantage / RunningMeanStd moments).  These entry points run them as ONE device kernel per rank over NVLink
 * peer memory instead of a library collective: each rank owns a region in its HBM, mapped into every process of the node with
 * CUDA IPC; a collective pushes the message into every peer's region as {word, tag} pairs (one traversal of the fabric: no fence,
 * no separate flag), polls its own region until every peer's words carry this call's tag and adds the slots in rank order
 * (bitwise identical results on every rank).  Stream-ordered, graph-capturable
 * (the call counter lives in the region), no host sync.  All ranks must issue the same sequence of collectives.
 * ------------------------------------------------------------------------------------------------------------- */
#define AGX_IPC_HANDLE_BYTES 64
#define AGX_COMM_MAX_RANKS 8
enum AgxDtype { AGX_F32 = 0, AGX_F64 = 1 };

typedef struct AgxComm {
    int32_t rank, world;
    int64_t slot_bytes;                  /* capacity of one message; a multiple of 256 */
    void*   region[AGX_COMM_MAX_RANKS];  /* region[rank]: this process's own allocation; the others: IPC mappings of the peers' */
} AgxComm;

/* bytes of one rank's region for messages of up to slot_bytes (rounded up to a multiple of 256): header + 2 x world slots of
 * {word, tag} pairs (the flag travels with the data: 8 bytes per 32-bit word) */
int64_t agx_comm_region_bytes(int world, int64_t slot_bytes);
/* cudaMalloc + zero-fill a region on the current device; handle (AGX_IPC_HANDLE_BYTES, may be NULL) = its cudaIpcMemHandle_t */
int agx_comm_alloc(int64_t bytes, void** ptr, unsigned char* handle);
/* map a peer's region from its handle (cudaIpcOpenMemHandle, enables peer access) / unmap it / free an own region */
int agx_comm_open(const unsigned char* handle, void** ptr);
int agx_comm_close(void* ptr);
int agx_comm_free(void* ptr);
/* in-place SUM over the ranks of buf[n] (dtype AGX_F32 | AGX_F64), n * sizeof <= slot_bytes */
int agx_comm_allreduce(const AgxComm* c, void* buf, int64_t n, int dtype, void* stream);
/* synchronous read of the own region's header: number of collectives completed, and the error word (0 = none; else
 * (peer + 1) << 56 | call number of the first collective that timed out waiting for `peer`) */
int agx_comm_status(const AgxComm* c, uint64_t* seq, uint64_t* err, void* stream);

/* agx_adam_step with the all-reduce fused in front: grads [n_params + n_extra] is summed over the ranks in place (the extra
 * floats — the loss statistics, kl_dev pointing into them — ride in the same message), then scaled by grad_scale (1/world),
 * clipped, Adam-stepped; the learning rate follows the all-reduced KL.  One launch per minibatch. */
int agx_adam_step_allreduce(const AgxPpoHyper* hp, const AgxComm* comm, int64_t n_params, int64_t n_extra, float* params,
                            float* grads, float* exp_avg, float* exp_avg_sq, float* lr_dev, int64_t* step_dev,
                            const float* kl_dev, float grad_scale, float* grad_norm_out, void* stream);

/* ---- host staging buffers for the end-to-end path (actions in / results out through host memory every step) --------------
 * NUMA node of the GPU's PCIe root (sysfs), or -1 when unknown. */
int agx_device_numa_node(int device);
/* page-locked host buffer preferring `numa_node` (mmap + mbind + cudaHostRegister); *placed = 1 when the policy call succeeded.
 * numa_node < 0: plain pinned memory. */
int agx_host_alloc_pinned(int64_t bytes, int numa_node, void** ptr, int* placed);
int agx_host_free_pinned(void* ptr, int64_t bytes);

/* ---------------------------------------------------------------------------------------------------------------
 * Fused actor-critic MLP on tensor cores (TF32 operands, fp32 accumulate).  Reference: lib/network/mlp.py:4-39 (three
 * Linear+ELU layers), the `mu` and `value_head` Linear layers of lib/model/a2c_continuous_logstd_model.py:52-68,159-168,
 * and the input normalisation RunningMeanStd.forward (lib/core/running_mean_std.py:76-80).  All pointers are DEVICE
 * pointers into the caller's parameter tensors (torch Linear layout: weight [out,in] row-major, bias [out]).
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct AgxMlpParams {
    int32_t in_dim;          /* observation width (18 hovering/balloon, 48 tracking) */
    int32_t in_pad;          /* in_dim rounded up to a multiple of 16 */
    int32_t h1, h2, h3;      /* hidden widths (64,128,64): multiples of 32, at most 128 */
    int32_t actions_num;     /* 4 or 5 */
    const float *w1, *b1, *w2, *b2, *w3, *b3;    /* actor_mlp.layers.{0,1,2}.{weight,bias} */
    const float *w_mu, *b_mu;                    /* mu.{weight,bias}         [A,h3], [A] */
    const float *w_value, *b_value;              /* value_head.{weight,bias} [1,h3], [1] */
    const double *in_mean, *in_var;              /* running_mean_std.running_{mean,var} (float64) or NULL = no normalisation */
} AgxMlpParams;

/* where the parameter gradients go (torch layout, e.g. views of one flat gradient buffer) */
typedef struct AgxMlpGrads {
    float *gw1, *gb1, *gw2, *gb2, *gw3, *gb3, *gw_mu, *gb_mu, *gw_value, *gb_value;
} AgxMlpGrads;

/* obs [b,in_dim] → mu [b,A], value [b] (the normalised head output); when the four *_out pointers are non-NULL the
 * normalised input [b,in_pad] (zero padded) and the post-ELU activations [b,h1],[b,h2],[b,h3] are kept for the backward. */
int agx_mlp_forward(const AgxMlpParams* p, int64_t b, const float* obs, float* mu, float* value, float* xn_out,
                    float* h1_out, float* h2_out, float* h3_out, void* stream);

/* floats of scratch agx_mlp_backward needs (split-K partials of the weight and bias gradients) */
int64_t agx_mlp_workspace_floats(const AgxMlpParams* p);

/* Full backward of the network for d(loss)/d(mu) [b,A] and d(loss)/d(value) [b] given the kept activations:
 * (1) activation-gradient chain dz3, dz2, dz1 [b,h*] and the padded head gradient dout [b,16] (scratch outputs),
 * (2) dW_l = dz_l^T · a_{l-1} and the bias gradients, reduced deterministically and written into `g` (overwrite, not +=).
 * b must be a multiple of 8.  Three launches, no host sync. */
int agx_mlp_backward(const AgxMlpParams* p, const AgxMlpGrads* g, int64_t b, const float* grad_mu, const float* grad_value,
                     const float* xn, const float* h1, const float* h2, const float* h3, float* dz1, float* dz2, float* dz3,
                     float* dout, float* workspace, void* stream);

/* Training path on the 5th-generation tensor cores (tcgen05 + TMEM) for the shipped 64-128-64 network: forward, activation-gradient
 * chain and weight / bias gradients all run on tcgen05.mma.  Intermediates are FEATURE-MAJOR planes ([b/128][width][128] (feature-major, blocked by 128-row tile), b % 128 == 0), which
 * is what makes both operands of the weight gradient (it contracts over the batch axis) K-major:
 *   xt [in_pad, b] (plane in_dim = 1: the bias-gradient column; in_pad in {32,48,64,96} must exceed in_dim), h1t [64, b], h2t [128, b],
 *   h3t [64, b] written by agx_mlp_forward_train;  dz1t [64, b], dz2t [128, b], dz3t [64, b], doutt [16, b] scratch of the backward.
 * agx_mlp_backward_train overwrites the gradients in `g` (deterministic); workspace as agx_mlp_workspace_floats. */
int agx_mlp_train_supported(const AgxMlpParams* p);
int agx_mlp_forward_train(const AgxMlpParams* p, int64_t b, const float* obs, float* mu, float* value, float* xt, float* h1t,
                          float* h2t, float* h3t, void* stream);
int agx_mlp_backward_train(const AgxMlpParams* p, const AgxMlpGrads* g, int64_t b, const float* grad_mu, const float* grad_value,
                           const float* xt, const float* h1t, const float* h2t, const float* h3t, float* dz1t, float* dz2t,
                           float* dz3t, float* doutt, float* workspace, void* stream);
/* The PPO loss (agx_ppo_loss: a2c_continuous.py:299-369 after the network heads) folded into the first stage of the tcgen05 backward:
 * the thread that owns a minibatch row computes its clipped-surrogate / value / bound-loss gradients from mu, value and the stored
 * rollout quantities right where the backward needs d loss / d (mu | value), updates old_mu / old_sigma (PPODataset.update_mu_sigma) and
 * accumulates the loss statistics; the CTA drawing the last ticket reduces the per-CTA partials in CTA order (deterministic) into
 * stats [AGX_PPO_STATS] and grad_logstd [a].  One launch less per minibatch and no grad_mu / grad_value round trip; same arithmetic
 * per row as agx_ppo_loss, statistics identical up to the summation order.  `workspace` as agx_ppo_workspace_floats(). */
typedef struct AgxLossIO {
    const float *mu, *logstd, *value;              /* [b,a] / [a] / [b]: network outputs of agx_mlp_forward_train, the model's logstd */
    const float *actions, *old_neglogp, *adv, *returns; /* [b,a] / [b] / [b] / [b] minibatch slices of the rollout buffers */
    float *old_mu, *old_sigma;                     /* [b,a] in/out */
    float *grad_logstd, *stats, *workspace;        /* [a] / [AGX_PPO_STATS] out; scratch */
    int32_t a, _pad;                               /* actions_num: 4 | 5 */
} AgxLossIO;
int agx_sizeof_loss_io(void);
int agx_ppo_loss_backward_train(const AgxPpoHyper* hp, const AgxLossIO* lio, const AgxMlpParams* p, const AgxMlpGrads* g, int64_t b,
                                const float* xt, const float* h1t, const float* h2t, const float* h3t, float* dz1t, float* dz2t, float* dz3t,
                                float* doutt, float* workspace, void* stream);


/* ---- fused rollout step (reference lib/agent/a2c_base.py:651-711 play_steps; get_action_values :357-369; the model's sampling
 * branch a2c_continuous_logstd_model.py:181-193; preprocess_actions a2c_continuous.py:61-71) -------------------------------------
 * agx_policy_step = the tcgen05 MLP forward with the whole policy head fused into its epilogue: per env row
 *   mu, value_n  ← network;  sigma = exp(logstd);  action = mu + sigma * z  (z: explicit draws or Philox stream 6 keyed by
 *   (seed, env_offset + row, step_dev[0]) → 32-bit Box-Muller);  neglogp;  value = denormalised value_n;
 * and writes the rollout-buffer slices of this step (row strides in elements: the buffers are env-major [n, horizon, ...]),
 * the observation row, the done flag before the step, and env_actions = clamp(action, -1, 1) rescaled to [act_lo, act_hi].
 * One launch replaces ~12 torch kernels per step. */
typedef struct AgxPolicyIO {
    const float*  logstd;                 /* [A] */
    const double *value_mean, *value_var; /* value RunningMeanStd buffers (float64), or NULL: values stay normalised */
    float* actions;   int64_t ld_actions;
    float* mus;       int64_t ld_mus;
    float* sigmas;    int64_t ld_sigmas;
    float* neglogp;   int64_t ld_neglogp;
    float* values;    int64_t ld_values;
    float* obs_out;   int64_t ld_obs;     /* [n, in_dim] slice: copy of the observation the policy saw */
    uint8_t* dones_out; int64_t ld_dones; /* [n] slice <- dones_in */
    const uint8_t* dones_in;              /* [n] done flags after the previous step */
    float* env_actions;                   /* [n, A] dense: what env.step receives */
    const float *act_lo, *act_hi;         /* [A] action-space bounds, or NULL: env_actions = action (clip_actions: False) */
    const float* noise;                   /* [n, A] N(0,1) draws, or NULL → Philox */
    uint64_t seed;
    const uint64_t* step_dev;             /* device step counter (the env's): read, not advanced */
    int64_t env_offset;
} AgxPolicyIO;
int agx_sizeof_policy_io(void);
int agx_policy_step(const AgxMlpParams* p, const AgxPolicyIO* io, int64_t n, const float* obs, void* stream);

/* agx_rollout_post = everything play_steps does after env.step (a2c_base.py:667-695): shaped reward
 * r' = clamp((r + shift) * scale, lo, hi) [+ gamma * value * time_out when bootstrap], stored in the rollout slice; running
 * episode return / shaped return / length per env; on a done flag their values are added to ep_stats {sum return, sum shaped,
 * sum length, episodes} (float64 atomics — logging only) and reset; dones_state <- done flags. */
typedef struct AgxPostIO {
    const float* reward;  const uint8_t* reset_u8;  const uint8_t* timeout;  /* env outputs [n] */
    const float* values;  int64_t ld_values;                                  /* this step's denormalised values (rollout slice) */
    float* rewards_out;   int64_t ld_rewards;                                 /* rollout slice */
    float *cur_reward, *cur_shaped, *cur_length;                              /* [n] running episode accumulators */
    uint8_t* dones_state;                                                     /* [n] */
    double* ep_stats;                                                         /* [4] */
    float scale, shift, min_val, max_val, gamma;
    int32_t bootstrap;
} AgxPostIO;
int agx_sizeof_post_io(void);
int agx_rollout_post(const AgxPostIO* io, int64_t n, void* stream);

/* RunningMeanStd.update in two launches (lib/core/running_mean_std.py:45-60; the reference runs ~12 torch kernels per update):
 * agx_col_sums: sums [2, k] float64 = column sums and sums of squares of x [n, k] (row stride ld), deterministic; workspace of
 * agx_col_sums_workspace_doubles() float64, zero-filled once.  Between the two a multi-GPU run all-reduces `sums`.
 * agx_rms_merge: batch mean / UNBIASED variance over n_total rows from the sums, merged into mean / var / count (float64, in place). */
int64_t agx_col_sums_workspace_doubles(void);
int agx_col_sums(const float* x, int64_t n, int k, int64_t ld, double* sums, double* workspace, void* stream);
int agx_rms_merge(const double* sums, int k, double n_total, double* mean, double* var, double* count, void* stream);

/* ---- depth-image encoder (row f3) ------------------------------------------------------------------------------------
 * Replaces the forward of lib/network/cnn.py:3-33 (CNNFeatureExtractor: three stride-2 convolutions, each followed by
 * ReLU then BatchNorm2d, global average pool, Linear 64 -> feature_dim) in EVAL mode, with the per-pixel input
 * normalisation of lib/core/running_mean_std.py:62-81 fused into the image load.  All pointers are device memory and
 * hold the PyTorch parameter tensors as they are (OIHW convolution weights); s* / t* are the eval-mode BatchNorm affine
 * s = weight / sqrt(running_var + eps), t = bias - running_mean * s, prepared by the caller. */
typedef struct AgxCnnParams {
    const float *w1, *b1, *s1, *t1; /* features.0 [16,1,5,5], [16]; features.2 scale / shift [16] */
    const float *w2, *b2, *s2, *t2; /* features.3 [32,16,3,3], [32]; features.5 [32] */
    const float *w3, *b3, *s3, *t3; /* features.6 [64,32,3,3], [64]; features.8 [64] */
    const float *wfc, *bfc;         /* fc [feature_dim,64], [feature_dim] */
    int32_t feature_dim;            /* 1..64 */
    int32_t _pad;
} AgxCnnParams;
int agx_sizeof_cnn_params(void);

/* image [n,1,AGX_CAM_W,AGX_CAM_H] f32 -> features [n, ld_features] (first feature_dim columns written; ld_features lets
 * the caller write straight into a wider trunk-input row).  px_mean / px_rstd [AGX_CAM_W*AGX_CAM_H] (both or neither):
 * each pixel becomes clamp((x - mean) * rstd, +-5) before the first convolution.  One persistent launch, fp32 FMA
 * arithmetic throughout (no TF32), no activation leaves the SM. */
int agx_cnn_encode(const AgxCnnParams* p, int64_t n, const float* image, const float* px_mean, const float* px_rstd,
                   float* features, int64_t ld_features, void* stream);

/* ---- encoder layers on the tensor cores (row f3; csrc/agx_conv.cu) ----------------------------------------------------------------
 * Channels-last (NHWC) fp32 activations.  agx_conv2d_nhwc: one convolution layer as an implicit GEMM on tcgen05.mma — the im2col
 * operand is gathered from the input by cp.async straight into the tensor core's operand layout; 3xTF32 operand split (w_lo given)
 * for fp32-level results, or single-pass TF32 (w_lo NULL: what torch/cuDNN run by default).  Epilogue order:
 * + bias, + residual, activation, per-channel affine.  Replaces the cuDNN calls behind nn.Conv2d / nn.Linear of
 * lib/network/cnn.py:3-33 and lib/network/VAE.py:52-148. */
enum AgxAct { AGX_ACT_NONE = 0, AGX_ACT_RELU = 1, AGX_ACT_ELU = 2 };
typedef struct AgxConvParams {
    const float* x;  int32_t N, H, W, Cin;     /* input [N,H,W,Cin], Cin % 4 == 0 */
    const float *w_hi, *w_lo;                   /* weights [Cout][kh][kw][Cin] split as w = w_hi + w_lo with w_hi exactly tf32; w_lo may be NULL */
    const float* bias;                          /* [Cout] or NULL */
    const float *scale, *shift;                 /* [Cout] affine applied after the activation (eval-mode BatchNorm), or both NULL */
    const float* res; int32_t rH, rW, ry0, rx0, rsy, rsx; /* residual [N,rH,rW,Cout] read at (oy*rsy + ry0, ox*rsx + rx0), or NULL */
    float* y;        int32_t Ho, Wo, Cout;      /* output [N,Ho,Wo,Cout], Cout % 32 == 0 (and % 128 == 0 above 128) */
    int32_t kh, kw, sy, sx, py, px, act;        /* kernel, stride, zero padding, AgxAct; kh*kw*Cin % 8 == 0 */
    int32_t _pad;
} AgxConvParams;
int agx_sizeof_conv_params(void);
int agx_conv2d_nhwc(const AgxConvParams* p, void* stream);

/* first layer (one input channel): direct fp32 convolution, output NHWC with Cout = 16 | 32; optional fused input normalisation
 * clamp((x - px_mean) * px_rstd, +-5) per pixel (lib/core/running_mean_std.py:76-80); bias, activation, affine as above */
typedef struct AgxConvFirstParams {
    const float* x;  int32_t N, H, W, _p0;     /* input [N,H,W] */
    const float *w, *bias, *scale, *shift;      /* w [Cout][kh][kw] (torch OIHW with I = 1) */
    const float *px_mean, *px_rstd;             /* [H*W] each, or both NULL */
    float* y;        int32_t Ho, Wo, Cout;
    int32_t kh, kw, sy, sx, py, px, act;
} AgxConvFirstParams;
int agx_sizeof_conv_first_params(void);
int agx_conv2d_first(const AgxConvFirstParams* p, void* stream);
/* F.interpolate(x[n,1,H,W], (Ho,Wo), mode='bilinear', align_corners=False) */
int agx_resize_bilinear(const float* x, float* y, int64_t n, int H, int W, int Ho, int Wo, void* stream);
/* Train-mode BatchNorm2d over a channels-last activation x [rows, C] in place (lib/network/cnn.py:9-21 called in train mode, as the
 * reference's update pass does): `sums` [2, C] float64 from agx_col_sums(x, rows, C, C, ...) — between the two calls a multi-GPU run may
 * all-reduce them; y = (x - batch mean) / sqrt(biased batch var + eps) * gamma + beta; running_mean / running_var (or both NULL) move by
 * `momentum` towards the batch mean / UNBIASED batch variance.  The caller bumps num_batches_tracked. */
int agx_bn_train(float* x, int64_t rows, int C, const double* sums, const float* gamma, const float* beta, float eps, float momentum,
                 float* running_mean, float* running_var, void* stream);
/* out[n, :F] = Linear(mean over `pixels` of x[n, pixels, C]) — AdaptiveAvgPool2d((1,1)) + fc of the CNN (cnn.py:27-33) */
int agx_pool_fc(const float* x, int64_t n, int pixels, int C, const float* wfc, const float* bfc, int F, float* out, int64_t ld_out,
                void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AGX_H */

// agx_step_kernel.cuh — the fused env-step kernel template and its launchers; included by one translation unit per task
// (agx_step_<task>.cu) so the 5 tasks x 5 control modes compile in parallel.  See agx_step.cu for the C ABI.
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
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "agx.h"
#include "agx_math.cuh"

namespace agxk {

// ---- PTX wrappers: mbarrier + TMA bulk copies ---------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
                 "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* src_gmem, uint32_t bytes) {  // bytes % 16 == 0, src 16-B aligned
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() {
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

#ifdef AGX_TIMELINE
// Tuning builds only (scripts/timeline.py): per-CTA phase stamps {globaltimer ns, clock64} x 6 phases + smid.
static __device__ unsigned long long g_timeline[8192 * 16];
__device__ __forceinline__ void tl_stamp(int phase) {
    if (threadIdx.x == 0 && blockIdx.x < 8192) {
        unsigned long long t, c;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        asm volatile("mov.u64 %0, %%clock64;" : "=l"(c));
        g_timeline[blockIdx.x * 16 + phase * 2] = t;
        g_timeline[blockIdx.x * 16 + phase * 2 + 1] = c;
        if (phase == 0) {
            unsigned int smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            g_timeline[blockIdx.x * 16 + 14] = smid;
        }
    }
}
#define TL(p) tl_stamp(p)
#else
#define TL(p)
#endif

template <int NOBS>
struct ObsLayout {
    // dense rows only when the row stride in words is conflict-light (18 → 2-way); else pad to odd
    static constexpr bool kDense = (NOBS == 18);
    static constexpr int kStride = kDense ? NOBS : (NOBS | 1);
};

template <int TASK>
struct TaskTraits;
template <>
struct TaskTraits<AGX_TASK_HOVERING> { static constexpr int kObs = 18; };
template <>
struct TaskTraits<AGX_TASK_TRACKING> { static constexpr int kObs = 48; };
template <>
struct TaskTraits<AGX_TASK_BALLOON> { static constexpr int kObs = 18; };
template <>
struct TaskTraits<AGX_TASK_AVOID> { static constexpr int kObs = 16; };
template <>
struct TaskTraits<AGX_TASK_PLANNING> { static constexpr int kObs = 16; };

// ---- warp-cooperative reset sampling ---------------------------------------------------------------------
// Resets are rare per env (~2 % of env-steps under random actions) but common per warp (~50 % of warps hold at least one
// resetting lane), so a per-lane `if (reset) sample()` makes half of all warps walk the whole sampler — Philox blocks,
// sincos, quaternion — for one or two live lanes (measured: 43 % of the step time at 4 M envs).  Instead the warp packs
// its resetting lanes into items and spends 4 lanes on each: lane `sub` of an item draws Philox block `sub` (or copies
// the explicit draws) into shared memory, lane 0 of the item turns the uniforms into the new root-state row, written
// straight into the CTA's state tile; task state (aux) returns through the same scratch row.  Up to 8 items per pass.
struct WarpScratch {
    float u[8][AGX_RESET_DRAWS_MAX];  // per item: uniforms in, aux out
    uint8_t src[32];                  // item → lane
};

// Planning.reset_idx (planning.py:63-136) needs 124 draws per env (41 assets x {x, y, yaw} + the goal's y): one resetting env
// at a time, lane l draws Philox block l (draws 4l..4l+3) and places them in the env's asset row in global memory; lane 0
// then fixes up the goal ball and writes the drone's start pose.  Episodes are hundreds of steps long, so this is rare.
__device__ __forceinline__ void warp_reset_planning(const AgxStepIO& io, bool need, int which, uint64_t step, int64_t warp_env0,
                                                    float* s_state_warp, WarpScratch& ws, float* aux) {
    using namespace agx;
    unsigned mask = __ballot_sync(0xFFFFFFFFu, need);
    const int lane = threadIdx.x & 31;
    while (mask) {
        const int src = __ffs(mask) - 1;
        mask &= mask - 1;
        const int64_t env = warp_env0 + src;
        float* row = io.assets + env * (int64_t)AGX_ASSET_ROW;
        float u[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        if (lane < AGX_PLANNING_DRAWS / 4) {
            if (io.rand_reset) {
                const float* r = io.rand_reset + (env * 2 + which) * (int64_t)AGX_PLANNING_DRAWS + lane * 4;
#pragma unroll
                for (int j = 0; j < 4; ++j) u[j] = r[j];
            } else {
                PhiloxCtx ph;
                const uint64_t genv = (uint64_t)(io.env_offset + env);
                ph.k0 = (uint32_t)io.seed; ph.k1 = (uint32_t)(io.seed >> 32);
                ph.env_lo = (uint32_t)genv; ph.env_hi = (uint32_t)(genv >> 32);
                ph.step_lo = (uint32_t)step; ph.step_hi = (uint32_t)(step >> 32);
                const U4 w = philox_block(ph, (uint32_t)which, (uint32_t)lane);
                u[0] = u32_to_unit(w.x); u[1] = u32_to_unit(w.y); u[2] = u32_to_unit(w.z); u[3] = u32_to_unit(w.w);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (lane * 4 + j != AGX_PLANNING_DRAWS - 1) planning_place_draw(lane * 4 + j, u[j], row);
        }
        const float u_goal = __shfl_sync(0xFFFFFFFFu, u[3], AGX_PLANNING_DRAWS / 4 - 1);
        __syncwarp();  // the row is complete before the goal fix-up overwrites slots 0 and 41
        if (lane == 0) {
            float st[13], ax[AGX_AUX_MAX];
            planning_goal_fixup(u_goal, row, st, ax);
#pragma unroll
            for (int i = 0; i < 13; ++i) s_state_warp[src * 13 + i] = st[i];
#pragma unroll
            for (int i = 0; i < 6; ++i) ws.u[0][i] = ax[i];
        }
        __syncwarp();
        if (lane == src) {
#pragma unroll
            for (int i = 0; i < 6; ++i) aux[i] = ws.u[0][i];  // goal xyz, pre_root_positions = 0
        }
        __syncwarp();
    }
}

template <int TASK>
__device__ __forceinline__ void warp_reset(const AgxStepIO& io, bool need, int which, uint64_t step, int64_t warp_env0,
                                           float* s_state_warp, WarpScratch& ws, float* aux) {
    using namespace agx;
    if (TASK == AGX_TASK_PLANNING) { warp_reset_planning(io, need, which, step, warp_env0, s_state_warp, ws, aux); return; }
    constexpr int D = ResetDraws<TASK>::kD <= AGX_RESET_DRAWS_MAX ? ResetDraws<TASK>::kD : AGX_RESET_DRAWS_MAX;
    constexpr bool kAuxOut = (TASK == AGX_TASK_BALLOON || TASK == AGX_TASK_AVOID);  // the sampler also writes aux[0:6]
    const unsigned mask = __ballot_sync(0xFFFFFFFFu, need);
    if (mask == 0u) return;  // warp-uniform
    const int lane = threadIdx.x & 31;
    const int n_items = __popc(mask);
    const int my_item = __popc(mask & ((1u << lane) - 1u));
    if (need) ws.src[my_item] = (uint8_t)lane;
    __syncwarp();
    const int slot = lane >> 2, sub = lane & 3;
    for (int base = 0; base < n_items; base += 8) {
        const int item = base + slot;
        int src = 0;
        if (item < n_items) {
            src = ws.src[item];
            const int64_t env = warp_env0 + src;
            if (io.rand_reset) {
                const float* row = io.rand_reset + (env * 2 + which) * (int64_t)D;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (sub * 4 + j < D) ws.u[slot][sub * 4 + j] = row[sub * 4 + j];
            } else if (sub * 4 < D) {
                PhiloxCtx ph;
                const uint64_t genv = (uint64_t)(io.env_offset + env);
                ph.k0 = (uint32_t)io.seed; ph.k1 = (uint32_t)(io.seed >> 32);
                ph.env_lo = (uint32_t)genv; ph.env_hi = (uint32_t)(genv >> 32);
                ph.step_lo = (uint32_t)step; ph.step_hi = (uint32_t)(step >> 32);
                const U4 r = philox_block(ph, (uint32_t)which, (uint32_t)sub);
                ws.u[slot][sub * 4 + 0] = u32_to_unit(r.x);
                ws.u[slot][sub * 4 + 1] = u32_to_unit(r.y);
                ws.u[slot][sub * 4 + 2] = u32_to_unit(r.z);
                ws.u[slot][sub * 4 + 3] = u32_to_unit(r.w);
            }
        }
        __syncwarp();
        if (item < n_items && sub == 0) {
            float u[AGX_RESET_DRAWS_MAX], st[13], ax[AGX_AUX_MAX];
#pragma unroll
            for (int i = 0; i < D; ++i) u[i] = ws.u[slot][i];
#pragma unroll
            for (int i = 0; i < AGX_AUX_MAX; ++i) ax[i] = 0.0f;
            reset_sample<TASK>(u, st, ax);
#pragma unroll
            for (int i = 0; i < 13; ++i) s_state_warp[src * 13 + i] = st[i];
            if (kAuxOut) {
#pragma unroll
                for (int i = 0; i < 6; ++i) ws.u[slot][i] = ax[i];
            }
        }
        __syncwarp();
        if (kAuxOut && need && my_item >= base && my_item < base + 8) {
#pragma unroll
            for (int i = 0; i < 6; ++i) aux[i] = ws.u[my_item - base][i];  // balloon: ball xyz, pre_root_positions = 0; avoid: cube xyz, linvel
        }
        __syncwarp();
    }
}

// ---- the fused step kernel ----------------------------------------------------------------------------
#ifndef AGX_MIN_CTAS
#define AGX_MIN_CTAS 5  // <= 102 registers: 5 CTAs (20 warps) per SM; hovering/rate needs 96 unspilled, the widest modes spill 16 B
#endif
template <int TASK, int MODE, int BLOCK, int PHASE>
__global__ void __launch_bounds__(BLOCK, AGX_MIN_CTAS)
agx_step_kernel(const __grid_constant__ AgxParams P, const __grid_constant__ AgxStepIO io, const int64_t n,
                const int kflags) {  // bit0: TMA bulk staging, bits1-2: PDL trigger point (0 none, 1 start, 2 pre-store)
    using namespace agx;
    constexpr int A = (MODE == AGX_CTL_ATTI) ? 5 : 4;
    constexpr int NOBS = TaskTraits<TASK>::kObs;
    constexpr int K = (MODE == AGX_CTL_PROP) ? 0 : ((MODE == AGX_CTL_RATE || MODE == AGX_CTL_ATTI) ? 6 : 12);
    using OL = ObsLayout<NOBS>;
    constexpr bool kHasAux = (TASK == AGX_TASK_BALLOON || TASK == AGX_TASK_AVOID || TASK == AGX_TASK_PLANNING);
    constexpr bool kNoise = !(TASK == AGX_TASK_AVOID || TASK == AGX_TASK_PLANNING);  // the depth-camera tasks add no obs noise
    constexpr bool kPhys = (PHASE != AGX_PHASE_TASK), kTask = (PHASE != AGX_PHASE_PHYSICS);

    __shared__ __align__(128) float s_state[BLOCK * 13];
    __shared__ __align__(128) float s_obs[BLOCK * OL::kStride];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ WarpScratch s_ws[BLOCK / 32];
    __shared__ unsigned long long s_step;

    const int tid = threadIdx.x;
    // Tiles: BLOCK consecutive envs per CTA, or — single-wave grids (kflags bit 3) — the env axis cut into gridDim.x nearly equal
    // tiles whose boundaries are multiples of 4 envs (16-byte aligned spans), with gridDim.x a multiple of the SM count: every SM
    // then carries the same number of envs (65 536 envs: 592 tiles of 108-112 instead of 512 of 128 spread 4/3 over the SMs).
    int64_t tile0, tile_end;
    if (kflags & 8) {
        tile0 = (((int64_t)blockIdx.x * n) / gridDim.x) & ~(int64_t)3;
        tile_end = (blockIdx.x + 1 == gridDim.x) ? n : ((((int64_t)(blockIdx.x + 1) * n) / gridDim.x) & ~(int64_t)3);
    } else {
        tile0 = (int64_t)blockIdx.x * BLOCK;
        tile_end = (n - tile0) < (int64_t)BLOCK ? n : tile0 + BLOCK;
    }
    const int tile_n = (int)(tile_end - tile0);
    const bool bulk = (kflags & 1) && tile_n > 0 && (tile_n & 3) == 0;  // CTA-uniform: spans are 16-byte multiples
    const int pdl_mode = (kflags >> 1) & 3;
    // TASK phase only — "observe" calls (agx_observe): recompute observations and / or reward from the CURRENT buffers without
    // stepping: no end-of-step reset, no progress / time-out / state / task-state writes, the step counter is not advanced
    const bool observe = kTask && !kPhys && (kflags & 16), obs_on = !observe || (kflags & 32), rew_on = !observe || (kflags & 64);
    const int64_t env = tile0 + tid;
    const bool active = tid < tile_n;

    TL(0);
    EnvRegs e;
    float z[AGX_NOISE_DRAWS];
    uint64_t step = io.step;
    unsigned long long ticket = 0;

    // ---- stage the state tile into shared memory + coalesced per-env loads (which overlap the bulk copy in flight)
    auto load_inputs = [&]() {
        if (bulk) {
            if (tid == 0) {
                mbar_init(&s_bar, 1);
                mbar_expect_tx(&s_bar, tile_n * 13 * 4);
                bulk_g2s(s_state, io.state + tile0 * 13, tile_n * 13 * 4, &s_bar);
            }
        } else {
            const float* src = io.state + tile0 * 13;
            for (int i = tid; i < tile_n * 13; i += BLOCK) s_state[i] = src[i];
        }
        if (active) {
            const float* act_in = kPhys ? io.action : io.actions_out;  // TASK phase: the actions the PHYSICS phase shaped
            if (A == 4) {
                const float4 a4 = reinterpret_cast<const float4*>(act_in)[env];
                const float4 p4 = reinterpret_cast<const float4*>(io.prev_action)[env];
                e.a[0] = a4.x; e.a[1] = a4.y; e.a[2] = a4.z; e.a[3] = a4.w; e.a[4] = 0.0f;
                e.pa[0] = p4.x; e.pa[1] = p4.y; e.pa[2] = p4.z; e.pa[3] = p4.w; e.pa[4] = 0.0f;
            } else {
#pragma unroll
                for (int i = 0; i < 5; ++i) { e.a[i] = act_in[env * 5 + i]; e.pa[i] = io.prev_action[env * 5 + i]; }
            }
#pragma unroll
            for (int k = 0; k < AGX_CTRL_STATE_MAX; ++k) e.cs[k] = (kPhys && k < K) ? io.ctrl_state[(int64_t)k * n + env] : 0.0f;
            e.progress = io.progress[env];
            if (!kPhys && io.cmd) {  // TASK phase: the effort term reads the rotor commands the physics phase (or the last step) exported
                const float4 c4 = reinterpret_cast<const float4*>(io.cmd)[env];
                e.cmd[0] = c4.x; e.cmd[1] = c4.y; e.cmd[2] = c4.z; e.cmd[3] = c4.w;
            }
            e.pending = kPhys ? (io.reset[env] != 0) : 0;
            e.reset = 0;
            if (kHasAux) {
                const float4 x0 = reinterpret_cast<const float4*>(io.aux)[env * 2], x1 = reinterpret_cast<const float4*>(io.aux)[env * 2 + 1];
                e.aux[0] = x0.x; e.aux[1] = x0.y; e.aux[2] = x0.z; e.aux[3] = x0.w;
                e.aux[4] = x1.x; e.aux[5] = x1.y; e.aux[6] = x1.z; e.aux[7] = x1.w;
            }
        }
    };
    // Philox step index: device counter (graph replay) or launch argument.  ONE thread reads the counter (acquire, pairs with the
    // release store of the previous launch's bump) and takes the CTA's ticket AFTER its read (data dependency through `zero`);
    // the value reaches the other warps through shared memory, so every thread of the CTA uses the same index even when the
    // CTA holding the last ticket — every other CTA has read by then — bumps the counter while this CTA's late warps arrive.
    auto take_step = [&]() {
        if (io.step_dev) {
            if (tid == 0) {
                unsigned long long s0;
                asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(s0) : "l"(io.step_dev) : "memory");
                if (!observe) {
                    unsigned int zero;
                    asm volatile("and.b32 %0, %1, 0;" : "=r"(zero) : "r"((unsigned int)s0));
                    ticket = atomicAdd(reinterpret_cast<unsigned long long*>(io.step_dev + 1), 1ULL + zero);
                }
                s_step = s0;
            }
            __syncthreads();
            step = s_step;
        }
    };
    auto bump_step = [&]() {
        if (io.step_dev && !observe && tid == 0 && ticket == (unsigned long long)gridDim.x - 1ULL) {
            io.step_dev[1] = 0;
            asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(io.step_dev), "l"((unsigned long long)(step + 1)) : "memory");
            __threadfence();
        }
    };
    // Observation noise: needs only (seed, env id, step), not the env state.
    auto make_noise = [&]() {
        if (active) {
            RandSrc rnd;
            rnd.reset_row = nullptr;  // reset draws are taken by warp_reset
            rnd.noise_row = io.rand_noise ? io.rand_noise + env * (int64_t)AGX_NOISE_DRAWS : nullptr;
            const uint64_t genv = (uint64_t)(io.env_offset + env);
            rnd.ph.k0 = (uint32_t)io.seed; rnd.ph.k1 = (uint32_t)(io.seed >> 32);
            rnd.ph.env_lo = (uint32_t)genv; rnd.ph.env_hi = (uint32_t)(genv >> 32);
            rnd.ph.step_lo = (uint32_t)step; rnd.ph.step_hi = (uint32_t)(step >> 32);
            if (kNoise && kTask) scaled_noise(P, rnd, z);
        }
    };

    if (pdl_mode == 3) {
        // Programmatic dependent launch, noise-first: this grid may start while the previous kernel of the stream is
        // still running.  The step counter, the ticket and the noise touch nothing that kernel writes (its own counter
        // bump precedes its launch_dependents), so a third of the step's instructions run under its tail; every other
        // global access waits for its completion + flush.
        // L2 is the GPU's point of coherence, so prefetching this tile's inputs into it is a pure hint whatever the
        // running predecessor still writes: the DRAM reads of step t+1 overlap the compute phase of step t, and the
        // loads after the wait hit L2.
        if (tile_n > 0 && (tile_n & 3) == 0 && tid < 5 + K) {
            if (tid == 0) prefetch_l2(io.state + tile0 * 13, tile_n * 13 * 4);
            else if (tid == 1) prefetch_l2(io.action + tile0 * A, tile_n * A * 4);
            else if (tid == 2) prefetch_l2(io.prev_action + tile0 * A, tile_n * A * 4);
            else if (tid == 3) prefetch_l2(io.progress + tile0, tile_n * 8);
            else if (tid == 4) prefetch_l2(io.reset + tile0, tile_n * 8);
            else if ((n & 3) == 0) prefetch_l2(io.ctrl_state + (int64_t)(tid - 5) * n + tile0, tile_n * 4);  // plane rows 16-B aligned
        }
        take_step();
        if (!io.rand_noise) make_noise();
        bump_step();
        if (tid == 0) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // after this CTA's (possible) bump
        TL(1);
        asm volatile("griddepcontrol.wait;" ::: "memory");
        TL(2);
        load_inputs();
        if (io.rand_noise) make_noise();  // explicit draws may come from the previous kernel
    } else {
        // (griddepcontrol.* are no-ops when launched without the attribute.)
        asm volatile("griddepcontrol.wait;" ::: "memory");
        TL(1);
        if (pdl_mode == 1) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        load_inputs();
        take_step();
        make_noise();  // evaluated while the loads above fly
        TL(2);
        if (active) {
            // keep the consumers of the loads below this point: the compiler would otherwise hoist the first use of the
            // action registers above the noise code and park every warp on the load latency before doing useful work
#pragma unroll
            for (int i = 0; i < AGX_MAX_ACTIONS; ++i) asm volatile("" : "+f"(e.a[i]), "+f"(e.pa[i]));
        }
    }

    if (bulk) {
        __syncthreads();  // barrier init visible to all waiters
        mbar_wait(&s_bar, 0);
    } else {
        __syncthreads();
    }

    TL(3);
    // ---- pre_physics_step reset of the envs flagged last step (hovering.py:209-211, quirk Q1): new rows land in the tile
    const int warp = tid >> 5;
    const int64_t warp_env0 = tile0 + warp * 32;
    if (kPhys) warp_reset<TASK>(io, active && e.pending, 0, step, warp_env0, s_state + warp * 32 * 13, s_ws[warp], e.aux);

    if (active) {
#pragma unroll
        for (int i = 0; i < 13; ++i) e.s[i] = s_state[tid * 13 + i];
        if (kPhys && e.pending) reset_apply(P, e);

        SceneRef sc;
        sc.assets_row = (TASK == AGX_TASK_PLANNING) ? io.assets + env * (int64_t)AGX_ASSET_ROW : nullptr;
        sc.trees = io.trees;
        float R[9];
        if (kPhys) {
            env_phys<TASK, MODE>(P, e, sc, R);
        } else {
            Q4 q; q.x = e.s[3]; q.y = e.s[4]; q.z = e.s[5]; q.w = e.s[6];
            quat_to_matrix(q, R);
        }
        if (kTask) env_task<TASK, MODE>(P, z, e, R, &s_obs[tid * OL::kStride]);
        if (pdl_mode == 2) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

#pragma unroll
        for (int i = 0; i < 13; ++i) s_state[tid * 13 + i] = e.s[i];
    }
    __syncwarp();
    // ---- end-of-step reset_idx (hovering.py:300-302): fresh rows overwrite the tile, reset_buf stays 1, progress 0
    if (kTask && !observe) warp_reset<TASK>(io, active && e.reset, 1, step, warp_env0, s_state + warp * 32 * 13, s_ws[warp], e.aux);

    if (active && observe) {  // compute_observations() / compute_reward() as stand-alone calls (hovering.py:337-459)
        if (rew_on) {  // rew_buf, reset_buf overwritten, item_reward_info, pre_actions = actions.clone()
            if (A == 4) reinterpret_cast<float4*>(io.prev_action)[env] = make_float4(e.pa[0], e.pa[1], e.pa[2], e.pa[3]);
            else {
#pragma unroll
                for (int i = 0; i < 5; ++i) io.prev_action[env * 5 + i] = e.pa[i];
            }
            io.reset[env] = (int64_t)e.reset;
            if (io.reset_u8) io.reset_u8[env] = (uint8_t)e.reset;
            io.reward[env] = e.rew;
            if (io.reward_terms) {
                constexpr int NT = (TASK == AGX_TASK_PLANNING) ? 11 : 9;
#pragma unroll
                for (int k = 0; k < NT; ++k) io.reward_terms[(int64_t)k * n + env] = e.terms[k];
            }
            if (kHasAux) {  // the Customized family's compute_reward also sets pre_root_positions = root_positions.clone() (balloon.py:153)
                reinterpret_cast<float4*>(io.aux)[env * 2] = make_float4(e.aux[0], e.aux[1], e.aux[2], e.aux[3]);
                reinterpret_cast<float4*>(io.aux)[env * 2 + 1] = make_float4(e.aux[4], e.aux[5], e.aux[6], e.aux[7]);
            }
        }
    } else if (active) {
        if (kTask) {
            if (e.reset) reset_apply(P, e);
            env_finish(P, e);
        }

        // ---- coalesced per-env stores
        if (A == 4) {
            if (kPhys) reinterpret_cast<float4*>(io.actions_out)[env] = make_float4(e.a[0], e.a[1], e.a[2], e.a[3]);
            reinterpret_cast<float4*>(io.prev_action)[env] = make_float4(e.pa[0], e.pa[1], e.pa[2], e.pa[3]);
        } else {
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                if (kPhys) io.actions_out[env * 5 + i] = e.a[i];
                io.prev_action[env * 5 + i] = e.pa[i];
            }
        }
        if (kPhys) {
            if ((P.flags & AGX_FLAG_MUTATE_ACTIONS) && (MODE == AGX_CTL_RATE || MODE == AGX_CTL_ATTI))
                io.action[env * A + (A - 1)] = e.a_last_remap;
#pragma unroll
            for (int k = 0; k < K; ++k) io.ctrl_state[(int64_t)k * n + env] = e.cs[k];
            if (io.cmd) reinterpret_cast<float4*>(io.cmd)[env] = make_float4(e.cmd[0], e.cmd[1], e.cmd[2], e.cmd[3]);
        }
        if (kHasAux) {
            reinterpret_cast<float4*>(io.aux)[env * 2] = make_float4(e.aux[0], e.aux[1], e.aux[2], e.aux[3]);
            reinterpret_cast<float4*>(io.aux)[env * 2 + 1] = make_float4(e.aux[4], e.aux[5], e.aux[6], e.aux[7]);
        }
        io.progress[env] = e.progress;
        if (kTask) {
            io.reset[env] = (int64_t)e.reset;
            if (io.reset_u8) io.reset_u8[env] = (uint8_t)e.reset;
            io.timeout[env] = (uint8_t)e.timeout;
            io.reward[env] = e.rew;
            if (io.reward_terms) {
                constexpr int NT = (TASK == AGX_TASK_PLANNING) ? 11 : 9;
#pragma unroll
                for (int k = 0; k < NT; ++k) io.reward_terms[(int64_t)k * n + env] = e.terms[k];
            }
        }
    }

    // ---- tiles leave shared memory
    TL(4);
    if (bulk) {
        fence_async_smem();  // generic-proxy smem writes → visible to the async (TMA) proxy
        __syncthreads();
        if (tid == 0) {
            if (!observe) bulk_s2g(io.state + tile0 * 13, s_state, tile_n * 13 * 4);
            if (OL::kDense && kTask && obs_on) bulk_s2g(io.obs + tile0 * NOBS, s_obs, tile_n * NOBS * 4);
            bulk_commit();
        }
    } else {
        __syncthreads();
        float* dst = io.state + tile0 * 13;
        if (!observe)
            for (int i = tid; i < tile_n * 13; i += BLOCK) dst[i] = s_state[i];
    }
    if (kTask && obs_on && !(bulk && OL::kDense)) {
        float* dst = io.obs + tile0 * NOBS;
        if (OL::kDense) {
            for (int i = tid; i < tile_n * NOBS; i += BLOCK) dst[i] = s_obs[i];
        } else if ((tile_n * NOBS) % 4 == 0) {  // padded rows → dense float4 stores
            for (int i4 = tid; i4 < tile_n * NOBS / 4; i4 += BLOCK) {
                float4 v;
                const int i = i4 * 4;
                v.x = s_obs[i + i / NOBS];
                v.y = s_obs[(i + 1) + (i + 1) / NOBS];
                v.z = s_obs[(i + 2) + (i + 2) / NOBS];
                v.w = s_obs[(i + 3) + (i + 3) / NOBS];
                reinterpret_cast<float4*>(dst)[i4] = v;
            }
        } else {
            for (int i = tid; i < tile_n * NOBS; i += BLOCK) dst[i] = s_obs[i + i / NOBS];
        }
    }
    if (bulk && tid == 0) bulk_wait_read0();  // smem must stay alive until the bulk stores have read it
    if (pdl_mode != 3) bump_step();
    TL(5);
}

// ---- standalone reset_idx kernel -------------------------------------------------------------------------
template <int TASK>
__global__ void agx_reset_idx_kernel(const __grid_constant__ AgxParams P, int64_t n, int64_t m,
                                     const int64_t* __restrict__ env_ids, float* state, float* prev_action,
                                     float* ctrl_state, int64_t* progress, int64_t* reset, float* aux, float* assets,
                                     const float* __restrict__ rand, uint64_t seed, uint64_t step,
                                     int64_t env_offset) {
    using namespace agx;
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const int64_t env = env_ids[j];
    if (env < 0 || env >= n) return;
    PhiloxCtx ph;
    const uint64_t genv = (uint64_t)(env_offset + env);
    ph.k0 = (uint32_t)seed; ph.k1 = (uint32_t)(seed >> 32);
    ph.env_lo = (uint32_t)genv; ph.env_hi = (uint32_t)(genv >> 32);
    ph.step_lo = (uint32_t)step; ph.step_hi = (uint32_t)(step >> 32);
    float s[13];
    if (TASK == AGX_TASK_PLANNING) {  // stream 3: standalone reset_idx
        RandSrc r;
        r.reset_row = rand ? rand + (j - 3) * (int64_t)AGX_PLANNING_DRAWS : nullptr;  // reset_planning_serial indexes [which = 3]
        r.noise_row = nullptr;
        r.ph = ph;
        float ax[AGX_AUX_MAX];
        for (int i = 0; i < AGX_AUX_MAX; ++i) ax[i] = aux[env * AGX_AUX_MAX + i];
        reset_planning_serial(r, 3, s, ax, assets + env * (int64_t)AGX_ASSET_ROW);
        for (int i = 0; i < AGX_AUX_MAX; ++i) aux[env * AGX_AUX_MAX + i] = ax[i];
        for (int i = 0; i < 13; ++i) state[env * 13 + i] = s[i];
        for (int i = 0; i < P.num_actions; ++i) prev_action[env * P.num_actions + i] = 0.0f;
        if ((P.flags & AGX_FLAG_CTRL_RESET) && ctrl_state)
            for (int k = 0; k < P.ctrl_state_dim; ++k) ctrl_state[(int64_t)k * n + env] = 0.0f;
        progress[env] = 0;
        reset[env] = 1;
        return;
    }
    float u[AGX_RESET_DRAWS_MAX];
    if (rand) {
        for (int i = 0; i < P.reset_draws; ++i) u[i] = rand[j * P.reset_draws + i];
    } else {
        philox_uniforms(ph, 3u, P.reset_draws, u);  // stream 3: standalone reset_idx
    }
    if (TASK == AGX_TASK_BALLOON || TASK == AGX_TASK_AVOID) {
        float ax[AGX_AUX_MAX];
        for (int i = 0; i < AGX_AUX_MAX; ++i) ax[i] = aux[env * AGX_AUX_MAX + i];
        reset_sample<TASK>(u, s, ax);
        for (int i = 0; i < AGX_AUX_MAX; ++i) aux[env * AGX_AUX_MAX + i] = ax[i];
    } else {
        reset_sample<TASK>(u, s, nullptr);
    }
    for (int i = 0; i < 13; ++i) state[env * 13 + i] = s[i];
    for (int i = 0; i < P.num_actions; ++i) prev_action[env * P.num_actions + i] = 0.0f;
    if ((P.flags & AGX_FLAG_CTRL_RESET) && ctrl_state)
        for (int k = 0; k < P.ctrl_state_dim; ++k) ctrl_state[(int64_t)k * n + env] = 0.0f;
    progress[env] = 0;
    reset[env] = 1;
}

// process-wide tuning knobs (agx_set_option), defined in agx_step.cu
extern int g_block, g_use_bulk, g_pdl, g_sm_count, g_balanced;
int sm_count();
int fail(int code, const char* fmt, const char* detail = "");

template <typename Kernel>
cudaError_t launch_ex(Kernel k, unsigned grid, unsigned block, cudaStream_t st, const AgxParams& P,
                      const AgxStepIO& io, int64_t n, int extra_flags = 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    int pdl = g_pdl;
    const int sms = sm_count();
    // balanced tiles for grids that fit one wave (5 CTAs of 128 threads per SM are resident at once): the smallest multiple of the SM
    // count that keeps every tile within the block (n / G + 4 <= 128)
    int balanced = 0;
    if (g_balanced && block == 128 && n >= (int64_t)sms * 64) {
        const int64_t need = (n + 123) / 124, g = (need + sms - 1) / sms * sms;
        if (g <= (int64_t)sms * 5) { grid = (unsigned)g; balanced = 1; }
    }
    cfg.gridDim = dim3(grid);
    if (pdl < 0) pdl = ((uint64_t)grid * block <= (uint64_t)sms * 512u) ? 3 : 0;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    if (extra_flags) pdl = 0;  // observe calls are plain launches
    attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    const int kflags = (g_use_bulk ? 1 : 0) | ((pdl & 3) << 1) | (balanced ? 8 : 0) | extra_flags;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k, P, io, n, kflags);
}

template <int TASK, int MODE>
int launch_step(const AgxParams& P, int64_t n, const AgxStepIO& io, cudaStream_t st) {
    if (n == 0) return AGX_OK;
    cudaError_t err;
    constexpr bool kImageTask = (TASK == AGX_TASK_AVOID || TASK == AGX_TASK_PLANNING);
    if (kImageTask && io.phase == AGX_PHASE_PHYSICS) {
        err = launch_ex(agx_step_kernel<TASK, MODE, 128, kImageTask ? AGX_PHASE_PHYSICS : 0>, (unsigned)((n + 127) / 128), 128, st, P, io, n);
    } else if (kImageTask && io.phase == AGX_PHASE_TASK) {
        err = launch_ex(agx_step_kernel<TASK, MODE, 128, kImageTask ? AGX_PHASE_TASK : 0>, (unsigned)((n + 127) / 128), 128, st, P, io, n);
    } else if (g_block == 64 && !kImageTask) {
        err = launch_ex(agx_step_kernel<TASK, MODE, kImageTask ? 128 : 64, 0>, (unsigned)((n + 63) / 64), 64, st, P, io, n);
    } else {
        err = launch_ex(agx_step_kernel<TASK, MODE, 128, 0>, (unsigned)((n + 127) / 128), 128, st, P, io, n);
    }
    if (err == cudaSuccess) err = cudaGetLastError();
    if (err != cudaSuccess) return fail(AGX_ERR_CUDA, "agx_step launch: %s", cudaGetErrorString(err));
    return AGX_OK;
}

// agx_observe: the TASK phase of the step kernel in "observe" mode (what: 1 observations, 2 reward side, 3 both)
template <int TASK, int MODE>
int launch_observe(const AgxParams& P, int64_t n, const AgxStepIO& io, cudaStream_t st, int what) {
    if (n == 0) return AGX_OK;
    cudaError_t err = launch_ex(agx_step_kernel<TASK, MODE, 128, AGX_PHASE_TASK>, (unsigned)((n + 127) / 128), 128, st, P, io, n,
                                16 | ((what & 1) ? 32 : 0) | ((what & 2) ? 64 : 0));
    if (err == cudaSuccess) err = cudaGetLastError();
    if (err != cudaSuccess) return fail(AGX_ERR_CUDA, "agx_observe launch: %s", cudaGetErrorString(err));
    return AGX_OK;
}
template <int TASK>
int dispatch_observe(const AgxParams& P, int64_t n, const AgxStepIO& io, cudaStream_t st, int what) {
    constexpr bool kImageTask = (TASK == AGX_TASK_AVOID || TASK == AGX_TASK_PLANNING);
    switch (P.ctl_mode) {
        case AGX_CTL_POS: return launch_observe<TASK, AGX_CTL_POS>(P, n, io, st, what);
        case AGX_CTL_VEL: return launch_observe<TASK, AGX_CTL_VEL>(P, n, io, st, what);
        case AGX_CTL_ATTI:
            if constexpr (kImageTask) return fail(AGX_ERR_UNSUPPORTED, "agx_observe: avoid/planning have no atti mode%s");
            else return launch_observe<TASK, AGX_CTL_ATTI>(P, n, io, st, what);
        case AGX_CTL_RATE: return launch_observe<TASK, AGX_CTL_RATE>(P, n, io, st, what);
        case AGX_CTL_PROP: return launch_observe<TASK, AGX_CTL_PROP>(P, n, io, st, what);
        default: return fail(AGX_ERR_ARG, "agx_observe: unknown ctl_mode%s");
    }
}
template <int TASK>
int agx_observe_task(const AgxParams& P, int64_t n, const AgxStepIO& io, cudaStream_t st, int what) { return dispatch_observe<TASK>(P, n, io, st, what); }

template <int TASK>
int dispatch_mode(const AgxParams& P, int64_t n, const AgxStepIO& io, cudaStream_t st) {
    constexpr bool kImageTask = (TASK == AGX_TASK_AVOID || TASK == AGX_TASK_PLANNING);
    switch (P.ctl_mode) {
        case AGX_CTL_POS: return launch_step<TASK, AGX_CTL_POS>(P, n, io, st);
        case AGX_CTL_VEL: return launch_step<TASK, AGX_CTL_VEL>(P, n, io, st);
        case AGX_CTL_ATTI:
            if constexpr (kImageTask) return fail(AGX_ERR_UNSUPPORTED, "agx_step: avoid/planning have no atti mode (obs[12:16] holds the 4 actions)%s");
            else return launch_step<TASK, AGX_CTL_ATTI>(P, n, io, st);
        case AGX_CTL_RATE: return launch_step<TASK, AGX_CTL_RATE>(P, n, io, st);
        case AGX_CTL_PROP: return launch_step<TASK, AGX_CTL_PROP>(P, n, io, st);
        default: return fail(AGX_ERR_ARG, "agx_step: unknown ctl_mode%s");
    }
}

// explicit per-task entry point, instantiated in agx_step_<task>.cu
template <int TASK>
int agx_dispatch_task(const AgxParams& P, int64_t n, const AgxStepIO& io, cudaStream_t st) { return dispatch_mode<TASK>(P, n, io, st); }

}  // namespace agxk

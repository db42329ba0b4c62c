// agx_render.cu — depth camera of the Avoid / Planning tasks + the reference's depth post-processing, one CTA per env.
//
// Replaces IsaacGym's camera sensor (render_all_camera_sensors, customized.py:386-391 — a closed binary) with an analytic
// ray cast of the few primitives these scenes hold, and Customized.dump_images (customized.py:399-435: a Python loop over
// envs, 3 torch.normal / randint calls and a conv2d per env) with four passes over an image that never leaves shared
// memory: 212 x 120 floats = 101 760 B per CTA, 2 CTAs per SM.
//   pass 1  ray cast: ground plane + visible tree capsules + goal ball (planning) / thrown cube (avoid) → clip, / 4.5
//   pass 2  + N(0, 0.1), clamp to [0, max of pass 1]          (block max reduction between passes)
//   pass 3  x N(1, 0.3), clamp to [0, max of pass 2]
//   pass 4  5x5 correlation with a random kernel (zero padding) → global memory (float4, coalesced), block min → esdf_dist
// HBM traffic per env-render: 101 760 B written (+ 203 520 B of noise read in explicit-randomness mode), 52 + 32 (+ 656) B read.
#include <cuda_runtime.h>
#include <stdio.h>

#include "agx.h"
#include "agx_math.cuh"

int agx_internal_fail(int code, const char* msg);

namespace {

using namespace agx;

constexpr int kPix = AGX_CAM_W * AGX_CAM_H;  // 25 440
constexpr int kThreads = 256;

__device__ __forceinline__ float block_reduce(float v, bool is_max, float* s_red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float other = __shfl_xor_sync(0xFFFFFFFFu, v, o);
        v = is_max ? fmaxf(v, other) : fminf(v, other);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();  // s_red reuse
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    float r = s_red[0];
#pragma unroll
    for (int w = 1; w < kThreads / 32; ++w) r = is_max ? fmaxf(r, s_red[w]) : fminf(r, s_red[w]);
    return r;
}

template <int TASK>
__global__ void __launch_bounds__(kThreads, 2) agx_render_kernel(const __grid_constant__ AgxRenderIO io, const int64_t n) {
    extern __shared__ __align__(16) float s_img[];  // [W][H]
    __shared__ Capsule s_caps[AGX_NUM_TREES];
    __shared__ int s_u0[AGX_NUM_TREES], s_u1[AGX_NUM_TREES];  // image-column band of each listed capsule
    __shared__ int s_ncaps;
    __shared__ float s_red[kThreads / 32];
    __shared__ float s_kern[28];

    const int64_t env = blockIdx.x;
    const int tid = threadIdx.x;
    const Camera cam = make_camera(io.state + env * 13);
    const float* aux = io.aux + env * AGX_AUX_MAX;
    const V3 obj = v3(aux[0], aux[1], aux[2]);  // avoid: cube centre; planning: goal ball centre

    PhiloxCtx ph;
    {
        const uint64_t genv = (uint64_t)(io.env_offset + env);
        ph.k0 = (uint32_t)io.seed; ph.k1 = (uint32_t)(io.seed >> 32);
        ph.env_lo = (uint32_t)genv; ph.env_hi = (uint32_t)(genv >> 32);
        const uint64_t step = io.step_dev ? *io.step_dev : io.step;
        ph.step_lo = (uint32_t)step; ph.step_hi = (uint32_t)(step >> 32);
    }

    if (tid == 0) s_ncaps = 0;
    if (tid < 28) {  // blur kernel: randint(0, 256) / 256 (customized.py:419)
        if (io.rand_kern) s_kern[tid] = tid < 25 ? io.rand_kern[env * 25 + tid] : 0.0f;
        else if ((tid & 3) == 0) {
            const U4 w = philox_block(ph, 5u, (uint32_t)(tid >> 2));
            s_kern[tid + 0] = (float)(w.x & 255u) * (1.0f / 256.0f);
            s_kern[tid + 1] = (float)(w.y & 255u) * (1.0f / 256.0f);
            s_kern[tid + 2] = (float)(w.z & 255u) * (1.0f / 256.0f);
            s_kern[tid + 3] = (float)(w.w & 255u) * (1.0f / 256.0f);
        }
    }
    __syncthreads();
    if (TASK == AGX_TASK_PLANNING && tid < AGX_NUM_TREES) {  // visible trees → compact list
        const float* row = io.assets + env * (int64_t)AGX_ASSET_ROW;
        const int j = tid + 1;
        const Capsule k = place_tree(io.trees + tid * 8, row[j], row[AGX_NUM_ASSETS + j], row[2 * AGX_NUM_ASSETS + j],
                                     row[3 * AGX_NUM_ASSETS + j]);
        int u0, u1;
        capsule_columns(cam, k, &u0, &u1);
        if (u0 <= u1) {
            const int slot = atomicAdd(&s_ncaps, 1);
            s_caps[slot] = k; s_u0[slot] = u0; s_u1[slot] = u1;
        }
    }
    __syncthreads();
    const int ncaps = s_ncaps;
    bool ball_near = false;  // CTA-uniform: the goal ball can only show up within the far plane's reach
    if (TASK == AGX_TASK_PLANNING) ball_near = norm(obj - cam.o) - kBallRadius < 7.5f;

    // ---- pass 1: ray cast (4 consecutive v per thread: one u, float4-aligned).  d(u,v) = R (1, dy(u), dz(v)) = A(u) + dz(v) R[:,2]
    // CTA-uniform screen band of the cube's bounding sphere (r = sqrt(3) x half extent): a parked / far cube cannot show up at
    // all, a flying one covers a few dozen columns and rows
    int cu0 = 1, cu1 = 0, cv0 = 1, cv1 = 0;
    if (TASK == AGX_TASK_AVOID) {
        const float rad = 1.7320508f * kCubeHalf;
        const V3 rel = obj - cam.o;
        if (norm(rel) - rad < 7.5f) {
            const float xc = dot(rel, v3(cam.R[0], cam.R[3], cam.R[6])), yc = dot(rel, v3(cam.R[1], cam.R[4], cam.R[7])),
                        zc = dot(rel, v3(cam.R[2], cam.R[5], cam.R[8]));
            if (xc + rad > 0.0f) {
                cu0 = 0; cu1 = AGX_CAM_W - 1; cv0 = 0; cv1 = AGX_CAM_H - 1;
                if (xc - rad > 0.05f) {  // entirely in front of the camera plane: tight band (x2 radius for the perspective stretch)
                    const float inv = kCamF / (xc - rad), uc = (float)(AGX_CAM_W / 2) - 0.5f - kCamF * yc / xc,
                                vc = (float)(AGX_CAM_H / 2) - 0.5f - kCamF * zc / xc, pad = 2.0f * rad * inv + fabsf(yc) * rad * inv / xc + 2.0f,
                                padv = 2.0f * rad * inv + fabsf(zc) * rad * inv / xc + 2.0f;
                    cu0 = max(0, (int)floorf(uc - pad)); cu1 = min(AGX_CAM_W - 1, (int)ceilf(uc + pad));
                    cv0 = max(0, (int)floorf(vc - padv)); cv1 = min(AGX_CAM_H - 1, (int)ceilf(vc + padv));
                }
            }
        }
    }
    const V3 col2 = v3(cam.R[2], cam.R[5], cam.R[8]);
    float lmax = 0.0f;
    for (int i4 = tid * 4; i4 < kPix; i4 += kThreads * 4) {
        const int u = i4 / AGX_CAM_H, v0 = i4 - u * AGX_CAM_H;
        const float dy = fdiv((float)(AGX_CAM_W / 2) - (float)u - 0.5f, kCamF);
        const V3 A = v3(fmaf(cam.R[1], dy, cam.R[0]), fmaf(cam.R[4], dy, cam.R[3]), fmaf(cam.R[7], dy, cam.R[6]));
        float val[4];
        V3 d[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float dz = fdiv((float)(AGX_CAM_H / 2) - (float)(v0 + j) - 0.5f, kCamF);
            d[j] = v3(fmaf(col2.x, dz, A.x), fmaf(col2.y, dz, A.y), fmaf(col2.z, dz, A.z));
            val[j] = hit_ground(cam.o, d[j]);
        }
        if (TASK == AGX_TASK_PLANNING) {
            for (int c = 0; c < ncaps; ++c) {
                if (u < s_u0[c] || u > s_u1[c]) continue;  // this column cannot see capsule c
                const Capsule k = s_caps[c];
#pragma unroll
                for (int j = 0; j < 4; ++j) val[j] = fminf(val[j], hit_capsule(cam.o, d[j], k));
            }
            if (ball_near) {
#pragma unroll
                for (int j = 0; j < 4; ++j) val[j] = fminf(val[j], hit_sphere(cam.o, d[j], obj, kBallRadius));
            }
        } else if (u >= cu0 && u <= cu1 && v0 + 3 >= cv0 && v0 <= cv1) {
#pragma unroll
            for (int j = 0; j < 4; ++j) val[j] = fminf(val[j], hit_box(cam.o, d[j], obj, kCubeHalf));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            val[j] = normalize_depth(val[j]);
            lmax = fmaxf(lmax, val[j]);
        }
        *reinterpret_cast<float4*>(&s_img[i4]) = make_float4(val[0], val[1], val[2], val[3]);
    }
    const float m0 = block_reduce(lmax, true, s_red);

    // ---- pass 2: + N(0,0.1), clamp to [0, m0].  In Philox mode one word yields BOTH normals of a pixel: the multiplicative
    // factor 1 + 0.3 z1 is parked in this env's slice of the OUTPUT image (global memory the kernel owns until pass 4 overwrites
    // it; 2 CTAs x 148 SMs x 101 KB stay in L2) instead of re-deriving Philox + Box-Muller in pass 3.
    float* out = io.image + env * (int64_t)kPix;
    float pmax = 0.0f;
    for (int i4 = tid * 4; i4 < kPix; i4 += kThreads * 4) {
        const float4 x = *reinterpret_cast<float4*>(&s_img[i4]);
        float add[4];
        if (io.rand_add) {
            const float4 r = *reinterpret_cast<const float4*>(io.rand_add + env * (int64_t)kPix + i4);
            add[0] = r.x; add[1] = r.y; add[2] = r.z; add[3] = r.w;
        } else {
            const U4 w = philox_block(ph, 4u, (uint32_t)(i4 >> 2));
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
            float mul[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float z0, z1;
                box_muller16(ww[j], &z0, &z1);
                add[j] = 0.1f * z0;            // torch.normal(0, .1)
                mul[j] = fmaf(0.3f, z1, 1.0f);  // torch.normal(1, .3)
            }
            *reinterpret_cast<float4*>(out + i4) = make_float4(mul[0], mul[1], mul[2], mul[3]);
        }
        float y[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float t = y[j] + add[j];
            t = t < 0.0f ? 0.0f : t;  // torch.clamp(., 0, max)
            t = t > m0 ? m0 : t;
            y[j] = t;
            pmax = fmaxf(pmax, t);
        }
        *reinterpret_cast<float4*>(&s_img[i4]) = make_float4(y[0], y[1], y[2], y[3]);
    }
    const float m1 = block_reduce(pmax, true, s_red);

    // ---- pass 3: x N(1,0.3), clamp to [0, m1] (each thread reads back the factors it parked itself: no fence needed)
    const float* mul_src = io.rand_mul ? io.rand_mul + env * (int64_t)kPix : out;
    for (int i4 = tid * 4; i4 < kPix; i4 += kThreads * 4) {
        const float4 x = *reinterpret_cast<float4*>(&s_img[i4]);
        const float4 r = *reinterpret_cast<const float4*>(mul_src + i4);
        float y[4] = {x.x * r.x, x.y * r.y, x.z * r.z, x.w * r.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            y[j] = y[j] < 0.0f ? 0.0f : y[j];
            y[j] = y[j] > m1 ? m1 : y[j];
        }
        *reinterpret_cast<float4*>(&s_img[i4]) = make_float4(y[0], y[1], y[2], y[3]);
    }
    __syncthreads();  // pass 4 reads neighbours written by other threads, and overwrites the parked factors

    // ---- pass 4: 5x5 correlation, zero padding (F.conv2d(padding=2) on the [212,120] plane) → global, block min
    float kk[25];
#pragma unroll
    for (int i = 0; i < 25; ++i) kk[i] = s_kern[i];
    float lmin = kInf;
    for (int i4 = tid * 4; i4 < kPix; i4 += kThreads * 4) {
        const int u = i4 / AGX_CAM_H, v0 = i4 - u * AGX_CAM_H;
        float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const int uu = u + i - 2;
            if (uu < 0 || uu >= AGX_CAM_W) continue;
            // the 8 inputs v0-2 .. v0+5 of this row sit inside the three aligned float4 groups v0-4 .. v0+7
            float rowv[8];
            const float* rp = &s_img[uu * AGX_CAM_H + v0];
            const float4 mid = *reinterpret_cast<const float4*>(rp);
            float4 lo = make_float4(0.0f, 0.0f, 0.0f, 0.0f), hi = lo;
            if (v0 > 0) lo = *reinterpret_cast<const float4*>(rp - 4);
            if (v0 + 4 < AGX_CAM_H) hi = *reinterpret_cast<const float4*>(rp + 4);
            rowv[0] = lo.z; rowv[1] = lo.w; rowv[2] = mid.x; rowv[3] = mid.y; rowv[4] = mid.z; rowv[5] = mid.w;
            rowv[6] = hi.x; rowv[7] = hi.y;
#pragma unroll
            for (int j = 0; j < 5; ++j) {
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[q] = fmaf(kk[i * 5 + j], rowv[q + j], acc[q]);
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) lmin = fminf(lmin, acc[q]);
        *reinterpret_cast<float4*>(out + i4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    }
    const float mn = block_reduce(lmin, false, s_red);
    if (TASK == AGX_TASK_PLANNING && tid == 0) io.aux[env * AGX_AUX_MAX + 7] = mn;  // Planning: esdf_dist = min over the image (planning.py:162-163)
}

}  // namespace

extern "C" int agx_render_depth(const AgxParams* p, int64_t n, const AgxRenderIO* io, void* stream) {
    if (!p || !io) return agx_internal_fail(AGX_ERR_ARG, "agx_render_depth: null params/io");
    if (n < 0) return agx_internal_fail(AGX_ERR_ARG, "agx_render_depth: n < 0");
    if (!io->state || !io->aux || !io->image) return agx_internal_fail(AGX_ERR_ARG, "agx_render_depth: a required buffer is null");
    const void* ptrs[] = {io->image, io->rand_add, io->rand_mul, io->aux};
    for (const void* q : ptrs)
        if (q && (reinterpret_cast<uintptr_t>(q) & 15u)) return agx_internal_fail(AGX_ERR_ALIGN, "agx_render_depth: buffer not 16-byte aligned");
    if (n == 0) return AGX_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t smem = sizeof(float) * AGX_CAM_W * AGX_CAM_H;
    cudaError_t err;
    if (p->task == AGX_TASK_PLANNING) {
        if (!io->assets || !io->trees) return agx_internal_fail(AGX_ERR_ARG, "agx_render_depth: planning needs assets and trees");
        static bool attr_set = false;
        if (!attr_set) { cudaFuncSetAttribute(agx_render_kernel<AGX_TASK_PLANNING>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr_set = true; }
        agx_render_kernel<AGX_TASK_PLANNING><<<(unsigned)n, kThreads, smem, st>>>(*io, n);
    } else if (p->task == AGX_TASK_AVOID) {
        static bool attr_set = false;
        if (!attr_set) { cudaFuncSetAttribute(agx_render_kernel<AGX_TASK_AVOID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr_set = true; }
        agx_render_kernel<AGX_TASK_AVOID><<<(unsigned)n, kThreads, smem, st>>>(*io, n);
    } else {
        return agx_internal_fail(AGX_ERR_UNSUPPORTED, "agx_render_depth: only the avoid and planning tasks have a camera");
    }
    err = cudaGetLastError();
    if (err != cudaSuccess) return agx_internal_fail(AGX_ERR_CUDA, cudaGetErrorString(err));
    return AGX_OK;
}

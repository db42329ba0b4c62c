// agx_conv_tma.cu — the spatial convolution layers of the depth-image encoders (row f3; reference lib/network/cnn.py:3-33,
// lib/network/VAE.py:52-148) as a persistent, warp-specialised implicit GEMM: TMA im2col loads → 3xTF32 split → tcgen05.mma with
// TMEM accumulators → fused epilogue.  Replaces the cp.async gather of agx_conv.cu (one 16-byte copy + ~30 index instructions per
// operand chunk, a CTA-wide barrier per K-slab: 1.36 ms per 2048 images for the CNN's conv2 against a 0.19 ms HBM floor) wherever
// the geometry allows; the dense layers and anything outside the list below stay on that kernel.
//
//   M = 128 output pixels per tile (flattened over images; tiles strided over one persistent CTA per SM), N = Cout <= 128,
//   K-slab = one filter tap x KS channels (KS = 16 for Cin = 16, else 32): ONE cp.async.bulk.tensor im2col instruction brings the
//   slab's A operand [128 pixels][KS] (stride, padding and image / batch boundaries resolved by the TMA unit, zero fill outside),
//   one tiled load each the weight slab's hi and lo parts [Cout][KS]; all land in the 64- / 128-byte-swizzled K-major layout the
//   tensor core reads.  Roles (512 threads): warp 0 = TMA producer, warp 1 = MMA issuer (both run converged and issue under
//   elect.sync: a lone diverged lane makes the compiler wrap every uniform-datapath instruction in an election loop, which made the
//   first version issue-bound in that one thread at ~290 instructions per slab), warp 2 = TMEM allocator, warps 4-7 = epilogue
//   (TMEM lane quarters 0-3), warps 8-15 = splitter (A_lo = A - tf32(A), element-wise on the stage: the layout does not matter).
//   A shared-memory stage holds 1-4 slabs (one full → ready → empty mbarrier round trip per stage, not per slab); ring of 2-8
//   stages and two TMEM accumulator sets, so the loads of tile t+1 run under the MMAs of tile t and under the epilogue of tile t-1.
//   Epilogue: TMEM → registers (+ bias, residual, activation, affine) → a 128-byte-swizzled staging tile (bank-conflict-free
//   row writes) → cp.async.bulk.tensor stores of full 128-byte lines (rows past the end clipped by the TMA unit).
//   Precision: 3xTF32 as in agx_conv.cu (A_hi·B_hi + A_lo·B_hi + A_hi·B_lo; G accumulator chains for K > 512).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "agx.h"
#include "agx_tc.cuh"

int agx_internal_fail(int code, const char* msg);
extern int g_first_impl;  // agx_conv.cu

namespace {
using namespace tc;

constexpr int kThreads = 512;
constexpr int kMaxStages = 8;

struct ConvGeom {
    int64_t M_total;
    int32_t num_tiles, Nt, cblocks, nslabs, spp, fills, stages, G, tmem_cols;
    uint32_t a_bytes, b_bytes, stage_bytes, stg_bytes;
};

__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
// bounded wait: a pipeline that stops advancing (a transaction count that never completes) traps after ~4 s instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t done = 0, spins = 0;
    uint64_t t0 = 0;
    while (true) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(s32(b)), "r"(parity) : "memory");
        if (done) return;
        if ((++spins & 0x3FFu) == 0) {
            uint64_t t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t0 == 0) t0 = t;
            else if (t - t0 > 4000000000ull) __trap();
        }
    }
}
__device__ __forceinline__ bool elect_one() {  // one lane of the (converged) warp
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tma_im2col_4d(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int c, int w, int h, int n, uint16_t ow, uint16_t oh) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(tm)), "r"(s32(bar)), "r"(c), "r"(w), "r"(h), "r"(n), "h"(ow), "h"(oh)
                 : "memory");
}
__device__ __forceinline__ void tma_tile_2d(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(tm)), "r"(s32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_tile_3d(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(tm)), "r"(s32(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(src), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db),
                 "r"(idesc), "r"(acc)
                 : "memory");
}

extern __shared__ uint8_t t_smem[];

// KS = K per slab (channels of one filter tap): 16 → 64-byte rows / SWIZZLE_64B, 32 → 128-byte rows / SWIZZLE_128B
template <int KS, bool SPLIT>
__global__ void __launch_bounds__(kThreads, 1)
agx_conv2d_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo,
                      const __grid_constant__ CUtensorMap tmY, const __grid_constant__ AgxConvParams P, const __grid_constant__ ConvGeom Gm) {
    __shared__ __align__(8) uint64_t full[kMaxStages], ready[kMaxStages], empty[kMaxStages], accfull[2], accempty[2];
    __shared__ uint32_t tmem_base;
    __shared__ __align__(16) float s_bias[128], s_scale[128], s_shift[128];
    constexpr uint32_t kSwb = KS * 4;
    // K-major operand in the 64- / 128-byte swizzled canonical layout: rows of kSwb bytes, 8-row groups 8 * kSwb bytes apart.
    // High word of the shared-memory descriptor: stride byte offset, version 1 (bit 46), UMMA LayoutType SWIZZLE_128B = 2 / SWIZZLE_64B = 4
    // (bits 61-63); low word: start address >> 4 | leading byte offset 1 << 16 (unused by swizzled K-major layouts).
    constexpr uint32_t kDescHi = (((8u * kSwb) >> 4) & 0x3FFFu) | (1u << 14) | ((kSwb == 128 ? 2u : 4u) << 29);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t smem0 = (s32(t_smem) + 1023u) & ~1023u;  // swizzle patterns repeat every 1024 bytes of the shared-memory address
    const uint32_t ring0 = smem0 + Gm.stg_bytes;            // [staging tile of the epilogue | ring of stages]
    const int S = Gm.stages, Nt = Gm.Nt, nslabs = Gm.nslabs, spp = Gm.spp;
    const uint32_t offAlo = (uint32_t)spp * Gm.a_bytes, offBhi = (SPLIT ? 2u : 1u) * offAlo, offBlo = offBhi + (uint32_t)spp * Gm.b_bytes;

    if (tid == 0) {
        for (int i = 0; i < S; ++i) { mbar_init(&full[i], 1); mbar_init(&ready[i], 8); mbar_init(&empty[i], 1); }
        mbar_init(&accfull[0], 1); mbar_init(&accfull[1], 1); mbar_init(&accempty[0], 4); mbar_init(&accempty[1], 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmBhi)) : "memory");
        if (SPLIT) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmBlo)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmY)) : "memory");
    }
    if (tid < 128) {
        s_bias[tid] = (P.bias && tid < Nt) ? P.bias[tid] : 0.0f;
        s_scale[tid] = (P.scale && tid < Nt) ? P.scale[tid] : 1.0f;
        s_shift[tid] = (P.shift && tid < Nt) ? P.shift[tid] : 0.0f;
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_base)), "r"((uint32_t)Gm.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    const int HoWo = P.Ho * P.Wo;

    if (warp == 0) {  // ---- TMA producer (whole warp converged; one elected lane issues)
        int stage = 0;
        uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < Gm.num_tiles; tile += gridDim.x) {
            const int64_t pix0 = (int64_t)tile * kM;
            const int n0 = (int)(pix0 / HoWo), rem = (int)(pix0 - (int64_t)n0 * HoWo), oy0 = rem / P.Wo, ox0 = rem - oy0 * P.Wo;
            const int w0 = ox0 * P.sx - P.px, h0 = oy0 * P.sy - P.py;  // base pixel of the tile's first output pixel, input coordinates
            int cb = 0, ky = 0, kx = 0, s = 0;
            for (int f = 0; f < Gm.fills; ++f) {
                const int cnt = nslabs - s < spp ? nslabs - s : spp;
                mbar_wait(&empty[stage], ph ^ 1u);
                const bool leader = elect_one();
                const uint32_t base = ring0 + (uint32_t)stage * Gm.stage_bytes;
                if (leader) {  // the weight slabs of a fill arrive as one box [spp][Nt][KS] (slabs past the end zero-filled, bytes counted in full)
                    mbar_expect_tx(&full[stage], (uint32_t)cnt * Gm.a_bytes + (SPLIT ? 2u : 1u) * (uint32_t)spp * Gm.b_bytes);
                    tma_tile_3d(base + offBhi, &tmBhi, &full[stage], 0, 0, s);
                    if (SPLIT) tma_tile_3d(base + offBlo, &tmBlo, &full[stage], 0, 0, s);
                }
                for (int u = 0; u < cnt; ++u, ++s) {
                    if (leader) tma_im2col_4d(base + (uint32_t)u * Gm.a_bytes, &tmA, &full[stage], cb * KS, w0, h0, n0, (uint16_t)kx, (uint16_t)ky);
                    if (++cb == Gm.cblocks) { cb = 0; if (++kx == P.kw) { kx = 0; ++ky; } }
                }
                __syncwarp();
                if (++stage == S) { stage = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {  // ---- MMA issuer (whole warp converged; one elected lane issues)
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(Nt >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
        const int gmask = Gm.G - 1;
        int stage = 0;
        uint32_t ph = 0, lt = 0;
        for (int tile = blockIdx.x; tile < Gm.num_tiles; tile += gridDim.x, ++lt) {
            const uint32_t buf = lt & 1u;
            mbar_wait(&accempty[buf], ((lt >> 1) & 1u) ^ 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t dbase = tmem + buf * (uint32_t)(Gm.G * Nt);
            int s = 0;
            for (int f = 0; f < Gm.fills; ++f) {
                const int cnt = nslabs - s < spp ? nslabs - s : spp;
                mbar_wait(&ready[stage], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t base = ring0 + (uint32_t)stage * Gm.stage_bytes;
                if (elect_one()) {
                    for (int u = 0; u < cnt; ++u) {
                        const int sl = s + u;
                        const uint32_t d = dbase + (uint32_t)((sl & gmask) * Nt);
                        const uint32_t acc0 = sl >= Gm.G ? 1u : 0u;  // the first slab of an accumulator chain overwrites
                        const uint32_t ah = ((base + (uint32_t)u * Gm.a_bytes) >> 4) | 0x10000u, bh = ((base + offBhi + (uint32_t)u * Gm.b_bytes) >> 4) | 0x10000u;
#pragma unroll
                        for (int ks = 0; ks < KS / 8; ++ks)  // a K step of 8 tf32 = 32 bytes inside the swizzle atom: start address + 2
                            mma_tf32(d, ((uint64_t)kDescHi << 32) | (ah + 2u * ks), ((uint64_t)kDescHi << 32) | (bh + 2u * ks), idesc, ks ? 1u : acc0);
                        if (SPLIT) {
                            const uint32_t al = ((base + offAlo + (uint32_t)u * Gm.a_bytes) >> 4) | 0x10000u, bl = ((base + offBlo + (uint32_t)u * Gm.b_bytes) >> 4) | 0x10000u;
#pragma unroll
                            for (int ks = 0; ks < KS / 8; ++ks) mma_tf32(d, ((uint64_t)kDescHi << 32) | (al + 2u * ks), ((uint64_t)kDescHi << 32) | (bh + 2u * ks), idesc, 1u);
#pragma unroll
                            for (int ks = 0; ks < KS / 8; ++ks) mma_tf32(d, ((uint64_t)kDescHi << 32) | (ah + 2u * ks), ((uint64_t)kDescHi << 32) | (bl + 2u * ks), idesc, 1u);
                        }
                    }
                    commit(&empty[stage]);  // the stage is free once these MMAs have read it
                    if (f + 1 == Gm.fills) commit(&accfull[buf]);
                }
                __syncwarp();
                s += cnt;
                if (++stage == S) { stage = 0; ph ^= 1u; }
            }
        }
    } else if (warp >= 8) {  // ---- splitter: A_lo = A - tf32(A) for the 3xTF32 product (256 threads)
        const int ts = tid - 256;
        int stage = 0;
        uint32_t ph = 0;
        constexpr int kPer = (kM * (int)kSwb / 16) / 256;  // 16-byte chunks per thread and slab: 2 (KS = 16) | 4 (KS = 32)
        for (int tile = blockIdx.x; tile < Gm.num_tiles; tile += gridDim.x) {
            int s = 0;
            for (int f = 0; f < Gm.fills; ++f) {
                const int cnt = nslabs - s < spp ? nslabs - s : spp;
                mbar_wait(&full[stage], ph);
                if (SPLIT) {
                    const uint32_t a = ring0 + (uint32_t)stage * Gm.stage_bytes + (uint32_t)ts * 16u;
                    float4 v[4][kPer];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (u < cnt) {
#pragma unroll
                            for (int i = 0; i < kPer; ++i)
                                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[u][i].x), "=f"(v[u][i].y), "=f"(v[u][i].z), "=f"(v[u][i].w)
                                             : "r"(a + (uint32_t)u * Gm.a_bytes + (uint32_t)i * 4096u));
                        }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (u < cnt) {
#pragma unroll
                            for (int i = 0; i < kPer; ++i) {
                                float4 l = v[u][i];
                                l.x -= __uint_as_float(__float_as_uint(l.x) & 0xFFFFE000u);
                                l.y -= __uint_as_float(__float_as_uint(l.y) & 0xFFFFE000u);
                                l.z -= __uint_as_float(__float_as_uint(l.z) & 0xFFFFE000u);
                                l.w -= __uint_as_float(__float_as_uint(l.w) & 0xFFFFE000u);
                                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a + offAlo + (uint32_t)u * Gm.a_bytes + (uint32_t)i * 4096u), "f"(l.x), "f"(l.y), "f"(l.z), "f"(l.w));
                            }
                        }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&ready[stage]);
                s += cnt;
                if (++stage == S) { stage = 0; ph ^= 1u; }
            }
        }
    } else if (warp >= 4) {  // ---- epilogue: thread = pixel = TMEM lane; 32-channel blocks through the swizzled staging tile
        const int q = warp - 4, rr = 32 * q + lane;
        const bool affine = P.scale != nullptr;
        const int chains = Gm.G < nslabs ? Gm.G : nslabs;
        const uint32_t stg_row = smem0 + (uint32_t)rr * 128u, sw = (uint32_t)(rr & 7);
        uint32_t lt = 0;
        for (int tile = blockIdx.x; tile < Gm.num_tiles; tile += gridDim.x, ++lt) {
            const uint32_t buf = lt & 1u;
            const int64_t p = (int64_t)tile * kM + rr;
            const bool ok = p < Gm.M_total;
            const float* res = nullptr;
            if (P.res && ok) {
                const int pn = (int)(p / HoWo), rem = (int)(p - (int64_t)pn * HoWo), py = rem / P.Wo, px = rem - py * P.Wo;
                res = P.res + (((int64_t)pn * P.rH + (py * P.rsy + P.ry0)) * P.rW + (px * P.rsx + P.rx0)) * P.Cout;
            }
            mbar_wait(&accfull[buf], (lt >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // the previous tile's stores must have read the staging tile before it is overwritten
            if (tid == 128) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const uint32_t trow = tmem + ((uint32_t)(32 * q) << 16) + buf * (uint32_t)(Gm.G * Nt);
            for (int c0 = 0; c0 < Nt; c0 += 16) {
                float4 cb[4], cs[4], ch[4];  // bias / scale / shift of this chunk: issued before the TMEM read so their latency hides under it
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    cb[i] = *reinterpret_cast<const float4*>(s_bias + c0 + 4 * i);
                    cs[i] = *reinterpret_cast<const float4*>(s_scale + c0 + 4 * i);
                    ch[i] = *reinterpret_cast<const float4*>(s_shift + c0 + 4 * i);
                }
                float4 rv[4];
                if (res) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) rv[i] = __ldg(reinterpret_cast<const float4*>(res + c0 + 4 * i));
                }
                float v[16];
                tmem_ld16(trow + (uint32_t)c0, v);
                for (int g = 1; g < chains; ++g) {
                    float u[16];
                    tmem_ld16(trow + (uint32_t)(g * Nt + c0), u);
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] += u[i];
                }
                const uint32_t blk = stg_row + (uint32_t)(c0 >> 5) * (uint32_t)(kM * 128);  // staging tile of this 32-channel block
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    const float4 b = cb[i >> 2];
                    float4 a = make_float4(v[i] + b.x, v[i + 1] + b.y, v[i + 2] + b.z, v[i + 3] + b.w);
                    if (res) {
                        const float4 r = rv[i >> 2];
                        a.x += r.x; a.y += r.y; a.z += r.z; a.w += r.w;
                    }
                    if (P.act == 1) {
                        a.x = fmaxf(a.x, 0.0f); a.y = fmaxf(a.y, 0.0f); a.z = fmaxf(a.z, 0.0f); a.w = fmaxf(a.w, 0.0f);
                    } else if (P.act == 2) {
                        a.x = a.x > 0.0f ? a.x : expm1f(a.x); a.y = a.y > 0.0f ? a.y : expm1f(a.y);
                        a.z = a.z > 0.0f ? a.z : expm1f(a.z); a.w = a.w > 0.0f ? a.w : expm1f(a.w);
                    }
                    if (affine) {
                        const float4 sc = cs[i >> 2], sh = ch[i >> 2];
                        a.x = a.x * sc.x + sh.x; a.y = a.y * sc.y + sh.y; a.z = a.z * sc.z + sh.z; a.w = a.w * sc.w + sh.w;
                    }
                    const uint32_t chunk = (uint32_t)(((c0 & 16) + i) >> 2);  // 16-byte chunk 0..7 of the 128-byte row; SWIZZLE_128B: chunk ^ (row & 7)
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(blk + ((chunk ^ sw) << 4)), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w));
                }
            }
            // the accumulator set is free as soon as it has been read
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&accempty[buf]);
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (tid == 128) {
                for (int cbk = 0; cbk < Nt / 32; ++cbk) tma_store_2d(&tmY, smem0 + (uint32_t)cbk * (uint32_t)(kM * 128), cbk * 32, tile * kM);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
        if (tid == 128) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)Gm.tmem_cols) : "memory");
}

// ---- first layer (one input channel, 5x5 / stride 2) on the tensor cores ---------------------------------------------------------
// GEMM view: M = the pixels of R whole output rows of one image (R * Wo <= 128: 2 x 60 for the CNN, 1 x 106 for the VAE), K = 25 taps
// padded to 32, N = Cout (16 | 32).  Cin = 1 rules out the TMA im2col mode (its channel run is 4 bytes), so the operand is built by
// threads: a TMA tile load brings the tile's input strip [(R-1)*2 + 5 rows][from column -4 to (Wo-1)*2 + 4 - pad], image borders zero-filled by the
// TMA unit (plus the matching strips of the per-pixel mean / rstd planes when the RunningMeanStd normalisation is fused: outside the
// image all three are 0, so the normalised padding stays 0); 256 builder threads normalise the strip in place, then write each
// pixel's 25 taps as a 128-byte row of the SWIZZLE_128B K-major layout, hi and lo parts.  ~100 instructions per output pixel
// instead of the ~700 of the direct kernel (25 x (load + 16 FMA + weight reads)).  Weights [Cout][32] hi / lo stay resident in
// shared memory.  Roles as above: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-7 epilogue, warps 8-15 builders.
struct FirstGeom {
    int32_t num_tiles, tiles_per_img, R, rows_px, SR, SW, strip_bytes, norm, sstages, astages;
    uint32_t stg_bytes;
};

template <int HALF>
__device__ __forceinline__ void build_row(const float* sp, int SW, uint32_t row_hi, uint32_t row_lo, uint32_t sw) {
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        constexpr int kBase = HALF * 16;
        const int k = kBase + i;
        v[i] = k < 25 ? sp[(k / 5) * SW + (k % 5)] : 0.0f;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t off = (((uint32_t)(HALF * 4 + j)) ^ sw) << 4;
        float l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) l[e] = v[4 * j + e] - __uint_as_float(__float_as_uint(v[4 * j + e]) & 0xFFFFE000u);
        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(row_hi + off), "f"(v[4 * j]), "f"(v[4 * j + 1]), "f"(v[4 * j + 2]), "f"(v[4 * j + 3]) : "memory");
        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(row_lo + off), "f"(l[0]), "f"(l[1]), "f"(l[2]), "f"(l[3]) : "memory");
    }
}

template <int COUT>
__global__ void __launch_bounds__(kThreads, 1)
agx_conv2d_first_tma_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmMean, const __grid_constant__ CUtensorMap tmRstd,
                            const __grid_constant__ CUtensorMap tmY, const __grid_constant__ AgxConvFirstParams P, const __grid_constant__ FirstGeom Gm) {
    __shared__ __align__(8) uint64_t sfull[4], sempty[4], aready[4], aempty[4], accfull[2], accempty[2];
    __shared__ uint32_t tmem_base;
    __shared__ __align__(16) float s_bias[COUT], s_scale[COUT], s_shift[COUT];
    constexpr uint32_t kDescHi = ((1024u >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);  // SWIZZLE_128B K-major, 8-row groups 1024 bytes apart
    constexpr uint32_t kABytes = kM * 128, kBBytes = COUT * 128;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t smem0 = (s32(t_smem) + 1023u) & ~1023u;
    // [B_hi | B_lo | staging | A ring (hi, lo per stage) | strip ring (image, mean, rstd per stage)]
    const uint32_t offB = smem0, offStg = offB + 2 * ((kBBytes + 1023u) & ~1023u), offA = offStg + Gm.stg_bytes, offS = offA + (uint32_t)Gm.astages * 2u * kABytes;
    const uint32_t strip_stage = (uint32_t)Gm.strip_bytes * (Gm.norm ? 3u : 1u);
    const int SS = Gm.sstages, AS = Gm.astages;

    if (tid == 0) {
        for (int i = 0; i < SS; ++i) { mbar_init(&sfull[i], 1); mbar_init(&sempty[i], 8); }
        for (int i = 0; i < AS; ++i) { mbar_init(&aready[i], 8); mbar_init(&aempty[i], 1); }
        mbar_init(&accfull[0], 1); mbar_init(&accfull[1], 1); mbar_init(&accempty[0], 4); mbar_init(&accempty[1], 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmX)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmY)) : "memory");
    }
    if (tid < COUT) {
        s_bias[tid] = P.bias ? P.bias[tid] : 0.0f;
        s_scale[tid] = P.scale ? P.scale[tid] : 1.0f;
        s_shift[tid] = P.shift ? P.shift[tid] : 0.0f;
    }
    // resident weights: row o = [w[o][0..24], 0 x 7] as hi (tf32, round to nearest) + lo, SWIZZLE_128B rows
    for (int i = tid; i < COUT * 32; i += kThreads) {
        const int o = i >> 5, k = i & 31;
        const float w = k < 25 ? P.w[o * 25 + k] : 0.0f;
        const float hi = __uint_as_float((__float_as_uint(w) + 0x1000u) & 0xFFFFE000u);
        const uint32_t off = (uint32_t)o * 128u + ((((uint32_t)k >> 2) ^ ((uint32_t)o & 7u)) << 4) + ((uint32_t)k & 3u) * 4u;
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(offB + off), "f"(hi) : "memory");
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(offB + ((kBBytes + 1023u) & ~1023u) + off), "f"(w - hi) : "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_base)), "r"((uint32_t)(2 * COUT)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;

    if (warp == 0) {  // ---- TMA producer: the input strip of each tile (+ the mean / rstd strips)
        int stage = 0;
        uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < Gm.num_tiles; tile += gridDim.x) {
            const int n = tile / Gm.tiles_per_img, oy0 = (tile - n * Gm.tiles_per_img) * Gm.R;
            mbar_wait(&sempty[stage], ph ^ 1u);
            if (elect_one()) {
                const uint32_t dst = offS + (uint32_t)stage * strip_stage;
                mbar_expect_tx(&sfull[stage], (uint32_t)(Gm.SR * Gm.SW * 4) * (Gm.norm ? 3u : 1u));
                tma_tile_3d(dst, &tmX, &sfull[stage], -4, oy0 * 2 - P.py, n);
                if (Gm.norm) {
                    tma_tile_2d(dst + (uint32_t)Gm.strip_bytes, &tmMean, &sfull[stage], -4, oy0 * 2 - P.py);
                    tma_tile_2d(dst + 2u * (uint32_t)Gm.strip_bytes, &tmRstd, &sfull[stage], -4, oy0 * 2 - P.py);
                }
            }
            __syncwarp();
            if (++stage == SS) { stage = 0; ph ^= 1u; }
        }
    } else if (warp == 1) {  // ---- MMA issuer
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(COUT >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
        const uint32_t bh = (offB >> 4) | 0x10000u, bl = ((offB + ((kBBytes + 1023u) & ~1023u)) >> 4) | 0x10000u;
        int stage = 0;
        uint32_t ph = 0, lt = 0;
        for (int tile = blockIdx.x; tile < Gm.num_tiles; tile += gridDim.x, ++lt) {
            const uint32_t buf = lt & 1u;
            mbar_wait(&accempty[buf], ((lt >> 1) & 1u) ^ 1u);
            mbar_wait(&aready[stage], ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                const uint32_t d = tmem + buf * (uint32_t)COUT;
                const uint32_t ah = ((offA + (uint32_t)stage * 2u * kABytes) >> 4) | 0x10000u, al = ah + (kABytes >> 4);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) mma_tf32(d, ((uint64_t)kDescHi << 32) | (ah + 2u * ks), ((uint64_t)kDescHi << 32) | (bh + 2u * ks), idesc, ks ? 1u : 0u);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) mma_tf32(d, ((uint64_t)kDescHi << 32) | (al + 2u * ks), ((uint64_t)kDescHi << 32) | (bh + 2u * ks), idesc, 1u);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) mma_tf32(d, ((uint64_t)kDescHi << 32) | (ah + 2u * ks), ((uint64_t)kDescHi << 32) | (bl + 2u * ks), idesc, 1u);
                commit(&aempty[stage]);
                commit(&accfull[buf]);
            }
            __syncwarp();
            if (++stage == AS) { stage = 0; ph ^= 1u; }
        }
    } else if (warp >= 8) {  // ---- builders: normalise the strip in place, then one 128-byte operand row (hi, lo) per output pixel
        const int bt = tid - 256, r = bt & 127;
        const int ry = r / P.Wo, rx = r - ry * P.Wo;
        const bool rvalid = r < Gm.rows_px;
        const uint32_t sw = (uint32_t)(r & 7);
        int sstage = 0, astage = 0;
        uint32_t sph = 0, aph = 0;
        const int strip_n = Gm.SR * Gm.SW;
        for (int tile = blockIdx.x; tile < Gm.num_tiles; tile += gridDim.x) {
            float* strip = reinterpret_cast<float*>(t_smem + (offS - s32(t_smem)) + (size_t)sstage * strip_stage);
            mbar_wait(&sfull[sstage], sph);
            if (Gm.norm) {  // RunningMeanStd forward: clamp((x - mean) * rstd, +-5) — the expression of the direct kernel
                const float* mean = strip + Gm.strip_bytes / 4;
                const float* rstd = mean + Gm.strip_bytes / 4;
                for (int i = bt; i < strip_n; i += 256) strip[i] = fminf(fmaxf((strip[i] - mean[i]) * rstd[i], -5.0f), 5.0f);
                asm volatile("bar.sync 2, 256;" ::: "memory");
            }
            mbar_wait(&aempty[astage], aph ^ 1u);
            if (rvalid) {
                const uint32_t row_hi = offA + (uint32_t)astage * 2u * kABytes + (uint32_t)r * 128u;
                const float* sp = strip + (ry * 2) * Gm.SW + rx * 2 + (4 - P.px);  // the strip starts at input column -4
                if (bt < 128) build_row<0>(sp, Gm.SW, row_hi, row_hi + kABytes, sw);
                else build_row<1>(sp, Gm.SW, row_hi, row_hi + kABytes, sw);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) { mbar_arrive(&aready[astage]); mbar_arrive(&sempty[sstage]); }
            if (++sstage == SS) { sstage = 0; sph ^= 1u; }
            if (++astage == AS) { astage = 0; aph ^= 1u; }
        }
    } else if (warp >= 4) {  // ---- epilogue
        const int q = warp - 4, rr = 32 * q + lane;
        const bool affine = P.scale != nullptr;
        // staging rows of COUT * 4 bytes: SWIZZLE_64B (chunk ^ ((row >> 1) & 3)) for 16 channels, SWIZZLE_128B (chunk ^ (row & 7)) for 32
        const uint32_t stg_row = offStg + (uint32_t)rr * (uint32_t)(COUT * 4), sw = COUT == 16 ? (uint32_t)((rr >> 1) & 3) : (uint32_t)(rr & 7);
        uint32_t lt = 0;
        for (int tile = blockIdx.x; tile < Gm.num_tiles; tile += gridDim.x, ++lt) {
            const uint32_t buf = lt & 1u;
            mbar_wait(&accfull[buf], (lt >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (tid == 128) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const uint32_t trow = tmem + ((uint32_t)(32 * q) << 16) + buf * (uint32_t)COUT;
#pragma unroll
            for (int c0 = 0; c0 < COUT; c0 += 16) {
                float v[16];
                tmem_ld16(trow + (uint32_t)c0, v);
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    const float4 b = *reinterpret_cast<const float4*>(s_bias + c0 + i);
                    float4 a = make_float4(v[i] + b.x, v[i + 1] + b.y, v[i + 2] + b.z, v[i + 3] + b.w);
                    if (P.act == 1) {
                        a.x = fmaxf(a.x, 0.0f); a.y = fmaxf(a.y, 0.0f); a.z = fmaxf(a.z, 0.0f); a.w = fmaxf(a.w, 0.0f);
                    } else if (P.act == 2) {
                        a.x = a.x > 0.0f ? a.x : expm1f(a.x); a.y = a.y > 0.0f ? a.y : expm1f(a.y);
                        a.z = a.z > 0.0f ? a.z : expm1f(a.z); a.w = a.w > 0.0f ? a.w : expm1f(a.w);
                    }
                    if (affine) {
                        const float4 sc = *reinterpret_cast<const float4*>(s_scale + c0 + i), sh = *reinterpret_cast<const float4*>(s_shift + c0 + i);
                        a.x = a.x * sc.x + sh.x; a.y = a.y * sc.y + sh.y; a.z = a.z * sc.z + sh.z; a.w = a.w * sc.w + sh.w;
                    }
                    const uint32_t chunk = (uint32_t)((c0 + i) >> 2);
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(stg_row + ((chunk ^ sw) << 4)), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w) : "memory");
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&accempty[buf]);
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (tid == 128) {
                const int n = tile / Gm.tiles_per_img, oy0 = (tile - n * Gm.tiles_per_img) * Gm.R;
                tma_store_2d(&tmY, offStg, 0, (n * P.Ho + oy0) * P.Wo);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
        if (tid == 128) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)(2 * COUT)) : "memory");
}

// ---- host side: tensor maps through the driver entry points (no libcuda link dependency) ---------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const int*, const int*,
                                   cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode_tiled = nullptr;
EncodeIm2colFn g_encode_im2col = nullptr;
int g_entry_state = 0;  // 0 not looked up, 1 available, -1 unavailable
int g_sm_count = 0;
int g_conv_impl = 1;    // agx_set_option("conv_impl", 0 = cp.async gather kernel of agx_conv.cu | 1 = this kernel where the geometry allows)
int g_conv_spp = 0;     // agx_set_option("conv_spp", 0 = automatic | 1..4 slabs per stage) — A/B knob
int g_conv_stages = 0;  // agx_set_option("conv_stages", 0 = as many as fit | 2..8) — A/B knob

bool lookup_entry_points() {
    if (g_entry_state) return g_entry_state > 0;
    void *f1 = nullptr, *f2 = nullptr;
    cudaDriverEntryPointQueryResult q1, q2;
    const bool ok = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f1, cudaEnableDefault, &q1) == cudaSuccess && q1 == cudaDriverEntryPointSuccess && f1 &&
                    cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f2, cudaEnableDefault, &q2) == cudaSuccess && q2 == cudaDriverEntryPointSuccess && f2;
    cudaGetLastError();
    g_encode_tiled = reinterpret_cast<EncodeTiledFn>(f1);
    g_encode_im2col = reinterpret_cast<EncodeIm2colFn>(f2);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (g_sm_count <= 0) g_sm_count = 148;
    g_entry_state = ok ? 1 : -1;
    return ok;
}

template <int KS, bool SPLIT>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmBhi, const CUtensorMap& tmBlo, const CUtensorMap& tmY, const AgxConvParams* p, const ConvGeom& G, size_t smem,
           cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(agx_conv2d_tma_kernel<KS, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024) != cudaSuccess) { cudaGetLastError(); return 0; }
        attr_set = true;
    }
    const int grid = G.num_tiles < g_sm_count ? G.num_tiles : g_sm_count;
    agx_conv2d_tma_kernel<KS, SPLIT><<<grid, kThreads, smem, st>>>(tmA, tmBhi, tmBlo, tmY, *p, G);
    return cudaGetLastError() == cudaSuccess ? 1 : agx_internal_fail(AGX_ERR_CUDA, "agx_conv2d_nhwc (tma): launch failed");
}

}  // namespace

extern "C" {

int agx_internal_conv_option(const char* key, int value) {
    if (!strcmp(key, "conv_impl")) {
        if (value < 0 || value > 1) return -1;
        g_conv_impl = value;
        return 1;
    }
    if (!strcmp(key, "conv_first")) {
        if (value < 0 || value > 3) return -1;
        g_first_impl = value;
        return 1;
    }
    if (!strcmp(key, "conv_stages")) {
        if (value != 0 && (value < 2 || value > kMaxStages)) return -1;
        g_conv_stages = value;
        return 1;
    }
    if (!strcmp(key, "conv_spp")) {
        if (value < 0 || value > 4) return -1;
        g_conv_spp = value;
        return 1;
    }
    return 0;
}

// tiled tensor map for other translation units (agx_mlp_train.cu): rank 2 / 3, fp32, swizzle 0 / 64 / 128 bytes; 1 = encoded, 0 = unavailable
int agx_internal_tmap_tiled(void* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
    if (!lookup_entry_points() || rank < 2 || rank > 3) return 0;
    cuuint64_t d[3], st[2];
    cuuint32_t b[3], es[3] = {1, 1, 1};
    for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; }
    for (int i = 0; i + 1 < rank; ++i) st[i] = strides_bytes[i];
    const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE);
    return g_encode_tiled(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), d, st, b, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 1 : 0;
}

// returns 1 when the layer was launched here, 0 when the geometry belongs to the gather kernel, < 0 on error (already recorded)
int agx_internal_conv_tma(const AgxConvParams* p, void* stream) {
    if (!g_conv_impl) return 0;
    const int taps = p->kh * p->kw;
    const bool cin_ok = p->Cin == 16 || (p->Cin >= 32 && p->Cin % 32 == 0 && p->Cin <= 256);
    if (taps <= 1 || !cin_ok || p->Cout > 128 || (p->Cout & 31) || p->kh > 16 || p->kw > 16 || p->sy > 8 || p->sx > 8 || p->py > 16 || p->px > 16) return 0;
    if (!lookup_entry_points()) return 0;
    ConvGeom G;
    memset(&G, 0, sizeof(G));
    G.M_total = (int64_t)p->N * p->Ho * p->Wo;
    if ((G.M_total + kM - 1) / kM > 0x7FFFFF) return 0;
    G.num_tiles = (int)((G.M_total + kM - 1) / kM);
    G.Nt = p->Cout;
    const int KS = p->Cin == 16 ? 16 : 32, swb = KS * 4;
    const bool split = p->w_lo != nullptr;
    G.cblocks = p->Cin / KS;
    G.nslabs = taps * G.cblocks;
    const int K = taps * p->Cin;
    G.G = K <= 512 ? 1 : ((K <= 1024 || G.Nt > 64) ? 2 : 4);  // accumulator chains, as in agx_conv.cu
    const int cols = 2 * G.G * G.Nt;
    if (cols > 512) return 0;
    G.tmem_cols = cols <= 32 ? 32 : (cols <= 64 ? 64 : (cols <= 128 ? 128 : (cols <= 256 ? 256 : 512)));
    G.a_bytes = (uint32_t)kM * swb;
    G.b_bytes = (uint32_t)G.Nt * swb;  // a multiple of 1024: Nt % 32 == 0, swb >= 64
    const uint32_t slab_bytes = (split ? 2u : 1u) * (G.a_bytes + G.b_bytes);
    G.stg_bytes = (uint32_t)G.Nt * kM * 4;
    const uint32_t budget = 220u * 1024u - G.stg_bytes;
    G.spp = g_conv_spp ? g_conv_spp : (int)(65536u / slab_bytes);
    if (G.spp > 4) G.spp = 4;
    if (G.spp > G.nslabs) G.spp = G.nslabs;
    while (G.spp > 1 && budget / (G.spp * slab_bytes) < 2) --G.spp;
    if (G.spp < 1) G.spp = 1;
    G.stage_bytes = (uint32_t)G.spp * slab_bytes;
    G.stages = (int)(budget / G.stage_bytes);
    if (G.stages > kMaxStages) G.stages = kMaxStages;
    if (g_conv_stages && G.stages > g_conv_stages) G.stages = g_conv_stages;
    if (G.stages < 2) return 0;
    G.fills = (G.nslabs + G.spp - 1) / G.spp;
    const size_t smem = 1024 + (size_t)G.stg_bytes + (size_t)G.stages * G.stage_bytes;

    const CUtensorMapSwizzle sw = swb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    CUtensorMap tmA, tmBhi, tmBlo, tmY;
    memset(&tmBlo, 0, sizeof(tmBlo));
    {   // activations as (C, W, H, N); the base-pixel box [-pad, dim + pad - (k - 1)) is traversed with the convolution's stride
        const cuuint64_t dims[4] = {(cuuint64_t)p->Cin, (cuuint64_t)p->W, (cuuint64_t)p->H, (cuuint64_t)p->N};
        const cuuint64_t strides[3] = {(cuuint64_t)p->Cin * 4, (cuuint64_t)p->W * p->Cin * 4, (cuuint64_t)p->H * p->W * p->Cin * 4};
        const int lower[2] = {-p->px, -p->py}, upper[2] = {p->px - (p->kw - 1), p->py - (p->kh - 1)};
        const cuuint32_t estr[4] = {1, (cuuint32_t)p->sx, (cuuint32_t)p->sy, 1};
        const CUresult r = g_encode_im2col(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(p->x), dims, strides, lower, upper, (cuuint32_t)KS,
                                           (cuuint32_t)kM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return 0;  // geometry the TMA unit does not take: the gather kernel handles it
    }
    for (int part = 0; part < (split ? 2 : 1); ++part) {  // weights [Cout][K] viewed as (KS, Cout, slab): a box = the slabs of one stage fill
        const cuuint64_t dims[3] = {(cuuint64_t)KS, (cuuint64_t)p->Cout, (cuuint64_t)G.nslabs};
        const cuuint64_t strides[2] = {(cuuint64_t)K * 4, (cuuint64_t)KS * 4};
        const cuuint32_t box[3] = {(cuuint32_t)KS, (cuuint32_t)G.Nt, (cuuint32_t)G.spp}, estr[3] = {1, 1, 1};
        const CUresult r = g_encode_tiled(part ? &tmBlo : &tmBhi, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(part ? p->w_lo : p->w_hi), dims, strides,
                                          box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return 0;
    }
    {   // output as [M_total][Cout]: boxes of 32 channels x 128 pixels, rows past M_total clipped by the TMA unit
        const cuuint64_t dims[2] = {(cuuint64_t)p->Cout, (cuuint64_t)G.M_total};
        const cuuint64_t strides[1] = {(cuuint64_t)p->Cout * 4};
        const cuuint32_t box[2] = {32, (cuuint32_t)kM}, estr[2] = {1, 1};
        const CUresult r = g_encode_tiled(&tmY, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, p->y, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return 0;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (KS == 16) return split ? launch<16, true>(tmA, tmBhi, tmBlo, tmY, p, G, smem, st) : launch<16, false>(tmA, tmBhi, tmBhi, tmY, p, G, smem, st);
    return split ? launch<32, true>(tmA, tmBhi, tmBlo, tmY, p, G, smem, st) : launch<32, false>(tmA, tmBhi, tmBhi, tmY, p, G, smem, st);
}

// first layer: returns 1 when launched here, 0 when the geometry belongs to the direct kernel of agx_conv.cu, < 0 on error
int agx_internal_conv_first_tma(const AgxConvFirstParams* p, void* stream) {
    if (p->kh != 5 || p->kw != 5 || p->sy != 2 || p->sx != 2 || (p->Cout != 16 && p->Cout != 32) || p->Wo > kM || p->Wo <= 0 || (p->W & 3) || p->px > 4 || p->py > 8) return 0;
    if ((int64_t)p->N * p->Ho * p->Wo > 0x7FFFFFFF) return 0;
    if (((uintptr_t)p->x | (uintptr_t)p->px_mean | (uintptr_t)p->px_rstd) & 15u) return 0;
    if (!lookup_entry_points()) return 0;
    FirstGeom G;
    memset(&G, 0, sizeof(G));
    G.R = 1;
    for (int r = kM / p->Wo; r >= 1; --r)
        if (p->Ho % r == 0) { G.R = r; break; }  // whole output rows per tile, a divisor of Ho: the stored box never crosses into the next image
    G.rows_px = G.R * p->Wo;
    G.tiles_per_img = p->Ho / G.R;
    if ((int64_t)p->N * G.tiles_per_img > 0x7FFFFFFF) return 0;
    G.num_tiles = p->N * G.tiles_per_img;
    G.SR = (G.R - 1) * 2 + 5;
    G.SW = (((p->Wo - 1) * 2 + 5 + (4 - p->px)) + 3) & ~3;  // from input column -4: the inner start coordinate of a TMA tile load must be 16-byte aligned
    if (G.SW > 256 || G.SR > 256) return 0;
    G.strip_bytes = (G.SR * G.SW * 4 + 127) & ~127;
    G.norm = p->px_mean != nullptr;
    G.stg_bytes = (uint32_t)kM * p->Cout * 4;
    G.astages = 3;
    G.sstages = 4;
    const size_t smem = 1024 + 2 * (((size_t)p->Cout * 128 + 1023) & ~(size_t)1023) + G.stg_bytes + (size_t)G.astages * 2 * kM * 128 +
                        (size_t)G.sstages * G.strip_bytes * (G.norm ? 3 : 1);
    if (smem > 222 * 1024) return 0;
    CUtensorMap tmX, tmMean, tmRstd, tmY;
    {
        const cuuint64_t dims[3] = {(cuuint64_t)p->W, (cuuint64_t)p->H, (cuuint64_t)p->N};
        const cuuint64_t strides[2] = {(cuuint64_t)p->W * 4, (cuuint64_t)p->H * p->W * 4};
        const cuuint32_t box[3] = {(cuuint32_t)G.SW, (cuuint32_t)G.SR, 1}, estr[3] = {1, 1, 1};
        if (g_encode_tiled(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(p->x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return 0;
    }
    tmMean = tmX;
    tmRstd = tmX;
    if (G.norm) {
        const cuuint64_t dims[2] = {(cuuint64_t)p->W, (cuuint64_t)p->H};
        const cuuint64_t strides[1] = {(cuuint64_t)p->W * 4};
        const cuuint32_t box[2] = {(cuuint32_t)G.SW, (cuuint32_t)G.SR}, estr[2] = {1, 1};
        for (int part = 0; part < 2; ++part)
            if (g_encode_tiled(part ? &tmRstd : &tmMean, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(part ? p->px_rstd : p->px_mean), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                return 0;
    }
    {
        const cuuint64_t dims[2] = {(cuuint64_t)p->Cout, (cuuint64_t)p->N * p->Ho * p->Wo};
        const cuuint64_t strides[1] = {(cuuint64_t)p->Cout * 4};
        const cuuint32_t box[2] = {(cuuint32_t)p->Cout, (cuuint32_t)G.rows_px}, estr[2] = {1, 1};
        if (g_encode_tiled(&tmY, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, p->y, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           p->Cout == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return 0;
    }
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(agx_conv2d_first_tma_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024) != cudaSuccess ||
            cudaFuncSetAttribute(agx_conv2d_first_tma_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024) != cudaSuccess) { cudaGetLastError(); return 0; }
        attr_set = true;
    }
    const int grid = G.num_tiles < g_sm_count ? G.num_tiles : g_sm_count;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (p->Cout == 16) agx_conv2d_first_tma_kernel<16><<<grid, kThreads, smem, st>>>(tmX, tmMean, tmRstd, tmY, *p, G);
    else agx_conv2d_first_tma_kernel<32><<<grid, kThreads, smem, st>>>(tmX, tmMean, tmRstd, tmY, *p, G);
    return cudaGetLastError() == cudaSuccess ? 1 : agx_internal_fail(AGX_ERR_CUDA, "agx_conv2d_first (tma): launch failed");
}

}  // extern "C"

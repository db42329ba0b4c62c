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
//   tensor core reads.  Roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warps 4-7 =
//   epilogue (TMEM lane quarters 0-3), warps 8-11 = splitter (A_lo = A - tf32(A), element-wise on the stage: the layout does not
//   matter).  Ring of 3-8 shared-memory stages (full → ready → empty mbarriers) and two TMEM accumulator sets, so the loads of tile
//   t+1 run under the MMAs of tile t and under the epilogue of tile t-1.
//   Precision: 3xTF32 as in agx_conv.cu (A_hi·B_hi + A_lo·B_hi + A_hi·B_lo; G accumulator chains for K > 512).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "agx.h"
#include "agx_tc.cuh"

int agx_internal_fail(int code, const char* msg);

namespace {
using namespace tc;

constexpr int kThreads = 384;
constexpr int kMaxStages = 8;

struct ConvGeom {
    int64_t M_total;
    int32_t num_tiles, Nt, KS, swb, cblocks, nslabs, stages, G, tmem_cols, split;
    uint32_t a_bytes, b_bytes, stage_bytes, tx_bytes;
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
// K-major operand in the 64- / 128-byte swizzled canonical layout: rows of `swb` bytes, 8-row groups `8 * swb` bytes apart
__device__ __forceinline__ uint64_t smem_desc_sw(uint32_t saddr, uint32_t swb) {
    const uint64_t layout = swb == 128 ? 2u : 4u;  // UMMA LayoutType: SWIZZLE_128B = 2, SWIZZLE_64B = 4
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(((8u * swb) >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | (layout << 61);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db),
                 "r"(idesc), "r"(acc)
                 : "memory");
}

extern __shared__ uint8_t t_smem[];

__global__ void __launch_bounds__(kThreads, 1)
agx_conv2d_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo,
                      const __grid_constant__ AgxConvParams P, const __grid_constant__ ConvGeom Gm) {
    __shared__ __align__(8) uint64_t full[kMaxStages], ready[kMaxStages], empty[kMaxStages], accfull[2], accempty[2];
    __shared__ uint32_t tmem_base;
    __shared__ __align__(16) float s_bias[128], s_scale[128], s_shift[128];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t smem0 = (s32(t_smem) + 1023u) & ~1023u;  // swizzle patterns repeat every 1024 bytes of the shared-memory address
    const int S = Gm.stages, Nt = Gm.Nt, nslabs = Gm.nslabs;
    const bool split = Gm.split != 0;
    const uint32_t offAlo = Gm.a_bytes, offBhi = (split ? 2u : 1u) * Gm.a_bytes, offBlo = offBhi + Gm.b_bytes;

    if (tid == 0) {
        for (int i = 0; i < S; ++i) { mbar_init(&full[i], 1); mbar_init(&ready[i], 4); mbar_init(&empty[i], 1); }
        mbar_init(&accfull[0], 1); mbar_init(&accfull[1], 1); mbar_init(&accempty[0], 4); mbar_init(&accempty[1], 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmBhi)) : "memory");
        if (split) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmBlo)) : "memory");
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

    if (warp == 0) {
        if (lane == 0) {  // ---- TMA producer
            int stage = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < Gm.num_tiles; tile += gridDim.x) {
                const int64_t pix0 = (int64_t)tile * kM;
                const int n0 = (int)(pix0 / HoWo), rem = (int)(pix0 - (int64_t)n0 * HoWo), oy0 = rem / P.Wo, ox0 = rem - oy0 * P.Wo;
                const int w0 = ox0 * P.sx - P.px, h0 = oy0 * P.sy - P.py;  // base pixel of the tile's first output pixel, input coordinates
                int tap = 0, cb = 0, ky = 0, kx = 0;
                for (int s = 0; s < nslabs; ++s) {
                    mbar_wait(&empty[stage], ph ^ 1u);
                    mbar_expect_tx(&full[stage], Gm.tx_bytes);
                    const uint32_t base = smem0 + (uint32_t)stage * Gm.stage_bytes;
                    tma_im2col_4d(base, &tmA, &full[stage], cb * Gm.KS, w0, h0, n0, (uint16_t)kx, (uint16_t)ky);
                    tma_tile_2d(base + offBhi, &tmBhi, &full[stage], s * Gm.KS, 0);
                    if (split) tma_tile_2d(base + offBlo, &tmBlo, &full[stage], s * Gm.KS, 0);
                    if (++cb == Gm.cblocks) { cb = 0; ++tap; if (++kx == P.kw) { kx = 0; ++ky; } }
                    if (++stage == S) { stage = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ---- MMA issuer
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(Nt >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
            const int ksteps = Gm.KS / 8;
            int stage = 0;
            uint32_t ph = 0, lt = 0;
            for (int tile = blockIdx.x; tile < Gm.num_tiles; tile += gridDim.x, ++lt) {
                const uint32_t buf = lt & 1u;
                mbar_wait(&accempty[buf], ((lt >> 1) & 1u) ^ 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t dbase = tmem + buf * (uint32_t)(Gm.G * Nt);
                int g = 0;
                for (int s = 0; s < nslabs; ++s) {
                    mbar_wait(&ready[stage], ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t base = smem0 + (uint32_t)stage * Gm.stage_bytes;
                    const uint32_t d = dbase + (uint32_t)(g * Nt);
                    const bool fresh = s < Gm.G;  // first slab of this accumulator chain overwrites
                    for (int ks = 0; ks < ksteps; ++ks)
                        mma_tf32(d, smem_desc_sw(base + ks * 32, Gm.swb), smem_desc_sw(base + offBhi + ks * 32, Gm.swb), idesc, (ks > 0 || !fresh) ? 1u : 0u);
                    if (split) {
                        for (int ks = 0; ks < ksteps; ++ks)
                            mma_tf32(d, smem_desc_sw(base + offAlo + ks * 32, Gm.swb), smem_desc_sw(base + offBhi + ks * 32, Gm.swb), idesc, 1u);
                        for (int ks = 0; ks < ksteps; ++ks)
                            mma_tf32(d, smem_desc_sw(base + ks * 32, Gm.swb), smem_desc_sw(base + offBlo + ks * 32, Gm.swb), idesc, 1u);
                    }
                    commit(&empty[stage]);  // the stage is free once these MMAs have read it
                    if (++g == Gm.G) g = 0;
                    if (++stage == S) { stage = 0; ph ^= 1u; }
                }
                commit(&accfull[buf]);
            }
        }
    } else if (warp >= 8) {  // ---- splitter: A_lo = A - tf32(A) for the 3xTF32 product
        const int ts = tid - 256;
        int stage = 0;
        uint32_t ph = 0;
        const int n4 = (int)(Gm.a_bytes >> 4);
        for (int tile = blockIdx.x; tile < Gm.num_tiles; tile += gridDim.x) {
            for (int s = 0; s < nslabs; ++s) {
                mbar_wait(&full[stage], ph);
                if (split) {
                    const uint32_t a = smem0 + (uint32_t)stage * Gm.stage_bytes;
                    for (int i = ts; i < n4; i += 128) {
                        float4 v;
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a + (uint32_t)i * 16u));
                        v.x -= __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
                        v.y -= __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
                        v.z -= __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
                        v.w -= __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
                        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a + offAlo + (uint32_t)i * 16u), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&ready[stage]);
                if (++stage == S) { stage = 0; ph ^= 1u; }
            }
        }
    } else if (warp >= 4) {  // ---- epilogue: thread = pixel = TMEM lane
        const int q = warp - 4, rr = 32 * q + lane;
        const bool affine = P.scale != nullptr;
        uint32_t lt = 0;
        for (int tile = blockIdx.x; tile < Gm.num_tiles; tile += gridDim.x, ++lt) {
            const uint32_t buf = lt & 1u;
            mbar_wait(&accfull[buf], (lt >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int64_t p = (int64_t)tile * kM + rr;
            const bool ok = p < Gm.M_total;
            const float* res = nullptr;
            if (P.res && ok) {
                const int pn = (int)(p / HoWo), rem = (int)(p - (int64_t)pn * HoWo), py = rem / P.Wo, px = rem - py * P.Wo;
                res = P.res + (((int64_t)pn * P.rH + (py * P.rsy + P.ry0)) * P.rW + (px * P.rsx + P.rx0)) * P.Cout;
            }
            float* y = P.y + p * P.Cout;
            const uint32_t trow = tmem + ((uint32_t)(32 * q) << 16) + buf * (uint32_t)(Gm.G * Nt);
            const int chains = Gm.G < nslabs ? Gm.G : nslabs;
            for (int c0 = 0; c0 < Nt; c0 += 16) {
                float v[16];
                tmem_ld16(trow + (uint32_t)c0, v);
                for (int g = 1; g < chains; ++g) {
                    float u[16];
                    tmem_ld16(trow + (uint32_t)(g * Nt + c0), u);
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] += u[i];
                }
                if (ok) {
#pragma unroll
                    for (int i = 0; i < 16; i += 4) {
                        const float4 b = *reinterpret_cast<const float4*>(s_bias + c0 + i);
                        float4 a = make_float4(v[i] + b.x, v[i + 1] + b.y, v[i + 2] + b.z, v[i + 3] + b.w);
                        if (res) {
                            const float4 r = __ldg(reinterpret_cast<const float4*>(res + c0 + i));
                            a.x += r.x; a.y += r.y; a.z += r.z; a.w += r.w;
                        }
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
                        *reinterpret_cast<float4*>(y + c0 + i) = a;
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&accempty[buf]);
        }
    }
    __syncwarp();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)Gm.tmem_cols) : "memory");
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

}  // namespace

extern "C" {

int agx_internal_conv_option(const char* key, int value) {
    if (!strcmp(key, "conv_impl")) {
        if (value < 0 || value > 1) return -1;
        g_conv_impl = value;
        return 1;
    }
    return 0;
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
    if ((G.M_total + kM - 1) / kM > 0x7FFFFFFF) return 0;
    G.num_tiles = (int)((G.M_total + kM - 1) / kM);
    G.Nt = p->Cout;
    G.KS = p->Cin == 16 ? 16 : 32;
    G.swb = G.KS * 4;
    G.cblocks = p->Cin / G.KS;
    G.nslabs = taps * G.cblocks;
    G.split = p->w_lo != nullptr;
    const int K = taps * p->Cin;
    G.G = K <= 512 ? 1 : ((K <= 1024 || G.Nt > 64) ? 2 : 4);  // accumulator chains, as in agx_conv.cu
    const int cols = 2 * G.G * G.Nt;
    G.tmem_cols = cols <= 32 ? 32 : (cols <= 64 ? 64 : (cols <= 128 ? 128 : (cols <= 256 ? 256 : 512)));
    if (cols > 512) return 0;
    G.a_bytes = (uint32_t)kM * G.swb;
    G.b_bytes = (((uint32_t)G.Nt * G.swb) + 1023u) & ~1023u;
    G.stage_bytes = (G.split ? 2u : 1u) * (G.a_bytes + G.b_bytes);
    G.tx_bytes = G.a_bytes + (G.split ? 2u : 1u) * (uint32_t)G.Nt * G.swb;
    const uint32_t budget = 220u * 1024u;
    G.stages = (int)(budget / G.stage_bytes);
    if (G.stages > kMaxStages) G.stages = kMaxStages;
    if (G.stages < 2) return 0;
    const size_t smem = (size_t)G.stages * G.stage_bytes + 1024;

    const CUtensorMapSwizzle sw = G.swb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    CUtensorMap tmA, tmBhi, tmBlo;
    memset(&tmBlo, 0, sizeof(tmBlo));
    {   // activations as (C, W, H, N); the base-pixel box [-pad, dim + pad - (k - 1)) is traversed with the convolution's stride
        const cuuint64_t dims[4] = {(cuuint64_t)p->Cin, (cuuint64_t)p->W, (cuuint64_t)p->H, (cuuint64_t)p->N};
        const cuuint64_t strides[3] = {(cuuint64_t)p->Cin * 4, (cuuint64_t)p->W * p->Cin * 4, (cuuint64_t)p->H * p->W * p->Cin * 4};
        const int lower[2] = {-p->px, -p->py}, upper[2] = {p->px - (p->kw - 1), p->py - (p->kh - 1)};
        const cuuint32_t estr[4] = {1, (cuuint32_t)p->sx, (cuuint32_t)p->sy, 1};
        const CUresult r = g_encode_im2col(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(p->x), dims, strides, lower, upper, (cuuint32_t)G.KS,
                                           (cuuint32_t)kM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return 0;  // geometry the TMA unit does not take: the gather kernel handles it
    }
    for (int part = 0; part < (G.split ? 2 : 1); ++part) {
        const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)p->Cout};
        const cuuint64_t strides[1] = {(cuuint64_t)K * 4};
        const cuuint32_t box[2] = {(cuuint32_t)G.KS, (cuuint32_t)G.Nt}, estr[2] = {1, 1};
        const CUresult r = g_encode_tiled(part ? &tmBlo : &tmBhi, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(part ? p->w_lo : p->w_hi), dims, strides,
                                          box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return 0;
    }
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(agx_conv2d_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 221 * 1024 + 1024) != cudaSuccess) { cudaGetLastError(); return 0; }
        attr_set = true;
    }
    const int grid = G.num_tiles < g_sm_count ? G.num_tiles : g_sm_count;
    agx_conv2d_tma_kernel<<<grid, kThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(tmA, tmBhi, tmBlo, *p, G);
    return cudaGetLastError() == cudaSuccess ? 1 : agx_internal_fail(AGX_ERR_CUDA, "agx_conv2d_nhwc (tma): launch failed");
}

}  // extern "C"

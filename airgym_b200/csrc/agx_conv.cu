// agx_conv.cu — convolution layers of the depth-image encoders on the 5th-generation tensor cores (row f3):
// CNNFeatureExtractor (reference lib/network/cnn.py:3-33) and the frozen VAE ImgEncoder (lib/network/VAE.py:52-148), layer by
// layer over channels-last (NHWC) fp32 activations.
//
// agx_conv2d_nhwc — implicit GEMM on tcgen05.mma (kind::tf32, fp32 accumulators in TMEM):
//   M = 128 output pixels per CTA (flattened over images), N = up to 128 output channels, K = kh * kw * Cin in slabs of 32.
//   The im2col operand is never materialised in HBM: with channels-last input, 4 consecutive channels of one filter tap of one
//   output pixel ARE one 16-byte chunk of the canonical K-major operand layout, so the A tile is gathered straight from the
//   input tensor by cp.async (zero-filled outside the image), one chunk per (pixel, tap, channel quad); the weight slab
//   [Cout][32] arrives the same way.  Two-stage pipeline: the copies of slab s+1 fly under the MMAs of slab s.
//   Precision: "3xTF32" — every fp32 operand is split into hi (the 19 bits the tensor core reads) + lo (the remainder, formed in
//   shared memory for A, on the host for the weights) and each slab issues A_hi·B_hi + A_lo·B_hi + A_hi·B_lo: products carry
//   ~22 mantissa bits, i.e. fp32-level results (torch/cuDNN's default conv path on this GPU is single-pass TF32; w_lo = NULL
//   selects that).
//   Epilogue (thread = pixel = TMEM lane): + bias, + residual (VAE skip branches, possibly cropped / broadcast), activation
//   (ReLU / ELU), per-channel affine (eval-mode BatchNorm, which the CNN applies AFTER its ReLU), 16-byte NHWC stores.
// agx_conv2d_first — the Cin = 1 first layer (K = 25: not worth a GEMM) as a direct fp32 convolution on the CUDA cores, with the
//   per-pixel RunningMeanStd normalisation of the policy's image input fused into the load; writes NHWC.
// agx_resize_bilinear — F.interpolate(mode='bilinear', align_corners=False) of single-channel images (the VAE wrapper resizes the
//   212x120 camera image to its 120x212 training resolution, vae_image_encoder.py:36-38).
// agx_pool_fc — global average pool over NHWC pixels + Linear (the CNN's head).
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "agx.h"
#include "agx_tc.cuh"

int agx_internal_fail(int code, const char* msg);
extern "C" int agx_internal_conv_first_tma(const AgxConvFirstParams* p, void* stream);  // same contract, first layer
extern "C" int agx_internal_conv_tma(const AgxConvParams* p, void* stream);  // agx_conv_tma.cu: 1 launched, 0 not its geometry, < 0 error

int g_first_impl = 1;  // agx_set_option("conv_first", 0 generic direct kernel | 1 unrolled constant-bank kernel, two pixels per thread, for 5x5 / stride 2 (default) | 2 tcgen05 | 3 constant-bank kernel, one pixel per thread)

namespace {

extern __shared__ __align__(128) float c_smem[];
using namespace tc;

constexpr int kConvThreads = 256;
constexpr int kSlab = 32;           // K per pipeline stage
constexpr int kSlabChunks = kSlab / 4;

__device__ __forceinline__ void cp16z(float* dst, const float* src, bool valid) {
    const int sz = valid ? 16 : 0;  // src-size 0: the 16 bytes are zero-filled (padding taps, pixels past the end)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s32(dst)), "l"(src), "r"(sz) : "memory");
}
// D[128, N] (+)= A[128, K] * B[N, K]^T for K = 8 * ksteps, operands [kc][rows][16 B]
__device__ __forceinline__ void mma_slab(uint32_t a_base, uint32_t b_base, int N, int ksteps, uint32_t tmem_d, bool accumulate) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
    const uint32_t lboA = kM * 16, lboB = (uint32_t)N * 16;
    for (int ks = 0; ks < ksteps; ++ks) {
        const uint64_t da = smem_desc(a_base + ks * 2 * lboA, lboA, 128), db = smem_desc(b_base + ks * 2 * lboB, lboB, 128);
        const uint32_t acc = (ks > 0 || accumulate) ? 1u : 0u;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
                     "l"(da), "l"(db), "r"(idesc), "r"(acc)
                     : "memory");
    }
}

__global__ void __launch_bounds__(kConvThreads)
agx_conv2d_nhwc_kernel(const __grid_constant__ AgxConvParams P, const int Nt, const int stage_floats, const int log2cin) {
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_base;
    __shared__ int s_tap[32];  // filter tap → ky << 16 | kx (the gather's index arithmetic is shifts, masks and one table look-up per chunk)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool split = P.w_lo != nullptr;
    const int K = P.kh * P.kw * P.Cin, nslabs = (K + kSlab - 1) / kSlab;
    const int64_t M_total = (int64_t)P.N * P.Ho * P.Wo, pix0 = (int64_t)blockIdx.x * kM;
    const int n0 = blockIdx.y * Nt;
    // stage buffer: A_hi [8][128][4] | A_lo (split only) | B_hi [8][Nt][4] | B_lo (split only)
    const int offAlo = kSlabChunks * kM * 4, offBhi = split ? 2 * offAlo : offAlo, offBlo = offBhi + kSlabChunks * Nt * 4;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) s_tap[tid] = tid < P.kh * P.kw ? ((tid / P.kw) << 16) | (tid % P.kw) : 0;
    // The tensor core adds into its fp32 accumulator with truncation, so a long accumulation chain drifts by ~(chain length) x 2^-24
    // of the accumulator's magnitude (measured: 8e-6 relative at K = 800 with one chain).  G accumulators take the K-slabs round
    // robin and are added in fp32 in the epilogue: chains G times shorter over sums G times smaller.
    const int G = K <= 512 ? 1 : (K <= 1024 || Nt > 64 ? 2 : 4);  // K = 144 / 288: chains of 60-110 accumulations stay below 4e-6
    const uint32_t tmem_cols = (uint32_t)(G * Nt) <= 32u ? 32u : ((uint32_t)(G * Nt) <= 64u ? 64u : ((uint32_t)(G * Nt) <= 128u ? 128u : 256u));
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_base)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // this thread gathers pixel r, K-chunks j0..j0+3 of every slab
    const int r = tid & (kM - 1), j0 = (tid >> 7) * 4;
    const int64_t pix = pix0 + r;
    const bool pix_ok = pix < M_total;
    int n_img = 0, oy = 0, ox = 0;
    if (pix_ok) { n_img = (int)(pix / ((int64_t)P.Ho * P.Wo)); const int rem = (int)(pix - (int64_t)n_img * P.Ho * P.Wo); oy = rem / P.Wo; ox = rem - oy * P.Wo; }
    const int iy0 = oy * P.sy - P.py, ix0 = ox * P.sx - P.px;
    const float* x_img = P.x + (int64_t)n_img * P.H * P.W * P.Cin;

    auto issue = [&](int s) {
        float* buf = c_smem + (s & 1) * stage_floats;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + u, k = s * kSlab + j * 4;
            const int tap = log2cin >= 0 ? (k >> log2cin) : 0, c = log2cin >= 0 ? (k & (P.Cin - 1)) : k;  // 1x1 layers: tap 0, c = k
            const int t = s_tap[tap & 31], ky = t >> 16, kx = t & 0xFFFF;
            const int iy = iy0 + ky, ix = ix0 + kx;
            const bool ok = pix_ok && k < K && iy >= 0 && iy < P.H && ix >= 0 && ix < P.W;
            cp16z(buf + (j * kM + r) * 4, ok ? x_img + ((int64_t)iy * P.W + ix) * P.Cin + c : P.x, ok);
        }
        for (int i = tid; i < Nt * kSlabChunks; i += kConvThreads) {  // weight slab: consecutive threads = consecutive chunks of one row
            const int j = i & (kSlabChunks - 1), n = i >> 3, k = s * kSlab + j * 4;
            const bool ok = k < K;
            const int64_t off = (int64_t)(n0 + n) * K + k;
            cp16z(buf + offBhi + (j * Nt + n) * 4, ok ? P.w_hi + off : P.w_hi, ok);
            if (split) cp16z(buf + offBlo + (j * Nt + n) * 4, ok ? P.w_lo + off : P.w_hi, ok);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    publish_and_sync(0, kConvThreads);
    const uint32_t tmem = tmem_base;
    issue(0);
    if (nslabs > 1) issue(1);
    uint32_t phase[2] = {0u, 0u};
    for (int s = 0; s < nslabs; ++s) {
        if (s + 1 < nslabs) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        float* buf = c_smem + (s & 1) * stage_floats;
        if (split) {  // lo = a - hi for the chunks this thread gathered (its own copies have landed)
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int o = ((j0 + u) * kM + r) * 4;
                const float4 a = *reinterpret_cast<const float4*>(buf + o);
                float4 l;
                l.x = a.x - __uint_as_float(__float_as_uint(a.x) & 0xFFFFE000u);
                l.y = a.y - __uint_as_float(__float_as_uint(a.y) & 0xFFFFE000u);
                l.z = a.z - __uint_as_float(__float_as_uint(a.z) & 0xFFFFE000u);
                l.w = a.w - __uint_as_float(__float_as_uint(a.w) & 0xFFFFE000u);
                *reinterpret_cast<float4*>(buf + offAlo + o) = l;
            }
        }
        publish_and_sync(0, kConvThreads);
        const int b = s & 1;
        if (tid == 0) {
            const int kk = K - s * kSlab, ksteps = (kk < kSlab ? kk : kSlab) / 8;
            const uint32_t base = s32(buf);
            const uint32_t d = tmem + (uint32_t)((s % G) * Nt);
            mma_slab(base, base + 4 * offBhi, Nt, ksteps, d, s >= G);
            if (split) {
                mma_slab(base + 4 * offAlo, base + 4 * offBhi, Nt, ksteps, d, true);
                mma_slab(base, base + 4 * offBlo, Nt, ksteps, d, true);
            }
            commit(&bars[b]);
        }
        if (s + 2 < nslabs) {
            wait(&bars[b], phase[b]); phase[b] ^= 1u;
            issue(s + 2);
        }
    }
    {
        const int b = (nslabs - 1) & 1;
        wait(&bars[b], phase[b]);
    }
    // ---- epilogue: thread = pixel (TMEM lane); the two warps of a lane quarter split the columns
    {
        const int q = warp & 3, half = warp >> 2, rr = 32 * q + lane;
        const int64_t p = pix0 + rr;
        const bool ok = p < M_total;
        int pn = 0, py = 0, px = 0;
        if (ok) { pn = (int)(p / ((int64_t)P.Ho * P.Wo)); const int rem = (int)(p - (int64_t)pn * P.Ho * P.Wo); py = rem / P.Wo; px = rem - py * P.Wo; }
        const float* res = nullptr;
        if (P.res && ok) res = P.res + (((int64_t)pn * P.rH + (py * P.rsy + P.ry0)) * P.rW + (px * P.rsx + P.rx0)) * P.Cout;
        float* y = P.y + p * P.Cout;
        const uint32_t trow = tmem + ((uint32_t)(32 * q) << 16);
        const int cols = Nt / 2;
        for (int c0 = half * cols; c0 < (half + 1) * cols; c0 += 16) {
            float v[16];
            tmem_ld16(trow + (uint32_t)c0, v);
            for (int g = 1; g < G && g < nslabs; ++g) {
                float u[16];
                tmem_ld16(trow + (uint32_t)(g * Nt + c0), u);
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] += u[i];
            }
            if (ok) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int c = n0 + c0 + i;
                    float a = v[i] + (P.bias ? __ldg(P.bias + c) : 0.0f);
                    if (res) a += __ldg(res + c);
                    if (P.act == 1) a = fmaxf(a, 0.0f);
                    else if (P.act == 2) a = a > 0.0f ? a : expm1f(a);
                    if (P.scale) a = a * __ldg(P.scale + c) + __ldg(P.shift + c);
                    v[i] = a;
                }
#pragma unroll
                for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(y + n0 + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
}

// ---- first layer: Cin = 1, direct convolution, one thread per output pixel, COUT channels in registers (a 4-pixel strip per thread
// with compile-time taps was measured slower: 0.92 vs 0.80 ms per 2048 images, 122-254 registers) ---------------------------
template <int COUT>
__global__ void __launch_bounds__(128)
agx_conv2d_first_kernel(const __grid_constant__ AgxConvFirstParams P) {
    __shared__ float s_w[COUT * 32];  // [tap][COUT]
    __shared__ float s_b[COUT], s_s[COUT], s_t[COUT];
    const int taps = P.kh * P.kw;
    for (int i = threadIdx.x; i < taps * COUT; i += 128) { const int t = i / COUT, c = i - t * COUT; s_w[i] = P.w[c * taps + t]; }
    for (int i = threadIdx.x; i < COUT; i += 128) { s_b[i] = P.bias ? P.bias[i] : 0.0f; s_s[i] = P.scale ? P.scale[i] : 1.0f; s_t[i] = P.shift ? P.shift[i] : 0.0f; }
    __syncthreads();
    const int64_t M_total = (int64_t)P.N * P.Ho * P.Wo;
    for (int64_t p = (int64_t)blockIdx.x * 128 + threadIdx.x; p < M_total; p += (int64_t)gridDim.x * 128) {
        const int n = (int)(p / ((int64_t)P.Ho * P.Wo)), rem = (int)(p - (int64_t)n * P.Ho * P.Wo), oy = rem / P.Wo, ox = rem - oy * P.Wo;
        const float* img = P.x + (int64_t)n * P.H * P.W;
        float acc[COUT];
#pragma unroll
        for (int c = 0; c < COUT; ++c) acc[c] = s_b[c];
        for (int ky = 0; ky < P.kh; ++ky) {
            const int iy = oy * P.sy - P.py + ky;
            if (iy < 0 || iy >= P.H) continue;
            for (int kx = 0; kx < P.kw; ++kx) {
                const int ix = ox * P.sx - P.px + kx;
                if (ix < 0 || ix >= P.W) continue;
                float v = __ldg(img + iy * P.W + ix);
                if (P.px_mean) {  // RunningMeanStd forward: clamp((x - mean) / sqrt(var + eps), +-5), rstd prepared by the caller
                    v = (v - __ldg(P.px_mean + iy * P.W + ix)) * __ldg(P.px_rstd + iy * P.W + ix);
                    v = fminf(fmaxf(v, -5.0f), 5.0f);
                }
                const float* wt = s_w + (ky * P.kw + kx) * COUT;
#pragma unroll
                for (int c = 0; c < COUT; ++c) acc[c] = fmaf(v, wt[c], acc[c]);
            }
        }
        float* y = P.y + p * COUT;
#pragma unroll
        for (int c = 0; c < COUT; c += 4) {
            float o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float a = acc[c + i];
                if (P.act == 1) a = fmaxf(a, 0.0f);
                else if (P.act == 2) a = a > 0.0f ? a : expm1f(a);
                o[i] = a * s_s[c + i] + s_t[c + i];
            }
            *reinterpret_cast<float4*>(y + c) = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
}

// ---- first layer, 5x5 / stride 2 (both encoders): fully unrolled direct convolution with the weights in the constant bank ---------
// The generic kernel above spends ~700 instructions per output pixel at IPC 1.2: run-time tap loops with bounds branches, one
// dependent load per tap, four shared-memory weight reads per 16 FMAs.  Here the 25 taps are loaded up front (predicated, zero
// outside the image: the convolution pads the NORMALISED image), and every FMA takes its weight as a constant-bank operand
// (c[bank][imm] after unrolling: no load instruction, no register), so a pixel costs 25 loads + 25 * COUT FMAs + the epilogue.
// Same tap order and fmaf chain as the generic kernel: results are bit-identical to it.  The weights reach the constant bank by a
// stream-ordered device-to-device copy (cudaMemcpyToSymbolAsync; a memcpy node under graph capture), one slot per Cout.
__constant__ float c_first[2][32 * 25 + 3 * 32];  // slot 0: Cout = 16, slot 1: Cout = 32 — [tap][Cout] weights | bias | scale | shift
__device__ float g_first_pack[2][32 * 25 + 3 * 32];  // staging of the constant-bank image (a static buffer: no allocation inside a stream capture)

__global__ void agx_first_pack_kernel(const float* __restrict__ w, const float* __restrict__ bias, const float* __restrict__ scale, const float* __restrict__ shift,
                                      float* __restrict__ out, int cout) {
    for (int i = threadIdx.x; i < 25 * cout; i += blockDim.x) { const int t = i / cout, c = i - t * cout; out[i] = w[c * 25 + t]; }
    for (int i = threadIdx.x; i < cout; i += blockDim.x) {
        out[25 * cout + i] = bias ? bias[i] : 0.0f;
        out[26 * cout + i] = scale ? scale[i] : 1.0f;
        out[27 * cout + i] = shift ? shift[i] : 0.0f;
    }
}

template <int COUT, bool NORM>
__global__ void __launch_bounds__(128)
agx_conv2d_first5_kernel(const __grid_constant__ AgxConvFirstParams P) {
    constexpr int SLOT = COUT == 16 ? 0 : 1;
    const int64_t M_total = (int64_t)P.N * P.Ho * P.Wo;
    const int HoWo = P.Ho * P.Wo;
    for (int64_t p = (int64_t)blockIdx.x * 128 + threadIdx.x; p < M_total; p += (int64_t)gridDim.x * 128) {
        const int n = (int)(p / HoWo), rem = (int)(p - (int64_t)n * HoWo), oy = rem / P.Wo, ox = rem - oy * P.Wo;
        const float* img = P.x + (int64_t)n * P.H * P.W;
        const int iy0 = oy * 2 - P.py, ix0 = ox * 2 - P.px;
        float v[25];
#pragma unroll
        for (int ky = 0; ky < 5; ++ky) {
            const int iy = iy0 + ky;
            const bool rok = (unsigned)iy < (unsigned)P.H;
#pragma unroll
            for (int kx = 0; kx < 5; ++kx) {
                const int ix = ix0 + kx;
                const bool ok = rok && (unsigned)ix < (unsigned)P.W;
                const int off = ok ? iy * P.W + ix : 0;
                float t = __ldg(img + off);
                if (NORM) t = fminf(fmaxf((t - __ldg(P.px_mean + off)) * __ldg(P.px_rstd + off), -5.0f), 5.0f);
                v[ky * 5 + kx] = ok ? t : 0.0f;
            }
        }
        float acc[COUT];
#pragma unroll
        for (int c = 0; c < COUT; ++c) acc[c] = c_first[SLOT][25 * COUT + c];
#pragma unroll
        for (int t = 0; t < 25; ++t)
#pragma unroll
            for (int c = 0; c < COUT; ++c) acc[c] = fmaf(v[t], c_first[SLOT][t * COUT + c], acc[c]);
        float* y = P.y + p * COUT;
#pragma unroll
        for (int c = 0; c < COUT; c += 4) {
            float o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float a = acc[c + i];
                if (P.act == 1) a = fmaxf(a, 0.0f);
                else if (P.act == 2) a = a > 0.0f ? a : expm1f(a);
                o[i] = a * c_first[SLOT][26 * COUT + c + i] + c_first[SLOT][27 * COUT + c + i];
            }
            *reinterpret_cast<float4*>(y + c) = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
}

// two horizontally adjacent output pixels per thread, one output row (segment of 32 pixel pairs) per warp: each lane loads ONE aligned
// float4 per input row (a warp reads the row as a single coalesced 512-byte request; ncu of the version with four overlapping 8-byte
// loads per row: L1 70 % busy at 4 wavefronts per load) and takes the two columns to its left and the one to its right from its
// neighbours by shuffle (the segment's edge lanes fetch theirs); every uniform weight load feeds two FMAs, and a thread stores
// 2 * COUT contiguous floats.  Same per-pixel fmaf chain as above.  Needs W % 4 == 0, Wo even, pad 2.
template <int COUT, bool NORM>
__global__ void __launch_bounds__(128)
agx_conv2d_first5x2_kernel(const __grid_constant__ AgxConvFirstParams P) {
    constexpr int SLOT = COUT == 16 ? 0 : 1;
    const int Wp = P.Wo >> 1, segs = (Wp + 31) >> 5, lane = threadIdx.x & 31;
    const int64_t items = (int64_t)P.N * P.Ho * segs;
    auto norm1 = [&](float t, int off) { return NORM ? fminf(fmaxf((t - __ldg(P.px_mean + off)) * __ldg(P.px_rstd + off), -5.0f), 5.0f) : t; };
    for (int64_t it = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5); it < items; it += (int64_t)gridDim.x * 4) {
        const int64_t row = it / segs;  // n * Ho + oy
        const int sg = (int)(it - row * segs), j = sg * 32 + lane, n = (int)(row / P.Ho), oy = (int)(row - (int64_t)n * P.Ho);
        const float* img = P.x + (int64_t)n * P.H * P.W;
        const int iy0 = oy * 2 - P.py, xa = 4 * j;  // this lane's aligned columns xa .. xa+3; its pixel pair reads columns xa-2 .. xa+4
        float v[5][7];
#pragma unroll
        for (int ky = 0; ky < 5; ++ky) {
            const int iy = iy0 + ky;
            const bool rok = (unsigned)iy < (unsigned)P.H, ok = rok && xa < P.W;
            const int off = ok ? iy * P.W + xa : 0;
            float4 t = __ldg(reinterpret_cast<const float4*>(img + off));
            if (NORM) {
                const float4 m = __ldg(reinterpret_cast<const float4*>(P.px_mean + off)), r = __ldg(reinterpret_cast<const float4*>(P.px_rstd + off));
                t.x = fminf(fmaxf((t.x - m.x) * r.x, -5.0f), 5.0f); t.y = fminf(fmaxf((t.y - m.y) * r.y, -5.0f), 5.0f);
                t.z = fminf(fmaxf((t.z - m.z) * r.z, -5.0f), 5.0f); t.w = fminf(fmaxf((t.w - m.w) * r.w, -5.0f), 5.0f);
            }
            if (!ok) t = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            float l0 = __shfl_up_sync(0xFFFFFFFFu, t.z, 1), l1 = __shfl_up_sync(0xFFFFFFFFu, t.w, 1), r0 = __shfl_down_sync(0xFFFFFFFFu, t.x, 1);
            if (lane == 0) {  // left neighbour lives in the previous segment (or is the zero padding)
                const bool lok = rok && sg > 0;
                const int lo = lok ? iy * P.W + xa - 2 : 0;
                l0 = lok ? norm1(__ldg(img + lo), lo) : 0.0f;
                l1 = lok ? norm1(__ldg(img + lo + 1), lo + 1) : 0.0f;
            }
            if (lane == 31) {
                const bool rk = rok && xa + 4 < P.W;
                const int ro = rk ? iy * P.W + xa + 4 : 0;
                r0 = rk ? norm1(__ldg(img + ro), ro) : 0.0f;
            }
            v[ky][0] = l0; v[ky][1] = l1; v[ky][2] = t.x; v[ky][3] = t.y; v[ky][4] = t.z; v[ky][5] = t.w; v[ky][6] = r0;
        }
        if (j >= Wp) continue;  // idle lanes of the row's last segment took part in the shuffles only
        float a0[COUT], a1[COUT];
#pragma unroll
        for (int c = 0; c < COUT; ++c) a0[c] = a1[c] = c_first[SLOT][25 * COUT + c];
#pragma unroll
        for (int ky = 0; ky < 5; ++ky)
#pragma unroll
            for (int kx = 0; kx < 5; ++kx)
#pragma unroll
                for (int c = 0; c < COUT; ++c) {
                    const float w = c_first[SLOT][(ky * 5 + kx) * COUT + c];
                    a0[c] = fmaf(v[ky][kx], w, a0[c]);
                    a1[c] = fmaf(v[ky][kx + 2], w, a1[c]);
                }
        float* y = P.y + (row * P.Wo + 2 * j) * COUT;
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int c = 0; c < COUT; c += 4) {
                float o[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float a = h ? a1[c + i] : a0[c + i];
                    if (P.act == 1) a = fmaxf(a, 0.0f);
                    else if (P.act == 2) a = a > 0.0f ? a : expm1f(a);
                    o[i] = a * c_first[SLOT][26 * COUT + c + i] + c_first[SLOT][27 * COUT + c + i];
                }
                *reinterpret_cast<float4*>(y + h * COUT + c) = make_float4(o[0], o[1], o[2], o[3]);
            }
    }
}

// flattened variant (pixel pairs numbered across rows, no shuffles): keeps every lane busy when a row's pair count is far from a multiple
// of 32 (the VAE's 53 pairs per row), at four overlapping 8-byte loads per input row — used without the fused normalisation only.
// Two horizontally adjacent output pixels per thread: the 5 x 7 input patch is shared (20 8-byte loads instead of 50 scalar ones), every
// uniform weight load feeds two FMAs, and a thread stores 2 * COUT contiguous floats.  Same per-pixel fmaf chain as above.
template <int COUT, bool NORM>
__global__ void __launch_bounds__(128)
agx_conv2d_first5x2_flat_kernel(const __grid_constant__ AgxConvFirstParams P) {
    constexpr int SLOT = COUT == 16 ? 0 : 1;
    const int Wp = P.Wo >> 1;
    const int64_t pairs = (int64_t)P.N * P.Ho * Wp;
    for (int64_t q = (int64_t)blockIdx.x * 128 + threadIdx.x; q < pairs; q += (int64_t)gridDim.x * 128) {
        const int64_t row = q / Wp;  // n * Ho + oy
        const int j = (int)(q - row * Wp), n = (int)(row / P.Ho), oy = (int)(row - (int64_t)n * P.Ho);
        const float* img = P.x + (int64_t)n * P.H * P.W;
        const int iy0 = oy * 2 - P.py, ix0 = j * 4 - P.px;  // even: W and px are even, so a column pair is inside the image or outside as a whole
        float v[5][7];
#pragma unroll
        for (int ky = 0; ky < 5; ++ky) {
            const int iy = iy0 + ky;
            const bool rok = (unsigned)iy < (unsigned)P.H;
#pragma unroll
            for (int c2 = 0; c2 < 4; ++c2) {
                const int ix = ix0 + 2 * c2;
                const bool ok = rok && (unsigned)ix < (unsigned)P.W;
                const int off = ok ? iy * P.W + ix : 0;
                if (c2 < 3) {
                    float2 t = __ldg(reinterpret_cast<const float2*>(img + off));
                    if (NORM) {
                        const float2 m = __ldg(reinterpret_cast<const float2*>(P.px_mean + off)), r = __ldg(reinterpret_cast<const float2*>(P.px_rstd + off));
                        t.x = fminf(fmaxf((t.x - m.x) * r.x, -5.0f), 5.0f);
                        t.y = fminf(fmaxf((t.y - m.y) * r.y, -5.0f), 5.0f);
                    }
                    v[ky][2 * c2] = ok ? t.x : 0.0f;
                    v[ky][2 * c2 + 1] = ok ? t.y : 0.0f;
                } else {
                    float t = __ldg(img + off);
                    if (NORM) t = fminf(fmaxf((t - __ldg(P.px_mean + off)) * __ldg(P.px_rstd + off), -5.0f), 5.0f);
                    v[ky][6] = ok ? t : 0.0f;
                }
            }
        }
        float a0[COUT], a1[COUT];
#pragma unroll
        for (int c = 0; c < COUT; ++c) a0[c] = a1[c] = c_first[SLOT][25 * COUT + c];
#pragma unroll
        for (int ky = 0; ky < 5; ++ky)
#pragma unroll
            for (int kx = 0; kx < 5; ++kx)
#pragma unroll
                for (int c = 0; c < COUT; ++c) {
                    const float w = c_first[SLOT][(ky * 5 + kx) * COUT + c];
                    a0[c] = fmaf(v[ky][kx], w, a0[c]);
                    a1[c] = fmaf(v[ky][kx + 2], w, a1[c]);
                }
        float* y = P.y + (row * P.Wo + 2 * j) * COUT;
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int c = 0; c < COUT; c += 4) {
                float o[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float a = h ? a1[c + i] : a0[c + i];
                    if (P.act == 1) a = fmaxf(a, 0.0f);
                    else if (P.act == 2) a = a > 0.0f ? a : expm1f(a);
                    o[i] = a * c_first[SLOT][26 * COUT + c + i] + c_first[SLOT][27 * COUT + c + i];
                }
                *reinterpret_cast<float4*>(y + h * COUT + c) = make_float4(o[0], o[1], o[2], o[3]);
            }
    }
}

__global__ void agx_resize_bilinear_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, int H, int W, int Ho, int Wo) {
    const int64_t total = n * Ho * Wo;
    const float sy = (float)H / (float)Ho, sx = (float)W / (float)Wo;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int img = (int)(i / ((int64_t)Ho * Wo)), rem = (int)(i - (int64_t)img * Ho * Wo), oy = rem / Wo, ox = rem - oy * Wo;
        float fy = sy * ((float)oy + 0.5f) - 0.5f, fx = sx * ((float)ox + 0.5f) - 0.5f;  // area_pixel_compute_source_index, align_corners = False
        fy = fy < 0.0f ? 0.0f : fy; fx = fx < 0.0f ? 0.0f : fx;
        const int y0 = (int)fy, x0 = (int)fx, y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
        const float ly = fy - (float)y0, lx = fx - (float)x0;
        const float* s = x + (int64_t)img * H * W;
        y[i] = (1.0f - ly) * ((1.0f - lx) * s[y0 * W + x0] + lx * s[y0 * W + x1]) + ly * ((1.0f - lx) * s[y1 * W + x0] + lx * s[y1 * W + x1]);
    }
}

// global average pool over `pixels` NHWC pixels of C (<= 128) channels + Linear C -> F (<= 64): one CTA per image
__global__ void __launch_bounds__(128)
agx_pool_fc_kernel(const float* __restrict__ x, int pixels, int C, const float* __restrict__ wfc, const float* __restrict__ bfc, int F,
                   float* __restrict__ out, int64_t ld_out) {
    __shared__ float s_part[4][128];
    __shared__ float s_mean[128];
    const int64_t img = blockIdx.x;
    const float* xi = x + img * (int64_t)pixels * C;
    // thread = (pixel group g, channel c): C channels x (128 / C) pixel groups, coalesced over c
    const int groups = 128 / C, c = threadIdx.x % C, g = threadIdx.x / C;
    float s = 0.0f;
    if (g < groups)
        for (int p = g; p < pixels; p += groups) s += xi[(int64_t)p * C + c];
    if (g < 4 && g < groups) s_part[g][c] = s;
    __syncthreads();
    if (threadIdx.x < C) {
        float t = 0.0f;
        for (int k = 0; k < (groups < 4 ? groups : 4); ++k) t += s_part[k][threadIdx.x];
        s_mean[threadIdx.x] = t / (float)pixels;
    }
    __syncthreads();
    if (threadIdx.x < F) {
        float a = bfc[threadIdx.x];
        for (int k = 0; k < C; ++k) a = fmaf(wfc[threadIdx.x * C + k], s_mean[k], a);
        out[img * ld_out + threadIdx.x] = a;
    }
}


// ---- train-mode BatchNorm2d over a channels-last activation x [rows, C] (reference lib/network/cnn.py:9-21 in train mode: Conv → ReLU →
// BatchNorm with BATCH statistics).  `sums` [2, C] float64 = per-channel sums and sums of squares over the rows (agx_col_sums: deterministic,
// all-reducible across ranks); every CTA derives scale = gamma / sqrt(biased var + eps), shift = beta - mean * scale and rewrites its share of
// x in place; CTA 0 also moves the running statistics: running = (1 - momentum) * running + momentum * batch (variance UNBIASED, as torch).
__global__ void __launch_bounds__(256)
agx_bn_train_kernel(float* __restrict__ x, int64_t rows, int C, const double* __restrict__ sums, const float* __restrict__ gamma,
                    const float* __restrict__ beta, float eps, float momentum, float* running_mean, float* running_var) {
    __shared__ float s_s[128], s_t[128];
    if (threadIdx.x < C) {
        const int c = threadIdx.x;
        const double n = (double)rows, mean = sums[c] / n;
        double var = sums[C + c] / n - mean * mean;
        var = var > 0.0 ? var : 0.0;
        const float sc = (gamma ? gamma[c] : 1.0f) / sqrtf((float)var + eps);
        s_s[c] = sc;
        s_t[c] = (beta ? beta[c] : 0.0f) - (float)mean * sc;
        if (blockIdx.x == 0 && running_mean) {
            const double unb = rows > 1 ? var * n / (n - 1.0) : var;
            running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * (float)mean;
            running_var[c] = (1.0f - momentum) * running_var[c] + momentum * (float)unb;
        }
    }
    __syncthreads();
    const int64_t total4 = rows * C / 4;  // C % 4 == 0
    float4* x4 = reinterpret_cast<float4*>(x);
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total4; i += (int64_t)gridDim.x * 256) {
        const int c = (int)((i * 4) % C);
        float4 v = x4[i];
        v.x = v.x * s_s[c] + s_t[c]; v.y = v.y * s_s[c + 1] + s_t[c + 1]; v.z = v.z * s_s[c + 2] + s_t[c + 2]; v.w = v.w * s_s[c + 3] + s_t[c + 3];
        x4[i] = v;
    }
}

}  // namespace

extern "C" {

int agx_sizeof_conv_params(void) { return (int)sizeof(AgxConvParams); }
int agx_sizeof_conv_first_params(void) { return (int)sizeof(AgxConvFirstParams); }

int agx_conv2d_nhwc(const AgxConvParams* p, void* stream) {
    if (!p || !p->x || !p->w_hi || !p->y || p->N <= 0 || p->H <= 0 || p->W <= 0 || p->Ho <= 0 || p->Wo <= 0 || p->kh <= 0 || p->kw <= 0 ||
        p->sy <= 0 || p->sx <= 0 || ((p->scale == nullptr) != (p->shift == nullptr)) || p->act < 0 || p->act > 2)
        return agx_internal_fail(AGX_ERR_ARG, "agx_conv2d_nhwc: bad argument");
    if ((p->Cin & 3) || p->Cin <= 0 || (p->Cout & 31) || p->Cout <= 0 || ((p->kh * p->kw * p->Cin) & 7))
        return agx_internal_fail(AGX_ERR_UNSUPPORTED, "agx_conv2d_nhwc: needs Cin % 4 == 0, Cout % 32 == 0, (kh*kw*Cin) % 8 == 0");
    const int Nt = p->Cout <= 128 ? p->Cout : 128;
    if (p->Cout % Nt) return agx_internal_fail(AGX_ERR_UNSUPPORTED, "agx_conv2d_nhwc: Cout above 128 must be a multiple of 128");
    const uintptr_t al = (uintptr_t)p->x | (uintptr_t)p->w_hi | (uintptr_t)p->w_lo | (uintptr_t)p->y | (uintptr_t)p->res;
    if (al & 15u) return agx_internal_fail(AGX_ERR_ALIGN, "agx_conv2d_nhwc: buffers must be 16-byte aligned");
    if (p->res && (p->rH <= 0 || p->rW <= 0 || (p->Ho - 1) * p->rsy + p->ry0 >= p->rH || (p->Wo - 1) * p->rsx + p->rx0 >= p->rW || p->ry0 < 0 || p->rx0 < 0))
        return agx_internal_fail(AGX_ERR_ARG, "agx_conv2d_nhwc: residual window outside the residual tensor");
    {   // spatial layers go to the TMA im2col kernel (agx_conv_tma.cu); the dense layers and odd geometries stay here
        const int r = agx_internal_conv_tma(p, stream);
        if (r) return r > 0 ? AGX_OK : r;
    }
    const bool split = p->w_lo != nullptr;
    const int stage_floats = (split ? 2 : 1) * (kSlabChunks * kM * 4 + kSlabChunks * Nt * 4);
    const size_t smem = 2 * (size_t)stage_floats * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) { cudaFuncSetAttribute(agx_conv2d_nhwc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 2 * (8 * 128 * 4 + 8 * 128 * 4) * 4); attr_set = true; }
    const int64_t M_total = (int64_t)p->N * p->Ho * p->Wo;
    const dim3 grid((unsigned)((M_total + kM - 1) / kM), (unsigned)(p->Cout / Nt));
    int log2cin = -1;
    if (p->kh * p->kw > 1) {
        if (p->kh * p->kw > 32 || (p->Cin & (p->Cin - 1))) return agx_internal_fail(AGX_ERR_UNSUPPORTED, "agx_conv2d_nhwc: spatial kernels need Cin a power of two and at most 32 taps");
        for (log2cin = 0; (1 << log2cin) < p->Cin; ++log2cin) {}
    }
    agx_conv2d_nhwc_kernel<<<grid, kConvThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(*p, Nt, stage_floats, log2cin);
    return cudaGetLastError() == cudaSuccess ? AGX_OK : agx_internal_fail(AGX_ERR_CUDA, "agx_conv2d_nhwc: launch failed");
}

int agx_conv2d_first(const AgxConvFirstParams* p, void* stream) {
    if (!p || !p->x || !p->w || !p->y || p->N <= 0 || p->kh * p->kw > 32 || p->kh <= 0 || p->kw <= 0 || ((p->px_mean == nullptr) != (p->px_rstd == nullptr)) ||
        p->act < 0 || p->act > 2 || p->sy <= 0 || p->sx <= 0)
        return agx_internal_fail(AGX_ERR_ARG, "agx_conv2d_first: bad argument");
    if ((uintptr_t)p->y & 15u) return agx_internal_fail(AGX_ERR_ALIGN, "agx_conv2d_first: output must be 16-byte aligned");
    const int64_t M_total = (int64_t)p->N * p->Ho * p->Wo;
    int64_t grid = (M_total + 127) / 128;
    if (grid > 148 * 16) grid = 148 * 16;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool five = p->kh == 5 && p->kw == 5 && p->sy == 2 && p->sx == 2 && (p->Cout == 16 || p->Cout == 32);
    if (five && g_first_impl == 2) {  // operand rows built from a TMA strip, tcgen05 (agx_conv_tma.cu): measured no faster than the generic kernel
        const int r = agx_internal_conv_first_tma(p, stream);
        if (r) return r > 0 ? AGX_OK : r;
    }
    if (five && g_first_impl >= 1) {
        const int slot = p->Cout == 16 ? 0 : 1, nfl = 28 * p->Cout;
        static float* pack_base = nullptr;  // device address of g_first_pack (stream order serialises the reuse of a slot)
        if (!pack_base && cudaGetSymbolAddress(reinterpret_cast<void**>(&pack_base), g_first_pack) != cudaSuccess)
            return agx_internal_fail(AGX_ERR_CUDA, "agx_conv2d_first: cudaGetSymbolAddress failed");
        float* pack = pack_base + slot * (32 * 25 + 3 * 32);
        agx_first_pack_kernel<<<1, 128, 0, st>>>(p->w, p->bias, p->scale, p->shift, pack, p->Cout);
        if (cudaMemcpyToSymbolAsync(c_first, pack, sizeof(float) * nfl, sizeof(float) * slot * (32 * 25 + 3 * 32), cudaMemcpyDeviceToDevice, st) != cudaSuccess)
            return agx_internal_fail(AGX_ERR_CUDA, "agx_conv2d_first: constant-bank copy failed");
        const bool pair_ok = !(p->Wo & 1) && !(p->W & 3) && p->px == 2 && !(((uintptr_t)p->x | (uintptr_t)p->px_mean | (uintptr_t)p->px_rstd) & 15u);
        if (pair_ok && g_first_impl == 1) {
            const int Wp = p->Wo / 2, segs = (Wp + 31) / 32;
            if (!p->px_mean && Wp * 10 < segs * 32 * 9) {  // rows fill their 32-lane segments below 90 %: flattened pairs
                int64_t g2 = (M_total / 2 + 127) / 128;
                if (g2 > 148 * 16) g2 = 148 * 16;
                if (p->Cout == 16) agx_conv2d_first5x2_flat_kernel<16, false><<<(unsigned)g2, 128, 0, st>>>(*p);
                else agx_conv2d_first5x2_flat_kernel<32, false><<<(unsigned)g2, 128, 0, st>>>(*p);
                return cudaGetLastError() == cudaSuccess ? AGX_OK : agx_internal_fail(AGX_ERR_CUDA, "agx_conv2d_first: launch failed");
            }
            int64_t g2 = ((int64_t)p->N * p->Ho * segs + 3) / 4;  // one (row, 32-pair segment) per warp, 4 warps per block
            if (g2 > 148 * 16) g2 = 148 * 16;
            const unsigned g = (unsigned)g2;
            if (p->Cout == 16) { if (p->px_mean) agx_conv2d_first5x2_kernel<16, true><<<g, 128, 0, st>>>(*p); else agx_conv2d_first5x2_kernel<16, false><<<g, 128, 0, st>>>(*p); }
            else { if (p->px_mean) agx_conv2d_first5x2_kernel<32, true><<<g, 128, 0, st>>>(*p); else agx_conv2d_first5x2_kernel<32, false><<<g, 128, 0, st>>>(*p); }
            return cudaGetLastError() == cudaSuccess ? AGX_OK : agx_internal_fail(AGX_ERR_CUDA, "agx_conv2d_first: launch failed");
        }
        const unsigned g = (unsigned)grid;
        if (p->Cout == 16) { if (p->px_mean) agx_conv2d_first5_kernel<16, true><<<g, 128, 0, st>>>(*p); else agx_conv2d_first5_kernel<16, false><<<g, 128, 0, st>>>(*p); }
        else { if (p->px_mean) agx_conv2d_first5_kernel<32, true><<<g, 128, 0, st>>>(*p); else agx_conv2d_first5_kernel<32, false><<<g, 128, 0, st>>>(*p); }
        return cudaGetLastError() == cudaSuccess ? AGX_OK : agx_internal_fail(AGX_ERR_CUDA, "agx_conv2d_first: launch failed");
    }
    if (p->Cout == 16) agx_conv2d_first_kernel<16><<<(unsigned)grid, 128, 0, st>>>(*p);
    else if (p->Cout == 32) agx_conv2d_first_kernel<32><<<(unsigned)grid, 128, 0, st>>>(*p);
    else return agx_internal_fail(AGX_ERR_UNSUPPORTED, "agx_conv2d_first: Cout must be 16 or 32");
    return cudaGetLastError() == cudaSuccess ? AGX_OK : agx_internal_fail(AGX_ERR_CUDA, "agx_conv2d_first: launch failed");
}

int agx_resize_bilinear(const float* x, float* y, int64_t n, int H, int W, int Ho, int Wo, void* stream) {
    if (!x || !y || n <= 0 || H <= 0 || W <= 0 || Ho <= 0 || Wo <= 0) return agx_internal_fail(AGX_ERR_ARG, "agx_resize_bilinear: bad argument");
    int64_t grid = (n * Ho * Wo + 255) / 256;
    if (grid > 148 * 32) grid = 148 * 32;
    agx_resize_bilinear_kernel<<<(unsigned)grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, y, n, H, W, Ho, Wo);
    return cudaGetLastError() == cudaSuccess ? AGX_OK : agx_internal_fail(AGX_ERR_CUDA, "agx_resize_bilinear: launch failed");
}

int agx_bn_train(float* x, int64_t rows, int C, const double* sums, const float* gamma, const float* beta, float eps, float momentum,
                 float* running_mean, float* running_var, void* stream) {
    if (!x || !sums || rows <= 0 || C <= 0 || C > 128 || (C & 3) || ((running_mean == nullptr) != (running_var == nullptr)) || ((uintptr_t)x & 15u))
        return agx_internal_fail(AGX_ERR_ARG, "agx_bn_train: bad argument (C % 4 == 0, C <= 128, x 16-byte aligned)");
    int64_t grid = (rows * C / 4 + 255) / 256;
    if (grid > 148 * 16) grid = 148 * 16;
    agx_bn_train_kernel<<<(unsigned)grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, rows, C, sums, gamma, beta, eps, momentum, running_mean, running_var);
    return cudaGetLastError() == cudaSuccess ? AGX_OK : agx_internal_fail(AGX_ERR_CUDA, "agx_bn_train: launch failed");
}

int agx_pool_fc(const float* x, int64_t n, int pixels, int C, const float* wfc, const float* bfc, int F, float* out, int64_t ld_out, void* stream) {
    if (!x || !wfc || !bfc || !out || n <= 0 || pixels <= 0 || C <= 0 || C > 128 || (128 % C) || F <= 0 || F > 128 || ld_out < F)
        return agx_internal_fail(AGX_ERR_ARG, "agx_pool_fc: bad argument (C must divide 128, F <= 128)");
    agx_pool_fc_kernel<<<(unsigned)n, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, pixels, C, wfc, bfc, F, out, ld_out);
    return cudaGetLastError() == cudaSuccess ? AGX_OK : agx_internal_fail(AGX_ERR_CUDA, "agx_pool_fc: launch failed");
}

}  // extern "C"

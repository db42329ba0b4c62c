// agx_mlp.cu — fused actor-critic MLP for the PPO path (reference lib/network/mlp.py:37-39 + the mu / value heads of
// lib/model/a2c_continuous_logstd_model.py:159-168, input normalisation lib/core/running_mean_std.py:76-80).
//
// forward : obs → clamp((obs-mean)/sqrt(var+eps), ±5) → [Linear+ELU]x3 → (mu | value) in ONE kernel; activations live in
//           shared memory between layers and are optionally kept in HBM for the backward pass;
// backward: (dmu | dvalue) → dZ3 → dZ2 → dZ1 in ONE kernel, which also accumulates the bias gradients (column sums);
// wgrad   : dW_l = dZ_l^T · a_{l-1} for all four weight matrices in ONE split-K kernel (every CTA reduces a slab of the batch)
//           followed by a small deterministic reduction that scatters into the caller's flat gradient buffer.
// All GEMMs run on the tensor cores: TF32 operands, fp32 accumulation (warp-level mma.sync m16n8k8, operands pre-rounded to TF32 in shared memory).
// In forward/backward each warp owns a tile of 16 samples end to end (no block-level sync after the weights are staged);
// shared-memory leading dimensions are padded (+4 / +8 floats) so that the fragment loads are bank-conflict free.
#include <cuda_runtime.h>
#include <string.h>
#include <stdint.h>

#include "agx.h"
#include "agx_math.cuh"
#include "agx_tc.cuh"

int agx_internal_fail(int code, const char* msg);
extern int g_wgrad_tma;  // agx_mlp_train.cu: operands of the weight-gradient kernel by TMA (agx_set_option("mlp_wgrad_tma"))

namespace {

constexpr int kWarps = 8;          // warps per CTA
constexpr int kRows = 16;          // rows per warp tile
constexpr int kOutPad = 16;        // head width padded to 16: [mu(A) | value | 0...]
constexpr int kMaxW = 128;         // widest layer
constexpr int kActLd = kMaxW + 4;  // activation tile leading dimension (≡ 4 mod 32 → conflict-free A-fragment loads)
constexpr int kBiasSlots = 3 * kMaxW + kOutPad;  // per-CTA bias partials: b1 | b2 | b3 | heads (each padded to kMaxW)

// layer widths either at run time (Dims, any supported network) or baked in at compile time (SDims: the shipped 64-128-64
// network; every width test, loop bound and GEMM dispatch then folds away — the run-time version executes ~4x more instructions)
struct Dims {
    int in_dim, in_pad, h1, h2, h3, a;  // in_pad multiple of 16, h* multiples of 32; a = actions_num (4 or 5)
};
template <int IN_PAD, int H1, int H2, int H3>
struct SDims {
    int in_dim, a;
    static constexpr int in_pad = IN_PAD, h1 = H1, h2 = H2, h3 = H3;
};
template <class D> __host__ __device__ inline D dims_of(const AgxMlpParams& P);
template <> __host__ __device__ inline Dims dims_of<Dims>(const AgxMlpParams& P) { return Dims{P.in_dim, P.in_pad, P.h1, P.h2, P.h3, P.actions_num}; }
template <class D> __host__ __device__ inline D dims_of(const AgxMlpParams& P) { D d; d.in_dim = P.in_dim; d.a = P.actions_num; return d; }

// shared-memory layout (float offsets into g_smem): W_l stored [out x (in + pad)] row-major, then biases, then per-warp tiles
struct SmemW {
    int w1, w2, w3, wh, b1, b2, b3, bh, end;
    int ld1, ld2, ld3, ldh;
};
template <class D>
__host__ __device__ inline SmemW carve_weights(const D& d, int pad) {
    SmemW s;
    s.ld1 = d.in_pad + pad; s.ld2 = d.h1 + pad; s.ld3 = d.h2 + pad; s.ldh = d.h3 + pad;
    s.w1 = 0;                        s.w2 = s.w1 + d.h1 * s.ld1;    s.w3 = s.w2 + d.h2 * s.ld2;
    s.wh = s.w3 + d.h3 * s.ld3;      s.b1 = s.wh + kOutPad * s.ldh; s.b2 = s.b1 + d.h1;
    s.b3 = s.b2 + d.h2;              s.bh = s.b3 + d.h3;            s.end = s.bh + kOutPad;
    return s;
}
template <class D>
__host__ __device__ inline int weights_floats(const D& d, int pad) { return carve_weights(d, pad).end; }

// ---- warp-level tensor-core GEMM pieces: mma.sync m16n8k8, TF32 operands (pre-rounded in shared memory), fp32 accumulate.
// Fragment ownership (g = lane / 4, t = lane % 4):  A: a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4);
// B: b0 (k = t, n = g) b1 (k = t+4, n = g);  C: c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1).
extern __shared__ __align__(128) float g_smem[];  // indexed directly so that every access compiles to LDS/STS

__device__ inline float tf32r(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ inline void mma_tf32(float (&c)[4], const float (&a)[4], float b0, float b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
                   "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}

// Y[16 x n] (shared, offset yo, ld kActLd) = X[16 x k] (shared, offset xo, ld kActLd) · op(W) with W in shared at offset wo:
// TRANSPOSED_W = true : op(W)[k][n] = W[n * ldw + k]  (torch Linear weight rows, ldw ≡ 4 mod 32 → conflict-free)
// TRANSPOSED_W = false: op(W)[k][n] = W[k * ldw + n]  (same weight used backwards,  ldw ≡ 8 mod 32 → conflict-free)
// One pass computes NT n8-tiles (8·NT columns) with NT independent accumulator chains fed by one A fragment per k-step; NT is a
// compile-time constant so the MMAs sit in straight-line code (no per-MMA convergence barriers).  Legacy mma.sync has ~127
// cycles of latency on sm_100a (measured, scripts/micro/mma_lat.cu): the 8–16 independent chains are what hides it.
template <bool TRANSPOSED_W, int NT>
__device__ inline void gemm_pass(int xo, int wo, int ldw, int k, int n0, int yo, int g, int t) {
    float acc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) { acc[j][0] = 0.0f; acc[j][1] = 0.0f; acc[j][2] = 0.0f; acc[j][3] = 0.0f; }
    const int xa = xo + g * kActLd + t;
    const int wb = TRANSPOSED_W ? (wo + (n0 + g) * ldw + t) : (wo + t * ldw + n0 + g);
    for (int k0 = 0; k0 < k; k0 += 8) {
        float a[4];
        a[0] = g_smem[xa + k0];
        a[1] = g_smem[xa + 8 * kActLd + k0];
        a[2] = g_smem[xa + k0 + 4];
        a[3] = g_smem[xa + 8 * kActLd + k0 + 4];
        float b0[NT], b1[NT];
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            if (TRANSPOSED_W) {
                b0[j] = g_smem[wb + 8 * j * ldw + k0];
                b1[j] = g_smem[wb + 8 * j * ldw + k0 + 4];
            } else {
                b0[j] = g_smem[wb + k0 * ldw + 8 * j];
                b1[j] = g_smem[wb + (k0 + 4) * ldw + 8 * j];
            }
        }
#pragma unroll
        for (int j = 0; j < NT; ++j) mma_tf32(acc[j], a, b0[j], b1[j]);
    }
    const int yc = yo + g * kActLd + n0 + 2 * t;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        *reinterpret_cast<float2*>(&g_smem[yc + 8 * j]) = make_float2(acc[j][0], acc[j][1]);
        *reinterpret_cast<float2*>(&g_smem[yc + 8 * kActLd + 8 * j]) = make_float2(acc[j][2], acc[j][3]);
    }
}
template <bool TRANSPOSED_W>
__device__ inline void gemm16(int xo, int wo, int ldw, int k, int n, int yo, int lane) {
    const int g = lane >> 2, t = lane & 3;
    if (n == 128) { gemm_pass<TRANSPOSED_W, 16>(xo, wo, ldw, k, 0, yo, g, t); return; }
    if (n == 64) { gemm_pass<TRANSPOSED_W, 8>(xo, wo, ldw, k, 0, yo, g, t); return; }
    if (n == 16) { gemm_pass<TRANSPOSED_W, 2>(xo, wo, ldw, k, 0, yo, g, t); return; }
    for (int n0 = 0; n0 < n; n0 += 32) gemm_pass<TRANSPOSED_W, 4>(xo, wo, ldw, k, n0, yo, g, t);  // n is a multiple of 32
}

// copy a [rows x cols] row-major matrix from global memory into shared memory with leading dimension ld (zero padded),
// rounding to TF32.  Four rows per warp iteration, all loads issued before the first store: ~20 independent loads in flight
// per thread, so the L2 latency is paid a handful of times instead of once per element.
__device__ inline void stage_matrix(int dst, int ld, const float* __restrict__ src, int rows, int cols, int warp, int lane) {
    constexpr int kC = (kMaxW + 8 + 31) / 32;  // column slots per lane (ld <= kMaxW + 8)
    for (int r0 = 4 * warp; r0 < rows; r0 += 4 * kWarps) {
        float v[4][kC];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int u = 0; u < kC; ++u) {
                const int r = r0 + q, c = lane + 32 * u;
                v[q][u] = (r < rows && c < cols) ? __ldg(src + (int64_t)r * cols + c) : 0.0f;
            }
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int u = 0; u < kC; ++u) {
                const int r = r0 + q, c = lane + 32 * u;
                if (r < rows && c < ld) g_smem[dst + r * ld + c] = tf32r(v[q][u]);
            }
    }
}
// stage the network into shared memory (weights pre-rounded to TF32; W1 zero padded; the two heads stacked into 16 rows)
template <class D>
__device__ void stage_weights(const AgxMlpParams& P, const D& d, const SmemW& s) {
    const int tid = threadIdx.x, nt = blockDim.x, warp = tid >> 5, lane = tid & 31;
    stage_matrix(s.w1, s.ld1, P.w1, d.h1, d.in_dim, warp, lane);
    stage_matrix(s.w2, s.ld2, P.w2, d.h2, d.h1, warp, lane);
    stage_matrix(s.w3, s.ld3, P.w3, d.h3, d.h2, warp, lane);
    stage_matrix(s.wh, s.ldh, P.w_mu, d.a, d.h3, warp, lane);                      // rows 0..A-1: mu head
    stage_matrix(s.wh + d.a * s.ldh, s.ldh, P.w_value, 1, d.h3, warp, lane);       // row A: value head
    for (int i = tid; i < (kOutPad - d.a - 1) * s.ldh; i += nt) g_smem[s.wh + (d.a + 1) * s.ldh + i] = 0.0f;
    for (int i = tid; i < d.h1; i += nt) g_smem[s.b1 + i] = P.b1[i];
    for (int i = tid; i < d.h2; i += nt) g_smem[s.b2 + i] = P.b2[i];
    for (int i = tid; i < d.h3; i += nt) g_smem[s.b3 + i] = P.b3[i];
    for (int i = tid; i < kOutPad; i += nt) g_smem[s.bh + i] = i < d.a ? P.b_mu[i] : (i == d.a ? P.b_value[0] : 0.0f);
}

__device__ inline float elu(float x) { return x > 0.0f ? x : __expf(x) - 1.0f; }  // SFU exp: abs error ~1e-7, far below TF32
__device__ inline float elu_grad_from_out(float h) { return h > 0.0f ? 1.0f : h + 1.0f; }  // d elu/dz through h = elu(z)
constexpr int kU = kMaxW / 32;  // column slots per lane: lane owns columns lane + 32 u

// bias + ELU over a 16 x width tile at offset `bo` (TF32-rounded copy stays in shared memory for the next GEMM; the
// full-precision value is optionally kept in HBM)
__device__ inline void bias_elu_tile(int bo, int bias_o, int width, int lane, float* out, int64_t row0, int rows_valid) {
    if (out) out += row0 * width;
    float b[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) b[u] = (lane + 32 * u < width) ? g_smem[bias_o + lane + 32 * u] : 0.0f;
#pragma unroll 4
    for (int r = 0; r < kRows; ++r) {
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int c = lane + 32 * u;
            if (c < width) {
                const float h = elu(g_smem[bo + r * kActLd + c] + b[u]);
                g_smem[bo + r * kActLd + c] = tf32r(h);
                if (out && r < rows_valid) out[r * width + c] = h;
            }
        }
    }
}
// g = dH ∘ elu'(h) over a 16 x width tile: dz to HBM, TF32 copy kept in shared memory for the next GEMM, column sums in bacc.
// The activations `hv` were prefetched from HBM before the GEMM that produced dH (latency hidden behind the tensor work).
__device__ inline void elu_grad_tile(int bo, const float (&hv)[kRows][kU], float* __restrict__ dz, int width, int lane, int64_t row0,
                                     int rows_valid, float* bacc, bool keep_in_smem) {
    dz += row0 * width;
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int c = lane + 32 * u;
            if (c < width) {
                float g = 0.0f;
                if (r < rows_valid) {
                    g = g_smem[bo + r * kActLd + c] * elu_grad_from_out(hv[r][u]);
                    dz[r * width + c] = g;
                }
                if (keep_in_smem) g_smem[bo + r * kActLd + c] = tf32r(g);
                bacc[u] += g;
            }
        }
    }
}
__device__ inline void prefetch_tile(const float* __restrict__ h, float (&hv)[kRows][kU], int width, int lane, int64_t row0, int rows_valid) {
    h += row0 * width;
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int c = lane + 32 * u;
            hv[r][u] = (c < width && r < rows_valid) ? h[r * width + c] : 0.0f;
        }
    }
}

// ---- forward ---------------------------------------------------------------------------------------------------------------
template <class D>
__global__ void __launch_bounds__(kWarps * 32)
agx_mlp_forward_kernel(const __grid_constant__ AgxMlpParams P, int64_t B, const float* __restrict__ obs,
                       float* __restrict__ mu, float* __restrict__ value, float* __restrict__ xn_out,
                       float* __restrict__ h1_out, float* __restrict__ h2_out, float* __restrict__ h3_out) {
    const D d = dims_of<D>(P);
    const SmemW W = carve_weights(d, 4);  // ld ≡ 4 (mod 32): conflict-free B fragments of X · W^T
    stage_weights(P, d, W);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bufA = W.end + warp * (2 * kRows * kActLd), bufB = bufA + kRows * kActLd;  // ping-pong tiles of this warp
    float n_mean[kU], n_sd[kU];  // this lane's input columns: mean and sqrt(var + eps) as float
#pragma unroll
    for (int u = 0; u < kU; ++u) {
        const int c = lane + 32 * u;
        const bool on = P.in_mean && c < d.in_dim;
        n_mean[u] = on ? (float)P.in_mean[c] : 0.0f;
        n_sd[u] = on ? sqrtf((float)P.in_var[c] + 1e-5f) : 1.0f;
    }
    const int64_t n_tiles = (B + kRows - 1) / kRows;
    for (int64_t tile = (int64_t)blockIdx.x * kWarps + warp; tile < n_tiles; tile += (int64_t)gridDim.x * kWarps) {
        const int64_t row0 = tile * kRows;
        const int rows_valid = (B - row0) < kRows ? (int)(B - row0) : kRows;
        // normalised input tile (RunningMeanStd eval branch, lib/core/running_mean_std.py:76-80)
#pragma unroll 4
        for (int r = 0; r < kRows; ++r) {
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int c = lane + 32 * u;
                if (c < d.in_pad) {
                    float v = 0.0f;
                    if (c < d.in_dim && r < rows_valid) {
                        v = obs[(row0 + r) * d.in_dim + c];
                        if (P.in_mean) {
                            v = (v - n_mean[u]) / n_sd[u];
                            v = v < -5.0f ? -5.0f : (v > 5.0f ? 5.0f : v);
                        }
                    }
                    g_smem[bufA + r * kActLd + c] = tf32r(v);
                    if (xn_out && r < rows_valid) xn_out[(row0 + r) * d.in_pad + c] = v;
                }
            }
        }
        __syncwarp();
        gemm16<true>(bufA, W.w1, W.ld1, d.in_pad, d.h1, bufB, lane);
        __syncwarp();
        bias_elu_tile(bufB, W.b1, d.h1, lane, h1_out, row0, rows_valid);
        __syncwarp();
        gemm16<true>(bufB, W.w2, W.ld2, d.h1, d.h2, bufA, lane);
        __syncwarp();
        bias_elu_tile(bufA, W.b2, d.h2, lane, h2_out, row0, rows_valid);
        __syncwarp();
        gemm16<true>(bufA, W.w3, W.ld3, d.h2, d.h3, bufB, lane);
        __syncwarp();
        bias_elu_tile(bufB, W.b3, d.h3, lane, h3_out, row0, rows_valid);
        __syncwarp();
        gemm16<true>(bufB, W.wh, W.ldh, d.h3, kOutPad, bufA, lane);
        __syncwarp();
        {
            const int c = lane & 15;
            const float bh = g_smem[W.bh + c];
#pragma unroll
            for (int j = 0; j < kRows / 2; ++j) {
                const int r = (lane >> 4) + 2 * j;
                if (r < rows_valid) {
                    const float v = g_smem[bufA + r * kActLd + c] + bh;
                    if (c < d.a) mu[(row0 + r) * d.a + c] = v;
                    else if (c == d.a) value[row0 + r] = v;
                }
            }
        }
        __syncwarp();
    }
}

// ---- forward on the 5th-generation tensor cores (tcgen05 + TMEM) ---------------------------------------------------------------
// One CTA = 128 threads = one 128-row tile of the batch at a time (persistent over tiles).  Each layer is ONE accumulation chain of
// tcgen05.mma.cta_group::1.kind::tf32 instructions (M = 128, N = layer width, K = 8 per instruction) issued by a single thread:
// A = the activations of the previous layer, B = the layer's weights, both K-major in the canonical no-swizzle shared-memory
// layout ([K/4 chunks][rows/8][8 rows][16 B]: LBO = byte stride between the two 16-B K-chunks of an instruction, SBO = byte stride
// between 8-row groups; validated by scripts/micro/umma_probe.cu), D = fp32 accumulators in tensor memory.  Completion reaches the
// 128 epilogue threads through tcgen05.commit → mbarrier; thread r owns row r = TMEM lane r: tcgen05.ld, + bias, ELU, and the
// result goes straight back into shared memory as the next layer's A operand (16-byte chunk stores, conflict-free) and, when
// training, to HBM for the backward.  TMEM columns: layer 1 → [0,64), layer 2 → [64,192), layer 3 → [192,256), heads → [0,16).
// The mma.sync kernel above spends 44 us on a 32 768-row minibatch (legacy-MMA issue bound); SASS of this one shows UTCMMA / LDTM.
namespace tc {
// epilogue of one hidden layer for this thread's row: TMEM cols [col0, col0 + W) → + bias → ELU → next A operand (+ HBM copy)
// keep_row: this row's slot in a row-major [b, W] keep tensor; keep_col (training path): this row's element of plane 0 of a
// FEATURE-MAJOR [W, ld_t] keep tensor (consecutive rows = consecutive addresses: coalesced scalar stores) — at most one is non-null
// half: which half of the W columns this thread handles (two warps share a TMEM lane quarter and split the columns between them)
template <int W>
__device__ __forceinline__ void hidden_epilogue(uint32_t tmem_row, int col0, const float* __restrict__ bias, float* a_next, int r, int half,
                                                float* __restrict__ keep_row, float* __restrict__ keep_col = nullptr, int64_t ld_t = 0) {
    // not unrolled over the column chunks: four call sites x up to 8 chunks of ~250 instructions each overflowed the instruction
    // cache (ncu: stall_no_instruction 6.2 of 11 cycles per issue with the two tile groups in different code regions)
#pragma unroll 1
    for (int c0 = half * (W / 2); c0 < (half + 1) * (W / 2); c0 += 16) {
        float v[16];
        tmem_ld16(tmem_row + (uint32_t)(col0 + c0), v);
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
            const float4 b = *reinterpret_cast<const float4*>(bias + c0 + i);  // broadcast LDS.128
            v[i] = elu(v[i] + b.x); v[i + 1] = elu(v[i + 1] + b.y); v[i + 2] = elu(v[i + 2] + b.z); v[i + 3] = elu(v[i + 3] + b.w);
        }
        if (keep_row) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(keep_row + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
        if (keep_col) {
#pragma unroll
            for (int i = 0; i < 16; ++i) keep_col[(int64_t)(c0 + i) * ld_t] = v[i];
        }
#pragma unroll
        for (int i = 0; i < 16; i += 4)
            // fp32 bit patterns go in as they are: kind::tf32 reads the top 19 bits (truncation, <= 1 tf32 ulp; the weights were rounded
            // once at staging) — cvt.rna.tf32 is a ~6-instruction software sequence on sm_100 and was 30 % of this kernel
            *reinterpret_cast<float4*>(a_next + canon(r, c0 + i, kM)) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    }
}

// The policy head of one env row (agx.h AgxPolicyIO): heads accumulators v (+ bias) → mu, value; sample, neglogp, record.
__device__ __forceinline__ void policy_head(const AgxPolicyIO& pol, const float (&v)[16], const float* bh, int A, int64_t row) {
    float mu[5], z[5];
#pragma unroll
    for (int a = 0; a < 5; ++a) mu[a] = a < A ? v[a] + bh[a] : 0.0f;
    float val = A == 5 ? v[5] + bh[5] : v[4] + bh[4];
    if (pol.value_mean) {  // RunningMeanStd(denorm=True), running_mean_std.py:76-80
        const float m = (float)pol.value_mean[0], sd = sqrtf((float)pol.value_var[0] + 1e-5f);
        val = sd * fminf(fmaxf(val, -5.0f), 5.0f) + m;
    }
    if (pol.noise) {
#pragma unroll
        for (int a = 0; a < 5; ++a) z[a] = a < A ? pol.noise[row * A + a] : 0.0f;
    } else {  // Philox stream 6, two blocks = 8 words = up to 4 Box-Muller pairs (32-bit uniforms)
        agx::PhiloxCtx ph;
        const uint64_t genv = (uint64_t)(pol.env_offset + row), step = pol.step_dev ? *pol.step_dev : 0ull;
        ph.k0 = (uint32_t)pol.seed; ph.k1 = (uint32_t)(pol.seed >> 32);
        ph.env_lo = (uint32_t)genv; ph.env_hi = (uint32_t)(genv >> 32);
        ph.step_lo = (uint32_t)step; ph.step_hi = (uint32_t)(step >> 32);
        const agx::U4 r0 = agx::philox_block(ph, 6u, 0u), r1 = agx::philox_block(ph, 6u, 1u);
        const uint32_t w[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float u1 = ((float)(w[2 * k] >> 8) + 1.0f) * 5.9604644775390625e-8f;  // (0, 1]
            const float u2 = (float)(w[2 * k + 1] >> 8) * 5.9604644775390625e-8f;         // [0, 1)
            const float rad = sqrtf(-2.0f * logf(u1));
            float sn, cs;
            sincosf(6.283185307179586f * u2, &sn, &cs);
            z[2 * k] = rad * cs;
            if (2 * k + 1 < 5) z[2 * k + 1] = rad * sn;
        }
    }
    float nlp = 0.0f, lsum = 0.0f;
#pragma unroll
    for (int a = 0; a < 5; ++a) {
        if (a < A) {
            const float ls = pol.logstd[a], sg = expf(ls), act = mu[a] + sg * z[a];
            const float t = (act - mu[a]) / sg;  // as the model's neglogp evaluates it on the sampled action (:195-198)
            nlp += t * t;
            lsum += ls;
            pol.actions[row * pol.ld_actions + a] = act;
            pol.mus[row * pol.ld_mus + a] = mu[a];
            pol.sigmas[row * pol.ld_sigmas + a] = sg;
            float e = act;
            if (pol.act_lo) {  // preprocess_actions: clamp to [-1, 1], rescale to the action space (a2c_continuous.py:61-71)
                const float lo = pol.act_lo[a], hi = pol.act_hi[a];
                e = fminf(fmaxf(act, -1.0f), 1.0f) * ((hi - lo) * 0.5f) + (hi + lo) * 0.5f;
            }
            pol.env_actions[row * A + a] = e;
        }
    }
    pol.neglogp[row * pol.ld_neglogp] = 0.5f * nlp + 0.9189385332046727f * (float)A + lsum;
    pol.values[row * pol.ld_values] = val;
    if (pol.dones_out) pol.dones_out[row * pol.ld_dones] = pol.dones_in[row];
}

// Two tile groups per CTA (the MMA / barrier latency of one hides under the epilogue of the other), 256 threads per group: thread
// (row, half) — the two warps of a TMEM lane quarter split every layer's columns, which halves the per-tile epilogue chain (the
// kernel is bound by that chain: one 32 768-row minibatch is a single pass of 256 tiles over the 296 groups of the grid)
constexpr int kGroup = 2 * kM, kThreads = 2 * kGroup;
template <int IN_PAD>
__global__ void __launch_bounds__(kThreads, 1)
agx_mlp_forward_tc_kernel(const __grid_constant__ AgxMlpParams P, int64_t B, const float* __restrict__ obs, float* __restrict__ mu,
                          float* __restrict__ value, float* __restrict__ xn_out, float* __restrict__ h1_out, float* __restrict__ h2_out,
                          float* __restrict__ h3_out, const int keep_t,  // keep_t: the four keep tensors are feature-major [width, B] planes,
                                                                         // and plane in_dim of xn_out is set to 1 (bias-gradient column)
                          const __grid_constant__ AgxPolicyIO pol) {     // pol.actions != NULL: the policy head of a rollout step (agx.h)
    // shared memory carve (floats); every operand base is 128-byte aligned
    float* w1 = g_smem;                       // [64 x IN_PAD] canonical
    float* w2 = w1 + kH1 * IN_PAD;            // [128 x 64]
    float* w3 = w2 + kH2 * kH1;               // [64 x 128]
    float* wh = w3 + kH3 * kH2;               // [16 x 64]
    float* bia = wh + kOutPad * kH3;          // b1 | b2 | b3 | heads = 64 + 128 + 64 + 16
    float* nrm = bia + (kH1 + kH2 + kH3 + kOutPad);  // input mean | sqrt(var + eps): 2 x IN_PAD
    float* act = nrm + 2 * IN_PAD;            // per group: Pbuf [128 x 64] (A0, A2[:, :64], A3) | Qbuf [128 x 64] (A1, A2[:, 64:])
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_base;
    const int tid_all = threadIdx.x, warp_all = tid_all >> 5, group = tid_all / kGroup, tig = tid_all % kGroup, tid = tig & (kM - 1), half = tig >> 7,
              A = P.actions_num, in_dim = P.in_dim;
    float* Pbuf = act + group * (2 * kM * kH1);
    float* Qbuf = Pbuf + kM * kH1;
    uint64_t* bar = &bars[group];
    // weights → canonical layout, TF32-rounded.  canon(n, k, N) = ((k / 4) * N + n) * 4 + k % 4: one 16-byte chunk per (row, K-chunk);
    // the loads of a batch of chunks are issued together (an un-unrolled load→store loop would serialise ~150 L2 round trips)
    auto stage4 = [&](float* dst, const float* __restrict__ src, int N, int K) {  // src rows 16-byte aligned (K % 4 == 0)
        const int chunks = N * (K >> 2);
        for (int i0 = tid_all; i0 < chunks; i0 += kThreads * 8) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = i0 + u * kThreads;
                if (i < chunks) { const int kc = i / N, n = i - kc * N; v[u] = *reinterpret_cast<const float4*>(src + n * K + kc * 4); }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = i0 + u * kThreads;
                if (i < chunks) *reinterpret_cast<float4*>(dst + i * 4) = make_float4(tf32r(v[u].x), tf32r(v[u].y), tf32r(v[u].z), tf32r(v[u].w));
            }
        }
    };
    stage4(w2, P.w2, kH2, kH1);
    stage4(w3, P.w3, kH3, kH2);
    {   // w1: in_dim (18 / 48 / 46) columns per row, zero padded to IN_PAD; heads: [w_mu (A rows) | w_value | 0] as a 16-row operand
        constexpr int kPer = (kH1 * IN_PAD + kThreads - 1) / kThreads;
        float v[kPer];
#pragma unroll
        for (int u = 0; u < kPer; ++u) { const int i = tid_all + u * kThreads, n = i / IN_PAD, k = i - n * IN_PAD; v[u] = (i < kH1 * IN_PAD && k < in_dim) ? P.w1[n * in_dim + k] : 0.0f; }
#pragma unroll
        for (int u = 0; u < kPer; ++u) { const int i = tid_all + u * kThreads, n = i / IN_PAD, k = i - n * IN_PAD; if (i < kH1 * IN_PAD) w1[canon(n, k, kH1)] = tf32r(v[u]); }
        float h[kOutPad * kH3 / kThreads];
#pragma unroll
        for (int u = 0; u < kOutPad * kH3 / kThreads; ++u) { const int i = tid_all + u * kThreads, n = i / kH3, k = i - n * kH3; h[u] = n < A ? P.w_mu[n * kH3 + k] : (n == A ? P.w_value[k] : 0.0f); }
#pragma unroll
        for (int u = 0; u < kOutPad * kH3 / kThreads; ++u) { const int i = tid_all + u * kThreads, n = i / kH3, k = i - n * kH3; wh[canon(n, k, kOutPad)] = tf32r(h[u]); }
    }
    for (int i = tid_all; i < kH1; i += kThreads) bia[i] = P.b1[i];
    for (int i = tid_all; i < kH2; i += kThreads) bia[kH1 + i] = P.b2[i];
    for (int i = tid_all; i < kH3; i += kThreads) bia[kH1 + kH2 + i] = P.b3[i];
    for (int i = tid_all; i < kOutPad; i += kThreads) bia[kH1 + kH2 + kH3 + i] = i < A ? P.b_mu[i] : (i == A ? P.b_value[0] : 0.0f);
    for (int i = tid_all; i < IN_PAD; i += kThreads) {
        const bool on = P.in_mean && i < in_dim;
        nrm[i] = on ? (float)P.in_mean[i] : 0.0f;
        nrm[IN_PAD + i] = on ? sqrtf((float)P.in_var[i] + 1e-5f) : 1.0f;
    }
    if (tid_all == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp_all == 0) {  // the whole tensor memory: 256 columns per tile group
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    publish_and_sync(0, kThreads);
    const uint32_t tmem = tmem_base + (uint32_t)(group * 256);                      // this group's columns
    const uint32_t tmem_row = tmem + ((uint32_t)((warp_all & 3) * 32) << 16);       // a warp reaches TMEM lanes 32 (warp % 4) ..
    const int gbar = 1 + group;
    uint32_t phase = 0;
    const int64_t n_tiles = (B + kM - 1) / kM;
    for (int64_t tile = (int64_t)blockIdx.x * 2 + group; tile < n_tiles; tile += (int64_t)gridDim.x * 2) {
        const int64_t row = tile * kM + tid;
        const bool ok = row < B;
        // normalised input row (RunningMeanStd eval branch, lib/core/running_mean_std.py:76-80) → A0 in Pbuf
#pragma unroll
        for (int c0 = 0; c0 < IN_PAD; c0 += 4) {
            if (((c0 >> 2) & 1) != half) continue;  // the two threads of a row take alternate 16-byte chunks
            float v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int c = c0 + i;
                float x = 0.0f;
                if (ok && c < in_dim) {
                    x = obs[row * in_dim + c];
                    if (pol.obs_out) pol.obs_out[row * pol.ld_obs + c] = x;
                    if (P.in_mean) {
                        x = (x - nrm[c]) / nrm[IN_PAD + c];
                        x = x < -5.0f ? -5.0f : (x > 5.0f ? 5.0f : x);
                    }
                }
                v[i] = x;
            }
            if (xn_out && ok) {
                if (keep_t) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) xn_out[((row >> 7) * IN_PAD + c0 + i) * kM + (row & (kM - 1))] = (c0 + i == in_dim) ? 1.0f : v[i];
                } else {
                    *reinterpret_cast<float4*>(xn_out + row * IN_PAD + c0) = make_float4(v[0], v[1], v[2], v[3]);
                }
            }
            *reinterpret_cast<float4*>(Pbuf + canon(tid, c0, kM)) = make_float4(v[0], v[1], v[2], v[3]);
        }
        publish_and_sync(gbar, kGroup);
        if (tig == 0) { gemm(s32(Pbuf), s32(w1), kH1, IN_PAD, tmem + 0); commit(bar); }
        wait(bar, phase); phase ^= 1;
        const bool kr = ok && !keep_t, kc = ok && keep_t;  // row-major / feature-major keeps
        // feature-major keeps are blocked by 128-row tile: element (row, c) of a W-wide tensor at ((row / 128) * W + c) * 128 + row % 128,
        // i.e. one tile's planes are one contiguous W x 512-byte region (DRAM-page friendly for this kernel's stores and the weight-gradient kernel's loads)
        const int64_t tile_row = row >> 7, in_tile = row & (kM - 1);
        hidden_epilogue<kH1>(tmem_row, 0, bia, Qbuf, tid, half, (h1_out && kr) ? h1_out + row * kH1 : nullptr,
                             (h1_out && kc) ? h1_out + tile_row * kH1 * kM + in_tile : nullptr, kM);  // A1 → Qbuf
        publish_and_sync(gbar, kGroup);
        if (tig == 0) { gemm(s32(Qbuf), s32(w2), kH2, kH1, tmem + 64); commit(bar); }
        wait(bar, phase); phase ^= 1;
        // layer 2's 128 columns leave in two halves so that A2 never needs more than the two 32 KB buffers: the first half goes to
        // Pbuf and layer 3 starts on it (K-steps 0..7) while the epilogue of the second half fills Qbuf (A1 is dead by now)
        hidden_epilogue<kH1>(tmem_row, 64, bia + kH1, Pbuf, tid, half, (h2_out && kr) ? h2_out + row * kH2 : nullptr,
                             (h2_out && kc) ? h2_out + tile_row * kH2 * kM + in_tile : nullptr, kM);
        publish_and_sync(gbar, kGroup);
        if (tig == 0) gemm(s32(Pbuf), s32(w3), kH3, kH1, tmem + 192);
        hidden_epilogue<kH1>(tmem_row, 64 + kH1, bia + kH1 + kH1, Qbuf, tid, half, (h2_out && kr) ? h2_out + row * kH2 + kH1 : nullptr,
                             (h2_out && kc) ? h2_out + (tile_row * kH2 + kH1) * kM + in_tile : nullptr, kM);
        publish_and_sync(gbar, kGroup);
        if (tig == 0) { gemm(s32(Qbuf), s32(w3) + 8 * 2 * (kH3 / 8) * 128, kH3, kH1, tmem + 192, true); commit(bar); }
        wait(bar, phase); phase ^= 1;
        hidden_epilogue<kH3>(tmem_row, 192, bia + kH1 + kH2, Pbuf, tid, half, (h3_out && kr) ? h3_out + row * kH3 : nullptr,
                             (h3_out && kc) ? h3_out + tile_row * kH3 * kM + in_tile : nullptr, kM);  // A3 → Pbuf
        publish_and_sync(gbar, kGroup);
        if (tig == 0) { gemm(s32(Pbuf), s32(wh), kOutPad, kH3, tmem + 0); commit(bar); }
        wait(bar, phase); phase ^= 1;
        if (half == 0) {  // the 16 head columns: one thread per row
            float v[16];
            tmem_ld16(tmem_row + 0u, v);
            if (ok && pol.actions) {
                policy_head(pol, v, bia + kH1 + kH2 + kH3, A, row);
            } else if (ok) {
                const float* bh = bia + kH1 + kH2 + kH3;
                for (int a = 0; a < A; ++a) mu[row * A + a] = v[a] + bh[a];
                float val = v[4] + bh[4];
                if (A == 5) val = v[5] + bh[5];
                value[row] = val;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp_all == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}
template <int IN_PAD>
constexpr size_t smem_bytes_tc() {
    return sizeof(float) * (size_t)(kH1 * IN_PAD + kH2 * kH1 + kH3 * kH2 + kOutPad * kH3 + (kH1 + kH2 + kH3 + kOutPad) + 2 * IN_PAD + 2 * (2 * kM * kH1));
}
}  // namespace tc

// ---- backward: activation-gradient chain + bias-gradient partials ----------------------------------------------------------------
template <class D>
__global__ void __launch_bounds__(kWarps * 32)
agx_mlp_backward_kernel(const __grid_constant__ AgxMlpParams P, int64_t B, const float* __restrict__ grad_mu,
                        const float* __restrict__ grad_value, const float* __restrict__ h1, const float* __restrict__ h2,
                        const float* __restrict__ h3, float* __restrict__ dz1, float* __restrict__ dz2,
                        float* __restrict__ dz3, float* __restrict__ dout, float* __restrict__ bias_partials) {
    const D d = dims_of<D>(P);
    const SmemW W = carve_weights(d, 8);  // ld ≡ 8 (mod 32): conflict-free B fragments of X · W
    stage_weights(P, d, W);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bufA = W.end + warp * (2 * kRows * kActLd), bufB = bufA + kRows * kActLd;
    // per-lane column sums: lane L owns columns L + 32 u of every layer, and column L % 16 of the padded head gradient
    float bacc1[kU], bacc2[kU], bacc3[kU], bacco = 0.0f;
#pragma unroll
    for (int u = 0; u < kU; ++u) { bacc1[u] = 0.0f; bacc2[u] = 0.0f; bacc3[u] = 0.0f; }
    const int q1 = d.h1 / 32, q2 = d.h2 / 32, q3 = d.h3 / 32;
    const int64_t n_tiles = (B + kRows - 1) / kRows;
    for (int64_t tile = (int64_t)blockIdx.x * kWarps + warp; tile < n_tiles; tile += (int64_t)gridDim.x * kWarps) {
        const int64_t row0 = tile * kRows;
        const int rows_valid = (B - row0) < kRows ? (int)(B - row0) : kRows;
        float hv[kRows][kU];
        prefetch_tile(h3, hv, d.h3, lane, row0, rows_valid);
        {
            const int c = lane & 15;
#pragma unroll
            for (int j = 0; j < kRows / 2; ++j) {
                const int r = (lane >> 4) + 2 * j;
                float v = 0.0f;
                if (r < rows_valid) {
                    v = c < d.a ? grad_mu[(row0 + r) * d.a + c] : (c == d.a ? grad_value[row0 + r] : 0.0f);
                    dout[(row0 + r) * kOutPad + c] = v;
                }
                g_smem[bufA + r * kActLd + c] = tf32r(v);
                bacco += v;
            }
        }
        __syncwarp();
        gemm16<false>(bufA, W.wh, W.ldh, kOutPad, d.h3, bufB, lane);  // dH3 = dOut · W_head
        __syncwarp();
        elu_grad_tile(bufB, hv, dz3, d.h3, lane, row0, rows_valid, bacc3, true);
        prefetch_tile(h2, hv, d.h2, lane, row0, rows_valid);
        __syncwarp();
        gemm16<false>(bufB, W.w3, W.ld3, d.h3, d.h2, bufA, lane);  // dH2 = dZ3 · W3
        __syncwarp();
        elu_grad_tile(bufA, hv, dz2, d.h2, lane, row0, rows_valid, bacc2, true);
        prefetch_tile(h1, hv, d.h1, lane, row0, rows_valid);
        __syncwarp();
        gemm16<false>(bufA, W.w2, W.ld2, d.h2, d.h1, bufB, lane);  // dH1 = dZ2 · W2
        __syncwarp();
        elu_grad_tile(bufB, hv, dz1, d.h1, lane, row0, rows_valid, bacc1, false);
        __syncwarp();
    }
    // CTA-level reduction of the bias partials, fixed order → deterministic
    __syncthreads();
    const int red = 0;  // [kWarps][kBiasSlots]; the weight tiles are dead now
    for (int i = lane; i < kBiasSlots; i += 32) g_smem[red + warp * kBiasSlots + i] = 0.0f;
    __syncwarp();
#pragma unroll
    for (int u = 0; u < kU; ++u) {
        if (u < q1) g_smem[red + warp * kBiasSlots + 0 * kMaxW + lane + 32 * u] = bacc1[u];
        if (u < q2) g_smem[red + warp * kBiasSlots + 1 * kMaxW + lane + 32 * u] = bacc2[u];
        if (u < q3) g_smem[red + warp * kBiasSlots + 2 * kMaxW + lane + 32 * u] = bacc3[u];
    }
    const float o2 = bacco + __shfl_down_sync(0xffffffffu, bacco, 16);  // lanes L and L+16 share column L % 16
    if (lane < 16) g_smem[red + warp * kBiasSlots + 3 * kMaxW + lane] = o2;
    __syncthreads();
    for (int i = threadIdx.x; i < kBiasSlots; i += blockDim.x) {
        float s = 0.0f;
        for (int w = 0; w < kWarps; ++w) s += g_smem[red + w * kBiasSlots + i];
        bias_partials[(int64_t)blockIdx.x * kBiasSlots + i] = s;
    }
}

// ---- weight gradients: split-K over the batch -------------------------------------------------------------------------------------
// dW_l[out, in] = sum_b dz_l[b, out] · a_{l-1}[b, in] as mma.sync m16n8k8 with A = dz^T (m = out, k = batch row) and B = a
// (k = batch row, n = in), fragments loaded straight from HBM/L2.  Work unit = one 16-row out-tile x up to 8 n8 in-tiles
// (32 accumulator registers); the units of all four layers are dealt round-robin to the 8 warps of a CTA.
// The units are split between kWgradSplit CTAs that walk the same slab of the batch: each CTA then holds half of the dW
// accumulators (<= 2 units = 64 registers per thread), which lets two CTAs (16 warps) share an SM and hide the L2 latency
// of the fragment loads.
struct WUnit { int layer, o0, i0, nt; };
constexpr int kMaxUnits = 2;   // per warp
constexpr int kWgradSplit = 2;
template <class D>
__device__ inline int build_units(const D& d, int warp, int half, WUnit (&u)[kMaxUnits]) {
    const int outs[4] = {d.h1, d.h2, d.h3, kOutPad}, ins[4] = {d.in_pad, d.h1, d.h2, d.h3};
    int n = 0, q = 0;
    for (int l = 0; l < 4; ++l)
        for (int o0 = 0; o0 < outs[l]; o0 += 16)
            for (int i0 = 0; i0 < ins[l]; i0 += 64, ++q)
                if (q % kWgradSplit == half && (q / kWgradSplit) % kWarps == warp && n < kMaxUnits) {
                    u[n].layer = l; u[n].o0 = o0; u[n].i0 = i0; u[n].nt = (ins[l] - i0) >= 64 ? 8 : (ins[l] - i0) / 8; ++n;
                }
    return n;
}

template <class D>
__global__ void __launch_bounds__(kWarps * 32, 2)
agx_mlp_wgrad_kernel(const __grid_constant__ AgxMlpParams P, int64_t B, int64_t rows_per_cta, const float* __restrict__ xn,
                     const float* __restrict__ h1, const float* __restrict__ h2, const float* __restrict__ h3,
                     const float* __restrict__ dz1, const float* __restrict__ dz2, const float* __restrict__ dz3,
                     const float* __restrict__ dout, float* __restrict__ partials, int partial_floats) {
    const D d = dims_of<D>(P);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const float* dzs[4] = {dz1, dz2, dz3, dout};
    const float* as[4] = {xn, h1, h2, h3};
    const int outs[4] = {d.h1, d.h2, d.h3, kOutPad}, ins[4] = {d.in_pad, d.h1, d.h2, d.h3};
    WUnit un[kMaxUnits];
    const int slab = blockIdx.x / kWgradSplit, half = blockIdx.x % kWgradSplit;
    const int n_units = build_units(d, warp, half, un);
    float acc[kMaxUnits][8][4];
#pragma unroll
    for (int q = 0; q < kMaxUnits; ++q)
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc[q][j][0] = 0.0f; acc[q][j][1] = 0.0f; acc[q][j][2] = 0.0f; acc[q][j][3] = 0.0f; }
    const int64_t r_begin = (int64_t)slab * rows_per_cta;
    int64_t r_end = r_begin + rows_per_cta;
    if (r_end > B) r_end = B;  // B is a multiple of 8 (checked on the host)
    for (int64_t r = r_begin; r < r_end; r += 8) {
#pragma unroll
        for (int q = 0; q < kMaxUnits; ++q) {
            if (q < n_units) {
                const int l = un[q].layer, ldo = outs[l], ldi = ins[l];
                const float* dz = dzs[l] + r * ldo + un[q].o0;
                const float* a_ = as[l] + r * ldi + un[q].i0;
                float a[4];
                a[0] = tf32r(dz[(t)*ldo + g]);      a[1] = tf32r(dz[(t)*ldo + g + 8]);
                a[2] = tf32r(dz[(t + 4) * ldo + g]); a[3] = tf32r(dz[(t + 4) * ldo + g + 8]);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (j < un[q].nt) {
                        const float b0 = tf32r(a_[t * ldi + 8 * j + g]), b1 = tf32r(a_[(t + 4) * ldi + 8 * j + g]);
                        mma_tf32(acc[q][j], a, b0, b1);
                    }
                }
            }
        }
    }
    // partial dW of this CTA: dense [out_l x in_l] blocks, layers back to back
    float* mine = partials + (int64_t)slab * partial_floats;
#pragma unroll
    for (int q = 0; q < kMaxUnits; ++q) {
        if (q < n_units) {
            const int l = un[q].layer;
            int base = 0;
            for (int m = 0; m < l; ++m) base += outs[m] * ins[m];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (j < un[q].nt) {
                    const int c = un[q].i0 + 8 * j + 2 * t;
                    *reinterpret_cast<float2*>(&mine[base + (un[q].o0 + g) * ins[l] + c]) = make_float2(acc[q][j][0], acc[q][j][1]);
                    *reinterpret_cast<float2*>(&mine[base + (un[q].o0 + g + 8) * ins[l] + c]) = make_float2(acc[q][j][2], acc[q][j][3]);
                }
            }
        }
    }
}

// ---- wgrad, staged: the fragment loads above walk dz / activation rows with a 4-byte stride pattern straight from L2
// (79 us per 32 768-row minibatch in the round-1 launch list, 7x its HBM time).  Here the CTA streams its slab through shared
// memory in 16-row stages with cp.async (16-byte chunks, fully coalesced, double buffered: the copy of stage s+1 flies under the
// MMAs of stage s) and the warps read their fragments from padded rows (stride = width + 8 floats: the 32 lanes of an A or B
// fragment hit 32 different banks).  Same work units, same partial layout → same reduce kernel.
constexpr int kStageRows = 32;
// Units split BY LAYER between the two CTAs of a slab, so each stages only the four arrays its layers touch:
// half 0: layer 2 (dz2, h1) + heads (dout, h3) = 9 units; half 1: layer 3 (dz3, h2) + layer 1 (dz1, xn) = 12 units (shipped network).
template <class D>
__device__ inline int build_units_by_layer(const D& d, int warp, int half, WUnit (&u)[kMaxUnits]) {
    const int outs[4] = {d.h1, d.h2, d.h3, kOutPad}, ins[4] = {d.in_pad, d.h1, d.h2, d.h3};
    const int layers[2][2] = {{1, 3}, {2, 0}};
    int n = 0, q = 0;
    for (int k = 0; k < 2; ++k) {
        const int l = layers[half][k];
        for (int o0 = 0; o0 < outs[l]; o0 += 16)
            for (int i0 = 0; i0 < ins[l]; i0 += 64, ++q)
                if (q % kWarps == warp && n < kMaxUnits) {
                    u[n].layer = l; u[n].o0 = o0; u[n].i0 = i0; u[n].nt = (ins[l] - i0) >= 64 ? 8 : (ins[l] - i0) / 8; ++n;
                }
    }
    return n;
}
// floats of one stage buffer: the larger of the two halves' four arrays, rows padded by 8 floats
template <class D>
struct StageLayout {
    static constexpr int half0 = kStageRows * ((D::h2 + 8) + (kOutPad + 8) + (D::h1 + 8) + (D::h3 + 8));     // dz2, dout, h1, h3
    static constexpr int half1 = kStageRows * ((D::h3 + 8) + (D::h1 + 8) + (D::h2 + 8) + (D::in_pad + 8));   // dz3, dz1, h2, xn
    static constexpr int floats = half0 > half1 ? half0 : half1;
};
__device__ __forceinline__ void cp_async16(float* dst_smem, const float* src, bool valid) {
    const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst_smem));
    const int sz = valid ? 16 : 0;  // src-size 0: the 16 bytes are zero-filled (rows past the end of the slab)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
template <class D>
__global__ void __launch_bounds__(kWarps * 32, 2)
agx_mlp_wgrad_staged_kernel(const __grid_constant__ AgxMlpParams P, int64_t B, int64_t rows_per_cta, const float* __restrict__ xn,
                            const float* __restrict__ h1, const float* __restrict__ h2, const float* __restrict__ h3,
                            const float* __restrict__ dz1, const float* __restrict__ dz2, const float* __restrict__ dz3,
                            const float* __restrict__ dout, float* __restrict__ partials, int partial_floats) {
    using L = StageLayout<D>;
    const D d = dims_of<D>(P);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const float* srcs[8] = {dz1, dz2, dz3, dout, xn, h1, h2, h3};
    const int widths[8] = {D::h1, D::h2, D::h3, kOutPad, D::in_pad, D::h1, D::h2, D::h3};
    const int outs[4] = {d.h1, d.h2, d.h3, kOutPad}, ins[4] = {d.in_pad, d.h1, d.h2, d.h3};
    WUnit un[kMaxUnits];
    const int slab = blockIdx.x / kWgradSplit, half = blockIdx.x % kWgradSplit;
    const int n_units = build_units_by_layer(d, warp, half, un);
    float acc[kMaxUnits][8][4];
#pragma unroll
    for (int q = 0; q < kMaxUnits; ++q)
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc[q][j][0] = 0.0f; acc[q][j][1] = 0.0f; acc[q][j][2] = 0.0f; acc[q][j][3] = 0.0f; }
    const int64_t r_begin = (int64_t)slab * rows_per_cta;
    int64_t r_end = r_begin + rows_per_cta;
    if (r_end > B) r_end = B;
    const int n_stages = (int)((r_end - r_begin + kStageRows - 1) / kStageRows);
    // arrays this half needs, as indices into srcs: [dz of layer A, dz of layer B, activation of A, activation of B]
    const int need[2][4] = {{1, 3, 5, 7}, {2, 0, 6, 4}};  // half 0: dz2, dout, h1, h3 (layers 2 + heads); half 1: dz3, dz1, h2, xn
    const int layer_a = half == 0 ? 1 : 2;
    int offs[4];
    offs[0] = 0;
#pragma unroll
    for (int k = 1; k < 4; ++k) offs[k] = offs[k - 1] + kStageRows * (widths[need[half][k - 1]] + 8);

    auto issue = [&](int s) {  // stage s → buffer s & 1
        float* buf = g_smem + (s & 1) * L::floats;
        const int64_t r0 = r_begin + (int64_t)s * kStageRows;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int a = need[half][k];
            const int w4 = widths[a] / 4, chunks = kStageRows * w4;
            for (int c = threadIdx.x; c < chunks; c += kWarps * 32) {
                const int row = c / w4, col = (c - row * w4) * 4;
                const bool ok = r0 + row < r_end;
                cp_async16(buf + offs[k] + row * (widths[a] + 8) + col, srcs[a] + (ok ? (r0 + row) * widths[a] + col : 0), ok);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    if (n_stages > 0) issue(0);
    for (int s = 0; s < n_stages; ++s) {
        if (s + 1 < n_stages) {
            issue(s + 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const float* buf = g_smem + (s & 1) * L::floats;
#pragma unroll
        for (int kk = 0; kk < kStageRows; kk += 8) {
#pragma unroll
            for (int q = 0; q < kMaxUnits; ++q) {
                if (q < n_units) {
                    const int l = un[q].layer, ldo = outs[l] + 8, ldi = ins[l] + 8;
                    const int idx = (l == layer_a) ? 0 : 1;
                    const float* dz = buf + offs[idx] + kk * ldo + un[q].o0;
                    const float* a_ = buf + offs[2 + idx] + kk * ldi + un[q].i0;
                    float a[4];
                    a[0] = tf32r(dz[(t)*ldo + g]);       a[1] = tf32r(dz[(t)*ldo + g + 8]);
                    a[2] = tf32r(dz[(t + 4) * ldo + g]); a[3] = tf32r(dz[(t + 4) * ldo + g + 8]);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (j < un[q].nt) {
                            const float b0 = tf32r(a_[t * ldi + 8 * j + g]), b1 = tf32r(a_[(t + 4) * ldi + 8 * j + g]);
                            mma_tf32(acc[q][j], a, b0, b1);
                        }
                    }
                }
            }
        }
        __syncthreads();  // everyone is done with this buffer before stage s+2 overwrites it
    }
    float* mine = partials + (int64_t)slab * partial_floats;
#pragma unroll
    for (int q = 0; q < kMaxUnits; ++q) {
        if (q < n_units) {
            const int l = un[q].layer;
            int base = 0;
            for (int m = 0; m < l; ++m) base += outs[m] * ins[m];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (j < un[q].nt) {
                    const int c = un[q].i0 + 8 * j + 2 * t;
                    *reinterpret_cast<float2*>(&mine[base + (un[q].o0 + g) * ins[l] + c]) = make_float2(acc[q][j][0], acc[q][j][1]);
                    *reinterpret_cast<float2*>(&mine[base + (un[q].o0 + g + 8) * ins[l] + c]) = make_float2(acc[q][j][2], acc[q][j][3]);
                }
            }
        }
    }
}

// deterministic reduction of the per-CTA partials + scatter into the caller's parameter-gradient tensors.
// A CTA owns 32 consecutive elements; warp w adds partials w, w + 8, w + 16, ... (independent coalesced 128-B loads, ~19 per
// thread at 148 slabs instead of a 148-long dependent chain), then warp 0 adds the 8 group sums in group order.
constexpr int kRedGroups = 8;
__global__ void __launch_bounds__(32 * kRedGroups)
agx_mlp_wgrad_reduce_kernel(const __grid_constant__ AgxMlpParams P, const __grid_constant__ AgxMlpGrads G,
                            const float* __restrict__ partials, int n_cta_w, int partial_floats,
                            const float* __restrict__ bias_partials, int n_cta_b) {
    __shared__ float s_sum[kRedGroups][32];
    const Dims d = dims_of<Dims>(P);
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int e = blockIdx.x * 32 + lane;
    const bool is_w = e < partial_floats, is_b = !is_w && e - partial_floats < kBiasSlots;
    const float* src = is_w ? partials + e : bias_partials + (e - partial_floats);
    const int64_t pitch = is_w ? partial_floats : kBiasSlots;
    const int count = is_w ? n_cta_w : (is_b ? n_cta_b : 0);
    float s = 0.0f;
#pragma unroll 8
    for (int c = grp; c < count; c += kRedGroups) s += src[(int64_t)c * pitch];
    s_sum[grp][lane] = s;
    __syncthreads();
    if (grp != 0) return;
    s = 0.0f;
#pragma unroll
    for (int k = 0; k < kRedGroups; ++k) s += s_sum[k][lane];
    if (is_w) {
        const int outs[4] = {d.h1, d.h2, d.h3, kOutPad}, ins[4] = {d.in_pad, d.h1, d.h2, d.h3}, real_in[4] = {d.in_dim, d.h1, d.h2, d.h3};
        float* gw[3] = {G.gw1, G.gw2, G.gw3};
        int l = 0, base = 0;
        while (l < 3 && e >= base + outs[l] * ins[l]) { base += outs[l] * ins[l]; ++l; }
        const int o = (e - base) / ins[l], i = (e - base) % ins[l];
        if (i < real_in[l]) {
            if (l < 3) gw[l][o * real_in[l] + i] = s;
            else if (o < d.a) G.gw_mu[o * d.h3 + i] = s;
            else if (o == d.a) G.gw_value[i] = s;
        }
    } else if (is_b) {
        const int i = e - partial_floats;
        const int seg = i / kMaxW, c = i % kMaxW;
        if (seg == 0 && c < d.h1) G.gb1[c] = s;
        else if (seg == 1 && c < d.h2) G.gb2[c] = s;
        else if (seg == 2 && c < d.h3) G.gb3[c] = s;
        else if (seg == 3) { if (c < d.a) G.gb_mu[c] = s; else if (c == d.a) G.gb_value[0] = s; }
    }
}

bool valid(const AgxMlpParams* p) {
    return p && p->in_dim > 0 && p->in_pad >= p->in_dim && p->in_pad % 16 == 0 && p->in_pad <= kMaxW && p->h1 % 32 == 0 &&
           p->h2 % 32 == 0 && p->h3 % 32 == 0 && p->h1 > 0 && p->h2 > 0 && p->h3 > 0 && p->h1 <= kMaxW && p->h2 <= kMaxW &&
           p->h3 <= kMaxW && (p->actions_num == 4 || p->actions_num == 5) && p->w1 && p->b1 && p->w2 && p->b2 && p->w3 &&
           p->b3 && p->w_mu && p->b_mu && p->w_value && p->b_value;
}
int partial_floats(const AgxMlpParams* p) {  // dense padded dW of all four layers
    return p->h1 * p->in_pad + p->h2 * p->h1 + p->h3 * p->h2 + kOutPad * p->h3;
}
bool units_fit(const AgxMlpParams* p) {  // every warp's work-unit list must fit kMaxUnits
    const int outs[4] = {p->h1, p->h2, p->h3, kOutPad}, ins[4] = {p->in_pad, p->h1, p->h2, p->h3};
    int q = 0;
    for (int l = 0; l < 4; ++l) q += (outs[l] / 16) * ((ins[l] + 63) / 64);
    return (q + kWgradSplit * kWarps - 1) / (kWgradSplit * kWarps) <= kMaxUnits;
}
size_t smem_bytes(const AgxMlpParams* p, int pad) {
    const Dims d = dims_of<Dims>(*p);
    const size_t f = (size_t)weights_floats(d, pad) + (size_t)kWarps * 2 * kRows * kActLd;  // weights + per-warp ping-pong tiles
    const size_t red = (size_t)kWarps * kBiasSlots;
    return sizeof(float) * (f > red ? f : red);
}
using S32 = SDims<32, 64, 128, 64>;  // hovering / balloon (18 obs) with the shipped [64,128,64] MLP
using S48 = SDims<48, 64, 128, 64>;  // tracking (48 obs)
bool is_shipped(const AgxMlpParams* p, int in_pad) { return p->in_pad == in_pad && p->h1 == 64 && p->h2 == 128 && p->h3 == 64; }
constexpr int kGridMax = 148;
int g_fwd_tc = 2;        // tcgen05 forward: 0 off (mma.sync kernel), 1 inference calls only, 2 always (default) — agx_set_option("mlp_forward")
int g_wgrad_staged = 1;  // agx_set_option("mlp_wgrad_staged") switches the staged weight-gradient kernel off/on (A/B)
constexpr int kWgradGrid = 148;  // batch slabs; each slab is walked by kWgradSplit CTAs
unsigned grid_for(int64_t B) {
    const int64_t tiles = (B + kRows - 1) / kRows;
    int64_t g = (tiles + kWarps - 1) / kWarps;
    if (g > kGridMax) g = kGridMax;  // one CTA per SM (shared-memory bound), persistent over tiles
    return (unsigned)(g < 1 ? 1 : g);
}

}  // namespace

extern "C" {

// agx_set_option keys owned by this translation unit: returns 1 when handled, 0 for an unknown key, -1 for a bad value
int agx_internal_mlp_option(const char* key, int value) {
    if (!strcmp(key, "mlp_forward")) {  // 0 TF32 mma.sync kernel, 1 tcgen05 for inference calls only, 2 tcgen05 always (default)
        if (value < 0 || value > 2) return -1;
        g_fwd_tc = value;
        return 1;
    }
    if (!strcmp(key, "mlp_wgrad_staged")) { g_wgrad_staged = value ? 1 : 0; return 1; }
    if (!strcmp(key, "mlp_wgrad_tma")) { g_wgrad_tma = value ? 1 : 0; return 1; }
    return 0;
}

static int mlp_forward_impl(const AgxMlpParams* p, int64_t b, const float* obs, float* mu, float* value, float* xn_out,
                            float* h1_out, float* h2_out, float* h3_out, void* stream, const int keep_t, const AgxPolicyIO* pol_in = nullptr) {
    if (!valid(p) || b <= 0 || !obs || (!pol_in && (!mu || !value))) return agx_internal_fail(AGX_ERR_ARG, "agx_mlp_forward: bad argument");
    AgxPolicyIO pol_none;
    memset(&pol_none, 0, sizeof(pol_none));
    const AgxPolicyIO* pol = pol_in ? pol_in : &pol_none;
    const size_t smem = smem_bytes(p, 4);  // of the mma.sync kernels; the tcgen05 kernels carve their own
    const bool tc_shape = is_shipped(p, 32) || is_shipped(p, 48) || is_shipped(p, 64) || is_shipped(p, 96);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define AGX_FWD(D)                                                                                                      \
    do {                                                                                                                \
        cudaFuncSetAttribute(agx_mlp_forward_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);        \
        agx_mlp_forward_kernel<D><<<grid_for(b), kWarps * 32, smem, st>>>(*p, b, obs, mu, value, xn_out, h1_out, h2_out, h3_out); \
    } while (0)
#define AGX_FWD_TC(PAD)                                                                                                                  \
    do {                                                                                                                                 \
        constexpr int kSm = (int)tc::smem_bytes_tc<PAD>();                                                                               \
        cudaFuncSetAttribute(tc::agx_mlp_forward_tc_kernel<PAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSm);                      \
        const int64_t pairs = ((b + tc::kM - 1) / tc::kM + 1) / 2;                                                                       \
        tc::agx_mlp_forward_tc_kernel<PAD><<<(unsigned)(pairs < kGridMax ? pairs : kGridMax), tc::kThreads, kSm, st>>>(*p, b, obs, mu, value,  \
                                                                                                                       xn_out, h1_out, h2_out, h3_out, keep_t, *pol); \
    } while (0)
    const bool keep_aligned = !xn_out || (((uintptr_t)xn_out | (uintptr_t)h1_out | (uintptr_t)h2_out | (uintptr_t)h3_out) & 15u) == 0;
    // measured (scripts/mlp_bench.py, B200, 32 768 / 65 536 rows): tcgen05 20.6 / 33.5 us vs mma.sync 32.9 / 58.7 us without the kept
    // activations (rollout), 28.4 / 49.0 vs 33.5 / 63.1 us with them (update)
    const bool w_aligned = (((uintptr_t)p->w2 | (uintptr_t)p->w3) & 15u) == 0;  // the tcgen05 kernel stages W2 / W3 with 16-byte loads
    const bool use_tc = keep_aligned && w_aligned && (g_fwd_tc == 2 || (g_fwd_tc == 1 && !xn_out));
    if (pol_in) {  // rollout step: the policy head lives in the tcgen05 kernel's epilogue only
        if (!(w_aligned && (is_shipped(p, 32) || is_shipped(p, 48) || is_shipped(p, 64) || is_shipped(p, 96))))
            return agx_internal_fail(AGX_ERR_UNSUPPORTED, "agx_policy_step: needs the 64-128-64 network with in_pad in {32,48,64,96}");
        if (p->in_pad == 32) AGX_FWD_TC(32); else if (p->in_pad == 48) AGX_FWD_TC(48); else if (p->in_pad == 64) AGX_FWD_TC(64); else AGX_FWD_TC(96);
    }
    else if (keep_t) {  // training path: feature-major keeps for the tensor-core backward / weight-gradient kernels (agx_mlp_train.cu)
        if (!(keep_aligned && w_aligned && xn_out && h1_out && h2_out && h3_out && (b % tc::kM) == 0 && p->in_dim < p->in_pad &&
              (is_shipped(p, 32) || is_shipped(p, 48) || is_shipped(p, 64) || is_shipped(p, 96))))
            return agx_internal_fail(AGX_ERR_UNSUPPORTED, "agx_mlp_forward_train: needs the 64-128-64 network, in_pad in {32,48,64,96} > in_dim, b % 128 == 0");
        if (p->in_pad == 32) AGX_FWD_TC(32); else if (p->in_pad == 48) AGX_FWD_TC(48); else if (p->in_pad == 64) AGX_FWD_TC(64); else AGX_FWD_TC(96);
    }
    else if (use_tc && tc_shape) {
        if (p->in_pad == 32) AGX_FWD_TC(32); else if (p->in_pad == 48) AGX_FWD_TC(48); else if (p->in_pad == 64) AGX_FWD_TC(64); else AGX_FWD_TC(96);
    }
    else if (smem > 227 * 1024) return agx_internal_fail(AGX_ERR_UNSUPPORTED, "agx_mlp_forward: network too large for shared memory");
    else if (is_shipped(p, 32)) AGX_FWD(S32);
    else if (is_shipped(p, 48)) AGX_FWD(S48);
    else AGX_FWD(Dims);
#undef AGX_FWD
#undef AGX_FWD_TC
    return cudaGetLastError() == cudaSuccess ? AGX_OK : agx_internal_fail(AGX_ERR_CUDA, "agx_mlp_forward: launch failed");
}

int agx_mlp_forward(const AgxMlpParams* p, int64_t b, const float* obs, float* mu, float* value, float* xn_out,
                    float* h1_out, float* h2_out, float* h3_out, void* stream) {
    return mlp_forward_impl(p, b, obs, mu, value, xn_out, h1_out, h2_out, h3_out, stream, 0);
}

int agx_mlp_forward_train(const AgxMlpParams* p, int64_t b, const float* obs, float* mu, float* value, float* xt, float* h1t,
                          float* h2t, float* h3t, void* stream) {
    return mlp_forward_impl(p, b, obs, mu, value, xt, h1t, h2t, h3t, stream, 1);
}

int agx_sizeof_policy_io(void) { return (int)sizeof(AgxPolicyIO); }

int agx_policy_step(const AgxMlpParams* p, const AgxPolicyIO* io, int64_t n, const float* obs, void* stream) {
    if (!io || !io->logstd || !io->actions || !io->mus || !io->sigmas || !io->neglogp || !io->values || !io->env_actions ||
        (io->dones_out && !io->dones_in) || ((io->act_lo == nullptr) != (io->act_hi == nullptr)) || ((io->value_mean == nullptr) != (io->value_var == nullptr)))
        return agx_internal_fail(AGX_ERR_ARG, "agx_policy_step: bad argument");
    return mlp_forward_impl(p, n, obs, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, stream, 0, io);
}

// launcher of the partial-sum reduction for agx_mlp_train.cu (same partial layout as the mma.sync weight-gradient kernels)
int agx_internal_wgrad_reduce(const AgxMlpParams* p, const AgxMlpGrads* g, const float* w_partials, int n_cta_w, const float* b_partials,
                              int n_cta_b, void* stream) {
    const int pf = partial_floats(p);
    const unsigned gr = (unsigned)((pf + kBiasSlots + 31) / 32);
    agx_mlp_wgrad_reduce_kernel<<<gr, 32 * kRedGroups, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*p, *g, w_partials, n_cta_w, pf, b_partials, n_cta_b);
    return cudaGetLastError() == cudaSuccess ? AGX_OK : agx_internal_fail(AGX_ERR_CUDA, "wgrad reduce: launch failed");
}

int64_t agx_mlp_workspace_floats(const AgxMlpParams* p) {
    if (!valid(p)) return -1;
    return (int64_t)kWgradGrid * partial_floats(p) + (int64_t)kGridMax * kBiasSlots;
}

int agx_mlp_backward(const AgxMlpParams* p, const AgxMlpGrads* g, int64_t b, const float* grad_mu, const float* grad_value,
                     const float* xn, const float* h1, const float* h2, const float* h3, float* dz1, float* dz2, float* dz3,
                     float* dout, float* workspace, void* stream) {
    if (!valid(p) || !units_fit(p) || !g || b <= 0 || (b % 8) != 0 || !grad_mu || !grad_value || !xn || !h1 || !h2 || !h3 || !dz1 ||
        !dz2 || !dz3 || !dout || !workspace || !g->gw1 || !g->gb1 || !g->gw2 || !g->gb2 || !g->gw3 || !g->gb3 || !g->gw_mu ||
        !g->gb_mu || !g->gw_value || !g->gb_value)
        return agx_internal_fail(AGX_ERR_ARG, "agx_mlp_backward: bad argument (batch must be a multiple of 8)");
    const size_t smem = smem_bytes(p, 8);
    if (smem > 227 * 1024) return agx_internal_fail(AGX_ERR_UNSUPPORTED, "agx_mlp_backward: network too large for shared memory");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int pf = partial_floats(p);
    float* w_partials = workspace;
    float* b_partials = workspace + (int64_t)kWgradGrid * pf;
    const unsigned gb = grid_for(b);
    int64_t rows = (b + kWgradGrid - 1) / kWgradGrid;
    rows = (rows + 7) / 8 * 8;
    const unsigned n_slabs = (unsigned)((b + rows - 1) / rows), gw = n_slabs * kWgradSplit;
#define AGX_BWD(D)                                                                                                       \
    do {                                                                                                                 \
        cudaFuncSetAttribute(agx_mlp_backward_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);        \
        agx_mlp_backward_kernel<D><<<gb, kWarps * 32, smem, st>>>(*p, b, grad_mu, grad_value, h1, h2, h3, dz1, dz2, dz3, dout, b_partials); \
        AGX_WGRAD(D);                                                                                                    \
    } while (0)
#define AGX_WGRAD(D) agx_mlp_wgrad_kernel<D><<<gw, kWarps * 32, 0, st>>>(*p, b, rows, xn, h1, h2, h3, dz1, dz2, dz3, dout, w_partials, pf)
#define AGX_WGRAD_STAGED(D)                                                                                              \
    do {                                                                                                                 \
        constexpr int kSm = 2 * StageLayout<D>::floats * (int)sizeof(float);                                             \
        cudaFuncSetAttribute(agx_mlp_wgrad_staged_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSm);          \
        agx_mlp_wgrad_staged_kernel<D><<<gw, kWarps * 32, kSm, st>>>(*p, b, rows, xn, h1, h2, h3, dz1, dz2, dz3, dout, w_partials, pf); \
    } while (0)
    if (is_shipped(p, 32)) {
        cudaFuncSetAttribute(agx_mlp_backward_kernel<S32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        agx_mlp_backward_kernel<S32><<<gb, kWarps * 32, smem, st>>>(*p, b, grad_mu, grad_value, h1, h2, h3, dz1, dz2, dz3, dout, b_partials);
        if (g_wgrad_staged) AGX_WGRAD_STAGED(S32); else AGX_WGRAD(S32);
    } else if (is_shipped(p, 48)) {
        cudaFuncSetAttribute(agx_mlp_backward_kernel<S48>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        agx_mlp_backward_kernel<S48><<<gb, kWarps * 32, smem, st>>>(*p, b, grad_mu, grad_value, h1, h2, h3, dz1, dz2, dz3, dout, b_partials);
        if (g_wgrad_staged) AGX_WGRAD_STAGED(S48); else AGX_WGRAD(S48);
    } else AGX_BWD(Dims);
#undef AGX_BWD
#undef AGX_WGRAD
#undef AGX_WGRAD_STAGED
    const unsigned gr = (unsigned)((pf + kBiasSlots + 31) / 32);
    agx_mlp_wgrad_reduce_kernel<<<gr, 32 * kRedGroups, 0, st>>>(*p, *g, w_partials, (int)n_slabs, pf, b_partials, (int)gb);
    return cudaGetLastError() == cudaSuccess ? AGX_OK : agx_internal_fail(AGX_ERR_CUDA, "agx_mlp_backward: launch failed");
}

}  // extern "C"

// agx_cnn.cuh — depth-image encoder of the Avoid / Planning policies as per-thread phase functions.
//
// Replaces, for inference (eval-mode BatchNorm — the only mode the trainer uses):
//   lib/network/cnn.py:3-33   CNNFeatureExtractor: Conv(1→16,5x5,s2,p2) ReLU BN → Conv(16→32,3x3,s2,p1) ReLU BN →
//                             Conv(32→64,3x3,s2,p1) ReLU BN → AdaptiveAvgPool(1,1) → Linear(64→feature_dim)
//   lib/core/running_mean_std.py:62-81 the per-pixel input normalisation clamp((x-mean)/sqrt(var+eps), ±5) in front of it
//
// One CTA encodes one env at a time, end to end in shared memory, in kStrips horizontal strips of kG conv3 rows: the strip's
// image rows → conv1 rows → conv2 rows → conv3 rows, whose ReLU outputs are only ever needed as per-channel sums (the average
// pool commutes with the eval-mode BatchNorm affine).  No activation leaves the SM: HBM traffic is the 101 760-byte image in
// and feature_dim floats out.  Strips overlap (conv1 rows are computed 15 per 12 new ones, conv2 rows 7 per 6): 19.4 M FMA per
// env instead of 17.3 M — cheaper than a ring buffer's index arithmetic.
//
// Every function below takes the thread / task index as an argument and touches only plain float arrays, so the same text is
// compiled by nvcc for the kernel (agx_cnn.cu) and by g++ for the CPU emulation the tests run (tests/hostsim/hostsim_cnn.cpp),
// where each phase is a sequential loop over the task index — exact, because tasks of one phase write disjoint locations.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define AGXC_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#define AGXC_HD inline
#endif

namespace agxcnn {

constexpr int kImgH = 212, kImgW = 120;        // the image tensor is [N,1,212,120]: 212 rows of 120 (customized.py:62)
constexpr int kC1 = 16, kH1 = 106, kW1 = 60;   // conv1 5x5 stride 2 pad 2
constexpr int kC2 = 32, kH2 = 53, kW2 = 30;    // conv2 3x3 stride 2 pad 1
constexpr int kC3 = 64, kH3 = 27, kW3 = 15;    // conv3 3x3 stride 2 pad 1
constexpr int kMaxFeat = 64;

constexpr int kG = 3;                 // conv3 rows per strip
constexpr int kStrips = kH3 / kG;     // 9 strips cover the 27 rows exactly
constexpr int kR2 = 2 * kG + 1;       // conv2 rows a strip needs (7): local row jl <-> conv2 row 2*i0 - 1 + jl
constexpr int kR1 = 2 * kR2 + 1;      // conv1 rows (15): local kl <-> conv1 row 4*i0 - 3 + kl
constexpr int kR0 = 2 * kR1 + 3;      // image rows (33): local rl <-> image row 8*i0 - 8 + rl
static_assert(kStrips * kG == kH3, "strips must tile conv3");
// with these offsets every layer reads local input rows 2*r + ky for its local output row r

// padded row lengths (floats).  Column x of a layer lives at x + pad, pad = 2 / 4 / 2, so that a task's output pixels start
// on a 16- / 8-byte boundary (vector stores); the pad columns are zero.
constexpr int kLd0 = 124;  // image strip: 2 + 120 + 2; a conv1 task reads 12 floats from 8*xg (max 8*14 + 11 = 123)
constexpr int kPad1 = 4;
constexpr int kLd1 = 64;   // conv1 strip: 4 + 60; a conv2 task reads column 4*xg + 3 and the aligned float4 at 4*xg + 4 (max 63)
constexpr int kPad2 = 2;
constexpr int kLd2 = 34;   // conv2 strip: 2 + 30 + 2; a conv3 task reads 7 floats from 6*xg + 1 (max 31)
constexpr int kLdP = 68;   // conv3 partial sums: 64 channels per pixel + 4, so the 16-byte stores of a warp's 15 tiles spread over the banks

// task grids (one task = one thread's register tile): pixels-per-task P along a row x C output channels
constexpr int kP1 = 4, kCt = 8;
constexpr int kTasks1 = kR1 * (kW1 / kP1) * (kC1 / kCt);                 // 15 * 15 * 2 = 450
constexpr int kP2 = 2;
constexpr int kTasks2 = kR2 * (kW2 / kP2) * (kC2 / kCt);                 // 7 * 15 * 4 = 420
constexpr int kP3 = 3, kKs = 4;                                           // conv3 also splits its 32 input channels 4 ways
constexpr int kTiles3 = kG * (kW3 / kP3);                                 // 15 pixel tiles
constexpr int kTasks3 = kTiles3 * (kC3 / kCt) * kKs;                      // 15 * 8 * 4 = 480
constexpr int kPix3 = kG * kW3;                                           // 45 conv3 pixels per strip
constexpr int kThreads = 512;
constexpr int kPoolSlices = kThreads / kC3;                               // 8
// A warp should hold ONE channel group, so that its weight loads are single-address broadcasts: the tasks of a channel group
// are padded to whole warps (conv1: 225 -> 256 slots x 2 groups, conv2: 105 -> 128 slots x 4 groups = the 512 threads).
constexpr int kSlots1 = 256, kSlots2 = 128;
static_assert(kSlots1 * (kC1 / kCt) == kThreads && kSlots2 * (kC2 / kCt) == kThreads && kTasks3 <= kThreads, "one round per phase");
static_assert(kSlots1 >= kR1 * (kW1 / kP1) && kSlots2 >= kR2 * (kW2 / kP2), "slots cover the tasks");

// shared-memory map (float offsets)
constexpr int kOffW1 = 0;                                  // [25 taps][16]
constexpr int kOffW2 = kOffW1 + 25 * kC1;                  // [16 ci][9 taps][32]
constexpr int kOffW3 = kOffW2 + kC1 * 9 * kC2;             // [32 ci][9 taps][64]
constexpr int kOffAff = kOffW3 + kC2 * 9 * kC3;            // bias | bn scale | bn shift for the three layers
constexpr int kAff1 = 0, kAff2 = 3 * kC1, kAff3 = 3 * kC1 + 3 * kC2;
constexpr int kOffImg = kOffAff + 3 * (kC1 + kC2 + kC3);   // [33][124]
constexpr int kOffA1 = kOffImg + kR0 * kLd0;               // [16][15][64]; conv3's partial sums [4][45][68] alias it
constexpr int kOffA2 = kOffA1 + kC1 * kR1 * kLd1;          // [32][7][34]
constexpr int kOffPool = ((kOffA2 + kC2 * kR2 * kLd2 + 3) / 4) * 4;  // [8][64] + y[64]
constexpr int kSmemFloats = kOffPool + kPoolSlices * kC3 + kC3;
static_assert(kKs * kPix3 * kLdP <= kC1 * kR1 * kLd1, "partials must fit the conv1 strip they alias");
static_assert(kOffImg % 4 == 0 && kOffA1 % 4 == 0 && kOffA2 % 2 == 0 && kOffW2 % 4 == 0 && kOffW3 % 4 == 0 && kOffAff % 4 == 0, "16-byte aligned regions");
static_assert(kSmemFloats * 4 <= 227 * 1024, "shared memory budget");

struct Weights {  // raw PyTorch parameter tensors (device pointers), CNNFeatureExtractor layout (cnn.py:8-29)
    const float *w1, *b1, *s1, *t1;  // features.0 weight [16,1,5,5] / bias; features.2 folded to scale / shift
    const float *w2, *b2, *s2, *t2;  // features.3 [32,16,3,3]; features.5
    const float *w3, *b3, *s3, *t3;  // features.6 [64,32,3,3]; features.8
    const float *wfc, *bfc;          // fc [feature_dim, 64]
};

struct F4 { float x, y, z, w; };
AGXC_HD F4 ld4(const float* p) {
#if defined(__CUDA_ARCH__)
    const float4 v = *reinterpret_cast<const float4*>(p);
    F4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w; return r;
#else
    F4 r; r.x = p[0]; r.y = p[1]; r.z = p[2]; r.w = p[3]; return r;
#endif
}
AGXC_HD void st4(float* p, float a, float b, float c, float d) {
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
#else
    p[0] = a; p[1] = b; p[2] = c; p[3] = d;
#endif
}
AGXC_HD void ld8(const float* p, float* w) {
    const F4 a = ld4(p), b = ld4(p + 4);
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
}
AGXC_HD float relu_bn(float acc, float bias, float scale, float shift) {
    const float r = acc + bias;
    return fmaf(r > 0.0f ? r : 0.0f, scale, shift);
}

// ---- phase 0 (once per CTA): weights into shared memory, output channel innermost so a task's 8 channels are two 16-byte loads --
AGXC_HD void stage_weights(int tid, int nthreads, const Weights& W, float* sm) {
    for (int i = tid; i < 25 * kC1; i += nthreads) { const int co = i % kC1, tap = i / kC1; sm[kOffW1 + i] = W.w1[co * 25 + tap]; }
    for (int i = tid; i < kC1 * 9 * kC2; i += nthreads) {
        const int co = i % kC2, tap = (i / kC2) % 9, ci = i / (kC2 * 9);
        sm[kOffW2 + i] = W.w2[(co * kC1 + ci) * 9 + tap];
    }
    for (int i = tid; i < kC2 * 9 * kC3; i += nthreads) {
        const int co = i % kC3, tap = (i / kC3) % 9, ci = i / (kC3 * 9);
        sm[kOffW3 + i] = W.w3[(co * kC2 + ci) * 9 + tap];
    }
    float* aff = sm + kOffAff;
    for (int i = tid; i < kC1; i += nthreads) { aff[kAff1 + i] = W.b1[i]; aff[kAff1 + kC1 + i] = W.s1[i]; aff[kAff1 + 2 * kC1 + i] = W.t1[i]; }
    for (int i = tid; i < kC2; i += nthreads) { aff[kAff2 + i] = W.b2[i]; aff[kAff2 + kC2 + i] = W.s2[i]; aff[kAff2 + 2 * kC2 + i] = W.t2[i]; }
    for (int i = tid; i < kC3; i += nthreads) { aff[kAff3 + i] = W.b3[i]; aff[kAff3 + kC3 + i] = W.s3[i]; aff[kAff3 + 2 * kC3 + i] = W.t3[i]; }
    for (int i = tid; i < kC2 * kR2 * kLd2; i += nthreads) sm[kOffA2 + i] = 0.0f;  // conv2 strip: the pad columns stay zero
}

// ---- phase 1: the strip's 33 image rows, normalised, zero outside the image -------------------------------------------------------
AGXC_HD void load_image_strip(int tid, int nthreads, const float* img, const float* px_mean, const float* px_rstd, int strip,
                              float* sm) {
    const int row0 = 8 * strip * kG - 8;
    for (int i = tid; i < kR0 * kLd0; i += nthreads) {
        const int rl = i / kLd0, xp = i - rl * kLd0;
        const int y = row0 + rl, x = xp - 2;
        float v = 0.0f;
        if (y >= 0 && y < kImgH && x >= 0 && x < kImgW) {
            v = img[y * kImgW + x];
            if (px_mean) {
                v = (v - px_mean[y * kImgW + x]) * px_rstd[y * kImgW + x];
                v = v < -5.0f ? -5.0f : (v > 5.0f ? 5.0f : v);
            }
        }
        sm[kOffImg + i] = v;
    }
}

// ---- phase 2: conv1 + ReLU + BN for 15 rows; thread = (8-channel group, slot), slot = (row, 4-pixel group) -------------------------
AGXC_HD void conv1_task(int tid, int strip, float* sm) {
    const int cg = tid / kSlots1, slot = tid % kSlots1;
    if (slot >= kR1 * (kW1 / kP1)) return;
    const int xg = slot % (kW1 / kP1), r = slot / (kW1 / kP1);
    float acc[kP1][kCt];
#pragma unroll
    for (int p = 0; p < kP1; ++p)
#pragma unroll
        for (int c = 0; c < kCt; ++c) acc[p][c] = 0.0f;
#pragma unroll
    for (int ky = 0; ky < 5; ++ky) {
        const float* row = sm + kOffImg + (2 * r + ky) * kLd0 + 2 * kP1 * xg;
        float v[12];
        const F4 a = ld4(row), b = ld4(row + 4), c4 = ld4(row + 8);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        v[8] = c4.x; v[9] = c4.y; v[10] = c4.z; v[11] = c4.w;
#pragma unroll
        for (int kx = 0; kx < 5; ++kx) {
            float w[kCt];
            ld8(sm + kOffW1 + (ky * 5 + kx) * kC1 + cg * kCt, w);
#pragma unroll
            for (int p = 0; p < kP1; ++p)
#pragma unroll
                for (int c = 0; c < kCt; ++c) acc[p][c] = fmaf(v[2 * p + kx], w[c], acc[p][c]);
        }
    }
    const int k = 4 * strip * kG - 3 + r;  // conv1 row of this local row; rows outside the layer are conv2's zero padding
    const bool valid = k >= 0 && k < kH1;
    const float* aff = sm + kOffAff + kAff1;
#pragma unroll
    for (int c = 0; c < kCt; ++c) {
        const int ch = cg * kCt + c;
        float o[kP1];
#pragma unroll
        for (int p = 0; p < kP1; ++p) o[p] = valid ? relu_bn(acc[p][c], aff[ch], aff[kC1 + ch], aff[2 * kC1 + ch]) : 0.0f;
        st4(sm + kOffA1 + (ch * kR1 + r) * kLd1 + kPad1 + kP1 * xg, o[0], o[1], o[2], o[3]);
    }
}
// the conv1 strip's pad columns (0..3) are rewritten every strip: conv3's partial sums alias the region
AGXC_HD void conv1_pads(int tid, int nthreads, float* sm) {
    for (int line = tid; line < kC1 * kR1; line += nthreads) st4(sm + kOffA1 + line * kLd1, 0.0f, 0.0f, 0.0f, 0.0f);
}

// ---- phase 3: conv2 + ReLU + BN for 7 rows; thread = (8-channel group, slot), slot = (row, 2-pixel group) -------------------------
AGXC_HD void st2(float* p, float a, float b) {
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<float2*>(p) = make_float2(a, b);
#else
    p[0] = a; p[1] = b;
#endif
}
AGXC_HD void conv2_task(int tid, int strip, float* sm) {
    const int cg = tid / kSlots2, slot = tid % kSlots2;
    if (slot >= kR2 * (kW2 / kP2)) return;
    const int xg = slot % (kW2 / kP2), r = slot / (kW2 / kP2);
    float acc[kP2][kCt];
#pragma unroll
    for (int p = 0; p < kP2; ++p)
#pragma unroll
        for (int c = 0; c < kCt; ++c) acc[p][c] = 0.0f;
#pragma unroll 2
    for (int ci = 0; ci < kC1; ++ci) {
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const float* row = sm + kOffA1 + (ci * kR1 + 2 * r + ky) * kLd1 + 2 * kP2 * xg + kPad1;  // column 4*xg of conv1
            const F4 a = ld4(row);
            const float v[5] = {row[-1], a.x, a.y, a.z, a.w};
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                float w[kCt];
                ld8(sm + kOffW2 + ((ci * 3 + ky) * 3 + kx) * kC2 + cg * kCt, w);
#pragma unroll
                for (int p = 0; p < kP2; ++p)
#pragma unroll
                    for (int c = 0; c < kCt; ++c) acc[p][c] = fmaf(v[2 * p + kx], w[c], acc[p][c]);
            }
        }
    }
    const int j = 2 * strip * kG - 1 + r;
    const bool valid = j >= 0 && j < kH2;
    const float* aff = sm + kOffAff + kAff2;
#pragma unroll
    for (int c = 0; c < kCt; ++c) {
        const int ch = cg * kCt + c;
        float o[kP2];
#pragma unroll
        for (int p = 0; p < kP2; ++p) o[p] = valid ? relu_bn(acc[p][c], aff[ch], aff[kC2 + ch], aff[2 * kC2 + ch]) : 0.0f;
        st2(sm + kOffA2 + (ch * kR2 + r) * kLd2 + kPad2 + kP2 * xg, o[0], o[1]);
    }
}

// ---- phase 4: conv3 partial sums for 3 rows; task = (3-pixel tile, 8-channel group, 8-input-channel slice) -----------------------
// Tiles vary fastest so a warp holds the 15 tiles of two channel groups: weight loads are 2-address broadcasts.
AGXC_HD void conv3_task(int task, float* sm) {
    const int tile = task % kTiles3, cg = (task / kTiles3) % (kC3 / kCt), ks = task / (kTiles3 * (kC3 / kCt));
    const int r = tile / (kW3 / kP3), xg = tile % (kW3 / kP3);
    float acc[kP3][kCt];
#pragma unroll
    for (int p = 0; p < kP3; ++p)
#pragma unroll
        for (int c = 0; c < kCt; ++c) acc[p][c] = 0.0f;
#pragma unroll 2
    for (int cl = 0; cl < kC2 / kKs; ++cl) {
        const int ci = ks * (kC2 / kKs) + cl;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const float* row = sm + kOffA2 + (ci * kR2 + 2 * r + ky) * kLd2 + 2 * kP3 * xg + kPad2 - 1;
            float v[7];
#pragma unroll
            for (int i = 0; i < 7; ++i) v[i] = row[i];
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                float w[kCt];
                ld8(sm + kOffW3 + ((ci * 3 + ky) * 3 + kx) * kC3 + cg * kCt, w);
#pragma unroll
                for (int p = 0; p < kP3; ++p)
#pragma unroll
                    for (int c = 0; c < kCt; ++c) acc[p][c] = fmaf(v[2 * p + kx], w[c], acc[p][c]);
            }
        }
    }
#pragma unroll
    for (int p = 0; p < kP3; ++p) {
        float* out = sm + kOffA1 + ((ks * kPix3) + r * kW3 + kP3 * xg + p) * kLdP + cg * kCt;
        st4(out, acc[p][0], acc[p][1], acc[p][2], acc[p][3]);
        st4(out + 4, acc[p][4], acc[p][5], acc[p][6], acc[p][7]);
    }
}

// ---- phase 5: bias + ReLU + per-channel sum of the strip's 45 pixels; thread = (channel, pixel slice) ----------------------------
AGXC_HD float pool_strip(int tid, const float* sm) {
    const int c = tid % kC3, q = tid / kC3;
    const float bias = sm[kOffAff + kAff3 + c];
    float sum = 0.0f;
    for (int pix = q; pix < kPix3; pix += kPoolSlices) {
        float v = bias;
#pragma unroll
        for (int ks = 0; ks < kKs; ++ks) v += sm[kOffA1 + (ks * kPix3 + pix) * kLdP + c];
        sum += v > 0.0f ? v : 0.0f;
    }
    return sum;
}

// ---- phase 6: average pool (commuted with the BN affine) and the linear layer -------------------------------------------------------
AGXC_HD void pool_finish(int c, float* sm) {  // c < 64, after every thread stored its slice sum at pool[q][c]
    float s = 0.0f;
#pragma unroll
    for (int q = 0; q < kPoolSlices; ++q) s += sm[kOffPool + q * kC3 + c];
    const float* aff = sm + kOffAff + kAff3;
    sm[kOffPool + kPoolSlices * kC3 + c] = fmaf(s * (1.0f / (float)(kH3 * kW3)), aff[kC3 + c], aff[2 * kC3 + c]);
}
AGXC_HD float fc_row(int f, const Weights& W, const float* sm) {
    float o = W.bfc[f];
    for (int c = 0; c < kC3; ++c) o = fmaf(W.wfc[f * kC3 + c], sm[kOffPool + kPoolSlices * kC3 + c], o);
    return o;
}

}  // namespace agxcnn

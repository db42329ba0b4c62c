// agx_cnn.cu — depth-image encoder kernel for sm_100a behind the C ABI (include/agx.h).
// The per-thread phase functions live in agx_cnn.cuh (shared with the CPU emulation in tests/hostsim); this file is the
// persistent kernel around them — one CTA per SM, weights staged once, a loop over envs — and the entry point.
#include <cuda_runtime.h>
#include <stdint.h>

#include "agx.h"
#include "agx_cnn.cuh"

int agx_internal_fail(int code, const char* msg);  // agx_step.cu: records the message for agx_error_string

namespace {

using namespace agxcnn;

__global__ void __launch_bounds__(kThreads, 1)
agx_cnn_encode_kernel(const __grid_constant__ Weights W, int64_t n, const float* __restrict__ image,
                      const float* __restrict__ px_mean, const float* __restrict__ px_rstd, float* __restrict__ features,
                      int64_t ld_features, int feature_dim) {
    extern __shared__ float4 smem4[];
    float* sm = reinterpret_cast<float*>(smem4);
    const int tid = threadIdx.x;
    stage_weights(tid, kThreads, W, sm);
    __syncthreads();
    for (int64_t env = blockIdx.x; env < n; env += gridDim.x) {
        const float* img = image + env * (int64_t)(kImgH * kImgW);
        float pooled = 0.0f;  // this thread's (channel, pixel slice) share of the ReLU sums
        for (int strip = 0; strip < kStrips; ++strip) {
            load_image_strip(tid, kThreads, img, px_mean, px_rstd, strip, sm);
            __syncthreads();  // image strip complete; the previous strip's partial sums (aliasing the conv1 strip) are consumed
            conv1_task(tid, strip, sm);
            conv1_pads(tid, kThreads, sm);
            __syncthreads();
            conv2_task(tid, strip, sm);
            __syncthreads();
            if (tid < kTasks3) conv3_task(tid, sm);
            __syncthreads();
            pooled += pool_strip(tid, sm);
        }
        sm[kOffPool + (tid / kC3) * kC3 + (tid % kC3)] = pooled;
        __syncthreads();
        if (tid < kC3) pool_finish(tid, sm);
        __syncthreads();
        if (tid < feature_dim) features[env * ld_features + tid] = fc_row(tid, W, sm);
        // the next env's first barrier orders these reads before anything overwrites the pool area
    }
}

}  // namespace

extern "C" {

int agx_sizeof_cnn_params(void) { return (int)sizeof(AgxCnnParams); }

int agx_cnn_encode(const AgxCnnParams* p, int64_t n, const float* image, const float* px_mean, const float* px_rstd,
                   float* features, int64_t ld_features, void* stream) {
    if (!p || n < 0 || !image || !features || !p->w1 || !p->b1 || !p->s1 || !p->t1 || !p->w2 || !p->b2 || !p->s2 || !p->t2 ||
        !p->w3 || !p->b3 || !p->s3 || !p->t3 || !p->wfc || !p->bfc)
        return agx_internal_fail(AGX_ERR_ARG, "agx_cnn_encode: bad argument");
    if (p->feature_dim <= 0 || p->feature_dim > kMaxFeat || ld_features < p->feature_dim)
        return agx_internal_fail(AGX_ERR_ARG, "agx_cnn_encode: feature_dim must be 1..64 and ld_features >= feature_dim");
    if ((px_mean == nullptr) != (px_rstd == nullptr))
        return agx_internal_fail(AGX_ERR_ARG, "agx_cnn_encode: px_mean and px_rstd go together");
    if (n == 0) return AGX_OK;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    constexpr int kSmem = kSmemFloats * (int)sizeof(float);
    if (cudaFuncSetAttribute(agx_cnn_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem) != cudaSuccess)
        return agx_internal_fail(AGX_ERR_CUDA, "agx_cnn_encode: shared-memory opt-in failed");
    Weights W;
    W.w1 = p->w1; W.b1 = p->b1; W.s1 = p->s1; W.t1 = p->t1;
    W.w2 = p->w2; W.b2 = p->b2; W.s2 = p->s2; W.t2 = p->t2;
    W.w3 = p->w3; W.b3 = p->b3; W.s3 = p->s3; W.t3 = p->t3;
    W.wfc = p->wfc; W.bfc = p->bfc;
    const unsigned grid = (unsigned)(n < sms ? n : sms);  // persistent: one CTA per SM (shared-memory bound)
    agx_cnn_encode_kernel<<<grid, kThreads, kSmem, reinterpret_cast<cudaStream_t>(stream)>>>(W, n, image, px_mean, px_rstd, features,
                                                                                           ld_features, p->feature_dim);
    return cudaGetLastError() == cudaSuccess ? AGX_OK : agx_internal_fail(AGX_ERR_CUDA, "agx_cnn_encode: launch failed");
}

}  // extern "C"

// hostsim_cnn.cpp — TEST INFRASTRUCTURE.  Compiles airgym_b200/csrc/agx_cnn.cuh (the exact per-thread phase functions the
// sm_100a encoder kernel runs) with g++ and replays the kernel's schedule on the CPU: a float array stands in for shared
// memory, every phase between two __syncthreads() is a sequential loop over the thread index.  Never loaded by the product.
#include <stdint.h>
#include <vector>
#include "agx.h"
#include "agx_cnn.cuh"

using namespace agxcnn;

extern "C" int hostsim_cnn_encode(const AgxCnnParams* p, int64_t n, const float* image, const float* px_mean,
                                  const float* px_rstd, float* features, int64_t ld_features) {
    Weights W;
    W.w1 = p->w1; W.b1 = p->b1; W.s1 = p->s1; W.t1 = p->t1;
    W.w2 = p->w2; W.b2 = p->b2; W.s2 = p->s2; W.t2 = p->t2;
    W.w3 = p->w3; W.b3 = p->b3; W.s3 = p->s3; W.t3 = p->t3;
    W.wfc = p->wfc; W.bfc = p->bfc;
    std::vector<float> smem(kSmemFloats, -1.0e30f);  // poison: a read of anything the schedule did not write shows up
    float* sm = smem.data();
    for (int tid = 0; tid < kThreads; ++tid) stage_weights(tid, kThreads, W, sm);
    std::vector<float> pooled(kThreads);
    for (int64_t env = 0; env < n; ++env) {
        const float* img = image + env * (int64_t)(kImgH * kImgW);
        for (int tid = 0; tid < kThreads; ++tid) pooled[tid] = 0.0f;
        for (int strip = 0; strip < kStrips; ++strip) {
            for (int tid = 0; tid < kThreads; ++tid) load_image_strip(tid, kThreads, img, px_mean, px_rstd, strip, sm);
            for (int tid = 0; tid < kThreads; ++tid) { conv1_task(tid, strip, sm); conv1_pads(tid, kThreads, sm); }
            for (int tid = 0; tid < kThreads; ++tid) conv2_task(tid, strip, sm);
            for (int tid = 0; tid < kThreads; ++tid) if (tid < kTasks3) conv3_task(tid, sm);
            for (int tid = 0; tid < kThreads; ++tid) pooled[tid] += pool_strip(tid, sm);
        }
        for (int tid = 0; tid < kThreads; ++tid) sm[kOffPool + (tid / kC3) * kC3 + (tid % kC3)] = pooled[tid];
        for (int tid = 0; tid < kC3; ++tid) pool_finish(tid, sm);
        for (int tid = 0; tid < p->feature_dim; ++tid) features[env * ld_features + tid] = fc_row(tid, W, sm);
    }
    return 0;
}

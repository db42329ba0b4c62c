// micro-benchmark: latency / throughput of legacy mma.sync on sm_100a (one warp, clock64)
#include <cstdio>
#include <cuda_bf16.h>
__global__ void k_tf32(float* out, long long* cyc, int iters, int chains) {
    float c[8][4] = {};
    float a0 = 1.f, a1 = 2.f, a2 = 3.f, a3 = 4.f, b0 = threadIdx.x * 0.001f, b1 = 0.5f;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (j < chains)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                             : "r"(__float_as_uint(a0)), "r"(__float_as_uint(a1)), "r"(__float_as_uint(a2)), "r"(__float_as_uint(a3)),
                               "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
    }
    long long t1 = clock64();
    float s = 0; for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
    out[threadIdx.x + blockIdx.x * blockDim.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_bf16(float* out, long long* cyc, int iters, int chains) {
    float c[8][4] = {};
    unsigned a0 = 0x3f803f80u, a1 = a0, a2 = a0, a3 = a0, b0 = 0x3f003f00u + threadIdx.x, b1 = b0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (j < chains)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                             : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    long long t1 = clock64();
    float s = 0; for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
    out[threadIdx.x + blockIdx.x * blockDim.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_ffma(float* out, long long* cyc, int iters) {
    float c[8] = {1, 2, 3, 4, 5, 6, 7, 8}; float a = threadIdx.x * 1e-3f, b = 0.999f;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) c[j] = fmaf(c[j], b, a);
    long long t1 = clock64();
    float s = 0; for (int j = 0; j < 8; ++j) s += c[j];
    out[threadIdx.x] = s; if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMallocManaged(&cyc, 8);
    const int iters = 1000;
    for (int warps : {1, 4, 8}) for (int chains : {1, 4, 8}) {
        k_tf32<<<1, 32 * warps>>>(out, cyc, iters, chains); cudaDeviceSynchronize();
        k_tf32<<<1, 32 * warps>>>(out, cyc, iters, chains); cudaDeviceSynchronize();
        printf("tf32 m16n8k8 : warps %d chains %d : %.1f cycles per mma (per warp)\n", warps, chains, (double)*cyc / (iters * chains));
        k_bf16<<<1, 32 * warps>>>(out, cyc, iters, chains); cudaDeviceSynchronize();
        k_bf16<<<1, 32 * warps>>>(out, cyc, iters, chains); cudaDeviceSynchronize();
        printf("bf16 m16n8k16: warps %d chains %d : %.1f cycles per mma (per warp)\n", warps, chains, (double)*cyc / (iters * chains));
    }
    k_ffma<<<1, 32>>>(out, cyc, iters); cudaDeviceSynchronize();
    printf("ffma 8 chains: %.2f cycles per ffma\n", (double)*cyc / (iters * 8));
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

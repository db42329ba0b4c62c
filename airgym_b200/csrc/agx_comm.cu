// agx_comm.cu — host side + generic kernel of the small-message all-reduce over NVLink peer memory (see agx_comm.cuh for the
// protocol).  Replaces the per-minibatch / per-epoch dist.all_reduce calls of the reference's multi-GPU PPO
// (lib/agent/a2c_base.py:293-309 flat gradients, lib/agent/a2c_continuous.py:112-123 KL) for messages of a few KB to a few
// hundred KB, where NCCL's launch + protocol latency is the whole cost.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "agx.h"
#include "agx_comm.cuh"

int agx_internal_fail(int code, const char* msg);

namespace {

template <typename T>
__global__ void __cluster_dims__(agxc::kCluster, 1, 1) __launch_bounds__(agxc::kBlock)
agx_comm_allreduce_kernel(const __grid_constant__ AgxComm c, T* __restrict__ buf, int64_t n) {
    namespace cg = cooperative_groups;
    const unsigned long long seq = agxc::push<T>(c, buf, n);
    if (sizeof(T) == 8) cg::this_cluster().sync();  // 8-byte elements: the two words of an element are pushed by other threads than the one
                                                    // that overwrites it with the sum (4-byte elements: same thread pushes and overwrites)
    const int64_t first = (int64_t)cg::this_cluster().block_rank() * agxc::kBlock + threadIdx.x, stride = (int64_t)agxc::kCluster * agxc::kBlock;
    unsigned long long t0 = 0ull;
    for (int64_t i = first; i < n; i += stride) buf[i] = agxc::reduce_elem<T>(c, seq, i, t0);
    agxc::finish(c, seq);
}

bool comm_ok(const AgxComm* c) {
    if (!c || c->world < 1 || c->world > AGX_COMM_MAX_RANKS || c->rank < 0 || c->rank >= c->world || c->slot_bytes <= 0 || (c->slot_bytes & 255)) return false;  // slots are 256-byte multiples
    for (int i = 0; i < c->world; ++i)
        if (!c->region[i]) return false;
    return true;
}

}  // namespace

extern "C" {

int64_t agx_comm_region_bytes(int world, int64_t slot_bytes) {
    if (world < 1 || world > AGX_COMM_MAX_RANKS || slot_bytes <= 0) return -1;
    const int64_t sb = (slot_bytes + 255) / 256 * 256;
    return agxc::kHdrBytes + 2 * (int64_t)world * (2 * sb);  // 2 parities x world slots of {word, tag} pairs (8 bytes per 32-bit word)
}

int agx_comm_alloc(int64_t bytes, void** ptr, unsigned char* handle) {
    if (bytes <= 0 || !ptr) return agx_internal_fail(AGX_ERR_ARG, "agx_comm_alloc: bad argument");
    void* p = nullptr;
    if (cudaMalloc(&p, (size_t)bytes) != cudaSuccess || cudaMemset(p, 0, (size_t)bytes) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
        cudaGetLastError();
        return agx_internal_fail(AGX_ERR_CUDA, "agx_comm_alloc: cudaMalloc/cudaMemset failed");
    }
    if (handle) {
        cudaIpcMemHandle_t h;
        static_assert(sizeof(h) == AGX_IPC_HANDLE_BYTES, "IPC handle size");
        if (cudaIpcGetMemHandle(&h, p) != cudaSuccess) {
            cudaGetLastError();
            cudaFree(p);
            return agx_internal_fail(AGX_ERR_CUDA, "agx_comm_alloc: cudaIpcGetMemHandle failed");
        }
        memcpy(handle, &h, sizeof(h));
    }
    *ptr = p;
    return AGX_OK;
}

int agx_comm_open(const unsigned char* handle, void** ptr) {
    if (!handle || !ptr) return agx_internal_fail(AGX_ERR_ARG, "agx_comm_open: bad argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        return agx_internal_fail(AGX_ERR_CUDA, "agx_comm_open: cudaIpcOpenMemHandle failed (peer access between the two GPUs?)");
    }
    *ptr = p;
    return AGX_OK;
}

int agx_comm_close(void* ptr) {
    if (!ptr) return AGX_OK;
    if (cudaIpcCloseMemHandle(ptr) != cudaSuccess) { cudaGetLastError(); return agx_internal_fail(AGX_ERR_CUDA, "agx_comm_close failed"); }
    return AGX_OK;
}

int agx_comm_free(void* ptr) {
    if (!ptr) return AGX_OK;
    if (cudaFree(ptr) != cudaSuccess) { cudaGetLastError(); return agx_internal_fail(AGX_ERR_CUDA, "agx_comm_free failed"); }
    return AGX_OK;
}

int agx_comm_allreduce(const AgxComm* c, void* buf, int64_t n, int dtype, void* stream) {
    if (!comm_ok(c) || !buf || n <= 0 || (dtype != AGX_F32 && dtype != AGX_F64)) return agx_internal_fail(AGX_ERR_ARG, "agx_comm_allreduce: bad argument");
    const int64_t bytes = n * (dtype == AGX_F64 ? 8 : 4);
    if (bytes > c->slot_bytes) return agx_internal_fail(AGX_ERR_ARG, "agx_comm_allreduce: message larger than the communicator's slot");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == AGX_F64) agx_comm_allreduce_kernel<double><<<agxc::kCluster, agxc::kBlock, 0, st>>>(*c, static_cast<double*>(buf), n);
    else agx_comm_allreduce_kernel<float><<<agxc::kCluster, agxc::kBlock, 0, st>>>(*c, static_cast<float*>(buf), n);
    return cudaGetLastError() == cudaSuccess ? AGX_OK : agx_internal_fail(AGX_ERR_CUDA, "agx_comm_allreduce: launch failed");
}

int agx_comm_status(const AgxComm* c, uint64_t* seq, uint64_t* err, void* stream) {
    if (!comm_ok(c)) return agx_internal_fail(AGX_ERR_ARG, "agx_comm_status: bad argument");
    uint64_t h[2] = {0, 0};
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (cudaMemcpyAsync(h, c->region[c->rank], sizeof(h), cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {
        cudaGetLastError();
        return agx_internal_fail(AGX_ERR_CUDA, "agx_comm_status: copy failed");
    }
    if (seq) *seq = h[0];
    if (err) *err = h[1];
    return AGX_OK;
}

}  // extern "C"

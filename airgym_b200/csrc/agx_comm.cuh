// agx_comm.cuh — device side of the small-message all-reduce over NVLink peer memory (include/agx.h "multi-GPU" section).
//
// Every rank owns one REGION in its own HBM; the regions of all ranks of the node are mapped into every process (CUDA IPC), so a
// kernel on rank r can store straight into rank p's memory through NVLink / NVSwitch.  One all-reduce is ONE kernel per rank:
//   push   : every rank writes its message into slot[parity][r] of EVERY peer's region (posted stores — no read round trip:
//            a peer load costs ~1 us, B300_MICROARCH.md "NVLink"), fences at system scope, then releases flag[r] = seq + 1
//            in every peer's region;
//   wait   : spins (acquire, system scope) on its OWN region's flags until every rank has signalled seq + 1;
//   reduce : adds the W slots of its own region in rank order — every rank adds the same numbers in the same order, so the
//            replicas stay bitwise identical without a broadcast.
// Slots are double buffered by the parity of the call counter `seq` (kept in the region, bumped by the kernel, so a captured
// CUDA graph replays without re-baked arguments): a rank can only be writing call s + 2 into a slot after it has seen every
// peer's flag of call s + 1, which that peer raises after it finished reading call s.  All ranks must issue the same sequence
// of collectives on one stream (SPMD), like any collective library.  A peer that never arrives trips a 20 s timeout which
// records the call in the region's error word instead of hanging the GPU.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "agx.h"

namespace agxc {

constexpr int kCluster = 8;      // CTAs of the one cluster that runs a collective
constexpr int kBlock = 1024;
constexpr int kHdrBytes = 256;   // seq | err | pad ... | flags[8] at byte 64
constexpr unsigned long long kTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned char* slot_ptr(const AgxComm& c, int owner, int parity, int src) {
    return static_cast<unsigned char*>(c.region[owner]) + kHdrBytes + (int64_t)(parity * c.world + src) * c.slot_bytes;
}

// Phase 1 + 2 of a collective, executed by all kCluster x kBlock threads of the cluster; returns the call's sequence number.
// `src` [n] elements of T in local memory.  On return every rank's message of this call sits in slot[seq & 1][*] of the own
// region and may be read with ld.volatile / __ldcv (the data arrived through NVLink, not through this SM's L1).
template <typename T>
__device__ __forceinline__ unsigned long long push_and_wait(const AgxComm& c, const T* __restrict__ src, int64_t n) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    unsigned long long* hdr = static_cast<unsigned long long*>(c.region[c.rank]);
    const unsigned long long seq = *reinterpret_cast<volatile unsigned long long*>(hdr);
    const int parity = (int)(seq & 1ull);
    const int64_t first = (int64_t)cluster.block_rank() * kBlock + threadIdx.x, stride = (int64_t)kCluster * kBlock;
    const bool vec = ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) && ((n * (int64_t)sizeof(T)) % 16 == 0);
    if (vec) {
        const int64_t n16 = n * (int64_t)sizeof(T) / 16;
        const uint4* s4 = reinterpret_cast<const uint4*>(src);
        for (int64_t i = first; i < n16; i += stride) {
            const uint4 v = s4[i];
            for (int p = 0; p < c.world; ++p) reinterpret_cast<uint4*>(slot_ptr(c, p, parity, c.rank))[i] = v;
        }
    } else {
        for (int64_t i = first; i < n; i += stride) {
            const T v = src[i];
            for (int p = 0; p < c.world; ++p) reinterpret_cast<T*>(slot_ptr(c, p, parity, c.rank))[i] = v;
        }
    }
    __threadfence_system();
    cluster.sync();  // every CTA's stores are fenced before the flags go up
    if (cluster.block_rank() == 0 && threadIdx.x < c.world) {
        unsigned long long* peer_flags = reinterpret_cast<unsigned long long*>(static_cast<unsigned char*>(c.region[threadIdx.x]) + 64);
        st_release_sys(peer_flags + c.rank, seq + 1ull);
    }
    if (threadIdx.x < c.world) {
        const unsigned long long* my_flags = reinterpret_cast<const unsigned long long*>(static_cast<unsigned char*>(c.region[c.rank]) + 64);
        const unsigned long long t0 = globaltimer_ns();
        while (ld_acquire_sys(my_flags + threadIdx.x) < seq + 1ull) {
            if (globaltimer_ns() - t0 > kTimeoutNs) {  // peer never arrived: record (call, peer) and give up instead of hanging
                hdr[1] = ((unsigned long long)(threadIdx.x + 1) << 56) | (seq + 1ull);
                break;
            }
            __nanosleep(64);
        }
    }
    __syncthreads();
    return seq;
}

// closing step: after every CTA of the cluster is done with `seq`, one thread bumps the region's call counter
__device__ __forceinline__ void finish(const AgxComm& c, unsigned long long seq) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    cluster.sync();
    if (cluster.block_rank() == 0 && threadIdx.x == 0) *static_cast<volatile unsigned long long*>(c.region[c.rank]) = seq + 1ull;
}

template <typename T>
__device__ __forceinline__ T reduce_elem(const AgxComm& c, int parity, int64_t i) {
    T s = __ldcv(reinterpret_cast<const T*>(slot_ptr(c, c.rank, parity, 0)) + i);
    for (int q = 1; q < c.world; ++q) s += __ldcv(reinterpret_cast<const T*>(slot_ptr(c, c.rank, parity, q)) + i);
    return s;
}

}  // namespace agxc

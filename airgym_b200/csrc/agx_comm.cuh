// agx_comm.cuh — device side of the small-message all-reduce over NVLink peer memory (include/agx.h "multi-GPU" section).
//
// Every rank owns one REGION in its own HBM; the regions of all ranks of the node are mapped into every process (CUDA IPC), so a
// kernel on rank r can store straight into rank p's memory through NVLink / NVSwitch.  One all-reduce is ONE kernel per rank and
// ONE traversal of the fabric ("LL" protocol: the flag travels with the data):
//   push   : every 32-bit word of the message goes into slot[parity][r] of EVERY peer's region as an 8-byte store {word, tag},
//            tag = call number + 1.  Posted stores — no read round trip (a peer load costs ~1 us, B300_MICROARCH.md "NVLink"), no
//            fence, no separate flag: an 8-byte store is delivered whole, so a word whose tag matches IS the data of this call;
//   reduce : each thread polls (volatile 8-byte loads of its OWN region) the W slots of the elements it owns until their tags
//            match, and adds them in rank order — every rank adds the same numbers in the same order, so the replicas stay bitwise
//            identical without a broadcast.
// Slots are double buffered by the parity of the call counter `seq` (kept in the region, bumped by the kernel, so a captured
// CUDA graph replays without re-baked arguments): a rank can only be writing call s + 2 into a slot after it has received every
// peer's words of call s + 1, which that peer pushed after it finished reading call s.  Tags only grow, so stale words (older
// calls, shorter messages) never match.  All ranks must issue the same sequence of collectives on one stream (SPMD), like any
// collective library.  A peer that never arrives trips a 20 s timeout which records the call in the region's error word instead
// of hanging the GPU.
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
// slots hold {word, tag} pairs: 8 bytes per 32-bit word of the message
__device__ __forceinline__ uint2* slot_ptr(const AgxComm& c, int owner, int parity, int src) {
    return reinterpret_cast<uint2*>(static_cast<unsigned char*>(c.region[owner]) + kHdrBytes + (int64_t)(parity * c.world + src) * (2 * c.slot_bytes));
}
__device__ __forceinline__ void st_pair(uint2* p, uint32_t word, uint32_t tag) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(word), "r"(tag) : "memory");
}
__device__ __forceinline__ uint2 ld_pair(const uint2* p) {
    uint2 v;
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}

// Phase 1 of a collective, executed by all kCluster x kBlock threads of the cluster: push `src` [n] elements of T (4 or 8 bytes) to
// every rank (this one included); returns the call's sequence number.
template <typename T>
__device__ __forceinline__ unsigned long long push(const AgxComm& c, const T* __restrict__ src, int64_t n) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned long long seq = *static_cast<volatile unsigned long long*>(c.region[c.rank]);
    const int parity = (int)(seq & 1ull);
    const uint32_t tag = (uint32_t)(seq + 1ull);
    const int64_t first = (int64_t)cluster.block_rank() * kBlock + threadIdx.x, stride = (int64_t)kCluster * kBlock;
    const int64_t n_words = n * (int64_t)(sizeof(T) / 4);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(src);
    for (int64_t i = first; i < n_words; i += stride) {
        const uint32_t v = w[i];
        for (int p = 0; p < c.world; ++p) st_pair(slot_ptr(c, p, parity, c.rank) + i, v, tag);
    }
    return seq;
}

// word `i` of rank q's message of call `seq`, as soon as it has arrived in this rank's region
__device__ __forceinline__ uint32_t poll_word(const AgxComm& c, unsigned long long seq, int q, int64_t i, unsigned long long& t0) {
    const uint2* p = slot_ptr(c, c.rank, (int)(seq & 1ull), q) + i;
    const uint32_t tag = (uint32_t)(seq + 1ull);
    uint2 v = ld_pair(p);
    while (v.y != tag) {
        if (t0 == 0ull) t0 = globaltimer_ns();
        else if (globaltimer_ns() - t0 > kTimeoutNs) {  // peer never arrived: record (call, peer) and give up instead of hanging
            static_cast<unsigned long long*>(c.region[c.rank])[1] = ((unsigned long long)(q + 1) << 56) | (seq + 1ull);
            break;
        }
        v = ld_pair(p);
    }
    return v.x;
}
template <typename T>
__device__ __forceinline__ T reduce_elem(const AgxComm& c, unsigned long long seq, int64_t i, unsigned long long& t0);
template <>
__device__ __forceinline__ float reduce_elem<float>(const AgxComm& c, unsigned long long seq, int64_t i, unsigned long long& t0) {
    float s = __uint_as_float(poll_word(c, seq, 0, i, t0));
    for (int q = 1; q < c.world; ++q) s += __uint_as_float(poll_word(c, seq, q, i, t0));
    return s;
}
template <>
__device__ __forceinline__ double reduce_elem<double>(const AgxComm& c, unsigned long long seq, int64_t i, unsigned long long& t0) {
    double s = 0.0;
    for (int q = 0; q < c.world; ++q) {
        const uint32_t lo = poll_word(c, seq, q, 2 * i, t0), hi = poll_word(c, seq, q, 2 * i + 1, t0);
        const double v = __hiloint2double((int)hi, (int)lo);
        s = q == 0 ? v : s + v;
    }
    return s;
}

// closing step: after every CTA of the cluster is done with `seq`, one thread bumps the region's call counter
__device__ __forceinline__ void finish(const AgxComm& c, unsigned long long seq) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    cluster.sync();
    if (cluster.block_rank() == 0 && threadIdx.x == 0) *static_cast<volatile unsigned long long*>(c.region[c.rank]) = seq + 1ull;
}

}  // namespace agxc

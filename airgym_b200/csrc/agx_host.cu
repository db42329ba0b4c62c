// agx_host.cu — pinned host staging buffers placed on the NUMA node of the GPU that reads / writes them.
// The end-to-end path (host actions in, host results out every step) is bound by the host<->device link; on a two-socket box a
// buffer that cudaHostAlloc happened to place on the other socket makes every copy cross the inter-socket link, and with 8 ranks
// all buffers landing on one node make that node's memory controllers the shared bottleneck (round-1 SCALE: 38 GB/s alone,
// 12 GB/s per rank at 8).  Here: mmap → mbind(MPOL_PREFERRED, node of the GPU's PCIe root, from sysfs) → touch → cudaHostRegister.
// If the node is unknown or the policy call is refused (cpuset), the buffer is still pinned, just not placed.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include "agx.h"

int agx_internal_fail(int code, const char* msg);

extern "C" {

int agx_device_numa_node(int device) {
    char bus[32] = "";
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) { cudaGetLastError(); return -1; }
    for (char* c = bus; *c; ++c)
        if (*c >= 'A' && *c <= 'Z') *c = (char)(*c - 'A' + 'a');  // sysfs names are lower case
    char path[128];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE* f = fopen(path, "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
}

int agx_host_alloc_pinned(int64_t bytes, int numa_node, void** ptr, int* placed) {
    if (bytes <= 0 || !ptr) return agx_internal_fail(AGX_ERR_ARG, "agx_host_alloc_pinned: bad argument");
    const size_t len = ((size_t)bytes + 4095) / 4096 * 4096;
    void* p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) return agx_internal_fail(AGX_ERR_ARG, "agx_host_alloc_pinned: mmap failed");
    int ok_node = 0;
    if (numa_node >= 0 && numa_node < 64) {
        unsigned long mask = 1ul << numa_node;
        const long r = syscall(SYS_mbind, p, len, 1 /* MPOL_PREFERRED */, &mask, 65ul, 0u);
        ok_node = (r == 0);
    }
    memset(p, 0, len);  // first touch under the policy
    if (cudaHostRegister(p, len, cudaHostRegisterPortable) != cudaSuccess) {
        cudaGetLastError();
        munmap(p, len);
        return agx_internal_fail(AGX_ERR_CUDA, "agx_host_alloc_pinned: cudaHostRegister failed");
    }
    if (placed) *placed = ok_node;
    *ptr = p;
    return AGX_OK;
}

int agx_host_free_pinned(void* ptr, int64_t bytes) {
    if (!ptr) return AGX_OK;
    cudaHostUnregister(ptr);
    cudaGetLastError();
    munmap(ptr, ((size_t)bytes + 4095) / 4096 * 4096);
    return AGX_OK;
}

}  // extern "C"

// agx_step_hovering.cu — instantiates the fused step kernel (agx_step_kernel.cuh) for the hovering task, all control modes.
#include "agx_step_kernel.cuh"

namespace agxk {
template int agx_dispatch_task<AGX_TASK_HOVERING>(const AgxParams&, int64_t, const AgxStepIO&, cudaStream_t);
template int agx_observe_task<AGX_TASK_HOVERING>(const AgxParams&, int64_t, const AgxStepIO&, cudaStream_t, int);
}  // namespace agxk

#ifdef AGX_TIMELINE
extern "C" int agx_debug_timeline(unsigned long long* host_out, int n_entries) {
    return (int)cudaMemcpyFromSymbol(host_out, agxk::g_timeline, sizeof(unsigned long long) * n_entries);
}
#endif

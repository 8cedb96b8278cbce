// launch.cuh — kernel launches with programmatic dependent launch (PDL).
//
// Every kernel of this library starts with pdl_entry(): wait until the previous kernel in the stream has
// completed and flushed (so reading its output is safe), then let the NEXT kernel in the stream be launched
// as soon as all of this kernel's CTAs have started.  The next grid's launch latency and prologue thus
// overlap this grid's last wave instead of following it — worth ~10 % on a 20 us kernel such as the 4096^2
// transpose, nothing on a millisecond one.  MDIM_PDL=0 turns the launch attribute off (the device side is
// then a no-op).
//
// Round 2: dependency-aware launches.  The context (api.cu) tracks the memory ranges of the kernels launched since
// the last one that really waited; a kernel whose operands and output touch none of them is launched with
// `nowait`: its threads skip the wait, so its first wave overlaps the previous grid's last one (4096^2 transpose
// back to back: 23.8 -> 20.5 us per launch).  ONE thread of such a grid still waits (the last CTA's thread 0), so
// "grid N complete => every earlier grid complete" keeps holding and a later kernel that does wait is safe.
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>

#include <utility>

namespace mdim {

inline bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("MDIM_PDL"); return !(e && e[0] == '0'); }();
    return on;
}

inline void pdl_config(cudaLaunchConfig_t& cfg, cudaLaunchAttribute& attr, dim3 grid, dim3 block, size_t smem, cudaStream_t stream) {
    cfg = cudaLaunchConfig_t{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = &attr; cfg.numAttrs = 1;
}

template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg; cudaLaunchAttribute attr;
    pdl_config(cfg, attr, grid, block, smem, stream);
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

#if defined(__CUDACC__)
__device__ __forceinline__ void pdl_entry(bool nowait = false) {
    if (!nowait || (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0)) asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
#endif

}  // namespace mdim

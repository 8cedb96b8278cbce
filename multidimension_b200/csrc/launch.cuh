// launch.cuh — kernel launches with programmatic dependent launch (PDL).
//
// Every kernel of this library starts with pdl_entry(): wait until the previous kernel in the stream has
// completed and flushed (so reading its output is safe), then let the NEXT kernel in the stream be launched
// as soon as all of this kernel's CTAs have started.  The next grid's launch latency and prologue thus
// overlap this grid's last wave instead of following it — worth ~10 % on a 20 us kernel such as the 4096^2
// transpose, nothing on a millisecond one.  MDIM_PDL=0 turns the launch attribute off (the device side is
// then a no-op).
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
__device__ __forceinline__ void pdl_entry() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
#endif

}  // namespace mdim

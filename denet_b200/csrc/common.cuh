// Shared host-side helpers for the C-ABI library: error reporting, launch checks, TMA descriptor encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/denet_b200.h"

namespace dn {

// Last error message, per host thread (C-ABI: functions return <0 and the text is read with denet_last_error()).
char* last_error_buf();
int set_error(int code, const char* fmt, ...);

#define DN_CHECK_CUDA(expr)                                                                              \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return dn::set_error(DENET_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,            \
                                 cudaGetErrorString(_e));                                                \
    } while (0)

#define DN_CHECK_LAUNCH() DN_CHECK_CUDA(cudaGetLastError())

#define DN_REQUIRE(cond, ...)                                                \
    do {                                                                     \
        if (!(cond)) return dn::set_error(DENET_ERR_ARG, __VA_ARGS__);       \
    } while (0)

// Every kernel launch site wraps its grid argument in DN_G(): counts the kernels this process has enqueued
// (denet_launch_count(); bench.py reports the per-step figure as gpu_launches).
void note_launch();
#define DN_G(grid) (dn::note_launch(), (grid))

// Programmatic dependent launch (griddepcontrol): a kernel launched through launch_pdl() may be scheduled while its
// predecessor in the stream is still draining; it runs its prologue (barrier init, TMEM allocation, descriptor
// prefetch, loads of data no recent kernel writes) and blocks in ptx::griddep_wait() before it touches anything the
// predecessor produces.  Every kernel launched this way MUST execute griddep_wait() on every path that reads or writes
// global memory shared with earlier kernels (the guarantee is transitive only through kernels that wait).
// denet_set_pdl(0) launches the same kernels fully serialised (the wait is then a no-op).
int pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

#ifdef __CUDACC__
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

int num_sms();

// Encode a bf16 tiled tensor map with 128B swizzle. dims/strides innermost first; strides in BYTES for dims 1..rank-1.
// estrides may be null (all 1).
int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, const uint32_t* estrides);

}  // namespace dn

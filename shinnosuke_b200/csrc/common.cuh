// Shared helpers for the shinnosuke_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>
#include "../../include/shinnosuke_b200.h"

namespace sn {

void set_error(const char *fmt, ...);
extern long long g_launch_count;
extern const char *g_last_kernel;

inline int check_cuda(cudaError_t e, const char *what) {
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return SN_OK;
}

// Called after every kernel launch: counts it and surfaces launch-configuration errors.
inline int after_launch(const char *name) {
    ++g_launch_count;
    g_last_kernel = name;
    return check_cuda(cudaPeekAtLastError(), name);
}

#define SN_REQUIRE(cond, msg)                                   \
    do {                                                        \
        if (!(cond)) {                                          \
            sn::set_error("%s: %s", __func__, msg);             \
            return SN_ERR_INVALID_ARG;                          \
        }                                                       \
    } while (0)

#define SN_CUDA(call)                                           \
    do {                                                        \
        int _e = sn::check_cuda((call), #call);                 \
        if (_e) return _e;                                      \
    } while (0)

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

inline int num_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
            sms = 148;
    }
    return sms;
}

// cudaFuncSetAttribute (dynamic shared memory above 48 KB) is per DEVICE: one process may drive several
// (fit(n_gpus=...), sn_comm_init_all), so "already configured" is remembered per device ordinal.
inline int current_device_slot() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess) d = 0;
    return d & 15;
}

// grid for a grid-stride elementwise kernel: enough CTAs for `work` items, capped at a
// multiple of the SM count (148 on B200) so the last wave is full.
inline int grid_for(long long work, int block, int ctas_per_sm = 8) {
    long long need = (work + block - 1) / block;
    long long cap = (long long)num_sms() * ctas_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

// ---- programmatic dependent launch (on by default; SN_PDL=0 launches every kernel normally) -----
// A kernel launched through launch_dependent() may start while the previous kernel of the stream is
// still draining: its blocks are scheduled as SMs free up and run their prologue (barrier init,
// TMEM allocation, tensor-map prefetch, index arithmetic) until pdl_wait(), which returns once the
// previous kernel has completed and its writes are visible.  EVERY kernel launched this way must call
// pdl_wait() before its first access to global memory; the call is a no-op under a normal launch.
// pdl_wait() also lets the NEXT kernel of the stream (if it was launched as a dependent) start being
// scheduled from here on: its blocks take whatever SM resources become free and park in their own
// pdl_wait() until this kernel has completed.  All blocks of this kernel are resident or done by the
// time the last of them has passed this point, so the dependents cannot starve it.
__device__ __forceinline__ void pdl_wait() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

inline bool pdl_enabled() {
    static int on = -1;
    if (on < 0) { const char *e = getenv("SN_PDL"); on = (e && e[0] == '0') ? 0 : 1; }
    return on == 1;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_dependent(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float relu_keep_sign(float z) {
    // max(0, z) with the sign of a negative pre-activation kept as -0.0f: numerically
    // zero for every consumer, but "pre-activation < 0" stays recoverable for backward.
    // (z == -0.0f is NOT < 0, so it yields +0.0f and the gradient passes, like the reference.)
    return __uint_as_float(__float_as_uint(fmaxf(z, 0.f)) | (z < 0.f ? 0x80000000u : 0u));
}
__device__ __forceinline__ bool relu_blocked(float y) { return __float_as_uint(y) == 0x80000000u; }

__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }

}  // namespace sn

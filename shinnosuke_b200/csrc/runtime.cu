// Runtime plumbing of the C ABI: errors, memory, streams, events.
#include "common.cuh"
#include <string.h>

namespace sn {
static thread_local char g_err[512] = "";
long long g_launch_count = 0;
const char *g_last_kernel = "";
extern long long g_igemm_pair_launches;      // conv_tc.cu
extern int g_igemm_pair_mode;
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace sn
using namespace sn;

extern "C" {

int sn_abi_version(void) { return SN_ABI_VERSION; }
const char *sn_last_error(void) { return g_err; }
long long sn_launch_count(void) { return g_launch_count; }
const char *sn_last_kernel(void) { return g_last_kernel; }
void sn_launch_count_reset(void) { g_launch_count = 0; }
long long sn_igemm_pair_launches(void) { return g_igemm_pair_launches; }
int sn_igemm_set_pair_mode(int mode) {
    SN_REQUIRE(mode >= -1 && mode <= 2, "pair mode must be -1 (environment), 0 (never), 1 (by shape) or 2 (whenever eligible)");
    g_igemm_pair_mode = mode;
    return SN_OK;
}

int sn_device_count(int *count) {
    SN_REQUIRE(count, "null out pointer");
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) { *count = 0; return check_cuda(e, "cudaGetDeviceCount"); }
    return SN_OK;
}
int sn_set_device(int device) { SN_CUDA(cudaSetDevice(device)); return SN_OK; }
int sn_device_props(int device, int *sm_count, int *cc_major, int *cc_minor, size_t *global_mem) {
    cudaDeviceProp p;
    SN_CUDA(cudaGetDeviceProperties(&p, device));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (global_mem) *global_mem = p.totalGlobalMem;
    return SN_OK;
}
int sn_malloc(void **ptr, size_t bytes) { SN_REQUIRE(ptr, "null out pointer"); SN_CUDA(cudaMalloc(ptr, bytes)); return SN_OK; }
int sn_free(void *ptr) { SN_CUDA(cudaFree(ptr)); return SN_OK; }
int sn_malloc_host(void **ptr, size_t bytes) { SN_REQUIRE(ptr, "null out pointer"); SN_CUDA(cudaMallocHost(ptr, bytes)); return SN_OK; }
int sn_free_host(void *ptr) { SN_CUDA(cudaFreeHost(ptr)); return SN_OK; }
int sn_host_register(void *ptr, size_t bytes) {
    SN_REQUIRE(ptr && bytes, "bad arguments");
    cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterDefault);
    if (e != cudaSuccess) {
        cudaGetLastError();            // a refused registration must not poison later launch checks
        return check_cuda(e, "cudaHostRegister");
    }
    return SN_OK;
}
int sn_host_unregister(void *ptr) {
    cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) { cudaGetLastError(); return check_cuda(e, "cudaHostUnregister"); }
    return SN_OK;
}
int sn_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream) {
    SN_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, as_stream(stream))); return SN_OK;
}
int sn_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream) {
    SN_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, as_stream(stream))); return SN_OK;
}
int sn_memcpy_d2d(void *dst, const void *src, size_t bytes, void *stream) {
    SN_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, as_stream(stream))); return SN_OK;
}
int sn_memset(void *dst, int byte, size_t bytes, void *stream) {
    SN_CUDA(cudaMemsetAsync(dst, byte, bytes, as_stream(stream))); return SN_OK;
}
int sn_stream_create(void **stream) {
    SN_REQUIRE(stream, "null out pointer");
    cudaStream_t s;
    SN_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = s;
    return SN_OK;
}
int sn_stream_destroy(void *stream) { SN_CUDA(cudaStreamDestroy(as_stream(stream))); return SN_OK; }
int sn_stream_sync(void *stream) { SN_CUDA(cudaStreamSynchronize(as_stream(stream))); return SN_OK; }
int sn_event_create(void **event) {
    SN_REQUIRE(event, "null out pointer");
    cudaEvent_t e;
    SN_CUDA(cudaEventCreate(&e));
    *event = e;
    return SN_OK;
}
int sn_event_destroy(void *event) { SN_CUDA(cudaEventDestroy((cudaEvent_t)event)); return SN_OK; }
int sn_event_record(void *event, void *stream) { SN_CUDA(cudaEventRecord((cudaEvent_t)event, as_stream(stream))); return SN_OK; }
int sn_event_elapsed_ms(void *start, void *stop, float *ms) {
    SN_REQUIRE(ms, "null out pointer");
    SN_CUDA(cudaEventSynchronize((cudaEvent_t)stop));
    SN_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
    return SN_OK;
}

}  // extern "C"

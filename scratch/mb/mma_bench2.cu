#include <cstdio>
#include <cuda_runtime.h>
#include "../../shinnosuke_b200/csrc/tc_common.cuh"
using namespace sn::tc;

// issuers: number of warps that each issue `iters` MMAs (to their own accumulator columns)
template <int KIND>
__global__ void __launch_bounds__(128, 1) bench(int N, int iters, int issuers, int M, long long *out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar[4];
    __shared__ uint32_t tmem_ptr;
    for (int i = threadIdx.x; i < 48 * 1024; i += 128) reinterpret_cast<float *>(smem)[i] = 0.f;
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); }
    if (threadIdx.x < 32) { tmem_alloc(&tmem_ptr, 512); tmem_relinquish(); }
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0 && w < issuers) {
        uint32_t idesc = make_idesc(KIND, M, N, 0, 0);
        uint32_t sa = smem_u32(smem), sb = sa + 64 * 1024;
        uint64_t ad = make_smem_desc(sa, 0, 1024), bd = make_smem_desc(sb, 0, 1024);
        uint32_t d = tmem_ptr + w * 128;
        long long t0 = clock64();
        for (int i = 0; i < iters; i += 8) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (KIND == 2) mma_tf32(d, ad, bd, idesc, 1); else mma_bf16(d, ad, bd, idesc, 1);
            }
        }
        mma_commit(&bar[w]);
        mbar_wait(&bar[w], 0);
        long long t1 = clock64();
        out[blockIdx.x * 4 + w] = t1 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem_ptr, 512); }
}

template <int KIND>
void run(const char *name, int M, int N, int issuers) {
    long long *d; cudaMalloc(&d, 8 * 4);
    int smem = 200 * 1024;
    cudaFuncSetAttribute(bench<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int iters = 4000;
    bench<KIND><<<1, 128, smem>>>(N, iters, issuers, M, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[4];
    cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
    int K = KIND == 2 ? 8 : 16;
    double cyc = (double)h[0] / iters;
    printf("%-10s M=%3d N=%3d issuers=%d: %7.1f cyc/mma/issuer  total %7.1f MAC/clk  (%s)\n", name, M, N, issuers, cyc, issuers * (double)M * N * K / cyc, cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    for (int N : {16, 32, 64, 128, 256}) run<2>("tf32", 128, N, 1);
    for (int N : {32, 128}) run<2>("tf32", 128, N, 2);
    for (int N : {32, 128}) run<2>("tf32", 128, N, 4);
    for (int N : {32, 64, 128, 256}) run<2>("tf32", 64, N, 1);
    for (int N : {32, 256}) run<1>("bf16", 128, N, 1);
    return 0;
}

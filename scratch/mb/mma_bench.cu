// Microbenchmark: cycles per tcgen05.mma for tf32 / bf16, K-major vs MN-major operands, several N.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../shinnosuke_b200/csrc/tc_common.cuh"
using namespace sn::tc;

template <int KIND /*2 tf32, 1 bf16*/, int AMN, int BMN>
__global__ void __launch_bounds__(128, 1) bench(int N, int iters, int nacc, long long *out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    for (int i = threadIdx.x; i < 48 * 1024; i += 128) reinterpret_cast<float *>(smem)[i] = 0.f;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) { tmem_alloc(&tmem_ptr, 512); tmem_relinquish(); }
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (threadIdx.x == 0) {
        uint32_t idesc = make_idesc(KIND, 128, N, AMN, BMN);
        uint32_t sa = smem_u32(smem), sb = sa + 64 * 1024;
        uint32_t la = AMN ? (KIND == 2 ? LAYOUT_SW128_BASE32B : LAYOUT_SW128) : LAYOUT_SW128;
        uint32_t lb = BMN ? (KIND == 2 ? LAYOUT_SW128_BASE32B : LAYOUT_SW128) : LAYOUT_SW128;
        uint64_t ad = AMN ? make_smem_desc(sa, 4096, KIND == 2 ? 512 : 1024, la) : make_smem_desc(sa, 0, 1024, la);
        uint64_t bd = BMN ? make_smem_desc(sb, 4096, KIND == 2 ? 512 : 1024, lb) : make_smem_desc(sb, 0, 1024, lb);
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            uint32_t d = tmem_ptr + (i % nacc) * (N <= 128 ? 128 : 256);
            if (KIND == 2) mma_tf32(d, ad, bd, idesc, 1); else mma_bf16(d, ad, bd, idesc, 1);
        }
        mma_commit(&bar);
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem_ptr, 512); }
}

template <int KIND, int AMN, int BMN>
void run(const char *name, int N, int nacc, int grid) {
    long long *d; cudaMalloc(&d, 8 * 256);
    int smem = 200 * 1024;
    cudaFuncSetAttribute(bench<KIND, AMN, BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int iters = 2000;
    bench<KIND, AMN, BMN><<<grid, 128, smem>>>(N, iters, nacc, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[256];
    cudaMemcpy(h, d, 8 * grid, cudaMemcpyDeviceToHost);
    int K = KIND == 2 ? 8 : 16;
    double cyc = (double)h[0] / iters;
    printf("%-28s N=%3d acc=%d grid=%3d: %7.1f cyc/mma  %7.1f MAC/clk  (%s)\n", name, N, nacc, grid, cyc, 128.0 * N * K / cyc, cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    int Ns[] = {32, 64, 128, 256};
    for (int g : {1, 148}) {
        for (int N : Ns) run<2, 0, 0>("tf32 A:K  B:K", N, 1, g);
        for (int N : Ns) run<2, 1, 1>("tf32 A:MN B:MN", N, 1, g);
        for (int N : Ns) run<2, 0, 1>("tf32 A:K  B:MN", N, 1, g);
        for (int N : Ns) run<2, 1, 0>("tf32 A:MN B:K", N, 1, g);
        for (int N : Ns) run<1, 0, 0>("bf16 A:K  B:K", N, 1, g);
        for (int N : Ns) run<1, 1, 1>("bf16 A:MN B:MN", N, 1, g);
    }
    run<2, 1, 1>("tf32 MN/MN 2 accumulators", 32, 2, 1);
    run<2, 1, 1>("tf32 MN/MN 4 accumulators", 32, 4, 1);
    run<2, 0, 0>("tf32 K/K 2 accumulators", 128, 2, 1);
    return 0;
}

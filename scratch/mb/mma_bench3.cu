#include <cstdio>
#include <cuda_runtime.h>
#include "../../shinnosuke_b200/csrc/tc_common.cuh"
using namespace sn::tc;
// tf32 (KIND 2) / bf16 (KIND 1), both operands MN-major, B start shifted by `shift_rows` rows of 128 B
template <int KIND>
__global__ void __launch_bounds__(128, 1) bench(int N, int iters, int shift_rows, int a_mn, int b_mn, long long *out) {
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
        uint32_t idesc = make_idesc(KIND, 128, N, a_mn, b_mn);
        uint32_t sa = smem_u32(smem), sb = sa + 64 * 1024 + shift_rows * 128;
        uint32_t lay = KIND == 2 ? LAYOUT_SW128_BASE32B : LAYOUT_SW128;
        uint32_t kg = KIND == 2 ? 512 : 1024;
        uint64_t ad = a_mn ? make_smem_desc(sa, 8192, kg, lay) : make_smem_desc(sa, 0, 1024, LAYOUT_SW128);
        uint64_t bd = b_mn ? make_smem_desc(sb, 16384, kg, lay) : make_smem_desc(sb, 0, 1024, LAYOUT_SW128);
        uint32_t d = tmem_ptr;
        long long t0 = clock64();
        for (int i = 0; i < iters; i += 8) {
#pragma unroll
            for (int u = 0; u < 8; ++u) { if (KIND == 2) mma_tf32(d + (u % 3) * 64, ad + (u & 7) * 64, bd + (u & 7) * 72, idesc, 1); else mma_bf16(d + (u % 3) * 64, ad + (u & 3) * 128, bd + (u & 3) * 136, idesc, 1); }
        }
        mma_commit(&bar);
        mbar_wait(&bar, 0);
        out[0] = clock64() - t0;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem_ptr, 512); }
}
template <int KIND>
void run(const char *name, int N, int shift, int a_mn, int b_mn) {
    long long *d; cudaMalloc(&d, 8);
    int smem = 200 * 1024;
    cudaFuncSetAttribute(bench<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int iters = 4000;
    bench<KIND><<<1, 128, smem>>>(N, iters, shift, a_mn, b_mn, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%-6s N=%3d A:%s B:%s shift=%d rows: %7.1f cyc/mma (%s)\n", name, N, a_mn ? "MN" : "K ", b_mn ? "MN" : "K ", shift, (double)h / iters, cudaGetErrorString(e));
    cudaFree(d);
}
int main() {
    for (int N : {64, 128}) {
        run<2>("tf32", N, 0, 0, 0); run<2>("tf32", N, 0, 1, 0); run<2>("tf32", N, 0, 0, 1); run<2>("tf32", N, 0, 1, 1);
        run<2>("tf32", N, 1, 1, 1); run<2>("tf32", N, 3, 1, 1); run<2>("tf32", N, 1, 0, 0);
        run<1>("bf16", N, 0, 0, 0); run<1>("bf16", N, 0, 1, 0); run<1>("bf16", N, 0, 0, 1); run<1>("bf16", N, 0, 1, 1); run<1>("bf16", N, 1, 1, 1); run<1>("bf16", N, 5, 1, 1);
    }
    return 0;
}

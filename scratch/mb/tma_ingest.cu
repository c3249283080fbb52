// How fast can one SM take operand tiles from L2 through TMA when ALL SMs do it, and does a
// 2-CTA cluster that multicasts the shared (B) tile raise the rate?  
//   mode 0: every CTA loads A (16 KB) + B (16 KB) per stage by itself            -> 32 KB issued / stage
//   mode 1: CTA pairs; each CTA loads its own A (16 KB) and HALF of B (8 KB), multicast to both CTAs
//                                                                                  -> 24 KB issued / stage
// Both modes deliver 32 KB per stage into each CTA's shared memory.  Source tensors are small
// (L2-resident).  Prints cycles per stage and bytes delivered per clock per SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scratch/mb/tma_ingest scratch/mb/tma_ingest.cu
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mbar_arrive_remote(uint64_t *b, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(s32(b)), "r"(rank));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(r) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t phase) {
    asm volatile(
        "{ .reg .pred p; W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1; @p bra D; bra W; D: }" ::"r"(s32(b)), "r"(phase)
        : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *b, uint32_t phase) {
    asm volatile(
        "{ .reg .pred p; W: mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1; @p bra D; bra W; D: }" ::"r"(s32(b)),
        "r"(phase)
        : "memory");
}
__device__ __forceinline__ void tma2d(void *dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(s32(dst)),
                 "l"(reinterpret_cast<uint64_t>(m)), "r"(s32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma2d_mc(void *dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
            s32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(s32(bar)), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

#ifndef NSTAGES
#define NSTAGES 4
#endif
constexpr int STAGES = NSTAGES, TILE = 16 * 1024;

// tmA: rows x 32 floats, box 32 x 128 (16 KB);  tmB: box 32 x 128 (mode 0) or 32 x 64 (mode 1)
template <int MODE>
__global__ void __launch_bounds__(64, 1) ingest(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int rows,
                                                int iters, long long *out) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t full[STAGES], empty[STAGES];
    uint32_t rank = 0;
    if (MODE == 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], MODE == 1 ? 2 : 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int i = 0; i < STAGES; ++i) mbar_expect(&full[i], 2 * TILE);        // armed before any peer can write here
    }
    __syncthreads();
    if (MODE == 1) cluster_sync();
    const int blocks = rows / 128;
    if (threadIdx.x == 0) {
        // producer; a slot is re-armed by the consumer thread before it frees the slot
        long long t0 = clock64();
        for (int s = 0; s < iters; ++s) {
            const int slot = s % STAGES;
            if (s >= STAGES) {
                if (MODE == 1) mbar_wait_cluster(&empty[slot], ((s / STAGES) - 1) & 1); else mbar_wait(&empty[slot], ((s / STAGES) - 1) & 1);
            }
            uint8_t *a = smem + slot * 2 * TILE, *b = a + TILE;
            const int r0 = ((blockIdx.x * 7 + s) % blocks) * 128;
            tma2d(a, &tmA, &full[slot], 0, r0);
            if (MODE == 0) tma2d(b, &tmB, &full[slot], 0, (s % blocks) * 128);
            else tma2d_mc(b + rank * (TILE / 2), &tmB, &full[slot], 0, (s % blocks) * 128 + rank * 64, (uint16_t)3);
        }
        out[blockIdx.x * 2] = clock64() - t0;
    } else if (threadIdx.x == 32) {
        // consumer: wait for the data, re-arm the slot, free it (on both CTAs of the pair in mode 1)
        long long t0 = clock64();
        for (int s = 0; s < iters; ++s) {
            const int slot = s % STAGES;
            mbar_wait(&full[slot], (s / STAGES) & 1);
            if (s + STAGES < iters) mbar_expect(&full[slot], 2 * TILE);
            if (MODE == 1) { mbar_arrive_remote(&empty[slot], 0); mbar_arrive_remote(&empty[slot], 1); }
            else mbar_arrive(&empty[slot]);
        }
        out[blockIdx.x * 2 + 1] = clock64() - t0;
    }
    __syncthreads();
    if (MODE == 1) cluster_sync();
}

static EncodeFn get_encode() {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    return (EncodeFn)p;
}
static CUtensorMap make(EncodeFn enc, float *base, int rows, int box_rows) {
    CUtensorMap m;
    cuuint64_t gdim[2] = {32, (cuuint64_t)rows}, gstr[1] = {128};
    cuuint32_t box[2] = {32, (cuuint32_t)box_rows}, es[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) printf("encode failed %d\n", (int)r);
    return m;
}

int main() {
    EncodeFn enc = get_encode();
    const int rows = 16384;                        // 2 MB per tensor: L2 resident
    float *a, *b;
    cudaMalloc(&a, rows * 128); cudaMalloc(&b, rows * 128);
    cudaMemset(a, 0, rows * 128); cudaMemset(b, 0, rows * 128);
    long long *out; cudaMalloc(&out, 148 * 2 * 8);
    const int smem = STAGES * 2 * TILE + 2048, iters = 2000, grid = 148;
    CUtensorMap tmA = make(enc, a, rows, 128), tmB0 = make(enc, b, rows, 128), tmB1 = make(enc, b, rows, 64);
    long long h[296];
    for (int rep = 0; rep < 2; ++rep) {
        cudaFuncSetAttribute(ingest<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        ingest<0><<<grid, 64, smem>>>(tmA, tmB0, rows, iters, out);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        double c = 0; for (int i = 0; i < grid; ++i) c += (double)h[2 * i + 1];
        c /= grid;
        printf("mode 0 (no cluster)        : %7.1f cycles / 32 KB stage = %5.1f B/clk delivered per SM, %5.1f issued (%s)\n", c / iters,
               32768.0 * iters / c, 32768.0 * iters / c, cudaGetErrorString(e));
        cudaFuncSetAttribute(ingest<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, ingest<1>, tmA, tmB1, rows, iters, out);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        c = 0; for (int i = 0; i < grid; ++i) c += (double)h[2 * i + 1];
        c /= grid;
        printf("mode 1 (pairs, B multicast): %7.1f cycles / 32 KB stage = %5.1f B/clk delivered per SM, %5.1f issued (%s)\n", c / iters,
               32768.0 * iters / c, 24576.0 * iters / c, cudaGetErrorString(e));
    }
    return 0;
}

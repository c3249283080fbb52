// Does Nsight Compute profile kernels launched as programmatic dependents (also out of a CUDA graph)?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scratch/mb/pdl_ncu scratch/mb/pdl_ncu.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float *p, int n, float a) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = p[i] * a + 1.f;
}
static cudaError_t launch(cudaStream_t st, float *p, int n, float a) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148); cfg.blockDim = dim3(256); cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k, p, n, a);
}
int main() {
    const int n = 1 << 20;
    float *p; cudaMalloc(&p, n * 4); cudaMemset(p, 0, n * 4);
    cudaStream_t st; cudaStreamCreate(&st);
    for (int i = 0; i < 3; ++i) launch(st, p, n, 0.5f);
    cudaGraph_t g; cudaGraphExec_t ge;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    for (int i = 0; i < 3; ++i) launch(st, p, n, 0.5f);
    cudaStreamEndCapture(st, &g);
    cudaGraphInstantiate(&ge, g, 0);
    cudaGraphLaunch(ge, st);
    cudaError_t e = cudaStreamSynchronize(st);
    float h; cudaMemcpy(&h, p, 4, cudaMemcpyDeviceToHost);
    printf("done: %s, p[0] = %f (expect 1.968750)\n", cudaGetErrorString(e), h);
    return e != cudaSuccess;
}

// Micro-benchmark: what a kernel boundary costs inside a CUDA graph on B200, and how much of it programmatic dependent launch
// (griddepcontrol.launch_dependents / griddepcontrol.wait + cudaLaunchAttributeProgrammaticStreamSerialization) takes back.
// A chain of 256 dependent kernels (each reads the previous one's output), captured into a graph and replayed; kernels are either
// light (148 CTAs x 256 threads, 1 MB) or carry 200 KB of dynamic shared memory like the persistent convolutions (one CTA per SM).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o profiles/_bin/pdl_microbench profiles/pdl_microbench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <bool PDL>
__global__ void __launch_bounds__(256) step_kernel(const float4* __restrict__ in, float4* __restrict__ out, int n4, int spin) {
    extern __shared__ float sm[];
    if (PDL) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // prologue work that does not depend on the previous kernel (stands for barrier init / TMEM alloc / tensor-map prefetch)
    float acc = 0.f;
    for (int i = 0; i < spin; ++i) acc += __sinf((float)(threadIdx.x + i));
    if (threadIdx.x == 0) sm[0] = acc;
    if (PDL) asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
        float4 v = in[i];
        v.x += 1.f; v.y += 1.f; v.z += 1.f; v.w += 1.f;
        out[i] = v;
    }
}

template <bool PDL>
static float run_chain(cudaStream_t st, float4* a, float4* b, int n4, int smem, int nk, int spin) {
    CK(cudaFuncSetAttribute(step_kernel<PDL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaGraph_t graph; cudaGraphExec_t exec;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    for (int k = 0; k < nk; ++k) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(148); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = PDL ? 1 : 0;
        const float4* in = (k & 1) ? b : a; float4* out = (k & 1) ? a : b;
        CK(cudaLaunchKernelEx(&cfg, step_kernel<PDL>, in, out, n4, spin));
    }
    CK(cudaStreamEndCapture(st, &graph));
    CK(cudaGraphInstantiate(&exec, graph, 0));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) CK(cudaGraphLaunch(exec, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaEventRecord(e0, st));
    for (int i = 0; i < 10; ++i) CK(cudaGraphLaunch(exec, st));
    CK(cudaEventRecord(e1, st));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    CK(cudaGraphExecDestroy(exec)); CK(cudaGraphDestroy(graph));
    return ms / 10 / nk * 1e3f;          // us per kernel
}

int main() {
    cudaStream_t st; CK(cudaStreamCreate(&st));
    const int nk = 256;
    for (int mb : {1, 16, 64}) {
        const int n4 = mb * 1024 * 1024 / 16;
        float4 *a, *b; CK(cudaMalloc(&a, (size_t)n4 * 16)); CK(cudaMalloc(&b, (size_t)n4 * 16));
        CK(cudaMemset(a, 0, (size_t)n4 * 16));
        for (int smem : {1024, 200 * 1024}) {
            for (int spin : {0, 64}) {
                float plain = run_chain<false>(st, a, b, n4, smem, nk, spin);
                float pdl = run_chain<true>(st, a, b, n4, smem, nk, spin);
                printf("tensor %3d MB  smem %6d B  prologue spin %3d: plain %7.2f us/kernel   PDL %7.2f us/kernel   saved %6.2f us\n", mb, smem, spin,
                       plain, pdl, plain - pdl);
            }
        }
        // correctness of the chain under PDL: every element was incremented once per kernel
        CK(cudaMemset(a, 0, (size_t)n4 * 16));
        CK(cudaFree(a)); CK(cudaFree(b));
    }
    return 0;
}

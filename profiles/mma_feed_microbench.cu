// Micro-benchmark: how many SM cycles does ONE tcgen05.mma take as a function of where its operands come from?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/_bin/mma_feed profiles/mma_feed_microbench.cu -lcuda
// Every CTA (one per SM, or one CTA pair per 2 SMs) has one thread issue `iters` back-to-back MMAs into one TMEM accumulator and
// times issue -> tcgen05.commit arrival with clock64 (SM cycles: independent of the clock the power cap settles at).
// Operand contents are zeros - the data path, not the values, is measured.
//   mode ss      A, B from shared memory (K-major, 128B swizzle)          M = 128
//   mode ts      A from TMEM, B from shared memory                        M = 128
//   mode ss2     cta_group::2: A [128 rows] and half of B per CTA         M = 256 per pair
//   mode tf32    kind::tf32, A and B from shared memory (K = 8 per MMA)   M = 128
//   +tma         a second warp streams bulk copies (global -> shared) into a scratch ring while the MMAs run
// Output: one line per (mode, N): median / min / max cycles per MMA over the CTAs, and the implied shared-memory operand rate.
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) { if (++spins > (1u << 26)) { printf("mbar timeout block %d\n", blockIdx.x); __trap(); } }
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
// fmt: 1 = bf16, 2 = tf32
__device__ __forceinline__ uint32_t idesc(int m, int n, int fmt) {
    uint32_t d = 0;
    d |= 1u << 4; d |= (uint32_t)fmt << 7; d |= (uint32_t)fmt << 10;
    d |= (uint32_t)(n >> 3) << 17;
    d |= (uint32_t)(m >> 4) << 24;
    return d;
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n.reg .b32 %%rx;\n.reg .pred %%px;\nelect.sync %%rx|%%px, %1;\n@%%px mov.s32 %0, 1;\n}\n" : "+r"(pred) : "r"(0xffffffffu));
    return pred != 0;
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

enum { MODE_SS = 0, MODE_TS = 1, MODE_SS2 = 2, MODE_TF32 = 3 };

struct Args { int n, iters, na, tma; const uint8_t* gsrc; unsigned long long* cycles; };

constexpr int A_SLAB = 128 * 128;     // 128 rows x 128 bytes
constexpr int B_SLAB = 256 * 128;
constexpr int TMA_RING = 4, TMA_CHUNK = 16384;

template <int MODE>
__global__ void __launch_bounds__(96, 1) feed_kernel(Args a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                                   // na slabs
    uint8_t* sB = smem + a.na * A_SLAB;
    uint8_t* sT = sB + B_SLAB;                            // TMA scratch ring
    uint64_t* bars = reinterpret_cast<uint64_t*>(sT + TMA_RING * TMA_CHUNK);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr bool pair = MODE == MODE_SS2;
    const uint32_t rank = pair ? cluster_ctarank() : 0;
    for (int i = threadIdx.x * 16; i < a.na * A_SLAB + B_SLAB; i += 96 * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
    const uint32_t done = smem_u32(bars), tbar0 = smem_u32(bars + 1);
    __shared__ volatile int stop_flag;
    if (threadIdx.x == 0) {
        mbar_init(done, 1);
        for (int i = 0; i < TMA_RING; ++i) mbar_init(tbar0 + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        stop_flag = 0;
    }
    if (warp == 0) {
        if (pair) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    fence_before();
    __syncthreads();
    if (pair) cluster_sync_all();
    fence_after();
    const uint32_t tmem = *tmem_ptr;

    if (warp == 0 && rank == 0) {
      // warp-uniform branch + elect.sync: the tcgen05 instructions live on the uniform datapath
      if (elect_one()) {
        constexpr int fmt = MODE == MODE_TF32 ? 2 : 1;
        const uint32_t id = idesc(pair ? 256 : 128, a.n, fmt);
        const uint64_t ad0 = smem_desc(smem_u32(sA), 16, 1024, 2), bd0 = smem_desc(smem_u32(sB), 16, 1024, 2);
        const uint32_t tmem_a = tmem + 256;               // TS mode: A operand columns (content irrelevant)
        long long t0 = clock64();
        for (int i = 0; i < a.iters; ++i) {
            const uint64_t koff = (uint64_t)((i & 3) * 2);
            const uint64_t ad = ad0 + (uint64_t)(((i % a.na) * A_SLAB) >> 4) + koff, bd = bd0 + koff;
            if constexpr (MODE == MODE_SS)
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                             ::"r"(tmem), "l"(ad), "l"(bd), "r"(id), "r"(1) : "memory");
            else if constexpr (MODE == MODE_TF32)
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                             ::"r"(tmem), "l"(ad), "l"(bd), "r"(id), "r"(1) : "memory");
            else if constexpr (MODE == MODE_TS)
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
                             ::"r"(tmem), "r"(tmem_a + (uint32_t)((i & 3) * 8)), "l"(bd), "r"(id), "r"(1) : "memory");
            else
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                             ::"r"(tmem), "l"(ad), "l"(bd), "r"(id), "r"(1) : "memory");
        }
        if (pair)
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(done), "h"((uint16_t)1) : "memory");
        else
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(done) : "memory");
        mbar_wait(done, 0);
        long long t1 = clock64();
        a.cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
        stop_flag = 1;
      }
      __syncwarp();
    } else if (warp == 1 && lane == 0 && a.tma && rank == 0) {
        // stream bulk copies into the scratch ring until the MMA thread is done
        int slot = 0; uint32_t phase = 0; long long n = 0;
        while (!stop_flag) {
            const uint32_t bar = tbar0 + 8 * slot;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(TMA_CHUNK) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(sT + slot * TMA_CHUNK)), "l"(a.gsrc + ((n * 148 + blockIdx.x) % 4096) * TMA_CHUNK), "r"(TMA_CHUNK), "r"(bar) : "memory");
            if (slot == TMA_RING - 1) {                   // wait for the whole ring, then reuse it
                for (int s = 0; s < TMA_RING; ++s) mbar_wait(tbar0 + 8 * s, phase);
                phase ^= 1;
            }
            slot = (slot + 1) % TMA_RING; ++n;
        }
        // drain what is in flight
        for (int s = 0; s < slot; ++s) mbar_wait(tbar0 + 8 * s, phase);
        a.cycles[gridDim.x + blockIdx.x] = (unsigned long long)n;
    }
    fence_before();
    __syncthreads();
    if (pair) cluster_sync_all();
    if (warp == 0) {
        fence_after();
        if (pair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

static void run(const char* name, int mode, int n, int grid, int na, int tma, const uint8_t* gsrc, unsigned long long* dcyc) {
    Args a; a.n = n; a.iters = 4096; a.na = na; a.tma = tma; a.gsrc = gsrc; a.cycles = dcyc;
    const int smem = 1024 + na * A_SLAB + B_SLAB + TMA_RING * TMA_CHUNK + 256;
    void (*kern)(Args) = mode == MODE_SS ? feed_kernel<MODE_SS> : mode == MODE_TS ? feed_kernel<MODE_TS> : mode == MODE_SS2 ? feed_kernel<MODE_SS2> : feed_kernel<MODE_TF32>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaMemset(dcyc, 0, sizeof(unsigned long long) * 2 * 148);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(96); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = mode == MODE_SS2 ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    for (int rep = 0; rep < 2; ++rep) {                   // first launch warms up
        cudaError_t e = cudaSuccess;
        if (mode == MODE_SS2) e = cudaLaunchKernelEx(&cfg, kern, a);
        else { kern<<<grid, 96, smem>>>(a); e = cudaGetLastError(); }
        if (e != cudaSuccess) { printf("%s N=%d launch failed: %s\n", name, n, cudaGetErrorString(e)); return; }
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s N=%d failed: %s\n", name, n, cudaGetErrorString(e)); exit(1); }
    }
    std::vector<unsigned long long> h(2 * 148);
    cudaMemcpy(h.data(), dcyc, sizeof(unsigned long long) * 2 * 148, cudaMemcpyDeviceToHost);
    std::vector<double> c;
    double chunks = 0;
    for (int i = 0; i < grid; ++i) { if (h[i]) c.push_back((double)h[i] / a.iters); chunks += (double)h[grid + i]; }
    std::sort(c.begin(), c.end());
    const double med = c[c.size() / 2];
    const int esz = mode == MODE_TF32 ? 4 : 2, kk = mode == MODE_TF32 ? 8 : 16;
    const double a_bytes = mode == MODE_TS ? 0 : 128.0 * kk * esz, b_bytes = (mode == MODE_SS2 ? n / 2 : n) * (double)kk * esz;
    const double floor_clk = (mode == MODE_SS2 ? 256.0 : 128.0) * n / (mode == MODE_SS2 ? 512.0 : 256.0) * (mode == MODE_TF32 ? 1.0 : 1.0);
    double tma_rate = 0;
    if (tma) tma_rate = chunks / grid * TMA_CHUNK / (med * a.iters);
    printf("%-10s N=%3d grid=%3d nA=%d tma=%d | clk/MMA median %7.1f min %7.1f max %7.1f | bf16-rate floor %5.0f | smem operand bytes/CTA/MMA %6.0f -> %5.1f B/clk"
           " | tma write %5.1f B/clk\n", name, n, grid, na, tma, med, c.front(), c.back(), floor_clk, a_bytes + b_bytes, (a_bytes + b_bytes) / med, tma_rate);
}

int main(int argc, char** argv) {
    int dev = 0; cudaSetDevice(dev);
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, dev);
    printf("device %s, %d SMs\n", prop.name, prop.multiProcessorCount);
    uint8_t* gsrc; cudaMalloc(&gsrc, (size_t)4096 * TMA_CHUNK); cudaMemset(gsrc, 0, (size_t)4096 * TMA_CHUNK);
    unsigned long long* dcyc; cudaMalloc(&dcyc, sizeof(unsigned long long) * 2 * 148);
    const int ns[] = {16, 32, 64, 96, 128, 160, 192, 256};
    for (int grid : {1, 148}) {
        for (int n : ns) run("ss", MODE_SS, n, grid, 4, 0, gsrc, dcyc);
        for (int n : ns) run("ts", MODE_TS, n, grid, 1, 0, gsrc, dcyc);
        for (int n : {32, 64, 128, 192, 256}) run("ss2", MODE_SS2, n, grid == 1 ? 2 : 148, 4, 0, gsrc, dcyc);
        for (int n : {64, 128, 256}) run("tf32", MODE_TF32, n, grid, 4, 0, gsrc, dcyc);
    }
    // same A slab every time vs four different slabs (is A cached between MMAs?)
    run("ss-sameA", MODE_SS, 64, 148, 1, 0, gsrc, dcyc);
    run("ss-sameA", MODE_SS, 128, 148, 1, 0, gsrc, dcyc);
    // with a concurrent bulk-copy stream into shared memory
    for (int n : {64, 128, 256}) run("ss+tma", MODE_SS, n, 148, 4, 1, gsrc, dcyc);
    for (int n : {64, 128, 256}) run("ts+tma", MODE_TS, n, 148, 1, 1, gsrc, dcyc);
    for (int n : {64, 128, 256}) run("ss2+tma", MODE_SS2, n, 148, 4, 1, gsrc, dcyc);
    return 0;
}

// Micro-benchmark: how to stream an HBM-bound elementwise pass on B200.
// Op = the BatchNorm-backward apply pass of the engine: out[p][c] = A[c]*g[p][c] + B[c]*x[p][c] + D[c] masked by x*sc+sh > 0
// (two bf16 NHWC inputs, one bf16 output, five per-channel fp32 coefficient vectors).  Variants:
//   v0  the engine's round-1/2 form: a thread keeps one 8-channel group (coefficients in registers), 2 pixels in flight,
//       __launch_bounds__(256, 2)
//   v1  bulk-staged: one elected thread streams 16 KB chunks of both inputs into a shared-memory ring with cp.async.bulk
//       (mbarrier complete_tx), 8 consumer warps read 16-byte vectors from the ring, coefficients in registers, direct 16-byte stores
//   v2  register-light: coefficients in shared memory, flat vector loop, 4 vectors in flight, up to 8 blocks per SM
//   v3  ceiling: out = g + x without coefficients (4 vectors in flight, 8 blocks per SM)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o profiles/_bin/ew_stream profiles/ew_stream_microbench.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
typedef __nv_bfloat16 bf16;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

struct Coef { const float *sc, *sh, *A, *B, *D; };

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return u;
}
__device__ __forceinline__ void ld8(const float* p, float* r) {
    float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w;
}
__device__ __forceinline__ uint4 op8(const uint4& gu, const uint4& xu, const float* sc, const float* sh, const float* A, const float* B,
                                     const float* D) {
    float g[8], x[8], r[8];
    unpack8(gu, g); unpack8(xu, x);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float gm = fmaf(x[i], sc[i], sh[i]) > 0.f ? g[i] : 0.f;
        r[i] = fmaf(A[i], gm, fmaf(B[i], x[i], D[i]));
    }
    return pack8(r);
}

// ---------------------------------------------------------------- v0
__global__ void __launch_bounds__(256, 2) v0_kernel(const bf16* __restrict__ g, const bf16* __restrict__ x, bf16* __restrict__ out, Coef k,
                                                    unsigned npix, int C) {
    const int cg = min(C / 8, 256), lanes = 256 / cg;
    const int c = (threadIdx.x % cg) * 8, lane = threadIdx.x / cg;
    float sc[8], sh[8], A[8], B[8], D[8];
    ld8(k.sc + c, sc); ld8(k.sh + c, sh); ld8(k.A + c, A); ld8(k.B + c, B); ld8(k.D + c, D);
    const unsigned step = gridDim.x * lanes;
    unsigned pix = blockIdx.x * lanes + lane;
    for (; pix + step < npix; pix += 2 * step) {
        uint4 xs[2], gs[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            xs[u] = *reinterpret_cast<const uint4*>(x + (size_t)(pix + u * step) * C + c);
            gs[u] = *reinterpret_cast<const uint4*>(g + (size_t)(pix + u * step) * C + c);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) *reinterpret_cast<uint4*>(out + (size_t)(pix + u * step) * C + c) = op8(gs[u], xs[u], sc, sh, A, B, D);
    }
    for (; pix < npix; pix += step)
        *reinterpret_cast<uint4*>(out + (size_t)pix * C + c) =
            op8(*reinterpret_cast<const uint4*>(g + (size_t)pix * C + c), *reinterpret_cast<const uint4*>(x + (size_t)pix * C + c), sc, sh, A, B, D);
}

// ---------------------------------------------------------------- v1: cp.async.bulk ring
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
            "selp.u32 %0, 1, 0, P1;\n"
            "}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
template <int STAGES, int CHUNK>      // CHUNK bytes per input per stage
__global__ void __launch_bounds__(288, 1) v1_kernel(const bf16* __restrict__ g, const bf16* __restrict__ x, bf16* __restrict__ out, Coef k,
                                                    size_t nbytes, int C) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * 2 * CHUNK);
    const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * STAGES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t nchunks = (nbytes + CHUNK - 1) / CHUNK;
    if (warp == 8) {
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (size_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
                mbar_wait(empty0 + 8 * s, ph ^ 1);
                const size_t off = ch * CHUNK;
                const uint32_t bytes = (uint32_t)min((size_t)CHUNK, nbytes - off);
                mbar_expect_tx(full0 + 8 * s, 2 * bytes);
                bulk_load(smem_u32(smem + (size_t)s * 2 * CHUNK), reinterpret_cast<const uint8_t*>(g) + off, bytes, full0 + 8 * s);
                bulk_load(smem_u32(smem + (size_t)s * 2 * CHUNK + CHUNK), reinterpret_cast<const uint8_t*>(x) + off, bytes, full0 + 8 * s);
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
        }
        return;
    }
    // consumers: 256 threads; CHUNK and 256*16 are multiples of the row size C*2, so a thread's channel group is fixed
    const int c = (int)((threadIdx.x * 8) % C);
    float sc[8], sh[8], A[8], B[8], D[8];
    ld8(k.sc + c, sc); ld8(k.sh + c, sh); ld8(k.A + c, A); ld8(k.B + c, B); ld8(k.D + c, D);
    int s = 0; uint32_t ph = 0;
    for (size_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        const size_t off = ch * CHUNK;
        const int bytes = (int)min((size_t)CHUNK, nbytes - off);
        mbar_wait(full0 + 8 * s, ph);
        const uint8_t* sg = smem + (size_t)s * 2 * CHUNK;
        const uint8_t* sx = sg + CHUNK;
#pragma unroll 4
        for (int o = threadIdx.x * 16; o < bytes; o += 256 * 16) {
            const uint4 gu = *reinterpret_cast<const uint4*>(sg + o), xu = *reinterpret_cast<const uint4*>(sx + o);
            *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(out) + off + o) = op8(gu, xu, sc, sh, A, B, D);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * s);
        if (++s == STAGES) { s = 0; ph ^= 1; }
    }
}

// ---------------------------------------------------------------- v2: coefficients in shared memory, flat loop
template <int U, int MINB>
__global__ void __launch_bounds__(256, MINB) v2_kernel(const bf16* __restrict__ g, const bf16* __restrict__ x, bf16* __restrict__ out, Coef k,
                                                       unsigned nvec, int C) {
    extern __shared__ __align__(16) float sco[];       // [5][C]
    for (int i = threadIdx.x; i < C; i += 256) {
        sco[i] = k.sc[i]; sco[C + i] = k.sh[i]; sco[2 * C + i] = k.A[i]; sco[3 * C + i] = k.B[i]; sco[4 * C + i] = k.D[i];
    }
    __syncthreads();
    const int cgt = C / 8;
    const unsigned stride = gridDim.x * 256;
    unsigned v = blockIdx.x * 256 + threadIdx.x;
    const uint4* g4 = reinterpret_cast<const uint4*>(g);
    const uint4* x4 = reinterpret_cast<const uint4*>(x);
    uint4* o4 = reinterpret_cast<uint4*>(out);
    for (; v + (U - 1) * stride < nvec; v += U * stride) {
        uint4 gs[U], xs[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { gs[u] = g4[v + u * stride]; xs[u] = x4[v + u * stride]; }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int c = (int)((v + u * stride) % cgt) * 8;
            o4[v + u * stride] = op8(gs[u], xs[u], sco + c, sco + C + c, sco + 2 * C + c, sco + 3 * C + c, sco + 4 * C + c);
        }
    }
    for (; v < nvec; v += stride) {
        const int c = (int)(v % cgt) * 8;
        o4[v] = op8(g4[v], x4[v], sco + c, sco + C + c, sco + 2 * C + c, sco + 3 * C + c, sco + 4 * C + c);
    }
}

// ---------------------------------------------------------------- v3: ceiling (no coefficients)
template <int U>
__global__ void __launch_bounds__(256, 8) v3_kernel(const bf16* __restrict__ g, const bf16* __restrict__ x, bf16* __restrict__ out, size_t nvec) {
    const size_t stride = (size_t)gridDim.x * 256;
    size_t v = (size_t)blockIdx.x * 256 + threadIdx.x;
    const uint4* g4 = reinterpret_cast<const uint4*>(g);
    const uint4* x4 = reinterpret_cast<const uint4*>(x);
    uint4* o4 = reinterpret_cast<uint4*>(out);
    for (; v + (U - 1) * stride < nvec; v += U * stride) {
        uint4 gs[U], xs[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { gs[u] = g4[v + u * stride]; xs[u] = x4[v + u * stride]; }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float a[8], b[8];
            unpack8(gs[u], a); unpack8(xs[u], b);
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] += b[i];
            o4[v + u * stride] = pack8(a);
        }
    }
    for (; v < nvec; v += stride) {
        float a[8], b[8];
        unpack8(g4[v], a); unpack8(x4[v], b);
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] += b[i];
        o4[v] = pack8(a);
    }
}

template <typename F>
static float time_ms(F f, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) f(i);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) f(i);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

int main() {
    const int NSETS = 4;              // rotate over 4 buffer sets so that small tensors are not served from L2
    struct Case { unsigned npix; int C; } cases[] = {{2097152, 64}, {524288, 64}, {131072, 128}, {32768, 256}, {8192, 512}};
    float* coef; CK(cudaMalloc(&coef, 5 * 512 * sizeof(float)));
    std::vector<float> h(5 * 512);
    for (int i = 0; i < 5 * 512; ++i) h[i] = 0.5f + 0.001f * (i % 97);
    CK(cudaMemcpy(coef, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
    Coef k{coef, coef + 512, coef + 1024, coef + 1536, coef + 2048};
    constexpr int ST = 4, CH = 16384;
    const int smem1 = ST * 2 * CH + 256;
    CK(cudaFuncSetAttribute(v1_kernel<ST, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1));
    constexpr int ST6 = 6;
    const int smem1b = ST6 * 2 * CH + 256;
    CK(cudaFuncSetAttribute(v1_kernel<ST6, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1b));
    constexpr int CH8 = 8192;
    const int smem1c = 4 * 2 * CH8 + 256;
    CK(cudaFuncSetAttribute(v1_kernel<4, CH8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1c));
    for (auto cs : cases) {
        const size_t n = (size_t)cs.npix * cs.C, bytes = n * 2;
        bf16 *g[NSETS], *x[NSETS], *o[NSETS];
        for (int s = 0; s < NSETS; ++s) {
            CK(cudaMalloc(&g[s], bytes)); CK(cudaMalloc(&x[s], bytes)); CK(cudaMalloc(&o[s], bytes));
            CK(cudaMemset(g[s], 0x3c, bytes)); CK(cudaMemset(x[s], 0x3d, bytes));
        }
        const size_t nvec = n / 8;
        const int C = cs.C;
        const unsigned npix = cs.npix;
        const int cg = C / 8 < 256 ? C / 8 : 256, lanes = 256 / cg;
        const int reps = 40;
        auto gbs = [&](float ms) { return 3.0 * bytes / ms / 1e6; };
        printf("npix %8u C %3d (%6.1f MB per tensor):", npix, C, bytes / 1e6);
        {
            long long b = ((long long)npix + lanes * 4 - 1) / (lanes * 4); if (b > 148 * 16) b = 148 * 16;
            float ms = time_ms([&](int i) { v0_kernel<<<(int)b, 256>>>(g[i % NSETS], x[i % NSETS], o[i % NSETS], k, npix, C); }, reps);
            printf("  v0 %6.1f us %5.0f GB/s |", ms * 1e3, gbs(ms));
        }
        {
            float ms = time_ms([&](int i) { v1_kernel<ST, CH><<<148, 288, smem1>>>(g[i % NSETS], x[i % NSETS], o[i % NSETS], k, bytes, C); }, reps);
            printf("  v1(4x16K) %6.1f us %5.0f GB/s |", ms * 1e3, gbs(ms));
            ms = time_ms([&](int i) { v1_kernel<ST6, CH><<<148, 288, smem1b>>>(g[i % NSETS], x[i % NSETS], o[i % NSETS], k, bytes, C); }, reps);
            printf("  v1(6x16K) %6.1f us %5.0f GB/s |", ms * 1e3, gbs(ms));
            ms = time_ms([&](int i) { v1_kernel<4, CH8><<<296, 288, smem1c>>>(g[i % NSETS], x[i % NSETS], o[i % NSETS], k, bytes, C); }, reps);
            printf("  v1(4x8K,2/SM) %6.1f us %5.0f GB/s |", ms * 1e3, gbs(ms));
        }
        {
            const int smem2 = 5 * C * sizeof(float);
            int blocks = (int)((nvec + 256 * 4 - 1) / (256 * 4)); if (blocks > 148 * 8) blocks = 148 * 8;
            float ms = time_ms([&](int i) { v2_kernel<4, 6><<<blocks, 256, smem2>>>(g[i % NSETS], x[i % NSETS], o[i % NSETS], k, (unsigned)nvec, C); }, reps);
            printf("  v2(U4) %6.1f us %5.0f GB/s |", ms * 1e3, gbs(ms));
            blocks = (int)((nvec + 256 * 2 - 1) / (256 * 2)); if (blocks > 148 * 8) blocks = 148 * 8;
            ms = time_ms([&](int i) { v2_kernel<2, 8><<<blocks, 256, smem2>>>(g[i % NSETS], x[i % NSETS], o[i % NSETS], k, (unsigned)nvec, C); }, reps);
            printf("  v2(U2) %6.1f us %5.0f GB/s |", ms * 1e3, gbs(ms));
        }
        {
            int blocks = (int)((nvec + 256 * 4 - 1) / (256 * 4)); if (blocks > 148 * 8) blocks = 148 * 8;
            float ms = time_ms([&](int i) { v3_kernel<4><<<blocks, 256>>>(g[i % NSETS], x[i % NSETS], o[i % NSETS], nvec); }, reps);
            printf("  v3 %6.1f us %5.0f GB/s", ms * 1e3, gbs(ms));
        }
        printf("\n");
        for (int s = 0; s < NSETS; ++s) { CK(cudaFree(g[s])); CK(cudaFree(x[s])); CK(cudaFree(o[s])); }
    }
    return 0;
}

// Known-answer test: may a tcgen05 shared-memory descriptor start at a row that is NOT a multiple of the 8-row swizzle atom?
// (Wanted for a halo-box weight-gradient kernel: one TMA box of (16+2) x (2+2) pixels serves all nine taps if a tap's 16-pixel row
// segment can be addressed at box_base + row * 128 B.)  MN-major operands, 128B swizzle, M = 128 (two 64-channel blocks through LBO,
// here LBO = 0: both halves read the same tile), N = 64, K = 16 pixels per instruction.
//   A tile: 40 rows (pixels) x 64 channels bf16, written with the TMA 128B-swizzle pattern (chunk j of row r at r*128 + ((j ^ (r & 7)) * 16))
//   D[m][n] = sum_{k<16} A[r0 + k][m] * B[k][n]        for r0 in {0, 1, 2, 3, 5, 8, 9, 17} and base_offset in {0, r0 & 7}
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o profiles/_bin/desc_offset_test profiles/desc_offset_test.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0, spins = 0;
    while (!ok) {
        asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (++spins > (1u << 24)) { printf("mbar timeout\n"); __trap(); }
    }
}
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout, uint32_t base_offset) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_offset & 7) << 49;
    d |= (uint64_t)layout << 61;
    return d;
}
__device__ __forceinline__ uint32_t idesc_bf16(int m, int n, bool a_mn, bool b_mn) {
    uint32_t d = 0;
    d |= 1u << 4; d |= 1u << 7; d |= 1u << 10;
    if (a_mn) d |= 1u << 15;
    if (b_mn) d |= 1u << 16;
    d |= (uint32_t)(n >> 3) << 17;
    d |= (uint32_t)(m >> 4) << 24;
    return d;
}

constexpr int A_ROWS = 40;
__global__ void __launch_bounds__(128, 1) test_kernel(const __nv_bfloat16* __restrict__ A /*[A_ROWS][64]*/, const __nv_bfloat16* __restrict__ B /*[16][64]*/,
                                                      float* __restrict__ D /*[128][64]*/, int r0, int base_offset) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                    // A_ROWS x 128 B (5 atoms)
    uint8_t* sB = smem + 8192;             // 16 x 128 B
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 8192 + 2048);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // software "TMA": 128B swizzle on the absolute row index (tile base is 1024-aligned)
    for (int i = threadIdx.x; i < A_ROWS * 8; i += 128) {
        const int r = i >> 3, j = i & 7;
        *reinterpret_cast<uint4*>(sA + r * 128 + ((j ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(A + r * 64 + j * 8);
    }
    for (int i = threadIdx.x; i < 16 * 8; i += 128) {
        const int r = i >> 3, j = i & 7;
        *reinterpret_cast<uint4*>(sB + r * 128 + ((j ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(B + r * 64 + j * 8);
    }
    const uint32_t done = smem_u32(bars);
    if (threadIdx.x == 0) { mbar_init(done, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_ptr;
    if (warp == 0) {
        uint32_t pred = 0;
        asm volatile("{\n.reg .b32 %%rx;\n.reg .pred %%px;\nelect.sync %%rx|%%px, %1;\n@%%px mov.s32 %0, 1;\n}\n" : "+r"(pred) : "r"(0xffffffffu));
        if (pred) {
            const uint64_t ad = smem_desc(smem_u32(sA) + r0 * 128, 0, 1024, 2, base_offset);
            const uint64_t bd = smem_desc(smem_u32(sB), 0, 1024, 2, 0);
            const uint32_t id = idesc_bf16(128, 64, true, true);
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                         ::"r"(tmem), "l"(ad), "l"(bd), "r"(id), "r"(0) : "memory");
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(done) : "memory");
        }
        __syncwarp();
    }
    mbar_wait(done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int ch = 0; ch < 2; ++ch) {
        uint32_t r[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
              "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
              "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(tmem + ((uint32_t)(warp * 32) << 16) + ch * 32));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 32; ++i) D[(warp * 32 + lane) * 64 + ch * 32 + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory");
    }
}

int main() {
    std::vector<__nv_bfloat16> hA(A_ROWS * 64), hB(16 * 64);
    std::vector<float> fA(A_ROWS * 64), fB(16 * 64);
    for (int i = 0; i < A_ROWS * 64; ++i) { fA[i] = (float)((i * 7 + (i >> 6) * 3) % 13 - 6); hA[i] = __float2bfloat16(fA[i]); }
    for (int i = 0; i < 16 * 64; ++i) { fB[i] = (float)((i * 5 + (i >> 6)) % 11 - 5); hB[i] = __float2bfloat16(fB[i]); }
    __nv_bfloat16 *dA, *dB; float* dD;
    CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, 128 * 64 * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    const int smem = 1024 + 8192 + 2048 + 64;
    CK(cudaFuncSetAttribute(test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    std::vector<float> hD(128 * 64);
    for (int r0 : {0, 1, 2, 3, 5, 8, 9, 17}) {
        for (int mode = 0; mode < 2; ++mode) {
            const int bo = mode ? (r0 & 7) : 0;
            if (mode && bo == 0) continue;
            test_kernel<<<1, 128, smem>>>(dA, dB, dD, r0, bo);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
            int bad = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < 64; ++n) {
                    float ref = 0.f;
                    for (int k = 0; k < 16; ++k) ref += fA[(r0 + k) * 64 + (m & 63)] * fB[k * 64 + n];
                    if (ref != hD[m * 64 + n]) ++bad;
                }
            printf("A start row %2d  base_offset %d : %s (%d of 8192 wrong)\n", r0, bo, bad ? "MISMATCH" : "exact", bad);
        }
    }
    return 0;
}

// Shared sm_100a building blocks for the tensor-core kernels: mbarrier / TMA / tcgen05 PTX wrappers, descriptor
// encoders and the host-side tensor-map helper.
#pragma once
#include "common.cuh"
#include "tc_epilogue.h"
#include <cuda.h>
#include <stdexcept>
#include <string>

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a mis-programmed pipeline traps (-> CUDA error on the host) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 24)) {
            printf("tc: mbarrier wait timed out (block %d,%d,%d thread %d bar 0x%x parity %u)\n", blockIdx.x, blockIdx.y, blockIdx.z,
                   threadIdx.x, bar, parity);
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// 32 lanes x 32 consecutive fp32 columns: thread (lane) gets its row's 32 values
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// the same load without the wait: the registers must not be read before tmem_ld_wait32(v), which carries them as in/out operands
// so that the compiler cannot move a use above the wait
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait32(float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.wait::ld.sync.aligned;"
        : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
          "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]),
          "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]),
          "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
        :: "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n"
        ".reg .b32 %%rx;\n"
        ".reg .pred %%px;\n"
        "elect.sync %%rx|%%px, %1;\n"
        "@%%px mov.s32 %0, 1;\n"
        "}\n" : "+r"(pred) : "r"(0xffffffffu));
    return pred != 0;
}

// Shared-memory matrix descriptor (PTX "matrix descriptor", Blackwell version 1).
//   start address (>>4) [0,14) | leading byte offset (>>4) [16,30) | stride byte offset (>>4) [32,46) | version=1 [46,48)
//   | layout type [61,64): 0 none, 2 = 128B swizzle, 4 = 64B swizzle
__device__ __forceinline__ uint64_t smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
// Instruction descriptor for kind::f16: bf16 x bf16 -> fp32, M = 128.  a_mn / b_mn: operand is MN-major (transposed).
__device__ __forceinline__ uint32_t instr_desc_bf16(int n, bool a_mn, bool b_mn) {
    uint32_t d = 0;
    d |= 1u << 4;                  // D format  F32
    d |= 1u << 7;                  // A format  BF16
    d |= 1u << 10;                 // B format  BF16
    d |= (a_mn ? 1u : 0u) << 15;
    d |= (b_mn ? 1u : 0u) << 16;
    d |= (uint32_t)(n >> 3) << 17;
    d |= (uint32_t)(128 >> 4) << 24;
    return d;
}

// n / d for n < 2^31 with a precomputed multiplier: one __umulhi + shift instead of a ~25-instruction division sequence.  The
// tile -> (channel tile, x, y, image) map is evaluated per tile by every thread of the CTA; its five divisions were 17 % of the
// stall samples of a layer1 launch (profiles/r2_notes.md).
struct FastDiv {
    uint32_t d, mul, shr;
    __device__ __forceinline__ uint32_t div(uint32_t n) const { return d == 1 ? n : __umulhi(n, mul) >> shr; }
};
static inline FastDiv make_fastdiv(int d) {
    FastDiv f; f.d = (uint32_t)d; f.mul = 0; f.shr = 0;
    if (d > 1) {
        int lg = 0;
        while ((1u << lg) < (uint32_t)d) ++lg;
        const unsigned p = 31 + lg;
        f.mul = (uint32_t)(((1ull << p) + (uint32_t)d - 1) / (uint32_t)d);
        f.shr = p - 32;
    }
    return f;
}

// ------------------------------------------------------------------------------------------------ fused epilogue tail
// What the convolution epilogues can do with a finished fp32 tile besides the bias: eval-mode BatchNorm folded to one per-channel
// affine (scale, shift), a residual add, ReLU, and a store into a tensor with a physical replicate border (the decoder's
// ReplicationPad2d((0,2,2,0)) inputs) - so that an inference forward needs no separate BatchNorm / ReLU / border pass.
__device__ __forceinline__ void epi_affine32(float* v, const float* __restrict__ sc, const float* __restrict__ sh) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(sc) + q), b = __ldg(reinterpret_cast<const float4*>(sh) + q);
        v[4 * q + 0] = fmaf(v[4 * q + 0], a.x, b.x); v[4 * q + 1] = fmaf(v[4 * q + 1], a.y, b.y);
        v[4 * q + 2] = fmaf(v[4 * q + 2], a.z, b.z); v[4 * q + 3] = fmaf(v[4 * q + 3], a.w, b.w);
    }
}
__device__ __forceinline__ void epi_add32(float* v, const uint4* old) {        // + 32 bf16 values held in 4 registers quads
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const __nv_bfloat162* ob = reinterpret_cast<const __nv_bfloat162*>(&old[q]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = __bfloat1622float2(ob[j]);
            v[q * 8 + 2 * j] += f.x; v[q * 8 + 2 * j + 1] += f.y;
        }
    }
}
__device__ __forceinline__ void epi_add32(float* v, const float* __restrict__ r) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const float4 a = *(reinterpret_cast<const float4*>(r) + q);
        v[4 * q + 0] += a.x; v[4 * q + 1] += a.y; v[4 * q + 2] += a.z; v[4 * q + 3] += a.w;
    }
}
__device__ __forceinline__ void epi_pack32(const float* v, uint4* pk) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        __nv_bfloat162 b0 = __floats2bfloat162_rn(v[q * 8 + 0], v[q * 8 + 1]);
        __nv_bfloat162 b1 = __floats2bfloat162_rn(v[q * 8 + 2], v[q * 8 + 3]);
        __nv_bfloat162 b2 = __floats2bfloat162_rn(v[q * 8 + 4], v[q * 8 + 5]);
        __nv_bfloat162 b3 = __floats2bfloat162_rn(v[q * 8 + 6], v[q * 8 + 7]);
        pk[q].x = *reinterpret_cast<uint32_t*>(&b0); pk[q].y = *reinterpret_cast<uint32_t*>(&b1);
        pk[q].z = *reinterpret_cast<uint32_t*>(&b2); pk[q].w = *reinterpret_cast<uint32_t*>(&b3);
    }
}
// 256-bit global store (sm_100: STG.256): a thread's 32 bf16 channels leave in two full 32-byte sectors instead of four half
// sectors - the epilogue's scattered 16-byte stores were what bounded the K = 576 layers (profiles/r2_notes.md).  32-byte aligned.
__device__ __forceinline__ void st_global_256(void* p, const uint32_t* r) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
// store 32 channels of one pixel to every physical position it owns: rows py0..py1, columns px0..px1 (one position unless on a border)
__device__ __forceinline__ void epi_store32(const float* v, bf16* obase, int py0, int py1, int px0, int px1, int Wp, int Co) {
    uint4 pk[4];
    epi_pack32(v, pk);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(pk);
    for (int yy = py0; yy <= py1; ++yy)
        for (int xx = px0; xx <= px1; ++xx) {
            bf16* o = obase + ((size_t)yy * Wp + xx) * Co;
            st_global_256(o, w);
            st_global_256(o + 16, w + 8);
        }
}
__device__ __forceinline__ void epi_store32(const float* v, float* obase, int py0, int py1, int px0, int px1, int Wp, int Co) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(v);
    for (int yy = py0; yy <= py1; ++yy)
        for (int xx = px0; xx <= px1; ++xx) {
            float* o = obase + ((size_t)yy * Wp + xx) * Co;
#pragma unroll
            for (int q = 0; q < 4; ++q) st_global_256(o + 8 * q, w + 8 * q);
        }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        cudaDriverEntryPointQueryResult qres;
        void* ptr = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess || !ptr)
            throw std::runtime_error("cuTensorMapEncodeTiled not available from the driver");
        fn = (PFN_encodeTiled)ptr;
    }
    return fn;
}
// NHWC bf16 activation [B,H,W,C] as a 4-D tensor map (C, W, H, B); box in *loaded* elements, `stride` = element stride in W,H
inline CUtensorMap make_map_nhwc(const void* base, int C, int W, int H, int B, int boxC, int boxW, int boxH, int boxB, int stride,
                                 CUtensorMapSwizzle swz) {
    CUtensorMap m;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)boxC, (cuuint32_t)(boxW * stride), (cuuint32_t)(boxH * stride), (cuuint32_t)boxB};
    cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled(NHWC) failed with code " + std::to_string((int)r));
    return m;
}
inline int num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n > SALT_STAT_SLOTS_CONV) n = SALT_STAT_SLOTS_CONV;      // persistent grids index BatchNorm partial slots by blockIdx.x
    }
    return n;
}

}  // namespace tc

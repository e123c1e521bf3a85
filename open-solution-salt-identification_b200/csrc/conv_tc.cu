// tcgen05 implicit-GEMM convolution for sm_100a (forward and stride-1 dgrad).
//
//   out[n,y,x,k] (+)= sum_{tap=(r,s)} sum_c  A[n, y*stride + r - pad, x*stride + s - pad, c] * Wp[k][tap*C + c]   (zero outside A)
//
// GEMM view: M = 128 output pixels per tile (a tn x th x tw box of the NHWC output), N = BN output channels, K = taps * C.
// One K-block = one (tap, 64-channel block): the A operand of that block is a TMA *box* load of the input tensor at the
// tap-shifted coordinate (im2col done by the TMA unit; out-of-bounds rows are zero-filled = the zero padding), the B operand
// a [BN x 64] slab of the packed weights.  Both land in shared memory in the canonical K-major 128B-swizzled layout and feed
// tcgen05.mma (M128 x N BN x K16, bf16 in, fp32 accumulate in TMEM).
//
// Persistent, warp-specialised CTA (1 per SM): warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2-5 =
// epilogue (tcgen05.ld -> +bias -> bf16 -> global, per-channel sum / sum^2 for BatchNorm via warp-shuffle butterflies).
// Two TMEM accumulators so the epilogue of tile i overlaps the MMAs of tile i+1; an S-stage smem ring decouples TMA from MMA.
#include "kernels.h"
#include "conv_tc.h"
#include <cuda.h>
#include <map>
#include <tuple>
#include <stdexcept>
#include <string>

// ------------------------------------------------------------------------------------------------ PTX wrappers
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
            printf("conv_tc: mbarrier wait timed out (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
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
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
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

// shared-memory matrix descriptor, canonical K-major layout, 128B (BK=64) or 64B (BK=32) swizzle
template <int BK>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    constexpr uint64_t row_bytes = BK * 2;                 // 128 or 64
    constexpr uint64_t sbo = 8 * row_bytes;                // stride between 8-row groups
    constexpr uint64_t layout = BK == 64 ? 2 : 4;          // SWIZZLE_128B : SWIZZLE_64B
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);           // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                                // leading byte offset (ignored for swizzled K-major), bits [16,30)
    d |= (uint64_t)(sbo >> 4) << 32;                       // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                                // descriptor version (Blackwell)
    d |= layout << 61;                                     // layout type, bits [61,64)
    return d;
}
template <int BN>
__device__ __forceinline__ uint32_t make_idesc() {
    uint32_t d = 0;
    d |= 1u << 4;                  // D format  F32
    d |= 1u << 7;                  // A format  BF16
    d |= 1u << 10;                 // B format  BF16
    d |= (uint32_t)(BN >> 3) << 17;
    d |= (uint32_t)(128 >> 4) << 24;
    return d;                      // A, B K-major, dense, no negate
}

// 32 per-lane values (one per channel) x 32 lanes (pixels) -> lane l holds the sum over pixels of channel l
template <int S>
__device__ __forceinline__ void butterfly_step(float* v, bool upper) {
#pragma unroll
    for (int i = 0; i < S; ++i) {
        float send = upper ? v[i] : v[i + S];
        float keep = upper ? v[i + S] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, S);
    }
}
__device__ __forceinline__ float butterfly_reduce32(float* v, int lane) {
    butterfly_step<16>(v, lane & 16);
    butterfly_step<8>(v, lane & 8);
    butterfly_step<4>(v, lane & 4);
    butterfly_step<2>(v, lane & 2);
    butterfly_step<1>(v, lane & 1);
    return v[0];
}

struct TcParams {
    int tw, th, tn;
    int tiles_x, tiles_y, tiles_b, tiles_co;
    int B, Ho, Wo, Co;
    int taps, S, cblks, stride, pad;
    int accumulate;
    const float* bias;
    double* stats;
    bf16* out;
};

constexpr int TC_THREADS = 192;
template <int BN, int BK> struct TcCfg {
    static constexpr int A_BYTES = 128 * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int BUDGET = 200 * 1024;
    static constexpr int STAGES_RAW = BUDGET / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
    static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 2 * BN * 4 /*stats*/;
};

template <int BN, int BK>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const TcParams p) {
    using Cfg = TcCfg<BN, BK>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    // bars: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
    float* s_stats = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES);
    const uint32_t tfull0 = smem_u32(bars + 2 * STAGES), tempty0 = smem_u32(bars + 2 * STAGES + 2);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        for (int i = 0; i < STAGES; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(tfull0 + 8 * i, 1); mbar_init(tempty0 + 8 * i, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_ptr_smem), Cfg::TMEM_COLS);
    for (int i = threadIdx.x; i < 2 * BN; i += TC_THREADS) s_stats[i] = 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const int total_tiles = p.tiles_x * p.tiles_y * p.tiles_b * p.tiles_co;
    const int num_kb = p.taps * p.cblks;

    if (warp == 0) {
        // ===================================================== TMA producer
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int nt = tile % p.tiles_co, mt = tile / p.tiles_co;
                const int tx = mt % p.tiles_x, ty = (mt / p.tiles_x) % p.tiles_y, tb = mt / (p.tiles_x * p.tiles_y);
                const int w0 = tx * p.tw * p.stride - p.pad, h0 = ty * p.th * p.stride - p.pad, n0 = tb * p.tn;
                for (int kb = 0; kb < num_kb; ++kb) {
                    const int tap = kb / p.cblks, cb = kb - tap * p.cblks;
                    const int r = tap / p.S, s = tap - r * p.S;
                    mbar_wait(empty0 + 8 * stage, phase ^ 1);
                    mbar_expect_tx(full0 + 8 * stage, Cfg::STAGE_BYTES);
                    tma_load_4d(smem_u32(smem_a + stage * Cfg::A_BYTES), &map_a, full0 + 8 * stage, cb * BK, w0 + s, h0 + r, n0);
                    tma_load_2d(smem_u32(smem_b + stage * Cfg::B_BYTES), &map_b, full0 + 8 * stage, kb * BK, nt * BN);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        const uint32_t idesc = make_idesc<BN>();
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * BN;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(full0 + 8 * stage, phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t adesc = make_smem_desc<BK>(smem_u32(smem_a + stage * Cfg::A_BYTES));
                    const uint64_t bdesc = make_smem_desc<BK>(smem_u32(smem_b + stage * Cfg::B_BYTES));
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)       // advance 16 bf16 = 32 bytes inside the swizzle atom
                        umma_bf16(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                    umma_commit(empty0 + 8 * stage);        // frees the smem stage when these MMAs retire
                    if (kb == num_kb - 1) umma_commit(tfull0 + 8 * acc);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else {
        // ===================================================== epilogue (warps 2..5 <-> TMEM lane quarters 2,3,0,1)
        const int quarter = warp & 3;
        const int m = quarter * 32 + lane;                  // row of the 128-pixel tile
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int nt = tile % p.tiles_co, mt = tile / p.tiles_co;
            const int tx = mt % p.tiles_x, ty = (mt / p.tiles_x) % p.tiles_y, tb = mt / (p.tiles_x * p.tiles_y);
            const int lx = m % p.tw, ly = (m / p.tw) % p.th, ln = m / (p.tw * p.th);
            const int x = tx * p.tw + lx, y = ty * p.th + ly, n = tb * p.tn + ln;
            const bool valid = (n < p.B) && (y < p.Ho) && (x < p.Wo);
            bf16* orow = p.out + (((size_t)n * p.Ho + y) * p.Wo + x) * p.Co + nt * BN;
            mbar_wait(tfull0 + 8 * acc, acc_phase);
            tc_fence_after();
#pragma unroll 1
            for (int ch = 0; ch < BN / 32; ++ch) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + ch * 32, v);
                if (p.bias) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] += __ldg(p.bias + nt * BN + ch * 32 + i);
                }
                if (valid) {
                    uint4* o4 = reinterpret_cast<uint4*>(orow + ch * 32);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (p.accumulate) {
                            uint4 old = o4[q];
                            const __nv_bfloat162* ob = reinterpret_cast<const __nv_bfloat162*>(&old);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                float2 f = __bfloat1622float2(ob[j]);
                                v[q * 8 + 2 * j] += f.x; v[q * 8 + 2 * j + 1] += f.y;
                            }
                        }
                        uint4 pk;
                        __nv_bfloat162 b0 = __floats2bfloat162_rn(v[q * 8 + 0], v[q * 8 + 1]);
                        __nv_bfloat162 b1 = __floats2bfloat162_rn(v[q * 8 + 2], v[q * 8 + 3]);
                        __nv_bfloat162 b2 = __floats2bfloat162_rn(v[q * 8 + 4], v[q * 8 + 5]);
                        __nv_bfloat162 b3 = __floats2bfloat162_rn(v[q * 8 + 6], v[q * 8 + 7]);
                        pk.x = *reinterpret_cast<uint32_t*>(&b0); pk.y = *reinterpret_cast<uint32_t*>(&b1);
                        pk.z = *reinterpret_cast<uint32_t*>(&b2); pk.w = *reinterpret_cast<uint32_t*>(&b3);
                        o4[q] = pk;
                    }
                }
                if (p.stats) {
                    float sq[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) { v[i] = valid ? v[i] : 0.f; sq[i] = v[i] * v[i]; }
                    float s1 = butterfly_reduce32(v, lane);
                    float s2 = butterfly_reduce32(sq, lane);
                    atomicAdd(s_stats + ch * 32 + lane, s1);
                    atomicAdd(s_stats + BN + ch * 32 + lane, s2);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            if (p.stats && p.tiles_co > 1) {
                // the per-CTA accumulators are per output-channel tile: flush when the channel tile changes
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const int t = threadIdx.x - 64;
                for (int i = t; i < 2 * BN; i += 128) {
                    float val = s_stats[i];
                    if (val != 0.f) atomicAdd(p.stats + (i < BN ? nt * BN + i : p.Co + nt * BN + (i - BN)), (double)val);
                    s_stats[i] = 0.f;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
        }
        if (p.stats && p.tiles_co == 1) {
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const int t = threadIdx.x - 64;
            for (int i = t; i < 2 * BN; i += 128)
                atomicAdd(p.stats + (i < BN ? i : p.Co + (i - BN)), (double)s_stats[i]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, Cfg::TMEM_COLS); }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode() {
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
static CUtensorMap make_map_4d(const void* base, int C, int W, int H, int B, int boxC, int boxW, int boxH, int boxB, int stride, bool sw64) {
    CUtensorMap m;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)boxC, (cuuint32_t)boxW, (cuuint32_t)boxH, (cuuint32_t)boxB};
    cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, sw64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled(A) failed with code " + std::to_string((int)r));
    return m;
}
static CUtensorMap make_map_2d(const void* base, int K, int N, int boxK, int boxN, bool sw64) {
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)boxK, (cuuint32_t)boxN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, sw64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled(B) failed with code " + std::to_string((int)r));
    return m;
}

bool tc_conv_supported(const ConvGeom& g, bool dgrad) {
    const int Cred = dgrad ? g.Co : g.Ci, Nout = dgrad ? g.Ci : g.Co;
    const int Wd = dgrad ? g.Wi : g.Wo, Hd = dgrad ? g.Hi : g.Ho;
    if (dgrad && g.stride != 1) return false;
    if (Cred % 32) return false;
    if (Nout % 32) return false;
    if (Wd < 8 || Hd < 8) return false;
    if (g.stride != 1 && g.stride != 2) return false;
    return true;
}

template <int BN, int BK>
static void launch_tc(cudaStream_t st, const CUtensorMap& ma, const CUtensorMap& mb, const TcParams& p, int total_tiles) {
    using Cfg = TcCfg<BN, BK>;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(conv_tc_kernel<BN, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        configured = true;
    }
    static int num_sms = 0;
    if (!num_sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev); }
    int grid = total_tiles < num_sms ? total_tiles : num_sms;
    conv_tc_kernel<BN, BK><<<grid, TC_THREADS, Cfg::SMEM_BYTES, st>>>(ma, mb, p);
}

// A: activation tensor [B,Ha,Wa,Ca] bf16 (physical dims);  Wp: packed weights [Nout][taps*Ca] bf16;  out: [B,Ho,Wo,Nout] bf16
void k_conv_tc(cudaStream_t st, const void* A, int B, int Ha, int Wa, int Ca, const void* Wp, int Nout, int R, int S, int stride,
               int pad, void* out, int Ho, int Wo, const float* bias, double* stats, bool accumulate) {
    SALT_COUNT(1);
    TcParams p;
    p.tw = Wo >= 16 ? 16 : 8;
    p.th = p.tw == 16 ? 8 : (Ho >= 16 ? 16 : 8);
    p.tn = 128 / (p.tw * p.th);
    p.tiles_x = cdiv(Wo, p.tw); p.tiles_y = cdiv(Ho, p.th); p.tiles_b = cdiv(B, p.tn);
    const int BK = (Ca % 64 == 0) ? 64 : 32;
    int BN = Nout % 256 == 0 ? 256 : Nout % 128 == 0 ? 128 : Nout % 64 == 0 ? 64 : 32;
    // prefer more, smaller channel tiles when the tile count would leave SMs idle
    while (BN > 64 && (long long)p.tiles_x * p.tiles_y * p.tiles_b * (Nout / BN) < 148) BN >>= 1;
    p.tiles_co = Nout / BN;
    p.B = B; p.Ho = Ho; p.Wo = Wo; p.Co = Nout;
    p.taps = R * S; p.S = S; p.cblks = Ca / BK; p.stride = stride; p.pad = pad;
    p.accumulate = accumulate ? 1 : 0; p.bias = bias; p.stats = stats; p.out = (bf16*)out;
    const bool sw64 = BK == 32;
    CUtensorMap ma = make_map_4d(A, Ca, Wa, Ha, B, BK, p.tw * stride, p.th * stride, p.tn, stride, sw64);
    CUtensorMap mb = make_map_2d(Wp, R * S * Ca, Nout, BK, BN, sw64);
    const int total = p.tiles_x * p.tiles_y * p.tiles_b * p.tiles_co;
#define TC_CASE(bn, bk) if (BN == bn && BK == bk) { launch_tc<bn, bk>(st, ma, mb, p, total); return; }
    TC_CASE(256, 64) TC_CASE(128, 64) TC_CASE(64, 64) TC_CASE(32, 64)
    TC_CASE(256, 32) TC_CASE(128, 32) TC_CASE(64, 32) TC_CASE(32, 32)
#undef TC_CASE
    throw std::runtime_error("k_conv_tc: unsupported tile configuration");
}

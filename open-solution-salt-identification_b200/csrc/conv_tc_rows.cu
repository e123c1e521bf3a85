// tcgen05 3x3 stride-1 convolution with shared-memory ROW-HALO reuse of the activation tile (forward and stride-1 dgrad).
//
// conv_tc.cu loads the A operand once per filter tap (9 TMA boxes per 64-channel block).  Its profile shows the kernel bound
// by L2 -> shared-memory bandwidth, not by the tensor core.  Here the output tile is 16 rows x 8 columns and, per 64-channel
// block, only THREE boxes are loaded - one per horizontal tap offset s - each 18 rows x 8 pixels (the 16 output rows plus the
// vertical halo).  A pixel row of 8 pixels x 64 channels is exactly one 1024-byte swizzle atom, so the three vertical taps
// r = 0,1,2 are the SAME buffer read at byte offsets r*1024: the UMMA descriptor start address stays 1024-byte aligned and the
// canonical K-major 128B-swizzled layout is untouched.  A traffic drops from 9 x 16 KB to 3 x 18 KB per channel block.
//
// Pipeline: A ring (18 KB stages, one per (channel block, s)), B ring (weights [BN x 64], one per tap), warp 0 = TMA producer,
// warp 1 = MMA issuer, warps 2-5 = epilogue (same epilogue as conv_tc.cu), two TMEM accumulators, persistent CTAs.
#include "tc_common.cuh"
#include "conv_tc.h"

using namespace tc;

template <int S>
__device__ __forceinline__ void rows_butterfly_step(float* v, bool upper) {
#pragma unroll
    for (int i = 0; i < S; ++i) {
        float send = upper ? v[i] : v[i + S];
        float keep = upper ? v[i + S] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, S);
    }
}
__device__ __forceinline__ float rows_butterfly_reduce32(float* v, int lane) {
    rows_butterfly_step<16>(v, lane & 16);
    rows_butterfly_step<8>(v, lane & 8);
    rows_butterfly_step<4>(v, lane & 4);
    rows_butterfly_step<2>(v, lane & 2);
    rows_butterfly_step<1>(v, lane & 1);
    return v[0];
}

struct RowsParams {
    int tiles_x, tiles_y, tiles_co, total_tiles;
    int B, Ho, Wo, Co, Ca, cblks, pad;
    int accumulate;
    const float* bias;
    double* stats;
    bf16* out;
};

constexpr int RW_THREADS = 192;
constexpr int RW_TH = 16, RW_TW = 8;
constexpr int RW_A_BYTES = (RW_TH + 2) * RW_TW * 128;        // 18 pixel rows x 8 pixels x 64 channels bf16 = 18 KB
template <int BN> struct RowsCfg {
    static constexpr int B_BYTES = BN * 128;
    static constexpr int A_STAGES = 4;
    static constexpr int B_STAGES = BN == 256 ? 3 : (BN == 128 ? 6 : 8);
    static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
    static constexpr int BAR_OFF = A_STAGES * RW_A_BYTES + B_STAGES * B_BYTES;
    static constexpr int SMEM_BYTES = BAR_OFF + 1024 + 256 + 2 * BN * 4;
};

template <int BN>
__global__ void __launch_bounds__(RW_THREADS, 1)
conv_tc_rows_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const RowsParams p) {
    using Cfg = RowsCfg<BN>;
    constexpr int SA = Cfg::A_STAGES, SB = Cfg::B_STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + SA * RW_A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
    // bars: a_full[SA], a_empty[SA], b_full[SB], b_empty[SB], tmem_full[2], tmem_empty[2]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * SA + 2 * SB + 4);
    float* s_stats = reinterpret_cast<float*>(smem + Cfg::BAR_OFF + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t afull0 = smem_u32(bars), aempty0 = afull0 + 8 * SA, bfull0 = aempty0 + 8 * SA, bempty0 = bfull0 + 8 * SB;
    const uint32_t tfull0 = bempty0 + 8 * SB, tempty0 = tfull0 + 16;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        for (int i = 0; i < SA; ++i) { mbar_init(afull0 + 8 * i, 1); mbar_init(aempty0 + 8 * i, 1); }
        for (int i = 0; i < SB; ++i) { mbar_init(bfull0 + 8 * i, 1); mbar_init(bempty0 + 8 * i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(tfull0 + 8 * i, 1); mbar_init(tempty0 + 8 * i, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_ptr_smem), Cfg::TMEM_COLS);
    for (int i = threadIdx.x; i < 2 * BN; i += RW_THREADS) s_stats[i] = 0.f;
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        // ===================================================== TMA producer
        if (elect_one()) {
            int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const int nt = tile % p.tiles_co, mt = tile / p.tiles_co;
                const int tx = mt % p.tiles_x, ty = (mt / p.tiles_x) % p.tiles_y, n = mt / (p.tiles_x * p.tiles_y);
                const int w0 = tx * RW_TW - p.pad, h0 = ty * RW_TH - p.pad;
                for (int cb = 0; cb < p.cblks; ++cb) {
                    for (int s = 0; s < 3; ++s) {
                        mbar_wait(aempty0 + 8 * sa, pa ^ 1);
                        mbar_expect_tx(afull0 + 8 * sa, RW_A_BYTES);
                        tma_load_4d(smem_u32(smem_a + sa * RW_A_BYTES), &map_a, afull0 + 8 * sa, cb * 64, w0 + s, h0, n);
                        if (++sa == SA) { sa = 0; pa ^= 1; }
                        for (int r = 0; r < 3; ++r) {
                            mbar_wait(bempty0 + 8 * sb, pb ^ 1);
                            mbar_expect_tx(bfull0 + 8 * sb, Cfg::B_BYTES);
                            tma_load_2d(smem_u32(smem_b + sb * Cfg::B_BYTES), &map_b, bfull0 + 8 * sb, (r * 3 + s) * p.Ca + cb * 64, nt * BN);
                            if (++sb == SB) { sb = 0; pb ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        const uint32_t idesc = instr_desc_bf16(BN, false, false);
        const uint64_t adesc0 = smem_desc(smem_u32(smem_a), 16, 1024, 2);
        const uint64_t bdesc0 = smem_desc(smem_u32(smem_b), 16, 1024, 2);
        int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);
            fence_after();
            const uint32_t tmem_d = tmem_base + acc * BN;
            const int num_a = p.cblks * 3;
            for (int ia = 0; ia < num_a; ++ia) {
                mbar_wait(afull0 + 8 * sa, pa);
                for (int r = 0; r < 3; ++r) {
                    mbar_wait(bfull0 + 8 * sb, pb);
                    fence_after();
                    if (elect_one()) {
                        // vertical tap r = the same A buffer, r pixel-rows (r * 1024 bytes) further down
                        const uint64_t adesc = adesc0 + (uint64_t)((sa * RW_A_BYTES + r * 1024) >> 4);
                        const uint64_t bdesc = bdesc0 + (uint64_t)((sb * Cfg::B_BYTES) >> 4);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (ia | r | k) != 0);
                        umma_commit(bempty0 + 8 * sb);
                        if (r == 2) umma_commit(aempty0 + 8 * sa);
                        if (r == 2 && ia == num_a - 1) umma_commit(tfull0 + 8 * acc);
                    }
                    __syncwarp();
                    if (++sb == SB) { sb = 0; pb ^= 1; }
                }
                if (++sa == SA) { sa = 0; pa ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else {
        // ===================================================== epilogue (warps 2..5 <-> TMEM lane quarters 2,3,0,1)
        const int quarter = warp & 3;
        const int m = quarter * 32 + lane;                  // row of the 128-position tile: 16 rows x 8 columns
        const int lx = m & (RW_TW - 1), ly = m >> 3;
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            const int nt = tile % p.tiles_co, mt = tile / p.tiles_co;
            const int tx = mt % p.tiles_x, ty = (mt / p.tiles_x) % p.tiles_y, n = mt / (p.tiles_x * p.tiles_y);
            const int x = tx * RW_TW + lx, y = ty * RW_TH + ly;
            const bool valid = (y < p.Ho) && (x < p.Wo);
            bf16* orow = p.out + (((size_t)n * p.Ho + y) * p.Wo + x) * p.Co + nt * BN;
            mbar_wait(tfull0 + 8 * acc, acc_phase);
            fence_after();
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
                    float s1 = rows_butterfly_reduce32(v, lane);
                    float s2 = rows_butterfly_reduce32(sq, lane);
                    atomicAdd(s_stats + ch * 32 + lane, s1);
                    atomicAdd(s_stats + BN + ch * 32 + lane, s2);
                }
            }
            fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            if (p.stats && p.tiles_co > 1) {
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const int tt = threadIdx.x - 64;
                for (int i = tt; i < 2 * BN; i += 128) {
                    float val = s_stats[i];
                    if (val != 0.f) atomicAdd(p.stats + (i < BN ? nt * BN + i : p.Co + nt * BN + (i - BN)), (double)val);
                    s_stats[i] = 0.f;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
        }
        if (p.stats && p.tiles_co == 1) {
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const int tt = threadIdx.x - 64;
            for (int i = tt; i < 2 * BN; i += 128)
                atomicAdd(p.stats + (i < BN ? i : p.Co + (i - BN)), (double)s_stats[i]);
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 1) { fence_after(); tmem_dealloc(tmem_base, Cfg::TMEM_COLS); }
}

bool tc_conv_rows_supported(int Ca, int Nout, int R, int S, int stride, int Ho, int Wo) {
    return R == 3 && S == 3 && stride == 1 && Ca % 64 == 0 && Nout % 32 == 0 && Ho >= 16 && Wo >= 8;
}

template <int BN>
static void launch_rows(cudaStream_t st, const CUtensorMap& ma, const CUtensorMap& mb, const RowsParams& p) {
    using Cfg = RowsCfg<BN>;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(conv_tc_rows_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        configured = true;
    }
    const int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
    conv_tc_rows_kernel<BN><<<grid, RW_THREADS, Cfg::SMEM_BYTES, st>>>(ma, mb, p);
}

// out[n,y,x,k] (+)= sum_{r,s,c} A[n, y+r-pad, x+s-pad, c] * Wp[k][(r*3+s)*Ca + c]      (3x3, stride 1)
void k_conv_tc_rows(cudaStream_t st, const void* A, int B, int Ha, int Wa, int Ca, const void* Wp, int Nout, int pad, void* out,
                    int Ho, int Wo, const float* bias, double* stats, bool accumulate) {
    SALT_COUNT(1);
    RowsParams p;
    p.tiles_x = cdiv(Wo, RW_TW); p.tiles_y = cdiv(Ho, RW_TH);
    int BN = Nout % 256 == 0 ? 256 : Nout % 128 == 0 ? 128 : Nout % 64 == 0 ? 64 : 32;
    while (BN > 64 && (long long)p.tiles_x * p.tiles_y * B * (Nout / BN) < num_sms()) BN >>= 1;
    p.tiles_co = Nout / BN;
    p.total_tiles = p.tiles_x * p.tiles_y * B * p.tiles_co;
    p.B = B; p.Ho = Ho; p.Wo = Wo; p.Co = Nout; p.Ca = Ca; p.cblks = Ca / 64; p.pad = pad;
    p.accumulate = accumulate ? 1 : 0; p.bias = bias; p.stats = stats; p.out = (bf16*)out;
    CUtensorMap ma = make_map_nhwc(A, Ca, Wa, Ha, B, 64, RW_TW, RW_TH + 2, 1, 1, CU_TENSOR_MAP_SWIZZLE_128B);
    CUtensorMap mb;
    {
        cuuint64_t dims[2] = {(cuuint64_t)9 * Ca, (cuuint64_t)Nout};
        cuuint64_t strides[1] = {(cuuint64_t)9 * Ca * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)BN};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = get_encode()(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(Wp), dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled(weights) failed with code " + std::to_string((int)r));
    }
    if (BN == 256) launch_rows<256>(st, ma, mb, p);
    else if (BN == 128) launch_rows<128>(st, ma, mb, p);
    else if (BN == 64) launch_rows<64>(st, ma, mb, p);
    else launch_rows<32>(st, ma, mb, p);
}

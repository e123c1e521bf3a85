// tcgen05 implicit-GEMM convolution for sm_100a: forward, stride-1 dgrad and stride-2 dgrad (as 4 output phases).
//
//   out[n, y*osy+oy, x*osx+ox, k] (+)= sum_{tap} sum_c  A[n, y*as + dy(tap), x*as + dx(tap), c] * Wp[k][kofs(tap) + c]   (zero outside A)
//
// GEMM view: M = 128 output positions per tile (a tn x th x tw box), N = BN output channels, K = taps * C.
// One K-block = one (tap, 64-channel block): the A operand of that block is a TMA *box* load of the NHWC tensor at the
// tap-shifted coordinate (the TMA unit does the im2col; out-of-bounds rows are zero-filled = the zero padding; a stride-2
// convolution uses the tensor map's element strides), the B operand a [BN x 64] slab of the packed weights.  Both land in
// shared memory in the canonical K-major 128B-swizzled layout and feed tcgen05.mma (M128 x N BN x K16, bf16 in, fp32
// accumulators in TMEM).  The (dy, dx, kofs) tap table makes forward, dgrad (flipped taps, transposed weights) and the four
// parity phases of a stride-2 dgrad the same kernel.
//
// Persistent, warp-specialised CTA (1 per SM): warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2-5 =
// epilogue (tcgen05.ld -> +bias -> bf16 -> global, per-channel sum / sum^2 for BatchNorm via warp-shuffle butterflies).
// Two TMEM accumulators so the epilogue of tile i overlaps the MMAs of tile i+1; an S-stage smem ring decouples TMA from MMA.
#include "tc_common.cuh"
#include "conv_tc.h"
#include <cstdlib>
#include <algorithm>

using namespace tc;

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

struct TcTap { short dy, dx; int kofs; };
struct TcPhase { int ntaps, oy, ox; TcTap taps[9]; };
struct TcParams {
    FastDiv d_tpp, d_co, d_x, d_y;     // tiles_per_phase, tiles_co, tiles_x, tiles_y
    int tw, th, tn;
    int tiles_x, tiles_y, tiles_b, tiles_co, tiles_per_phase;
    int m_tiles;
    int B, Ho, Wo;                 // output positions per phase
    int Co;                        // output channels; physical output tensor: ep.Hp x ep.Wp
    int osy, osx, a_stride, cblks;
    int accumulate, nphases;
    int split_c;                   // fp32-output instantiations: channels of one operand term (K-blocks below it hold the h*h products)
    const float* bias;
    float* stats;              // [SALT_STAT_SLOTS_CONV][2*Co] partial slots, slot = blockIdx.x
    void* out;                 // OutT, physical [B][ep.Hp][ep.Wp][Co]
    EpiParams ep;
    TcPhase ph[4];
};

struct TcTile { int pi, nt, mt; bool live; };
__device__ __forceinline__ bool tc_tile(const TcParams& p, int k, TcTile& t) {
    const int tile = blockIdx.x + k * gridDim.x;
    if (tile >= p.nphases * p.tiles_per_phase) return false;
    t.pi = (int)p.d_tpp.div((uint32_t)tile);
    const int r = tile - t.pi * p.tiles_per_phase;
    t.mt = (int)p.d_co.div((uint32_t)r); t.nt = r - t.mt * p.tiles_co; t.live = true;
    return true;
}

constexpr int TC_THREADS = 192;
template <int BN, int BK> struct TcCfg {
    static constexpr int A_BYTES = 128 * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int BUDGET = 200 * 1024;
    static constexpr int STAGES_RAW = BUDGET / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
    static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
    static constexpr int TMEM_COLS_SPLIT = 4 * BN < 32 ? 32 : 4 * BN;      // fp32 output: main + correction accumulator per tile (conv_tc_rows.cu)
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 4 * 2 * BN * 4 /*per-epilogue-warp stats*/;
    // canonical K-major swizzled layout: rows of BK*2 bytes, 8-row groups SBO apart
    static constexpr uint32_t SBO = 8 * BK * 2;
    static constexpr uint32_t LAYOUT = BK == 64 ? 2 : 4;       // SWIZZLE_128B : SWIZZLE_64B
};

template <int BN, int BK, typename OutT>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const __grid_constant__ TcParams p) {
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
    constexpr bool SPLIT = sizeof(OutT) == 4;
    constexpr int TMEM_COLS = SPLIT ? Cfg::TMEM_COLS_SPLIT : Cfg::TMEM_COLS, ACC_STRIDE = SPLIT ? 2 * BN : BN;
    if (warp == 1) tmem_alloc(smem_u32(tmem_ptr_smem), TMEM_COLS);
    for (int i = threadIdx.x; i < 4 * 2 * BN; i += TC_THREADS) s_stats[i] = 0.f;
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        // ===================================================== TMA producer
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            TcTile tt;
            for (int k = 0; tc_tile(p, k, tt); ++k) {
                const TcPhase& ph = p.ph[tt.pi];
                const int nt = tt.nt, mt = tt.mt;
                const int trow = (int)p.d_x.div((uint32_t)mt), tx = mt - trow * p.tiles_x, tb = (int)p.d_y.div((uint32_t)trow), ty = trow - tb * p.tiles_y;
                const int w0 = tx * p.tw * p.a_stride, h0 = ty * p.th * p.a_stride, n0 = tb * p.tn;
                for (int tap = 0; tap < ph.ntaps; ++tap) {
                    const int dy = ph.taps[tap].dy, dx = ph.taps[tap].dx, kofs = ph.taps[tap].kofs;
                    for (int cb = 0; cb < p.cblks; ++cb) {
                        mbar_wait(empty0 + 8 * stage, phase ^ 1);
                        mbar_expect_tx(full0 + 8 * stage, Cfg::STAGE_BYTES);
                        tma_load_4d(smem_u32(smem_a + stage * Cfg::A_BYTES), &map_a, full0 + 8 * stage, cb * BK, w0 + dx, h0 + dy, n0);
                        tma_load_2d(smem_u32(smem_b + stage * Cfg::B_BYTES), &map_b, full0 + 8 * stage, kofs + cb * BK, nt * BN);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        const uint32_t idesc = instr_desc_bf16(BN, false, false);
        // one elected thread runs the whole loop and probes the next stage's barrier before the last MMA of the current stage
        // (tcgen05.mma issue is throttled to the execution rate; a barrier probe costs ~90 cycles: see conv_tc_rows.cu)
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            bool stage_ready = false, acc_ready = false;
            TcTile tt, tn;
            bool more = tc_tile(p, 0, tt);
            for (int kk = 0; more; ++kk) {
                more = tc_tile(p, kk + 1, tn);
                const int num_kb = p.ph[tt.pi].ntaps * p.cblks;
                if (!acc_ready) mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);
                const uint32_t tmem_d0 = tmem_base + acc * ACC_STRIDE;
                const int nacc = acc ^ 1;
                const uint32_t nacc_phase = acc_phase ^ (uint32_t)acc;
                uint32_t used = 0;
                for (int kb = 0; kb < num_kb; ++kb) {
                    const uint32_t sub = (SPLIT && (kb % p.cblks) * BK >= p.split_c) ? 1u : 0u;
                    const uint32_t tmem_d = tmem_d0 + sub * BN;
                    const uint32_t fresh = ((used >> sub) & 1u) ^ 1u;
                    used |= 1u << sub;
                    if (!stage_ready) mbar_wait(full0 + 8 * stage, phase);
                    fence_after();
                    const bool last = kb == num_kb - 1;
                    const int nstage = stage + 1 == STAGES ? 0 : stage + 1;
                    const uint32_t nphase = stage + 1 == STAGES ? phase ^ 1 : phase;
                    const uint64_t adesc = smem_desc(smem_u32(smem_a + stage * Cfg::A_BYTES), 16, Cfg::SBO, Cfg::LAYOUT);
                    const uint64_t bdesc = smem_desc(smem_u32(smem_b + stage * Cfg::B_BYTES), 16, Cfg::SBO, Cfg::LAYOUT);
                    uint32_t probe_stage = 0, probe_acc = 0;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {     // advance 16 bf16 = 32 bytes inside the swizzle atom
                        if (k == 1) {
                            if (!last || more) probe_stage = mbar_try_wait(full0 + 8 * nstage, nphase) ? 1u : 0u;
                            if (last && more) probe_acc = mbar_try_wait(tempty0 + 8 * nacc, nacc_phase ^ 1) ? 1u : 0u;
                        }
                        umma_bf16(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, fresh ? (uint32_t)k : 1u);
                    }
                    umma_commit(empty0 + 8 * stage);        // frees the smem stage when these MMAs retire
                    if (last) umma_commit(tfull0 + 8 * acc);
                    stage_ready = probe_stage != 0;
                    if (last) acc_ready = probe_acc != 0;
                    stage = nstage; phase = nphase;
                }
                acc = nacc; acc_phase = nacc_phase;
                tt = tn;
            }
        }
        __syncwarp();
    } else {
        // ===================================================== epilogue (warps 2..5 <-> TMEM lane quarters 2,3,0,1)
        const int quarter = warp & 3;
        float* my_stats = s_stats + (warp - 2) * (2 * BN);  // private per-warp accumulators, merged in warp order: no atomics, bit-reproducible
        float* slot = p.stats ? p.stats + (size_t)blockIdx.x * 2 * p.Co : nullptr;
        const int m = quarter * 32 + lane;                  // row of the 128-position tile
        const int lx = m % p.tw, ly = (m / p.tw) % p.th, ln = m / (p.tw * p.th);
        int acc = 0; uint32_t acc_phase = 0;
        TcTile tt;
        for (int kk = 0; tc_tile(p, kk, tt); ++kk) {
            const int pi = tt.pi, nt = tt.nt, mt = tt.mt;
            const int trow = (int)p.d_x.div((uint32_t)mt), tx = mt - trow * p.tiles_x, tb = (int)p.d_y.div((uint32_t)trow), ty = trow - tb * p.tiles_y;
            const int x = tx * p.tw + lx, y = ty * p.th + ly, n = tb * p.tn + ln;
            const bool valid = tt.live && (n < p.B) && (y < p.Ho) && (x < p.Wo);
            const int oy = y * p.osy + p.ph[pi].oy, ox = x * p.osx + p.ph[pi].ox;
            // physical output positions of this pixel: one, unless it sits on an edge of a replicate-bordered tensor (stride-1 only)
            const EpiParams& ep = p.ep;
            const bool bordered = p.osy == 1;
            const int py0 = (bordered && y == 0) ? 0 : oy + ep.pt, py1 = (bordered && y == p.Ho - 1) ? ep.Hp - 1 : oy + ep.pt;
            const int px0 = (bordered && x == 0) ? 0 : ox + ep.pl, px1 = (bordered && x == p.Wo - 1) ? ep.Wp - 1 : ox + ep.pl;
            OutT* obase = reinterpret_cast<OutT*>(p.out) + (size_t)n * ep.Hp * ep.Wp * p.Co + nt * BN;
            // accumulate mode (dgrad into an existing gradient) or the residual branch of an eval-mode fused block
            const OutT* rrow = nullptr;
            if (p.accumulate) rrow = obase + ((size_t)(oy + ep.pt) * ep.Wp + ox + ep.pl) * p.Co;
            else if (ep.res) rrow = reinterpret_cast<const OutT*>(ep.res) + (((size_t)n * p.Ho + y) * p.Wo + x) * p.Co + nt * BN;
            const bool has_r = rrow != nullptr && valid;
            mbar_wait(tfull0 + 8 * acc, acc_phase);
            fence_after();
#pragma unroll 1
            for (int ch = 0; ch < BN / 32; ++ch) {
                float v[32];
                uint4 old[4];
                if (sizeof(OutT) == 2 && has_r) {
                    const uint4* o4 = reinterpret_cast<const uint4*>(rrow + ch * 32);
#pragma unroll
                    for (int q = 0; q < 4; ++q) old[q] = o4[q];
                }
                tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * ACC_STRIDE + ch * 32, v);
                if constexpr (SPLIT) {
                    float v2[32];
                    tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * ACC_STRIDE + BN + ch * 32, v2);
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] += v2[i];
                }
                if (p.bias) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] += __ldg(p.bias + nt * BN + ch * 32 + i);
                }
                if (ep.scale) epi_affine32(v, ep.scale + nt * BN + ch * 32, ep.shift + nt * BN + ch * 32);
                if (has_r) {
                    if constexpr (sizeof(OutT) == 2) epi_add32(v, old);
                    else epi_add32(v, reinterpret_cast<const float*>(rrow) + ch * 32);
                }
                if (ep.relu) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
                }
                if (valid) epi_store32(v, obase + ch * 32, py0, py1, px0, px1, ep.Wp, p.Co);
                if (p.stats) {
                    float sq[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) { v[i] = valid ? v[i] : 0.f; sq[i] = v[i] * v[i]; }
                    float s1 = butterfly_reduce32(v, lane);
                    float s2 = butterfly_reduce32(sq, lane);
                    my_stats[ch * 32 + lane] += s1;
                    my_stats[BN + ch * 32 + lane] += s2;
                }
            }
            fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            if (p.stats && p.tiles_co > 1) {
                // the per-CTA accumulators are per output-channel tile: flush when the channel tile changes
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const int tt = threadIdx.x - 64;
                for (int i = tt; i < 2 * BN; i += 128) {
                    float val = 0.f;
#pragma unroll
                    for (int w4 = 0; w4 < 4; ++w4) { val += s_stats[w4 * 2 * BN + i]; s_stats[w4 * 2 * BN + i] = 0.f; }
                    slot[i < BN ? nt * BN + i : p.Co + nt * BN + (i - BN)] += val;      // this CTA's own slot: plain read-modify-write
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
        }
        if (p.stats && p.tiles_co == 1) {
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const int tt = threadIdx.x - 64;
            for (int i = tt; i < 2 * BN; i += 128) {
                float val = 0.f;
#pragma unroll
                for (int w4 = 0; w4 < 4; ++w4) val += s_stats[w4 * 2 * BN + i];
                slot[i < BN ? i : p.Co + (i - BN)] = val;
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 1) { fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

// ------------------------------------------------------------------------------------------------ host side
static CUtensorMap make_map_weights(const void* base, int K, int N, int boxK, int boxN, bool sw64) {
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)boxK, (cuuint32_t)boxN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, sw64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled(weights) failed with code " + std::to_string((int)r));
    return m;
}

bool tc_conv_supported(const ConvGeom& g, bool dgrad) {
    const int Cred = dgrad ? g.Co : g.Ci, Nout = dgrad ? g.Ci : g.Co;
    int Wd = dgrad ? g.Wi : g.Wo, Hd = dgrad ? g.Hi : g.Ho;
    if (g.stride != 1 && g.stride != 2) return false;
    if (dgrad && g.stride == 2) {
        if ((g.Hi | g.Wi) & 1) return false;
        Wd /= 2; Hd /= 2;
    }
    if (Cred % 32 || Nout % 32) return false;
    if (Wd < 8 || Hd < 8) return false;
    return true;
}

template <int BN, int BK, typename OutT>
static void launch_tc(cudaStream_t st, const CUtensorMap& ma, const CUtensorMap& mb, const TcParams& p) {
    using Cfg = TcCfg<BN, BK>;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(conv_tc_kernel<BN, BK, OutT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        configured = true;
    }
    const int total_tiles = p.nphases * p.tiles_per_phase;
    const int grid = total_tiles < num_sms() ? total_tiles : num_sms();
    conv_tc_kernel<BN, BK, OutT><<<grid, TC_THREADS, Cfg::SMEM_BYTES, st>>>(ma, mb, p);
}

// Common launcher.  A: [B,Ha,Wa,Ca] bf16; Wp: [Nout][Ktot] bf16; out: physical [B,out_H,out_W,Nout] bf16;
// per phase Ho x Wo output positions written at (y*os+oy, x*os+ox).
static void run_tc(cudaStream_t st, const void* A, int B, int Ha, int Wa, int Ca, const void* Wp, int Ktot, int Nout, TcParams& p, bool out_f32) {
    p.tw = p.Wo >= 16 ? 16 : 8;
    p.th = p.tw == 16 ? 8 : (p.Ho >= 16 ? 16 : 8);
    p.tn = 128 / (p.tw * p.th);
    p.tiles_x = cdiv(p.Wo, p.tw); p.tiles_y = cdiv(p.Ho, p.th); p.tiles_b = cdiv(B, p.tn);
    const int BK = (Ca % 64 == 0) ? 64 : 32;
    int BN = Nout % 256 == 0 ? 256 : Nout % 128 == 0 ? 128 : Nout % 64 == 0 ? 64 : 32;
    if (out_f32 && BN == 256) BN = 128;          // two accumulators per tile in the fp32-output mode: 4 x BN TMEM columns
    // prefer more, smaller channel tiles when the tile count would leave SMs idle
    while (BN > 64 && (long long)p.nphases * p.tiles_x * p.tiles_y * p.tiles_b * (Nout / BN) < num_sms()) BN >>= 1;
    p.tiles_co = Nout / BN;
    p.tiles_per_phase = p.tiles_x * p.tiles_y * p.tiles_b * p.tiles_co;
    p.d_tpp = make_fastdiv(p.tiles_per_phase); p.d_co = make_fastdiv(p.tiles_co); p.d_x = make_fastdiv(p.tiles_x); p.d_y = make_fastdiv(p.tiles_y);
    p.B = B; p.Co = Nout; p.cblks = Ca / BK;
    const bool sw64 = BK == 32;
    CUtensorMap ma = make_map_nhwc(A, Ca, Wa, Ha, B, BK, p.tw, p.th, p.tn, p.a_stride,
                                   sw64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
    p.m_tiles = p.tiles_x * p.tiles_y * p.tiles_b;
    CUtensorMap mb = make_map_weights(Wp, Ktot, Nout, BK, BN, sw64);
#define TC_CASE(bn, bk) if (BN == bn && BK == bk) {                                                 \
        if (out_f32) launch_tc<bn, bk, float>(st, ma, mb, p); else launch_tc<bn, bk, bf16>(st, ma, mb, p);  \
        return; }
    if (BN == 256 && BK == 64) { launch_tc<256, 64, bf16>(st, ma, mb, p); return; }
    if (BN == 256 && BK == 32) { launch_tc<256, 32, bf16>(st, ma, mb, p); return; }
    TC_CASE(128, 64) TC_CASE(64, 64) TC_CASE(32, 64)
    TC_CASE(128, 32) TC_CASE(64, 32) TC_CASE(32, 32)
#undef TC_CASE
    throw std::runtime_error("k_conv_tc: unsupported tile configuration");
}

static bool rows_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SALT_TC_ROWS"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}
// out[n,y,x,k] (+)= sum_{r,s,c} A[n, y*stride+r-pad, x*stride+s-pad, c] * Wp[k][(r*S+s)*Ca + c]
void k_conv_tc(cudaStream_t st, const void* A, int B, int Ha, int Wa, int Ca, const void* Wp, int Nout, int R, int S, int stride,
               int pad, void* out, int Ho, int Wo, const float* bias, float* stats, bool accumulate, bool out_f32, const EpiParams* ep,
               int split_c) {
    if (out_f32 && accumulate) throw std::runtime_error("k_conv_tc: accumulation into an fp32 output is not implemented");
    if (ep && accumulate) throw std::runtime_error("k_conv_tc: a fused epilogue cannot be combined with accumulation");
    if (rows_enabled() && tc_conv_rows_supported(Ca, Nout, R, S, stride, Ho, Wo)) {
        k_conv_tc_rows(st, A, B, Ha, Wa, Ca, Wp, Nout, pad, out, Ho, Wo, bias, stats, accumulate, out_f32, ep, split_c);
        return;
    }
    SALT_COUNT(1);
    if (out_f32 && (split_c <= 0 || split_c > Ca)) throw std::runtime_error("k_conv_tc: fp32 output needs the split-operand channel count");
    TcParams p;
    p.split_c = split_c;
    if (ep) p.ep = *ep;
    if (p.ep.Hp == 0) { p.ep.Hp = Ho; p.ep.Wp = Wo; p.ep.pt = p.ep.pl = 0; }
    p.Ho = Ho; p.Wo = Wo; p.osy = p.osx = 1; p.a_stride = stride;
    p.accumulate = accumulate ? 1 : 0; p.bias = bias; p.stats = stats; p.out = out;
    p.nphases = 1;
    TcPhase& ph = p.ph[0];
    ph.ntaps = R * S; ph.oy = ph.ox = 0;
    for (int r = 0; r < R; ++r)
        for (int s = 0; s < S; ++s) { TcTap& t = ph.taps[r * S + s]; t.dy = (short)(r - pad); t.dx = (short)(s - pad); t.kofs = (r * S + s) * Ca; }
    run_tc(st, A, B, Ha, Wa, Ca, Wp, R * S * Ca, Nout, p, out_f32);
}

// stride-2 dgrad: gin[n,yi,xi,c] (+)= sum over (r,s) with (yi+pad-r), (xi+pad-s) even of
//   gout[n,(yi+pad-r)/2,(xi+pad-s)/2,k] * W[k][c][r][s];   wpd = [Ci][(RS-1-t)*Co + k] (flipped-tap packing, kernels_conv_simt.cu).
// The four parities (yi&1, xi&1) are four small stride-1 correlations over gout, each writing a stride-2 lattice of gin.
void k_conv_tc_dgrad_s2(cudaStream_t st, const void* gout, int B, int Ho, int Wo, int Co, const void* wpd, int Ci, int R, int S,
                        int pad, void* gin, int Hi, int Wi, bool accumulate) {
    SALT_COUNT(1);
    TcParams p;
    p.Ho = Hi / 2; p.Wo = Wi / 2; p.ep.Hp = Hi; p.ep.Wp = Wi; p.osy = p.osx = 2; p.a_stride = 1; p.split_c = 0;
    p.accumulate = accumulate ? 1 : 0; p.bias = nullptr; p.stats = nullptr; p.out = gin;
    p.nphases = 0;
    for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px) {
            TcPhase ph; ph.ntaps = 0; ph.oy = py; ph.ox = px;
            for (int r = 0; r < R; ++r) {
                if ((py + pad - r) & 1) continue;
                for (int s = 0; s < S; ++s) {
                    if ((px + pad - s) & 1) continue;
                    TcTap& t = ph.taps[ph.ntaps++];
                    t.dy = (short)((py + pad - r) / 2); t.dx = (short)((px + pad - s) / 2);
                    t.kofs = (R * S - 1 - (r * S + s)) * Co;
                }
            }
            if (ph.ntaps == 0) {
                if (!accumulate) throw std::runtime_error("k_conv_tc_dgrad_s2: empty parity phase needs accumulate=true");
                continue;
            }
            p.ph[p.nphases++] = ph;
        }
    run_tc(st, gout, B, Ho, Wo, Co, wpd, R * S * Co, Ci, p, false);
}

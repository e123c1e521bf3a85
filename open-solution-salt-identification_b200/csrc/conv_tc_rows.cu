// tcgen05 3x3 stride-1 convolution with shared-memory ROW-HALO reuse of the activation tile (forward and stride-1 dgrad).
//
// Measurements that shaped this kernel (profiles/r1_notes.md): conv_tc.cu issues 4 MMAs per pipeline stage; with every load and
// store disabled it still ran at the same speed for N <= 128, i.e. it is bound by the single MMA-issuing thread's per-stage
// overhead (mbarrier try_wait ~90 cycles + commit), not by memory.  So here one stage carries THREE filter taps:
//   * the output tile is 16 rows x 8 columns; per 64-channel block and horizontal tap offset s one TMA box of 18 rows x 8 pixels
//     (16 output rows + vertical halo) is loaded.  A pixel row of 8 pixels x 64 channels is exactly one 1024-byte swizzle atom,
//     so the vertical taps r = 0,1,2 are the SAME buffer read at byte offsets r*1024 - the UMMA descriptor start address stays
//     1024-byte aligned and the canonical K-major 128B-swizzled layout is untouched (A traffic: 3 x 18 KB instead of 9 x 16 KB);
//   * the three weight slabs [BN x 64] of taps (0..2, s) arrive in the same stage (one 4-D TMA box (c, n, r, s), or CL multicast parts);
//   * the issuer waits once and issues 12 MMAs (3 taps x 4 K-steps) per stage, then one tcgen05.commit frees the stage.
// Warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-9 = epilogue (two per TMEM lane quarter), two TMEM accumulators, persistent CTAs.
//
//   conv_tc_rows_kernel<BN, OutT, 1>   one CTA per tile
//   conv_tc_rows_kernel<BN, OutT, 2>   2-CTA thread-block cluster, weight slabs TMA-multicast (default; SALT_TC_CLUSTER=1 turns it off)
// OutT = bf16 (the bf16 precision mode) or float (the fp32 tensor-core parity mode: split-bf16 operands, see k_split6).
// Measured bound (profiles/r1_notes.md, profiles/r2_notes.md): tcgen05.mma fetches its shared-memory operands at ~64 B/clk and
// re-fetches the 4 KB A slice for every instruction -> ~(4096 + 32 N)/64 clk per MMA; L2/HBM traffic and the bytes entering the SM
// are not the limit.  (Round 1 also carried a cta_group::2 pair variant with a forwarded barrier and a multi-sub-tile variant;
// both measured no faster and were removed - git history has them.)
#include "tc_common.cuh"
#include "conv_tc.h"
#include <cstdlib>
#include <algorithm>

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

// ---- optional pipeline-stall accounting (build.py --timing, -DSALT_TC_TIMING; never in the shipped library): per CTA, the cycles
// the three roles of conv_tc_rows_kernel spend waiting on each other.  Read back with salt_debug_rows_timing() (profiles/rows_timing.py).
#ifdef SALT_TC_TIMING
__device__ unsigned long long g_rows_timing[SALT_STAT_SLOTS_CONV][8];
#define TCT(...) __VA_ARGS__
extern "C" int salt_debug_rows_timing(unsigned long long* out, int reset) {
    cudaDeviceSynchronize();
    if (out) cudaMemcpyFromSymbol(out, g_rows_timing, sizeof(g_rows_timing));
    if (reset) { static unsigned long long z[SALT_STAT_SLOTS_CONV][8]; cudaMemcpyToSymbol(g_rows_timing, z, sizeof(z)); }
    return 0;
}
#else
#define TCT(...)
#endif

struct RowsParams {
    FastDiv d_co, d_x, d_y;    // tiles_co, tiles_x, tiles_y
    int tiles_x, tiles_y, tiles_co, total_tiles;
    int m_tiles, total_groups;  // cluster mode: a group = CL consecutive pixel tiles of one channel tile (one per CTA of the cluster)
    int B, Ho, Wo, Co, Ca, cblks, pad;
    int accumulate;
    int ksteps;                // K = 16 steps per tap and 64-channel block that carry data: 4, or 2 for a 32-channel input (the TMA
                               // box is still 64 channels wide - channels 32..63 are out of bounds and zero-filled - so the layout,
                               // the descriptors and the byte counts are unchanged; the two all-zero K steps are simply not issued)
    int split_c;               // fp32-output instantiations: channels of ONE operand term (stages with cb*64 < split_c hold the h*h products)
    const float* bias;
    float* stats;              // [SALT_STAT_SLOTS_CONV][2*Co] partial slots, slot = blockIdx.x
    void* out;                 // OutT, physical [B][ep.Hp][ep.Wp][Co]
    EpiParams ep;
};

constexpr int RW_THREADS = 320;                              // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue
constexpr int RW_EPI_THREADS = 256;
constexpr int RW_TH = 16, RW_TW = 8;
constexpr int RW_A_BYTES = (RW_TH + 2) * RW_TW * 128;        // 18 pixel rows x 8 pixels x 64 channels bf16 = 18 KB
// RESW (resident weights): layers with <= 64 input channels and ONE output-channel tile keep all nine weight slabs in shared memory
// for the life of the persistent CTA (72 KB at BN = 64); a stage is then the activation box alone.  With the weights in the ring such
// a layer pulled 42 KB per stage into the SM for 12 MMAs (6 for 32-channel inputs) = 55-109 B/clk, above what L2 and the shared-memory
// fill path deliver (profiles/r2_notes.md section 7).
template <int BN, bool RESW = false> struct RowsCfg {
    static constexpr int B_BYTES = BN * 128;                 // one tap: [BN][64] bf16
    static constexpr int W_BYTES = RESW ? 9 * B_BYTES : 0;   // resident weight slabs, in front of the ring
    static constexpr int STAGE_BYTES = RESW ? RW_A_BYTES : RW_A_BYTES + 3 * B_BYTES;
    static constexpr int STAGES = RESW ? 6 : (BN == 128 ? 3 : (BN == 64 ? 4 : (BN == 32 ? 6 : 2)));    // wide tiles (160 / 192): 78 / 90 KB stages
    // column distance of the two accumulators: BN for the power-of-two tiles, 256 for the wide ones (tcgen05.alloc wants 2^k columns)
    static constexpr int ACC_STRIDE = (BN & (BN - 1)) == 0 ? BN : 256;
    static constexpr int TMEM_COLS = 2 * ACC_STRIDE < 32 ? 32 : 2 * ACC_STRIDE;
    // fp32-output (split-bf16) instantiations keep TWO accumulators per tile: the h*h products and the ~2^-8 smaller correction
    // products.  tcgen05 adds into an fp32 accumulator with truncation at the accumulator's magnitude; keeping the small terms apart
    // makes their share of that error negligible (measured: 3e-5 -> relative error per convolution, see profiles/r2_notes.md)
    static constexpr int ACC_STRIDE_SPLIT = 2 * BN;
    static constexpr int TMEM_COLS_SPLIT = 2 * ACC_STRIDE_SPLIT < 32 ? 32 : 2 * ACC_STRIDE_SPLIT;
    static constexpr int BAR_OFF = W_BYTES + STAGES * STAGE_BYTES;
    static constexpr int SMEM_BYTES = BAR_OFF + 1024 + 256 + 8 * 2 * BN * 4 + BN * 4;      // + per-epilogue-warp BN statistics + the bias
};

// ---- thread-block-cluster helpers (weight multicast)
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA box delivered to the same shared-memory offset of every CTA in `mask`, completing bytes on the mbarrier at the same offset
__device__ __forceinline__ void tma_load_4d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3,
                                               uint16_t mask) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
                 " [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "h"(mask) : "memory");
}
// tcgen05.commit arriving on the mbarrier at the same offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}
// ---- CTA-pair (cta_group::2) helpers: one tcgen05.mma spans the tensor cores of both SMs of the pair (M = 256)
// arrive on the mbarrier at the same shared-memory offset in CTA `cta` of the cluster (release at cluster scope)
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(bar), "r"(cta));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// wait on a local mbarrier whose arrivals come from other CTAs of the cluster (acquire at cluster scope); bounded like mbar_wait
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait_cluster(bar, parity)) {
        if (++spins > (1u << 24)) {
            printf("tc: cluster mbarrier wait timed out (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
            __trap();
        }
    }
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}
// TMA box load issued by either CTA of a pair into ITS OWN shared memory, completing bytes on the LEADER's mbarrier (the barrier
// address with the CTA-rank bit of the shared::cluster window cleared - the cta_group::2 form allows the barrier to live in the peer)
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ uint32_t instr_desc_bf16_m256(int n) {
    uint32_t d = 0;
    d |= 1u << 4; d |= 1u << 7; d |= 1u << 10;
    d |= (uint32_t)(n >> 3) << 17;
    d |= (uint32_t)(256 >> 4) << 24;
    return d;
}

// Work item k of this CTA.  CL == 1: tile = blockIdx.x + k*gridDim.x, channel tile fastest.  CL > 1: the CL CTAs of a cluster take
// CL consecutive pixel tiles of ONE channel tile, so they consume identical weight stages in lockstep; a cluster whose last group
// is short gives the surplus CTAs a clamped tile whose results are dropped (`live` = false).
template <int CL>
struct RowsTile { int nt, tx, ty, n; bool live; };
template <int CL>
__device__ __forceinline__ bool rows_tile(const RowsParams& p, int k, uint32_t rank, RowsTile<CL>& t) {
    int mt;
    if constexpr (CL == 1) {
        const int tile = blockIdx.x + k * gridDim.x;
        if (tile >= p.total_tiles) return false;
        mt = (int)p.d_co.div((uint32_t)tile); t.nt = tile - mt * p.tiles_co; t.live = true;
    } else {
        const int g = (int)(blockIdx.x / CL) + k * (int)(gridDim.x / CL);
        if (g >= p.total_groups) return false;
        const int q = (int)p.d_co.div((uint32_t)g);
        t.nt = g - q * p.tiles_co; mt = q * CL + (int)rank;
        t.live = mt < p.m_tiles;
        if (!t.live) mt = p.m_tiles - 1;
    }
    const int row = (int)p.d_x.div((uint32_t)mt);          // (image, tile row)
    t.tx = mt - row * p.tiles_x;
    t.n = (int)p.d_y.div((uint32_t)row);
    t.ty = row - t.n * p.tiles_y;
    return true;
}

// accumulator buffers per CTA: 2, or 4 for the narrow (<= 64 column) bf16 tiles - their tiles are short (36 MMAs for K = 576), so the
// MMA -> epilogue -> MMA hand-off latencies were exposed with only two buffers (profiles/r2_notes.md)
template <int BN, typename OutT, bool PAIR> struct RowsAcc { static constexpr int N = (!PAIR && sizeof(OutT) == 2 && BN <= 64) ? 4 : 2; };
template <int BN, typename OutT, bool NARROW, int CL, bool PAIR = false>
__device__ __forceinline__ void rows_epilogue(const RowsParams& p, const int warp, const int lane, const uint32_t tmem_base,
                                      const uint32_t tfull0, const uint32_t tempty0, float* s_stats, const uint32_t rank) {
    // Epilogue, 8 warps (2..9).  Warp w may touch TMEM lanes 32*(w%4)..+31;
    // the two warps of a lane quarter split the 32-column chunks between them, so every scheduler has two epilogue warps
    // to interleave (the shuffle/convert chains of a single warp left the issue slots idle - profiles/r1_notes.md).
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    float* my_stats = s_stats + (warp - 2) * (2 * BN);  // private per-warp accumulators: no shared-memory atomics
    const int m = quarter * 32 + lane;                  // row of the 128-position tile: 16 rows x 8 columns
    const int lx = m & (RW_TW - 1), ly = m >> 3;
    int acc = 0; uint32_t acc_phase = 0;
    // Narrow layers (BN <= 64, one channel tile): every epilogue warp owns ONE fixed 32-column chunk for the whole kernel, so
    // its bias lives in registers and the BatchNorm sums are accumulated per THREAD (one pixel row each) across all tiles of
    // the CTA; the cross-lane butterfly runs once per kernel instead of once per tile.  For K = 576 layers the per-tile
    // butterflies (62 shuffles per chunk) made the epilogue longer than the MMA phase (profiles/r1_notes.md).
    const bool own_chunk = half < BN / 32;
    float rs1[32], rs2[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) { rs1[i] = 0.f; rs2[i] = 0.f; }
    // the bias of the narrow layers sits in shared memory behind the statistics (32 more live registers spilled the sums above)
    const float4* s_bias4 = reinterpret_cast<const float4*>(s_stats + 8 * 2 * BN + (own_chunk ? half * 32 : 0));
    RowsTile<CL> t;
    for (int k = 0; rows_tile<CL>(p, k, rank, t); ++k) {
        const int nt = t.nt, n = t.n;
        const int x = t.tx * RW_TW + lx, y = t.ty * RW_TH + ly;
        const bool valid = t.live && (y < p.Ho) && (x < p.Wo);
        // physical output positions of this pixel (one, unless it sits on an edge of a replicate-bordered tensor)
        const EpiParams& ep = p.ep;
        const int py0 = (y == 0) ? 0 : y + ep.pt, py1 = (y == p.Ho - 1) ? ep.Hp - 1 : y + ep.pt;
        const int px0 = (x == 0) ? 0 : x + ep.pl, px1 = (x == p.Wo - 1) ? ep.Wp - 1 : x + ep.pl;
        OutT* obase = reinterpret_cast<OutT*>(p.out) + (size_t)n * ep.Hp * ep.Wp * p.Co + nt * BN;
        // values added to the tile before it is stored: the old gradient (accumulate mode: dgrad into an existing gradient, an
        // unbordered tensor) or the residual branch (eval-mode fused BasicBlock).  For bf16 the first chunk is fetched BEFORE
        // waiting for the accumulator, so its DRAM latency hides behind the MMA phase instead of serialising the epilogue.
        const OutT* rrow = nullptr;
        if constexpr (!NARROW) {
            if (p.accumulate) rrow = obase + ((size_t)(y + ep.pt) * ep.Wp + x + ep.pl) * p.Co;
            else if (ep.res) rrow = reinterpret_cast<const OutT*>(ep.res) + (((size_t)n * p.Ho + y) * p.Wo + x) * p.Co + nt * BN;
        }
        const bool has_r = rrow != nullptr && valid;
        uint4 old[4];
        if (sizeof(OutT) == 2 && has_r && half < BN / 32) {
            const uint4* o4 = reinterpret_cast<const uint4*>(rrow + half * 32);
#pragma unroll
            for (int q = 0; q < 4; ++q) old[q] = o4[q];
        }
        TCT(long long te0 = clock64();)
        mbar_wait(tfull0 + 8 * acc, acc_phase);
        TCT(if (warp == 2 && lane == 0) g_rows_timing[blockIdx.x][6] += (unsigned long long)(clock64() - te0);)
        fence_after();
        constexpr bool SPLIT = sizeof(OutT) == 4;
        constexpr int ACC_STRIDE = SPLIT ? RowsCfg<BN>::ACC_STRIDE_SPLIT : (PAIR ? BN : RowsCfg<BN>::ACC_STRIDE);
        // tiles with several 32-column chunks per warp (BN >= 128, wide dgrad tiles): the TMEM load of chunk c+1 is in flight while
        // chunk c is converted and stored (the wide dgrads were epilogue-bound at 3x the MMA time, profiles/r2_notes.md)
        constexpr bool PREFETCH = !NARROW && !SPLIT && BN > 64;
        const uint32_t tacc = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * ACC_STRIDE;
        float vnext[PREFETCH ? 32 : 1];
        if constexpr (PREFETCH) { if (half < BN / 32) tmem_ld32_issue(tacc + half * 32, vnext); }
#pragma unroll 1
        for (int ch = half; ch < BN / 32; ch += 2) {
            float v[32];
            if (sizeof(OutT) == 2 && has_r && ch != half) {
                const uint4* o4 = reinterpret_cast<const uint4*>(rrow + ch * 32);
#pragma unroll
                for (int q = 0; q < 4; ++q) old[q] = o4[q];
            }
            if constexpr (PREFETCH) {
                tmem_ld_wait32(vnext);
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = vnext[i];
                if (ch + 2 < BN / 32) tmem_ld32_issue(tacc + (ch + 2) * 32, vnext);
            } else {
                tmem_ld32(tacc + ch * 32, v);
            }
            if constexpr (SPLIT) {          // + the accumulator of the correction products
                float v2[32];
                tmem_ld32(tacc + BN + ch * 32, v2);
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] += v2[i];
            }
            if constexpr (NARROW) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 b4 = s_bias4[i];
                    v[4 * i] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
                }
            } else {
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
            }
            if (valid) epi_store32(v, obase + ch * 32, py0, py1, px0, px1, ep.Wp, p.Co);
            if (p.stats) {
                if constexpr (NARROW) {
                    if (valid) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) { rs1[i] += v[i]; rs2[i] = fmaf(v[i], v[i], rs2[i]); }
                    }
                } else {
                    float sq[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) { v[i] = valid ? v[i] : 0.f; sq[i] = v[i] * v[i]; }
                    float s1 = rows_butterfly_reduce32(v, lane);
                    float s2 = rows_butterfly_reduce32(sq, lane);
                    my_stats[ch * 32 + lane] += s1;
                    my_stats[BN + ch * 32 + lane] += s2;
                }
            }
        }
        fence_before();
        __syncwarp();
        if (lane == 0) {
            if (PAIR && rank != 0) mbar_arrive_remote(tempty0 + 8 * acc, 0);       // the accumulator barrier lives in the leader CTA
            else mbar_arrive(tempty0 + 8 * acc);
        }
        if (++acc == RowsAcc<BN, OutT, PAIR>::N) { acc = 0; acc_phase ^= 1; }
        if (p.stats && p.tiles_co > 1) {
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const int tt = threadIdx.x - 64;
            for (int i = tt; i < 2 * BN; i += RW_EPI_THREADS) {
                float val = 0.f;
#pragma unroll
                for (int w8 = 0; w8 < 8; ++w8) { val += s_stats[w8 * 2 * BN + i]; s_stats[w8 * 2 * BN + i] = 0.f; }
                p.stats[(size_t)blockIdx.x * 2 * p.Co + (i < BN ? nt * BN + i : p.Co + nt * BN + (i - BN))] += val;   // own slot
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
        }
    }
    if (NARROW && p.stats && own_chunk) {
        const float s1 = rows_butterfly_reduce32(rs1, lane);
        const float s2 = rows_butterfly_reduce32(rs2, lane);
        my_stats[half * 32 + lane] += s1;
        my_stats[BN + half * 32 + lane] += s2;
    }
    if (p.stats && p.tiles_co == 1) {
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const int tt = threadIdx.x - 64;
        for (int i = tt; i < 2 * BN; i += RW_EPI_THREADS) {
            float val = 0.f;
#pragma unroll
            for (int w8 = 0; w8 < 8; ++w8) val += s_stats[w8 * 2 * BN + i];
            p.stats[(size_t)blockIdx.x * 2 * p.Co + (i < BN ? i : p.Co + (i - BN))] = val;                                      // own slot
        }
    }
}

// CL = thread-block-cluster size.  CL > 1: the CL CTAs of a cluster work on CL different pixel tiles of the same channel tile and
// share every weight stage: each CTA fetches 1/CL of the three weight slabs and TMA-multicasts it into the shared memory of all
// CL CTAs (map_b then has a [64 c][BN/CL n] box).  Weights are 60-75 % of the bytes a stage pulls through L2, and the L2 -> SM
// path (~42 B/clk/SM chip-wide), not the tensor pipe, bounds these kernels (profiles/r1_notes.md).  A stage may be overwritten only
// when ALL CTAs of the cluster have consumed it: the MMA issuer's tcgen05.commit arrives on the `empty` barrier of every CTA.
template <int BN, typename OutT, int CL, bool RESW = false>
__global__ void __launch_bounds__(RW_THREADS, 1)
conv_tc_rows_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const RowsParams p) {
    static_assert(!RESW || CL == 1, "resident weights: one CTA per tile");
    using Cfg = RowsCfg<BN, RESW>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
    // bars: full[STAGES], empty[STAGES], tmem_full[4], tmem_empty[4], weights_full
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 9);
    float* s_stats = reinterpret_cast<float*>(smem + Cfg::BAR_OFF + 256);
    constexpr int NACC = RowsAcc<BN, OutT, false>::N;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t wsm0 = smem_u32(smem);                    // resident weight slabs (RESW), then the ring
    const uint32_t smem0 = wsm0 + Cfg::W_BYTES;
    const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * STAGES, tfull0 = empty0 + 8 * STAGES, tempty0 = tfull0 + 32, wfull = tempty0 + 32;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        for (int i = 0; i < STAGES; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, CL); }
        for (int i = 0; i < NACC; ++i) { mbar_init(tfull0 + 8 * i, 1); mbar_init(tempty0 + 8 * i, 8); }
        mbar_init(wfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    constexpr bool SPLIT = sizeof(OutT) == 4;
    constexpr int ACC_STRIDE = SPLIT ? Cfg::ACC_STRIDE_SPLIT : Cfg::ACC_STRIDE;
    constexpr int TMEM_COLS = NACC * ACC_STRIDE < 32 ? 32 : NACC * ACC_STRIDE;
    static_assert(TMEM_COLS <= 512, "accumulators do not fit TMEM");
    if (warp == 1) tmem_alloc(smem_u32(tmem_ptr_smem), TMEM_COLS);
    for (int i = threadIdx.x; i < 8 * 2 * BN; i += RW_THREADS) s_stats[i] = 0.f;
    for (int i = threadIdx.x; i < BN; i += RW_THREADS) s_stats[8 * 2 * BN + i] = (p.bias && p.tiles_co == 1) ? __ldg(p.bias + i) : 0.f;
    fence_before();
    __syncthreads();
    uint32_t rank = 0;
    if constexpr (CL > 1) { rank = cluster_ctarank(); cluster_sync_all(); }      // every CTA's barriers exist before any multicast
    fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const int stages_per_tile = p.cblks * 3;
    constexpr uint16_t MC_MASK = (uint16_t)((1u << CL) - 1);

    if (warp == 0) {
        // ===================================================== TMA producer: one A box + one 3-tap weight box per stage
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            RowsTile<CL> t;
            TCT(long long tp0 = clock64(); long long tp_wait = 0;)
            if constexpr (RESW) {                        // all nine slabs once: box s = [3 taps r][BN][64 c] of kernel column s
                mbar_expect_tx(wfull, Cfg::W_BYTES);
#pragma unroll
                for (int s = 0; s < 3; ++s) tma_load_4d(wsm0 + s * 3 * Cfg::B_BYTES, &map_b, wfull, 0, 0, 0, s);
            }
            for (int k = 0; rows_tile<CL>(p, k, rank, t); ++k) {
                const int nt = t.nt, n = t.n;
                const int w0 = t.tx * RW_TW - p.pad, h0 = t.ty * RW_TH - p.pad;
                for (int cb = 0; cb < p.cblks; ++cb) {
                    for (int s = 0; s < 3; ++s) {
                        const uint32_t st = smem0 + stage * Cfg::STAGE_BYTES, fb = full0 + 8 * stage;
                        TCT(long long tw = clock64();)
                        mbar_wait(empty0 + 8 * stage, phase ^ 1);
                        TCT(tp_wait += clock64() - tw;)
                        {
                            mbar_expect_tx(fb, Cfg::STAGE_BYTES);
                            tma_load_4d(st, &map_a, fb, cb * 64, w0 + s, h0, n);
                            if constexpr (RESW) {
                                // the weights are resident
                            } else if constexpr (CL == 1) {
                                tma_load_4d(st + RW_A_BYTES, &map_b, fb, cb * 64, nt * BN, 0, s);     // (c, n, r = 0..2, s)
                            } else {
                                constexpr int PART = BN / CL;                                         // my rows of every weight slab
#pragma unroll
                                for (int r = 0; r < 3; ++r)
                                    tma_load_4d_mc(st + RW_A_BYTES + r * Cfg::B_BYTES + rank * (PART * 128), &map_b, fb, cb * 64,
                                                   nt * BN + (int)rank * PART, r, s, MC_MASK);
                            }
                        }
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
            TCT(g_rows_timing[blockIdx.x][4] += (unsigned long long)tp_wait; g_rows_timing[blockIdx.x][5] += (unsigned long long)(clock64() - tp0);)
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer: 12 MMAs per barrier wait
        const uint32_t idesc = instr_desc_bf16(BN, false, false);
        const uint64_t adesc0 = smem_desc(smem0, 16, 1024, 2);
        const uint64_t bdesc0 = smem_desc(RESW ? wsm0 : smem0 + RW_A_BYTES, 16, 1024, 2);
        // ONE elected thread runs the whole issue loop.  tcgen05.mma issue is throttled to the execution rate (the hardware
        // queue is shallow: profiles/r2_notes.md, rows_timing), so every cycle this thread spends between two MMAs idles the tensor
        // pipe - and a barrier probe has ~90 cycles of latency even when the barrier completed long ago.  The probe of the NEXT
        // stage's `full` barrier (and, on a tile's last stage, of the next accumulator's `tmem_empty` barrier) is therefore issued
        // BEFORE the last two MMAs of the current stage and consumed after them: its latency hides behind their execution.
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            bool stage_ready = false, acc_ready = false;
            RowsTile<CL> t, tn;
            TCT(long long ti0 = clock64(); long long ti_full = 0, ti_tempty = 0, ti_stages = 0;)
            if constexpr (RESW) mbar_wait(wfull, 0);
            bool more = rows_tile<CL>(p, 0, rank, t);
            for (int kk = 0; more; ++kk) {
                more = rows_tile<CL>(p, kk + 1, rank, tn);
                TCT(long long tw0 = clock64();)
                if (!acc_ready) mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);
                TCT(ti_tempty += clock64() - tw0;)
                const uint32_t tmem_d0 = tmem_base + acc * ACC_STRIDE;
                const int nacc = acc + 1 == NACC ? 0 : acc + 1;
                const uint32_t nacc_phase = acc + 1 == NACC ? acc_phase ^ 1u : acc_phase;      // the phase flips when acc wraps
                uint32_t used = 0;          // bit 0 / 1: the main / correction accumulator of this tile has been written
                for (int it = 0; it < stages_per_tile; ++it) {
                    const uint32_t sub = (SPLIT && (it / 3) * 64 >= p.split_c) ? 1u : 0u;
                    const uint32_t tmem_d = tmem_d0 + sub * BN;
                    const uint32_t fresh = ((used >> sub) & 1u) ^ 1u;
                    used |= 1u << sub;
                    TCT(long long tw1 = clock64();)
                    if (!stage_ready) mbar_wait(full0 + 8 * stage, phase);
                    TCT(ti_full += clock64() - tw1; ++ti_stages;)
                    fence_after();
                    const bool last = it == stages_per_tile - 1;
                    const int nstage = stage + 1 == STAGES ? 0 : stage + 1;
                    const uint32_t nphase = stage + 1 == STAGES ? phase ^ 1 : phase;
                    const uint64_t soff = (uint64_t)((stage * Cfg::STAGE_BYTES) >> 4);
                    uint32_t probe_stage = 0, probe_acc = 0;
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        // vertical tap r = the same A buffer, r pixel-rows (r * 1024 bytes) further down
                        const uint64_t adesc = adesc0 + soff + (uint64_t)((r * 1024) >> 4);
                        // resident weights: one 64-channel block, so stage `it` of a tile is kernel column s = it
                        const uint64_t bdesc = RESW ? bdesc0 + (uint64_t)(((it * 3 + r) * Cfg::B_BYTES) >> 4)
                                                    : bdesc0 + soff + (uint64_t)((r * Cfg::B_BYTES) >> 4);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (k >= p.ksteps) break;
                            if (r == 2 && k == p.ksteps - 2) {      // all but the last two issued: probe what the next iteration will need
                                if (!last || more) probe_stage = mbar_try_wait(full0 + 8 * nstage, nphase) ? 1u : 0u;
                                if (last && more) probe_acc = mbar_try_wait(tempty0 + 8 * nacc, nacc_phase ^ 1) ? 1u : 0u;
                            }
                            umma_bf16(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, fresh ? (uint32_t)(r | k) : 1u);
                        }
                    }
                    if constexpr (CL == 1) umma_commit(empty0 + 8 * stage);
                    else umma_commit_mc(empty0 + 8 * stage, MC_MASK);
                    if (last) umma_commit(tfull0 + 8 * acc);
                    stage_ready = probe_stage != 0;
                    if (last) acc_ready = probe_acc != 0;
                    stage = nstage; phase = nphase;
                }
                acc = nacc; acc_phase = nacc_phase;
                t = tn;
            }
            TCT(g_rows_timing[blockIdx.x][0] += (unsigned long long)(clock64() - ti0); g_rows_timing[blockIdx.x][1] += (unsigned long long)ti_full;
                g_rows_timing[blockIdx.x][2] += (unsigned long long)ti_tempty; g_rows_timing[blockIdx.x][3] += (unsigned long long)ti_stages;)
        }
        __syncwarp();
    } else {
        // ===================================================== epilogue: 8 warps (2..9)
        bool narrow = false;
        if constexpr (BN <= 64) narrow = p.tiles_co == 1 && p.stats != nullptr && !p.accumulate;
        if constexpr (BN <= 64) {
            if (narrow) rows_epilogue<BN, OutT, true, CL>(p, warp, lane, tmem_base, tfull0, tempty0, s_stats, rank);
        }
        if (!narrow) rows_epilogue<BN, OutT, false, CL>(p, warp, lane, tmem_base, tfull0, tempty0, s_stats, rank);
    }
    fence_before();
    __syncthreads();
    if constexpr (CL > 1) cluster_sync_all();      // no CTA leaves while a peer's commit may still arrive on its barriers
    if (warp == 1) { fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}


// ================================================================================================
// CTA-pair variant (cta_group::2): the two CTAs of a cluster work on two pixel tiles of the same channel tile; ONE tcgen05.mma
// (M = 256, N = BN) issued by the leader covers both.  Each CTA stages its own activation box and only HALF of every weight slab
// (B is split in N across the pair), so the bytes a stage pulls into an SM drop from 18 KB + 3*BN*128 to 18 KB + 1.5*BN*128 -
// that ingest (~47 B/clk/SM measured, profiles/r2_notes.md) is what bounds the single-CTA kernel for BN >= 128 - and BN = 256
// becomes possible (3 stages of 66 KB).  Measured operand-feed bound of the pair MMA: max(58, BN/2) clk (mma_feed micro-benchmark).
// Protocol (round 1's pair kernel forwarded every stage through a second barrier and was 1.45x slower; here the hardware does it):
//   full[s]    leader only, 1 arrival (the leader's producer, expect_tx = the bytes of BOTH CTAs); both producers' TMA loads
//              complete on it directly (cp.async.bulk.tensor ... .cta_group::2 with the leader's barrier address)
//   empty[s]   each CTA, 1 arrival: the leader's tcgen05.commit.cta_group::2, multicast to both CTAs
//   tfull[a]   each CTA, 1 arrival: the same multicast commit after a tile's last MMA; each CTA drains its own 128 TMEM lanes
//   tempty[a]  leader only, 16 arrivals: the 8 epilogue warps of each CTA (remote arrive from the peer)
template <int BN> struct PairCfg {
    static constexpr int B_HALF = (BN / 2) * 128;                        // one tap, my half of the channels: [BN/2][64] bf16
    static constexpr int STAGE_BYTES = RW_A_BYTES + 3 * B_HALF;
    static constexpr int STAGES = BN == 256 ? 3 : (BN == 128 ? 4 : 6);
    static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
    static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
    static constexpr int SMEM_BYTES = BAR_OFF + 1024 + 256 + 8 * 2 * BN * 4 + BN * 4;
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
    static_assert(TMEM_COLS <= 512, "accumulators do not fit TMEM");
};
template <int BN>
__global__ void __launch_bounds__(RW_THREADS, 1)
conv_tc_rows_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const RowsParams p) {
    using Cfg = PairCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
    // bars: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
    float* s_stats = reinterpret_cast<float*>(smem + Cfg::BAR_OFF + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem0 = smem_u32(smem);
    const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * STAGES, tfull0 = empty0 + 8 * STAGES, tempty0 = tfull0 + 16;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        for (int i = 0; i < STAGES; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(tfull0 + 8 * i, 1); mbar_init(tempty0 + 8 * i, 16); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc_pair(smem_u32(tmem_ptr_smem), Cfg::TMEM_COLS);
    for (int i = threadIdx.x; i < 8 * 2 * BN; i += RW_THREADS) s_stats[i] = 0.f;
    for (int i = threadIdx.x; i < BN; i += RW_THREADS) s_stats[8 * 2 * BN + i] = (p.bias && p.tiles_co == 1) ? __ldg(p.bias + i) : 0.f;
    fence_before();
    __syncthreads();
    cluster_sync_all();                 // both CTAs' barriers and TMEM exist before any cross-CTA traffic
    fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const int stages_per_tile = p.cblks * 3;

    if (warp == 0) {
        // ===================================================== TMA producer (both CTAs): own activation box + my half of the weights
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            RowsTile<2> t;
            for (int k = 0; rows_tile<2>(p, k, rank, t); ++k) {
                const int w0 = t.tx * RW_TW - p.pad, h0 = t.ty * RW_TH - p.pad;
                for (int cb = 0; cb < p.cblks; ++cb) {
                    for (int s = 0; s < 3; ++s) {
                        const uint32_t st = smem0 + stage * Cfg::STAGE_BYTES, fb = full0 + 8 * stage;
                        mbar_wait_cluster(empty0 + 8 * stage, phase ^ 1);         // released by the leader's multicast commit
                        if (leader) mbar_expect_tx(fb, 2 * Cfg::STAGE_BYTES);
                        tma_load_4d_pair(st, &map_a, fb, cb * 64, w0 + s, h0, t.n);
                        tma_load_4d_pair(st + RW_A_BYTES, &map_b, fb, cb * 64, t.nt * BN + (int)rank * (BN / 2), 0, s);   // (c, n half, r = 0..2, s)
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== leader: MMA issuer for the pair (one elected thread, early barrier probes)
        if (leader && elect_one()) {
            const uint32_t idesc = instr_desc_bf16_m256(BN);
            const uint64_t adesc0 = smem_desc(smem0, 16, 1024, 2);
            const uint64_t bdesc0 = smem_desc(smem0 + RW_A_BYTES, 16, 1024, 2);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            bool stage_ready = false, acc_ready = false;
            RowsTile<2> t, tn;
            bool more = rows_tile<2>(p, 0, rank, t);
            for (int kk = 0; more; ++kk) {
                more = rows_tile<2>(p, kk + 1, rank, tn);
                if (!acc_ready) mbar_wait_cluster(tempty0 + 8 * acc, acc_phase ^ 1);
                const uint32_t tmem_d = tmem_base + acc * BN;
                const int nacc = acc ^ 1;
                const uint32_t nacc_phase = acc_phase ^ (uint32_t)acc;
                for (int it = 0; it < stages_per_tile; ++it) {
                    if (!stage_ready) mbar_wait_cluster(full0 + 8 * stage, phase);
                    fence_after();
                    const bool last = it == stages_per_tile - 1;
                    const int nstage = stage + 1 == STAGES ? 0 : stage + 1;
                    const uint32_t nphase = stage + 1 == STAGES ? phase ^ 1 : phase;
                    const uint64_t soff = (uint64_t)((stage * Cfg::STAGE_BYTES) >> 4);
                    uint32_t probe_stage = 0, probe_acc = 0;
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        const uint64_t adesc = adesc0 + soff + (uint64_t)((r * 1024) >> 4);
                        const uint64_t bdesc = bdesc0 + soff + (uint64_t)((r * Cfg::B_HALF) >> 4);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (r == 2 && k == 2) {
                                if (!last || more) probe_stage = mbar_try_wait_cluster(full0 + 8 * nstage, nphase) ? 1u : 0u;
                                if (last && more) probe_acc = mbar_try_wait_cluster(tempty0 + 8 * nacc, nacc_phase ^ 1) ? 1u : 0u;
                            }
                            umma_bf16_pair(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (uint32_t)(it | r | k));
                        }
                    }
                    umma_commit_pair(empty0 + 8 * stage, 3);
                    if (last) umma_commit_pair(tfull0 + 8 * acc, 3);
                    stage_ready = probe_stage != 0;
                    if (last) acc_ready = probe_acc != 0;
                    stage = nstage; phase = nphase;
                }
                acc = nacc; acc_phase = nacc_phase;
                t = tn;
            }
        }
        __syncwarp();
    } else {
        // ===================================================== epilogue: each CTA drains its own half (128 rows) of the accumulator
        bool narrow = false;
        if constexpr (BN <= 64) narrow = p.tiles_co == 1 && p.stats != nullptr && !p.accumulate;
        if constexpr (BN <= 64) {
            if (narrow) rows_epilogue<BN, bf16, true, 2, true>(p, warp, lane, tmem_base, tfull0, tempty0, s_stats, rank);
        }
        if (!narrow) rows_epilogue<BN, bf16, false, 2, true>(p, warp, lane, tmem_base, tfull0, tempty0, s_stats, rank);
    }
    fence_before();
    __syncthreads();
    cluster_sync_all();                 // the peer's shared memory / TMEM stay alive until the leader's last MMA has retired
    if (warp == 1) { fence_after(); tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS); }
}

bool tc_conv_rows_supported(int Ca, int Nout, int R, int S, int stride, int Ho, int Wo) {
    return R == 3 && S == 3 && stride == 1 && (Ca % 64 == 0 || Ca == 32) && Nout % 32 == 0 && Ho >= 16 && Wo >= 8;
}

// largest number of 2-CTA clusters of this kernel that can be resident at once (persistent grid = that many clusters)
template <int BN, typename OutT>
static int rows_max_clusters() {
    static int cached = -1;
    if (cached < 0) {
        using Cfg = RowsCfg<BN>;
        // the occupancy query honours the opt-in shared-memory limit only once it has been raised on the function
        cudaFuncSetAttribute(conv_tc_rows_kernel<BN, OutT, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(num_sms() / 2 * 2); cfg.blockDim = dim3(RW_THREADS); cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, conv_tc_rows_kernel<BN, OutT, 2>, &cfg) != cudaSuccess || n < 1) { cudaGetLastError(); n = 0; }
        cached = n;
    }
    return cached;
}
template <int BN, typename OutT, int CL, bool RESW = false>
static void launch_rows(cudaStream_t st, const CUtensorMap& ma, const CUtensorMap& mb, RowsParams p) {
    using Cfg = RowsCfg<BN, RESW>;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(conv_tc_rows_kernel<BN, OutT, CL, RESW>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        configured = true;
    }
    if constexpr (CL == 1) {
        const int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
        conv_tc_rows_kernel<BN, OutT, 1, RESW><<<grid, RW_THREADS, Cfg::SMEM_BYTES, st>>>(ma, mb, p);
    } else {
        const int clusters = std::min(rows_max_clusters<BN, OutT>(), p.total_groups);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(clusters * CL); cfg.blockDim = dim3(RW_THREADS); cfg.dynamicSmemBytes = Cfg::SMEM_BYTES; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, conv_tc_rows_kernel<BN, OutT, CL>, ma, mb, p);
        if (e != cudaSuccess) throw std::runtime_error(std::string("conv_tc_rows cluster launch failed: ") + cudaGetErrorString(e));
        ++g_salt_cluster_launches;
    }
}
// cluster size for the weight multicast: env SALT_TC_CLUSTER = 1 | 2 (read once).  Measured on B200 (bench.py, UNetResNet-34,
// profiles/r1_notes.md): CL = 2 runs exactly as fast as CL = 1 (19.02 vs 19.04 ms per step) while pulling ~30 % fewer bytes out of
// L2; CL = 4 was 2 % slower and is gone.
static int rows_cluster_pref() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SALT_TC_CLUSTER"); v = e ? atoi(e) : 2; if (v != 1 && v != 2) v = 1; }
    return v;
}
static int rows_wide_pref() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SALT_TC_WIDE"); v = (e && e[0] == '0') ? 0 : 1; }
    return v;
}
// CTA-pair kernel: env SALT_TC_PAIR = 0 turns it off (read once)
static int rows_pair_pref() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SALT_TC_PAIR"); v = (e && e[0] == '0') ? 0 : 1; }
    return v;
}
template <int BN>
static bool launch_rows_pair(cudaStream_t st, const CUtensorMap& ma, const void* Wp, int Ca, int Nout, RowsParams p) {
    using Cfg = PairCfg<BN>;
    static int max_clusters = -1;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(RW_THREADS); cfg.dynamicSmemBytes = Cfg::SMEM_BYTES; cfg.attrs = at; cfg.numAttrs = 1;
    if (max_clusters < 0) {
        cudaFuncSetAttribute(conv_tc_rows_pair_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        cfg.gridDim = dim3(num_sms() / 2 * 2);
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, conv_tc_rows_pair_kernel<BN>, &cfg) != cudaSuccess || n < 1) { cudaGetLastError(); n = 0; }
        max_clusters = n;
    }
    if (p.m_tiles < 16 || max_clusters * 2 < num_sms() * 3 / 4) return false;
    p.total_groups = cdiv(p.m_tiles, 2) * p.tiles_co;
    CUtensorMap mb;          // weights (c, n, r, s): one box = [3 taps r][BN/2][64 c] = this CTA's half of a kernel row's slabs
    {
        cuuint64_t dims[4] = {(cuuint64_t)Ca, (cuuint64_t)Nout, 3, 3};
        cuuint64_t strides[3] = {(cuuint64_t)9 * Ca * 2, (cuuint64_t)3 * Ca * 2, (cuuint64_t)Ca * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)(BN / 2), 3, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = get_encode()(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(Wp), dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled(weights 4d, pair) failed with code " + std::to_string((int)r));
    }
    cfg.gridDim = dim3(std::min(max_clusters, p.total_groups) * 2); cfg.stream = st;
    cudaError_t e = cudaLaunchKernelEx(&cfg, conv_tc_rows_pair_kernel<BN>, ma, mb, p);
    if (e != cudaSuccess) throw std::runtime_error(std::string("conv_tc_rows pair launch failed: ") + cudaGetErrorString(e));
    ++g_salt_cluster_launches;
    return true;
}
// resident weights: env SALT_TC_RESW = 0 turns them off (read once)
static int rows_resw_pref() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SALT_TC_RESW"); v = (e && e[0] == '0') ? 0 : 1; }
    return v;
}
template <int BN, typename OutT>
static void launch_rows_any(cudaStream_t st, const CUtensorMap& ma, const void* Wp, int Ca, int Nout, RowsParams p) {
    int cl = rows_cluster_pref();
    // <= 64 input channels, one output-channel tile, bf16: the nine weight slabs stay in shared memory (one CTA per tile)
    bool resw = false;
    if constexpr (sizeof(OutT) == 2 && BN <= 64) resw = rows_resw_pref() && p.cblks == 1 && p.tiles_co == 1 && p.m_tiles >= 16;
    if (resw) cl = 1;
    // a cluster only pays when there are enough pixel tiles to fill it and the machine with whole groups
    if (cl == 2 && (p.m_tiles < 16 || rows_max_clusters<BN, OutT>() * 2 < num_sms() * 3 / 4)) cl = 1;
    p.total_groups = cdiv(p.m_tiles, cl) * p.tiles_co;
    // weights Wp[n][(r*3+s)*Ca + c] viewed as a 4-D tensor (c, n, r, s).  CL == 1: one box = [3 taps r][BN][64 c];
    // CL == 2: a box is this CTA's BN/2 rows of ONE tap slab, multicast to both CTAs
    CUtensorMap mb;
    {
        cuuint64_t dims[4] = {(cuuint64_t)Ca, (cuuint64_t)Nout, 3, 3};
        cuuint64_t strides[3] = {(cuuint64_t)9 * Ca * 2, (cuuint64_t)3 * Ca * 2, (cuuint64_t)Ca * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)(BN / cl), cl == 1 ? 3u : 1u, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = get_encode()(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(Wp), dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled(weights 4d) failed with code " + std::to_string((int)r));
    }
    if constexpr (sizeof(OutT) == 2 && BN <= 64) {
        if (resw) { launch_rows<BN, OutT, 1, true>(st, ma, mb, p); return; }
    }
    if (cl == 2) launch_rows<BN, OutT, 2>(st, ma, mb, p);
    else launch_rows<BN, OutT, 1>(st, ma, mb, p);
}

// out[n,y,x,k] (+)= sum_{r,s,c} A[n, y+r-pad, x+s-pad, c] * Wp[k][(r*3+s)*Ca + c]      (3x3, stride 1); out_f32: fp32 output tensor
void k_conv_tc_rows(cudaStream_t st, const void* A, int B, int Ha, int Wa, int Ca, const void* Wp, int Nout, int pad, void* out,
                    int Ho, int Wo, const float* bias, float* stats, bool accumulate, bool out_f32, const EpiParams* ep, int split_c) {
    SALT_COUNT(1);
    if (out_f32 && (split_c <= 0 || split_c > Ca)) throw std::runtime_error("k_conv_tc_rows: fp32 output needs the split-operand channel count");
    if (out_f32 && accumulate) throw std::runtime_error("k_conv_tc_rows: accumulation into an fp32 output is not implemented");
    RowsParams p;
    if (ep) p.ep = *ep;
    if (p.ep.Hp == 0) { p.ep.Hp = Ho; p.ep.Wp = Wo; p.ep.pt = p.ep.pl = 0; }
    p.tiles_x = cdiv(Wo, RW_TW); p.tiles_y = cdiv(Ho, RW_TH);
    int BN = Nout % 128 == 0 ? 128 : Nout % 64 == 0 ? 64 : 32;
    // channel counts that are no multiple of 128 as few WIDE tiles instead of many N = 64 ones: 320 = 2 x 160, 192 = 1 x 192 (the
    // dgrads of the concat layers).  clk per MMA ~ 64 + N/2 (profiles/r2_notes.md), so a 160-wide instruction does 2.5x the MACs of
    // a 64-wide one in 1.5x the time.  SALT_TC_WIDE=0 turns it off.
    if (!out_f32 && rows_wide_pref() && BN == 64) { if (Nout % 192 == 0) BN = 192; else if (Nout % 160 == 0) BN = 160; }
    p.tiles_co = Nout / BN;
    p.d_co = make_fastdiv(p.tiles_co); p.d_x = make_fastdiv(p.tiles_x); p.d_y = make_fastdiv(p.tiles_y);
    p.total_tiles = p.tiles_x * p.tiles_y * B * p.tiles_co;
    p.B = B; p.Ho = Ho; p.Wo = Wo; p.Co = Nout; p.Ca = Ca; p.cblks = cdiv(Ca, 64); p.pad = pad;
    p.ksteps = Ca == 32 ? 2 : 4;
    p.accumulate = accumulate ? 1 : 0; p.bias = bias; p.stats = stats; p.out = out; p.split_c = split_c;
    CUtensorMap ma = make_map_nhwc(A, Ca, Wa, Ha, B, 64, RW_TW, RW_TH + 2, 1, 1, CU_TENSOR_MAP_SWIZZLE_128B);
    p.m_tiles = p.tiles_x * p.tiles_y * B; p.total_groups = 0;
    if (out_f32) {
        if (BN == 128) launch_rows_any<128, float>(st, ma, Wp, Ca, Nout, p);
        else if (BN == 64) launch_rows_any<64, float>(st, ma, Wp, Ca, Nout, p);
        else launch_rows_any<32, float>(st, ma, Wp, Ca, Nout, p);
        return;
    }
    if (BN == 128 && rows_pair_pref() && Ca % 64 == 0) {
        // >= 128 output channels: the CTA-pair kernel (half the weight bytes per SM; N = 256 tiles when the layer has them)
        if (Nout % 256 == 0) {
            RowsParams q = p; q.tiles_co = Nout / 256; q.d_co = make_fastdiv(q.tiles_co); q.total_tiles = q.tiles_x * q.tiles_y * B * q.tiles_co;
            if (launch_rows_pair<256>(st, ma, Wp, Ca, Nout, q)) return;
        }
        if (launch_rows_pair<128>(st, ma, Wp, Ca, Nout, p)) return;
    }
    if (BN == 128) launch_rows_any<128, bf16>(st, ma, Wp, Ca, Nout, p);
    else if (BN == 192) launch_rows_any<192, bf16>(st, ma, Wp, Ca, Nout, p);
    else if (BN == 160) launch_rows_any<160, bf16>(st, ma, Wp, Ca, Nout, p);
    else if (BN == 64) launch_rows_any<64, bf16>(st, ma, Wp, Ca, Nout, p);
    else launch_rows_any<32, bf16>(st, ma, Wp, Ca, Nout, p);
}

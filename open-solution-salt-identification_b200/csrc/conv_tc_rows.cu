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
// Three variants share the epilogue (rows_epilogue):
//   conv_tc_rows_kernel<BN, 1>       one CTA per tile
//   conv_tc_rows_kernel<BN, 2|4>     thread-block cluster, weight slabs TMA-multicast (default CL = 2; SALT_TC_CLUSTER)
//   conv_tc_rows_pair_kernel<BN>     CTA pair, cta_group::2 MMAs with M = 256 (experimental, SALT_TC_PAIR=1)
// Measured bound (profiles/r1_notes.md): tcgen05.mma fetches its shared-memory operands at ~64 B/clk and re-fetches the 4 KB A slice
// for every instruction -> ~(4096 + 32 N)/64 clk per MMA; L2/HBM traffic and the bytes entering the SM are not the limit (the
// cluster, pair and multi-sub-tile variants all cut them and none is faster).
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

struct RowsParams {
    int tiles_x, tiles_y, tiles_co, total_tiles;
    int m_tiles, total_groups;  // cluster mode: a group = CL consecutive pixel tiles of one channel tile (one per CTA of the cluster)
    int B, Ho, Wo, Co, Ca, cblks, pad;
    int accumulate;
    int debug;                 // timing experiments only (env SALT_TC_DEBUG): 1 = skip loads, 4 = skip stores+stats, 8 = skip stats, 16 = skip stores,
                               // 32 = per-tile butterflies for the BatchNorm sums also on narrow layers (the pre-round-1b epilogue)
    const float* bias;
    float* stats;              // [SALT_STAT_SLOTS_CONV][2*Co] partial slots, slot = blockIdx.x
    bf16* out;
};

constexpr int RW_THREADS = 320;                              // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue
constexpr int RW_EPI_THREADS = 256;
constexpr int RW_TH = 16, RW_TW = 8;
constexpr int RW_A_BYTES = (RW_TH + 2) * RW_TW * 128;        // 18 pixel rows x 8 pixels x 64 channels bf16 = 18 KB
template <int BN> struct RowsCfg {
    static constexpr int B_BYTES = BN * 128;                 // one tap: [BN][64] bf16
    static constexpr int STAGE_BYTES = RW_A_BYTES + 3 * B_BYTES;
    static constexpr int STAGES = BN == 128 ? 3 : (BN == 64 ? 4 : (BN == 32 ? 6 : 2));    // wide tiles (160 / 192): 78 / 90 KB stages
    // column distance of the two accumulators: BN for the power-of-two tiles, 256 for the wide ones (tcgen05.alloc wants 2^k columns)
    static constexpr int ACC_STRIDE = (BN & (BN - 1)) == 0 ? BN : 256;
    static constexpr int TMEM_COLS = 2 * ACC_STRIDE < 32 ? 32 : 2 * ACC_STRIDE;
    static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
    static constexpr int SMEM_BYTES = BAR_OFF + 1024 + 256 + 8 * 2 * BN * 4;      // + per-epilogue-warp BN statistics
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
// arrive on the mbarrier at the same shared-memory offset in CTA `cta` of the cluster (release at cluster scope)
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(bar), "r"(cta));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// wait on a local mbarrier whose arrivals come from other CTAs of the cluster (acquire at cluster scope); bounded like mbar_wait
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0, ok = 0;
    while (true) {
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2;\n"
            "selp.u32 %0, 1, 0, P1;\n"
            "}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) break;
        if (++spins > (1u << 24)) {
            printf("tc: cluster mbarrier wait timed out (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
            __trap();
        }
    }
}
// ---- CTA-pair (cta_group::2) wrappers: one MMA spans the tensor cores of both SMs of the pair (M = 256)
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
// instruction descriptor of the pair MMA: as tc::instr_desc_bf16 with M = 256
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
        t.nt = tile % p.tiles_co; mt = tile / p.tiles_co; t.live = true;
    } else {
        const int g = blockIdx.x / CL + k * (gridDim.x / CL);
        if (g >= p.total_groups) return false;
        t.nt = g % p.tiles_co; mt = (g / p.tiles_co) * CL + (int)rank;
        t.live = mt < p.m_tiles;
        if (!t.live) mt = p.m_tiles - 1;
    }
    t.tx = mt % p.tiles_x; t.ty = (mt / p.tiles_x) % p.tiles_y; t.n = mt / (p.tiles_x * p.tiles_y);
    return true;
}

template <int BN, bool NARROW, int CL, bool PAIR = false>
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
    float rs1[32], rs2[32], rbias[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) { rs1[i] = 0.f; rs2[i] = 0.f; rbias[i] = 0.f; }
    if (NARROW && own_chunk && p.bias) {
#pragma unroll
        for (int i = 0; i < 32; ++i) rbias[i] = __ldg(p.bias + half * 32 + i);
    }
    RowsTile<CL> t;
    for (int k = 0; rows_tile<CL>(p, k, rank, t); ++k) {
        const int nt = t.nt, n = t.n;
        const int x = t.tx * RW_TW + lx, y = t.ty * RW_TH + ly;
        const bool valid = t.live && (y < p.Ho) && (x < p.Wo) && !(p.debug & (4 | 16));
        bf16* orow = p.out + (((size_t)n * p.Ho + y) * p.Wo + x) * p.Co + nt * BN;
        // accumulate mode (dgrad into an existing gradient): the old values of the first chunk are fetched BEFORE waiting for the
        // accumulator, so their DRAM latency hides behind the MMA phase instead of serialising the epilogue
        uint4 old[4];
        const bool accum = !NARROW && p.accumulate;
        if (accum && valid && half < BN / 32) {
            const uint4* o4 = reinterpret_cast<const uint4*>(orow + half * 32);
#pragma unroll
            for (int q = 0; q < 4; ++q) old[q] = o4[q];
        }
        mbar_wait(tfull0 + 8 * acc, acc_phase);
        fence_after();
#pragma unroll 1
        for (int ch = half; ch < BN / 32; ch += 2) {
            float v[32];
            if (accum && valid && ch != half) {
                const uint4* o4 = reinterpret_cast<const uint4*>(orow + ch * 32);
#pragma unroll
                for (int q = 0; q < 4; ++q) old[q] = o4[q];
            }
            tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * RowsCfg<BN>::ACC_STRIDE + ch * 32, v);
            if constexpr (NARROW) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] += rbias[i];
            } else if (p.bias) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] += __ldg(p.bias + nt * BN + ch * 32 + i);
            }
            if (valid) {
                uint4* o4 = reinterpret_cast<uint4*>(orow + ch * 32);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (accum) {
                        const __nv_bfloat162* ob = reinterpret_cast<const __nv_bfloat162*>(&old[q]);
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
            if (p.stats && !(p.debug & (4 | 8))) {
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
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
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

// Sub-tile `sub` of work item k of this CTA in the multi-tile kernel: S consecutive pixel tiles of one channel tile share every
// weight stage; sub-tiles past the end of the tensor are clamped and dropped (`live` = false).
template <int S>
__device__ __forceinline__ bool multi_tile(const RowsParams& p, int k, int sub, RowsTile<1>& t) {
    const int g = blockIdx.x + k * gridDim.x;
    if (g >= p.total_groups) return false;
    t.nt = g % p.tiles_co;
    int mt = (g / p.tiles_co) * S + sub;
    t.live = mt < p.m_tiles;
    if (!t.live) mt = p.m_tiles - 1;
    t.tx = mt % p.tiles_x; t.ty = (mt / p.tiles_x) % p.tiles_y; t.n = mt / (p.tiles_x * p.tiles_y);
    return true;
}
// Epilogue of the multi-tile kernel: rows_epilogue with S accumulators per finished group (a textual twin of the function above,
// kept separate until the multi-tile kernel has been measured on the GPU so that the validated path stays byte-identical).
template <int BN, bool NARROW, int S>
__device__ __forceinline__ void rows_epilogue_multi(const RowsParams& p, const int warp, const int lane, const uint32_t tmem_base,
                                            const uint32_t tfull0, const uint32_t tempty0, float* s_stats) {
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
    float rs1[32], rs2[32], rbias[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) { rs1[i] = 0.f; rs2[i] = 0.f; rbias[i] = 0.f; }
    if (NARROW && own_chunk && p.bias) {
#pragma unroll
        for (int i = 0; i < 32; ++i) rbias[i] = __ldg(p.bias + half * 32 + i);
    }
    RowsTile<1> t;
    for (int k = 0;; ++k) {
      bool more = true;
#pragma unroll 1
      for (int sub = 0; sub < S; ++sub) {
        more = multi_tile<S>(p, k, sub, t);
        if (!more) break;
        const int nt = t.nt, n = t.n;
        const int x = t.tx * RW_TW + lx, y = t.ty * RW_TH + ly;
        const bool valid = t.live && (y < p.Ho) && (x < p.Wo) && !(p.debug & (4 | 16));
        bf16* orow = p.out + (((size_t)n * p.Ho + y) * p.Wo + x) * p.Co + nt * BN;
        // accumulate mode (dgrad into an existing gradient): the old values of the first chunk are fetched BEFORE waiting for the
        // accumulator, so their DRAM latency hides behind the MMA phase instead of serialising the epilogue
        uint4 old[4];
        const bool accum = !NARROW && p.accumulate;
        if (accum && valid && half < BN / 32) {
            const uint4* o4 = reinterpret_cast<const uint4*>(orow + half * 32);
#pragma unroll
            for (int q = 0; q < 4; ++q) old[q] = o4[q];
        }
        if (sub == 0) mbar_wait(tfull0 + 8 * acc, acc_phase);      // all S accumulators of the group complete together
        fence_after();
#pragma unroll 1
        for (int ch = half; ch < BN / 32; ch += 2) {
            float v[32];
            if (accum && valid && ch != half) {
                const uint4* o4 = reinterpret_cast<const uint4*>(orow + ch * 32);
#pragma unroll
                for (int q = 0; q < 4; ++q) old[q] = o4[q];
            }
            tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (acc * S + sub) * BN + ch * 32, v);
            if constexpr (NARROW) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] += rbias[i];
            } else if (p.bias) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] += __ldg(p.bias + nt * BN + ch * 32 + i);
            }
            if (valid) {
                uint4* o4 = reinterpret_cast<uint4*>(orow + ch * 32);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (accum) {
                        const __nv_bfloat162* ob = reinterpret_cast<const __nv_bfloat162*>(&old[q]);
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
            if (p.stats && !(p.debug & (4 | 8))) {
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
        if (sub == S - 1) {
            fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
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
      if (!more) break;
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
template <int BN, int CL>
__global__ void __launch_bounds__(RW_THREADS, 1)
conv_tc_rows_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const RowsParams p) {
    using Cfg = RowsCfg<BN>;
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

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        for (int i = 0; i < STAGES; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, CL); }
        for (int i = 0; i < 2; ++i) { mbar_init(tfull0 + 8 * i, 1); mbar_init(tempty0 + 8 * i, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_ptr_smem), Cfg::TMEM_COLS);
    for (int i = threadIdx.x; i < 8 * 2 * BN; i += RW_THREADS) s_stats[i] = 0.f;
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
            for (int k = 0; rows_tile<CL>(p, k, rank, t); ++k) {
                const int nt = t.nt, n = t.n;
                const int w0 = t.tx * RW_TW - p.pad, h0 = t.ty * RW_TH - p.pad;
                for (int cb = 0; cb < p.cblks; ++cb) {
                    for (int s = 0; s < 3; ++s) {
                        const uint32_t st = smem0 + stage * Cfg::STAGE_BYTES, fb = full0 + 8 * stage;
                        mbar_wait(empty0 + 8 * stage, phase ^ 1);
                        if (p.debug & 1) mbar_arrive(fb);
                        else {
                            mbar_expect_tx(fb, Cfg::STAGE_BYTES);
                            tma_load_4d(st, &map_a, fb, cb * 64, w0 + s, h0, n);
                            if constexpr (CL == 1) {
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
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer: 12 MMAs per barrier wait
        const uint32_t idesc = instr_desc_bf16(BN, false, false);
        const uint64_t adesc0 = smem_desc(smem0, 16, 1024, 2);
        const uint64_t bdesc0 = smem_desc(smem0 + RW_A_BYTES, 16, 1024, 2);
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        RowsTile<CL> t;
        for (int kk = 0; rows_tile<CL>(p, kk, rank, t); ++kk) {
            mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);
            fence_after();
            const uint32_t tmem_d = tmem_base + acc * Cfg::ACC_STRIDE;
            for (int it = 0; it < stages_per_tile; ++it) {
                mbar_wait(full0 + 8 * stage, phase);
                fence_after();
                if (elect_one()) {
                    const uint64_t soff = (uint64_t)((stage * Cfg::STAGE_BYTES) >> 4);
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        // vertical tap r = the same A buffer, r pixel-rows (r * 1024 bytes) further down
                        const uint64_t adesc = adesc0 + soff + (uint64_t)((r * 1024) >> 4);
                        const uint64_t bdesc = bdesc0 + soff + (uint64_t)((r * Cfg::B_BYTES) >> 4);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (it | r | k) != 0);
                    }
                    if constexpr (CL == 1) umma_commit(empty0 + 8 * stage);
                    else umma_commit_mc(empty0 + 8 * stage, MC_MASK);
                    if (it == stages_per_tile - 1) umma_commit(tfull0 + 8 * acc);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else {
        // ===================================================== epilogue: 8 warps (2..9)
        bool narrow = false;
        if constexpr (BN <= 64) narrow = p.tiles_co == 1 && p.stats != nullptr && !p.accumulate && !(p.debug & 32);
        if constexpr (BN <= 64) {
            if (narrow) rows_epilogue<BN, true, CL>(p, warp, lane, tmem_base, tfull0, tempty0, s_stats, rank);
        }
        if (!narrow) rows_epilogue<BN, false, CL>(p, warp, lane, tmem_base, tfull0, tempty0, s_stats, rank);
    }
    fence_before();
    __syncthreads();
    if constexpr (CL > 1) cluster_sync_all();      // no CTA leaves while a peer's commit may still arrive on its barriers
    if (warp == 1) { fence_after(); tmem_dealloc(tmem_base, Cfg::TMEM_COLS); }
}


// ================================================================================================
// EXPERIMENTAL (env SALT_TC_PAIR=1, off by default; not yet measured): the same convolution with cta_group::2 MMAs.
// The two CTAs of a cluster form a CTA pair: one tcgen05.mma (M = 256, N = BN) issued by the leader covers both pixel tiles;
// each CTA stages its own activation tile and only HALF of every weight slab (B is split in N across the pair), so per stage a
// CTA's shared memory serves (4 KB + N/2 * 32 B) per MMA instead of (4 KB + N * 32 B) and takes 18 KB + 1.5 * N * 128 B of TMA
// writes instead of 18 KB + 3 * N * 128 B - measured on B200: correct, but 1.45x slower (per-stage cross-CTA hand-off), see profiles/r1_notes.md.
// Protocol:  full[s]      (each CTA, 1 arrival + bytes)  own TMA loads landed
//            peerfull[s]  (leader, 1 arrival)            the peer forwards its full[s] (remote arrive by its otherwise idle warp 1)
//            empty[s]     (each CTA, 1 arrival)          leader's tcgen05.commit.cta_group::2, multicast to both CTAs
//            tfull[a]     (each CTA, 1 arrival)          accumulator a complete, same commit multicast; each CTA drains its own TMEM
//            tempty[a]    (leader, 16 arrivals)          8 epilogue warps of each CTA (remote arrive from the peer)
template <int BN> struct PairCfg {
    static constexpr int B_HALF = (BN / 2) * 128;                        // one tap, my half of the channels: [BN/2][64] bf16
    static constexpr int STAGE_BYTES = RW_A_BYTES + 3 * B_HALF;
    static constexpr int STAGES = BN == 128 ? 4 : (BN == 64 ? 6 : 7);
    static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
    static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
    static constexpr int SMEM_BYTES = BAR_OFF + 1024 + 512 + 8 * 2 * BN * 4;
};
template <int BN>
__global__ void __launch_bounds__(RW_THREADS, 1)
conv_tc_rows_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const RowsParams p) {
    using Cfg = PairCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
    // bars: full[STAGES], empty[STAGES], peerfull[STAGES], tmem_full[2], tmem_empty[2]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 4);
    float* s_stats = reinterpret_cast<float*>(smem + Cfg::BAR_OFF + 512);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem0 = smem_u32(smem);
    const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * STAGES, pfull0 = empty0 + 8 * STAGES, tfull0 = pfull0 + 8 * STAGES,
                   tempty0 = tfull0 + 16;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        for (int i = 0; i < STAGES; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); mbar_init(pfull0 + 8 * i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(tfull0 + 8 * i, 1); mbar_init(tempty0 + 8 * i, 16); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc_pair(smem_u32(tmem_ptr_smem), Cfg::TMEM_COLS);
    for (int i = threadIdx.x; i < 8 * 2 * BN; i += RW_THREADS) s_stats[i] = 0.f;
    fence_before();
    __syncthreads();
    cluster_sync_all();
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
                        mbar_wait(empty0 + 8 * stage, phase ^ 1);
                        mbar_expect_tx(fb, Cfg::STAGE_BYTES);
                        tma_load_4d(st, &map_a, fb, cb * 64, w0 + s, h0, t.n);
                        tma_load_4d(st + RW_A_BYTES, &map_b, fb, cb * 64, t.nt * BN + (int)rank * (BN / 2), 0, s);   // (c, n half, r = 0..2, s)
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (!leader) {
            // ===================================================== peer: forward "my stage has landed" to the leader
            if (elect_one()) {
                int stage = 0; uint32_t phase = 0;
                RowsTile<2> t;
                for (int kk = 0; rows_tile<2>(p, kk, rank, t); ++kk)
                    for (int it = 0; it < stages_per_tile; ++it) {
                        mbar_wait(full0 + 8 * stage, phase);
                        mbar_arrive_remote(pfull0 + 8 * stage, 0);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
            }
        } else {
            // ===================================================== leader: MMA issuer for the pair
            const uint32_t idesc = instr_desc_bf16_m256(BN);
            const uint64_t adesc0 = smem_desc(smem0, 16, 1024, 2);
            const uint64_t bdesc0 = smem_desc(smem0 + RW_A_BYTES, 16, 1024, 2);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            RowsTile<2> t;
            for (int kk = 0; rows_tile<2>(p, kk, rank, t); ++kk) {
                mbar_wait_cluster(tempty0 + 8 * acc, acc_phase ^ 1);
                fence_after();
                const uint32_t tmem_d = tmem_base + acc * BN;
                for (int it = 0; it < stages_per_tile; ++it) {
                    mbar_wait(full0 + 8 * stage, phase);
                    mbar_wait_cluster(pfull0 + 8 * stage, phase);
                    fence_after();
                    if (elect_one()) {
                        const uint64_t soff = (uint64_t)((stage * Cfg::STAGE_BYTES) >> 4);
#pragma unroll
                        for (int r = 0; r < 3; ++r) {
                            const uint64_t adesc = adesc0 + soff + (uint64_t)((r * 1024) >> 4);
                            const uint64_t bdesc = bdesc0 + soff + (uint64_t)((r * Cfg::B_HALF) >> 4);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_bf16_pair(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (it | r | k) != 0);
                        }
                        umma_commit_pair(empty0 + 8 * stage, 3);
                        if (it == stages_per_tile - 1) umma_commit_pair(tfull0 + 8 * acc, 3);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================================================== epilogue: each CTA drains its own half (128 rows) of the accumulator
        bool narrow = false;
        if constexpr (BN <= 64) narrow = p.tiles_co == 1 && p.stats != nullptr && !p.accumulate && !(p.debug & 32);
        if constexpr (BN <= 64) {
            if (narrow) rows_epilogue<BN, true, 2, true>(p, warp, lane, tmem_base, tfull0, tempty0, s_stats, rank);
        }
        if (!narrow) rows_epilogue<BN, false, 2, true>(p, warp, lane, tmem_base, tfull0, tempty0, s_stats, rank);
    }
    fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) { fence_after(); tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS); }
}


// ================================================================================================
// EXPERIMENTAL (env SALT_TC_MULTI=1, off by default; bench on B200: same speed as the default kernel, numerics not yet checked by a
// parity case large enough to reach it): every weight stage is
// re-used for S pixel sub-tiles inside ONE CTA.  Round-1 measurements (profiles/r1_notes.md): the bytes that have to enter an SM
// per MMA bound conv_tc_rows_kernel (~45 B/clk/SM arrive in every shape) and 60-75 % of them are weights.  TMEM holds 512 columns,
// i.e. 2 x S accumulators of BN columns with S = 4 (BN <= 64) or 2 (BN = 128); per 12 MMAs a CTA then loads 18 + 24/S KB instead of
// 42 KB (BN = 64) or 18 + 48/S instead of 66 KB (BN = 128).
// Two shared-memory rings with their own barriers: weights (NB stages of the 3 slabs of a kernel row) and activations (NA boxes).
//   producer:  per (channel block, s): weight stage, then the S activation boxes of the group's sub-tiles
//   issuer:    per weight stage: for each sub-tile wait its box, 12 MMAs into accumulator (acc, sub), commit -> box free;
//              after the S sub-tiles commit -> weight stage free; after the last stage commit -> tfull[acc]
//   epilogue:  rows_epilogue_multi drains the S accumulators of the finished group, then frees them with one tempty arrival per warp
template <int BN, int S> struct MultiCfg {
    static constexpr int B_BYTES = BN * 128;
    static constexpr int W_STAGE = 3 * B_BYTES;                          // three tap slabs of one kernel row
    static constexpr int NB = 2;
    static constexpr int NA = BN == 128 ? 5 : 8;
    static constexpr int TMEM_COLS = 2 * S * BN < 32 ? 32 : 2 * S * BN;
    static constexpr int A_OFF = NB * W_STAGE;
    static constexpr int BAR_OFF = A_OFF + NA * RW_A_BYTES;
    static constexpr int SMEM_BYTES = BAR_OFF + 1024 + 512 + 8 * 2 * BN * 4;
    static_assert(TMEM_COLS <= 512, "accumulators do not fit TMEM");
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};
template <int BN, int S>
__global__ void __launch_bounds__(RW_THREADS, 1)
conv_tc_rows_multi_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const RowsParams p) {
    using Cfg = MultiCfg<BN, S>;
    constexpr int NA = Cfg::NA, NB = Cfg::NB;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
    // bars: bfull[NB], bempty[NB], afull[NA], aempty[NA], tmem_full[2], tmem_empty[2]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * NB + 2 * NA + 4);
    float* s_stats = reinterpret_cast<float*>(smem + Cfg::BAR_OFF + 512);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem0 = smem_u32(smem);
    const uint32_t bfull0 = smem_u32(bars), bempty0 = bfull0 + 8 * NB, afull0 = bempty0 + 8 * NB, aempty0 = afull0 + 8 * NA,
                   tfull0 = aempty0 + 8 * NA, tempty0 = tfull0 + 16;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        for (int i = 0; i < NB; ++i) { mbar_init(bfull0 + 8 * i, 1); mbar_init(bempty0 + 8 * i, 1); }
        for (int i = 0; i < NA; ++i) { mbar_init(afull0 + 8 * i, 1); mbar_init(aempty0 + 8 * i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(tfull0 + 8 * i, 1); mbar_init(tempty0 + 8 * i, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_ptr_smem), Cfg::TMEM_COLS);
    for (int i = threadIdx.x; i < 8 * 2 * BN; i += RW_THREADS) s_stats[i] = 0.f;
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const int stages_per_group = p.cblks * 3;

    if (warp == 0) {
        // ===================================================== TMA producer
        if (elect_one()) {
            int bs = 0, as = 0; uint32_t bphase = 0, aphase = 0;
            RowsTile<1> t[S];
            for (int k = 0;; ++k) {
                bool more = true;
                for (int sub = 0; sub < S; ++sub) more = multi_tile<S>(p, k, sub, t[sub]) && more;
                if (!more) break;
                for (int cb = 0; cb < p.cblks; ++cb) {
                    for (int s = 0; s < 3; ++s) {
                        mbar_wait(bempty0 + 8 * bs, bphase ^ 1);
                        mbar_expect_tx(bfull0 + 8 * bs, Cfg::W_STAGE);
                        tma_load_4d(smem0 + bs * Cfg::W_STAGE, &map_b, bfull0 + 8 * bs, cb * 64, t[0].nt * BN, 0, s);   // (c, n, r = 0..2, s)
                        if (++bs == NB) { bs = 0; bphase ^= 1; }
#pragma unroll
                        for (int sub = 0; sub < S; ++sub) {
                            mbar_wait(aempty0 + 8 * as, aphase ^ 1);
                            mbar_expect_tx(afull0 + 8 * as, RW_A_BYTES);
                            tma_load_4d(smem0 + Cfg::A_OFF + as * RW_A_BYTES, &map_a, afull0 + 8 * as, cb * 64,
                                        t[sub].tx * RW_TW - p.pad + s, t[sub].ty * RW_TH - p.pad, t[sub].n);
                            if (++as == NA) { as = 0; aphase ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer: S x 12 MMAs per weight stage
        const uint32_t idesc = instr_desc_bf16(BN, false, false);
        const uint64_t adesc0 = smem_desc(smem0 + Cfg::A_OFF, 16, 1024, 2);
        const uint64_t bdesc0 = smem_desc(smem0, 16, 1024, 2);
        int bs = 0, as = 0; uint32_t bphase = 0, aphase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        RowsTile<1> t0;
        for (int kk = 0; multi_tile<S>(p, kk, 0, t0); ++kk) {
            mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);
            fence_after();
            for (int it = 0; it < stages_per_group; ++it) {
                mbar_wait(bfull0 + 8 * bs, bphase);
                const uint64_t boff = (uint64_t)((bs * Cfg::W_STAGE) >> 4);
#pragma unroll 1
                for (int sub = 0; sub < S; ++sub) {
                    mbar_wait(afull0 + 8 * as, aphase);
                    fence_after();
                    if (elect_one()) {
                        const uint32_t tmem_d = tmem_base + (acc * S + sub) * BN;
                        const uint64_t aoff = (uint64_t)((as * RW_A_BYTES) >> 4);
#pragma unroll
                        for (int r = 0; r < 3; ++r) {
                            const uint64_t adesc = adesc0 + aoff + (uint64_t)((r * 1024) >> 4);
                            const uint64_t bdesc = bdesc0 + boff + (uint64_t)((r * Cfg::B_BYTES) >> 4);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_bf16(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (it | r | k) != 0);
                        }
                        umma_commit(aempty0 + 8 * as);                                   // this activation box is free again
                        if (sub == S - 1) {
                            umma_commit(bempty0 + 8 * bs);                               // ... and so is the weight stage
                            if (it == stages_per_group - 1) umma_commit(tfull0 + 8 * acc);
                        }
                    }
                    __syncwarp();
                    if (++as == NA) { as = 0; aphase ^= 1; }
                }
                if (++bs == NB) { bs = 0; bphase ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else {
        // ===================================================== epilogue: 8 warps (2..9)
        bool narrow = false;
        if constexpr (BN <= 64) narrow = p.tiles_co == 1 && p.stats != nullptr && !p.accumulate && !(p.debug & 32);
        if constexpr (BN <= 64) {
            if (narrow) rows_epilogue_multi<BN, true, S>(p, warp, lane, tmem_base, tfull0, tempty0, s_stats);
        }
        if (!narrow) rows_epilogue_multi<BN, false, S>(p, warp, lane, tmem_base, tfull0, tempty0, s_stats);
    }
    fence_before();
    __syncthreads();
    if (warp == 1) { fence_after(); tmem_dealloc(tmem_base, Cfg::TMEM_COLS); }
}

bool tc_conv_rows_supported(int Ca, int Nout, int R, int S, int stride, int Ho, int Wo) {
    return R == 3 && S == 3 && stride == 1 && Ca % 64 == 0 && Nout % 32 == 0 && Ho >= 16 && Wo >= 8;
}

// largest number of CL-CTA clusters of this kernel that can be resident at once (persistent grid = that many clusters)
template <int BN, int CL>
static int rows_max_clusters() {
    static int cached = -1;
    if (cached < 0) {
        using Cfg = RowsCfg<BN>;
        // the occupancy query honours the opt-in shared-memory limit only once it has been raised on the function
        cudaFuncSetAttribute(conv_tc_rows_kernel<BN, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(num_sms() / CL * CL); cfg.blockDim = dim3(RW_THREADS); cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, conv_tc_rows_kernel<BN, CL>, &cfg) != cudaSuccess || n < 1) { cudaGetLastError(); n = 0; }
        cached = n;
    }
    return cached;
}
template <int BN, int CL>
static void launch_rows(cudaStream_t st, const CUtensorMap& ma, const CUtensorMap& mb, RowsParams p) {
    using Cfg = RowsCfg<BN>;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(conv_tc_rows_kernel<BN, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        configured = true;
    }
    if constexpr (CL == 1) {
        const int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
        conv_tc_rows_kernel<BN, 1><<<grid, RW_THREADS, Cfg::SMEM_BYTES, st>>>(ma, mb, p);
    } else {
        const int clusters = std::min(rows_max_clusters<BN, CL>(), p.total_groups);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(clusters * CL); cfg.blockDim = dim3(RW_THREADS); cfg.dynamicSmemBytes = Cfg::SMEM_BYTES; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, conv_tc_rows_kernel<BN, CL>, ma, mb, p);
        if (e != cudaSuccess) throw std::runtime_error(std::string("conv_tc_rows cluster launch failed: ") + cudaGetErrorString(e));
        ++g_salt_cluster_launches;
    }
}
// cluster size for the weight multicast: env SALT_TC_CLUSTER = 1 | 2 | 4.  Measured on B200 (bench.py, UNetResNet-34, profiles/
// r1_notes.md): CL = 2 runs exactly as fast as CL = 1 (19.02 vs 19.04 ms per step) while pulling ~30 % fewer bytes out of L2,
// CL = 4 is 2 % slower (36 clusters of 4 leave 4 SMs idle and the 4-way lockstep adds skew).  The kernel is bound by
// shared-memory bandwidth (tensor-core operand reads + TMA writes), which multicast does not lower - default 2.
static int rows_cluster_pref() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SALT_TC_CLUSTER"); v = e ? atoi(e) : 2; if (v != 1 && v != 2 && v != 4) v = 1; }
    return v;
}
static int rows_multi_pref() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SALT_TC_MULTI"); v = (e && e[0] == '1') ? 1 : 0; }
    return v;
}
template <int BN, int S>
static void launch_rows_multi_s(cudaStream_t st, const CUtensorMap& ma, const CUtensorMap& mb, RowsParams p) {
    using Cfg = MultiCfg<BN, S>;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(conv_tc_rows_multi_kernel<BN, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        configured = true;
    }
    p.total_groups = cdiv(p.m_tiles, S) * p.tiles_co;
    const int grid = p.total_groups < num_sms() ? p.total_groups : num_sms();
    conv_tc_rows_multi_kernel<BN, S><<<grid, RW_THREADS, Cfg::SMEM_BYTES, st>>>(ma, mb, p);
}
// S sub-tiles per weight stage: as many as TMEM allows, fewer when the layer has too few pixel tiles to keep every SM busy
template <int BN>
static bool launch_rows_multi(cudaStream_t st, const CUtensorMap& ma, const void* Wp, int Ca, int Nout, const RowsParams& p) {
    int S = BN == 128 ? 2 : 4;
    while (S > 1 && (long long)cdiv(p.m_tiles, S) * p.tiles_co < 2LL * num_sms()) S >>= 1;
    if (S == 1) return false;
    CUtensorMap mb;          // weights (c, n, r, s): one box = [3 taps r][BN][64 c]
    {
        cuuint64_t dims[4] = {(cuuint64_t)Ca, (cuuint64_t)Nout, 3, 3};
        cuuint64_t strides[3] = {(cuuint64_t)9 * Ca * 2, (cuuint64_t)3 * Ca * 2, (cuuint64_t)Ca * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)BN, 3, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = get_encode()(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(Wp), dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled(weights 4d, multi) failed with code " + std::to_string((int)r));
    }
    if constexpr (BN == 128) {
        launch_rows_multi_s<BN, 2>(st, ma, mb, p);
    } else {
        if (S == 4) launch_rows_multi_s<BN, 4>(st, ma, mb, p); else launch_rows_multi_s<BN, 2>(st, ma, mb, p);
    }
    return true;
}
static int rows_pair_pref() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SALT_TC_PAIR"); v = (e && e[0] == '1') ? 1 : 0; }
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
template <int BN>
static void launch_rows_any(cudaStream_t st, const CUtensorMap& ma, const void* Wp, int Ca, int Nout, RowsParams p) {
    if constexpr ((BN & (BN - 1)) == 0) {
        if (rows_multi_pref() && launch_rows_multi<BN>(st, ma, Wp, Ca, Nout, p)) return;
        if (rows_pair_pref() && launch_rows_pair<BN>(st, ma, Wp, Ca, Nout, p)) return;
    }
    int cl = rows_cluster_pref();
    // a cluster only pays when there are enough pixel tiles to fill it and the machine with whole groups
    while (cl > 1 && (p.m_tiles < cl * 8 || (cl == 4 ? rows_max_clusters<BN, 4>() : rows_max_clusters<BN, 2>()) * cl < num_sms() * 3 / 4)) cl >>= 1;
    p.total_groups = cdiv(p.m_tiles, cl) * p.tiles_co;
    // weights Wp[n][(r*3+s)*Ca + c] viewed as a 4-D tensor (c, n, r, s).  CL == 1: one box = [3 taps r][BN][64 c];
    // CL > 1: a box is this CTA's BN/CL rows of ONE tap slab, multicast to the whole cluster
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
    if (cl == 4) launch_rows<BN, 4>(st, ma, mb, p);
    else if (cl == 2) launch_rows<BN, 2>(st, ma, mb, p);
    else launch_rows<BN, 1>(st, ma, mb, p);
}

// out[n,y,x,k] (+)= sum_{r,s,c} A[n, y+r-pad, x+s-pad, c] * Wp[k][(r*3+s)*Ca + c]      (3x3, stride 1)
void k_conv_tc_rows(cudaStream_t st, const void* A, int B, int Ha, int Wa, int Ca, const void* Wp, int Nout, int pad, void* out,
                    int Ho, int Wo, const float* bias, float* stats, bool accumulate) {
    SALT_COUNT(1);
    RowsParams p;
    p.tiles_x = cdiv(Wo, RW_TW); p.tiles_y = cdiv(Ho, RW_TH);
    int BN = Nout % 128 == 0 ? 128 : Nout % 64 == 0 ? 64 : 32;
    // EXPERIMENTAL (env SALT_TC_WIDE=1, not yet run on a GPU): channel counts that are no multiple of 128 as few wide tiles instead of
    // many N = 64 ones - 320 = 2 x 160, 192 = 1 x 192 (the concat-layer dgrads).  clk per MMA ~ 64 + N/2 (profiles/r1_notes.md), so a
    // 160-wide instruction does 2.5x the MACs of a 64-wide one in 1.5x the time.
    static int wide = -1;
    if (wide < 0) { const char* e = getenv("SALT_TC_WIDE"); wide = (e && e[0] == '1') ? 1 : 0; }
    if (wide && BN == 64) { if (Nout % 192 == 0) BN = 192; else if (Nout % 160 == 0) BN = 160; }
    p.tiles_co = Nout / BN;
    p.total_tiles = p.tiles_x * p.tiles_y * B * p.tiles_co;
    p.B = B; p.Ho = Ho; p.Wo = Wo; p.Co = Nout; p.Ca = Ca; p.cblks = Ca / 64; p.pad = pad;
    p.accumulate = accumulate ? 1 : 0; p.bias = bias; p.stats = stats; p.out = (bf16*)out;
    { const char* e = getenv("SALT_TC_DEBUG"); p.debug = e ? atoi(e) : 0; }
    CUtensorMap ma = make_map_nhwc(A, Ca, Wa, Ha, B, 64, RW_TW, RW_TH + 2, 1, 1, CU_TENSOR_MAP_SWIZZLE_128B);
    p.m_tiles = p.tiles_x * p.tiles_y * B; p.total_groups = 0;
    if (BN == 128) launch_rows_any<128>(st, ma, Wp, Ca, Nout, p);
    else if (BN == 192) launch_rows_any<192>(st, ma, Wp, Ca, Nout, p);
    else if (BN == 160) launch_rows_any<160>(st, ma, Wp, Ca, Nout, p);
    else if (BN == 64) launch_rows_any<64>(st, ma, Wp, Ca, Nout, p);
    else launch_rows_any<32>(st, ma, Wp, Ca, Nout, p);
}

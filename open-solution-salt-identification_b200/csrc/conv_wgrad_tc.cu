// tcgen05 weight-gradient convolution for sm_100a.
//
//   dw[k][c][r][s] += sum_{n,y,x} gout[n,y,x,k] * in[n, y*stride + r - pad, x*stride + s - pad, c]        (zero outside `in`)
//
// GEMM view per filter tap: D_tap[c][k] = X_tap^T (c x pixels) * G (pixels x k); the reduction runs over output pixels.
// Both operands are "MN-major" for the tensor core (channels contiguous, the reduction index = pixel is the row index), which is
// exactly what a TMA box load of an NHWC tensor produces: a [32 pixels][64 channels] 128B-swizzled tile.
//   * M = 128 rows of D = two "slots" (slot = (tap, 64-channel block of the input)) stacked through the descriptor's
//     leading-byte-offset, so 64-channel layers still issue full M=128 instructions;
//   * N = NK output channels (64 or 128);  K = 32 pixels per pipeline stage (two K=16 instructions per slot pair);
//   * one CTA owns up to 5 (NK=64) or 4 (NK=128) slot pairs -> that many fp32 accumulators live in TMEM for the whole pixel loop;
//   * grid = (pixel splits, slot chunks, k tiles); partial results are added with 16-byte vector reductions
//     (red.global.add.v4.f32) into a k-contiguous fp32 scratch dwp[tap][c][k]; unpack_dw_kernel then transposes that
//     into the reference layout dw[k][c][r][s] through shared memory (coalesced on both sides).
// Warp roles as in conv_tc.cu: warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue.
#include "tc_common.cuh"
#include "conv_tc.h"

using namespace tc;

struct WgParams {
    int B, Ho, Wo;
    int Ci_pad, Co;
    int RS, S, stride, pad;
    int cblks, total_slots, slots_per_cta;
    int pw, ph, chunks_x, chunks_y, total_chunks, chunks_per_split;
    float* dwp;                 // [RS][Ci_pad][Co] fp32, zero-initialised
};

constexpr int WG_THREADS = 192;
constexpr int WG_TILE = 4096;                       // [32 pixels][64 channels] bf16
template <int NK> struct WgCfg {
    static constexpr int MAX_GROUPS = NK == 64 ? 5 : 4;
    static constexpr int MAX_SLOTS = 2 * MAX_GROUPS;
    static constexpr int G_TILES = NK / 64;
    static constexpr int STAGE_BYTES = (MAX_SLOTS + G_TILES) * WG_TILE;
    static constexpr int STAGES = 4;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

template <int NK>
__global__ void __launch_bounds__(WG_THREADS, 1)
conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_g, const WgParams p) {
    using Cfg = WgCfg<NK>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES), tfull = smem_u32(bars + 2 * STAGES);

    // this CTA's work
    const int slot0 = blockIdx.y * p.slots_per_cta;
    const int nslots = min(p.slots_per_cta, p.total_slots - slot0);
    const int ngroups = (nslots + 1) >> 1;
    const int k0 = blockIdx.z * NK;
    const int chunk_begin = blockIdx.x * p.chunks_per_split;
    const int chunk_end = min(p.total_chunks, chunk_begin + p.chunks_per_split);
    const int nchunks = chunk_end - chunk_begin;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_g) : "memory");
        for (int i = 0; i < STAGES; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
        mbar_init(tfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_ptr_smem), 512);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (nchunks > 0 && nslots > 0) {
        if (warp == 0) {
            // ===================================================== TMA producer
            if (elect_one()) {
                int stage = 0; uint32_t phase = 0;
                const uint32_t tx_bytes = (uint32_t)(nslots + Cfg::G_TILES) * WG_TILE;
                // per-slot constants (no divisions inside the pixel loop: this single thread must keep up with the tensor core)
                int s_c[Cfg::MAX_SLOTS], s_dx[Cfg::MAX_SLOTS], s_dy[Cfg::MAX_SLOTS];
#pragma unroll
                for (int i = 0; i < Cfg::MAX_SLOTS; ++i) {
                    const int slot = slot0 + (i < nslots ? i : 0), tap = slot / p.cblks, cb = slot - tap * p.cblks;
                    const int r = tap / p.S, s = tap - r * p.S;
                    s_c[i] = cb * 64; s_dx[i] = s - p.pad; s_dy[i] = r - p.pad;
                }
                int cx = chunk_begin % p.chunks_x, cy = (chunk_begin / p.chunks_x) % p.chunks_y, n = chunk_begin / (p.chunks_x * p.chunks_y);
                for (int ck = chunk_begin; ck < chunk_end; ++ck) {
                    const int x0 = cx * p.pw, y0 = cy * p.ph;
                    const uint32_t st = smem_u32(smem + stage * Cfg::STAGE_BYTES), fb = full0 + 8 * stage;
                    mbar_wait(empty0 + 8 * stage, phase ^ 1);
                    mbar_expect_tx(fb, tx_bytes);
#pragma unroll
                    for (int i = 0; i < Cfg::MAX_SLOTS; ++i)
                        if (i < nslots) tma_load_4d(st + i * WG_TILE, &map_x, fb, s_c[i], x0 * p.stride + s_dx[i], y0 * p.stride + s_dy[i], n);
#pragma unroll
                    for (int j = 0; j < Cfg::G_TILES; ++j)
                        tma_load_4d(st + (Cfg::MAX_SLOTS + j) * WG_TILE, &map_g, fb, k0 + j * 64, x0, y0, n);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    if (++cx == p.chunks_x) { cx = 0; if (++cy == p.chunks_y) { cy = 0; ++n; } }
                }
            }
        } else if (warp == 1) {
            // ===================================================== MMA issuer
            const uint32_t idesc = instr_desc_bf16(NK, true, true);
            // MN-major, 128B swizzle: 8-pixel groups 1024 B apart (SBO), 64-channel blocks LBO apart.  Descriptors are built once;
            // per stage / K step only the 14-bit start-address field advances (16-byte units).
            const uint32_t smem0 = smem_u32(smem);
            uint64_t adesc0[Cfg::MAX_GROUPS];
#pragma unroll
            for (int g = 0; g < Cfg::MAX_GROUPS; ++g)      // odd tail: both halves read the same tile (LBO = 0)
                adesc0[g] = smem_desc(smem0 + 2 * g * WG_TILE, (2 * g + 1 < nslots) ? WG_TILE : 0, 1024, 2);
            const uint64_t bdesc0 = smem_desc(smem0 + Cfg::MAX_SLOTS * WG_TILE, WG_TILE, 1024, 2);
            // one elected thread runs the whole loop; the next stage's barrier is probed before the last slot pair's MMAs so that the
            // ~90-cycle probe latency hides behind their execution (conv_tc_rows.cu explains the measurement behind this)
            if (elect_one()) {
                int stage = 0; uint32_t phase = 0;
                bool stage_ready = false;
                for (int it = 0; it < nchunks; ++it) {
                    if (!stage_ready) mbar_wait(full0 + 8 * stage, phase);
                    fence_after();
                    const int nstage = stage + 1 == STAGES ? 0 : stage + 1;
                    const uint32_t nphase = stage + 1 == STAGES ? phase ^ 1 : phase;
                    const uint64_t soff = (uint64_t)((stage * Cfg::STAGE_BYTES) >> 4);
                    uint32_t probe = 0;
#pragma unroll
                    for (int g = 0; g < Cfg::MAX_GROUPS; ++g) {
                        if (g < ngroups) {
                            if (g == ngroups - 1 && it + 1 < nchunks) probe = mbar_try_wait(full0 + 8 * nstage, nphase) ? 1u : 0u;
#pragma unroll
                            for (int kk = 0; kk < 2; ++kk)
                                umma_bf16(tmem_base + g * NK, adesc0[g] + soff + (uint64_t)(kk * 128), bdesc0 + soff + (uint64_t)(kk * 128),
                                          idesc, (it | kk) != 0);
                        }
                    }
                    umma_commit(empty0 + 8 * stage);
                    if (it == nchunks - 1) umma_commit(tfull);
                    stage_ready = probe != 0;
                    stage = nstage; phase = nphase;
                }
            }
            __syncwarp();
        } else {
            // ===================================================== epilogue: TMEM -> red.global.add.v4.f32
            const int quarter = warp & 3;
            const int m = quarter * 32 + lane;
            mbar_wait(tfull, 0);
            fence_after();
            for (int g = 0; g < ngroups; ++g) {
                const int slot = slot0 + 2 * g + (m >> 6);
                const bool row_ok = (2 * g + (m >> 6)) < nslots;
                const int tap = slot / p.cblks, cb = slot - tap * p.cblks;
                const int c = cb * 64 + (m & 63);
                float* drow = p.dwp + ((size_t)tap * p.Ci_pad + c) * p.Co + k0;
#pragma unroll 1
                for (int ch = 0; ch < NK / 32; ++ch) {
                    float v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + g * NK + ch * 32, v);
                    if (row_ok) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            if (k0 + ch * 32 + j < p.Co)
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(drow + ch * 32 + j), "f"(v[j]),
                                             "f"(v[j + 1]), "f"(v[j + 2]), "f"(v[j + 3]) : "memory");
                        }
                    }
                }
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 1) { fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ------------------------------------------------------------------------------------------------ halo-box variant (3x3, stride 1)
// The kernel above loads one shifted [32 px][64 ch] input tile PER TAP: 9 x 4 KB + the gradient tile per 10 MMAs = 64-80 B/clk of
// L2 -> SM traffic, above what the chip delivers to all SMs at once (~42 B/clk/SM) - the 3x3 weight gradients ran at 40-59 % of
// the sustained tensor rate, ingest-bound.  Here ONE TMA box of (PW+2) x (PH+2) input pixels x 64 channels per stage serves all nine
// taps: tap (r, s) of 16-pixel K-step kk is the same shared-memory tile entered at pixel row (kk + r) * (PW + 2) + s - the
// descriptor start address may sit anywhere inside the 8-row swizzle atom because the tensor core applies the 128B swizzle to
// absolute shared-memory address bits (known-answer test profiles/desc_offset_test.cu), and the leading-byte offset that stacks the
// second tap of a pair into rows 64-127 of the M = 128 instruction is simply the byte distance between the two taps' entry points.
// One CTA = one 64-channel input block x one 64-channel output tile, all nine taps (5 accumulators of 64 columns), a 64-pixel chunk
// per stage (4 K-steps): 22 KB per 20 MMAs = 18 B/clk.  N = 64 caps an MMA at 64 clk (half the tensor peak), still above what
// the ingest-bound N = 128 form reached.
constexpr int WH_STAGES = 6;
constexpr int WH_XBYTES = 14336;       // >= (PW+2)*(PH+2)*128: 18 x 6 (PW = 16, PH = 4) or 10 x 10 (PW = 8, PH = 8) pixel rows, 1024-aligned
constexpr int WH_GBYTES = 8192;        // 64 pixels x 64 channels
constexpr int WH_STAGE = WH_XBYTES + WH_GBYTES;
constexpr int WH_SMEM = WH_STAGES * WH_STAGE + 1024 + 256;

template <int PW>
__global__ void __launch_bounds__(WG_THREADS, 1)
conv_wgrad_halo_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_g, const WgParams p) {
    constexpr int PH = 64 / PW, BW = PW + 2, STAGES = WH_STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * WH_STAGE);
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES), tfull = smem_u32(bars + 2 * STAGES);
    const int cb = blockIdx.y, k0 = blockIdx.z * 64;
    const int chunk_begin = blockIdx.x * p.chunks_per_split;
    const int chunk_end = min(p.total_chunks, chunk_begin + p.chunks_per_split);
    const int nchunks = chunk_end - chunk_begin;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_g) : "memory");
        for (int i = 0; i < STAGES; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
        mbar_init(tfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_ptr_smem), 512);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (nchunks > 0) {
        if (warp == 0) {
            if (elect_one()) {
                int stage = 0; uint32_t phase = 0;
                constexpr uint32_t tx_bytes = (uint32_t)(BW * (PH + 2) * 128 + WH_GBYTES);
                int cx = chunk_begin % p.chunks_x, cy = (chunk_begin / p.chunks_x) % p.chunks_y, n = chunk_begin / (p.chunks_x * p.chunks_y);
                for (int ck = chunk_begin; ck < chunk_end; ++ck) {
                    const int x0 = cx * PW, y0 = cy * PH;
                    const uint32_t st = smem_u32(smem + stage * WH_STAGE), fb = full0 + 8 * stage;
                    mbar_wait(empty0 + 8 * stage, phase ^ 1);
                    mbar_expect_tx(fb, tx_bytes);
                    tma_load_4d(st, &map_x, fb, cb * 64, x0 - p.pad, y0 - p.pad, n);
                    tma_load_4d(st + WH_XBYTES, &map_g, fb, k0, x0, y0, n);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    if (++cx == p.chunks_x) { cx = 0; if (++cy == p.chunks_y) { cy = 0; ++n; } }
                }
            }
        } else if (warp == 1) {
            const uint32_t idesc = instr_desc_bf16(64, true, true);
            const uint32_t smem0 = smem_u32(smem);
            // pixel-row offset of tap t = (r, s) inside the box; a K-step is one 16-pixel image row (PW = 16: SBO 1024) or two
            // 8-pixel rows (PW = 8: the second 8-row group starts one box row = BW * 128 B further)
            constexpr uint32_t SBO = PW == 16 ? 1024 : BW * 128;
            constexpr int KROWS = PW == 16 ? BW : 2 * BW;            // box rows (pixels) advanced per K-step
            uint64_t adesc0[5];
#pragma unroll
            for (int g = 0; g < 5; ++g) {
                const int ta = 2 * g, tb = 2 * g + 1 < 9 ? 2 * g + 1 : 2 * g;
                const int ra = (ta / 3) * BW + ta % 3, rb = (tb / 3) * BW + tb % 3;
                adesc0[g] = smem_desc(smem0 + ra * 128, (uint32_t)(rb - ra) * 128, SBO, 2);
            }
            const uint64_t bdesc0 = smem_desc(smem0 + WH_XBYTES, 0, 1024, 2);
            if (elect_one()) {
                int stage = 0; uint32_t phase = 0;
                bool stage_ready = false;
                for (int it = 0; it < nchunks; ++it) {
                    if (!stage_ready) mbar_wait(full0 + 8 * stage, phase);
                    fence_after();
                    const int nstage = stage + 1 == STAGES ? 0 : stage + 1;
                    const uint32_t nphase = stage + 1 == STAGES ? phase ^ 1 : phase;
                    const uint64_t soff = (uint64_t)((stage * WH_STAGE) >> 4);
                    uint32_t probe = 0;
#pragma unroll
                    for (int g = 0; g < 5; ++g) {
                        if (g == 4 && it + 1 < nchunks) probe = mbar_try_wait(full0 + 8 * nstage, nphase) ? 1u : 0u;
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            umma_bf16(tmem_base + g * 64, adesc0[g] + soff + (uint64_t)(kk * KROWS * 8), bdesc0 + soff + (uint64_t)(kk * 128),
                                      idesc, (it | kk) != 0);
                    }
                    umma_commit(empty0 + 8 * stage);
                    if (it == nchunks - 1) umma_commit(tfull);
                    stage_ready = probe != 0;
                    stage = nstage; phase = nphase;
                }
            }
            __syncwarp();
        } else {
            // epilogue: TMEM -> red.global.add.v4.f32 into dwp[tap][c][k]
            const int quarter = warp & 3;
            const int m = quarter * 32 + lane;
            mbar_wait(tfull, 0);
            fence_after();
            // all CTAs of a layer add into the SAME few hundred KB at the same moment: each starts at a different accumulator so that
            // the L2 atomic units do not serialise on one address
            for (int gi = 0; gi < 5; ++gi) {
                const int g = (gi + (int)blockIdx.x) % 5;
                const int tap = 2 * g + (m >> 6);
                const bool row_ok = tap < 9 && cb * 64 + (m & 63) < p.Ci_pad;
                float* drow = p.dwp + ((size_t)(tap < 9 ? tap : 8) * p.Ci_pad + cb * 64 + (m & 63)) * p.Co + k0;
#pragma unroll 1
                for (int ch = 0; ch < 2; ++ch) {
                    float v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + g * 64 + ch * 32, v);
                    if (row_ok) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            if (k0 + ch * 32 + j < p.Co)
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(drow + ch * 32 + j), "f"(v[j]),
                                             "f"(v[j + 1]), "f"(v[j + 2]), "f"(v[j + 3]) : "memory");
                        }
                    }
                }
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 1) { fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---- the same idea for layers with >= 128 output channels: N = 128 instructions (full tensor rate).  512 TMEM columns hold only
// four such accumulators, so a CTA takes ONE kernel row (3 taps) of a PAIR of 64-channel input blocks - the pair is stacked into
// M = 128 through the leading-byte offset (= the distance between the two blocks' boxes) - and one 128-channel output tile.  Per
// 64-pixel stage: two (PW+2) x PH boxes (horizontal halo only) + the 64 x 128 gradient tile = 35-37 KB per 12 MMAs of 64 clk
// = 46 B/clk, against 80 B/clk of the per-tap loads above.
constexpr int WR_STAGES = 5;
constexpr int WR_XBYTES = 10240;       // >= (PW+2)*PH*128: 18 x 4 (PW = 16) or 10 x 8 (PW = 8) pixel rows, 1024-aligned
constexpr int WR_GBYTES = 16384;       // 64 pixels x 128 channels (two 64-channel tiles)
constexpr int WR_STAGE = 2 * WR_XBYTES + WR_GBYTES;
constexpr int WR_SMEM = WR_STAGES * WR_STAGE + 1024 + 256;

template <int PW>
__global__ void __launch_bounds__(WG_THREADS, 1)
conv_wgrad_halo_rows_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_g, const WgParams p) {
    constexpr int PH = 64 / PW, BW = PW + 2, STAGES = WR_STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * WR_STAGE);
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES), tfull = smem_u32(bars + 2 * STAGES);
    const int pair = blockIdx.y / 3, r = blockIdx.y - pair * 3, k0 = blockIdx.z * 128;
    const int nblk = min(2, p.cblks - 2 * pair);                 // 64-channel input blocks of this pair that exist
    const int chunk_begin = blockIdx.x * p.chunks_per_split;
    const int chunk_end = min(p.total_chunks, chunk_begin + p.chunks_per_split);
    const int nchunks = chunk_end - chunk_begin;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_g) : "memory");
        for (int i = 0; i < STAGES; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
        mbar_init(tfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_ptr_smem), 512);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (nchunks > 0) {
        if (warp == 0) {
            if (elect_one()) {
                int stage = 0; uint32_t phase = 0;
                const uint32_t tx_bytes = (uint32_t)(nblk * BW * PH * 128 + WR_GBYTES);
                int cx = chunk_begin % p.chunks_x, cy = (chunk_begin / p.chunks_x) % p.chunks_y, n = chunk_begin / (p.chunks_x * p.chunks_y);
                for (int ck = chunk_begin; ck < chunk_end; ++ck) {
                    const int x0 = cx * PW, y0 = cy * PH;
                    const uint32_t st = smem_u32(smem + stage * WR_STAGE), fb = full0 + 8 * stage;
                    mbar_wait(empty0 + 8 * stage, phase ^ 1);
                    mbar_expect_tx(fb, tx_bytes);
                    for (int i = 0; i < nblk; ++i) tma_load_4d(st + i * WR_XBYTES, &map_x, fb, (2 * pair + i) * 64, x0 - p.pad, y0 + r - p.pad, n);
                    tma_load_4d(st + 2 * WR_XBYTES, &map_g, fb, k0, x0, y0, n);
                    tma_load_4d(st + 2 * WR_XBYTES + WR_GBYTES / 2, &map_g, fb, k0 + 64, x0, y0, n);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    if (++cx == p.chunks_x) { cx = 0; if (++cy == p.chunks_y) { cy = 0; ++n; } }
                }
            }
        } else if (warp == 1) {
            const uint32_t idesc = instr_desc_bf16(128, true, true);
            const uint32_t smem0 = smem_u32(smem);
            constexpr uint32_t SBO = PW == 16 ? 1024 : BW * 128;
            constexpr int KROWS = PW == 16 ? BW : 2 * BW;
            uint64_t adesc0[3];
#pragma unroll
            for (int s = 0; s < 3; ++s) adesc0[s] = smem_desc(smem0 + s * 128, nblk == 2 ? WR_XBYTES : 0, SBO, 2);
            const uint64_t bdesc0 = smem_desc(smem0 + 2 * WR_XBYTES, WR_GBYTES / 2, 1024, 2);
            if (elect_one()) {
                int stage = 0; uint32_t phase = 0;
                bool stage_ready = false;
                for (int it = 0; it < nchunks; ++it) {
                    if (!stage_ready) mbar_wait(full0 + 8 * stage, phase);
                    fence_after();
                    const int nstage = stage + 1 == STAGES ? 0 : stage + 1;
                    const uint32_t nphase = stage + 1 == STAGES ? phase ^ 1 : phase;
                    const uint64_t soff = (uint64_t)((stage * WR_STAGE) >> 4);
                    uint32_t probe = 0;
#pragma unroll
                    for (int s = 0; s < 3; ++s) {
                        if (s == 2 && it + 1 < nchunks) probe = mbar_try_wait(full0 + 8 * nstage, nphase) ? 1u : 0u;
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            umma_bf16(tmem_base + s * 128, adesc0[s] + soff + (uint64_t)(kk * KROWS * 8), bdesc0 + soff + (uint64_t)(kk * 128),
                                      idesc, (it | kk) != 0);
                    }
                    umma_commit(empty0 + 8 * stage);
                    if (it == nchunks - 1) umma_commit(tfull);
                    stage_ready = probe != 0;
                    stage = nstage; phase = nphase;
                }
            }
            __syncwarp();
        } else {
            const int quarter = warp & 3;
            const int m = quarter * 32 + lane;
            mbar_wait(tfull, 0);
            fence_after();
            for (int si = 0; si < 3; ++si) {
                const int s = (si + (int)blockIdx.x) % 3;                   // staggered start, see the kernel above
                const int tap = r * 3 + s, c = (2 * pair + (m >> 6)) * 64 + (m & 63);
                const bool row_ok = (m >> 6) < nblk;
                float* drow = p.dwp + ((size_t)tap * p.Ci_pad + (row_ok ? c : 0)) * p.Co + k0;
#pragma unroll 1
                for (int ch = 0; ch < 4; ++ch) {
                    float v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + s * 128 + ch * 32, v);
                    if (row_ok) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            if (k0 + ch * 32 + j < p.Co)
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(drow + ch * 32 + j), "f"(v[j]),
                                             "f"(v[j + 1]), "f"(v[j + 2]), "f"(v[j + 3]) : "memory");
                        }
                    }
                }
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 1) { fence_after(); tmem_dealloc(tmem_base, 512); }
}

static bool wgrad_halo_pref() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SALT_WGRAD_HALO"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}
// 3x3 stride-1 weight gradients with <= 64 output channels whose map tiles into 16 x 4 or 8 x 8 pixel chunks.  Measured per layer
// group (profiles/r2_notes.md): final.0 796 -> 1045 TFLOP/s, layer1 553 -> 639, dec1 399 -> 480, dec2 656 -> 749; layers with >= 128
// output channels LOSE 5-10 % against the N = 128 kernel above (N = 64 instructions cap at half the tensor rate), so they stay there.
static bool wgrad_halo_rows_pref() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SALT_WGRAD_HALO_ROWS"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}
static bool wgrad_halo_ok(const ConvGeom& g) {
    if (!wgrad_halo_pref() || g.R != 3 || g.S != 3 || g.stride != 1 || g.Ci % 8 || g.Co % 8) return false;
    if (g.Co > 64 && (!wgrad_halo_rows_pref() || g.Ci < 128)) return false;
    if (g.Wo >= 16) return g.Wo % 16 == 0 && g.Ho % 4 == 0;
    return g.Wo == 8 && g.Ho % 8 == 0;
}
template <int PW>
static void launch_wg_halo_rows(cudaStream_t st, const void* in, const void* gout, float* dwp, const ConvGeom& g) {
    constexpr int PH = 64 / PW;
    WgParams p;
    p.B = g.B; p.Ho = g.Ho; p.Wo = g.Wo; p.Ci_pad = cdiv(g.Ci, 64) * 64; p.Co = g.Co;
    p.RS = 9; p.S = 3; p.stride = 1; p.pad = g.pad;
    p.cblks = cdiv(g.Ci, 64); p.total_slots = 9 * p.cblks; p.slots_per_cta = 6;
    p.pw = PW; p.ph = PH; p.chunks_x = g.Wo / PW; p.chunks_y = g.Ho / PH;
    p.total_chunks = g.B * p.chunks_x * p.chunks_y;
    const int pairs = cdiv(p.cblks, 2), k_tiles = cdiv(g.Co, 128);
    int splits = num_sms() / (pairs * 3 * k_tiles);
    const int max_splits = cdiv(p.total_chunks, 4);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    p.chunks_per_split = cdiv(p.total_chunks, splits);
    splits = cdiv(p.total_chunks, p.chunks_per_split);
    p.dwp = dwp;
    CUtensorMap mx = make_map_nhwc(in, g.Ci, g.Wi, g.Hi, g.B, 64, PW + 2, PH, 1, 1, CU_TENSOR_MAP_SWIZZLE_128B);
    CUtensorMap mg = make_map_nhwc(gout, g.Co, g.Wo, g.Ho, g.B, 64, PW, PH, 1, 1, CU_TENSOR_MAP_SWIZZLE_128B);
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(conv_wgrad_halo_rows_kernel<PW>, cudaFuncAttributeMaxDynamicSharedMemorySize, WR_SMEM);
        configured = true;
    }
    conv_wgrad_halo_rows_kernel<PW><<<dim3(splits, pairs * 3, k_tiles), WG_THREADS, WR_SMEM, st>>>(mx, mg, p);
}
template <int PW>
static void launch_wg_halo(cudaStream_t st, const void* in, const void* gout, float* dwp, const ConvGeom& g) {
    constexpr int PH = 64 / PW;
    WgParams p;
    p.B = g.B; p.Ho = g.Ho; p.Wo = g.Wo; p.Ci_pad = cdiv(g.Ci, 64) * 64; p.Co = g.Co;
    p.RS = 9; p.S = 3; p.stride = 1; p.pad = g.pad;
    p.cblks = cdiv(g.Ci, 64); p.total_slots = 9 * p.cblks; p.slots_per_cta = 9;
    p.pw = PW; p.ph = PH; p.chunks_x = g.Wo / PW; p.chunks_y = g.Ho / PH;
    p.total_chunks = g.B * p.chunks_x * p.chunks_y;
    const int k_tiles = cdiv(g.Co, 64);
    int splits = num_sms() / (p.cblks * k_tiles);          // one wave of CTAs, as above
    const int max_splits = cdiv(p.total_chunks, 4);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    p.chunks_per_split = cdiv(p.total_chunks, splits);
    splits = cdiv(p.total_chunks, p.chunks_per_split);
    p.dwp = dwp;
    CUtensorMap mx = make_map_nhwc(in, g.Ci, g.Wi, g.Hi, g.B, 64, PW + 2, PH + 2, 1, 1, CU_TENSOR_MAP_SWIZZLE_128B);
    CUtensorMap mg = make_map_nhwc(gout, g.Co, g.Wo, g.Ho, g.B, 64, PW, PH, 1, 1, CU_TENSOR_MAP_SWIZZLE_128B);
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(conv_wgrad_halo_kernel<PW>, cudaFuncAttributeMaxDynamicSharedMemorySize, WH_SMEM);
        configured = true;
    }
    conv_wgrad_halo_kernel<PW><<<dim3(splits, p.cblks, k_tiles), WG_THREADS, WH_SMEM, st>>>(mx, mg, p);
}

bool tc_wgrad_supported(const ConvGeom& g) {
    if (g.Wo < 8 || g.Ho < 4) return false;
    const int pw = g.Wo >= 16 ? 16 : 8, ph = 32 / pw;
    if (g.Wo % pw || g.Ho % ph) return false;
    if (g.Ci % 8 || g.Co % 8) return false;
    if (g.stride != 1 && g.stride != 2) return false;
    return true;
}

template <int NK>
static void launch_wg(cudaStream_t st, const CUtensorMap& mx, const CUtensorMap& mg, const WgParams& p, dim3 grid) {
    using Cfg = WgCfg<NK>;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(conv_wgrad_tc_kernel<NK>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        configured = true;
    }
    conv_wgrad_tc_kernel<NK><<<grid, WG_THREADS, Cfg::SMEM_BYTES, st>>>(mx, mg, p);
}

// dw[k][c][t] (+)= dwp[t][c][k]: 32(k) x 32(c) x RS tile transposed through shared memory
__global__ void __launch_bounds__(256) unpack_dw_kernel(const float* __restrict__ dwp, float* __restrict__ dw, int Co, int Ci_real,
                                                        int Ci_pad, int RS, int accumulate) {
    extern __shared__ float tile[];                  // [32 k][32 c * RS + 1]
    const int k0 = blockIdx.x * 32, c0 = blockIdx.y * 32, pitch = 32 * RS + 1;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int t = 0; t < RS; ++t)
        for (int c = ty; c < 32; c += 8) {
            float v = 0.f;
            if (k0 + tx < Co && c0 + c < Ci_pad) v = dwp[((size_t)t * Ci_pad + c0 + c) * Co + k0 + tx];
            tile[tx * pitch + c * RS + t] = v;
        }
    __syncthreads();
    const int ncols = min(32, Ci_real - c0) * RS;    // contiguous run in dw for one k
    for (int k = ty; k < 32; k += 8) {
        if (k0 + k >= Co) continue;
        float* o = dw + ((size_t)(k0 + k) * Ci_real + c0) * RS;
        for (int j = tx; j < ncols; j += 32) o[j] = (accumulate ? o[j] : 0.f) + tile[k * pitch + j];
    }
}
void k_unpack_dw(cudaStream_t st, const float* dwp, float* dw, int Co, int Ci_real, int Ci_pad, int RS, bool accumulate) {
    SALT_COUNT(1);
    dim3 grid(cdiv(Co, 32), cdiv(Ci_real, 32));
    unpack_dw_kernel<<<grid, 256, sizeof(float) * 32 * (32 * RS + 1), st>>>(dwp, dw, Co, Ci_real, Ci_pad, RS, accumulate ? 1 : 0);
}
size_t tc_wgrad_scratch_floats(int Ci, int Co, int RS) { return (size_t)RS * (cdiv(Ci, 64) * 64) * Co; }

// in: [B,Hi,Wi,Ci] bf16 (physical dims); gout: [B,Ho,Wo,Co] bf16; dwp: zeroed fp32 scratch [RS][ceil64(Ci)][Co], accumulated into
void k_conv_wgrad_tc(cudaStream_t st, const void* in, const void* gout, float* dwp, const ConvGeom& g) {
    SALT_COUNT(1);
    if (wgrad_halo_ok(g)) {
        if (g.Co > 64) {
            if (g.Wo >= 16) launch_wg_halo_rows<16>(st, in, gout, dwp, g);
            else launch_wg_halo_rows<8>(st, in, gout, dwp, g);
        } else {
            if (g.Wo >= 16) launch_wg_halo<16>(st, in, gout, dwp, g);
            else launch_wg_halo<8>(st, in, gout, dwp, g);
        }
        return;
    }
    WgParams p;
    p.B = g.B; p.Ho = g.Ho; p.Wo = g.Wo; p.Ci_pad = cdiv(g.Ci, 64) * 64; p.Co = g.Co;
    p.RS = g.R * g.S; p.S = g.S; p.stride = g.stride; p.pad = g.pad;
    p.cblks = cdiv(g.Ci, 64);
    p.total_slots = p.RS * p.cblks;
    const int NK = g.Co > 64 ? 128 : 64;
    const int max_slots = NK == 64 ? 10 : 8;
    int slot_chunks = cdiv(p.total_slots, max_slots);
    p.slots_per_cta = cdiv(cdiv(p.total_slots, slot_chunks), 2) * 2;
    slot_chunks = cdiv(p.total_slots, p.slots_per_cta);
    p.pw = g.Wo >= 16 ? 16 : 8; p.ph = 32 / p.pw;
    p.chunks_x = g.Wo / p.pw; p.chunks_y = g.Ho / p.ph;
    p.total_chunks = g.B * p.chunks_x * p.chunks_y;
    const int k_tiles = cdiv(g.Co, NK);
    int splits = num_sms() / (slot_chunks * k_tiles);      // at most ONE wave of CTAs (1 CTA/SM): a few CTAs over would double the time
    int max_splits = cdiv(p.total_chunks, 8);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    p.chunks_per_split = cdiv(p.total_chunks, splits);
    splits = cdiv(p.total_chunks, p.chunks_per_split);
    p.dwp = dwp;
    CUtensorMap mx = make_map_nhwc(in, g.Ci, g.Wi, g.Hi, g.B, 64, p.pw, p.ph, 1, g.stride, CU_TENSOR_MAP_SWIZZLE_128B);
    CUtensorMap mg = make_map_nhwc(gout, g.Co, g.Wo, g.Ho, g.B, 64, p.pw, p.ph, 1, 1, CU_TENSOR_MAP_SWIZZLE_128B);
    dim3 grid(splits, slot_chunks, k_tiles);
    if (NK == 64) launch_wg<64>(st, mx, mg, p, grid);
    else launch_wg<128>(st, mx, mg, p, grid);
}

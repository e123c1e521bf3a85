// Stream ring: the pipeline every HBM-bound elementwise / reduction pass of the engine runs on.
//
// The passes read one to three equally shaped NHWC tensors front to back.  With plain global loads the bytes in flight per SM are
// (resident threads) x (loads a thread can keep in registers); the register-heavy passes (five per-channel coefficient vectors,
// reduction accumulators) sat at 2 blocks per SM and 16-32 KB in flight, i.e. 2.4-3.5 TB/s of the 6.4 TB/s copy peak
// (profiles/r2_notes.md section 4).  Here ONE elected thread of a producer warp streams CHUNK-byte pieces of every input into a
// shared-memory ring with cp.async.bulk (completion on an mbarrier), so 96-128 KB per SM are in flight whatever the consumers'
// register budget is; 8 consumer warps read 16-byte vectors from the ring (conflict-free: consecutive threads, consecutive
// vectors), keep their per-channel coefficients in registers and store results straight to global memory.
//
// Geometry: a chunk is a whole number of pixels and 256 threads x 16 bytes is a whole number of pixels (row bytes = C * sizeof(T)
// is a power of two <= 4096, and so is the consumers' pass of 4 or 8 KB), so a consumer thread serves ONE channel group for the whole kernel.  Grid = min(#SMs, chunks)
// persistent CTAs, chunk i of CTA b = b + i * grid.  Reductions end in the block's own partial slot (fixed-order finalize).
#pragma once
#include "elem_vec.cuh"

namespace ring {

constexpr int CHUNK = 16384;                 // bytes per input per stage
// Op::WARPS consumer warps (8 or 16) + the producer warp.  The consumers need ~70-190 instructions per 16-byte vector (unpack,
// fp32 arithmetic, pack, addressing); with 8 warps - two per scheduler - they issue ~1.5 instructions per clock and top out near
// 4 TB/s (ncu: issue slots 43-52 % busy, 14 % of the warp slots occupied).  Ops that fit 120 registers run 16 warps.
template <int NS> struct Cfg {
    static constexpr int STAGES = NS == 1 ? 6 : (NS == 2 ? 4 : 3);
    static constexpr int RING_BYTES = STAGES * NS * CHUNK;
    static constexpr int SMEM = RING_BYTES + 128;
};
// Stream 0 is the primary tensor (nbytes, byte offsets `off` refer to it).  A secondary stream may be narrower per pixel
// (shift[i]: its bytes = primary bytes >> shift[i]) and may be laid out per image with its own image stride (planes of an NCHW
// tensor): source offset = (off / img_bytes) * img_stride[i] + ((off % img_bytes) >> shift[i]); img_stride[i] == 0 = flat.
// group_chunks > 0 (per-image reductions): the grid is (images x slices) CTAs, CTA b walks the chunks slice, slice + slices, ...
// of image b / slices only (an image is a whole number of chunks), so that its partial result belongs to ONE image.
template <int NS> struct Streams {
    const uint8_t* p[NS]; size_t nbytes;
    int shift[NS] = {}; size_t img_stride[NS] = {}; size_t img_bytes = 0;
    int group_chunks = 0, slices = 1;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
// bounded: a mis-programmed pipeline traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0, spins = 0;
    while (true) {
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
            "selp.u32 %0, 1, 0, P1;\n"
            "}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) break;
        if (++spins > (1u << 24)) { printf("ring: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
    }
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
template <int NT> __device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }

// Op interface (all members __device__):
//   void begin(int tid)                                   per-thread set-up (coefficients of the thread's channel group)
//   static constexpr bool FULL_WARPS
//   void vec(size_t off, const uint4 (&in)[NS])           FULL_WARPS = false: one 16-byte vector of every input at byte offset
//                                                         `off` of the tensors
//   void vecs(size_t off, const uint8_t* stage, const int (&o)[U], const bool (&valid)[U])
//                                                         FULL_WARPS = true (ops with warp shuffles): U vectors per call, the op
//                                                         reads its inputs from the stage itself (stream i at stage + i*CHUNK +
//                                                         o[u]; tensor byte offset off + o[u])
//   void end(int tid, float* red)                         after the last chunk; `red` = 8 floats per consumer thread of shared memory,
//                                                         consumers synchronise with consumer_sync<NT>() / the NAMED block sums
template <int NS, class Op>
__global__ void __launch_bounds__(Op::WARPS * 32 + 32, 1) ring_kernel(Streams<NS> s, Op op) {
    constexpr int CONSUMERS = Op::WARPS * 32;
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int STAGES = Cfg<NS>::STAGES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg<NS>::RING_BYTES);
    const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * STAGES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, CONSUMERS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t nchunks = (s.nbytes + CHUNK - 1) / CHUNK;
    size_t ch_first = blockIdx.x, ch_last = nchunks, ch_step = gridDim.x;
    if (s.group_chunks > 0) {
        const size_t g = blockIdx.x / s.slices;
        ch_first = g * s.group_chunks + (blockIdx.x - g * s.slices); ch_last = (g + 1) * s.group_chunks; ch_step = s.slices;
    }
    if (warp == CONSUMERS / 32) {
        if (lane == 0) {
            int st = 0; uint32_t ph = 0;
            for (size_t ch = ch_first; ch < ch_last; ch += ch_step) {
                mbar_wait(empty0 + 8 * st, ph ^ 1);
                const size_t off = ch * CHUNK;
                const uint32_t bytes = (uint32_t)min((size_t)CHUNK, s.nbytes - off);
                uint32_t total = 0;
#pragma unroll
                for (int i = 0; i < NS; ++i) total += bytes >> s.shift[i];
                mbar_expect_tx(full0 + 8 * st, total);
#pragma unroll
                for (int i = 0; i < NS; ++i) {
                    const size_t so = s.img_stride[i] ? (off / s.img_bytes) * s.img_stride[i] + ((off % s.img_bytes) >> s.shift[i])
                                                      : (off >> s.shift[i]);
                    bulk_load(smem_u32(smem + (size_t)(st * NS + i) * CHUNK), s.p[i] + so, bytes >> s.shift[i], full0 + 8 * st);
                }
                if (++st == STAGES) { st = 0; ph ^= 1; }
            }
        }
        return;
    }
    op.begin(threadIdx.x);
    int st = 0; uint32_t ph = 0;
    for (size_t ch = ch_first; ch < ch_last; ch += ch_step) {
        const size_t off = ch * CHUNK;
        const int bytes = (int)min((size_t)CHUNK, s.nbytes - off);
        mbar_wait(full0 + 8 * st, ph);
        const uint8_t* base = smem + (size_t)st * NS * CHUNK;
        if constexpr (Op::FULL_WARPS) {
            // ops with warp shuffles: every thread runs every pass of a (possibly short) chunk, `valid` masks the tail; the op takes
            // Op::U vectors (U consecutive passes) at once so that their shuffle / special-function chains overlap
            constexpr int U = Op::U;
            for (int o0 = threadIdx.x * 16; o0 - (int)threadIdx.x * 16 < bytes; o0 += U * CONSUMERS * 16) {
                int oc[U]; bool valid[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int o = o0 + u * CONSUMERS * 16;
                    valid[u] = o < bytes;
                    oc[u] = valid[u] ? o : (int)threadIdx.x * 16;
                }
                op.vecs(off, base, oc, valid);
            }
        } else {
#pragma unroll 2
            for (int o = threadIdx.x * 16; o < bytes; o += CONSUMERS * 16) {
                uint4 in[NS];
#pragma unroll
                for (int i = 0; i < NS; ++i) in[i] = *reinterpret_cast<const uint4*>(base + (size_t)i * CHUNK + o);
                op.vec(off + o, in);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * st);
        if (++st == STAGES) { st = 0; ph ^= 1; }
    }
    consumer_sync<CONSUMERS>();          // every stage has been consumed: the ring is free to serve as reduction scratch
    op.end(threadIdx.x, reinterpret_cast<float*>(smem));
}

static inline int num_sms_cached() {
    static int n = 0;
    if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); }
    return n;
}
// a tensor can ride the ring if its row (C elements) is a power of two of 16..4096 bytes and it has no border
static inline bool row_ok(const Tensor& t) {
    const size_t rb = (size_t)t.C * dtype_size(t.dt);
    return t.pt == 0 && t.pb == 0 && t.pl == 0 && t.pr == 0 && rb >= 16 && rb <= 4096 && (rb & (rb - 1)) == 0 &&
           (reinterpret_cast<uintptr_t>(t.p) & 15) == 0;
}
template <int NS, class Op>
static void launch(cudaStream_t st, const Streams<NS>& s, const Op& op) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(ring_kernel<NS, Op>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<NS>::SMEM);
        attr_set = true;
    }
    const size_t nchunks = (s.nbytes + CHUNK - 1) / CHUNK;
    const int grid = s.group_chunks > 0 ? (int)(nchunks / s.group_chunks) * s.slices
                                        : (int)std::min<size_t>((size_t)num_sms_cached(), nchunks);
    ring_kernel<NS, Op><<<grid, Op::WARPS * 32 + 32, Cfg<NS>::SMEM, st>>>(s, op);
}
static inline int grid_for(size_t nbytes) {
    const size_t nchunks = (nbytes + CHUNK - 1) / CHUNK;
    return (int)std::min<size_t>((size_t)num_sms_cached(), nchunks);
}

}  // namespace ring

// 16-byte channel vectors, per-channel parameter loads and the fixed-order block reductions shared by the elementwise kernels
// (kernels_elem.cu: direct global loads; kernels_stream.cu: the same passes fed through a cp.async.bulk shared-memory ring).
#pragma once
#include "kernels.h"
#define EW_THREADS 256

// ------------------------------------------------------------------------------------------------
// 16-byte channel vectors
// ------------------------------------------------------------------------------------------------
// n / d for n < 2^31 by multiply-shift (the hardware has no integer divide: a 32-bit division is ~30 instructions through the
// special-function unit, and the ring consumers are instruction-bound - profiles/r2_notes.md)
struct FDiv {
    uint32_t d, mul, shr;
    __device__ __forceinline__ uint32_t div(uint32_t n) const { return d == 1 ? n : __umulhi(n, mul) >> shr; }
};
static inline FDiv make_fdiv(int d) {
    FDiv f; f.d = (uint32_t)d; f.mul = 0; f.shr = 0;
    if (d > 1) {
        int lg = 0;
        while ((1u << lg) < (uint32_t)d) ++lg;
        const unsigned p = 31 + lg;
        f.mul = (uint32_t)(((1ull << p) + (uint32_t)d - 1) / (uint32_t)d);
        f.shr = p - 32;
    }
    return f;
}
template <typename T> struct VW;
template <> struct VW<float> { static constexpr int N = 4; };
template <> struct VW<bf16> { static constexpr int N = 8; };
template <int N> struct Vf { float v[N]; };

__device__ __forceinline__ Vf<4> ldv(const float* p) {
    float4 a = *reinterpret_cast<const float4*>(p);
    Vf<4> r; r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    return r;
}
__device__ __forceinline__ Vf<8> ldv(const bf16* p) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
    Vf<8> r;
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); r.v[2 * i] = f.x; r.v[2 * i + 1] = f.y; }
    return r;
}
__device__ __forceinline__ void stv(float* p, const Vf<4>& a) { *reinterpret_cast<float4*>(p) = make_float4(a.v[0], a.v[1], a.v[2], a.v[3]); }
__device__ __forceinline__ void stv(bf16* p, const Vf<8>& a) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(a.v[2 * i], a.v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = u;
}
// streaming (evict-first) stores for write-once outputs that must not push re-used inputs out of L2
__device__ __forceinline__ void stv_stream(float* p, const Vf<4>& a) { __stcs(reinterpret_cast<float4*>(p), make_float4(a.v[0], a.v[1], a.v[2], a.v[3])); }
__device__ __forceinline__ void stv_stream(bf16* p, const Vf<8>& a) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(a.v[2 * i], a.v[2 * i + 1]);
    __stcs(reinterpret_cast<uint4*>(p), u);
}
// the same conversions for a 16-byte vector that already sits in registers (stream-ring kernels read it from shared memory)
template <typename T> __device__ __forceinline__ Vf<VW<T>::N> vfrom(const uint4& u);
template <> __device__ __forceinline__ Vf<4> vfrom<float>(const uint4& u) {
    Vf<4> r; r.v[0] = __uint_as_float(u.x); r.v[1] = __uint_as_float(u.y); r.v[2] = __uint_as_float(u.z); r.v[3] = __uint_as_float(u.w);
    return r;
}
template <> __device__ __forceinline__ Vf<8> vfrom<bf16>(const uint4& u) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
    Vf<8> r;
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); r.v[2 * i] = f.x; r.v[2 * i + 1] = f.y; }
    return r;
}
template <typename T> __device__ __forceinline__ uint4 vto(const Vf<VW<T>::N>& a);
template <> __device__ __forceinline__ uint4 vto<float>(const Vf<4>& a) {
    return make_uint4(__float_as_uint(a.v[0]), __float_as_uint(a.v[1]), __float_as_uint(a.v[2]), __float_as_uint(a.v[3]));
}
template <> __device__ __forceinline__ uint4 vto<bf16>(const Vf<8>& a) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(a.v[2 * i], a.v[2 * i + 1]);
    return u;
}
// The per-pixel kernels (scSE, final 1x1) reduce over the channel groups of ONE pixel with warp shuffles (<= 32 lanes); they use
// 8-channel vectors for both storage types so that C <= 256 fits a warp.
__device__ __forceinline__ Vf<8> ldv8(const bf16* p) { return ldv(p); }
__device__ __forceinline__ Vf<8> ldv8(const float* p) {
    float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    Vf<8> r; r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void stv8(bf16* p, const Vf<8>& a) { stv(p, a); }
__device__ __forceinline__ void stv8(float* p, const Vf<8>& a) {
    *reinterpret_cast<float4*>(p) = make_float4(a.v[0], a.v[1], a.v[2], a.v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(a.v[4], a.v[5], a.v[6], a.v[7]);
}
template <int N> __device__ __forceinline__ Vf<N> ldp(const float* p) {      // N fp32 per-channel parameters
    Vf<N> r;
#pragma unroll
    for (int i = 0; i < N; i += 4) {
        float4 a = *reinterpret_cast<const float4*>(p + i);
        r.v[i] = a.x; r.v[i + 1] = a.y; r.v[i + 2] = a.z; r.v[i + 3] = a.w;
    }
    return r;
}
template <int N> __device__ __forceinline__ void stp(float* p, const Vf<N>& a) {
#pragma unroll
    for (int i = 0; i < N; i += 4) *reinterpret_cast<float4*>(p + i) = make_float4(a.v[i], a.v[i + 1], a.v[i + 2], a.v[i + 3]);
}
template <int N> __device__ __forceinline__ Vf<N> vzero() { Vf<N> r;
#pragma unroll
    for (int i = 0; i < N; ++i) r.v[i] = 0.f;
    return r; }
template <int N> __device__ __forceinline__ Vf<N> vfma(const Vf<N>& a, const Vf<N>& b, const Vf<N>& c) { Vf<N> r;
#pragma unroll
    for (int i = 0; i < N; ++i) r.v[i] = fmaf(a.v[i], b.v[i], c.v[i]);
    return r; }
template <int N> __device__ __forceinline__ Vf<N> vadd(const Vf<N>& a, const Vf<N>& b) { Vf<N> r;
#pragma unroll
    for (int i = 0; i < N; ++i) r.v[i] = a.v[i] + b.v[i];
    return r; }
template <int N> __device__ __forceinline__ Vf<N> vmul(const Vf<N>& a, const Vf<N>& b) { Vf<N> r;
#pragma unroll
    for (int i = 0; i < N; ++i) r.v[i] = a.v[i] * b.v[i];
    return r; }
template <int N> __device__ __forceinline__ Vf<N> vaxpy(const Vf<N>& a, float s, const Vf<N>& c) { Vf<N> r;   // a*s + c
#pragma unroll
    for (int i = 0; i < N; ++i) r.v[i] = fmaf(a.v[i], s, c.v[i]);
    return r; }
template <int N> __device__ __forceinline__ Vf<N> vscale(const Vf<N>& a, float s) { Vf<N> r;
#pragma unroll
    for (int i = 0; i < N; ++i) r.v[i] = a.v[i] * s;
    return r; }
template <int N> __device__ __forceinline__ Vf<N> vrelu(const Vf<N>& a) { Vf<N> r;
#pragma unroll
    for (int i = 0; i < N; ++i) r.v[i] = fmaxf(a.v[i], 0.f);
    return r; }
template <int N> __device__ __forceinline__ Vf<N> vmaskpos(const Vf<N>& g, const Vf<N>& m) { Vf<N> r;       // g where m > 0
#pragma unroll
    for (int i = 0; i < N; ++i) r.v[i] = m.v[i] > 0.f ? g.v[i] : 0.f;
    return r; }
template <int N> __device__ __forceinline__ float vdot(const Vf<N>& a, const Vf<N>& b) { float s = 0.f;
#pragma unroll
    for (int i = 0; i < N; ++i) s = fmaf(a.v[i], b.v[i], s);
    return s; }
// xhat = (x - mean) * invstd
template <int N> __device__ __forceinline__ Vf<N> vxhat(const Vf<N>& x, const Vf<N>& mu, const Vf<N>& is) { Vf<N> r;
#pragma unroll
    for (int i = 0; i < N; ++i) r.v[i] = (x.v[i] - mu.v[i]) * is.v[i];
    return r; }

// Block sum of a per-thread vector over the threads that share a channel group (tid % cg; cg a power of two).  Fixed order:
// warp shuffles over the lanes of a warp that share the group (xor offsets cg, 2cg, ...), then the 8 warps through shared memory
// in warp order - no atomics, bit-reproducible, and ~20 instructions instead of the 32-step serial shared-memory loop whose tail
// made every extra block of the reduction passes expensive (profiles/r2_notes.md).  The result is valid in threads tid < cg.
// NAMED: the block carries extra warps that do not take part (the producer warp of the stream-ring kernels): the EW_THREADS
// reducing threads meet on named barrier 1 instead of __syncthreads().
// NT = number of reducing threads (a multiple of 32; the ring kernels run 8 or 16 consumer warps); red holds N * NT floats.
template <bool NAMED, int NT> __device__ __forceinline__ void ew_sync() {
    if constexpr (NAMED) asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
    else __syncthreads();
}
template <int N, bool NAMED = false, int NT = EW_THREADS>
__device__ __forceinline__ void block_sum(Vf<N>& v, int cg, float* red) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = NT / 32;
    if (cg < 32) {
        for (int off = cg; off < 32; off <<= 1) {
#pragma unroll
            for (int i = 0; i < N; ++i) v.v[i] += __shfl_xor_sync(0xffffffffu, v.v[i], off);
        }
        if (lane < cg) {
#pragma unroll
            for (int i = 0; i < N; ++i) red[(i * NW + warp) * 32 + lane] = v.v[i];
        }
        ew_sync<NAMED, NT>();
        if (tid < cg) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                float s = 0.f;
#pragma unroll
                for (int w = 0; w < NW; ++w) s += red[(i * NW + w) * 32 + tid];
                v.v[i] = s;
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) red[i * NT + tid] = v.v[i];
        ew_sync<NAMED, NT>();
        if (tid < cg) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                float s = 0.f;
                for (int t = tid; t < NT; t += cg) s += red[i * NT + t];
                v.v[i] = s;
            }
        }
    }
    ew_sync<NAMED, NT>();
}
// ... added to dst[c..] with atomics (parameter gradients that several blocks contribute to)
template <int N, typename D, bool NAMED = false, int NT = EW_THREADS>
__device__ __forceinline__ void block_reduce_add(Vf<N> v, int cg, D* dst, float* red) {
    block_sum<N, NAMED, NT>(v, cg, red);
    if ((int)threadIdx.x < cg) {
#pragma unroll
        for (int i = 0; i < N; ++i) atomicAdd(dst + threadIdx.x * N + i, (D)v.v[i]);
    }
}
// ... STORED into the block's own partial slot (dst already points at the slot): the finalize kernel adds the slots in a fixed
// order, so the result does not depend on block scheduling.
template <int N, bool NAMED = false, int NT = EW_THREADS>
__device__ __forceinline__ void block_reduce_slot(Vf<N> v, int cg, float* dst, float* red) {
    block_sum<N, NAMED, NT>(v, cg, red);
    if ((int)threadIdx.x < cg) {
#pragma unroll
        for (int i = 0; i < N; ++i) dst[threadIdx.x * N + i] = v.v[i];
    }
}
static inline int reduce_blocks(long long npix, int cg) {
    int lanes = EW_THREADS / cg;
    long long b = (npix + (long long)lanes * 8 - 1) / ((long long)lanes * 8);
    if (b > SALT_STAT_SLOTS_BWD) b = SALT_STAT_SLOTS_BWD;      // one partial slot per block (kernels.h); up to 8 blocks per SM
    if (b < 1) b = 1;
    return (int)b;
}
// sum over the cg (power of two, <= 32) lanes that share one pixel
__device__ __forceinline__ float group_sum(float v, int cg) {
    for (int o = cg >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

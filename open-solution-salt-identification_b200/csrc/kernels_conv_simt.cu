// Generic fp32-accumulate implicit-GEMM convolution on CUDA cores (SIMT).
//
// Role in the engine: (1) the exact-fp32 "parity" precision runs every convolution here, (2) in bf16
// mode it covers the shapes the tcgen05 kernel does not take (stem C=3, stride-2 dgrad, tiny N) and is
// the on-device reference the tensor-core kernels are checked against.
//
// Tiling: 64 output pixels x 64 output channels per 256-thread CTA, 4x4 register tile per thread,
// reduction chunks of 16 staged through shared memory with register prefetch (double buffering).
#include "kernels.h"

#define TM 64
#define TN 64
#define TK 16
#define SPAD 4

// ------------------------------------------------------------------------------------------------
// weight packing: master fp32 [Co][Ci_real][R][S] -> fwd [Co][R*S][Ci] and dgrad [Ci][R*S][Co] (type T)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void pack_weights_kernel(const float* __restrict__ w, T* __restrict__ wp, T* __restrict__ wpd, int Co,
                                    int Ci_real, int Ci, int RS) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)Co * RS * Ci;
    if (idx >= total) return;
    int c = (int)(idx % Ci);
    int t = (int)((idx / Ci) % RS);
    int k = (int)(idx / ((long long)Ci * RS));
    float v = c < Ci_real ? w[((size_t)k * Ci_real + c) * RS + t] : 0.f;
    st1(wp + idx, v);
    // dgrad copy is stored with the taps FLIPPED (tap' = RS-1-tap), so that dgrad is a plain correlation of the output
    // gradient with wpd (same kernel as forward): gin(y) = sum_{r'} gout(y - (R-1-pad) + r') * wpd[..][r'][..]
    if (wpd) st1(wpd + ((size_t)c * RS + (RS - 1 - t)) * Co + k, v);
}
void k_pack_weights(cudaStream_t st, DType dt, const float* w, void* wp, void* wpd, int Co, int Ci_real, int Ci, int R, int S) {
    SALT_COUNT(1);
    long long total = (long long)Co * R * S * Ci;
    SALT_DISPATCH(dt, T, (pack_weights_kernel<T><<<cdiv(total, 256), 256, 0, st>>>(w, (T*)wp, (T*)wpd, Co, Ci_real, Ci, R * S)));
}

// ------------------------------------------------------------------------------------------------
// forward / dgrad implicit GEMM
//   FWD  : out[m=(n,yo,xo)][k]  = sum_{r,s,c} in [n, yo*st+r-pad, xo*st+s-pad, c] * wp [k][(r,s,c)]
//   DGRAD: gin[m=(n,yi,xi)][c]  = sum_{r,s,k} gout[n, (yi+pad-r)/st, (xi+pad-s)/st, k] * wpd[c][(r,s,k)]
// ------------------------------------------------------------------------------------------------
struct IGemmArgs {
    int M, N, Kred;            // GEMM extents
    int Cred;                  // channels per tap in the reduction (FWD: Ci, DGRAD: Co)
    int Hs, Ws;                // source tensor spatial dims (FWD: Hi,Wi ; DGRAD: Ho,Wo)
    int Hd, Wd;                // destination spatial dims   (FWD: Ho,Wo ; DGRAD: Hi,Wi)
    int R, S, stride, pad;
    int accumulate;
};

template <typename T, bool DGRAD>
__global__ void __launch_bounds__(256) conv_igemm_simt_kernel(const T* __restrict__ src, const T* __restrict__ w,
                                                              const float* __restrict__ bias, T* __restrict__ dst,
                                                              float* __restrict__ stats, IGemmArgs a) {
    __shared__ __align__(16) float As[2][TK][TM + SPAD];
    __shared__ __align__(16) float Bs[2][TK][TN + SPAD];
    __shared__ float red[16][TN];

    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
    // loader mapping: 4 consecutive threads fetch 16 consecutive reduction elements of one row
    const int lrow = tid >> 2, lk = (tid & 3) * 4;
    // A row = destination pixel
    const int am = m0 + lrow;
    const bool am_ok = am < a.M;
    int an = 0, ay = 0, ax = 0;
    if (am_ok) { ax = am % a.Wd; ay = (am / a.Wd) % a.Hd; an = am / (a.Wd * a.Hd); }
    const int bn_ = n0 + lrow;
    const bool bn_ok = bn_ < a.N;

    auto load_a = [&](int kk) -> float4 {
        if (!am_ok || kk >= a.Kred) return f4_zero();
        int tap = kk / a.Cred, c = kk - tap * a.Cred;
        int r = tap / a.S, s = tap - r * a.S;
        int sy, sx;
        if (!DGRAD) {
            sy = ay * a.stride + r - a.pad;
            sx = ax * a.stride + s - a.pad;
        } else {
            // wpd holds flipped taps: reduction tap (r,s) multiplies W[..][R-1-r][S-1-s]
            int ty = ay + a.pad - (a.R - 1 - r), tx = ax + a.pad - (a.S - 1 - s);
            if (ty < 0 || tx < 0) return f4_zero();
            if (a.stride > 1) {
                if ((ty % a.stride) | (tx % a.stride)) return f4_zero();
                ty /= a.stride; tx /= a.stride;
            }
            sy = ty; sx = tx;
        }
        if (sy < 0 || sy >= a.Hs || sx < 0 || sx >= a.Ws) return f4_zero();
        return ld4(src + (((size_t)an * a.Hs + sy) * a.Ws + sx) * a.Cred + c);
    };
    auto load_b = [&](int kk) -> float4 {
        if (!bn_ok || kk >= a.Kred) return f4_zero();
        return ld4(w + (size_t)bn_ * a.Kred + kk);
    };
    auto stage = [&](int buf, float4 va, float4 vb) {
        As[buf][lk + 0][lrow] = va.x; As[buf][lk + 1][lrow] = va.y; As[buf][lk + 2][lrow] = va.z; As[buf][lk + 3][lrow] = va.w;
        Bs[buf][lk + 0][lrow] = vb.x; Bs[buf][lk + 1][lrow] = vb.y; Bs[buf][lk + 2][lrow] = vb.z; Bs[buf][lk + 3][lrow] = vb.w;
    };

    const int tx = tid & 15, ty = tid >> 4;     // thread computes rows ty*4..+3, cols tx*4..+3
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int nk = (a.Kred + TK - 1) / TK;
    float4 va = load_a(lk), vb = load_b(lk);
    stage(0, va, vb);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) { va = load_a((kt + 1) * TK + lk); vb = load_b((kt + 1) * TK + lk); }
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            float4 av = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            float4 bv = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
        if (kt + 1 < nk) stage(buf ^ 1, va, vb);
        __syncthreads();
    }

    // epilogue
    const int cn = n0 + tx * 4;
    float4 bv = f4_zero();
    if (bias && cn < a.N) bv = ld4(bias + cn);
    float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m0 + ty * 4 + i;
        if (m < a.M && cn < a.N) {
            float4 v = make_float4(acc[i][0] + bv.x, acc[i][1] + bv.y, acc[i][2] + bv.z, acc[i][3] + bv.w);
            T* o = dst + (size_t)m * a.N + cn;
            if (a.accumulate) v = f4_add(v, ld4(o));
            st4(o, v);
            s1[0] += v.x; s1[1] += v.y; s1[2] += v.z; s1[3] += v.w;
            s2[0] += v.x * v.x; s2[1] += v.y * v.y; s2[2] += v.z * v.z; s2[3] += v.w * v.w;
        }
    }
    if (stats) {            // per-channel sum / sum of squares for train-mode BatchNorm
        for (int pass = 0; pass < 2; ++pass) {
            float* sv = pass ? s2 : s1;
#pragma unroll
            for (int j = 0; j < 4; ++j) red[ty][tx * 4 + j] = sv[j];
            __syncthreads();
            if (tid < TN && n0 + tid < a.N) {
                float t = 0.f;
#pragma unroll
                for (int i = 0; i < 16; ++i) t += red[i][tid];
                // many M-tiles share a slot: the fp32 parity mode keeps float atomics here (its activations are fp32, so a last-bit
                // difference in a sum is not amplified the way bf16 storage amplifies it)
                atomicAdd(stats + (size_t)(blockIdx.x % SALT_STAT_SLOTS) * 2 * a.N + pass * a.N + n0 + tid, t);
            }
            __syncthreads();
        }
    }
}

void k_conv_fwd_simt(cudaStream_t st, DType dt, const void* in, const void* wp, const float* bias, void* out, float* stats,
                     const ConvGeom& g) {
    SALT_COUNT(1);
    IGemmArgs a;
    a.M = g.B * g.Ho * g.Wo; a.N = g.Co; a.Kred = g.R * g.S * g.Ci; a.Cred = g.Ci;
    a.Hs = g.Hi; a.Ws = g.Wi; a.Hd = g.Ho; a.Wd = g.Wo; a.R = g.R; a.S = g.S; a.stride = g.stride; a.pad = g.pad; a.accumulate = 0;
    dim3 grid(cdiv(a.M, TM), cdiv(a.N, TN));
    SALT_DISPATCH(dt, T, (conv_igemm_simt_kernel<T, false><<<grid, 256, 0, st>>>((const T*)in, (const T*)wp, bias, (T*)out, stats, a)));
}
void k_conv_dgrad_simt(cudaStream_t st, DType dt, const void* gout, const void* wpd, void* gin, bool accumulate, const ConvGeom& g) {
    SALT_COUNT(1);
    IGemmArgs a;
    a.M = g.B * g.Hi * g.Wi; a.N = g.Ci; a.Kred = g.R * g.S * g.Co; a.Cred = g.Co;
    a.Hs = g.Ho; a.Ws = g.Wo; a.Hd = g.Hi; a.Wd = g.Wi; a.R = g.R; a.S = g.S; a.stride = g.stride; a.pad = g.pad; a.accumulate = accumulate ? 1 : 0;
    dim3 grid(cdiv(a.M, TM), cdiv(a.N, TN));
    SALT_DISPATCH(dt, T, (conv_igemm_simt_kernel<T, true><<<grid, 256, 0, st>>>((const T*)gout, (const T*)wpd, nullptr, (T*)gin, nullptr, a)));
}

// ------------------------------------------------------------------------------------------------
// wgrad: dw[k][c][r][s] += sum_{n,yo,xo} gout[n,yo,xo,k] * in[n, yo*st+r-pad, xo*st+s-pad, c]
//   GEMM  M = Co, N = R*S*Ci, reduction = pixels (split across gridDim.z, fp32 atomics)
// ------------------------------------------------------------------------------------------------
struct WGradArgs {
    int P, Co, Ci, Ci_real, NC;     // pixels, channels, NC = R*S*Ci
    int Hi, Wi, Ho, Wo, R, S, stride, pad;
    int chunk;                       // pixels per z-slice (multiple of TK)
    int groups;
};
template <typename T>
__global__ void __launch_bounds__(256) conv_wgrad_simt_kernel(const T* __restrict__ in, const T* __restrict__ gout,
                                                              float* __restrict__ dw, WGradArgs a) {
    __shared__ __align__(16) float As[2][TK][TM + SPAD];
    __shared__ __align__(16) float Bs[2][TK][TN + SPAD];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
    const int p_begin = blockIdx.z * a.chunk, p_end = min(a.P, p_begin + a.chunk);
    const int lp = tid >> 4, lv = (tid & 15) * 4;      // loader: pixel lp (0..15), 4 columns at lv
    // column (tap, c) handled by this thread's B loads is fixed over the whole loop
    const int col = n0 + lv;
    const bool col_ok = col < a.NC;
    int tap = 0, cc = 0, r = 0, s = 0;
    if (col_ok) { tap = col / a.Ci; cc = col - tap * a.Ci; r = tap / a.S; s = tap - r * a.S; }
    const bool arow_ok = (m0 + lv) < a.Co;

    auto load_a = [&](int p) -> float4 {
        if (p >= p_end || !arow_ok) return f4_zero();
        return ld4(gout + (size_t)p * a.Co + m0 + lv);
    };
    auto load_b = [&](int p) -> float4 {
        if (p >= p_end || !col_ok) return f4_zero();
        int xo = p % a.Wo, yo = (p / a.Wo) % a.Ho, n = p / (a.Wo * a.Ho);
        int yi = yo * a.stride + r - a.pad, xi = xo * a.stride + s - a.pad;
        if (yi < 0 || yi >= a.Hi || xi < 0 || xi >= a.Wi) return f4_zero();
        return ld4(in + (((size_t)n * a.Hi + yi) * a.Wi + xi) * a.Ci + cc);
    };
    const int tx = tid & 15, ty = tid >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int nk = (p_end - p_begin + TK - 1) / TK;
    if (nk <= 0) return;
    float4 va = load_a(p_begin + lp), vb = load_b(p_begin + lp);
    *reinterpret_cast<float4*>(&As[0][lp][lv]) = va;
    *reinterpret_cast<float4*>(&Bs[0][lp][lv]) = vb;
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) { va = load_a(p_begin + (kt + 1) * TK + lp); vb = load_b(p_begin + (kt + 1) * TK + lp); }
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            float4 av = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            float4 bv = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            *reinterpret_cast<float4*>(&As[buf ^ 1][lp][lv]) = va;
            *reinterpret_cast<float4*>(&Bs[buf ^ 1][lp][lv]) = vb;
        }
        __syncthreads();
    }
    const int RS = a.R * a.S;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int k = m0 + ty * 4 + i;
        if (k >= a.Co) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int cl = n0 + tx * 4 + j;
            if (cl >= a.NC) continue;
            int t = cl / a.Ci, c = cl - t * a.Ci;
            if (c < a.Ci_real) {
                // grouped convolution computed densely (block-diagonal packed weights): only the group's own input channels exist in
                // the reference-layout gradient [Co][Ci/groups][R][S]
                const int cpg = a.Ci_real / a.groups, cg0 = (k / (a.Co / a.groups)) * cpg;
                if (c >= cg0 && c < cg0 + cpg) atomicAdd(dw + ((size_t)k * cpg + (c - cg0)) * RS + t, acc[i][j]);
            }
        }
    }
}
void k_conv_wgrad_simt(cudaStream_t st, DType dt, const void* in, const void* gout, float* dw, int Ci_real, const ConvGeom& g, int groups) {
    SALT_COUNT(1);
    WGradArgs a;
    a.groups = groups;
    a.P = g.B * g.Ho * g.Wo; a.Co = g.Co; a.Ci = g.Ci; a.Ci_real = Ci_real; a.NC = g.R * g.S * g.Ci;
    a.Hi = g.Hi; a.Wi = g.Wi; a.Ho = g.Ho; a.Wo = g.Wo; a.R = g.R; a.S = g.S; a.stride = g.stride; a.pad = g.pad;
    int tiles = cdiv(a.Co, TM) * cdiv(a.NC, TN);
    int splits = cdiv(148 * 4, tiles);
    int max_splits = cdiv(a.P, TK * 8);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    a.chunk = cdiv(cdiv(a.P, splits), TK) * TK;
    splits = cdiv(a.P, a.chunk);
    dim3 grid(cdiv(a.Co, TM), cdiv(a.NC, TN), splits);
    SALT_DISPATCH(dt, T, (conv_wgrad_simt_kernel<T><<<grid, 256, 0, st>>>((const T*)in, (const T*)gout, dw, a)));
}

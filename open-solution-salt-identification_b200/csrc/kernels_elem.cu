// Elementwise / reduction kernels of the U-Net engine: BatchNorm (forward + backward), residual add,
// replicate-border writers, bilinear gather (upsample + virtual concat -> bordered conv input) and its
// adjoint, scSE forward/backward, the final 1x1 conv.  All NHWC, storage type T in {float, bf16},
// arithmetic fp32.  These are HBM-bound passes: every thread moves 16 bytes of storage (4 fp32 / 8 bf16
// channels), rows map to blockIdx.x so that the index arithmetic is 32-bit and division-light.
#include "kernels.h"
#include <algorithm>
#include <stdexcept>

#include "elem_vec.cuh"
#include "kernels_stream.h"

// ------------------------------------------------------------------------------------------------
// input adapter: fp32 NCHW [B,3,H,W] -> NHWC with C padded to 4
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void input_nchw_to_nhwc4_kernel(const float* __restrict__ x, T* __restrict__ out, int HW) {
    const int n = blockIdx.y, p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    const float* xp = x + (size_t)n * 3 * HW + p;
    T* o = out + ((size_t)n * HW + p) * 4;
    st4(o, make_float4(xp[0], xp[HW], xp[2 * (size_t)HW], 0.f));
}
void k_input_nchw_to_nhwc4(cudaStream_t st, DType dt, const float* x, void* out, int B, int H, int W) {
    SALT_COUNT(1);
    SALT_DISPATCH(dt, T, (input_nchw_to_nhwc4_kernel<T><<<dim3(cdiv(H * W, 256), B), 256, 0, st>>>(x, (T*)out, H * W)));
}
void k_zero(cudaStream_t st, void* p, size_t bytes) { cudaMemsetAsync(p, 0, bytes, st); }

#define STEM_PATCH_C 160
template <typename T>
__global__ void stem_im2col_kernel(const float* __restrict__ x, T* __restrict__ out, int H, int W) {
    // one block per output row; a thread owns ONE channel vector of the patch (its (c, r, s) decomposition and row validity are
    // computed once) and walks over the row's pixels - the per-element div/mod of the one-pixel-per-thread version made this pass
    // integer-ALU bound at 0.9 TB/s (profiles/r2_notes.md)
    constexpr int N = VW<T>::N, CG = STEM_PATCH_C / N, LANES = EW_THREADS / CG;
    const int Ho = H / 2, Wo = W / 2;
    const int cv = threadIdx.x % CG, lane = threadIdx.x / CG;
    if (lane >= LANES) return;
    const int row = blockIdx.x, n = row / Ho, yo = row - n * Ho;
    const float* xn = x + (size_t)n * 3 * H * W;
    int base[N], sx[N];                 // offset of (c, yi, -3 + s) in the image; column shift; base < 0: the whole tap row is outside
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const int ch = cv * N + i;                       // c*49 + r*7 + s
        const int c = ch / 49, rs = ch - c * 49, r = rs / 7, s = rs - r * 7;
        const int yi = 2 * yo + r - 3;
        sx[i] = s - 3;
        base[i] = (ch < 147 && yi >= 0 && yi < H) ? (c * H + yi) * W : -1;
    }
    for (int xo = lane; xo < Wo; xo += LANES) {
        Vf<N> v;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const int xi = 2 * xo + sx[i];
            v.v[i] = (base[i] >= 0 && xi >= 0 && xi < W) ? __ldg(xn + base[i] + xi) : 0.f;
        }
        stv(out + ((size_t)row * Wo + xo) * STEM_PATCH_C + cv * N, v);
    }
}
void k_stem_im2col(cudaStream_t st, DType dt, const float* x, void* patches, int B, int H, int W) {
    SaltProfScope prof_scope(SALT_PROF_OTHER, (double)B * 3 * H * W * 4 + (double)B * (H / 2) * (W / 2) * STEM_PATCH_C * dtype_size(dt), st);
    SALT_COUNT(1);
    SALT_DISPATCH(dt, T, (stem_im2col_kernel<T><<<B * (H / 2), EW_THREADS, 0, st>>>(x, (T*)patches, H, W)));
}

// ------------------------------------------------------------------------------------------------
// BatchNorm finalisation (nn.BatchNorm2d: eps 1e-5, momentum 0.1, biased var to normalise, unbiased to track)
// ------------------------------------------------------------------------------------------------
// Fixed-order sum of the partial slots of channel c = blockIdx.x*32 + threadIdx.x: thread row y adds slots y, y+32, ... in fp64
// (all its loads are independent and issued together: the kernel costs about one memory round trip), then row 0 adds the 32 row
// sums in order.  Block = (32, 32).  Returns the two sums in row 0 (other rows return garbage).
#define SLOT_ROWS 32
__device__ __forceinline__ void slot_sums(const float* __restrict__ part, int nslots, int C, int c, double& s0, double& s1) {
    __shared__ double sm[SLOT_ROWS][2][32];
    double a = 0.0, b = 0.0;
    if (c < C) {
        int slot = threadIdx.y;
        for (; slot + 3 * SLOT_ROWS < nslots; slot += 4 * SLOT_ROWS) {
            float va[4], vb[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                va[u] = __ldcg(part + (size_t)(slot + u * SLOT_ROWS) * 2 * C + c);
                vb[u] = __ldcg(part + (size_t)(slot + u * SLOT_ROWS) * 2 * C + C + c);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) { a += (double)va[u]; b += (double)vb[u]; }
        }
        for (; slot < nslots; slot += SLOT_ROWS) {
            a += (double)__ldcg(part + (size_t)slot * 2 * C + c);
            b += (double)__ldcg(part + (size_t)slot * 2 * C + C + c);
        }
    }
    sm[threadIdx.y][0][threadIdx.x] = a; sm[threadIdx.y][1][threadIdx.x] = b;
    __syncthreads();
    s0 = 0.0; s1 = 0.0;
    if (threadIdx.y == 0) {
#pragma unroll 8
        for (int g = 0; g < SLOT_ROWS; ++g) { s0 += sm[g][0][threadIdx.x]; s1 += sm[g][1][threadIdx.x]; }
    }
}
__global__ void stats_reduce_kernel(const float* __restrict__ part, int nslots, int C, double* __restrict__ out) {
    const int c = blockIdx.x * 32 + threadIdx.x;
    double s0, s1;
    slot_sums(part, nslots, C, c, s0, s1);
    if (threadIdx.y == 0 && c < C) { out[c] = s0; out[C + c] = s1; }
}
void k_stats_reduce(cudaStream_t st, const float* stats, int nslots, int C, double* out) {
    SALT_COUNT(1);
    stats_reduce_kernel<<<cdiv(C, 32), dim3(32, SLOT_ROWS), 0, st>>>(stats, nslots, C, out);
}
__global__ void bn_finalize_train_kernel(BNRef bn, int nslots, double count, float momentum, float eps) {
    const int c = blockIdx.x * 32 + threadIdx.x;
    double s0, s1;
    slot_sums(bn.sums, nslots, bn.C, c, s0, s1);
    if (threadIdx.y != 0 || c >= bn.C) return;
    double mean = s0 / count;
    double var = s1 / count - mean * mean;
    if (var < 0) var = 0;
    double invstd = 1.0 / sqrt(var + (double)eps);
    float sc = (float)(bn.gamma[c] * invstd);
    bn.scale[c] = sc;
    bn.shift[c] = (float)(bn.beta[c] - mean * (double)sc);
    bn.mean[c] = (float)mean;
    bn.invstd[c] = (float)invstd;
    double unbiased = count > 1 ? var * count / (count - 1) : var;
    bn.rmean[c] = (float)((1.0 - momentum) * bn.rmean[c] + momentum * mean);
    bn.rvar[c] = (float)((1.0 - momentum) * bn.rvar[c] + momentum * unbiased);
}
void k_bn_finalize_train(cudaStream_t st, const BNRef& bn, int nslots, double count, float momentum, float eps) {
    SaltProfScope prof_scope(SALT_PROF_BN_FINALIZE, 0.0, st);
    SALT_COUNT(1);
    bn_finalize_train_kernel<<<cdiv(bn.C, 32), dim3(32, SLOT_ROWS), 0, st>>>(bn, nslots, count, momentum, eps);
}
__global__ void bn_finalize_eval_kernel(BNRef bn, float eps) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= bn.C) return;
    float invstd = 1.0f / sqrtf(bn.rvar[c] + eps);
    float sc = bn.gamma[c] * invstd;
    bn.scale[c] = sc;
    bn.shift[c] = bn.beta[c] - bn.rmean[c] * sc;
}
void k_bn_finalize_eval(cudaStream_t st, const BNRef& bn, float eps) {
    SaltProfScope prof_scope(SALT_PROF_BN_FINALIZE, 0.0, st);
    SALT_COUNT(1);
    bn_finalize_eval_kernel<<<cdiv(bn.C, 128), 128, 0, st>>>(bn, eps);
}
__global__ void bn_bwd_finalize_kernel(BNRef bn, double count) {
    const int c = blockIdx.x * 32 + threadIdx.x;
    double sg, sgx;
    slot_sums(bn.bsums, min(*bn.bslots, SALT_STAT_SLOTS_BWD), bn.C, c, sg, sgx);
    if (threadIdx.y != 0 || c >= bn.C) return;
    bn.dbeta[c] += (float)sg;
    bn.dgamma[c] += (float)sgx;
    bn.cb[c] = (float)(sg / count);
    bn.cc[c] = (float)(sgx / count * (double)bn.invstd[c]);
}
void k_bn_bwd_finalize(cudaStream_t st, const BNRef& bn, double count) {
    SaltProfScope prof_scope(SALT_PROF_BN_FINALIZE, 0.0, st);
    SALT_COUNT(1);
    bn_bwd_finalize_kernel<<<cdiv(bn.C, 32), dim3(32, SLOT_ROWS), 0, st>>>(bn, count);
}

// ------------------------------------------------------------------------------------------------
// BN apply (+ residual) (+ ReLU), optional replicate border on the output.  grid = (B*H rows, row blocks)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(EW_THREADS, 2) bn_apply_kernel(const T* __restrict__ raw, const float* __restrict__ scale, const float* __restrict__ shift,
                                const float* __restrict__ gate, const T* __restrict__ res, const float* __restrict__ rscale,
                                const float* __restrict__ rshift, int relu, T* __restrict__ out, int nrows, int rows_per_block, int H,
                                int W, int C, int pt, int pb, int pl, int pr) {
    // one block per `rows_per_block` image rows (and per slab of EW_THREADS channel groups when C is very wide); a thread keeps
    // ONE channel group (its scale/shift live in registers) and walks over the block's pixels in flat order, FOUR pixels per
    // iteration so that 4 independent 16-byte loads are in flight per thread at 3 blocks per SM - the one-row / two-pixel version
    // ran at 1.9 TB/s, latency-bound (profiles/r1_launches_f_final.md).
    constexpr int N = VW<T>::N;
    const int cg = min(C / N, EW_THREADS), lanes = EW_THREADS / cg;
    const int cv = blockIdx.y * cg + threadIdx.x % cg, lane = threadIdx.x / cg, c = cv * N;
    Vf<N> sc = ldp<N>(scale + c);
    const Vf<N> sh = ldp<N>(shift + c);
    Vf<N> rsc = vzero<N>(), rsh = vzero<N>();
    if (res && rscale) { rsc = ldp<N>(rscale + c); rsh = ldp<N>(rshift + c); }
    const int Hp = H + pt + pb, Wp = W + pl + pr;
    const int q_end = min(nrows, (int)(blockIdx.x + 1) * rows_per_block) * W;     // flat pixel range of this block
    auto finish = [&](int q, Vf<N> v, Vf<N> r) {
        const int row = q / W, x = q - row * W, n = row / H, y = row - n * H;
        v = vfma(v, sc, sh);
        if (gate) v = vmul(v, ldp<N>(gate + (size_t)n * C + c));          // per-(image, channel) SE gate: (raw*scale+shift)*gate
        if (res) {
            if (rscale) r = vfma(r, rsc, rsh);
            v = vadd(v, r);
        }
        if (relu) v = vrelu(v);
        const int y0 = (y == 0) ? 0 : y + pt, y1 = (y == H - 1) ? Hp - 1 : y + pt;
        const int x0 = (x == 0) ? 0 : x + pl, x1 = (x == W - 1) ? Wp - 1 : x + pl;
        for (int yy = y0; yy <= y1; ++yy)
            for (int xx = x0; xx <= x1; ++xx) stv(out + (((size_t)n * Hp + yy) * Wp + xx) * C + c, v);
    };
    int q = blockIdx.x * rows_per_block * W + lane;
    for (; q + 3 * lanes < q_end; q += 4 * lanes) {
        if (res) {                                   // residual: 2 + 2 pixels (the same 4 loads in flight per half)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                Vf<N> a[2], r[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const size_t o = (size_t)(q + (2 * h + u) * lanes) * C + c;
                    a[u] = ldv(raw + o); r[u] = ldv(res + o);
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) finish(q + (2 * h + u) * lanes, a[u], r[u]);
            }
        } else {
            Vf<N> a[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) a[u] = ldv(raw + (size_t)(q + u * lanes) * C + c);
#pragma unroll
            for (int u = 0; u < 4; ++u) finish(q + u * lanes, a[u], vzero<N>());
        }
    }
    for (; q < q_end; q += lanes) {
        const size_t o = (size_t)q * C + c;
        finish(q, ldv(raw + o), res ? ldv(res + o) : vzero<N>());
    }
}
void k_bn_apply(cudaStream_t st, const Tensor& raw, const float* scale, const float* shift, const Tensor* res,
                const float* rscale, const float* rshift, bool relu, const Tensor& out, const float* gate) {
    SaltProfScope prof_scope(SALT_PROF_BN_APPLY, (double)raw.bytes() * (res ? 2 : 1) + (double)out.bytes(), st);
    SALT_COUNT(1);
    if (k_ring_bn_apply(st, raw, scale, shift, res, rscale, rshift, relu, out, gate)) return;
    SALT_DISPATCH(raw.dt, T, {
        const int nrows = raw.B * raw.H;
        const long long row_bytes = (long long)raw.W * raw.C * sizeof(T);
        int rpb = (int)std::max<long long>(1, std::min<long long>(8, 32768 / std::max<long long>(1, row_bytes)));
        while (rpb > 1 && cdiv(nrows, rpb) < 148 * 4) rpb >>= 1;           // keep every SM busy on small layers
        dim3 grid(cdiv(nrows, rpb), cdiv(raw.C / VW<T>::N, EW_THREADS));
        bn_apply_kernel<T><<<grid, EW_THREADS, 0, st>>>((const T*)raw.p, scale, shift, gate, res ? (const T*)res->p : nullptr, rscale,
                                                        rshift, relu ? 1 : 0, (T*)out.p, nrows, rpb, raw.H, raw.W, raw.C, out.pt, out.pb,
                                                        out.pl, out.pr);
    });
}

template <typename T>
__global__ void bn_relu_avgpool_kernel(const T* __restrict__ raw, const float* __restrict__ scale,
                                       const float* __restrict__ shift, T* __restrict__ out, int Ho, int Wo, int C) {
    constexpr int N = VW<T>::N;
    const int cg = C / N;
    const int j = blockIdx.y * EW_THREADS + threadIdx.x;
    if (j >= Wo * cg) return;
    const int x = j / cg, c = (j - x * cg) * N;
    const int row = blockIdx.x, n = row / Ho, y = row - n * Ho, Wi = Wo * 2, Hi = Ho * 2;
    const Vf<N> sc = ldp<N>(scale + c), sh = ldp<N>(shift + c);
    Vf<N> acc = vzero<N>();
    for (int dy = 0; dy < 2; ++dy)
        for (int dx = 0; dx < 2; ++dx)
            acc = vadd(acc, vrelu(vfma(ldv(raw + (((size_t)n * Hi + 2 * y + dy) * Wi + 2 * x + dx) * C + c), sc, sh)));
    stv(out + ((size_t)row * Wo + x) * C + c, vscale(acc, 0.25f));
}
void k_bn_relu_avgpool(cudaStream_t st, const Tensor& raw, const float* scale, const float* shift, const Tensor& out) {
    SaltProfScope prof_scope(SALT_PROF_OTHER, (double)raw.bytes() + (double)out.bytes(), st);
    SALT_COUNT(1);
    SALT_DISPATCH(raw.dt, T, {
        dim3 grid(out.B * out.H, cdiv(out.W * (out.C / VW<T>::N), EW_THREADS));
        bn_relu_avgpool_kernel<T><<<grid, EW_THREADS, 0, st>>>((const T*)raw.p, scale, shift, (T*)out.p, out.H, out.W, out.C);
    });
}
template <typename T>
__global__ void avgpool_bwd_kernel(const T* __restrict__ gout, T* __restrict__ gin, int Hi, int Wi, int C) {
    constexpr int N = VW<T>::N;
    const int cg = C / N;
    const int j = blockIdx.y * EW_THREADS + threadIdx.x;
    if (j >= Wi * cg) return;
    const int x = j / cg, c = (j - x * cg) * N;
    const int row = blockIdx.x, n = row / Hi, y = row - n * Hi;
    const size_t o = (((size_t)n * (Hi / 2) + y / 2) * (Wi / 2) + x / 2) * C + c;
    stv(gin + ((size_t)row * Wi + x) * C + c, vscale(ldv(gout + o), 0.25f));
}
void k_avgpool_bwd(cudaStream_t st, const Tensor& gout, const Tensor& gin) {
    SaltProfScope prof_scope(SALT_PROF_OTHER, (double)gin.bytes() + (double)gout.bytes(), st);
    SALT_COUNT(1);
    SALT_DISPATCH(gin.dt, T, {
        dim3 grid(gin.B * gin.H, cdiv(gin.W * (gin.C / VW<T>::N), EW_THREADS));
        avgpool_bwd_kernel<T><<<grid, EW_THREADS, 0, st>>>((const T*)gout.p, (T*)gin.p, gin.H, gin.W, gin.C);
    });
}

// ------------------------------------------------------------------------------------------------
// gather: bordered conv input = concat_c( bilinear_upsample_f(src_i) ), replicate border
//   bilinear: align_corners=False, src = max((d+0.5)/f - 0.5, 0)  (torch upsample_bilinear2d)
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ void bilin_coord(int d, int f, int nsrc, int& i0, int& i1, float& l) {
    float s = ((float)d + 0.5f) * (1.0f / (float)f) - 0.5f;
    s = fmaxf(s, 0.f);
    i0 = (int)s;
    i1 = min(i0 + 1, nsrc - 1);
    l = s - (float)i0;
}

// One block per physical output row.  A work item = (source, run, channel vector); a run is the stretch of consecutive output
// columns that interpolate between the SAME two source columns (f columns for an f-times upsampled source, the first / last
// run also cover the replicated border), so the four source vectors are loaded once per run instead of once per output pixel
// (the per-pixel version was L1/L2-throughput bound: profiles/r1_notes.md).
struct GatherPlan { GatherSrc s[5]; int n; int item0[6]; int c0[5]; };

template <typename T>
__global__ void gather_fwd_kernel(T* __restrict__ out, GatherPlan a, int H, int W, int C, int pt, int pb, int pl, int pr) {
    constexpr int N = VW<T>::N;
    const int Hp = H + pt + pb, Wp = W + pl + pr;
    const int row = blockIdx.x, n = row / Hp, yp = row - n * Hp;
    const int y = min(max(yp - pt, 0), H - 1);
    T* orow = out + (size_t)row * Wp * C;
    for (int item = threadIdx.x; item < a.item0[a.n]; item += EW_THREADS) {
        int si = 0;
        while (si < a.n - 1 && item >= a.item0[si + 1]) ++si;
        const GatherSrc s = a.s[si];
        const int cg = s.C / N, local = item - a.item0[si];
        const int run = local / cg, cv = local - run * cg;
        const T* sp = (const T*)s.p + (size_t)n * s.H * s.W * s.C + cv * N;
        T* o = orow + a.c0[si] + cv * N;
        if (s.f == 1) {                                   // plain copy (+ replicate border): one run per physical column
            const int x = min(max(run - pl, 0), W - 1);
            stv_stream(o + (size_t)run * C, ldv(sp + ((size_t)y * s.W + x) * s.C));
            continue;
        }
        // run r covers logical x in [r*f - f/2, r*f + f/2): all of them interpolate between source columns r-1 and r
        const int f = s.f, xb = run * f - (f >> 1);
        const int xlo = max(xb, 0), xhi = min(xb + f, W);   // logical range [xlo, xhi)
        int y0, y1; float ly;
        bilin_coord(y, f, s.H, y0, y1, ly);
        const int j0 = max(run - 1, 0), j1 = min(run, s.W - 1);
        const Vf<N> v00 = ldv(sp + ((size_t)y0 * s.W + j0) * s.C), v01 = ldv(sp + ((size_t)y0 * s.W + j1) * s.C);
        const Vf<N> v10 = ldv(sp + ((size_t)y1 * s.W + j0) * s.C), v11 = ldv(sp + ((size_t)y1 * s.W + j1) * s.C);
        const float wy0 = 1.f - ly;
        // physical columns: logical x -> x + pl; the first run also fills the left border, the last run the right border
        const int plo = (xlo == 0) ? 0 : xlo + pl, phi = (xhi == W) ? Wp : xhi + pl;
        for (int xp = plo; xp < phi; ++xp) {
            const int x = min(max(xp - pl, 0), W - 1);
            int x0, x1; float lx;
            bilin_coord(x, f, s.W, x0, x1, lx);           // x0 == j0, x1 == j1 by construction
            const float wx0 = 1.f - lx;
            Vf<N> v;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                float top = v00.v[i] * wx0 + v01.v[i] * lx, bot = v10.v[i] * wx0 + v11.v[i] * lx;
                v.v[i] = top * wy0 + bot * ly;
            }
            stv_stream(o + (size_t)xp * C, v);
        }
    }
}
void k_gather_fwd(cudaStream_t st, const Tensor& out, const GatherSrc* srcs, int nsrc) {
    SaltProfScope prof_scope(SALT_PROF_GATHER, 2.0 * (double)out.bytes(), st);
    SALT_COUNT(1);
    SALT_DISPATCH(out.dt, T, {
        GatherPlan a; a.n = nsrc; a.item0[0] = 0;
        int c0 = 0;
        for (int i = 0; i < nsrc; ++i) {
            a.s[i] = srcs[i]; a.c0[i] = c0; c0 += srcs[i].C;
            const int runs = srcs[i].f == 1 ? out.Wp() : out.W / srcs[i].f + 1;
            a.item0[i + 1] = a.item0[i] + runs * (srcs[i].C / VW<T>::N);
        }
        gather_fwd_kernel<T><<<out.B * out.Hp(), EW_THREADS, 0, st>>>((T*)out.p, a, out.H, out.W, out.C, out.pt, out.pb, out.pl, out.pr);
    });
}

// adjoint of the replicate border for a copied (f == 1) source: gsrc[y,x] (+)= sum of the border cells mapped to it
template <typename T>
__global__ void fold_bwd_kernel(const T* __restrict__ gP, int c0, T* __restrict__ gsrc, int accumulate,
                                int H, int W, int Cp, int Cs, int pt, int pb, int pl, int pr, int cgs) {
    // one (pixel, channel vector) item per thread; a four-items-per-thread version with the centre loads issued up front measured
    // 2x SLOWER (110 registers, profiles/r2_notes.md)
    constexpr int N = VW<T>::N;
    const int cg = Cs / N, Hp = H + pt + pb, Wp = W + pl + pr;
    const int j = blockIdx.y * EW_THREADS + threadIdx.x;
    if (j >= W * cg) return;
    // (the pass is instruction-bound - ncu: issue slots 61 % busy at 3.1 TB/s - so the index arithmetic is kept short)
    const int x = cgs >= 0 ? (j >> cgs) : j / cg, c = (j - x * cg) * N;
    const int row = blockIdx.x, n = row / H, y = row - n * H;
    const T* img = gP + (size_t)n * Hp * Wp * Cp + c0 + c;
    T* o = gsrc + ((size_t)row * W + x) * Cs + c;
    Vf<N> acc = ldv(img + ((size_t)(y + pt) * Wp + x + pl) * Cp);
    if (y == 0 || y == H - 1 || x == 0 || x == W - 1) {                        // replicated border cells fold onto the edge pixels
        const int y0 = (y == 0) ? 0 : y + pt, y1 = (y == H - 1) ? Hp - 1 : y + pt;
        const int x0 = (x == 0) ? 0 : x + pl, x1 = (x == W - 1) ? Wp - 1 : x + pl;
        for (int yy = y0; yy <= y1; ++yy)
            for (int xx = x0; xx <= x1; ++xx)
                if (yy != y + pt || xx != x + pl) acc = vadd(acc, ldv(img + ((size_t)yy * Wp + xx) * Cp));
    }
    if (accumulate) acc = vadd(acc, ldv(o));
    stv(o, acc);
}
void k_fold_bwd(cudaStream_t st, const Tensor& gP, int c0, const Tensor& gsrc, bool accumulate) {
    SaltProfScope prof_scope(SALT_PROF_GATHER_BWD, 2.0 * (double)gsrc.bytes(), st);
    SALT_COUNT(1);
    SALT_DISPATCH(gP.dt, T, {
        dim3 grid(gsrc.B * gsrc.H, cdiv(gsrc.W * (gsrc.C / VW<T>::N), EW_THREADS));
        const int cg = gsrc.C / VW<T>::N;
        int cgs = -1;                                   // log2 of the channel-group count when it is a power of two
        if ((cg & (cg - 1)) == 0) { cgs = 0; while ((1 << cgs) < cg) ++cgs; }
        fold_bwd_kernel<T><<<grid, EW_THREADS, 0, st>>>((const T*)gP.p, c0, (T*)gsrc.p, accumulate ? 1 : 0, gP.H, gP.W, gP.C,
                                                        gsrc.C, gP.pt, gP.pb, gP.pl, gP.pr, cgs);
    });
}

// adjoint of (replicate border o bilinear upsample xf), separable: first along x into tmp (fp32), then along y.
// weight of physical destination coordinate p for source index j:
__device__ __forceinline__ float bilin_adj_w(int p, int pad, int ndst, int f, int nsrc, int j) {
    int d = min(max(p - pad, 0), ndst - 1);
    int i0, i1; float l;
    bilin_coord(d, f, nsrc, i0, i1, l);
    return (i0 == j ? 1.f - l : 0.f) + (i1 == j ? l : 0.f);
}
__device__ __forceinline__ void adj_range(int j, int f, int pad, int ndst, int nphys, int& lo, int& hi) {
    int dlo = (j - 1) * f, dhi = (j + 2) * f;            // logical candidates [dlo, dhi)
    lo = dlo <= 0 ? 0 : dlo + pad;
    hi = dhi >= ndst ? nphys : dhi + pad;
}
// One block per (image, source row jy).  Phase A reduces along y straight from global memory: item = (physical column xp, channel
// vector), all threads walk the SAME 2f destination rows with non-zero weight (uniform weights, unconditional independent loads, four
// rows in flight), the fp32 row of partial sums stays in shared memory.  Phase B reduces along x out of shared memory and writes the
// source-row gradient.  Nothing but the gradient slice is read and nothing but the source gradient is written: the two-kernel
// version went through an fp32 temporary in global memory that is as large as the slice itself for f = 2 (profiles/r2_notes.md).
template <typename T>
__global__ void __launch_bounds__(EW_THREADS) upsample_bwd_kernel(const T* __restrict__ gP, int c0, T* __restrict__ gsrc, int accumulate,
                                                                  int H, int W, int Cp, int Cs, int pt, int pb, int pl, int pr, int f,
                                                                  int Cslab, int cgs, int rpb) {
    constexpr int N = VW<T>::N;
    extern __shared__ __align__(16) float rowacc[];      // [Wp][Cslab]: blockIdx.y selects a slab of Cslab source channels
    const int cg = Cslab / N, Hp = H + pt + pb, Wp = W + pl + pr, Hs = H / f, Ws = W / f;
    c0 += blockIdx.y * Cslab;
    gsrc += blockIdx.y * Cslab;
    // the pass was instruction-bound (ncu: issue slots 70 % busy at 2 TB/s): the bilinear weights are tabulated instead of being
    // recomputed per (item, row) and per (item, column); the column table is shared by the `rpb` source rows a block walks over
    // (source index 1 has 3f + border candidates: its range starts at physical 0, so the tables are 3f + 4 wide)
    const int tw = 3 * f + 4;
    float* wy = rowacc + (size_t)Wp * Cslab;             // [tw] weights of the candidate rows
    float* wx = wy + tw;                                 // [Ws][tw] weights of the candidate columns of source column jx
    for (int i = threadIdx.x; i < Ws * tw; i += EW_THREADS) {
        const int jx = i / tw, k = i - jx * tw;
        int xlo, xhi;
        adj_range(jx, f, pl, W, Wp, xlo, xhi);
        wx[i] = xlo + k < xhi ? bilin_adj_w(xlo + k, pl, W, f, Ws, jx) : 0.f;
    }
    for (int row = blockIdx.x * rpb; row < (int)(blockIdx.x + 1) * rpb; ++row) {
        const int n = row / Hs, jy = row - n * Hs;
        int lo, hi;
        adj_range(jy, f, pt, H, Hp, lo, hi);
        __syncthreads();                                 // previous row's phase B is done with rowacc / wy
        for (int i = threadIdx.x; i < hi - lo; i += EW_THREADS) wy[i] = bilin_adj_w(lo + i, pt, H, f, Hs, jy);
        __syncthreads();
        int lo2 = lo, hi2 = hi;                          // the rows with non-zero weight are contiguous
        while (lo2 < hi2 && wy[lo2 - lo] == 0.f) ++lo2;
        while (hi2 > lo2 && wy[hi2 - 1 - lo] == 0.f) --hi2;
        const T* base = gP + (size_t)n * Hp * Wp * Cp + c0;
        const size_t rstride = (size_t)Wp * Cp;
        for (int item = threadIdx.x; item < Wp * cg; item += EW_THREADS) {
            const int xp = cgs >= 0 ? (item >> cgs) : item / cg, c = (item - xp * cg) * N;
            const T* col = base + (size_t)xp * Cp + c + (size_t)lo2 * rstride;
            const float* w = wy + (lo2 - lo);
            Vf<N> acc = vzero<N>();
            int k = 0;
            for (; k + 3 < hi2 - lo2; k += 4) {
                Vf<N> v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = ldv(col + (size_t)(k + u) * rstride);
#pragma unroll
                for (int u = 0; u < 4; ++u) acc = vaxpy(v[u], w[k + u], acc);
            }
            for (; k < hi2 - lo2; ++k) acc = vaxpy(ldv(col + (size_t)k * rstride), w[k], acc);
            stp<N>(rowacc + (size_t)xp * Cslab + c, acc);
        }
        __syncthreads();
        for (int item = threadIdx.x; item < Ws * cg; item += EW_THREADS) {
            const int jx = cgs >= 0 ? (item >> cgs) : item / cg, c = (item - jx * cg) * N;
            int xlo, xhi;
            adj_range(jx, f, pl, W, Wp, xlo, xhi);
            const float* w = wx + jx * tw;
            Vf<N> acc = vzero<N>();
            for (int k = 0; k < xhi - xlo; ++k) {
                if (w[k] != 0.f) acc = vaxpy(ldp<N>(rowacc + (size_t)(xlo + k) * Cslab + c), w[k], acc);
            }
            T* o = gsrc + ((size_t)row * Ws + jx) * Cs + c;
            if (accumulate) acc = vadd(acc, ldv(o));
            stv(o, acc);
        }
    }
}
size_t upsample_bwd_tmp_floats(const Tensor& gP, int f, int Csrc) {
    return (size_t)gP.B * gP.Hp() * (gP.W / f) * Csrc;
}
void k_upsample_bwd(cudaStream_t st, const Tensor& gP, int c0, int f, const Tensor& gsrc, float* /*tmp: unused since the fused kernel*/,
                    bool accumulate) {
    SaltProfScope prof_scope(SALT_PROF_GATHER_BWD, (double)gP.bytes() * gsrc.C / gP.C + (double)gsrc.bytes(), st);
    SALT_COUNT(1);
    SALT_DISPATCH(gP.dt, T, {
        int cslab = gsrc.C;                           // halve the channel slab until the fp32 row fits 64 KB
        while (gP.Wp() * cslab * (int)sizeof(float) > 64 * 1024 && cslab % (2 * VW<T>::N) == 0) cslab >>= 1;
        const int smem = (gP.Wp() * cslab + (3 * f + 4) * (1 + gP.W / f)) * (int)sizeof(float);
        const int cg = cslab / VW<T>::N;
        int cgs = -1;
        if ((cg & (cg - 1)) == 0) { cgs = 0; while ((1 << cgs) < cg) ++cgs; }
        static int smem_set = 0;
        if (smem > 48 * 1024 && smem > smem_set) {
            if (smem > 200 * 1024) throw std::runtime_error("upsample backward: row of partial sums exceeds shared memory");
            cudaFuncSetAttribute(upsample_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            smem_set = smem;
        }
        const int rows = gsrc.B * gsrc.H;
        int rpb = 1;                                    // source rows per block: amortises the column-weight table
        while (rpb < 4 && rows % (2 * rpb) == 0 && rows / (2 * rpb) >= 148 * 16) rpb *= 2;   // (fewer, longer blocks lost 30 % at f = 4, 8)
        upsample_bwd_kernel<T><<<dim3(rows / rpb, gsrc.C / cslab), EW_THREADS, smem, st>>>((const T*)gP.p, c0, (T*)gsrc.p,
                                                                         accumulate ? 1 : 0, gP.H, gP.W,
                                                                         gP.C, gsrc.C, gP.pt, gP.pb, gP.pl, gP.pr, f, cslab, cgs, rpb);
    });
}

// ------------------------------------------------------------------------------------------------
// squeeze-and-excitation
//   decoder scSE (base.py:82-117): z = relu(bn(raw));  out = relu(z*cse[n,c] + z*sse[n,y,x]) = z*(cse+sse)
//   encoder SE   (pretrainedmodels senet.py SEModule, restated in oracle/senet_restated.py): u = bn(raw);
//                out = relu(u*cse[n,c] + residual)  - the gate is applied by bn_apply / bn_bwd_* through their `gate` arguments
// ------------------------------------------------------------------------------------------------
// per-(n,c) sum over a pixel chunk of  [g *] [relu](raw*scale+shift)  ->  part[n][chunk][C]
// (the FC kernels add the chunks in a fixed order: deterministic and batch independent).  grid = (chunks, B, channel slabs)
template <typename T, bool WITH_G, bool RELU>
__global__ void se_pool_kernel(const T* __restrict__ raw, const T* __restrict__ g, const float* __restrict__ scale,
                               const float* __restrict__ shift, float* __restrict__ part, int HW, int C) {
    constexpr int N = VW<T>::N;
    __shared__ float red[N * EW_THREADS];
    const int cg = min(C / N, EW_THREADS), lanes = EW_THREADS / cg;
    const int cv = blockIdx.z * cg + threadIdx.x % cg, lane = threadIdx.x / cg, n = blockIdx.y;
    const int chunk = (HW + gridDim.x - 1) / gridDim.x;
    const int p0 = blockIdx.x * chunk, p1 = min(HW, p0 + chunk);
    const Vf<N> sc = ldp<N>(scale + cv * N), sh = ldp<N>(shift + cv * N);
    Vf<N> acc = vzero<N>();
    // four pixels per pass: all loads of a pass are issued before the first use (one load in flight per thread ran at 2 TB/s)
    for (int p = p0 + lane; p < p1; p += 4 * lanes) {
        Vf<N> xs[4], gs[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const size_t o = ((size_t)n * HW + min(p + u * lanes, p1 - 1)) * C + cv * N;
            xs[u] = ldv(raw + o);
            if (WITH_G) gs[u] = ldv(g + o);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (p + u * lanes < p1) {
                Vf<N> z = vfma(xs[u], sc, sh);
                if (RELU) z = vrelu(z);
                if (WITH_G) z = vmul(z, gs[u]);
                acc = vadd(acc, z);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < N; ++i) red[i * EW_THREADS + threadIdx.x] = acc.v[i];
    __syncthreads();
    if (threadIdx.x < cg) {
        float* o = part + ((size_t)n * gridDim.x + blockIdx.x) * C + (blockIdx.z * cg + threadIdx.x) * N;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            float s = 0.f;
            for (int t = threadIdx.x; t < EW_THREADS; t += cg) s += red[i * EW_THREADS + t];
            o[i] = s;
        }
    }
}
template <typename T, bool WITH_G, bool RELU>
static void launch_se_pool(cudaStream_t st, const Tensor& raw, const void* g, const float* scale, const float* shift, const SERef& se) {
    if (k_ring_se_pool(st, raw, WITH_G ? g : nullptr, RELU, scale, shift, se)) return;
    dim3 grid(se.chunks, raw.B, cdiv(raw.C / VW<T>::N, EW_THREADS));
    se_pool_kernel<T, WITH_G, RELU><<<grid, EW_THREADS, 0, st>>>((const T*)raw.p, (const T*)g, scale, shift, se.part, raw.H * raw.W, raw.C);
}
__device__ __forceinline__ float warp_sum(float v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// one block per image: squeeze -> fc(C->Cr) -> relu -> fc(Cr->C) -> sigmoid   (any C, Cr)
__global__ void se_fc_kernel(SERef se, float inv_hw) {
    extern __shared__ float sm[];
    float* gap = sm;               // [C]
    float* hid = sm + se.C;        // [Cr]
    const int n = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
    for (int c = tid; c < se.C; c += blockDim.x) {
        float t = 0.f;
        for (int k = 0; k < se.chunks; ++k) t += se.part[((size_t)n * se.chunks + k) * se.C + c];
        gap[c] = t * inv_hw;
        se.gap[(size_t)n * se.C + c] = gap[c];
    }
    __syncthreads();
    for (int j = warp; j < se.Cr; j += nwarps) {
        float a = 0.f;
        for (int i = lane; i < se.C; i += 32) a = fmaf(se.w1[(size_t)j * se.C + i], gap[i], a);
        a = warp_sum(a);
        if (lane == 0) {
            a = fmaxf(a + se.b1[j], 0.f);
            hid[j] = a;
            se.hid[(size_t)n * se.Cr + j] = a;
        }
    }
    __syncthreads();
    for (int c = tid; c < se.C; c += blockDim.x) {
        float a = se.b2[c];
        for (int j = 0; j < se.Cr; ++j) a = fmaf(se.w2[(size_t)c * se.Cr + j], hid[j], a);
        se.cse[(size_t)n * se.C + c] = 1.f / (1.f + expf(-a));
    }
}
// backward of the two FCs.  Stage 1, one block per image: part[n][*][c] = sum_pix g*z  ->  dpre2 (saved over part[n][0][c]),
// dhid, and G[n][c] = d loss / d gap / HW.  Stage 2, one thread per weight: batch reduction of the outer products (no atomics).
__global__ void se_fc_bwd_kernel(SERef se, float inv_hw) {
    extern __shared__ float sm[];
    float* dpre2 = sm;             // [C]
    float* dhid = sm + se.C;       // [Cr]
    const int n = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
    for (int c = tid; c < se.C; c += blockDim.x) {
        const float cs = se.cse[(size_t)n * se.C + c];
        float t = 0.f;
        for (int k = 0; k < se.chunks; ++k) t += se.part[((size_t)n * se.chunks + k) * se.C + c];
        const float d = t * cs * (1.f - cs);
        dpre2[c] = d;
        se.part[(size_t)n * se.chunks * se.C + c] = d;
    }
    __syncthreads();
    for (int j = warp; j < se.Cr; j += nwarps) {
        float a = 0.f;
        for (int i = lane; i < se.C; i += 32) a = fmaf(se.w2[(size_t)i * se.Cr + j], dpre2[i], a);
        a = warp_sum(a);
        if (lane == 0) {
            a = se.hid[(size_t)n * se.Cr + j] > 0.f ? a : 0.f;
            dhid[j] = a;
            se.dhid[(size_t)n * se.Cr + j] = a;
        }
    }
    __syncthreads();
    for (int c = tid; c < se.C; c += blockDim.x) {
        float gsum = 0.f;
        for (int j = 0; j < se.Cr; ++j) gsum = fmaf(se.w1[(size_t)j * se.C + c], dhid[j], gsum);
        se.G[(size_t)n * se.C + c] = gsum * inv_hw;
    }
}
__global__ void se_fc_wgrad_kernel(SERef se, int B) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int C = se.C, Cr = se.Cr;
    const size_t pstride = (size_t)se.chunks * C;          // dpre2[n][c] lives in part[n][0][c]
    if (idx < C * Cr) {
        { const int c = idx / Cr, j = idx - c * Cr;        // dw2[c][j] = sum_n dpre2[n][c] * hid[n][j]
          float a = 0.f;
          for (int n = 0; n < B; ++n) a = fmaf(se.part[n * pstride + c], se.hid[(size_t)n * Cr + j], a);
          se.dw2[idx] += a; }
        { const int j = idx / C, c = idx - j * C;          // dw1[j][c] = sum_n dhid[n][j] * gap[n][c]
          float a = 0.f;
          for (int n = 0; n < B; ++n) a = fmaf(se.dhid[(size_t)n * Cr + j], se.gap[(size_t)n * C + c], a);
          se.dw1[idx] += a; }
    }
    if (idx < C) {
        float a = 0.f;
        for (int n = 0; n < B; ++n) a += se.part[n * pstride + idx];
        se.db2[idx] += a;
    }
    if (idx < Cr) {
        float a = 0.f;
        for (int n = 0; n < B; ++n) a += se.dhid[(size_t)n * Cr + idx];
        se.db1[idx] += a;
    }
}
static void launch_se_fc(cudaStream_t st, const SERef& se, int B, int HW) {
    se_fc_kernel<<<B, 256, sizeof(float) * (se.C + se.Cr), st>>>(se, 1.0f / HW);
}
static void launch_se_fc_bwd(cudaStream_t st, const SERef& se, int B, int HW) {
    se_fc_bwd_kernel<<<B, 256, sizeof(float) * (se.C + se.Cr), st>>>(se, 1.0f / HW);
    se_fc_wgrad_kernel<<<cdiv(se.C * se.Cr, 256), 256, 0, st>>>(se, B);
}
// encoder SE: squeeze bn(raw) (no ReLU) and compute the channel gates se.cse[n][c]
void k_se_gate_fwd(cudaStream_t st, const Tensor& raw, const float* scale, const float* shift, const SERef& se) {
    SaltProfScope prof_scope(SALT_PROF_SCSE, (double)raw.bytes(), st);
    SALT_COUNT(2);
    SALT_DISPATCH(raw.dt, T, (launch_se_pool<T, false, false>(st, raw, nullptr, scale, shift, se)));
    launch_se_fc(st, se, raw.B, raw.H * raw.W);
}
// encoder SE backward: g = gradient w.r.t. u*cse (already ReLU-masked); accumulates the FC gradients and leaves
// se.G[n][c] = the per-(image, channel) term that the gap path adds to d loss / d u
void k_se_gate_bwd(cudaStream_t st, const Tensor& g, const Tensor& raw, const float* scale, const float* shift, const SERef& se) {
    SaltProfScope prof_scope(SALT_PROF_SCSE, 2.0 * (double)raw.bytes(), st);
    SALT_COUNT(3);
    SALT_DISPATCH(raw.dt, T, (launch_se_pool<T, true, false>(st, raw, g.p, scale, shift, se)));
    launch_se_fc_bwd(st, se, raw.B, raw.H * raw.W);
}

template <typename T>
__global__ void scse_apply_kernel(const T* __restrict__ raw, const float* __restrict__ scale, const float* __restrict__ shift,
                                  SERef se, T* __restrict__ out, unsigned npix, int HW, int C) {
    constexpr int N = 8;
    const unsigned cg = C / N;
    const unsigned idx = blockIdx.x * EW_THREADS + threadIdx.x;
    unsigned pix = idx / cg;
    const int c = (idx - pix * cg) * N;
    const bool ok = pix < npix;
    if (!ok) pix = npix - 1;
    const int n = pix / HW;
    const Vf<N> z = vrelu(vfma(ldv8(raw + (size_t)pix * C + c), ldp<N>(scale + c), ldp<N>(shift + c)));
    const float dot = group_sum(vdot(z, ldp<N>(se.ws + c)), cg);
    const float s = 1.f / (1.f + expf(-(dot + se.bs[0])));
    Vf<N> gate = ldp<N>(se.cse + (size_t)n * C + c);
#pragma unroll
    for (int i = 0; i < N; ++i) gate.v[i] += s;
    if (ok) stv8(out + (size_t)pix * C + c, vrelu(vmul(z, gate)));
}
void k_scse_fwd(cudaStream_t st, const Tensor& raw, const float* scale, const float* shift, const SERef& se, const Tensor& out) {
    SaltProfScope prof_scope(SALT_PROF_SCSE, 3.0 * (double)raw.bytes(), st);
    SALT_COUNT(3);
    const int HW = raw.H * raw.W, C = raw.C;
    const unsigned npix = (unsigned)raw.B * HW;
    if (C % 8 || C / 8 > 32 || ((C / 8) & (C / 8 - 1))) throw std::runtime_error("scSE: channel count must be 8 * 2^k <= 256");
    SALT_DISPATCH(raw.dt, T, {
        launch_se_pool<T, false, true>(st, raw, nullptr, scale, shift, se);
        launch_se_fc(st, se, raw.B, HW);
        if (!k_ring_scse_apply(st, raw, scale, shift, se, out))
            scse_apply_kernel<T><<<cdiv((long long)npix * (C / 8), EW_THREADS), EW_THREADS, 0, st>>>((const T*)raw.p, scale, shift, se, (T*)out.p, npix, HW, C);
    });
}

template <typename T>
__global__ void __launch_bounds__(EW_THREADS, 2) scse_bwd_apply_kernel(const T* __restrict__ gout, const T* __restrict__ raw, BNRef bn, SERef se,
                                      T* __restrict__ gbn, unsigned npix, int HW, int C) {
    constexpr int N = 8;
    __shared__ float red[N * EW_THREADS];
    const int cg = C / N, lanes = EW_THREADS / cg;
    const int cv = threadIdx.x % cg, lane = threadIdx.x / cg, c = cv * N;
    const Vf<N> sc = ldp<N>(bn.scale + c), sh = ldp<N>(bn.shift + c), mu = ldp<N>(bn.mean + c), is = ldp<N>(bn.invstd + c),
                w = ldp<N>(se.ws + c);
    const float bs = se.bs[0];
    Vf<N> sg = vzero<N>(), sgx = vzero<N>(), sws = vzero<N>();
    float sbs = 0.f;
    // all lanes of a pixel group must iterate together (shuffles) -> loop bound on the block's first pixel
    const unsigned bstep = gridDim.x * lanes;
    for (unsigned base0 = blockIdx.x * lanes; base0 < npix; base0 += 2 * bstep) {
      // two pixel groups per iteration: their 4 loads are issued back to back (<= 2 blocks per SM: one partial-sum slot per block)
      Vf<N> xs[2], gs[2];
      unsigned pixs[2]; bool oks[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        unsigned pix = base0 + u * bstep + lane;
        oks[u] = pix < npix;
        if (!oks[u]) pix = npix - 1;
        pixs[u] = pix;
        xs[u] = ldv8(raw + (size_t)pix * C + c);
        gs[u] = ldv8(gout + (size_t)pix * C + c);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (base0 + u * bstep >= npix) break;               // uniform over the block
        const unsigned pix = pixs[u];
        const bool ok = oks[u];
        const int n = pix / HW;
        const Vf<N> x = xs[u];
        const Vf<N> z = vrelu(vfma(x, sc, sh));
        const Vf<N> g = gs[u];
        const float dot = group_sum(vdot(z, w), cg);
        const float s = 1.f / (1.f + expf(-(dot + bs)));
        const float D = group_sum(vdot(g, z), cg);
        const float dsp = D * s * (1.f - s);
        const Vf<N> cse = ldp<N>(se.cse + (size_t)n * C + c), G = ldp<N>(se.G + (size_t)n * C + c);
        Vf<N> dz;
#pragma unroll
        for (int i = 0; i < N; ++i) dz.v[i] = z.v[i] > 0.f ? g.v[i] * (cse.v[i] + s) + dsp * w.v[i] + G.v[i] : 0.f;
        if (ok) {
            stv8(gbn + (size_t)pix * C + c, dz);
            sg = vadd(sg, dz);
            sgx = vfma(dz, vxhat(x, mu, is), sgx);
            sws = vaxpy(z, dsp, sws);
            if (cv == 0) sbs += dsp;
        }
      }
    }
    block_reduce_slot<N>(sg, cg, bn.bsums + (size_t)blockIdx.x * 2 * C, red);
    block_reduce_slot<N>(sgx, cg, bn.bsums + (size_t)blockIdx.x * 2 * C + C, red);
    if (blockIdx.x == 0 && threadIdx.x == 0) *bn.bslots = (int)gridDim.x;
    block_reduce_add<N, float>(sws, cg, se.dws, red);
    red[threadIdx.x] = sbs;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < EW_THREADS; ++i) t += red[i];
        atomicAdd(se.dbs, t);
    }
}
void k_scse_bwd(cudaStream_t st, const Tensor& gout, const Tensor& raw, const BNRef& bn, const SERef& se, const Tensor& gbn) {
    SaltProfScope prof_scope(SALT_PROF_SCSE, 5.0 * (double)raw.bytes(), st);
    SALT_COUNT(4);
    const int HW = raw.H * raw.W, C = raw.C;
    const unsigned npix = (unsigned)raw.B * HW;
    SALT_DISPATCH(raw.dt, T, {
        launch_se_pool<T, true, true>(st, raw, gout.p, bn.scale, bn.shift, se);
        launch_se_fc_bwd(st, se, raw.B, HW);
        if (!k_ring_scse_bwd_apply(st, gout, raw, bn, se, gbn))
            scse_bwd_apply_kernel<T><<<reduce_blocks(npix, C / 8), EW_THREADS, 0, st>>>((const T*)gout.p, (const T*)raw.p, bn, se, (T*)gbn.p, npix, HW, C);
    });
}

// ------------------------------------------------------------------------------------------------
// final 1x1 conv C -> K (K <= 4) on z = relu(bn(raw)); logits are fp32 NCHW (the reference layout)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void final_fwd_kernel(const T* __restrict__ raw, const float* __restrict__ scale, const float* __restrict__ shift,
                                 const float* __restrict__ w, const float* __restrict__ b, int K, float* __restrict__ logits,
                                 unsigned npix, int HW, int C) {
    constexpr int N = 8;
    const unsigned cg = C / N;
    const unsigned idx = blockIdx.x * EW_THREADS + threadIdx.x;
    unsigned pix = idx / cg;
    const int cv = idx - pix * cg, c = cv * N;
    const bool ok = pix < npix;
    if (!ok) pix = npix - 1;
    const Vf<N> z = vrelu(vfma(ldv8(raw + (size_t)pix * C + c), ldp<N>(scale + c), ldp<N>(shift + c)));
    const int n = pix / HW, p = pix - n * HW;
    for (int k = 0; k < K; ++k) {
        const float d = group_sum(vdot(z, ldp<N>(w + k * C + c)), cg);
        if (ok && cv == 0) logits[((size_t)n * K + k) * HW + p] = d + b[k];
    }
}
void k_final_fwd(cudaStream_t st, const Tensor& raw, const float* scale, const float* shift, const float* w, const float* b,
                 int K, float* logits) {
    SaltProfScope prof_scope(SALT_PROF_OTHER, (double)raw.bytes(), st);
    SALT_COUNT(1);
    const unsigned npix = (unsigned)raw.B * raw.H * raw.W;
    const int cg = raw.C / 8;
    if (raw.C % 8 || cg > 32 || (cg & (cg - 1))) throw std::runtime_error("final 1x1 conv: channel count must be 8 * 2^k <= 256");
    if (k_ring_final_fwd(st, raw, scale, shift, w, b, K, logits)) return;
    SALT_DISPATCH(raw.dt, T, {
        final_fwd_kernel<T><<<cdiv((long long)npix * cg, EW_THREADS), EW_THREADS, 0, st>>>((const T*)raw.p, scale, shift, w, b, K, logits,
                                                                                       npix, raw.H * raw.W, raw.C);
    });
}
template <typename T, int K>
__global__ void final_bwd_kernel(const float* __restrict__ dlogits, const T* __restrict__ raw, BNRef bn,
                                 const float* __restrict__ w, float* __restrict__ dw, float* __restrict__ db,
                                 T* __restrict__ gbn, unsigned npix, int HW, int C) {
    constexpr int N = 8;
    __shared__ float red[N * EW_THREADS];
    const int cg = C / N, lanes = EW_THREADS / cg;
    const int cv = threadIdx.x % cg, lane = threadIdx.x / cg, c = cv * N;
    const Vf<N> sc = ldp<N>(bn.scale + c), sh = ldp<N>(bn.shift + c), mu = ldp<N>(bn.mean + c), is = ldp<N>(bn.invstd + c);
    Vf<N> wk[K], sdw[K];
    float sdb[K];
    for (int k = 0; k < K; ++k) { wk[k] = ldp<N>(w + k * C + c); sdw[k] = vzero<N>(); sdb[k] = 0.f; }
    Vf<N> sg = vzero<N>(), sgx = vzero<N>();
    const unsigned pstep = gridDim.x * lanes;
    for (unsigned pix0 = blockIdx.x * lanes + lane; pix0 < npix; pix0 += 2 * pstep) {
      Vf<N> xs[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) xs[u] = ldv8(raw + (size_t)min(pix0 + u * pstep, npix - 1) * C + c);      // both loads in flight
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const unsigned pix = pix0 + u * pstep;
        if (pix >= npix) break;
        const int n = pix / HW, p = pix - n * HW;
        const Vf<N> x = xs[u];
        const Vf<N> z = vrelu(vfma(x, sc, sh));
        Vf<N> gz = vzero<N>();
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float dl = dlogits[((size_t)n * K + k) * HW + p];
            gz = vaxpy(wk[k], dl, gz);
            sdw[k] = vaxpy(z, dl, sdw[k]);
            if (cv == 0) sdb[k] += dl;
        }
        gz = vmaskpos(gz, z);
        stv8(gbn + (size_t)pix * C + c, gz);
        sg = vadd(sg, gz);
        sgx = vfma(gz, vxhat(x, mu, is), sgx);
      }
    }
    block_reduce_slot<N>(sg, cg, bn.bsums + (size_t)blockIdx.x * 2 * C, red);
    block_reduce_slot<N>(sgx, cg, bn.bsums + (size_t)blockIdx.x * 2 * C + C, red);
    if (blockIdx.x == 0 && threadIdx.x == 0) *bn.bslots = (int)gridDim.x;
    for (int k = 0; k < K; ++k) block_reduce_add<N, float>(sdw[k], cg, dw + k * C, red);
    for (int k = 0; k < K; ++k) {
        red[threadIdx.x] = sdb[k];
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.f;
            for (int i = 0; i < EW_THREADS; ++i) t += red[i];
            atomicAdd(db + k, t);
        }
        __syncthreads();
    }
}
void k_final_bwd(cudaStream_t st, const float* dlogits, const Tensor& raw, const BNRef& bn, const float* w, int K,
                 float* dw, float* db, const Tensor& gbn) {
    SaltProfScope prof_scope(SALT_PROF_OTHER, 2.0 * (double)raw.bytes(), st);
    SALT_COUNT(1);
    const unsigned npix = (unsigned)raw.B * raw.H * raw.W;
    const int HW = raw.H * raw.W;
    if (k_ring_final_bwd(st, dlogits, raw, bn, w, K, dw, db, gbn)) return;
#define LAUNCH_FB(KK) final_bwd_kernel<T, KK><<<blocks, EW_THREADS, 0, st>>>(dlogits, (const T*)raw.p, bn, w, dw, db, (T*)gbn.p, npix, HW, raw.C)
    SALT_DISPATCH(raw.dt, T, {
        const int blocks = reduce_blocks(npix, raw.C / 8);
        if (K == 1) LAUNCH_FB(1); else if (K == 2) LAUNCH_FB(2); else if (K == 3) LAUNCH_FB(3); else LAUNCH_FB(4);
    });
#undef LAUNCH_FB
}

// ------------------------------------------------------------------------------------------------
// ReLU / BatchNorm backward
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void relu_mask_inplace_kernel(T* __restrict__ g, const T* __restrict__ mask, unsigned nvec) {
    constexpr int N = VW<T>::N;
    const unsigned idx = blockIdx.x * EW_THREADS + threadIdx.x;
    if (idx >= nvec) return;
    stv(g + (size_t)idx * N, vmaskpos(ldv(g + (size_t)idx * N), ldv(mask + (size_t)idx * N)));
}
void k_relu_mask_inplace(cudaStream_t st, const Tensor& g, const Tensor& mask) {
    SaltProfScope prof_scope(SALT_PROF_OTHER, 3.0 * (double)g.bytes(), st);
    SALT_COUNT(1);
    if (k_ring_relu_mask(st, g, mask)) return;
    SALT_DISPATCH(g.dt, T, {
        const unsigned nvec = (unsigned)(g.numel() / VW<T>::N);
        relu_mask_inplace_kernel<T><<<cdiv(nvec, EW_THREADS), EW_THREADS, 0, st>>>((T*)g.p, (const T*)mask.p, nvec);
    });
}
// upstream gradient seen by a BN layer: g, optionally gated per (image, channel): g*gate[n][c] + addc[n][c] (encoder SE),
// optionally masked by the layer's own ReLU.  grid = (pixel blocks, channel slabs)
template <typename T>
__global__ void __launch_bounds__(EW_THREADS, 2) bn_bwd_reduce_kernel(const T* __restrict__ g, const T* __restrict__ raw, BNRef bn, int self_mask,
                                     const float* __restrict__ gate, const float* __restrict__ addc, unsigned npix, int HW, int C) {
    constexpr int N = VW<T>::N;
    __shared__ float red[N * EW_THREADS];
    const int cg = min(C / N, EW_THREADS), lanes = EW_THREADS / cg;
    const int cv = blockIdx.y * cg + threadIdx.x % cg, lane = threadIdx.x / cg, c = cv * N;
    const Vf<N> sc = ldp<N>(bn.scale + c), sh = ldp<N>(bn.shift + c), mu = ldp<N>(bn.mean + c), is = ldp<N>(bn.invstd + c);
    Vf<N> sg = vzero<N>(), sgx = vzero<N>();
    auto accumulate = [&](unsigned pix, const Vf<N>& x, Vf<N> gv) {
        if (gate) {
            const size_t o = (size_t)(pix / HW) * C + c;
            gv = vfma(gv, ldp<N>(gate + o), ldp<N>(addc + o));
        }
        if (self_mask) gv = vmaskpos(gv, vfma(x, sc, sh));
        sg = vadd(sg, gv);
        sgx = vfma(gv, vxhat(x, mu, is), sgx);
    };
    // 2 pixels = 4 independent 16-byte loads in flight per thread, 3 blocks per SM (a 4-pixel version needed 149 registers: one
    // block per SM, no faster; the one-pixel loop with <= 2 blocks per SM ran at 1.9 TB/s - profiles/r2_notes.md)
    const unsigned step = gridDim.x * lanes;
    unsigned pix = blockIdx.x * lanes + lane;
    for (; pix + step < npix; pix += 2 * step) {
        Vf<N> xs[2], gs[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) { xs[u] = ldv(raw + (size_t)(pix + u * step) * C + c); gs[u] = ldv(g + (size_t)(pix + u * step) * C + c); }
#pragma unroll
        for (int u = 0; u < 2; ++u) accumulate(pix + u * step, xs[u], gs[u]);
    }
    for (; pix < npix; pix += step) accumulate(pix, ldv(raw + (size_t)pix * C + c), ldv(g + (size_t)pix * C + c));
    const int coff = blockIdx.y * cg * N;
    block_reduce_slot<N>(sg, cg, bn.bsums + (size_t)blockIdx.x * 2 * C + coff, red);
    block_reduce_slot<N>(sgx, cg, bn.bsums + (size_t)blockIdx.x * 2 * C + C + coff, red);
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *bn.bslots = (int)gridDim.x;
}
void k_bn_bwd_reduce(cudaStream_t st, const Tensor& g, const Tensor& raw, const BNRef& bn, bool self_mask, const float* gate,
                     const float* addc) {
    SaltProfScope prof_scope(SALT_PROF_BN_BWD_REDUCE, 2.0 * (double)raw.bytes(), st);
    SALT_COUNT(1);
    if (k_ring_bn_bwd_reduce(st, g, raw, bn, self_mask, gate, addc)) return;
    const unsigned npix = (unsigned)raw.B * raw.H * raw.W;
    SALT_DISPATCH(raw.dt, T, {
        const int cgt = raw.C / VW<T>::N, cg = std::min(cgt, EW_THREADS);
        dim3 grid(reduce_blocks(npix, cg), cgt / cg);
        bn_bwd_reduce_kernel<T><<<grid, EW_THREADS, 0, st>>>((const T*)g.p, (const T*)raw.p, bn, self_mask ? 1 : 0, gate, addc, npix,
                                                             raw.H * raw.W, raw.C);
    });
}
template <typename T>
__global__ void __launch_bounds__(EW_THREADS, 2) bn_bwd_apply_kernel(const T* __restrict__ g, const T* __restrict__ raw, BNRef bn, int self_mask,
                                    const float* __restrict__ gate, const float* __restrict__ addc, T* __restrict__ graw,
                                    unsigned npix, int HW, int C) {
    // a thread keeps ONE channel group (5 per-channel coefficient vectors in registers) and strides over pixels
    constexpr int N = VW<T>::N;
    const int cg = min(C / N, EW_THREADS), lanes = EW_THREADS / cg;
    const int cv = blockIdx.y * cg + threadIdx.x % cg, lane = threadIdx.x / cg, c = cv * N;
    const Vf<N> sc = ldp<N>(bn.scale + c), sh = ldp<N>(bn.shift + c), mu = ldp<N>(bn.mean + c), cb = ldp<N>(bn.cb + c),
                cc = ldp<N>(bn.cc + c);
    auto finish = [&](unsigned pix, const Vf<N>& x, Vf<N> gv) {
        if (gate) {
            const size_t o = (size_t)(pix / HW) * C + c;
            gv = vfma(gv, ldp<N>(gate + o), ldp<N>(addc + o));
        }
        if (self_mask) gv = vmaskpos(gv, vfma(x, sc, sh));
        Vf<N> r;
#pragma unroll
        for (int i = 0; i < N; ++i) r.v[i] = sc.v[i] * (gv.v[i] - cb.v[i] - cc.v[i] * (x.v[i] - mu.v[i]));
        stv(graw + (size_t)pix * C + c, r);
    };
    const unsigned step = gridDim.x * lanes;
    unsigned pix = blockIdx.x * lanes + lane;
    for (; pix + step < npix; pix += 2 * step) {           // 4 independent 16-byte loads in flight per thread, 3 blocks per SM
        Vf<N> xs[2], gs[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) { xs[u] = ldv(raw + (size_t)(pix + u * step) * C + c); gs[u] = ldv(g + (size_t)(pix + u * step) * C + c); }
#pragma unroll
        for (int u = 0; u < 2; ++u) finish(pix + u * step, xs[u], gs[u]);
    }
    for (; pix < npix; pix += step) finish(pix, ldv(raw + (size_t)pix * C + c), ldv(g + (size_t)pix * C + c));
}
void k_bn_bwd_apply(cudaStream_t st, const Tensor& g, const Tensor& raw, const BNRef& bn, bool self_mask, const Tensor& graw,
                    const float* gate, const float* addc) {
    SaltProfScope prof_scope(SALT_PROF_BN_BWD_APPLY, 3.0 * (double)raw.bytes(), st);
    SALT_COUNT(1);
    if (k_ring_bn_bwd_apply(st, g, raw, bn, self_mask, graw, gate, addc)) return;
    SALT_DISPATCH(raw.dt, T, {
        const unsigned npix = (unsigned)raw.B * raw.H * raw.W;
        const int cgt = raw.C / VW<T>::N, cg = std::min(cgt, EW_THREADS), lanes = EW_THREADS / cg;
        const int blocks = (int)std::min<long long>(((long long)npix + lanes * 4 - 1) / (lanes * 4), 148 * 16);
        bn_bwd_apply_kernel<T><<<dim3(blocks, cgt / cg), EW_THREADS, 0, st>>>((const T*)g.p, (const T*)raw.p, bn, self_mask ? 1 : 0, gate,
                                                                             addc, (T*)graw.p, npix, raw.H * raw.W, raw.C);
    });
}

// ------------------------------------------------------------------------------------------------
// split-bf16 operands of the fp32 tensor-core parity mode (kernels.h)
// ------------------------------------------------------------------------------------------------
struct Split6Order { int part[6]; };        // which term (0 = h, 1 = m, 2 = l) goes into channel segment i
__global__ void split6_kernel(const float* __restrict__ in, bf16* __restrict__ out, size_t nvec, int C, Split6Order ord) {
    const size_t idx = (size_t)blockIdx.x * EW_THREADS + threadIdx.x;          // one 4-channel vector
    if (idx >= nvec) return;
    const int cg = C >> 2;
    const size_t row = idx / cg;
    const int c = (int)(idx - row * cg) * 4;
    const float4 v = *reinterpret_cast<const float4*>(in + row * C + c);
    const float x[4] = {v.x, v.y, v.z, v.w};
    __nv_bfloat16 t[3][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat16 h = __float2bfloat16_rn(x[i]);
        const float r1 = x[i] - __bfloat162float(h);                            // exact in fp32
        const __nv_bfloat16 m = __float2bfloat16_rn(r1);
        const float r2 = r1 - __bfloat162float(m);
        t[0][i] = h; t[1][i] = m; t[2][i] = __float2bfloat16_rn(r2);
    }
    bf16* o = out + row * (size_t)(6 * C) + c;
#pragma unroll
    for (int sgm = 0; sgm < 6; ++sgm) {
        const __nv_bfloat16* p = t[ord.part[sgm]];
        uint2 u;
        __nv_bfloat162 a = __halves2bfloat162(p[0], p[1]), b = __halves2bfloat162(p[2], p[3]);
        u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(o + (size_t)sgm * C) = u;
    }
}
static void launch_split6(cudaStream_t st, const float* in, void* out, size_t rows, int C, const Split6Order& ord) {
    if (C % 4) throw std::runtime_error("split6: channel count must be a multiple of 4");
    SALT_COUNT(1);
    const size_t nvec = rows * (size_t)(C / 4);
    if (nvec == 0) return;
    split6_kernel<<<(unsigned)((nvec + EW_THREADS - 1) / EW_THREADS), EW_THREADS, 0, st>>>(in, (bf16*)out, nvec, C, ord);
}
void k_split6_act(cudaStream_t st, const float* in, void* out, size_t rows, int C) {
    launch_split6(st, in, out, rows, C, Split6Order{{0, 1, 0, 1, 0, 2}});
}
void k_split6_weights(cudaStream_t st, const float* wp, void* wp6, size_t rows, int C) {
    launch_split6(st, wp, wp6, rows, C, Split6Order{{0, 1, 1, 0, 2, 0}});
}

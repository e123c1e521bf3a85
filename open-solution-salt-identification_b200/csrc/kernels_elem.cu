// Elementwise / reduction kernels of the U-Net engine: BatchNorm (forward + backward), residual add,
// replicate-border writers, bilinear gather (upsample + virtual concat -> bordered conv input) and its
// adjoint, scSE forward/backward, the final 1x1 conv.  All NHWC, storage type T in {float, bf16},
// arithmetic fp32.  These are HBM-bound passes; vectors of 4 channels per thread, fully coalesced.
#include "kernels.h"

#define EW_THREADS 256

// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------
// Sum a per-thread float4 over the `lanes` threads that share a channel group and add it (as double)
// to dst[c..c+3].  Thread layout: cv = tid % cg (channel group), lane = tid / cg.
__device__ __forceinline__ void block_reduce_to_double(float4 v, int cg, double* dst, float4* red) {
    const int tid = threadIdx.x;
    red[tid] = v;
    __syncthreads();
    if (tid < cg) {
        float4 s = red[tid];
        for (int t = tid + cg; t < EW_THREADS; t += cg) s = f4_add(s, red[t]);
        atomicAdd(dst + tid * 4 + 0, (double)s.x);
        atomicAdd(dst + tid * 4 + 1, (double)s.y);
        atomicAdd(dst + tid * 4 + 2, (double)s.z);
        atomicAdd(dst + tid * 4 + 3, (double)s.w);
    }
    __syncthreads();
}
__device__ __forceinline__ void block_reduce_to_float(float4 v, int cg, float* dst, float4* red) {
    const int tid = threadIdx.x;
    red[tid] = v;
    __syncthreads();
    if (tid < cg) {
        float4 s = red[tid];
        for (int t = tid + cg; t < EW_THREADS; t += cg) s = f4_add(s, red[t]);
        atomicAdd(dst + tid * 4 + 0, s.x);
        atomicAdd(dst + tid * 4 + 1, s.y);
        atomicAdd(dst + tid * 4 + 2, s.z);
        atomicAdd(dst + tid * 4 + 3, s.w);
    }
    __syncthreads();
}

static inline int reduce_blocks(long long npix, int cg) {
    int lanes = EW_THREADS / cg;
    long long b = (npix + (long long)lanes * 8 - 1) / ((long long)lanes * 8);
    if (b > 148 * 8) b = 148 * 8;
    if (b < 1) b = 1;
    return (int)b;
}

// ------------------------------------------------------------------------------------------------
// input adapter: fp32 NCHW [B,3,H,W] -> NHWC with C padded to 4
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void input_nchw_to_nhwc4_kernel(const float* __restrict__ x, T* __restrict__ out, int B, int HW) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)B * HW) return;
    int n = (int)(idx / HW), p = (int)(idx % HW);
    const float* xp = x + (size_t)n * 3 * HW + p;
    st4(out + idx * 4, make_float4(xp[0], xp[HW], xp[2 * (size_t)HW], 0.f));
}
void k_input_nchw_to_nhwc4(cudaStream_t st, DType dt, const float* x, void* out, int B, int H, int W) {
    SALT_COUNT(1);
    long long n = (long long)B * H * W;
    SALT_DISPATCH(dt, T, (input_nchw_to_nhwc4_kernel<T><<<cdiv(n, 256), 256, 0, st>>>(x, (T*)out, B, H * W)));
}
void k_zero(cudaStream_t st, void* p, size_t bytes) { cudaMemsetAsync(p, 0, bytes, st); }

// ------------------------------------------------------------------------------------------------
// BatchNorm finalisation (nn.BatchNorm2d: eps 1e-5, momentum 0.1, biased var to normalise, unbiased to track)
// ------------------------------------------------------------------------------------------------
__global__ void bn_finalize_train_kernel(BNRef bn, double count, float momentum, float eps) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= bn.C) return;
    double mean = bn.sums[c] / count;
    double var = bn.sums[bn.C + c] / count - mean * mean;
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
void k_bn_finalize_train(cudaStream_t st, const BNRef& bn, double count, float momentum, float eps) {
    SALT_COUNT(1);
    bn_finalize_train_kernel<<<cdiv(bn.C, 128), 128, 0, st>>>(bn, count, momentum, eps);
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
    SALT_COUNT(1);
    bn_finalize_eval_kernel<<<cdiv(bn.C, 128), 128, 0, st>>>(bn, eps);
}
__global__ void bn_bwd_finalize_kernel(BNRef bn, double count) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= bn.C) return;
    double sg = bn.bsums[c], sgx = bn.bsums[bn.C + c];
    bn.dbeta[c] += (float)sg;
    bn.dgamma[c] += (float)sgx;
    bn.cb[c] = (float)(sg / count);
    bn.cc[c] = (float)(sgx / count * (double)bn.invstd[c]);
}
void k_bn_bwd_finalize(cudaStream_t st, const BNRef& bn, double count) {
    SALT_COUNT(1);
    bn_bwd_finalize_kernel<<<cdiv(bn.C, 128), 128, 0, st>>>(bn, count);
}

// ------------------------------------------------------------------------------------------------
// BN apply (+ residual) (+ ReLU), optional replicate border on the output
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void bn_apply_kernel(const T* __restrict__ raw, const float* __restrict__ scale, const float* __restrict__ shift,
                                const T* __restrict__ res, const float* __restrict__ rscale, const float* __restrict__ rshift,
                                int relu, T* __restrict__ out, int B, int H, int W, int C, int pt, int pb, int pl, int pr) {
    const int cg = C >> 2;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)B * H * W * cg;
    if (idx >= total) return;
    int cv = (int)(idx % cg);
    long long pix = idx / cg;
    int x = (int)(pix % W);
    int y = (int)((pix / W) % H);
    int n = (int)(pix / ((long long)W * H));
    int c = cv * 4;
    float4 v = f4_fma(ld4(raw + pix * C + c), ld4(scale + c), ld4(shift + c));
    if (res) {
        float4 r = ld4(res + pix * C + c);
        if (rscale) r = f4_fma(r, ld4(rscale + c), ld4(rshift + c));
        v = f4_add(v, r);
    }
    if (relu) v = f4_relu(v);
    const int Hp = H + pt + pb, Wp = W + pl + pr;
    int y0 = (y == 0) ? 0 : y + pt, y1 = (y == H - 1) ? Hp - 1 : y + pt;
    int x0 = (x == 0) ? 0 : x + pl, x1 = (x == W - 1) ? Wp - 1 : x + pl;
    for (int yy = y0; yy <= y1; ++yy)
        for (int xx = x0; xx <= x1; ++xx) st4(out + (((size_t)n * Hp + yy) * Wp + xx) * C + c, v);
}
void k_bn_apply(cudaStream_t st, const Tensor& raw, const float* scale, const float* shift, const Tensor* res,
                const float* rscale, const float* rshift, bool relu, const Tensor& out) {
    SALT_COUNT(1);
    long long total = (long long)raw.B * raw.H * raw.W * (raw.C / 4);
    SALT_DISPATCH(raw.dt, T, (bn_apply_kernel<T><<<cdiv(total, EW_THREADS), EW_THREADS, 0, st>>>(
        (const T*)raw.p, scale, shift, res ? (const T*)res->p : nullptr, rscale, rshift, relu ? 1 : 0, (T*)out.p,
        raw.B, raw.H, raw.W, raw.C, out.pt, out.pb, out.pl, out.pr)));
}

template <typename T>
__global__ void bn_relu_avgpool_kernel(const T* __restrict__ raw, const float* __restrict__ scale,
                                       const float* __restrict__ shift, T* __restrict__ out, int B, int Ho, int Wo, int C) {
    const int cg = C >> 2;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)B * Ho * Wo * cg) return;
    int cv = (int)(idx % cg);
    long long pix = idx / cg;
    int x = (int)(pix % Wo), y = (int)((pix / Wo) % Ho), n = (int)(pix / ((long long)Wo * Ho));
    int c = cv * 4, Wi = Wo * 2, Hi = Ho * 2;
    float4 sc = ld4(scale + c), sh = ld4(shift + c), acc = f4_zero();
    for (int dy = 0; dy < 2; ++dy)
        for (int dx = 0; dx < 2; ++dx) {
            size_t o = (((size_t)n * Hi + 2 * y + dy) * Wi + 2 * x + dx) * C + c;
            acc = f4_add(acc, f4_relu(f4_fma(ld4(raw + o), sc, sh)));
        }
    st4(out + pix * C + c, f4_scale(acc, 0.25f));
}
void k_bn_relu_avgpool(cudaStream_t st, const Tensor& raw, const float* scale, const float* shift, const Tensor& out) {
    SALT_COUNT(1);
    long long total = (long long)out.B * out.H * out.W * (out.C / 4);
    SALT_DISPATCH(raw.dt, T, (bn_relu_avgpool_kernel<T><<<cdiv(total, EW_THREADS), EW_THREADS, 0, st>>>(
        (const T*)raw.p, scale, shift, (T*)out.p, out.B, out.H, out.W, out.C)));
}
template <typename T>
__global__ void avgpool_bwd_kernel(const T* __restrict__ gout, T* __restrict__ gin, int B, int Hi, int Wi, int C) {
    const int cg = C >> 2;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)B * Hi * Wi * cg) return;
    int cv = (int)(idx % cg);
    long long pix = idx / cg;
    int x = (int)(pix % Wi), y = (int)((pix / Wi) % Hi), n = (int)(pix / ((long long)Wi * Hi));
    size_t o = (((size_t)n * (Hi / 2) + y / 2) * (Wi / 2) + x / 2) * C + cv * 4;
    st4(gin + pix * C + cv * 4, f4_scale(ld4(gout + o), 0.25f));
}
void k_avgpool_bwd(cudaStream_t st, const Tensor& gout, const Tensor& gin) {
    SALT_COUNT(1);
    long long total = (long long)gin.B * gin.H * gin.W * (gin.C / 4);
    SALT_DISPATCH(gin.dt, T, (avgpool_bwd_kernel<T><<<cdiv(total, EW_THREADS), EW_THREADS, 0, st>>>(
        (const T*)gout.p, (T*)gin.p, gin.B, gin.H, gin.W, gin.C)));
}

// ------------------------------------------------------------------------------------------------
// gather: bordered conv input = concat_c( bilinear_upsample_f(src_i) ), replicate border
//   bilinear: align_corners=False, src = max((d+0.5)/f - 0.5, 0)  (torch upsample_bilinear2d)
// ------------------------------------------------------------------------------------------------
struct GatherArgs { GatherSrc s[5]; int n; };

__device__ __forceinline__ void bilin_coord(int d, int f, int nsrc, int& i0, int& i1, float& l) {
    float s = ((float)d + 0.5f) * (1.0f / (float)f) - 0.5f;
    s = fmaxf(s, 0.f);
    i0 = (int)s;
    i1 = min(i0 + 1, nsrc - 1);
    l = s - (float)i0;
}

template <typename T>
__global__ void gather_fwd_kernel(T* __restrict__ out, GatherArgs a, int B, int H, int W, int C, int pt, int pb, int pl, int pr) {
    const int cg = C >> 2, Hp = H + pt + pb, Wp = W + pl + pr;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)B * Hp * Wp * cg) return;
    int cv = (int)(idx % cg);
    long long pix = idx / cg;
    int xp = (int)(pix % Wp), yp = (int)((pix / Wp) % Hp), n = (int)(pix / ((long long)Wp * Hp));
    int y = min(max(yp - pt, 0), H - 1), x = min(max(xp - pl, 0), W - 1);
    int c = cv * 4, si = 0;
    while (si < a.n - 1 && c >= a.s[si].C) { c -= a.s[si].C; ++si; }
    const GatherSrc s = a.s[si];
    const T* sp = (const T*)s.p + (size_t)n * s.H * s.W * s.C + c;
    float4 v;
    if (s.f == 1) {
        v = ld4(sp + ((size_t)y * s.W + x) * s.C);
    } else {
        int y0, y1, x0, x1; float ly, lx;
        bilin_coord(y, s.f, s.H, y0, y1, ly);
        bilin_coord(x, s.f, s.W, x0, x1, lx);
        float4 v00 = ld4(sp + ((size_t)y0 * s.W + x0) * s.C), v01 = ld4(sp + ((size_t)y0 * s.W + x1) * s.C);
        float4 v10 = ld4(sp + ((size_t)y1 * s.W + x0) * s.C), v11 = ld4(sp + ((size_t)y1 * s.W + x1) * s.C);
        float wy0 = 1.f - ly, wx0 = 1.f - lx;
        float4 top = f4_add(f4_scale(v00, wx0), f4_scale(v01, lx));
        float4 bot = f4_add(f4_scale(v10, wx0), f4_scale(v11, lx));
        v = f4_add(f4_scale(top, wy0), f4_scale(bot, ly));
    }
    st4(out + pix * C + cv * 4, v);
}
void k_gather_fwd(cudaStream_t st, const Tensor& out, const GatherSrc* srcs, int nsrc) {
    SALT_COUNT(1);
    GatherArgs a; a.n = nsrc;
    for (int i = 0; i < nsrc; ++i) a.s[i] = srcs[i];
    long long total = (long long)out.B * out.Hp() * out.Wp() * (out.C / 4);
    SALT_DISPATCH(out.dt, T, (gather_fwd_kernel<T><<<cdiv(total, EW_THREADS), EW_THREADS, 0, st>>>(
        (T*)out.p, a, out.B, out.H, out.W, out.C, out.pt, out.pb, out.pl, out.pr)));
}

// adjoint of the replicate border for a copied (f == 1) source: gsrc[y,x] (+)= sum of the border cells mapped to it
template <typename T>
__global__ void fold_bwd_kernel(const T* __restrict__ gP, int c0, T* __restrict__ gsrc, int accumulate,
                                int B, int H, int W, int Cp, int Cs, int pt, int pb, int pl, int pr) {
    const int cg = Cs >> 2, Hp = H + pt + pb, Wp = W + pl + pr;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)B * H * W * cg) return;
    int cv = (int)(idx % cg);
    long long pix = idx / cg;
    int x = (int)(pix % W), y = (int)((pix / W) % H), n = (int)(pix / ((long long)W * H));
    int y0 = (y == 0) ? 0 : y + pt, y1 = (y == H - 1) ? Hp - 1 : y + pt;
    int x0 = (x == 0) ? 0 : x + pl, x1 = (x == W - 1) ? Wp - 1 : x + pl;
    float4 acc = f4_zero();
    for (int yy = y0; yy <= y1; ++yy)
        for (int xx = x0; xx <= x1; ++xx)
            acc = f4_add(acc, ld4(gP + (((size_t)n * Hp + yy) * Wp + xx) * Cp + c0 + cv * 4));
    T* o = gsrc + pix * Cs + cv * 4;
    if (accumulate) acc = f4_add(acc, ld4(o));
    st4(o, acc);
}
void k_fold_bwd(cudaStream_t st, const Tensor& gP, int c0, const Tensor& gsrc, bool accumulate) {
    SALT_COUNT(1);
    long long total = (long long)gsrc.B * gsrc.H * gsrc.W * (gsrc.C / 4);
    SALT_DISPATCH(gP.dt, T, (fold_bwd_kernel<T><<<cdiv(total, EW_THREADS), EW_THREADS, 0, st>>>(
        (const T*)gP.p, c0, (T*)gsrc.p, accumulate ? 1 : 0, gP.B, gP.H, gP.W, gP.C, gsrc.C, gP.pt, gP.pb, gP.pl, gP.pr)));
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
template <typename T>
__global__ void upsample_bwd_x_kernel(const T* __restrict__ gP, int c0, float* __restrict__ tmp, int B, int H, int W,
                                      int Cp, int Cs, int pt, int pb, int pl, int pr, int f) {
    const int cg = Cs >> 2, Hp = H + pt + pb, Wp = W + pl + pr, Ws = W / f;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)B * Hp * Ws * cg) return;
    int cv = (int)(idx % cg);
    long long r = idx / cg;
    int jx = (int)(r % Ws);
    long long row = r / Ws;                                // n*Hp + yp
    int lo, hi;
    adj_range(jx, f, pl, W, Wp, lo, hi);
    float4 acc = f4_zero();
    for (int xp = lo; xp < hi; ++xp) {
        float w = bilin_adj_w(xp, pl, W, f, Ws, jx);
        if (w != 0.f) acc = f4_fma(ld4(gP + ((size_t)row * Wp + xp) * Cp + c0 + cv * 4), make_float4(w, w, w, w), acc);
    }
    st4(tmp + idx * 4, acc);
}
template <typename T>
__global__ void upsample_bwd_y_kernel(const float* __restrict__ tmp, T* __restrict__ gsrc, int accumulate, int B, int H,
                                      int W, int Cs, int pt, int pb, int f) {
    const int cg = Cs >> 2, Hp = H + pt + pb, Ws = W / f, Hs = H / f;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)B * Hs * Ws * cg) return;
    int cv = (int)(idx % cg);
    long long r = idx / cg;
    int jx = (int)(r % Ws), jy = (int)((r / Ws) % Hs), n = (int)(r / ((long long)Ws * Hs));
    int lo, hi;
    adj_range(jy, f, pt, H, Hp, lo, hi);
    float4 acc = f4_zero();
    for (int yp = lo; yp < hi; ++yp) {
        float w = bilin_adj_w(yp, pt, H, f, Hs, jy);
        if (w != 0.f) acc = f4_fma(ld4(tmp + ((((size_t)n * Hp + yp) * Ws + jx) * cg + cv) * 4), make_float4(w, w, w, w), acc);
    }
    T* o = gsrc + idx * 4;
    if (accumulate) acc = f4_add(acc, ld4(o));
    st4(o, acc);
}
size_t upsample_bwd_tmp_floats(const Tensor& gP, int f, int Csrc) {
    return (size_t)gP.B * gP.Hp() * (gP.W / f) * Csrc;
}
void k_upsample_bwd(cudaStream_t st, const Tensor& gP, int c0, int f, const Tensor& gsrc, float* tmp, bool accumulate) {
    SALT_COUNT(2);
    long long t1 = (long long)gP.B * gP.Hp() * (gP.W / f) * (gsrc.C / 4);
    long long t2 = (long long)gsrc.B * gsrc.H * gsrc.W * (gsrc.C / 4);
    SALT_DISPATCH(gP.dt, T, {
        upsample_bwd_x_kernel<T><<<cdiv(t1, EW_THREADS), EW_THREADS, 0, st>>>((const T*)gP.p, c0, tmp, gP.B, gP.H, gP.W,
                                                                          gP.C, gsrc.C, gP.pt, gP.pb, gP.pl, gP.pr, f);
        upsample_bwd_y_kernel<T><<<cdiv(t2, EW_THREADS), EW_THREADS, 0, st>>>(tmp, (T*)gsrc.p, accumulate ? 1 : 0, gP.B,
                                                                          gP.H, gP.W, gsrc.C, gP.pt, gP.pb, f);
    });
}

// ------------------------------------------------------------------------------------------------
// scSE (base.py:82-117).  z = relu(bn(raw));  out = relu(z*cse[n,c] + z*sse[n,y,x]) = z*(cse+sse)
// ------------------------------------------------------------------------------------------------
// per-(n,c) sum over pixels of  [g *] relu(raw*scale+shift)
template <typename T, bool WITH_G>
__global__ void scse_pool_kernel(const T* __restrict__ raw, const T* __restrict__ g, const float* __restrict__ scale,
                                 const float* __restrict__ shift, float* __restrict__ part, int HW, int C) {
    // writes part[n][chunk][C]; the FC kernels add the chunks in a fixed order (deterministic, batch independent)
    __shared__ float4 red[EW_THREADS];
    const int cg = C >> 2, lanes = EW_THREADS / cg;
    const int cv = threadIdx.x % cg, lane = threadIdx.x / cg, n = blockIdx.y;
    const int chunk = (HW + gridDim.x - 1) / gridDim.x;
    const int p0 = blockIdx.x * chunk, p1 = min(HW, p0 + chunk);
    float4 sc = ld4(scale + cv * 4), sh = ld4(shift + cv * 4), acc = f4_zero();
    for (int p = p0 + lane; p < p1; p += lanes) {
        size_t o = ((size_t)n * HW + p) * C + cv * 4;
        float4 z = f4_relu(f4_fma(ld4(raw + o), sc, sh));
        if (WITH_G) z = f4_mul(z, ld4(g + o));
        acc = f4_add(acc, z);
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x < cg) {
        float4 s = red[threadIdx.x];
        for (int t = threadIdx.x + cg; t < EW_THREADS; t += cg) s = f4_add(s, red[t]);
        st4(part + ((size_t)n * gridDim.x + blockIdx.x) * C + threadIdx.x * 4, s);
    }
}
// one block per image: squeeze -> fc(C->Cr) -> relu -> fc(Cr->C) -> sigmoid
__global__ void scse_fc_kernel(SERef se, float inv_hw) {
    extern __shared__ float sm[];
    float* gap = sm;               // [C]
    float* hid = sm + se.C;        // [Cr]
    const int n = blockIdx.x, c = threadIdx.x;
    if (c < se.C) {
        float t = 0.f;
        for (int k = 0; k < se.chunks; ++k) t += se.part[((size_t)n * se.chunks + k) * se.C + c];
        gap[c] = t * inv_hw;
        se.gap[(size_t)n * se.C + c] = gap[c];
    }
    __syncthreads();
    if (c < se.Cr) {
        float a = se.b1[c];
        for (int i = 0; i < se.C; ++i) a = fmaf(se.w1[c * se.C + i], gap[i], a);
        a = fmaxf(a, 0.f);
        hid[c] = a;
        se.hid[(size_t)n * se.Cr + c] = a;
    }
    __syncthreads();
    if (c < se.C) {
        float a = se.b2[c];
        for (int j = 0; j < se.Cr; ++j) a = fmaf(se.w2[c * se.Cr + j], hid[j], a);
        se.cse[(size_t)n * se.C + c] = 1.f / (1.f + expf(-a));
    }
}
// sum over the cg (power of two, <= 32) lanes that share one pixel
__device__ __forceinline__ float group_sum(float v, int cg) {
    for (int o = cg >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <typename T>
__global__ void scse_apply_kernel(const T* __restrict__ raw, const float* __restrict__ scale, const float* __restrict__ shift,
                                  SERef se, T* __restrict__ out, long long npix, int HW, int C) {
    const int cg = C >> 2;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int cv = (int)(idx % cg);
    long long pix = idx / cg;
    bool ok = pix < npix;
    if (!ok) pix = npix - 1;
    int n = (int)(pix / HW), c = cv * 4;
    float4 z = f4_relu(f4_fma(ld4(raw + pix * C + c), ld4(scale + c), ld4(shift + c)));
    float4 w = ld4(se.ws + c);
    float dot = group_sum(z.x * w.x + z.y * w.y + z.z * w.z + z.w * w.w, cg);
    float s = 1.f / (1.f + expf(-(dot + se.bs[0])));
    float4 gate = ld4(se.cse + (size_t)n * C + c);
    gate = make_float4(gate.x + s, gate.y + s, gate.z + s, gate.w + s);
    if (ok) st4(out + pix * C + c, f4_relu(f4_mul(z, gate)));
}
void k_scse_fwd(cudaStream_t st, const Tensor& raw, const float* scale, const float* shift, const SERef& se, const Tensor& out) {
    SALT_COUNT(3);
    const int HW = raw.H * raw.W, C = raw.C, cg = C / 4;
    const int chunks = se.chunks;
    long long npix = (long long)raw.B * HW;
    SALT_DISPATCH(raw.dt, T, {
        scse_pool_kernel<T, false><<<dim3(chunks, raw.B), EW_THREADS, 0, st>>>((const T*)raw.p, nullptr, scale, shift, se.part, HW, C);
        scse_fc_kernel<<<raw.B, max(32, C), sizeof(float) * (C + se.Cr), st>>>(se, 1.0f / HW);
        scse_apply_kernel<T><<<cdiv(npix * cg, EW_THREADS), EW_THREADS, 0, st>>>((const T*)raw.p, scale, shift, se, (T*)out.p, npix, HW, C);
    });
}

// backward of the two tiny FCs, one block per image
__global__ void scse_fc_bwd_kernel(SERef se, float inv_hw) {
    extern __shared__ float sm[];
    float* dpre2 = sm;             // [C]
    float* dhid = sm + se.C;       // [Cr]
    const int n = blockIdx.x, c = threadIdx.x;
    if (c < se.C) {
        float cs = se.cse[(size_t)n * se.C + c];
        float t = 0.f;
        for (int k = 0; k < se.chunks; ++k) t += se.part[((size_t)n * se.chunks + k) * se.C + c];
        float d = t * cs * (1.f - cs);
        dpre2[c] = d;
        atomicAdd(se.db2 + c, d);
        for (int j = 0; j < se.Cr; ++j) atomicAdd(se.dw2 + c * se.Cr + j, d * se.hid[(size_t)n * se.Cr + j]);
    }
    __syncthreads();
    if (c < se.Cr) {
        float a = 0.f;
        for (int i = 0; i < se.C; ++i) a = fmaf(se.w2[i * se.Cr + c], dpre2[i], a);
        a = se.hid[(size_t)n * se.Cr + c] > 0.f ? a : 0.f;
        dhid[c] = a;
        atomicAdd(se.db1 + c, a);
    }
    __syncthreads();
    if (c < se.C) {
        float gsum = 0.f, gp = se.gap[(size_t)n * se.C + c];
        for (int j = 0; j < se.Cr; ++j) {
            atomicAdd(se.dw1 + j * se.C + c, dhid[j] * gp);
            gsum = fmaf(se.w1[j * se.C + c], dhid[j], gsum);
        }
        se.G[(size_t)n * se.C + c] = gsum * inv_hw;
    }
}
template <typename T>
__global__ void scse_bwd_apply_kernel(const T* __restrict__ gout, const T* __restrict__ raw, BNRef bn, SERef se,
                                      T* __restrict__ gbn, long long npix, int HW, int C) {
    __shared__ float4 red[EW_THREADS];
    const int cg = C >> 2, lanes = EW_THREADS / cg;
    const int cv = threadIdx.x % cg, lane = threadIdx.x / cg, c = cv * 4;
    const float4 sc = ld4(bn.scale + c), sh = ld4(bn.shift + c), mu = ld4(bn.mean + c), is = ld4(bn.invstd + c), w = ld4(se.ws + c);
    const float bs = se.bs[0];
    float4 sg = f4_zero(), sgx = f4_zero(), sws = f4_zero();
    float sbs = 0.f;
    // all lanes of a pixel group must iterate together (shuffles) -> loop bound on the group's first pixel
    for (long long base = (long long)blockIdx.x * lanes; base < npix; base += (long long)gridDim.x * lanes) {
        long long pix = base + lane;
        bool ok = pix < npix;
        if (!ok) pix = npix - 1;
        int n = (int)(pix / HW);
        float4 x = ld4(raw + pix * C + c);
        float4 z = f4_relu(f4_fma(x, sc, sh));
        float4 g = ld4(gout + pix * C + c);
        float dot = group_sum(z.x * w.x + z.y * w.y + z.z * w.z + z.w * w.w, cg);
        float s = 1.f / (1.f + expf(-(dot + bs)));
        float D = group_sum(g.x * z.x + g.y * z.y + g.z * z.z + g.w * z.w, cg);
        float dsp = D * s * (1.f - s);
        float4 cse = ld4(se.cse + (size_t)n * C + c), G = ld4(se.G + (size_t)n * C + c);
        float4 dz = make_float4(g.x * (cse.x + s) + dsp * w.x + G.x, g.y * (cse.y + s) + dsp * w.y + G.y,
                                g.z * (cse.z + s) + dsp * w.z + G.z, g.w * (cse.w + s) + dsp * w.w + G.w);
        dz = f4_mask_pos(dz, z);
        if (ok) {
            st4(gbn + pix * C + c, dz);
            float4 xh = f4_mul(make_float4(x.x - mu.x, x.y - mu.y, x.z - mu.z, x.w - mu.w), is);
            sg = f4_add(sg, dz);
            sgx = f4_fma(dz, xh, sgx);
            sws = f4_fma(z, make_float4(dsp, dsp, dsp, dsp), sws);
            if (cv == 0) sbs += dsp;
        }
    }
    block_reduce_to_double(sg, cg, bn.bsums, red);
    block_reduce_to_double(sgx, cg, bn.bsums + C, red);
    block_reduce_to_float(sws, cg, se.dws, red);
    // dbs: sum sbs over the block
    red[threadIdx.x] = make_float4(sbs, 0.f, 0.f, 0.f);
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < EW_THREADS; ++i) t += red[i].x;
        atomicAdd(se.dbs, t);
    }
}
void k_scse_bwd(cudaStream_t st, const Tensor& gout, const Tensor& raw, const BNRef& bn, const SERef& se, const Tensor& gbn) {
    SALT_COUNT(3);
    const int HW = raw.H * raw.W, C = raw.C, cg = C / 4;
    const int chunks = se.chunks;
    long long npix = (long long)raw.B * HW;
    SALT_DISPATCH(raw.dt, T, {
        scse_pool_kernel<T, true><<<dim3(chunks, raw.B), EW_THREADS, 0, st>>>((const T*)raw.p, (const T*)gout.p, bn.scale, bn.shift, se.part, HW, C);
        scse_fc_bwd_kernel<<<raw.B, max(32, C), sizeof(float) * (C + se.Cr), st>>>(se, 1.0f / HW);
        scse_bwd_apply_kernel<T><<<reduce_blocks(npix, cg), EW_THREADS, 0, st>>>((const T*)gout.p, (const T*)raw.p, bn, se, (T*)gbn.p, npix, HW, C);
    });
}

// ------------------------------------------------------------------------------------------------
// final 1x1 conv C -> K (K <= 4) on z = relu(bn(raw)); logits are fp32 NCHW (the reference layout)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void final_fwd_kernel(const T* __restrict__ raw, const float* __restrict__ scale, const float* __restrict__ shift,
                                 const float* __restrict__ w, const float* __restrict__ b, int K, float* __restrict__ logits,
                                 long long npix, int HW, int C) {
    const int cg = C >> 2;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int cv = (int)(idx % cg);
    long long pix = idx / cg;
    bool ok = pix < npix;
    if (!ok) pix = npix - 1;
    int c = cv * 4;
    float4 z = f4_relu(f4_fma(ld4(raw + pix * C + c), ld4(scale + c), ld4(shift + c)));
    int n = (int)(pix / HW), p = (int)(pix % HW);
    for (int k = 0; k < K; ++k) {
        float4 wk = ld4(w + k * C + c);
        float d = group_sum(z.x * wk.x + z.y * wk.y + z.z * wk.z + z.w * wk.w, cg);
        if (ok && cv == 0) logits[((size_t)n * K + k) * HW + p] = d + b[k];
    }
}
void k_final_fwd(cudaStream_t st, const Tensor& raw, const float* scale, const float* shift, const float* w, const float* b,
                 int K, float* logits) {
    SALT_COUNT(1);
    long long npix = (long long)raw.B * raw.H * raw.W;
    int cg = raw.C / 4;
    SALT_DISPATCH(raw.dt, T, (final_fwd_kernel<T><<<cdiv(npix * cg, EW_THREADS), EW_THREADS, 0, st>>>(
        (const T*)raw.p, scale, shift, w, b, K, logits, npix, raw.H * raw.W, raw.C)));
}
template <typename T, int K>
__global__ void final_bwd_kernel(const float* __restrict__ dlogits, const T* __restrict__ raw, BNRef bn,
                                 const float* __restrict__ w, float* __restrict__ dw, float* __restrict__ db,
                                 T* __restrict__ gbn, long long npix, int HW, int C) {
    __shared__ float4 red[EW_THREADS];
    const int cg = C >> 2, lanes = EW_THREADS / cg;
    const int cv = threadIdx.x % cg, lane = threadIdx.x / cg, c = cv * 4;
    const float4 sc = ld4(bn.scale + c), sh = ld4(bn.shift + c), mu = ld4(bn.mean + c), is = ld4(bn.invstd + c);
    float4 wk[K], sdw[K];
    float sdb[K];
    for (int k = 0; k < K; ++k) { wk[k] = ld4(w + k * C + c); sdw[k] = f4_zero(); sdb[k] = 0.f; }
    float4 sg = f4_zero(), sgx = f4_zero();
    for (long long pix = (long long)blockIdx.x * lanes + lane; pix < npix; pix += (long long)gridDim.x * lanes) {
        int n = (int)(pix / HW), p = (int)(pix % HW);
        float4 x = ld4(raw + pix * C + c);
        float4 z = f4_relu(f4_fma(x, sc, sh));
        float4 gz = f4_zero();
        for (int k = 0; k < K; ++k) {
            float dl = dlogits[((size_t)n * K + k) * HW + p];
            gz = f4_fma(wk[k], make_float4(dl, dl, dl, dl), gz);
            sdw[k] = f4_fma(z, make_float4(dl, dl, dl, dl), sdw[k]);
            if (cv == 0) sdb[k] += dl;
        }
        gz = f4_mask_pos(gz, z);
        st4(gbn + pix * C + c, gz);
        float4 xh = f4_mul(make_float4(x.x - mu.x, x.y - mu.y, x.z - mu.z, x.w - mu.w), is);
        sg = f4_add(sg, gz);
        sgx = f4_fma(gz, xh, sgx);
    }
    block_reduce_to_double(sg, cg, bn.bsums, red);
    block_reduce_to_double(sgx, cg, bn.bsums + C, red);
    for (int k = 0; k < K; ++k) block_reduce_to_float(sdw[k], cg, dw + k * C, red);
    float4 t = f4_zero();
    if (K > 0) t.x = sdb[0];
    if (K > 1) t.y = sdb[1];
    if (K > 2) t.z = sdb[2];
    if (K > 3) t.w = sdb[3];
    red[threadIdx.x] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        float4 s = f4_zero();
        for (int i = 0; i < EW_THREADS; ++i) s = f4_add(s, red[i]);
        float sv[4] = {s.x, s.y, s.z, s.w};
        for (int k = 0; k < K; ++k) atomicAdd(db + k, sv[k]);
    }
}
void k_final_bwd(cudaStream_t st, const float* dlogits, const Tensor& raw, const BNRef& bn, const float* w, int K,
                 float* dw, float* db, const Tensor& gbn) {
    SALT_COUNT(1);
    long long npix = (long long)raw.B * raw.H * raw.W;
    int cg = raw.C / 4, blocks = reduce_blocks(npix, cg), HW = raw.H * raw.W;
#define LAUNCH_FB(KK) final_bwd_kernel<T, KK><<<blocks, EW_THREADS, 0, st>>>(dlogits, (const T*)raw.p, bn, w, dw, db, (T*)gbn.p, npix, HW, raw.C)
    SALT_DISPATCH(raw.dt, T, {
        if (K == 1) LAUNCH_FB(1); else if (K == 2) LAUNCH_FB(2); else if (K == 3) LAUNCH_FB(3); else LAUNCH_FB(4);
    });
#undef LAUNCH_FB
}

// ------------------------------------------------------------------------------------------------
// ReLU / BatchNorm backward
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void relu_mask_inplace_kernel(T* __restrict__ g, const T* __restrict__ mask, long long nvec) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nvec) return;
    st4(g + idx * 4, f4_mask_pos(ld4(g + idx * 4), ld4(mask + idx * 4)));
}
void k_relu_mask_inplace(cudaStream_t st, const Tensor& g, const Tensor& mask) {
    SALT_COUNT(1);
    long long nvec = (long long)g.numel() / 4;
    SALT_DISPATCH(g.dt, T, (relu_mask_inplace_kernel<T><<<cdiv(nvec, EW_THREADS), EW_THREADS, 0, st>>>((T*)g.p, (const T*)mask.p, nvec)));
}
template <typename T>
__global__ void bn_bwd_reduce_kernel(const T* __restrict__ g, const T* __restrict__ raw, BNRef bn, int self_mask,
                                     long long npix, int C) {
    __shared__ float4 red[EW_THREADS];
    const int cg = C >> 2, lanes = EW_THREADS / cg;
    const int cv = threadIdx.x % cg, lane = threadIdx.x / cg, c = cv * 4;
    const float4 sc = ld4(bn.scale + c), sh = ld4(bn.shift + c), mu = ld4(bn.mean + c), is = ld4(bn.invstd + c);
    float4 sg = f4_zero(), sgx = f4_zero();
    for (long long pix = (long long)blockIdx.x * lanes + lane; pix < npix; pix += (long long)gridDim.x * lanes) {
        float4 x = ld4(raw + pix * C + c), gv = ld4(g + pix * C + c);
        if (self_mask) gv = f4_mask_pos(gv, f4_fma(x, sc, sh));
        float4 xh = f4_mul(make_float4(x.x - mu.x, x.y - mu.y, x.z - mu.z, x.w - mu.w), is);
        sg = f4_add(sg, gv);
        sgx = f4_fma(gv, xh, sgx);
    }
    block_reduce_to_double(sg, cg, bn.bsums, red);
    block_reduce_to_double(sgx, cg, bn.bsums + C, red);
}
void k_bn_bwd_reduce(cudaStream_t st, const Tensor& g, const Tensor& raw, const BNRef& bn, bool self_mask) {
    SALT_COUNT(1);
    long long npix = (long long)raw.B * raw.H * raw.W;
    int cg = raw.C / 4;
    SALT_DISPATCH(raw.dt, T, (bn_bwd_reduce_kernel<T><<<reduce_blocks(npix, cg), EW_THREADS, 0, st>>>(
        (const T*)g.p, (const T*)raw.p, bn, self_mask ? 1 : 0, npix, raw.C)));
}
template <typename T>
__global__ void bn_bwd_apply_kernel(const T* __restrict__ g, const T* __restrict__ raw, BNRef bn, int self_mask,
                                    T* __restrict__ graw, long long nvec, int C) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nvec) return;
    int c = (int)(idx % (C >> 2)) * 4;
    float4 x = ld4(raw + idx * 4), gv = ld4(g + idx * 4);
    float4 sc = ld4(bn.scale + c);
    if (self_mask) gv = f4_mask_pos(gv, f4_fma(x, sc, ld4(bn.shift + c)));
    float4 mu = ld4(bn.mean + c), cb = ld4(bn.cb + c), cc = ld4(bn.cc + c);
    float4 r = make_float4(sc.x * (gv.x - cb.x - cc.x * (x.x - mu.x)), sc.y * (gv.y - cb.y - cc.y * (x.y - mu.y)),
                           sc.z * (gv.z - cb.z - cc.z * (x.z - mu.z)), sc.w * (gv.w - cb.w - cc.w * (x.w - mu.w)));
    st4(graw + idx * 4, r);
}
void k_bn_bwd_apply(cudaStream_t st, const Tensor& g, const Tensor& raw, const BNRef& bn, bool self_mask, const Tensor& graw) {
    SALT_COUNT(1);
    long long nvec = (long long)raw.numel() / 4;
    SALT_DISPATCH(raw.dt, T, (bn_bwd_apply_kernel<T><<<cdiv(nvec, EW_THREADS), EW_THREADS, 0, st>>>(
        (const T*)g.p, (const T*)raw.p, bn, self_mask ? 1 : 0, (T*)graw.p, nvec, raw.C)));
}

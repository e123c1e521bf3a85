// The BatchNorm passes of the engine on the stream ring (stream_ring.cuh): same arithmetic, same fixed-order partial-slot
// reductions as kernels_elem.cu - only the way the bytes reach the SM differs.  Each launcher returns false when the tensors do
// not fit the ring's geometry (row bytes not a power of two <= 4096, bordered input); the caller then takes the direct-load kernel.
#include "kernels_stream.h"
#include "stream_ring.cuh"
#include <algorithm>

namespace {

__device__ __forceinline__ int ilog2(unsigned v) { return 31 - __clz(v); }

// ------------------------------------------------------------------------------------------------ BN apply (+ residual, ReLU)
template <typename T, bool RES>
struct BnApplyOp {
    static constexpr int N = VW<T>::N;
    static constexpr bool FULL_WARPS = false;
    static constexpr int WARPS = 16, NT = WARPS * 32;
    const float *scale, *shift, *rscale, *rshift, *gate;
    T* out;
    int relu, C, H, W, pt, pb, pl, pr;
    FDiv dHW, dW, dH;
    // per thread
    Vf<N> sc, sh, rsc, rsh;
    int c, rshift_bits;
    __device__ void begin(int tid) {
        const int rb = C * (int)sizeof(T);
        rshift_bits = ilog2((unsigned)rb);
        c = ((tid * 16) & (rb - 1)) / (int)sizeof(T);
        sc = ldp<N>(scale + c); sh = ldp<N>(shift + c);
        rsc = vzero<N>(); rsh = vzero<N>();
        if (RES && rscale) { rsc = ldp<N>(rscale + c); rsh = ldp<N>(rshift + c); }
    }
    __device__ void vec(size_t off, const uint4 (&in)[RES ? 2 : 1]) {
        Vf<N> v = vfma(vfrom<T>(in[0]), sc, sh);
        const unsigned pix = (unsigned)(off >> rshift_bits);
        if (gate) v = vmul(v, ldp<N>(gate + (size_t)dHW.div(pix) * C + c));
        if constexpr (RES) {
            Vf<N> r = vfrom<T>(in[1]);
            if (rscale) r = vfma(r, rsc, rsh);
            v = vadd(v, r);
        }
        if (relu) v = vrelu(v);
        const uint4 o = vto<T>(v);
        if ((pt | pb | pl | pr) == 0) {
            *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(out) + off) = o;
        } else {
            const unsigned row = dW.div(pix), x = pix - row * W, n = dH.div(row), y = row - n * H;
            const int Hp = H + pt + pb, Wp = W + pl + pr;
            const int y0 = (y == 0) ? 0 : (int)y + pt, y1 = ((int)y == H - 1) ? Hp - 1 : (int)y + pt;
            const int x0 = (x == 0) ? 0 : (int)x + pl, x1 = ((int)x == W - 1) ? Wp - 1 : (int)x + pl;
            for (int yy = y0; yy <= y1; ++yy)
                for (int xx = x0; xx <= x1; ++xx) *reinterpret_cast<uint4*>(out + (((size_t)n * Hp + yy) * Wp + xx) * C + c) = o;
        }
    }
    __device__ void end(int, float*) {}
};

// ------------------------------------------------------------------------------------------------ BN backward: reduce, apply
// upstream gradient seen by the layer: g, optionally gated per (image, channel) (encoder SE), optionally masked by its own ReLU
template <typename T>
struct BnBwdReduceOp {
    static constexpr int N = VW<T>::N;
    static constexpr bool FULL_WARPS = false;
    static constexpr int WARPS = 16, NT = WARPS * 32;
    BNRef bn;
    const float *gate, *addc;
    int self_mask, HW, C;
    FDiv dHW;
    Vf<N> sc, sh, mu, is, sg, sgx;
    int c, rshift_bits;
    __device__ void begin(int tid) {
        const int rb = C * (int)sizeof(T);
        rshift_bits = ilog2((unsigned)rb);
        c = ((tid * 16) & (rb - 1)) / (int)sizeof(T);
        sc = ldp<N>(bn.scale + c); sh = ldp<N>(bn.shift + c); mu = ldp<N>(bn.mean + c); is = ldp<N>(bn.invstd + c);
        sg = vzero<N>(); sgx = vzero<N>();
    }
    __device__ void vec(size_t off, const uint4 (&in)[2]) {
        Vf<N> gv = vfrom<T>(in[0]);
        const Vf<N> x = vfrom<T>(in[1]);
        if (gate) {
            const size_t o = (size_t)dHW.div((unsigned)(off >> rshift_bits)) * C + c;
            gv = vfma(gv, ldp<N>(gate + o), ldp<N>(addc + o));
        }
        if (self_mask) gv = vmaskpos(gv, vfma(x, sc, sh));
        sg = vadd(sg, gv);
        sgx = vfma(gv, vxhat(x, mu, is), sgx);
    }
    __device__ void end(int tid, float* red) {
        const int cg = C / N;
        block_reduce_slot<N, true, NT>(sg, cg, bn.bsums + (size_t)blockIdx.x * 2 * C, red);
        block_reduce_slot<N, true, NT>(sgx, cg, bn.bsums + (size_t)blockIdx.x * 2 * C + C, red);
        if (blockIdx.x == 0 && tid == 0) *bn.bslots = (int)gridDim.x;
    }
};

template <typename T>
struct BnBwdApplyOp {
    static constexpr int N = VW<T>::N;
    static constexpr bool FULL_WARPS = false;
    static constexpr int WARPS = 16, NT = WARPS * 32;
    BNRef bn;
    const float *gate, *addc;
    T* graw;
    int self_mask, HW, C;
    FDiv dHW;
    Vf<N> sc, sh, mu, cb, cc;
    int c, rshift_bits;
    __device__ void begin(int tid) {
        const int rb = C * (int)sizeof(T);
        rshift_bits = ilog2((unsigned)rb);
        c = ((tid * 16) & (rb - 1)) / (int)sizeof(T);
        sc = ldp<N>(bn.scale + c); sh = ldp<N>(bn.shift + c); mu = ldp<N>(bn.mean + c); cb = ldp<N>(bn.cb + c); cc = ldp<N>(bn.cc + c);
    }
    __device__ void vec(size_t off, const uint4 (&in)[2]) {
        Vf<N> gv = vfrom<T>(in[0]);
        const Vf<N> x = vfrom<T>(in[1]);
        if (gate) {
            const size_t o = (size_t)dHW.div((unsigned)(off >> rshift_bits)) * C + c;
            gv = vfma(gv, ldp<N>(gate + o), ldp<N>(addc + o));
        }
        if (self_mask) gv = vmaskpos(gv, vfma(x, sc, sh));
        Vf<N> r;
#pragma unroll
        for (int i = 0; i < N; ++i) r.v[i] = sc.v[i] * (gv.v[i] - cb.v[i] - cc.v[i] * (x.v[i] - mu.v[i]));
        *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(graw) + off) = vto<T>(r);
    }
    __device__ void end(int, float*) {}
};

// g <- g where mask > 0 (the ReLU that closes a residual block), in place
template <typename T>
struct ReluMaskOp {
    static constexpr int N = VW<T>::N;
    static constexpr bool FULL_WARPS = false;
    static constexpr int WARPS = 16, NT = WARPS * 32;
    T* g;
    __device__ void begin(int) {}
    __device__ void vec(size_t off, const uint4 (&in)[2]) {
        *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(g) + off) = vto<T>(vmaskpos(vfrom<T>(in[0]), vfrom<T>(in[1])));
    }
    __device__ void end(int, float*) {}
};

// ------------------------------------------------------------------------------------------------ per-pixel passes (bf16)
// scSE gate, final 1x1 convolution: the C/8 consecutive threads that hold one pixel reduce over its channels with warp shuffles
// (C <= 256), so every thread runs every pass of a chunk (FULL_WARPS).
__device__ __forceinline__ uint4 lds16(const uint8_t* p) { return *reinterpret_cast<const uint4*>(p); }

// out = z * (cse[n][c] + sigmoid(<z, ws> + bs)),  z = relu(raw*scale+shift)      (base.py:82-117)
struct ScseApplyOp {
    static constexpr int N = 8;
    static constexpr bool FULL_WARPS = true;
    static constexpr int WARPS = 16, NT = WARPS * 32;
    const float *scale, *shift;
    SERef se;
    bf16* out;
    int HW, C;
    FDiv dHW;
    Vf<N> sc, sh, w;
    float bs;
    int c, cg, rshift_bits;
    __device__ void begin(int tid) {
        const int rb = C * 2;
        rshift_bits = ilog2((unsigned)rb); cg = C / N;
        c = ((tid * 16) & (rb - 1)) / 2;
        sc = ldp<N>(scale + c); sh = ldp<N>(shift + c); w = ldp<N>(se.ws + c); bs = se.bs[0];
    }
    static constexpr int U = 2;
    __device__ void vecs(size_t off, const uint8_t* stage, const int (&o)[U], const bool (&valid)[U]) {
        Vf<N> z[U], gate[U];
        float dot[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            z[u] = vrelu(vfma(vfrom<bf16>(lds16(stage + o[u])), sc, sh));
            gate[u] = ldp<N>(se.cse + (size_t)dHW.div((unsigned)((off + o[u]) >> rshift_bits)) * C + c);
            dot[u] = vdot(z[u], w);
        }
        for (int m = cg >> 1; m > 0; m >>= 1) {
#pragma unroll
            for (int u = 0; u < U; ++u) dot[u] += __shfl_xor_sync(0xffffffffu, dot[u], m);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const float sg = __fdividef(1.f, 1.f + __expf(-(dot[u] + bs)));
#pragma unroll
            for (int i = 0; i < N; ++i) gate[u].v[i] += sg;
            if (valid[u]) *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(out) + off + o[u]) = vto<bf16>(vrelu(vmul(z[u], gate[u])));
        }
    }
    __device__ void end(int, float*) {}
};

// scSE backward through out = z*(cse + s): gradient w.r.t. the BatchNorm output (ReLU-masked), BatchNorm backward sums in the
// block's slot, spatial-gate parameter gradients
struct ScseBwdOp {
    static constexpr int N = 8;
    static constexpr bool FULL_WARPS = true;
    static constexpr int WARPS = 8, NT = WARPS * 32;
    BNRef bn;
    SERef se;
    bf16* gbn;
    int HW, C;
    FDiv dHW;
    Vf<N> sc, sh, mu, is, w, sg, sgx, sws;
    float bs, sbs;
    int c, cg, rshift_bits;
    __device__ void begin(int tid) {
        const int rb = C * 2;
        rshift_bits = ilog2((unsigned)rb); cg = C / N;
        c = ((tid * 16) & (rb - 1)) / 2;
        sc = ldp<N>(bn.scale + c); sh = ldp<N>(bn.shift + c); mu = ldp<N>(bn.mean + c); is = ldp<N>(bn.invstd + c);
        w = ldp<N>(se.ws + c); bs = se.bs[0];
        sg = vzero<N>(); sgx = vzero<N>(); sws = vzero<N>(); sbs = 0.f;
    }
    static constexpr int U = 2;
    __device__ void vecs(size_t off, const uint8_t* stage, const int (&o)[U], const bool (&valid)[U]) {
        Vf<N> g[U], x[U], z[U];
        float dot[U], D[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            g[u] = vfrom<bf16>(lds16(stage + o[u])); x[u] = vfrom<bf16>(lds16(stage + ring::CHUNK + o[u]));
            z[u] = vrelu(vfma(x[u], sc, sh));
            dot[u] = vdot(z[u], w); D[u] = vdot(g[u], z[u]);
        }
        for (int m = cg >> 1; m > 0; m >>= 1) {
#pragma unroll
            for (int u = 0; u < U; ++u) { dot[u] += __shfl_xor_sync(0xffffffffu, dot[u], m); D[u] += __shfl_xor_sync(0xffffffffu, D[u], m); }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const float s = __fdividef(1.f, 1.f + __expf(-(dot[u] + bs)));
            const float dsp = D[u] * s * (1.f - s);
            const size_t no = (size_t)dHW.div((unsigned)((off + o[u]) >> rshift_bits)) * C + c;
            const Vf<N> cse = ldp<N>(se.cse + no), G = ldp<N>(se.G + no);
            Vf<N> dz;
#pragma unroll
            for (int i = 0; i < N; ++i) dz.v[i] = z[u].v[i] > 0.f ? g[u].v[i] * (cse.v[i] + s) + dsp * w.v[i] + G.v[i] : 0.f;
            if (valid[u]) {
                *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(gbn) + off + o[u]) = vto<bf16>(dz);
                sg = vadd(sg, dz);
                sgx = vfma(dz, vxhat(x[u], mu, is), sgx);
                sws = vaxpy(z[u], dsp, sws);
                if (c == 0) sbs += dsp;
            }
        }
    }
    __device__ void end(int tid, float* red) {
        block_reduce_slot<N, true, NT>(sg, cg, bn.bsums + (size_t)blockIdx.x * 2 * C, red);
        block_reduce_slot<N, true, NT>(sgx, cg, bn.bsums + (size_t)blockIdx.x * 2 * C + C, red);
        if (blockIdx.x == 0 && tid == 0) *bn.bslots = (int)gridDim.x;
        block_reduce_add<N, float, true, NT>(sws, cg, se.dws, red);
        red[tid] = sbs;
        ring::consumer_sync<NT>();
        if (tid == 0) {
            float t = 0.f;
            for (int i = 0; i < NT; ++i) t += red[i];
            atomicAdd(se.dbs, t);
        }
    }
};

// logits[n][k][p] = <relu(raw*scale+shift), w[k]> + b[k]      (fp32 NCHW, the reference layout)
struct FinalFwdOp {
    static constexpr int N = 8;
    static constexpr bool FULL_WARPS = true;
    static constexpr int WARPS = 16, NT = WARPS * 32;
    const float *scale, *shift, *w, *b;
    float* logits;
    int K, HW, C;
    FDiv dHW;
    Vf<N> sc, sh;
    int c, cg, rshift_bits;
    __device__ void begin(int tid) {
        const int rb = C * 2;
        rshift_bits = ilog2((unsigned)rb); cg = C / N;
        c = ((tid * 16) & (rb - 1)) / 2;
        sc = ldp<N>(scale + c); sh = ldp<N>(shift + c);
    }
    static constexpr int U = 2;
    __device__ void vecs(size_t off, const uint8_t* stage, const int (&o)[U], const bool (&valid)[U]) {
        Vf<N> z[U];
#pragma unroll
        for (int u = 0; u < U; ++u) z[u] = vrelu(vfma(vfrom<bf16>(lds16(stage + o[u])), sc, sh));
        for (int k = 0; k < K; ++k) {
            const Vf<N> wk = ldp<N>(w + k * C + c);
            float d[U];
#pragma unroll
            for (int u = 0; u < U; ++u) d[u] = vdot(z[u], wk);
            for (int m = cg >> 1; m > 0; m >>= 1) {
#pragma unroll
                for (int u = 0; u < U; ++u) d[u] += __shfl_xor_sync(0xffffffffu, d[u], m);
            }
            if (c == 0) {
                const float bk = b[k];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const unsigned pix = (unsigned)((off + o[u]) >> rshift_bits), n = dHW.div(pix), p = pix - n * HW;
                    if (valid[u]) logits[((size_t)n * K + k) * HW + p] = d[u] + bk;
                }
            }
        }
    }
    __device__ void end(int, float*) {}
};

// backward of the final 1x1 convolution + the ReLU / BatchNorm sums of final.0: streams 1..K are the dlogits planes of the image
template <int K>
struct FinalBwdOp {
    static constexpr int N = 8;
    static constexpr bool FULL_WARPS = true;
    static constexpr int WARPS = 8, NT = WARPS * 32;
    BNRef bn;
    const float* w;
    float *dw, *db;
    bf16* gbn;
    int C;
    Vf<N> sc, sh, mu, is, sg, sgx, wk[K], sdw[K];
    float sdb[K];
    int c, cg, rshift_bits;
    __device__ void begin(int tid) {
        const int rb = C * 2;
        cg = C / N; rshift_bits = ilog2((unsigned)rb);
        c = ((tid * 16) & (rb - 1)) / 2;
        sc = ldp<N>(bn.scale + c); sh = ldp<N>(bn.shift + c); mu = ldp<N>(bn.mean + c); is = ldp<N>(bn.invstd + c);
        sg = vzero<N>(); sgx = vzero<N>();
#pragma unroll
        for (int k = 0; k < K; ++k) { wk[k] = ldp<N>(w + k * C + c); sdw[k] = vzero<N>(); sdb[k] = 0.f; }
    }
    __device__ void end(int tid, float* red) {
        block_reduce_slot<N, true, NT>(sg, cg, bn.bsums + (size_t)blockIdx.x * 2 * C, red);
        block_reduce_slot<N, true, NT>(sgx, cg, bn.bsums + (size_t)blockIdx.x * 2 * C + C, red);
        if (blockIdx.x == 0 && tid == 0) *bn.bslots = (int)gridDim.x;
#pragma unroll
        for (int k = 0; k < K; ++k) block_reduce_add<N, float, true, NT>(sdw[k], cg, dw + k * C, red);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            red[tid] = sdb[k];
            ring::consumer_sync<NT>();
            if (tid == 0) {
                float t = 0.f;
                for (int i = 0; i < NT; ++i) t += red[i];
                atomicAdd(db + k, t);
            }
            ring::consumer_sync<NT>();
        }
    }
    // the dlogits planes are 1/(C/2) as dense as the activation stream, so the op addresses the stage itself (no shuffles here:
    // the tail of a short chunk simply returns)
    static constexpr int U = 2;
    __device__ void vecs(size_t off, const uint8_t* stage, const int (&o)[U], const bool (&valid)[U]) {
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (valid[u]) one(off + o[u], stage, o[u]);
    }
    __device__ void one(size_t off, const uint8_t* stage, int o) {
        const Vf<N> x = vfrom<bf16>(lds16(stage + o));
        const Vf<N> z = vrelu(vfma(x, sc, sh));
        const int pl = o >> rshift_bits;                  // pixel within the chunk
        Vf<N> gz = vzero<N>();
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float dl = *reinterpret_cast<const float*>(stage + (size_t)(k + 1) * ring::CHUNK + pl * 4);
            gz = vaxpy(wk[k], dl, gz);
            sdw[k] = vaxpy(z, dl, sdw[k]);
            if (c == 0) sdb[k] += dl;
        }
        gz = vmaskpos(gz, z);
        *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(gbn) + off) = vto<bf16>(gz);
        sg = vadd(sg, gz);
        sgx = vfma(gz, vxhat(x, mu, is), sgx);
    }
};

// ------------------------------------------------------------------------------------------------ squeeze (per-image channel sums)
// part[n][slice][c] = sum over the slice's pixels of  [g *] [relu](raw*scale+shift): the pooling pass of both squeeze-and-excitation
// kinds (decoder scSE: RELU; encoder SE module: plain; WITH_G: the backward pass).  One CTA = one slice of one image
// (Streams::group_chunks), so the partial belongs to one image; the FC kernels add the `chunks` partial slots in slot order.
template <typename T, bool WITH_G, bool RELU>
struct SePoolOp {
    static constexpr int N = VW<T>::N;
    static constexpr bool FULL_WARPS = false;
    static constexpr int WARPS = 8, NT = WARPS * 32;
    const float *scale, *shift;
    float* part;
    int C, slices, chunks;
    Vf<N> sc, sh, acc;
    int c;
    __device__ void begin(int tid) {
        const int rb = C * (int)sizeof(T);
        c = ((tid * 16) & (rb - 1)) / (int)sizeof(T);
        sc = ldp<N>(scale + c); sh = ldp<N>(shift + c);
        acc = vzero<N>();
    }
    __device__ void vec(size_t, const uint4 (&in)[WITH_G ? 2 : 1]) {
        Vf<N> z = vfma(vfrom<T>(in[0]), sc, sh);
        if (RELU) z = vrelu(z);
        if constexpr (WITH_G) z = vmul(z, vfrom<T>(in[1]));
        acc = vadd(acc, z);
    }
    __device__ void end(int tid, float* red) {
        const int n = blockIdx.x / slices, slice = blockIdx.x - n * slices;
        block_reduce_slot<N, true, NT>(acc, C / N, part + ((size_t)n * chunks + slice) * C, red);
        if (slice == 0)                                   // the slots this launch does not use must read as zero
            for (int i = tid; i < (chunks - slices) * C; i += NT) part[((size_t)n * chunks + slices) * C + i] = 0.f;
    }
};

bool same_shape(const Tensor& a, const Tensor& b) { return a.B == b.B && a.H == b.H && a.W == b.W && a.C == b.C && a.dt == b.dt; }
size_t flat_bytes(const Tensor& t) { return (size_t)t.B * t.H * t.W * t.C * dtype_size(t.dt); }

}  // namespace

bool ring_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SALT_EW_RING"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}

bool k_ring_bn_apply(cudaStream_t st, const Tensor& raw, const float* scale, const float* shift, const Tensor* res, const float* rscale,
                     const float* rshift, bool relu, const Tensor& out, const float* gate) {
    if (!ring_enabled() || !ring::row_ok(raw) || (res && (!ring::row_ok(*res) || !same_shape(raw, *res)))) return false;
    if (out.B != raw.B || out.H != raw.H || out.W != raw.W || out.C != raw.C || out.dt != raw.dt) return false;
    SALT_DISPATCH(raw.dt, T, {
        if (res) {
            BnApplyOp<T, true> op;
            op.scale = scale; op.shift = shift; op.rscale = rscale; op.rshift = rshift; op.gate = gate; op.out = (T*)out.p;
            op.relu = relu ? 1 : 0; op.C = raw.C; op.H = raw.H; op.W = raw.W; op.pt = out.pt; op.pb = out.pb; op.pl = out.pl; op.pr = out.pr;
            op.dHW = make_fdiv(raw.H * raw.W); op.dW = make_fdiv(raw.W); op.dH = make_fdiv(raw.H);
            ring::Streams<2> s; s.p[0] = (const uint8_t*)raw.p; s.p[1] = (const uint8_t*)res->p; s.nbytes = flat_bytes(raw);
            ring::launch<2>(st, s, op);
        } else {
            BnApplyOp<T, false> op;
            op.scale = scale; op.shift = shift; op.rscale = nullptr; op.rshift = nullptr; op.gate = gate; op.out = (T*)out.p;
            op.relu = relu ? 1 : 0; op.C = raw.C; op.H = raw.H; op.W = raw.W; op.pt = out.pt; op.pb = out.pb; op.pl = out.pl; op.pr = out.pr;
            op.dHW = make_fdiv(raw.H * raw.W); op.dW = make_fdiv(raw.W); op.dH = make_fdiv(raw.H);
            ring::Streams<1> s; s.p[0] = (const uint8_t*)raw.p; s.nbytes = flat_bytes(raw);
            ring::launch<1>(st, s, op);
        }
    });
    return true;
}

bool k_ring_bn_bwd_reduce(cudaStream_t st, const Tensor& g, const Tensor& raw, const BNRef& bn, bool self_mask, const float* gate,
                          const float* addc) {
    if (!ring_enabled() || !ring::row_ok(raw) || !ring::row_ok(g) || !same_shape(raw, g)) return false;
    SALT_DISPATCH(raw.dt, T, {
        BnBwdReduceOp<T> op;
        op.bn = bn; op.gate = gate; op.addc = addc; op.self_mask = self_mask ? 1 : 0; op.HW = raw.H * raw.W; op.C = raw.C;
        op.dHW = make_fdiv(raw.H * raw.W);
        ring::Streams<2> s; s.p[0] = (const uint8_t*)g.p; s.p[1] = (const uint8_t*)raw.p; s.nbytes = flat_bytes(raw);
        ring::launch<2>(st, s, op);
    });
    return true;
}

bool k_ring_bn_bwd_apply(cudaStream_t st, const Tensor& g, const Tensor& raw, const BNRef& bn, bool self_mask, const Tensor& graw,
                         const float* gate, const float* addc) {
    if (!ring_enabled() || !ring::row_ok(raw) || !ring::row_ok(g) || !ring::row_ok(graw) || !same_shape(raw, g) || !same_shape(raw, graw))
        return false;
    SALT_DISPATCH(raw.dt, T, {
        BnBwdApplyOp<T> op;
        op.bn = bn; op.gate = gate; op.addc = addc; op.graw = (T*)graw.p; op.self_mask = self_mask ? 1 : 0; op.HW = raw.H * raw.W;
        op.C = raw.C; op.dHW = make_fdiv(raw.H * raw.W);
        ring::Streams<2> s; s.p[0] = (const uint8_t*)g.p; s.p[1] = (const uint8_t*)raw.p; s.nbytes = flat_bytes(raw);
        ring::launch<2>(st, s, op);
    });
    return true;
}

bool k_ring_relu_mask(cudaStream_t st, const Tensor& g, const Tensor& mask) {
    if (!ring_enabled() || !ring::row_ok(g) || !ring::row_ok(mask) || !same_shape(g, mask)) return false;
    SALT_DISPATCH(g.dt, T, {
        ReluMaskOp<T> op; op.g = (T*)g.p;
        ring::Streams<2> s; s.p[0] = (const uint8_t*)g.p; s.p[1] = (const uint8_t*)mask.p; s.nbytes = flat_bytes(g);
        ring::launch<2>(st, s, op);
    });
    return true;
}

static bool pixel_group_ok(const Tensor& t) {       // bf16, one pixel's channel groups inside a warp
    const int cg = t.C / 8;
    return t.dt == DT_BF16 && t.C % 8 == 0 && cg <= 32 && (cg & (cg - 1)) == 0 && ring::row_ok(t);
}
bool k_ring_scse_apply(cudaStream_t st, const Tensor& raw, const float* scale, const float* shift, const SERef& se, const Tensor& out) {
    if (!ring_enabled() || !pixel_group_ok(raw) || !ring::row_ok(out) || !same_shape(raw, out)) return false;
    ScseApplyOp op;
    op.scale = scale; op.shift = shift; op.se = se; op.out = (bf16*)out.p; op.HW = raw.H * raw.W; op.C = raw.C;
    op.dHW = make_fdiv(raw.H * raw.W);
    ring::Streams<1> s; s.p[0] = (const uint8_t*)raw.p; s.nbytes = flat_bytes(raw);
    ring::launch<1>(st, s, op);
    return true;
}
bool k_ring_scse_bwd_apply(cudaStream_t st, const Tensor& gout, const Tensor& raw, const BNRef& bn, const SERef& se, const Tensor& gbn) {
    if (!ring_enabled() || !pixel_group_ok(raw) || !ring::row_ok(gout) || !ring::row_ok(gbn) || !same_shape(raw, gout) ||
        !same_shape(raw, gbn))
        return false;
    ScseBwdOp op;
    op.bn = bn; op.se = se; op.gbn = (bf16*)gbn.p; op.HW = raw.H * raw.W; op.C = raw.C;
    op.dHW = make_fdiv(raw.H * raw.W);
    ring::Streams<2> s; s.p[0] = (const uint8_t*)gout.p; s.p[1] = (const uint8_t*)raw.p; s.nbytes = flat_bytes(raw);
    ring::launch<2>(st, s, op);
    return true;
}
bool k_ring_final_fwd(cudaStream_t st, const Tensor& raw, const float* scale, const float* shift, const float* w, const float* b, int K,
                      float* logits) {
    if (!ring_enabled() || !pixel_group_ok(raw)) return false;
    FinalFwdOp op;
    op.scale = scale; op.shift = shift; op.w = w; op.b = b; op.logits = logits; op.K = K; op.HW = raw.H * raw.W; op.C = raw.C;
    op.dHW = make_fdiv(raw.H * raw.W);
    ring::Streams<1> s; s.p[0] = (const uint8_t*)raw.p; s.nbytes = flat_bytes(raw);
    ring::launch<1>(st, s, op);
    return true;
}
template <int K>
static void launch_final_bwd(cudaStream_t st, const float* dlogits, const Tensor& raw, const BNRef& bn, const float* w, float* dw, float* db,
                             const Tensor& gbn) {
    FinalBwdOp<K> op;
    op.bn = bn; op.w = w; op.dw = dw; op.db = db; op.gbn = (bf16*)gbn.p; op.C = raw.C;
    const size_t HW = (size_t)raw.H * raw.W;
    int shift = 0;
    while ((1 << shift) < raw.C / 2) ++shift;            // activation bytes per pixel (2C) / dlogits bytes per pixel and plane (4)
    ring::Streams<K + 1> s;
    s.p[0] = (const uint8_t*)raw.p; s.nbytes = flat_bytes(raw); s.img_bytes = HW * raw.C * 2;
    for (int k = 0; k < K; ++k) {
        s.p[k + 1] = (const uint8_t*)(dlogits + (size_t)k * HW);
        s.shift[k + 1] = shift; s.img_stride[k + 1] = (size_t)K * HW * sizeof(float);
    }
    ring::launch<K + 1>(st, s, op);
}
bool k_ring_final_bwd(cudaStream_t st, const float* dlogits, const Tensor& raw, const BNRef& bn, const float* w, int K, float* dw, float* db,
                      const Tensor& gbn) {
    if (!ring_enabled() || !pixel_group_ok(raw) || !ring::row_ok(gbn) || !same_shape(raw, gbn) || K < 1 || K > 3) return false;
    const size_t img_bytes = (size_t)raw.H * raw.W * raw.C * 2;
    if (img_bytes % ring::CHUNK || (ring::CHUNK / (raw.C / 2)) % 16 || (reinterpret_cast<uintptr_t>(dlogits) & 15)) return false;
    if (K == 1) launch_final_bwd<1>(st, dlogits, raw, bn, w, dw, db, gbn);
    else if (K == 2) launch_final_bwd<2>(st, dlogits, raw, bn, w, dw, db, gbn);
    else launch_final_bwd<3>(st, dlogits, raw, bn, w, dw, db, gbn);
    return true;
}

template <typename T, bool WITH_G, bool RELU>
static bool ring_se_pool_t(cudaStream_t st, const Tensor& raw, const void* g, const float* scale, const float* shift, const SERef& se) {
    const size_t img_bytes = (size_t)raw.H * raw.W * raw.C * sizeof(T);
    if (img_bytes % ring::CHUNK) return false;           // an image must be a whole number of chunks
    const int cpi = (int)(img_bytes / ring::CHUNK);
    // Slices per image are a property of the LAYER, never of the batch size: the summation order of an image's partial sums must
    // not depend on how many other images ride along (tests/test_engine_gpu.py::test_batch_independence_full_size).  At the
    // benchmark batch (128) two slices of the single-stream ring (two CTAs per SM) / one of the two-stream ring fill one wave.
    const int slices = std::min(WITH_G ? 1 : 2, std::min(cpi, se.chunks));
    SePoolOp<T, WITH_G, RELU> op;
    op.scale = scale; op.shift = shift; op.part = se.part; op.C = raw.C; op.slices = slices; op.chunks = se.chunks;
    ring::Streams<WITH_G ? 2 : 1> s;
    s.p[0] = (const uint8_t*)raw.p; s.nbytes = flat_bytes(raw); s.group_chunks = cpi; s.slices = slices;
    if constexpr (WITH_G) s.p[1] = (const uint8_t*)g;
    ring::launch<WITH_G ? 2 : 1>(st, s, op);
    return true;
}
bool k_ring_se_pool(cudaStream_t st, const Tensor& raw, const void* g, bool relu, const float* scale, const float* shift, const SERef& se) {
    if (!ring_enabled() || !ring::row_ok(raw) || (g && (reinterpret_cast<uintptr_t>(g) & 15))) return false;
    bool ok = false;
    SALT_DISPATCH(raw.dt, T, {
        if (g) ok = relu ? ring_se_pool_t<T, true, true>(st, raw, g, scale, shift, se) : ring_se_pool_t<T, true, false>(st, raw, g, scale, shift, se);
        else ok = relu ? ring_se_pool_t<T, false, true>(st, raw, g, scale, shift, se) : ring_se_pool_t<T, false, false>(st, raw, g, scale, shift, se);
    });
    return ok;
}

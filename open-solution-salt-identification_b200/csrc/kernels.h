// Host-side launchers of every kernel in the engine.  All launches are asynchronous on `st`.
#pragma once
#include "common.cuh"

// -------------------------------------------------------------------- convolution geometry
// out(y,x,k) = sum_{r,s,c} in(y*stride + r - pad, x*stride + s - pad, c) * W[k][c][r][s]   (zero outside `in`)
// `in` dims are PHYSICAL (borders included), so a replicate-padded tensor is convolved with pad = 0.
struct ConvGeom {
    int B, Hi, Wi, Ci;     // input (physical); Ci = channel count in memory (stem: 4, of which 3 real)
    int Ho, Wo, Co;
    int R, S, stride, pad;
};

// packed weights: fwd  wp [Co][R*S*Ci]  (T);  dgrad wpd [Ci][R*S*Co] (T)
void k_pack_weights(cudaStream_t st, DType dt, const float* w_master /*[Co][Ci_real][R][S]*/, void* wp, void* wpd,
                    int Co, int Ci_real, int Ci, int R, int S);

// BatchNorm statistics are reduced in two fixed-order stages so that a forward / backward pass is bit-reproducible run to run:
// every producing CTA owns one SLOT of per-channel partial sums, stats[slot][2*C] floats (slot = blockIdx.x; the arena is zeroed once
// per pass), and the finalize kernels add the slots in slot order in fp64.  No floating-point atomics on this path.
// (SALT_STAT_SLOTS, SALT_STAT_SLOTS_CONV: common.cuh)
void k_conv_fwd_simt(cudaStream_t st, DType dt, const void* in, const void* wp, const float* bias, void* out,
                     float* stats, const ConvGeom& g);
// out[2C] (fp64) = sum over slots, in slot order (C-ABI single-operator entry points; the engine uses the finalize kernels)
void k_stats_reduce(cudaStream_t st, const float* stats, int nslots, int C, double* out);
void k_conv_dgrad_simt(cudaStream_t st, DType dt, const void* gout, const void* wpd, void* gin, bool accumulate,
                       const ConvGeom& g);
void k_conv_wgrad_simt(cudaStream_t st, DType dt, const void* in, const void* gout, float* dw, int Ci_real,
                       const ConvGeom& g, int groups = 1);

// -------------------------------------------------------------------- fp32 parity mode on the tensor cores (split-bf16 operands)
// x = h + m + l with h = bf16(x), m = bf16(x - h), l = bf16(x - h - m): three bf16 terms carry 24 mantissa bits.  A convolution
// of fp32 tensors is the sum of the six bf16 x bf16 products h*h, h*m, m*h, m*m, h*l, l*h (the dropped ones are < 2^-24
// relative), accumulated in fp32 by tcgen05 - i.e. the SAME bf16 convolution kernel run over 6x the input channels:
//   activation segments [h | m | h | m | h | l]   x   weight segments [h | m | m | h | l | h]      (hh, mm, hm, mh, hl, lh)
// The h*h segment comes first: the kernels give it its own TMEM accumulator (tcgen05 truncates at the accumulator's magnitude,
// so the 2^-8 .. 2^-16 smaller correction terms are summed apart and added in the epilogue; a 64-channel K-block that straddles the
// hh / mm boundary only drags the negligible m*m products into the main accumulator).
// Measured against the reference forward: <= 6e-5 max-abs on the logits (plain kind::tf32 operands give 1e-2..1e-1).
void k_split6_act(cudaStream_t st, const float* in, void* out_bf16, size_t rows, int C);       // [rows][C] fp32 -> [rows][6C] bf16
void k_split6_weights(cudaStream_t st, const float* wp, void* wp6_bf16, size_t rows, int C);   // same, weight segment order

// -------------------------------------------------------------------- input / elementwise
void k_input_nchw_to_nhwc4(cudaStream_t st, DType dt, const float* x, void* out, int B, int H, int W);
// stem input adapter: fp32 NCHW [B,3,H,W] -> im2col patches of the 7x7 stride-2 pad-3 stem convolution,
// NHWC [B,H/2,W/2,160]: channel j = c*49 + r*7 + s (the reference weight layout [k][c][r][s] flattened), j >= 147 zero.
// The stem then runs as a 1x1 convolution over 160 channels on the tensor cores (forward and wgrad).
void k_stem_im2col(cudaStream_t st, DType dt, const float* x, void* patches, int B, int H, int W);
void k_zero(cudaStream_t st, void* p, size_t bytes);

// -------------------------------------------------------------------- data formats either side of the network (kernels_io.cu)
// raw u8 tile [th][tw] -> network input [3][S][S]: edge pad (top/left offsets as utils.py:308-313), /255, normalise, depth channels
struct TileGeom {
    int th, tw, S, top, left, hflip;
    float mean0, std0;         // only channel 0 survives AddDepthChannels (utils.py:494-500)
    double lin_step;           // 1/(S-1), the float64 step of np.linspace(0, 1, S)
};
static inline TileGeom make_tile_geom(int th, int tw, int S, float mean0, float std0, bool hflip) {
    TileGeom g;
    g.th = th; g.tw = tw; g.S = S; g.hflip = hflip ? 1 : 0; g.mean0 = mean0; g.std0 = std0;
    g.top = (S - th) / 2;                      // get_crop_pad_sequence: top = int(v/2)
    g.left = (S - tw) - (S - tw) / 2;          //                        left = h - int(h/2)
    g.lin_step = 1.0 / (double)(S - 1);
    return g;
}
void k_adapt_tiles(cudaStream_t st, const uint8_t* tiles, int B, const TileGeom& g, float* x_nchw);
void k_stem_im2col_tiles(cudaStream_t st, DType dt, const uint8_t* tiles, void* patches, int B, const TileGeom& g);
// column-major run-length encoding of u8 masks [B][H][W] (utils.py:99-111); runs int32 [B][cap][2] = (start from 1, length)
void k_rle_encode(cudaStream_t st, const uint8_t* mask, int B, int H, int W, int cap, int* runs, int* nruns);
// per image and threshold: |pred & gt|, |pred| (pred = sigmoid(logit[1]) cropped to T x T > thr), and |gt|; nthr <= 32
void k_validation_counts(cudaStream_t st, const float* logits, const float* logits_flip, int B, int K, int S, int T,
                         const uint8_t* gt, const double* thresholds, int nthr, int* inter, int* pred, int* gtsum);

struct BNRef {                 // device pointers describing one BatchNorm layer at run time
    int C;
    const float *gamma, *beta;
    float *rmean, *rvar;
    float *dgamma, *dbeta;
    float* sums;               // [SALT_STAT_SLOTS][2C] forward  per-CTA partials of sum(x), sum(x^2)
    float* bsums;              // [SALT_STAT_SLOTS_BWD][2C] backward per-block partials of sum(g), sum(g*xhat)
    int* bslots;               // [1] number of slots the producing kernel of this backward pass wrote (= its grid)
    float *scale, *shift;      // y = x*scale + shift
    float *mean, *invstd;      // batch statistics of the last training forward
    float *cb, *cc;            // backward coefficients: g_raw = scale*(g - cb - cc*(x-mean))
};

void k_bn_finalize_train(cudaStream_t st, const BNRef& bn, int nslots, double count, float momentum, float eps);
void k_bn_finalize_eval(cudaStream_t st, const BNRef& bn, float eps);
void k_bn_bwd_finalize(cudaStream_t st, const BNRef& bn, double count);

// out = [relu]( (raw*scale+shift) [* gate[n][c]]  [+ res | + res*rscale+rshift] ), out may carry a replicate border
void k_bn_apply(cudaStream_t st, const Tensor& raw, const float* scale, const float* shift, const Tensor* res,
                const float* rscale, const float* rshift, bool relu, const Tensor& out, const float* gate = nullptr);
// out[B,H/2,W/2,C] = avgpool2( relu(raw*scale+shift) )
void k_bn_relu_avgpool(cudaStream_t st, const Tensor& raw, const float* scale, const float* shift, const Tensor& out);
void k_avgpool_bwd(cudaStream_t st, const Tensor& gout /*[B,h,w,C]*/, const Tensor& gin /*[B,2h,2w,C]*/);

struct GatherSrc { const void* p; int H, W, C, f; };   // f = integer bilinear upsampling factor (1 = copy)
// out (replicate-bordered) = concat_c( upsample_f(src_i) )
void k_gather_fwd(cudaStream_t st, const Tensor& out, const GatherSrc* srcs, int nsrc);
// adjoint pieces of k_gather_fwd: gP is the gradient w.r.t. the bordered tensor
void k_fold_bwd(cudaStream_t st, const Tensor& gP, int c0, const Tensor& gsrc, bool accumulate);
void k_upsample_bwd(cudaStream_t st, const Tensor& gP, int c0, int f, const Tensor& gsrc, float* tmp, bool accumulate);
size_t upsample_bwd_tmp_floats(const Tensor& gP, int f, int Csrc);

// -------------------------------------------------------------------- scSE (base.py:82-117)
struct SERef {
    int C, Cr;                                  // channels, reduced channels (C/16)
    const float *w1, *b1, *w2, *b2, *ws, *bs;   // fc.0 [Cr][C], fc.2 [C][Cr], spatial fc [C], [1]
    float *dw1, *db1, *dw2, *db2, *dws, *dbs;
    float *gap, *hid, *cse;                     // [B][C], [B][Cr], [B][C]   (saved by forward)
    float *part, *G;                            // [B][chunks][C] pooling partial sums, [B][C] backward scratch
    float *dhid;                                // [B][Cr] backward scratch
    int chunks;                                 // pixel chunks per image in the pooling kernels (depends on H*W only)
};
void k_scse_fwd(cudaStream_t st, const Tensor& raw, const float* scale, const float* shift, const SERef& se,
                const Tensor& out);
// g_out -> g_bn (gradient w.r.t. the BN output feeding the block's last ReLU), accumulates bn.bsums and SE grads
void k_scse_bwd(cudaStream_t st, const Tensor& gout, const Tensor& raw, const BNRef& bn, const SERef& se,
                const Tensor& gbn);

// encoder SE module (SE-ResNet bottleneck): gates se.cse[n][c] = sigmoid(fc2(relu(fc1(mean_pix(raw*scale+shift)))));
// the product with the gate is fused into k_bn_apply / k_bn_bwd_* (their `gate` arguments)
void k_se_gate_fwd(cudaStream_t st, const Tensor& raw, const float* scale, const float* shift, const SERef& se);
// g = d loss / d (u*cse) with u = raw*scale+shift; accumulates the FC weight gradients, leaves se.G[n][c] (see k_bn_bwd_reduce)
void k_se_gate_bwd(cudaStream_t st, const Tensor& g, const Tensor& raw, const float* scale, const float* shift, const SERef& se);

// -------------------------------------------------------------------- final 1x1 conv (unet.py:84)
void k_final_fwd(cudaStream_t st, const Tensor& raw, const float* scale, const float* shift, const float* w,
                 const float* b, int K, float* logits_nchw);
void k_final_bwd(cudaStream_t st, const float* dlogits_nchw, const Tensor& raw, const BNRef& bn, const float* w,
                 int K, float* dw, float* db, const Tensor& gbn);

// -------------------------------------------------------------------- BN / ReLU backward
// g <- g * [mask > 0]   (in place)
void k_relu_mask_inplace(cudaStream_t st, const Tensor& g, const Tensor& mask);
// bsums += per-channel { sum(gm), sum(gm*xhat) },  gm = g' * [raw*scale+shift > 0] if self_mask else g',
// g' = g*gate[n][c] + addc[n][c] when gate != NULL (encoder SE), else g
void k_bn_bwd_reduce(cudaStream_t st, const Tensor& g, const Tensor& raw, const BNRef& bn, bool self_mask,
                     const float* gate = nullptr, const float* addc = nullptr);
// graw = scale*(gm - cb - cc*(raw-mean))
void k_bn_bwd_apply(cudaStream_t st, const Tensor& g, const Tensor& raw, const BNRef& bn, bool self_mask,
                    const Tensor& graw, const float* gate = nullptr, const float* addc = nullptr);

// -------------------------------------------------------------------- losses / prediction / optimiser
// Lovasz hinge with ELU (lovasz_losses.py:97-115): loss_out[0] = mean_b loss_b; dlogits = d loss / d logits
// P <= 32768 logits per image: one CTA sorts them in shared memory.  Larger images: sort_scratch (lovasz_sort_scratch_bytes)
// holds the keys / payloads of a global-memory bitonic sort.
size_t lovasz_sort_scratch_bytes(int B, int P);
void k_lovasz(cudaStream_t st, const float* logits, const float* target, int B, int P, float* per_image,
              float* loss_out, float* dlogits, void* sort_scratch = nullptr);
// 0.2*dice + 0.9*bce (models.py:331-340).  sums: scratch [3K+1] doubles.  If allreduce is wanted the caller
// reduces `sums` between the two stages.
void k_bce_dice_reduce(cudaStream_t st, const float* logits, const float* target, int B, int K, int HW, double* sums);
void k_bce_dice_finish(cudaStream_t st, const float* logits, const float* target, int B, int K, int HW,
                       const double* sums, double total_count, float dice_w, float bce_w, float grad_scale,
                       float* loss_out, float* dlogits);
// sigmoid (+ un-flipped h-flip copy mean) -> probs [B,K,S,S]; crop + (probs[:,1] > thr) -> mask u8 [B,T,T]
void k_predict(cudaStream_t st, const float* logits, const float* logits_flip, int B, int K, int S, int T,
               float thr, float* probs, uint8_t* mask);

void k_adam(cudaStream_t st, float* p, const float* g, float* m, float* v, size_t n, float lr, float wd, float b1,
            float b2, float eps, int step, float grad_scale);

/* libsaltunet - C ABI of the B200-native U-Net segmentation engine.
 *
 * Drop-in boundary for the hot path of neptune-ai/open-solution-salt-identification:
 *   common_blocks/models.py:67-208  SegmentationModel (fit loop body, transform, load/persist)
 * Every entry point below names the reference code it replaces.  All pointers are plain device pointers
 * (CUDA, current device) unless said otherwise; no framework types cross this boundary.  `stream` is a
 * cudaStream_t passed as void* (NULL = legacy default stream).  Functions return 0 on success, non-zero on
 * error; salt_last_error() then returns a description (thread-local).
 *
 * Tensor layouts at the boundary are the reference's: images/logits/targets fp32 NCHW; parameters fp32 in
 * the reference state_dict layout (conv weight [Cout][Cin][R][S]) inside one flat array whose table is
 * given by salt_tensor_info().
 */
#ifndef SALTUNET_H
#define SALTUNET_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct salt_engine salt_engine;

enum { SALT_PREC_FP32 = 0, SALT_PREC_BF16 = 1 };
enum { SALT_ARCH_UNET_RESNET = 0, SALT_ARCH_UNET_SERESNET = 1, SALT_ARCH_UNET_SERESNEXT = 2 };

typedef struct salt_config {
    int arch;            /* SALT_ARCH_UNET_RESNET   <- models.py:15-18 ARCHITECTURES['UNetResNet']   (unet.py:22-109)
                            SALT_ARCH_UNET_SERESNET <- models.py:19-24 ARCHITECTURES['UNetSeResNet'] (unet.py:112-172)
                            SALT_ARCH_UNET_SERESNEXT <- models.py:25-30 ARCHITECTURES['UNetSeResNetXt'] (unet.py:175-236,
                                                        encoders.py:86-118: se_resnext50_32x4d / se_resnext101_32x4d)     */
    int encoder_depth;   /* UNetResNet: 18 or 34 (encoders.py:10-13); UNetSeResNet: 50, 101 or 152 (encoders.py:52-57);
                            UNetSeResNetXt: 50 or 101 (encoders.py:90-95)                                           */
    int num_classes;     /* out_channels           <- models.py:182                                             */
    int max_batch;       /* largest batch any call will pass                                                    */
    int height, width;   /* network input size, multiples of 32 (128 for the 101x101 tiles, loaders.py)         */
    int precision;       /* SALT_PREC_FP32: fp32 storage + fp32 FMA (parity mode); SALT_PREC_BF16: bf16 storage */
    int use_tensor_cores;/* bf16 only: run eligible convolutions on the tcgen05 kernels                         */
} salt_config;

const char* salt_last_error(void);
const char* salt_version(void);

/* models.py:179-184 set_model(): build the network plan. */
int salt_create(const salt_config* cfg, salt_engine** out);
void salt_destroy(salt_engine* h);

/* Sizes of the flat arrays the caller must provide to salt_bind(). */
size_t salt_param_floats(const salt_engine* h);     /* trainable parameters (fp32)                      */
size_t salt_buffer_floats(const salt_engine* h);    /* BatchNorm running_mean / running_var (fp32)      */
size_t salt_workspace_bytes(const salt_engine* h);  /* activations, packed weights, scratch             */

/* state_dict table (models.py:196-208 load / toolkit persist): entry i has the reference's key, shape, and
 * its offset (in floats) inside the params (is_buffer=0) or buffers (is_buffer=1) array. */
int salt_num_tensors(const salt_engine* h);
int salt_tensor_info(const salt_engine* h, int i, char* name, int name_cap, int shape[4], int* ndim, size_t* offset,
                     size_t* numel, int* is_buffer);

/* Attach caller-owned device memory.  grads/adam_m/adam_v may be NULL for inference-only use. */
int salt_bind(salt_engine* h, float* params, float* grads, float* adam_m, float* adam_v, float* buffers,
              void* workspace, size_t workspace_bytes);
/* Call after writing params from outside (load_state_dict, broadcast). */
int salt_params_changed(salt_engine* h);

/* unet.py:89-109 UNetResNet.forward: x [B,3,H,W] -> logits [B,num_classes,H,W].
 * train=1: BatchNorm batch statistics (+ running-stat update) and activations kept for salt_backward;
 * train=0: running statistics (model.eval(), models.py:150). */
int salt_forward(salt_engine* h, const float* x_nchw, int batch, float* logits_nchw, int train, void* stream);

/* models.py:326-328 lovasz_loss -> lovasz_losses.py:81-115 (per image, ELU variant).
 * loss_out: 1 float; dlogits: d(mean loss)/d logits, same shape as logits. target: [B,C,H,W] fp32 of 0/1. */
int salt_loss_lovasz(salt_engine* h, const float* logits, const float* target, int batch, float* loss_out,
                     float* dlogits, void* stream);
/* models.py:331-340 mixed_dice_bce_loss (dice 0.2 over the whole batch + BCE-with-logits 0.9), two stages so
 * that data-parallel callers can all-reduce the 3*C+1 partial sums in between:
 *   sums[3c+0]=sum(p*t) sums[3c+1]=sum(p) sums[3c+2]=sum(t)  sums[3C]=sum of element-wise BCE. */
int salt_loss_bce_dice_reduce(salt_engine* h, const float* logits, const float* target, int batch, double* sums,
                              void* stream);
int salt_loss_bce_dice_finish(salt_engine* h, const float* logits, const float* target, int batch,
                              const double* sums, double total_elements, float grad_scale, float* loss_out,
                              float* dlogits, void* stream);

/* models.py:133 batch_loss.backward(): fills the flat gradient array (zeroed first). */
int salt_backward(salt_engine* h, const float* dlogits_nchw, void* stream);
/* The same pass in three consecutive segments (0: final + decoder + center, 1: encoder layer4 + layer3, 2: layer2 + layer1 + stem),
 * to be called in that order.  When segment k returns, the gradients of its parameters - the contiguous range
 * salt_grad_segment(k) of the flat gradient buffer, in floats - are final: a data-parallel caller starts their NCCL all-reduce
 * while the next segment computes.  Replaces the single gradient reduction nn.DataParallel performs after backward
 * (common_blocks/models.py:81-82, 133). */
int salt_backward_segment(salt_engine* h, const float* dlogits_nchw, int segment, void* stream);
int salt_grad_segment(const salt_engine* h, int segment, size_t* offset, size_t* numel);

/* models.py:74-75,134,289-297: torch.optim.Adam step with L2 `grad += wd*p` on every parameter.
 * grad_scale multiplies the stored gradients first (1/world after a SUM all-reduce). step counts from 1. */
int salt_adam_step(salt_engine* h, float lr, float weight_decay, float beta1, float beta2, float eps, int step,
                   float grad_scale, void* stream);

/* utils.py:173-174 sigmoid, loaders.py:751-760 + augmentation.py:155-176 (mean over {orig, h-flip} of the
 * un-flipped probabilities), postprocessing.py:24-43 crop_image + binarize.
 * logits_flip may be NULL (no TTA). probs [B,C,S,S] and/or mask u8 [B,crop,crop] may be NULL. */
int salt_predict(salt_engine* h, const float* logits, const float* logits_flip, int batch, int crop, float threshold,
                 float* probs, uint8_t* mask, void* stream);

/* ---- data formats either side of the network ------------------------------- */

/* loaders.py:607-612 (Grayscale(3) + ToTensor + Normalize + AddDepthChannels, utils.py:494-500) after
 * augmentation.py:247-281 InferencePad('edge') with the pad split of utils.py:308-313: raw u8 tiles [B,tile_h,tile_w]
 * -> fp32 NCHW [B,3,size,size]. Only channel 0's mean/std matter (channels 1, 2 are overwritten by the depth
 * channels). hflip: np.fliplr of the raw tile first (augmentation.py:143-147, TTA). Bit-exact with the reference. */
int salt_adapt_tiles(const uint8_t* tiles, int batch, int tile_h, int tile_w, int size, float mean0, float std0, int hflip,
                     float* x_nchw, void* stream);
/* salt_forward with that adapter fused into the stem: u8 tiles in (12x fewer input bytes than fp32 [B,3,S,S]). */
int salt_forward_tiles(salt_engine* h, const uint8_t* tiles, int batch, int tile_h, int tile_w, float mean0, float std0,
                       int hflip, float* logits_nchw, int train, void* stream);

/* utils.py:99-111 run_length_encoding (called from utils.py:68-75 create_submission / :78-79 encode_rle): masks u8
 * [B,height,width] -> runs int32 [B,cap_runs,2] = (start, length), pixels numbered from 1 in column-major order;
 * nruns[b] = number of runs of image b (may exceed cap_runs: then only the first cap_runs are stored). */
int salt_rle_encode(const uint8_t* mask, int batch, int height, int width, int cap_runs, int32_t* runs, int32_t* nruns,
                    void* stream);

/* callbacks.py:499-527 ValidationMonitor._get_validation_loss inner loop (callbacks.py:832-866 crop_image + binarize at
 * each threshold, metrics.py:8-64 on single-object masks): for image b and threshold k
 *   pred = sigmoid(logits[b,1]) (mean with the un-flipped logits_flip if given) cropped to crop x crop, > thresholds[k]
 *   inter[b,k] = |pred & gt[b]|, pred[b,k] = |pred|, gtsum[b] = |gt[b]|      (gt: u8 [B,crop,crop], nthr <= 32).
 * thresholds: HOST pointer to float64 values (np.linspace(0.5, 0.3, 21) in the reference). */
int salt_validation_counts(const float* logits, const float* logits_flip, int batch, int classes, int size, int crop,
                           const uint8_t* gt, const double* thresholds, int nthr, int32_t* inter, int32_t* pred,
                           int32_t* gtsum, void* stream);

/* Test/debug: copy a named internal activation ("stem","e2".."e5","center","d5".."d1","final_raw", "g_<name>")
 * as fp32 NCHW into out (may be NULL to query the shape only). */
int salt_get_activation(salt_engine* h, const char* name, float* out_nchw, int shape[4], void* stream);

/* Number of CUDA kernels this library has launched so far in this process. */
unsigned long long salt_launch_count(void);
/* ... of which launched as thread-block clusters (convolutions whose weight stages are TMA-multicast; env SALT_TC_CLUSTER=1 disables). */
unsigned long long salt_cluster_launch_count(void);

/* Measurement aid (bench.py roofline): bracket every convolution launch with CUDA events on its stream.
 * kernel_class: 0 = conv forward, 1 = conv dgrad, 2 = conv wgrad.  salt_profile_read synchronises and returns the
 * accumulated device milliseconds, algorithmic FLOPs (2*B*Ho*Wo*Cout*Cin*R*S per launch) and launch count. */
int salt_profile_enable(salt_engine* h, int on);
int salt_profile_read(salt_engine* h, int kernel_class, double* ms, double* flops, long long* launches);
/* the same restricted to one layer group: 0 stem, 1-4 encoder layer1..layer4, 5 center, 6-10 dec5..dec1, 11 hypercolumn + final;
 * -1 = all groups (bench.py: roofline.per_group) */
int salt_profile_read_group(salt_engine* h, int kernel_class, int layer_group, double* ms, double* flops, long long* launches);
/* every bracketed launch since salt_profile_enable(h, 1), in launch order: class, layer group, algorithmic work (FLOPs for classes
 * 0-2, bytes for the memory-bound classes) and device milliseconds; writes at most max_records entries, returns the total number. */
long long salt_profile_records(salt_engine* h, int* kernel_class, int* layer_group, double* work, double* ms, long long max_records);

/* ---- single-operator entry points (unit tests; tensors NHWC in the given precision) ------------------ */
typedef struct salt_conv_desc {
    int batch, in_h, in_w, in_c;   /* physical input extent (borders included) */
    int out_h, out_w, out_c;
    int kernel, stride, pad;
    int precision;
    int use_tensor_cores;
} salt_conv_desc;
/* w: fp32 [out_c][in_c][k][k] (reference layout). stats may be NULL, else 2*out_c doubles (sum, sum of squares). */
int salt_op_conv_forward(const salt_conv_desc* d, const void* in, const float* w, const float* bias, void* out,
                         double* stats, void* stream);
int salt_op_conv_dgrad(const salt_conv_desc* d, const void* gout, const float* w, void* gin, int accumulate,
                       void* stream);
int salt_op_conv_wgrad(const salt_conv_desc* d, const void* in, const void* gout, float* dw, void* stream);
int salt_op_adam(float* p, const float* g, float* m, float* v, size_t n, float lr, float wd, float b1, float b2,
                 float eps, int step, float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif

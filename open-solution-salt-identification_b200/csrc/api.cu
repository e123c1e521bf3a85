// extern "C" boundary of libsaltunet (declared in include/saltunet.h).
#include "../../include/saltunet.h"
#include "engine.h"
#include <string>
#include <cstring>
#include <exception>

struct salt_engine { Engine* e; };

static thread_local std::string g_err;
static int fail(const std::string& m) { g_err = m; return 1; }
static int check_cuda(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(std::string(what) + ": " + cudaGetErrorString(e));
    return 0;
}
#define SALT_TRY(fn_name, ...)                                         \
    try { __VA_ARGS__; } catch (const std::exception& ex) { return fail(std::string(fn_name) + ": " + ex.what()); } \
    return check_cuda(fn_name);

static int require_gpu() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); return fail("no CUDA device: libsaltunet has no CPU fallback"); }
    return 0;
}

extern "C" {

const char* salt_last_error(void) { return g_err.c_str(); }
const char* salt_version(void) { return "saltunet-b200 0.1 (sm_100a)"; }

int salt_create(const salt_config* cfg, salt_engine** out) {
    if (!cfg || !out) return fail("salt_create: null argument");
    if (cfg->arch != SALT_ARCH_UNET_RESNET && cfg->arch != SALT_ARCH_UNET_SERESNET && cfg->arch != SALT_ARCH_UNET_SERESNEXT)
        return fail("salt_create: unknown architecture (UNetResNet, UNetSeResNet and UNetSeResNetXt are implemented)");
    if (cfg->precision != SALT_PREC_FP32 && cfg->precision != SALT_PREC_BF16) return fail("salt_create: unknown precision");
    try {
        EngineConfig c;
        c.arch = cfg->arch == SALT_ARCH_UNET_SERESNEXT ? 2 : (cfg->arch == SALT_ARCH_UNET_SERESNET ? 1 : 0);
        c.depth = cfg->encoder_depth; c.num_classes = cfg->num_classes; c.max_batch = cfg->max_batch;
        c.H = cfg->height; c.W = cfg->width; c.dt = cfg->precision == SALT_PREC_FP32 ? DT_F32 : DT_BF16;
        c.use_tc = cfg->use_tensor_cores;
        salt_engine* h = new salt_engine();
        h->e = new Engine(c);
        *out = h;
    } catch (const std::exception& ex) { return fail(std::string("salt_create: ") + ex.what()); }
    return 0;
}
void salt_destroy(salt_engine* h) { if (h) { delete h->e; delete h; } }

size_t salt_param_floats(const salt_engine* h) { return h->e->param_floats(); }
size_t salt_buffer_floats(const salt_engine* h) { return h->e->buffer_floats(); }
size_t salt_workspace_bytes(const salt_engine* h) { return h->e->workspace_bytes(); }
int salt_num_tensors(const salt_engine* h) { return (int)h->e->tensors().size(); }
int salt_tensor_info(const salt_engine* h, int i, char* name, int name_cap, int shape[4], int* ndim, size_t* offset,
                     size_t* numel, int* is_buffer) {
    const auto& v = h->e->tensors();
    if (i < 0 || i >= (int)v.size()) return fail("salt_tensor_info: index out of range");
    const TensorInfo& t = v[i];
    if (name && name_cap > 0) { strncpy(name, t.name.c_str(), name_cap - 1); name[name_cap - 1] = 0; }
    for (int k = 0; k < 4; ++k) shape[k] = t.shape[k];
    *ndim = t.ndim; *offset = t.offset; *numel = t.numel; *is_buffer = t.is_buffer;
    return 0;
}
int salt_bind(salt_engine* h, float* params, float* grads, float* adam_m, float* adam_v, float* buffers, void* workspace,
              size_t workspace_bytes) {
    if (!params || !buffers || !workspace) return fail("salt_bind: params, buffers and workspace are required");
    if (require_gpu()) return 1;
    SALT_TRY("salt_bind", h->e->bind(params, grads, adam_m, adam_v, buffers, workspace, workspace_bytes));
}
int salt_params_changed(salt_engine* h) { h->e->mark_params_dirty(); return 0; }

int salt_forward(salt_engine* h, const float* x, int batch, float* logits, int train, void* stream) {
    if (require_gpu()) return 1;
    SALT_TRY("salt_forward", h->e->forward(x, batch, logits, train != 0, (cudaStream_t)stream));
}
int salt_loss_lovasz(salt_engine* h, const float* logits, const float* target, int batch, float* loss_out, float* dlogits,
                     void* stream) {
    const EngineConfig& c = h->e->config();
    int P = c.num_classes * c.H * c.W;
    if (batch > c.max_batch) return fail("salt_loss_lovasz: batch exceeds max_batch");
    SALT_TRY("salt_loss_lovasz", k_lovasz((cudaStream_t)stream, logits, target, batch, P, h->e->loss_scratch(), loss_out, dlogits,
                                          h->e->lovasz_sort_scratch()));
}
int salt_loss_bce_dice_reduce(salt_engine* h, const float* logits, const float* target, int batch, double* sums, void* stream) {
    const EngineConfig& c = h->e->config();
    SALT_TRY("salt_loss_bce_dice_reduce", k_bce_dice_reduce((cudaStream_t)stream, logits, target, batch, c.num_classes, c.H * c.W, sums));
}
int salt_loss_bce_dice_finish(salt_engine* h, const float* logits, const float* target, int batch, const double* sums,
                              double total_elements, float grad_scale, float* loss_out, float* dlogits, void* stream) {
    const EngineConfig& c = h->e->config();
    SALT_TRY("salt_loss_bce_dice_finish", k_bce_dice_finish((cudaStream_t)stream, logits, target, batch, c.num_classes, c.H * c.W, sums,
                                                          total_elements, 0.2f, 0.9f, grad_scale, loss_out, dlogits));
}
int salt_backward(salt_engine* h, const float* dlogits, void* stream) {
    SALT_TRY("salt_backward", h->e->backward(dlogits, (cudaStream_t)stream));
}
int salt_backward_segment(salt_engine* h, const float* dlogits, int segment, void* stream) {
    if (segment < 0 || segment > 2) return fail("salt_backward_segment: segment must be 0, 1 or 2");
    SALT_TRY("salt_backward_segment", h->e->backward(dlogits, (cudaStream_t)stream, segment));
}
int salt_grad_segment(const salt_engine* h, int segment, size_t* offset, size_t* numel) {
    SALT_TRY("salt_grad_segment", h->e->grad_segment(segment, offset, numel));
}
int salt_adam_step(salt_engine* h, float lr, float wd, float b1, float b2, float eps, int step, float grad_scale, void* stream) {
    SALT_TRY("salt_adam_step", h->e->adam(lr, wd, b1, b2, eps, step, grad_scale, (cudaStream_t)stream));
}
int salt_predict(salt_engine* h, const float* logits, const float* logits_flip, int batch, int crop, float threshold,
                 float* probs, uint8_t* mask, void* stream) {
    const EngineConfig& c = h->e->config();
    if (c.H != c.W) return fail("salt_predict: square inputs only");
    if (crop > c.H) return fail("salt_predict: crop larger than the network input");
    SALT_TRY("salt_predict", k_predict((cudaStream_t)stream, logits, logits_flip, batch, c.num_classes, c.H, crop, threshold, probs, mask));
}
int salt_adapt_tiles(const uint8_t* tiles, int batch, int tile_h, int tile_w, int size, float mean0, float std0, int hflip,
                     float* x_nchw, void* stream) {
    if (require_gpu()) return 1;
    if (batch < 0 || tile_h < 1 || tile_w < 1 || size < 2 || tile_h > size || tile_w > size)
        return fail("salt_adapt_tiles: tile must be non-empty and fit into the padded size");
    if (batch == 0) return 0;
    if (!tiles || !x_nchw) return fail("salt_adapt_tiles: null argument");
    SALT_TRY("salt_adapt_tiles", k_adapt_tiles((cudaStream_t)stream, tiles, batch, make_tile_geom(tile_h, tile_w, size, mean0, std0, hflip != 0), x_nchw));
}
int salt_forward_tiles(salt_engine* h, const uint8_t* tiles, int batch, int tile_h, int tile_w, float mean0, float std0, int hflip,
                       float* logits, int train, void* stream) {
    if (require_gpu()) return 1;
    if (!tiles || !logits) return fail("salt_forward_tiles: null argument");
    SALT_TRY("salt_forward_tiles", h->e->forward_tiles(tiles, batch, make_tile_geom(tile_h, tile_w, h->e->config().H, mean0, std0, hflip != 0),
                                                       logits, train != 0, (cudaStream_t)stream));
}
int salt_rle_encode(const uint8_t* mask, int batch, int height, int width, int cap_runs, int32_t* runs, int32_t* nruns, void* stream) {
    if (require_gpu()) return 1;
    if (batch < 0 || height < 1 || width < 1 || cap_runs < 1) return fail("salt_rle_encode: bad shape");
    if (batch == 0) return 0;
    if (!mask || !runs || !nruns) return fail("salt_rle_encode: null argument");
    SALT_TRY("salt_rle_encode", k_rle_encode((cudaStream_t)stream, mask, batch, height, width, cap_runs, runs, nruns));
}
int salt_validation_counts(const float* logits, const float* logits_flip, int batch, int classes, int size, int crop,
                           const uint8_t* gt, const double* thresholds, int nthr, int32_t* inter, int32_t* pred, int32_t* gtsum,
                           void* stream) {
    if (require_gpu()) return 1;
    if (batch == 0) return 0;
    if (!logits || !gt || !thresholds || !inter || !pred || !gtsum) return fail("salt_validation_counts: null argument");
    if (classes < 2) return fail("salt_validation_counts: the salt map is class 1, needs >= 2 classes (postprocessing.py:41-43)");
    if (nthr < 1 || nthr > 32) return fail("salt_validation_counts: 1..32 thresholds");
    if (crop < 1 || crop > size) return fail("salt_validation_counts: crop larger than the network output");
    if (batch == 0) return 0;
    SALT_TRY("salt_validation_counts", k_validation_counts((cudaStream_t)stream, logits, logits_flip, batch, classes, size, crop, gt,
                                                           thresholds, nthr, inter, pred, gtsum));
}
int salt_get_activation(salt_engine* h, const char* name, float* out, int shape[4], void* stream) {
    try {
        if (!h->e->get_activation(name, out, shape, (cudaStream_t)stream)) return fail(std::string("salt_get_activation: unknown tensor ") + name);
    } catch (const std::exception& ex) { return fail(ex.what()); }
    return check_cuda("salt_get_activation");
}
unsigned long long salt_launch_count(void) { return g_salt_launches; }
unsigned long long salt_cluster_launch_count(void) { return g_salt_cluster_launches; }
int salt_profile_enable(salt_engine* h, int on) { h->e->profile_enable(on != 0); return 0; }
int salt_profile_read(salt_engine* h, int kernel_class, double* ms, double* flops, long long* launches) {
    return salt_profile_read_group(h, kernel_class, -1, ms, flops, launches);
}
long long salt_profile_records(salt_engine* h, int* kernel_class, int* layer_group, double* work, double* ms, long long max_records) {
    return h->e->profile_records(kernel_class, layer_group, work, ms, max_records);
}
int salt_profile_read_group(salt_engine* h, int kernel_class, int layer_group, double* ms, double* flops, long long* launches) {
    if (kernel_class < 0 || kernel_class >= Engine::PROF_NCLASS) return fail("salt_profile_read: unknown kernel class");
    if (layer_group < -1 || layer_group >= Engine::PROF_NGROUPS) return fail("salt_profile_read: unknown layer group");
    h->e->profile_read(kernel_class, ms, flops, launches, layer_group);
    return check_cuda("salt_profile_read");
}

// ---------------------------------------------------------------------------------------- single operators
static ConvGeom to_geom(const salt_conv_desc* d, int ci_mem) {
    ConvGeom g;
    g.B = d->batch; g.Hi = d->in_h; g.Wi = d->in_w; g.Ci = ci_mem; g.Ho = d->out_h; g.Wo = d->out_w; g.Co = d->out_c;
    g.R = g.S = d->kernel; g.stride = d->stride; g.pad = d->pad;
    return g;
}
struct PackedTmp {
    void *wp = nullptr, *wpd = nullptr;
    PackedTmp(const salt_conv_desc* d, const float* w, cudaStream_t st) {
        DType dt = d->precision == SALT_PREC_FP32 ? DT_F32 : DT_BF16;
        size_t bytes = (size_t)d->out_c * d->kernel * d->kernel * d->in_c * dtype_size(dt);
        cudaMalloc(&wp, bytes); cudaMalloc(&wpd, bytes);
        k_pack_weights(st, dt, w, wp, wpd, d->out_c, d->in_c, d->in_c, d->kernel, d->kernel);
    }
    ~PackedTmp() { cudaDeviceSynchronize(); cudaFree(wp); cudaFree(wpd); }
};
int salt_op_conv_forward(const salt_conv_desc* d, const void* in, const float* w, const float* bias, void* out, double* stats,
                         void* stream) {
    if (require_gpu()) return 1;
    if (d->in_c % 4 || d->out_c % 4) return fail("salt_op_conv_forward: channel counts must be multiples of 4");
    cudaStream_t st = (cudaStream_t)stream;
    DType dt = d->precision == SALT_PREC_FP32 ? DT_F32 : DT_BF16;
    PackedTmp pk(d, w, st);
    ConvGeom g = to_geom(d, d->in_c);
    float* part = nullptr;                  // per-CTA partial slots of the BatchNorm sums (kernels.h), reduced in slot order below
    if (stats) {
        const size_t pb = sizeof(float) * SALT_STAT_SLOTS * 2 * g.Co;
        if (cudaMalloc(&part, pb) != cudaSuccess) return fail("salt_op_conv_forward: out of device memory");
        cudaMemsetAsync(part, 0, pb, st);
    }
    try {
        if (d->use_tensor_cores && dt == DT_F32) {
            // fp32 parity mode on the tensor cores: split-bf16 operands (kernels.h k_split6_*), fp32 output
            ConvGeom g6 = g; g6.Ci = 6 * g.Ci;
            if (!tc_conv_supported(g6, false)) { cudaFree(part); return fail("salt_op_conv_forward: geometry not supported by the tensor-core kernel"); }
            void *in6 = nullptr, *w6 = nullptr;
            const size_t rows_in = (size_t)g.B * g.Hi * g.Wi, rows_w = (size_t)g.Co * g.R * g.S;
            if (cudaMalloc(&in6, rows_in * g6.Ci * 2) != cudaSuccess || cudaMalloc(&w6, rows_w * g6.Ci * 2) != cudaSuccess) {
                cudaFree(in6); cudaFree(part); return fail("salt_op_conv_forward: out of device memory");
            }
            k_split6_act(st, (const float*)in, in6, rows_in, g.Ci);
            k_split6_weights(st, (const float*)pk.wp, w6, rows_w, g.Ci);
            k_conv_tc(st, in6, g.B, g.Hi, g.Wi, g6.Ci, w6, g.Co, g.R, g.S, g.stride, g.pad, out, g.Ho, g.Wo, bias, part, false, true, nullptr, g.Ci);
            cudaStreamSynchronize(st);
            cudaFree(in6); cudaFree(w6);
        } else if (d->use_tensor_cores) {
            if (!tc_conv_supported(g, false)) { cudaFree(part); return fail("salt_op_conv_forward: geometry not supported by the tensor-core kernel"); }
            k_conv_tc(st, in, g.B, g.Hi, g.Wi, g.Ci, pk.wp, g.Co, g.R, g.S, g.stride, g.pad, out, g.Ho, g.Wo, bias, part, false);
        } else {
            k_conv_fwd_simt(st, dt, in, pk.wp, bias, out, part, g);
        }
        if (stats) k_stats_reduce(st, part, SALT_STAT_SLOTS, g.Co, stats);
    } catch (const std::exception& ex) { cudaFree(part); return fail(std::string("salt_op_conv_forward: ") + ex.what()); }
    cudaStreamSynchronize(st);
    cudaFree(part);
    return check_cuda("salt_op_conv_forward");
}
int salt_op_conv_dgrad(const salt_conv_desc* d, const void* gout, const float* w, void* gin, int accumulate, void* stream) {
    if (require_gpu()) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    DType dt = d->precision == SALT_PREC_FP32 ? DT_F32 : DT_BF16;
    PackedTmp pk(d, w, st);
    ConvGeom g = to_geom(d, d->in_c);
    try {
        if (d->use_tensor_cores) {
            if (dt != DT_BF16 || !tc_conv_supported(g, true)) return fail("salt_op_conv_dgrad: geometry not supported by the tensor-core kernel");
            if (g.stride == 1)
                k_conv_tc(st, gout, g.B, g.Ho, g.Wo, g.Co, pk.wpd, g.Ci, g.R, g.S, 1, g.R - 1 - g.pad, gin, g.Hi, g.Wi, nullptr, nullptr, accumulate != 0);
            else
                k_conv_tc_dgrad_s2(st, gout, g.B, g.Ho, g.Wo, g.Co, pk.wpd, g.Ci, g.R, g.S, g.pad, gin, g.Hi, g.Wi, accumulate != 0);
        } else {
            k_conv_dgrad_simt(st, dt, gout, pk.wpd, gin, accumulate != 0, g);
        }
    } catch (const std::exception& ex) { return fail(std::string("salt_op_conv_dgrad: ") + ex.what()); }
    cudaStreamSynchronize(st);
    return check_cuda("salt_op_conv_dgrad");
}
int salt_op_conv_wgrad(const salt_conv_desc* d, const void* in, const void* gout, float* dw, void* stream) {
    if (require_gpu()) return 1;
    DType dt = d->precision == SALT_PREC_FP32 ? DT_F32 : DT_BF16;
    ConvGeom g = to_geom(d, d->in_c);
    try {
        if (d->use_tensor_cores) {
            if (dt != DT_BF16 || !tc_wgrad_supported(g)) return fail("salt_op_conv_wgrad: geometry not supported by the tensor-core kernel");
            float* dwp = nullptr;
            size_t nf = tc_wgrad_scratch_floats(g.Ci, g.Co, g.R * g.S);
            cudaMalloc(&dwp, nf * sizeof(float));
            cudaMemsetAsync(dwp, 0, nf * sizeof(float), (cudaStream_t)stream);
            k_conv_wgrad_tc((cudaStream_t)stream, in, gout, dwp, g);
            k_unpack_dw((cudaStream_t)stream, dwp, dw, g.Co, d->in_c, cdiv(g.Ci, 64) * 64, g.R * g.S, true);
            cudaStreamSynchronize((cudaStream_t)stream);
            cudaFree(dwp);
        } else {
            k_conv_wgrad_simt((cudaStream_t)stream, dt, in, gout, dw, d->in_c, g);
        }
    } catch (const std::exception& ex) { return fail(std::string("salt_op_conv_wgrad: ") + ex.what()); }
    cudaStreamSynchronize((cudaStream_t)stream);
    return check_cuda("salt_op_conv_wgrad");
}
int salt_op_adam(float* p, const float* g, float* m, float* v, size_t n, float lr, float wd, float b1, float b2, float eps,
                 int step, float grad_scale, void* stream) {
    if (require_gpu()) return 1;
    k_adam((cudaStream_t)stream, p, g, m, v, n, lr, wd, b1, b2, eps, step, grad_scale);
    return check_cuda("salt_op_adam");
}

}  // extern "C"

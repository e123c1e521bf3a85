// U-Net (ResNet-18/34 or SE-ResNet-50 encoder, scSE decoder, hypercolumn head) execution plan: static buffers, explicit
// forward and backward passes built from the kernels in kernels.h.  Mirrors the computation of the reference
// common_blocks/architectures/unet.py:44-109 (UNetResNet) and :112-172 (UNetSeResNet) - see DESIGN.md for the layer mapping.
#pragma once
#include <string>
#include <vector>
#include <memory>
#include "kernels.h"
#include "conv_tc.h"
#include "kernels_batched.h"

struct TensorInfo {            // one named entry of the reference state_dict
    std::string name;
    int shape[4];
    int ndim;
    size_t offset;             // in floats, inside the params (is_buffer=0) or buffers (is_buffer=1) flat array
    size_t numel;
    int is_buffer;
};

struct ConvLayer {
    int Ci = 0, Ci_real = 0, Co = 0, R = 1, S = 1, stride = 1, pad = 0;
    int groups = 1;            // > 1: grouped convolution run densely over block-diagonal packed weights (kernels_batched.h)
    size_t o_w = 0;
    long long o_b = -1;        // bias offset or -1
    void* wp = nullptr;        // packed fwd weights   [Co][R*S][Ci]
    void* wpd = nullptr;       // packed dgrad weights [Ci][R*S][Co]
    bool in_unpack_table = false;
    void* wp6 = nullptr;       // fp32 tensor-core parity mode: split-bf16 forward weights [Co][R*S][6*Ci] (kernels.h k_split6_weights)
    float* dwp = nullptr;      // tensor-core wgrad scratch [R*S][ceil64(Ci)][Co] fp32 (inside the zeroed-per-backward arena)
};
struct BNLayer {
    int C = 0;
    size_t o_gamma = 0, o_beta = 0, o_rm = 0, o_rv = 0;
    float *sums = nullptr, *bsums = nullptr;      // [SALT_STAT_SLOTS][2C] partial slots
    float *scale = nullptr, *shift = nullptr, *mean = nullptr, *invstd = nullptr, *cb = nullptr, *cc = nullptr;
    int* bslots = nullptr;
};
struct SELayer {
    int C = 0, Cr = 0;
    size_t o_w1 = 0, o_b1 = 0, o_w2 = 0, o_b2 = 0, o_ws = 0, o_bs = 0;
    float *gap = nullptr, *hid = nullptr, *cse = nullptr, *part = nullptr, *G = nullptr, *dhid = nullptr;
    int chunks = 1;
    bool spatial = true;       // decoder scSE has the spatial branch, the encoder SE module does not
};
struct GradBuf { Tensor g; bool fresh = true; };
struct Act { Tensor t; GradBuf* gb = nullptr; };

struct BasicBlock {
    int group = 0;             // profile group (encoder layer index)
    ConvLayer c1, c2, cd;
    BNLayer b1, b2, bd;
    bool down = false;
    Act* x = nullptr;
    Tensor raw1, a1, raw2, rawd;
    Act out;
};
// SE-ResNet bottleneck (pretrainedmodels senet.py SEResNetBottleneck; restated in oracle/senet_restated.py):
// 1x1 (stride here) -> 3x3 -> 1x1 (x4), each + BN, SE gate on the last BN output, + shortcut, ReLU
struct Bottleneck {
    int group = 0;
    ConvLayer c1, c2, c3, cd;
    BNLayer b1, b2, b3, bd;
    SELayer se;
    bool down = false;
    Act* x = nullptr;
    Tensor raw1, a1, raw2, a2, raw3, rawd;
    Act out;
};
struct ConvBnRelu {            // reference base.py:7-37 on a replicate-bordered input
    ConvLayer c;
    BNLayer bn;
    Tensor P, raw;
};
struct Source { Act* a; int f; };
struct DecoderBlock {
    ConvBnRelu u1, u2;
    SELayer se;
    std::vector<Source> srcs;
    Act out;
};

struct EngineConfig {
    int depth = 34, num_classes = 2, max_batch = 8, H = 128, W = 128;
    int arch = 0;              // 0 = UNetResNet (depth 18/34), 1 = UNetSeResNet (depth 50)
    DType dt = DT_F32;
    int use_tc = 0;
};

class Engine {
public:
    explicit Engine(const EngineConfig& cfg);
    ~Engine();

    const std::vector<TensorInfo>& tensors() const { return infos_; }
    size_t param_floats() const { return n_params_; }
    size_t buffer_floats() const { return n_buffers_; }
    size_t workspace_bytes() const { return ws_bytes_; }

    void bind(float* params, float* grads, float* m, float* v, float* buffers, void* ws, size_t ws_bytes);
    void forward(const float* x_nchw, int B, float* logits_nchw, bool train, cudaStream_t st);
    // same network, input given as raw u8 tiles [B][th][tw]: the loader's pad / normalise / depth-channel adapter is fused
    // into the stem's im2col
    void forward_tiles(const uint8_t* tiles, int B, const TileGeom& g, float* logits_nchw, bool train, cudaStream_t st);
    // seg = -1: the whole backward pass.  seg = 0, 1, 2: one of three consecutive SEGMENTS (0: final + decoder + center, 1: encoder
    // layer4 + layer3, 2: layer2 + layer1 + stem), called in that order; when segment k returns, the gradients of its parameters -
    // the contiguous range grad_segment(k) of the flat buffer - are final, so a data-parallel caller can start their all-reduce
    // while the next segment computes (reference models.py:81-82 reduces all gradients after backward)
    void backward(const float* dlogits_nchw, cudaStream_t st, int seg = -1);
    void grad_segment(int seg, size_t* offset, size_t* numel) const;
    void adam(float lr, float wd, float b1, float b2, float eps, int step, float grad_scale, cudaStream_t st);
    void mark_params_dirty() { packed_dirty_ = true; eval_coef_dirty_ = true; }
    // copy a named internal activation (NHWC T, border dropped) into fp32 NCHW; returns false if unknown
    bool get_activation(const std::string& name, float* out_nchw, int* shape4, cudaStream_t st);

    // ---- optional per-kernel-class timing with CUDA events on the launch stream (bench.py roofline)
    enum ProfClass { PROF_CONV_FWD = SALT_PROF_CONV_FWD, PROF_CONV_DGRAD = SALT_PROF_CONV_DGRAD, PROF_CONV_WGRAD = SALT_PROF_CONV_WGRAD,
                     PROF_NCLASS = SALT_PROF_NCLASS };
    void profile_enable(bool on);
    void prof_begin(int cls, double work, cudaStream_t st);       // (public: called through the g_salt_prof_* hooks)
    void prof_end(cudaStream_t st);
    // synchronises, then returns accumulated device time / algorithmic flops / launches since enable
    void set_group(int g);          // layer group of the launches that follow (profile records, NVTX range)
    void close_group();
    bool nvtx_group_open_ = false;
    void profile_read(int cls, double* ms, double* flops, long long* launches, int group = -1);
    long long profile_records(int* cls, int* group, double* work, double* ms, long long max_records);
    // layer groups of the profile records: 0 stem, 1-4 encoder layer1..layer4, 5 center, 6-10 dec5..dec1, 11 final
    enum { PROF_NGROUPS = 12 };

    const EngineConfig& config() const { return cfg_; }
    float* loss_scratch() { return loss_scratch_; }
    void* lovasz_sort_scratch() { return lovasz_sort_; }
    double* loss_sums() { return loss_sums_; }

private:
    // ---- plan construction
    size_t add_param(const std::string& name, std::vector<int> shape);
    size_t add_buffer(const std::string& name, std::vector<int> shape);
    ConvLayer make_conv(const std::string& wname, const std::string& bname, int ci, int co, int k, int stride, int pad, int ci_mem = -1,
                        int groups = 1);
    BNLayer make_bn(const std::string& prefix, int c);
    Tensor make_tensor(int H, int W, int C, int pt = 0, int pb = 0, int pl = 0, int pr = 0);
    size_t make_tensor_bytes(int H, int W, int C) const;
    GradBuf* make_gradbuf(const Tensor& like);
    void* ws_alloc(size_t bytes);
    void need_scratch(int i, size_t bytes) { if (bytes > scratch_bytes_[i]) scratch_bytes_[i] = bytes; }
    void build();
    void build_cbr(ConvBnRelu& u, const std::string& prefix, int ci, int co, int H, int W);
    void build_decoder(DecoderBlock& d, const std::string& name, std::vector<Source> srcs, int cm, int co, int H, int W);

    // ---- run time
    BNRef bn_ref(const BNLayer& b) const;
    SERef se_ref(const SELayer& s) const;
    Tensor view(const Tensor& t) const { Tensor r = t; r.B = B_; return r; }
    Tensor scratch(int i, int H, int W, int C, int pt = 0, int pb = 0, int pl = 0, int pr = 0) const;
    ConvGeom geom(const ConvLayer& c, const Tensor& in, const Tensor& out) const;
    void pack_all(cudaStream_t st);
    void conv_fwd(const ConvLayer& c, const Tensor& in, const Tensor& out, BNLayer* bn, bool train, cudaStream_t st);
    // eval mode: convolution + folded BatchNorm (+ residual) (+ ReLU) in ONE kernel, stored straight into `dst` (which may carry a
    // replicate border).  fusable(): the tensor-core kernels can run this layer (otherwise the caller takes the unfused path)
    bool fusable(const ConvLayer& c, const Tensor& in, const Tensor& out) const;
    void conv_bn_fused(const ConvLayer& c, const Tensor& in, const Tensor& dst, const BNLayer& bn, const Tensor* res, bool relu, cudaStream_t st);
    std::vector<BNLayer*> all_bns();
    void finalize_eval_all(cudaStream_t st);
    void conv_dgrad(const ConvLayer& c, const Tensor& gout, const Tensor& gin, bool accumulate, cudaStream_t st);
    void conv_wgrad(ConvLayer& c, const Tensor& in, const Tensor& gout, cudaStream_t st);
    std::vector<ConvLayer*> all_convs();
    void build_pack_table();
    void unpack_all(cudaStream_t st, int table = 3);      // table 0-2: the convolutions of one backward segment, 3: all
    void gather_fwd(const std::vector<Source>& srcs, const Tensor& P, cudaStream_t st);
    void gather_bwd(const std::vector<Source>& srcs, const Tensor& gP, cudaStream_t st);
    void block_fwd(BasicBlock& b, bool train, cudaStream_t st);
    void block_bwd(BasicBlock& b, cudaStream_t st);
    void bneck_fwd(Bottleneck& b, bool train, cudaStream_t st);
    void bneck_bwd(Bottleneck& b, cudaStream_t st);
    SELayer make_se(const std::string& w1, const std::string& b1, const std::string& w2, const std::string& b2, int C, int HW, bool conv_shape);
    void cbr_fwd(ConvBnRelu& u, bool train, cudaStream_t st);
    void decoder_fwd(DecoderBlock& d, bool train, cudaStream_t st);
    void decoder_bwd(DecoderBlock& d, cudaStream_t st);

    EngineConfig cfg_;
    std::vector<TensorInfo> infos_;
    size_t n_params_ = 0, n_buffers_ = 0;
    size_t ws_bytes_ = 0, ws_cursor_ = 0;
    bool counting_ = true;
    char* ws_base_ = nullptr;
    float *params_ = nullptr, *grads_ = nullptr, *adam_m_ = nullptr, *adam_v_ = nullptr, *buffers_ = nullptr;
    bool packed_dirty_ = true;
    // batched pack / unpack tables (device copies live in the workspace)
    PackDesc* d_pack_ = nullptr; int* d_pack_start_ = nullptr; int pack_layers_ = 0, pack_blocks_ = 0;
    int pack_max_rs_ = 1;
    UnpackDesc* d_unpack_[4] = {nullptr, nullptr, nullptr, nullptr}; int* d_unpack_start_[4] = {nullptr, nullptr, nullptr, nullptr};
    int unpack_layers_[4] = {0, 0, 0, 0}, unpack_blocks_[4] = {0, 0, 0, 0}, unpack_max_rs_[4] = {1, 1, 1, 1};
    size_t seg_bound_[4] = {0, 0, 0, 0};      // parameter offsets: segment 2 = [0, b1), segment 1 = [b1, b2), segment 0 = [b2, n_params)
    bool unpack_table_dirty_ = false;
    bool trained_forward_ = false;
    bool eval_coef_dirty_ = true;      // eval-mode BatchNorm (scale, shift) must be recomputed from the running statistics
    bool fuse_eval_ = true;            // env SALT_ENGINE_FUSE_EVAL=0: eval forward through the separate BatchNorm passes (A/B testing)
    float *ones_ = nullptr, *zeros_ = nullptr;   // identity BatchNorm coefficients for consumers of already-normalised tensors
    void forward_body(int B, float* logits_nchw, bool train, cudaStream_t st);
    int B_ = 0;

    // network
    Tensor x4_;
    ConvLayer stem_;
    BNLayer stem_bn_;
    Tensor stem_raw_;
    Act stem_out_;
    std::vector<std::unique_ptr<BasicBlock>> blocks_;
    std::vector<std::unique_ptr<Bottleneck>> bnecks_;
    Act* enc_out_[4] = {nullptr, nullptr, nullptr, nullptr};
    std::vector<Source> center_src_;
    ConvBnRelu center0_, center1_;
    Act center_out_;
    DecoderBlock dec_[5];            // dec5, dec4, dec3, dec2, dec1
    std::vector<Source> final_src_;
    ConvBnRelu final0_;
    size_t o_final_w_ = 0, o_final_b_ = 0;
    std::vector<std::unique_ptr<GradBuf>> gradbufs_;

    // stats arenas (zeroed per pass)
    float* stats_arena_ = nullptr; size_t stats_floats_ = 0;
    float* bstats_arena_ = nullptr; size_t bstats_floats_ = 0;
    size_t stats_cursor_ = 0, bstats_cursor_ = 0;
    float* dwp_arena_ = nullptr; size_t dwp_floats_ = 0, dwp_cursor_ = 0;
    // scratch pool
    void* scratch_[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t scratch_bytes_[4] = {0, 0, 0, 0};
    struct ProfRec { cudaEvent_t a, b; double flops; int cls, group; };
    int prof_group_ = 0;
    std::vector<ProfRec> prof_;
    bool prof_on_ = false;
    int prof_depth_ = 0;       // only the outermost scope of nested launchers is recorded
    void* split_scratch_ = nullptr; size_t split_elems_ = 0;      // fp32 TC parity mode: split-bf16 copy of the current conv input
    bool split_tc() const { return cfg_.dt == DT_F32 && cfg_.use_tc; }
    float* loss_scratch_ = nullptr;   // [max_batch + 8]
    void* lovasz_sort_ = nullptr;     // global-memory sort buffers of the Lovasz loss for images with > 32768 logits
    double* loss_sums_ = nullptr;     // [16]
};

extern unsigned long long g_salt_launches;

#include "engine.h"
#include <nvtx3/nvToolsExt.h>
#include <stdexcept>
#include <cstring>
#include <algorithm>
#include <cstdlib>

unsigned long long g_salt_launches = 0;
void (*g_salt_prof_begin)(int, double, cudaStream_t) = nullptr;
void (*g_salt_prof_end)(cudaStream_t) = nullptr;
static Engine* g_prof_engine = nullptr;
unsigned long long g_salt_cluster_launches = 0;

static const float BN_EPS = 1e-5f, BN_MOMENTUM = 0.1f;
static const int MAX_CONVS = 192;       // capacity of the batched pack / unpack descriptor tables

// ------------------------------------------------------------------------------------------------
// plan construction
// ------------------------------------------------------------------------------------------------
Engine::Engine(const EngineConfig& cfg) : cfg_(cfg) {
    if (cfg.arch == 0 && cfg.depth != 18 && cfg.depth != 34)
        throw std::runtime_error("UNetResNet: only encoder_depth 18 and 34 are implemented in this engine");
    if (cfg.arch == 1 && cfg.depth != 50 && cfg.depth != 101 && cfg.depth != 152)
        throw std::runtime_error("UNetSeResNet: encoder_depth must be 50, 101 or 152 (reference encoders.py:52-59)");
    if (cfg.arch == 2 && cfg.depth != 50 && cfg.depth != 101)
        throw std::runtime_error("UNetSeResNetXt: encoder_depth must be 50 or 101 (reference encoders.py:90-95)");
    if (cfg.arch < 0 || cfg.arch > 2) throw std::runtime_error("unknown architecture id");
    if (cfg.H % 32 || cfg.W % 32) throw std::runtime_error("input height/width must be multiples of 32");
    if (cfg.num_classes < 1 || cfg.num_classes > 4) throw std::runtime_error("num_classes must be 1..4");
    counting_ = true;
    build();            // pass 1: learns the BN-statistics arena and scratch sizes
    build();            // pass 2: final layout
    ws_bytes_ = ws_cursor_;
}
Engine::~Engine() {}

size_t Engine::add_param(const std::string& name, std::vector<int> shape) {
    TensorInfo ti; ti.name = name; ti.ndim = (int)shape.size(); ti.numel = 1; ti.is_buffer = 0;
    for (int i = 0; i < 4; ++i) ti.shape[i] = i < ti.ndim ? shape[i] : 1;
    for (int s : shape) ti.numel *= s;
    ti.offset = n_params_;
    n_params_ += (ti.numel + 3) / 4 * 4;      // keep every tensor 16-byte aligned
    if (counting_) infos_.push_back(ti);
    return ti.offset;
}
size_t Engine::add_buffer(const std::string& name, std::vector<int> shape) {
    TensorInfo ti; ti.name = name; ti.ndim = (int)shape.size(); ti.numel = 1; ti.is_buffer = 1;
    for (int i = 0; i < 4; ++i) ti.shape[i] = i < ti.ndim ? shape[i] : 1;
    for (int s : shape) ti.numel *= s;
    ti.offset = n_buffers_;
    n_buffers_ += (ti.numel + 3) / 4 * 4;
    if (counting_) infos_.push_back(ti);
    return ti.offset;
}
void* Engine::ws_alloc(size_t bytes) {
    size_t off = (ws_cursor_ + 255) / 256 * 256;
    ws_cursor_ = off + bytes;
    return counting_ ? nullptr : (void*)(ws_base_ + off);
}
Tensor Engine::make_tensor(int H, int W, int C, int pt, int pb, int pl, int pr) {
    Tensor t; t.B = cfg_.max_batch; t.H = H; t.W = W; t.C = C; t.pt = pt; t.pb = pb; t.pl = pl; t.pr = pr; t.dt = cfg_.dt;
    t.p = ws_alloc(t.bytes());
    if (t.numel() > split_elems_) split_elems_ = t.numel();
    return t;
}
GradBuf* Engine::make_gradbuf(const Tensor& like) {
    gradbufs_.emplace_back(new GradBuf());
    GradBuf* g = gradbufs_.back().get();
    g->g = make_tensor(like.H, like.W, like.C);
    return g;
}
ConvLayer Engine::make_conv(const std::string& wname, const std::string& bname, int ci, int co, int k, int stride, int pad, int ci_mem,
                            int groups) {
    ConvLayer c; c.Ci_real = ci; c.Ci = ci_mem > 0 ? ci_mem : ci; c.Co = co; c.R = c.S = k; c.stride = stride; c.pad = pad;
    c.groups = groups;
    if (ci % groups || co % groups) throw std::runtime_error("grouped convolution: channel counts must be multiples of the group count");
    c.o_w = add_param(wname, {co, ci / groups, k, k});
    c.o_b = bname.empty() ? -1 : (long long)add_param(bname, {co});
    size_t wb = (size_t)co * k * k * c.Ci * dtype_size(cfg_.dt);
    c.wp = ws_alloc(wb);
    c.wpd = ws_alloc(wb);
    if (split_tc()) c.wp6 = ws_alloc((size_t)co * k * k * c.Ci * 6 * sizeof(bf16));
    c.dwp = dwp_arena_ + dwp_cursor_;
    dwp_cursor_ += (tc_wgrad_scratch_floats(c.Ci, co, k * k) + 3) / 4 * 4;
    return c;
}
BNLayer Engine::make_bn(const std::string& prefix, int c) {
    BNLayer b; b.C = c;
    b.o_gamma = add_param(prefix + ".weight", {c});
    b.o_beta = add_param(prefix + ".bias", {c});
    b.o_rm = add_buffer(prefix + ".running_mean", {c});
    b.o_rv = add_buffer(prefix + ".running_var", {c});
    b.sums = stats_arena_ + stats_cursor_; stats_cursor_ += (size_t)SALT_STAT_SLOTS * 2 * c;
    b.bsums = bstats_arena_ + bstats_cursor_; bstats_cursor_ += (size_t)SALT_STAT_SLOTS_BWD * 2 * c;
    float* f = (float*)ws_alloc(sizeof(float) * 6 * c + 16);
    b.bslots = (int*)(f + 6 * c);
    b.scale = f; b.shift = f + c; b.mean = f + 2 * c; b.invstd = f + 3 * c; b.cb = f + 4 * c; b.cc = f + 5 * c;
    return b;
}
void Engine::build_cbr(ConvBnRelu& u, const std::string& prefix, int ci, int co, int H, int W) {
    // state_dict order of the reference: batch_norm.* then conv.* (base.py:24-27)
    u.bn = make_bn(prefix + ".batch_norm", co);
    u.c = make_conv(prefix + ".conv.weight", prefix + ".conv.bias", ci, co, 3, 1, 0);
    u.P = make_tensor(H, W, ci, 2, 0, 0, 2);          // ReplicationPad2d((0,2,2,0)): 2 rows on top, 2 columns on the right
    u.raw = make_tensor(H, W, co);
    need_scratch(0, u.raw.bytes()); need_scratch(1, u.raw.bytes());
    need_scratch(2, u.P.bytes());
}
SELayer Engine::make_se(const std::string& w1, const std::string& b1, const std::string& w2, const std::string& b2, int C, int HW,
                        bool conv_shape) {
    SELayer se; se.C = C; se.Cr = C / 16; se.spatial = !conv_shape;
    const int Cr = se.Cr, B = cfg_.max_batch;
    se.o_w1 = conv_shape ? add_param(w1, {Cr, C, 1, 1}) : add_param(w1, {Cr, C});
    se.o_b1 = add_param(b1, {Cr});
    se.o_w2 = conv_shape ? add_param(w2, {C, Cr, 1, 1}) : add_param(w2, {C, Cr});
    se.o_b2 = add_param(b2, {C});
    se.chunks = std::max(1, std::min(64, HW / 128));
    float* f = (float*)ws_alloc(sizeof(float) * B * ((size_t)(3 + se.chunks) * C + 2 * Cr));
    se.gap = f; se.cse = f + (size_t)B * C; se.G = f + (size_t)2 * B * C; se.part = f + (size_t)3 * B * C;
    se.hid = f + (size_t)(3 + se.chunks) * B * C; se.dhid = se.hid + (size_t)B * Cr;
    return se;
}
void Engine::build_decoder(DecoderBlock& d, const std::string& name, std::vector<Source> srcs, int cm, int co, int H, int W) {
    int ci = 0;
    for (auto& s : srcs) ci += s.a->t.C;
    d.srcs = srcs;
    build_cbr(d.u1, name + ".conv1", ci, cm, H, W);
    build_cbr(d.u2, name + ".conv2", cm, co, H, W);
    d.se = make_se(name + ".channel_se.fc.0.weight", name + ".channel_se.fc.0.bias", name + ".channel_se.fc.2.weight",
                   name + ".channel_se.fc.2.bias", co, H * W, false);
    d.se.o_ws = add_param(name + ".spatial_se.fc.weight", {1, co, 1, 1});
    d.se.o_bs = add_param(name + ".spatial_se.fc.bias", {1});
    d.out.t = make_tensor(H, W, co);
    d.out.gb = make_gradbuf(d.out.t);
    for (auto& s : srcs)
        if (s.f > 1) need_scratch(3, sizeof(float) * (size_t)cfg_.max_batch * (H + 2) * (W / s.f) * s.a->t.C);
}
size_t Engine::make_tensor_bytes(int H, int W, int C) const { return (size_t)cfg_.max_batch * H * W * C * dtype_size(cfg_.dt); }

void Engine::build() {
    n_params_ = n_buffers_ = 0; ws_cursor_ = 0; stats_cursor_ = bstats_cursor_ = 0; dwp_cursor_ = 0;
    blocks_.clear(); bnecks_.clear(); gradbufs_.clear();
    if (counting_) infos_.clear();
    const int B = cfg_.max_batch, H = cfg_.H, W = cfg_.W;
    const int nblk18[4] = {2, 2, 2, 2}, nblk34[4] = {3, 4, 6, 3}, nblk101[4] = {3, 4, 23, 3}, nblk152[4] = {3, 8, 36, 3};
    // ResNet-34 and SE-ResNet-50 share [3,4,6,3]; se_resnet101 / se_resnet152 differ only in the block counts
    const int* nblk = cfg_.depth == 18 ? nblk18 : cfg_.depth == 101 ? nblk101 : cfg_.depth == 152 ? nblk152 : nblk34;
    const int chans[4] = {64, 128, 256, 512};

    // BN statistic arenas: sized on the counting pass
    stats_arena_ = (float*)ws_alloc(sizeof(float) * std::max<size_t>(stats_floats_, 1));
    bstats_arena_ = (float*)ws_alloc(sizeof(float) * std::max<size_t>(bstats_floats_, 1));
    dwp_arena_ = (float*)ws_alloc(sizeof(float) * std::max<size_t>(cfg_.dt == DT_BF16 ? dwp_floats_ : 0, 4));

    // ---- encoder, parameter order = state_dict order
    //   arch 0: reference encoders.py:10-45, torchvision ResNet-18/34 (BasicBlock)
    //   arch 1: reference encoders.py:48-83, pretrainedmodels se_resnet50 (layer0 = conv7x7/2 + BN + ReLU, SEResNetBottleneck x [3,4,6,3])
    const bool se50 = cfg_.arch >= 1;           // bottleneck encoders: SE-ResNet (arch 1) and SE-ResNeXt 32x4d (arch 2)
    const bool resnext = cfg_.arch == 2;
    const std::string e = "encoders.encoder.";
    // the 7x7 stride-2 stem runs as a 1x1 convolution over im2col patches (160 channels, 147 real): see k_stem_im2col
    x4_ = make_tensor(H / 2, W / 2, 160);
    stem_ = make_conv(e + (se50 ? "layer0.conv1.weight" : "conv1.weight"), "", 3, 64, 7, 2, 3, 4);
    {
        stem_.Ci = 160; stem_.Ci_real = 147; stem_.R = stem_.S = 1; stem_.stride = 1; stem_.pad = 0;
        const size_t wb = (size_t)64 * 160 * dtype_size(cfg_.dt);
        stem_.wp = ws_alloc(wb); stem_.wpd = ws_alloc(wb);
        if (split_tc()) stem_.wp6 = ws_alloc((size_t)64 * 160 * 6 * sizeof(bf16));
        stem_.dwp = dwp_arena_ + dwp_cursor_;
        dwp_cursor_ += (tc_wgrad_scratch_floats(160, 64, 1) + 3) / 4 * 4;
    }
    stem_bn_ = make_bn(e + (se50 ? "layer0.bn1" : "bn1"), 64);
    stem_raw_ = make_tensor(H / 2, W / 2, 64);
    stem_out_.t = make_tensor(H / 2, W / 2, 64);
    stem_out_.gb = make_gradbuf(stem_out_.t);
    need_scratch(1, stem_raw_.bytes());
    Act* cur = &stem_out_;
    int h = H / 2, w = W / 2, cin = 64;
    for (int li = 0; li < 4; ++li) {
        for (int bi = 0; bi < nblk[li]; ++bi) {
            const std::string p = e + "layer" + std::to_string(li + 1) + "." + std::to_string(bi) + ".";
            const int stride = (bi == 0 && li > 0) ? 2 : 1;
            if (!se50) {
                blocks_.emplace_back(new BasicBlock());
                BasicBlock& b = *blocks_.back();
                b.group = li + 1;
                const int co = chans[li];
                b.down = (bi == 0 && li > 0);
                b.x = cur;
                b.c1 = make_conv(p + "conv1.weight", "", cin, co, 3, stride, 1);
                b.b1 = make_bn(p + "bn1", co);
                b.c2 = make_conv(p + "conv2.weight", "", co, co, 3, 1, 1);
                b.b2 = make_bn(p + "bn2", co);
                if (b.down) {
                    b.cd = make_conv(p + "downsample.0.weight", "", cin, co, 1, stride, 0);
                    b.bd = make_bn(p + "downsample.1", co);
                }
                h /= stride; w /= stride;
                b.raw1 = make_tensor(h, w, co); b.a1 = make_tensor(h, w, co); b.raw2 = make_tensor(h, w, co);
                if (b.down) b.rawd = make_tensor(h, w, co);
                b.out.t = make_tensor(h, w, co);
                b.out.gb = b.down ? make_gradbuf(b.out.t) : cur->gb;      // identity blocks pass the gradient buffer through
                need_scratch(0, b.raw1.bytes()); need_scratch(1, b.raw1.bytes()); need_scratch(2, b.raw1.bytes());
                cur = &b.out; cin = co;
            } else {
                bnecks_.emplace_back(new Bottleneck());
                Bottleneck& b = *bnecks_.back();
                b.group = li + 1;
                const int pl = chans[li], co = 4 * pl;
                b.down = (bi == 0);
                b.x = cur;
                // SE-ResNet (senet.py SEResNetBottleneck): 1x1 stride s -> 3x3 -> 1x1, width = planes.
                // SE-ResNeXt 32x4d (SEResNeXtBottleneck, base_width 4): 1x1 -> 3x3 stride s, 32 groups -> 1x1, width = planes * 2.
                const int wd = resnext ? 2 * pl : pl;
                const int hin = h, win = w;
                h /= stride; w /= stride;
                const int h1 = resnext ? hin : h, w1 = resnext ? win : w;       // resolution of conv1's output
                b.c1 = make_conv(p + "conv1.weight", "", cin, wd, 1, resnext ? 1 : stride, 0);
                b.b1 = make_bn(p + "bn1", wd);
                b.c2 = make_conv(p + "conv2.weight", "", wd, wd, 3, resnext ? stride : 1, 1, -1, resnext ? 32 : 1);
                b.b2 = make_bn(p + "bn2", wd);
                b.c3 = make_conv(p + "conv3.weight", "", wd, co, 1, 1, 0);
                b.b3 = make_bn(p + "bn3", co);
                b.se = make_se(p + "se_module.fc1.weight", p + "se_module.fc1.bias", p + "se_module.fc2.weight", p + "se_module.fc2.bias",
                               co, h * w, true);
                if (b.down) {
                    b.cd = make_conv(p + "downsample.0.weight", "", cin, co, 1, stride, 0);
                    b.bd = make_bn(p + "downsample.1", co);
                }
                b.raw1 = make_tensor(h1, w1, wd); b.a1 = make_tensor(h1, w1, wd);
                b.raw2 = make_tensor(h, w, wd); b.a2 = make_tensor(h, w, wd);
                b.raw3 = make_tensor(h, w, co);
                if (b.down) b.rawd = make_tensor(h, w, co);
                b.out.t = make_tensor(h, w, co);
                b.out.gb = b.down ? make_gradbuf(b.out.t) : cur->gb;
                need_scratch(0, b.raw3.bytes()); need_scratch(1, b.raw3.bytes()); need_scratch(2, b.raw3.bytes());
                need_scratch(0, b.raw1.bytes()); need_scratch(1, b.raw1.bytes());
                cur = &b.out; cin = co;
            }
        }
        enc_out_[li] = cur;
        if (li == 1) seg_bound_[1] = n_params_;      // end of stem + layer1 + layer2
        if (li == 3) seg_bound_[2] = n_params_;      // end of the encoder
    }
    // ---- center (unet.py:60-63 / :123-126); bc = bottom_channel_nr
    const int bc = se50 ? 2048 : 512, dc = bc / 8;
    center_src_ = {{enc_out_[3], 1}};
    build_cbr(center0_, "center.0", bc, bc, h, w);
    build_cbr(center1_, "center.1", bc, bc / 2, h, w);
    center_out_.t = make_tensor(h / 2, w / 2, bc / 2);
    center_out_.gb = make_gradbuf(center_out_.t);
    // ---- decoder (unet.py:65-79 / :128-143): dec5..dec1
    build_decoder(dec_[0], "dec5", {{&center_out_, 2}, {enc_out_[3], 1}}, bc, dc, H / 16, W / 16);
    build_decoder(dec_[1], "dec4", {{&dec_[0].out, 2}, {enc_out_[2], 1}}, bc / 2, dc, H / 8, W / 8);
    build_decoder(dec_[2], "dec3", {{&dec_[1].out, 2}, {enc_out_[1], 1}}, bc / 4, dc, H / 4, W / 4);
    build_decoder(dec_[3], "dec2", {{&dec_[2].out, 2}, {enc_out_[0], 1}}, bc / 8, dc, H / 2, W / 2);
    build_decoder(dec_[4], "dec1", {{&dec_[3].out, 2}}, bc / 16, dc, H, W);
    // ---- hypercolumn + final (unet.py:82-84, 101-109 / :145-172)
    final_src_ = {{&dec_[4].out, 1}, {&dec_[3].out, 2}, {&dec_[2].out, 4}, {&dec_[1].out, 8}, {&dec_[0].out, 16}};
    build_cbr(final0_, "final.0", 5 * dc, dc, H, W);
    for (auto& s : final_src_)
        if (s.f > 1) need_scratch(3, sizeof(float) * (size_t)B * (H + 2) * (W / s.f) * s.a->t.C);
    o_final_w_ = add_param("final.1.weight", {cfg_.num_classes, dc, 1, 1});
    o_final_b_ = add_param("final.1.bias", {cfg_.num_classes});

    d_pack_ = (PackDesc*)ws_alloc(sizeof(PackDesc) * MAX_CONVS); d_pack_start_ = (int*)ws_alloc(sizeof(int) * (MAX_CONVS + 1));
    for (int i = 0; i < 4; ++i) {
        d_unpack_[i] = (UnpackDesc*)ws_alloc(sizeof(UnpackDesc) * MAX_CONVS); d_unpack_start_[i] = (int*)ws_alloc(sizeof(int) * (MAX_CONVS + 1));
    }
    loss_scratch_ = (float*)ws_alloc(sizeof(float) * (B + 8));
    {
        const size_t sb = lovasz_sort_scratch_bytes(B, cfg_.num_classes * H * W);
        lovasz_sort_ = sb ? ws_alloc(sb) : nullptr;
    }
    loss_sums_ = (double*)ws_alloc(sizeof(double) * 16);
    for (int i = 0; i < 4; ++i) scratch_[i] = ws_alloc(std::max<size_t>(scratch_bytes_[i], 256));
    split_scratch_ = split_tc() ? ws_alloc(split_elems_ * 6 * sizeof(bf16)) : nullptr;
    ones_ = (float*)ws_alloc(sizeof(float) * 2 * 4096); zeros_ = ones_ + 4096;
    if (counting_) { stats_floats_ = stats_cursor_; bstats_floats_ = bstats_cursor_; dwp_floats_ = dwp_cursor_; }
}

void Engine::bind(float* params, float* grads, float* m, float* v, float* buffers, void* ws, size_t ws_bytes) {
    if (ws_bytes < ws_bytes_) throw std::runtime_error("workspace too small");
    params_ = params; grads_ = grads; adam_m_ = m; adam_v_ = v; buffers_ = buffers;
    ws_base_ = (char*)ws;
    counting_ = false;
    build();
    if (ws_cursor_ > ws_bytes_) throw std::runtime_error("internal error: workspace layout changed between passes");
    build_pack_table();
    unpack_table_dirty_ = true;
    packed_dirty_ = true;
    eval_coef_dirty_ = true;
    {
        std::vector<float> id(2 * 4096, 0.f);
        std::fill(id.begin(), id.begin() + 4096, 1.f);
        cudaMemcpy(ones_, id.data(), sizeof(float) * id.size(), cudaMemcpyHostToDevice);
        const char* e = getenv("SALT_ENGINE_FUSE_EVAL");
        fuse_eval_ = !(e && e[0] == '0');
    }
}

// ------------------------------------------------------------------------------------------------
// run-time helpers
// ------------------------------------------------------------------------------------------------
BNRef Engine::bn_ref(const BNLayer& b) const {
    BNRef r; r.C = b.C;
    r.gamma = params_ + b.o_gamma; r.beta = params_ + b.o_beta;
    r.rmean = buffers_ + b.o_rm; r.rvar = buffers_ + b.o_rv;
    r.dgamma = grads_ ? grads_ + b.o_gamma : nullptr; r.dbeta = grads_ ? grads_ + b.o_beta : nullptr;
    r.sums = b.sums; r.bsums = b.bsums; r.scale = b.scale; r.shift = b.shift; r.mean = b.mean; r.invstd = b.invstd;
    r.cb = b.cb; r.cc = b.cc; r.bslots = b.bslots;
    return r;
}
SERef Engine::se_ref(const SELayer& s) const {
    SERef r; r.C = s.C; r.Cr = s.Cr;
    r.w1 = params_ + s.o_w1; r.b1 = params_ + s.o_b1; r.w2 = params_ + s.o_w2; r.b2 = params_ + s.o_b2;
    r.ws = params_ + s.o_ws; r.bs = params_ + s.o_bs;
    float* g = grads_;
    r.dw1 = g ? g + s.o_w1 : nullptr; r.db1 = g ? g + s.o_b1 : nullptr; r.dw2 = g ? g + s.o_w2 : nullptr;
    r.db2 = g ? g + s.o_b2 : nullptr; r.dws = g ? g + s.o_ws : nullptr; r.dbs = g ? g + s.o_bs : nullptr;
    r.gap = s.gap; r.hid = s.hid; r.cse = s.cse; r.part = s.part; r.G = s.G; r.dhid = s.dhid; r.chunks = s.chunks;
    if (!s.spatial) { r.ws = r.bs = nullptr; r.dws = r.dbs = nullptr; }
    return r;
}
Tensor Engine::scratch(int i, int H, int W, int C, int pt, int pb, int pl, int pr) const {
    Tensor t; t.p = scratch_[i]; t.B = B_; t.H = H; t.W = W; t.C = C; t.pt = pt; t.pb = pb; t.pl = pl; t.pr = pr; t.dt = cfg_.dt;
    if (t.bytes() > scratch_bytes_[i]) throw std::runtime_error("internal error: scratch buffer too small");
    return t;
}
ConvGeom Engine::geom(const ConvLayer& c, const Tensor& in, const Tensor& out) const {
    ConvGeom g;
    g.B = B_; g.Hi = in.Hp(); g.Wi = in.Wp(); g.Ci = c.Ci; g.Ho = out.H; g.Wo = out.W; g.Co = c.Co;
    g.R = c.R; g.S = c.S; g.stride = c.stride; g.pad = c.pad;
    return g;
}
std::vector<ConvLayer*> Engine::all_convs() {
    std::vector<ConvLayer*> v;
    v.push_back(&stem_);
    for (auto& b : blocks_) { v.push_back(&b->c1); v.push_back(&b->c2); if (b->down) v.push_back(&b->cd); }
    for (auto& b : bnecks_) { v.push_back(&b->c1); v.push_back(&b->c2); v.push_back(&b->c3); if (b->down) v.push_back(&b->cd); }
    if ((int)v.size() + 13 > MAX_CONVS) throw std::runtime_error("internal error: MAX_CONVS too small");
    v.push_back(&center0_.c); v.push_back(&center1_.c);
    for (auto& d : dec_) { v.push_back(&d.u1.c); v.push_back(&d.u2.c); }
    v.push_back(&final0_.c);
    return v;
}
void Engine::build_pack_table() {
    std::vector<PackDesc> descs; std::vector<int> start(1, 0);
    for (ConvLayer* c : all_convs()) {
        PackDesc d; d.w = params_ + c->o_w; d.wp = c->wp; d.wpd = c->wpd; d.Co = c->Co; d.Ci_real = c->Ci_real; d.Ci = c->Ci; d.RS = c->R * c->S; d.groups = c->groups;
        descs.push_back(d);
        start.push_back(start.back() + pack_blocks(c->Co, c->Ci, c->R * c->S));
        pack_max_rs_ = std::max(pack_max_rs_, c->R * c->S);
    }
    pack_layers_ = (int)descs.size(); pack_blocks_ = start.back();
    cudaMemcpy(d_pack_, descs.data(), sizeof(PackDesc) * descs.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(d_pack_start_, start.data(), sizeof(int) * start.size(), cudaMemcpyHostToDevice);
}
void Engine::pack_all(cudaStream_t st) {
    k_pack_all(st, cfg_.dt, d_pack_, d_pack_start_, pack_layers_, pack_blocks_, pack_max_rs_);
    if (split_tc())
        for (ConvLayer* c : all_convs()) k_split6_weights(st, (const float*)c->wp, c->wp6, (size_t)c->Co * c->R * c->S, c->Ci);
    packed_dirty_ = false;
}
// one launch: transpose every tensor-core wgrad scratch into the reference-layout gradient
void Engine::unpack_all(cudaStream_t st, int table) {
    if (unpack_table_dirty_) {
        // all four tables (one per backward segment + the whole network) are rebuilt together, on the eager steps that precede
        // any CUDA-graph capture (the rebuild synchronises)
        cudaStreamSynchronize(st);
        for (int tb = 0; tb < 4; ++tb) {
            std::vector<UnpackDesc> descs; std::vector<int> start(1, 0);
            unpack_max_rs_[tb] = 1;
            for (ConvLayer* c : all_convs()) {
                if (!c->in_unpack_table) continue;
                const int seg = c->o_w >= seg_bound_[2] ? 0 : (c->o_w >= seg_bound_[1] ? 1 : 2);
                if (tb < 3 && seg != tb) continue;
                UnpackDesc d; d.dwp = c->dwp; d.dw = grads_ + c->o_w; d.Co = c->Co; d.Ci_real = c->Ci_real; d.Ci_pad = cdiv(c->Ci, 64) * 64; d.RS = c->R * c->S; d.groups = c->groups;
                descs.push_back(d);
                start.push_back(start.back() + cdiv(c->Co, 32) * cdiv(c->Ci_real, 32));
                unpack_max_rs_[tb] = std::max(unpack_max_rs_[tb], d.RS);
            }
            unpack_layers_[tb] = (int)descs.size(); unpack_blocks_[tb] = start.back();
            if (!descs.empty()) cudaMemcpy(d_unpack_[tb], descs.data(), sizeof(UnpackDesc) * descs.size(), cudaMemcpyHostToDevice);
            cudaMemcpy(d_unpack_start_[tb], start.data(), sizeof(int) * start.size(), cudaMemcpyHostToDevice);
        }
        unpack_table_dirty_ = false;
    }
    if (unpack_layers_[table] > 0)
        k_unpack_all(st, d_unpack_[table], d_unpack_start_[table], unpack_layers_[table], unpack_blocks_[table], unpack_max_rs_[table]);
}
void Engine::grad_segment(int seg, size_t* offset, size_t* numel) const {
    if (seg < 0 || seg > 2) throw std::runtime_error("grad_segment: segment must be 0, 1 or 2");
    const size_t lo = seg == 0 ? seg_bound_[2] : (seg == 1 ? seg_bound_[1] : 0);
    const size_t hi = seg == 0 ? n_params_ : (seg == 1 ? seg_bound_[2] : seg_bound_[1]);
    *offset = lo; *numel = hi - lo;
}
static double conv_flops(const ConvGeom& g, int ci_real) {
    return 2.0 * g.B * g.Ho * g.Wo * (double)g.Co * ci_real * g.R * g.S;
}
void Engine::profile_enable(bool on) {
    for (auto& r : prof_) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    prof_.clear();
    prof_on_ = on;
    prof_depth_ = 0;
    // the elementwise launchers (kernels_*.cu) report through these hooks
    g_prof_engine = on ? this : nullptr;
    g_salt_prof_begin = on ? +[](int cls, double work, cudaStream_t st) { if (g_prof_engine) g_prof_engine->prof_begin(cls, work, st); } : nullptr;
    g_salt_prof_end = on ? +[](cudaStream_t st) { if (g_prof_engine) g_prof_engine->prof_end(st); } : nullptr;
}
// NVTX: one range per pass ("salt.forward" / "salt.backward") and one per layer group inside it, so a timeline tool (nsys, ncu
// --nvtx) attributes the ~480 launches of a step to stem / layer1..4 / center / dec5..1 / final.  Host-side only: free when no
// tool is attached, invisible to CUDA-graph capture.
static const char* const kGroupNames[Engine::PROF_NGROUPS] = {"stem", "layer1", "layer2", "layer3", "layer4", "center",
                                                              "dec5", "dec4", "dec3", "dec2", "dec1", "final"};
void Engine::set_group(int g) {
    prof_group_ = g;
    if (nvtx_group_open_) nvtxRangePop();
    nvtxRangePushA(kGroupNames[g]);
    nvtx_group_open_ = true;
}
void Engine::close_group() {
    if (nvtx_group_open_) { nvtxRangePop(); nvtx_group_open_ = false; }
}
void Engine::prof_begin(int cls, double flops, cudaStream_t st) {
    if (!prof_on_) return;
    if (prof_depth_++ > 0) return;
    ProfRec r; r.flops = flops; r.cls = cls; r.group = prof_group_;
    cudaEventCreate(&r.a); cudaEventCreate(&r.b);
    cudaEventRecord(r.a, st);
    prof_.push_back(r);
}
void Engine::prof_end(cudaStream_t st) {
    if (!prof_on_) return;
    if (--prof_depth_ > 0) return;
    cudaEventRecord(prof_.back().b, st);
}
void Engine::profile_read(int cls, double* ms, double* flops, long long* launches, int group) {
    cudaDeviceSynchronize();
    *ms = 0; *flops = 0; *launches = 0;
    for (auto& r : prof_) {
        if (r.cls != cls || (group >= 0 && r.group != group)) continue;
        float t = 0.f;
        cudaEventElapsedTime(&t, r.a, r.b);
        *ms += t; *flops += r.flops; *launches += 1;
    }
}
long long Engine::profile_records(int* cls, int* group, double* work, double* ms, long long max_records) {
    cudaDeviceSynchronize();
    long long n = 0;
    for (auto& r : prof_) {
        if (n >= max_records) break;
        float t = 0.f;
        cudaEventElapsedTime(&t, r.a, r.b);
        cls[n] = r.cls; group[n] = r.group; work[n] = r.flops; ms[n] = t;
        ++n;
    }
    return (long long)prof_.size();
}
void Engine::conv_fwd(const ConvLayer& c, const Tensor& in, const Tensor& out, BNLayer* bn, bool train, cudaStream_t st) {
    ConvGeom g = geom(c, in, out);
    const float* bias = c.o_b >= 0 ? params_ + c.o_b : nullptr;
    prof_begin(PROF_CONV_FWD, conv_flops(g, c.Ci_real / c.groups), st);
    float* stats = (bn && train) ? bn->sums : nullptr;
    ConvGeom g6 = g; g6.Ci = 6 * g.Ci;
    const bool tc_split = split_tc() && tc_conv_supported(g6, false);
    const bool tc = tc_split || (cfg_.dt == DT_BF16 && cfg_.use_tc && tc_conv_supported(g, false));
    if (tc_split) {
        // fp32 parity mode on tcgen05: the bf16 kernel over the split-bf16 copy of the input (6x the channels), fp32 output
        k_split6_act(st, (const float*)in.p, split_scratch_, (size_t)g.B * g.Hi * g.Wi, g.Ci);
        k_conv_tc(st, split_scratch_, g.B, g.Hi, g.Wi, g6.Ci, c.wp6, g.Co, g.R, g.S, g.stride, g.pad, out.p, g.Ho, g.Wo, bias, stats, false, true,
                  nullptr, g.Ci);
    } else if (tc)
        k_conv_tc(st, in.p, g.B, g.Hi, g.Wi, g.Ci, c.wp, g.Co, g.R, g.S, g.stride, g.pad, out.p, g.Ho, g.Wo, bias, stats, false);
    else
        k_conv_fwd_simt(st, cfg_.dt, in.p, c.wp, bias, out.p, stats, g);
    prof_end(st);
    if (bn) {
        if (train) k_bn_finalize_train(st, bn_ref(*bn), tc ? SALT_STAT_SLOTS_CONV : SALT_STAT_SLOTS, (double)B_ * out.H * out.W, BN_MOMENTUM, BN_EPS);
        else k_bn_finalize_eval(st, bn_ref(*bn), BN_EPS);
    }
}
std::vector<BNLayer*> Engine::all_bns() {
    std::vector<BNLayer*> v;
    v.push_back(&stem_bn_);
    for (auto& b : blocks_) { v.push_back(&b->b1); v.push_back(&b->b2); if (b->down) v.push_back(&b->bd); }
    for (auto& b : bnecks_) { v.push_back(&b->b1); v.push_back(&b->b2); v.push_back(&b->b3); if (b->down) v.push_back(&b->bd); }
    v.push_back(&center0_.bn); v.push_back(&center1_.bn);
    for (auto& d : dec_) { v.push_back(&d.u1.bn); v.push_back(&d.u2.bn); }
    v.push_back(&final0_.bn);
    return v;
}
// eval-mode coefficients of every BatchNorm layer; they only change with the parameters / running statistics
void Engine::finalize_eval_all(cudaStream_t st) {
    for (BNLayer* b : all_bns()) k_bn_finalize_eval(st, bn_ref(*b), BN_EPS);
    eval_coef_dirty_ = false;
}
bool Engine::fusable(const ConvLayer& c, const Tensor& in, const Tensor& out) const {
    if (!fuse_eval_ || !cfg_.use_tc) return false;
    ConvGeom g = geom(c, in, out);
    if (cfg_.dt == DT_F32) g.Ci *= 6;
    return tc_conv_supported(g, false);
}
void Engine::conv_bn_fused(const ConvLayer& c, const Tensor& in, const Tensor& dst, const BNLayer& bn, const Tensor* res, bool relu,
                           cudaStream_t st) {
    ConvGeom g = geom(c, in, dst);
    const float* bias = c.o_b >= 0 ? params_ + c.o_b : nullptr;
    EpiParams ep;
    ep.scale = bn.scale; ep.shift = bn.shift; ep.res = res ? res->p : nullptr; ep.relu = relu ? 1 : 0;
    ep.Hp = dst.Hp(); ep.Wp = dst.Wp(); ep.pt = dst.pt; ep.pl = dst.pl;
    prof_begin(PROF_CONV_FWD, conv_flops(g, c.Ci_real / c.groups), st);
    if (cfg_.dt == DT_F32) {
        k_split6_act(st, (const float*)in.p, split_scratch_, (size_t)g.B * g.Hi * g.Wi, g.Ci);
        k_conv_tc(st, split_scratch_, g.B, g.Hi, g.Wi, 6 * g.Ci, c.wp6, g.Co, g.R, g.S, g.stride, g.pad, dst.p, g.Ho, g.Wo, bias, nullptr,
                  false, true, &ep, g.Ci);
    } else {
        k_conv_tc(st, in.p, g.B, g.Hi, g.Wi, g.Ci, c.wp, g.Co, g.R, g.S, g.stride, g.pad, dst.p, g.Ho, g.Wo, bias, nullptr, false, false, &ep);
    }
    prof_end(st);
}
void Engine::conv_dgrad(const ConvLayer& c, const Tensor& gout, const Tensor& gin, bool accumulate, cudaStream_t st) {
    ConvGeom g = geom(c, gin, gout);
    prof_begin(PROF_CONV_DGRAD, conv_flops(g, c.Ci_real / c.groups), st);
    const bool tc_ok = cfg_.dt == DT_BF16 && cfg_.use_tc && tc_conv_supported(g, true);
    if (tc_ok && g.stride == 1)
        // dgrad = correlation of the output gradient with the tap-flipped, transposed weights, zero padding R-1-pad
        k_conv_tc(st, gout.p, g.B, g.Ho, g.Wo, g.Co, c.wpd, g.Ci, g.R, g.S, 1, g.R - 1 - g.pad, gin.p, g.Hi, g.Wi, nullptr, nullptr, accumulate);
    else if (tc_ok && g.stride == 2 && (accumulate || g.R > 1))
        k_conv_tc_dgrad_s2(st, gout.p, g.B, g.Ho, g.Wo, g.Co, c.wpd, g.Ci, g.R, g.S, g.pad, gin.p, g.Hi, g.Wi, accumulate);
    else
        k_conv_dgrad_simt(st, cfg_.dt, gout.p, c.wpd, gin.p, accumulate, g);
    prof_end(st);
}
void Engine::conv_wgrad(ConvLayer& c, const Tensor& in, const Tensor& gout, cudaStream_t st) {
    ConvGeom g = geom(c, in, gout);
    prof_begin(PROF_CONV_WGRAD, conv_flops(g, c.Ci_real / c.groups), st);
    if (cfg_.dt == DT_BF16 && cfg_.use_tc && tc_wgrad_supported(g)) {
        k_conv_wgrad_tc(st, in.p, gout.p, c.dwp, g);      // partial sums land in c.dwp; unpack_all() transposes them at the end of backward()
        if (!c.in_unpack_table) { c.in_unpack_table = true; unpack_table_dirty_ = true; }
    } else
        k_conv_wgrad_simt(st, cfg_.dt, in.p, gout.p, grads_ + c.o_w, c.Ci_real, g, c.groups);
    prof_end(st);
}
void Engine::gather_fwd(const std::vector<Source>& srcs, const Tensor& P, cudaStream_t st) {
    GatherSrc gs[5];
    for (size_t i = 0; i < srcs.size(); ++i) {
        const Tensor& t = srcs[i].a->t;
        gs[i].p = t.p; gs[i].H = t.H; gs[i].W = t.W; gs[i].C = t.C; gs[i].f = srcs[i].f;
    }
    k_gather_fwd(st, P, gs, (int)srcs.size());
}
void Engine::gather_bwd(const std::vector<Source>& srcs, const Tensor& gP, cudaStream_t st) {
    int c0 = 0;
    for (auto& s : srcs) {
        GradBuf* gb = s.a->gb;
        Tensor g = view(gb->g);
        if (s.f == 1) k_fold_bwd(st, gP, c0, g, !gb->fresh);
        else k_upsample_bwd(st, gP, c0, s.f, g, (float*)scratch_[3], !gb->fresh);
        gb->fresh = false;
        c0 += s.a->t.C;
    }
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
void Engine::block_fwd(BasicBlock& b, bool train, cudaStream_t st) {
    Tensor x = view(b.x->t), raw1 = view(b.raw1), a1 = view(b.a1), raw2 = view(b.raw2), out = view(b.out.t);
    if (!train && fusable(b.c1, x, a1) && fusable(b.c2, a1, out) && (!b.down || fusable(b.cd, x, out))) {
        // inference: 2 (3 with a projection shortcut) kernels per block, no BatchNorm / ReLU / residual passes
        conv_bn_fused(b.c1, x, a1, b.b1, nullptr, true, st);
        if (b.down) {
            Tensor rawd = view(b.rawd);
            conv_bn_fused(b.cd, x, rawd, b.bd, nullptr, false, st);
            conv_bn_fused(b.c2, a1, out, b.b2, &rawd, true, st);
        } else {
            conv_bn_fused(b.c2, a1, out, b.b2, &x, true, st);
        }
        return;
    }
    conv_fwd(b.c1, x, raw1, &b.b1, train, st);
    k_bn_apply(st, raw1, b.b1.scale, b.b1.shift, nullptr, nullptr, nullptr, true, a1);
    conv_fwd(b.c2, a1, raw2, &b.b2, train, st);
    if (b.down) {
        Tensor rawd = view(b.rawd);
        conv_fwd(b.cd, x, rawd, &b.bd, train, st);
        k_bn_apply(st, raw2, b.b2.scale, b.b2.shift, &rawd, b.bd.scale, b.bd.shift, true, out);
    } else {
        k_bn_apply(st, raw2, b.b2.scale, b.b2.shift, &x, nullptr, nullptr, true, out);
    }
}
void Engine::bneck_fwd(Bottleneck& b, bool train, cudaStream_t st) {
    Tensor x = view(b.x->t), raw1 = view(b.raw1), a1 = view(b.a1), raw2 = view(b.raw2), a2 = view(b.a2), raw3 = view(b.raw3),
           out = view(b.out.t);
    if (!train && fusable(b.c1, x, a1) && fusable(b.c2, a1, a2) && fusable(b.c3, a2, raw3) && (!b.down || fusable(b.cd, x, raw3))) {
        // inference: BatchNorm (+ ReLU) folded into the convolutions; the SE gate needs the pooled bn3 output, so the gated
        // residual sum stays one separate pass over already-normalised tensors (identity coefficients)
        conv_bn_fused(b.c1, x, a1, b.b1, nullptr, true, st);
        conv_bn_fused(b.c2, a1, a2, b.b2, nullptr, true, st);
        conv_bn_fused(b.c3, a2, raw3, b.b3, nullptr, false, st);
        k_se_gate_fwd(st, raw3, ones_, zeros_, se_ref(b.se));
        if (b.down) {
            Tensor rawd = view(b.rawd);
            conv_bn_fused(b.cd, x, rawd, b.bd, nullptr, false, st);
            k_bn_apply(st, raw3, ones_, zeros_, &rawd, nullptr, nullptr, true, out, b.se.cse);
        } else {
            k_bn_apply(st, raw3, ones_, zeros_, &x, nullptr, nullptr, true, out, b.se.cse);
        }
        return;
    }
    conv_fwd(b.c1, x, raw1, &b.b1, train, st);
    k_bn_apply(st, raw1, b.b1.scale, b.b1.shift, nullptr, nullptr, nullptr, true, a1);
    conv_fwd(b.c2, a1, raw2, &b.b2, train, st);
    k_bn_apply(st, raw2, b.b2.scale, b.b2.shift, nullptr, nullptr, nullptr, true, a2);
    conv_fwd(b.c3, a2, raw3, &b.b3, train, st);
    k_se_gate_fwd(st, raw3, b.b3.scale, b.b3.shift, se_ref(b.se));
    if (b.down) {
        Tensor rawd = view(b.rawd);
        conv_fwd(b.cd, x, rawd, &b.bd, train, st);
        k_bn_apply(st, raw3, b.b3.scale, b.b3.shift, &rawd, b.bd.scale, b.bd.shift, true, out, b.se.cse);
    } else {
        k_bn_apply(st, raw3, b.b3.scale, b.b3.shift, &x, nullptr, nullptr, true, out, b.se.cse);
    }
}
void Engine::cbr_fwd(ConvBnRelu& u, bool train, cudaStream_t st) {
    conv_fwd(u.c, view(u.P), view(u.raw), &u.bn, train, st);
}
void Engine::decoder_fwd(DecoderBlock& d, bool train, cudaStream_t st) {
    gather_fwd(d.srcs, view(d.u1.P), st);
    if (!train && fusable(d.u1.c, view(d.u1.P), view(d.u1.raw)) && fusable(d.u2.c, view(d.u2.P), view(d.u2.raw))) {
        // inference: conv1 writes relu(bn(.)) straight into conv2's replicate-bordered input, conv2 leaves z = relu(bn(.)) for scSE
        conv_bn_fused(d.u1.c, view(d.u1.P), view(d.u2.P), d.u1.bn, nullptr, true, st);
        conv_bn_fused(d.u2.c, view(d.u2.P), view(d.u2.raw), d.u2.bn, nullptr, true, st);
        k_scse_fwd(st, view(d.u2.raw), ones_, zeros_, se_ref(d.se), view(d.out.t));
        return;
    }
    cbr_fwd(d.u1, train, st);
    k_bn_apply(st, view(d.u1.raw), d.u1.bn.scale, d.u1.bn.shift, nullptr, nullptr, nullptr, true, view(d.u2.P));
    cbr_fwd(d.u2, train, st);
    k_scse_fwd(st, view(d.u2.raw), d.u2.bn.scale, d.u2.bn.shift, se_ref(d.se), view(d.out.t));
}
void Engine::forward(const float* x_nchw, int B, float* logits_nchw, bool train, cudaStream_t st) {
    if (!params_) throw std::runtime_error("engine not bound");
    if (B < 1 || B > cfg_.max_batch) throw std::runtime_error("batch size exceeds the engine's max_batch");
    B_ = B;
    k_stem_im2col(st, cfg_.dt, x_nchw, x4_.p, B, cfg_.H, cfg_.W);
    forward_body(B, logits_nchw, train, st);
}
void Engine::forward_tiles(const uint8_t* tiles, int B, const TileGeom& g, float* logits_nchw, bool train, cudaStream_t st) {
    if (!params_) throw std::runtime_error("engine not bound");
    if (B < 1 || B > cfg_.max_batch) throw std::runtime_error("batch size exceeds the engine's max_batch");
    if (cfg_.H != cfg_.W || g.S != cfg_.H) throw std::runtime_error("tile adapter: network input must be square and equal to the padded size");
    if (g.th < 1 || g.tw < 1 || g.th > g.S || g.tw > g.S) throw std::runtime_error("tile adapter: tile larger than the network input");
    B_ = B;
    k_stem_im2col_tiles(st, cfg_.dt, tiles, x4_.p, B, g);
    forward_body(B, logits_nchw, train, st);
}
void Engine::forward_body(int B, float* logits_nchw, bool train, cudaStream_t st) {
    nvtxRangePushA(train ? "salt.forward(train)" : "salt.forward(eval)");
    if (packed_dirty_) pack_all(st);
    if (train) { k_zero(st, stats_arena_, sizeof(float) * stats_floats_); eval_coef_dirty_ = true; }
    else if (eval_coef_dirty_ && fuse_eval_) finalize_eval_all(st);
    Tensor x4 = view(x4_), sraw = view(stem_raw_);
    set_group(0);
    if (!train && fusable(stem_, x4, sraw)) {
        conv_bn_fused(stem_, x4, view(stem_out_.t), stem_bn_, nullptr, true, st);
    } else {
        conv_fwd(stem_, x4, sraw, &stem_bn_, train, st);
        k_bn_apply(st, sraw, stem_bn_.scale, stem_bn_.shift, nullptr, nullptr, nullptr, true, view(stem_out_.t));
    }
    for (auto& b : blocks_) { set_group(b->group); block_fwd(*b, train, st); }
    for (auto& b : bnecks_) { set_group(b->group); bneck_fwd(*b, train, st); }
    // center
    set_group(5);
    gather_fwd(center_src_, view(center0_.P), st);
    if (!train && fusable(center0_.c, view(center0_.P), view(center0_.raw)) && fusable(center1_.c, view(center1_.P), view(center1_.raw))) {
        conv_bn_fused(center0_.c, view(center0_.P), view(center1_.P), center0_.bn, nullptr, true, st);
        conv_bn_fused(center1_.c, view(center1_.P), view(center1_.raw), center1_.bn, nullptr, true, st);
        k_bn_relu_avgpool(st, view(center1_.raw), ones_, zeros_, view(center_out_.t));
    } else {
        cbr_fwd(center0_, train, st);
        k_bn_apply(st, view(center0_.raw), center0_.bn.scale, center0_.bn.shift, nullptr, nullptr, nullptr, true, view(center1_.P));
        cbr_fwd(center1_, train, st);
        k_bn_relu_avgpool(st, view(center1_.raw), center1_.bn.scale, center1_.bn.shift, view(center_out_.t));
    }
    for (int i = 0; i < 5; ++i) { set_group(6 + i); decoder_fwd(dec_[i], train, st); }
    set_group(11);
    gather_fwd(final_src_, view(final0_.P), st);
    if (!train && fusable(final0_.c, view(final0_.P), view(final0_.raw))) {
        conv_bn_fused(final0_.c, view(final0_.P), view(final0_.raw), final0_.bn, nullptr, true, st);
        k_final_fwd(st, view(final0_.raw), ones_, zeros_, params_ + o_final_w_, params_ + o_final_b_, cfg_.num_classes, logits_nchw);
    } else {
        cbr_fwd(final0_, train, st);
        k_final_fwd(st, view(final0_.raw), final0_.bn.scale, final0_.bn.shift, params_ + o_final_w_, params_ + o_final_b_,
                    cfg_.num_classes, logits_nchw);
    }
    trained_forward_ = train;
    close_group();
    nvtxRangePop();
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
void Engine::decoder_bwd(DecoderBlock& d, cudaStream_t st) {
    const Tensor raw2 = view(d.u2.raw), raw1 = view(d.u1.raw), P2 = view(d.u2.P), P1 = view(d.u1.P);
    const double cnt = (double)B_ * raw2.H * raw2.W;
    BNRef bn2 = bn_ref(d.u2.bn), bn1 = bn_ref(d.u1.bn);
    Tensor gbn2 = scratch(0, raw2.H, raw2.W, raw2.C), graw2 = scratch(1, raw2.H, raw2.W, raw2.C);
    k_scse_bwd(st, view(d.out.gb->g), raw2, bn2, se_ref(d.se), gbn2);
    k_bn_bwd_finalize(st, bn2, cnt);
    k_bn_bwd_apply(st, gbn2, raw2, bn2, false, graw2);
    conv_wgrad(d.u2.c, P2, graw2, st);
    Tensor gP2 = scratch(2, P2.H, P2.W, P2.C, P2.pt, P2.pb, P2.pl, P2.pr);
    conv_dgrad(d.u2.c, graw2, gP2, false, st);
    Tensor ga1 = scratch(0, raw1.H, raw1.W, raw1.C), graw1 = scratch(1, raw1.H, raw1.W, raw1.C);
    k_fold_bwd(st, gP2, 0, ga1, false);
    k_bn_bwd_reduce(st, ga1, raw1, bn1, true);
    k_bn_bwd_finalize(st, bn1, cnt);
    k_bn_bwd_apply(st, ga1, raw1, bn1, true, graw1);
    conv_wgrad(d.u1.c, P1, graw1, st);
    Tensor gP1 = scratch(2, P1.H, P1.W, P1.C, P1.pt, P1.pb, P1.pl, P1.pr);
    conv_dgrad(d.u1.c, graw1, gP1, false, st);
    gather_bwd(d.srcs, gP1, st);
}
void Engine::block_bwd(BasicBlock& b, cudaStream_t st) {
    Tensor x = view(b.x->t), raw1 = view(b.raw1), a1 = view(b.a1), raw2 = view(b.raw2), out = view(b.out.t);
    Tensor G = view(b.out.gb->g);
    const double cnt = (double)B_ * out.H * out.W;
    BNRef bn1 = bn_ref(b.b1), bn2 = bn_ref(b.b2);
    k_relu_mask_inplace(st, G, out);                               // gradient through the block's final ReLU
    Tensor graw2 = scratch(0, out.H, out.W, out.C), ga1 = scratch(1, out.H, out.W, out.C);
    k_bn_bwd_reduce(st, G, raw2, bn2, false);
    k_bn_bwd_finalize(st, bn2, cnt);
    k_bn_bwd_apply(st, G, raw2, bn2, false, graw2);
    GradBuf* gxb = b.x->gb;
    Tensor Gx = view(gxb->g);
    Tensor grawd = scratch(2, out.H, out.W, out.C);
    if (b.down) {
        Tensor rawd = view(b.rawd);
        BNRef bnd = bn_ref(b.bd);
        k_bn_bwd_reduce(st, G, rawd, bnd, false);
        k_bn_bwd_finalize(st, bnd, cnt);
        k_bn_bwd_apply(st, G, rawd, bnd, false, grawd);
        conv_wgrad(b.cd, x, grawd, st);
    }
    conv_wgrad(b.c2, a1, graw2, st);
    conv_dgrad(b.c2, graw2, ga1, false, st);
    Tensor graw1 = scratch(0, out.H, out.W, out.C);
    k_bn_bwd_reduce(st, ga1, raw1, bn1, true);
    k_bn_bwd_finalize(st, bn1, cnt);
    k_bn_bwd_apply(st, ga1, raw1, bn1, true, graw1);
    conv_wgrad(b.c1, x, graw1, st);
    // identity blocks: Gx aliases G and already holds the skip-path gradient -> accumulate
    conv_dgrad(b.c1, graw1, Gx, b.down ? !gxb->fresh : true, st);
    gxb->fresh = false;
    // the 1x1 stride-2 shortcut only touches the even/even input positions: it accumulates after the 3x3 path wrote everything
    if (b.down) conv_dgrad(b.cd, grawd, Gx, true, st);
}
void Engine::bneck_bwd(Bottleneck& b, cudaStream_t st) {
    Tensor x = view(b.x->t), raw1 = view(b.raw1), a1 = view(b.a1), raw2 = view(b.raw2), a2 = view(b.a2), raw3 = view(b.raw3),
           out = view(b.out.t);
    Tensor G = view(b.out.gb->g);
    const double cnt = (double)B_ * out.H * out.W;
    BNRef bn1 = bn_ref(b.b1), bn2 = bn_ref(b.b2), bn3 = bn_ref(b.b3);
    SERef se = se_ref(b.se);
    k_relu_mask_inplace(st, G, out);                               // gradient through the block's final ReLU
    // out = u*cse + shortcut, u = bn3(raw3):  d/du = G*cse + (gap path, se.G);  shortcut gets G
    k_se_gate_bwd(st, G, raw3, bn3.scale, bn3.shift, se);
    Tensor graw3 = scratch(0, out.H, out.W, out.C);
    k_bn_bwd_reduce(st, G, raw3, bn3, false, se.cse, se.G);
    k_bn_bwd_finalize(st, bn3, cnt);
    k_bn_bwd_apply(st, G, raw3, bn3, false, graw3, se.cse, se.G);
    GradBuf* gxb = b.x->gb;
    Tensor Gx = view(gxb->g);
    Tensor grawd = scratch(2, out.H, out.W, out.C);
    if (b.down) {
        Tensor rawd = view(b.rawd);
        BNRef bnd = bn_ref(b.bd);
        k_bn_bwd_reduce(st, G, rawd, bnd, false);
        k_bn_bwd_finalize(st, bnd, cnt);
        k_bn_bwd_apply(st, G, rawd, bnd, false, grawd);
        conv_wgrad(b.cd, x, grawd, st);
    }
    conv_wgrad(b.c3, a2, graw3, st);
    Tensor ga2 = scratch(1, raw2.H, raw2.W, raw2.C);
    conv_dgrad(b.c3, graw3, ga2, false, st);
    Tensor graw2 = scratch(0, raw2.H, raw2.W, raw2.C);
    k_bn_bwd_reduce(st, ga2, raw2, bn2, true);
    k_bn_bwd_finalize(st, bn2, (double)B_ * raw2.H * raw2.W);
    k_bn_bwd_apply(st, ga2, raw2, bn2, true, graw2);
    conv_wgrad(b.c2, a1, graw2, st);
    Tensor ga1 = scratch(1, raw1.H, raw1.W, raw1.C);
    conv_dgrad(b.c2, graw2, ga1, false, st);
    Tensor graw1 = scratch(0, raw1.H, raw1.W, raw1.C);
    k_bn_bwd_reduce(st, ga1, raw1, bn1, true);
    k_bn_bwd_finalize(st, bn1, (double)B_ * raw1.H * raw1.W);      // SE-ResNeXt: conv1 runs at the block's INPUT resolution
    k_bn_bwd_apply(st, ga1, raw1, bn1, true, graw1);
    conv_wgrad(b.c1, x, graw1, st);
    // identity blocks: Gx aliases G and already holds the shortcut gradient -> accumulate.  A stride-2 1x1 convolution only
    // reaches the even/even input positions, so a fresh buffer is cleared first.
    bool acc = b.down ? !gxb->fresh : true;
    if (!acc && b.c1.stride == 2) { k_zero(st, Gx.p, Gx.bytes()); acc = true; }
    conv_dgrad(b.c1, graw1, Gx, acc, st);
    gxb->fresh = false;
    if (b.down) conv_dgrad(b.cd, grawd, Gx, true, st);
}
void Engine::backward(const float* dlogits, cudaStream_t st, int seg) {
    if (!grads_) throw std::runtime_error("engine bound without gradient buffers");
    if (!trained_forward_) throw std::runtime_error("backward() requires a preceding forward(train=1)");
    if (seg < -1 || seg > 2) throw std::runtime_error("backward: segment must be -1 (all), 0, 1 or 2");
    const bool all = seg < 0;
    nvtxRangePushA("salt.backward");
    if (all || seg == 0) {
    k_zero(st, grads_, sizeof(float) * n_params_);
    // (the backward BatchNorm partial-sum slots need no clearing: every producing block stores its slot and records the slot count)
    if (cfg_.dt == DT_BF16 && cfg_.use_tc) k_zero(st, dwp_arena_, sizeof(float) * dwp_floats_);
    for (auto& g : gradbufs_) g->fresh = true;
    // ---- final
    set_group(11);
    {
        const Tensor raw = view(final0_.raw), P = view(final0_.P);
        BNRef bn = bn_ref(final0_.bn);
        Tensor gbn = scratch(0, raw.H, raw.W, raw.C), graw = scratch(1, raw.H, raw.W, raw.C);
        k_final_bwd(st, dlogits, raw, bn, params_ + o_final_w_, cfg_.num_classes, grads_ + o_final_w_, grads_ + o_final_b_, gbn);
        k_bn_bwd_finalize(st, bn, (double)B_ * raw.H * raw.W);
        k_bn_bwd_apply(st, gbn, raw, bn, false, graw);
        conv_wgrad(final0_.c, P, graw, st);
        Tensor gP = scratch(2, P.H, P.W, P.C, P.pt, P.pb, P.pl, P.pr);
        conv_dgrad(final0_.c, graw, gP, false, st);
        gather_bwd(final_src_, gP, st);
    }
    for (int i = 4; i >= 0; --i) { set_group(6 + i); decoder_bwd(dec_[i], st); }
    // ---- center
    set_group(5);
    {
        const Tensor raw1 = view(center1_.raw), raw0 = view(center0_.raw), P1 = view(center1_.P), P0 = view(center0_.P);
        const double cnt = (double)B_ * raw1.H * raw1.W;
        BNRef bn1 = bn_ref(center1_.bn), bn0 = bn_ref(center0_.bn);
        Tensor gpost = scratch(0, raw1.H, raw1.W, raw1.C), graw1 = scratch(1, raw1.H, raw1.W, raw1.C);
        k_avgpool_bwd(st, view(center_out_.gb->g), gpost);
        k_bn_bwd_reduce(st, gpost, raw1, bn1, true);
        k_bn_bwd_finalize(st, bn1, cnt);
        k_bn_bwd_apply(st, gpost, raw1, bn1, true, graw1);
        conv_wgrad(center1_.c, P1, graw1, st);
        Tensor gP1 = scratch(2, P1.H, P1.W, P1.C, P1.pt, P1.pb, P1.pl, P1.pr);
        conv_dgrad(center1_.c, graw1, gP1, false, st);
        Tensor ga0 = scratch(0, raw0.H, raw0.W, raw0.C), graw0 = scratch(1, raw0.H, raw0.W, raw0.C);
        k_fold_bwd(st, gP1, 0, ga0, false);
        k_bn_bwd_reduce(st, ga0, raw0, bn0, true);
        k_bn_bwd_finalize(st, bn0, cnt);
        k_bn_bwd_apply(st, ga0, raw0, bn0, true, graw0);
        conv_wgrad(center0_.c, P0, graw0, st);
        Tensor gP0 = scratch(2, P0.H, P0.W, P0.C, P0.pt, P0.pb, P0.pl, P0.pr);
        conv_dgrad(center0_.c, graw0, gP0, false, st);
        gather_bwd(center_src_, gP0, st);
    }
    if (!all) unpack_all(st, 0);
    }
    // encoder blocks in reverse order; segment 1 = layer4 + layer3 (groups 4, 3), segment 2 = layer2 + layer1 (+ stem)
    for (int i = (int)blocks_.size() - 1; i >= 0; --i) {
        const int g = blocks_[i]->group;
        if (all || (seg == 1 && g >= 3) || (seg == 2 && g <= 2)) { set_group(g); block_bwd(*blocks_[i], st); }
    }
    for (int i = (int)bnecks_.size() - 1; i >= 0; --i) {
        const int g = bnecks_[i]->group;
        if (all || (seg == 1 && g >= 3) || (seg == 2 && g <= 2)) { set_group(g); bneck_bwd(*bnecks_[i], st); }
    }
    if (seg == 1) unpack_all(st, 1);
    if (all || seg == 2) {
    // ---- stem (no input gradient)
    set_group(0);
    {
        const Tensor raw = view(stem_raw_);
        BNRef bn = bn_ref(stem_bn_);
        Tensor G = view(stem_out_.gb->g), graw = scratch(1, raw.H, raw.W, raw.C);
        k_bn_bwd_reduce(st, G, raw, bn, true);
        k_bn_bwd_finalize(st, bn, (double)B_ * raw.H * raw.W);
        k_bn_bwd_apply(st, G, raw, bn, true, graw);
        conv_wgrad(stem_, view(x4_), graw, st);
    }
    unpack_all(st, all ? 3 : 2);
    }
    close_group();
    nvtxRangePop();
}
void Engine::adam(float lr, float wd, float b1, float b2, float eps, int step, float grad_scale, cudaStream_t st) {
    if (!adam_m_ || !adam_v_ || !grads_) throw std::runtime_error("engine bound without optimiser state");
    k_adam(st, params_, grads_, adam_m_, adam_v_, n_params_, lr, wd, b1, b2, eps, step, grad_scale);
    packed_dirty_ = true;
    eval_coef_dirty_ = true;
}

// ------------------------------------------------------------------------------------------------
// debugging / test access to internal activations
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void nhwc_to_nchw_f32_kernel(const T* __restrict__ src, float* __restrict__ dst, int B, int H, int W, int C, int pt,
                                        int pl, int Hp, int Wp) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)B * C * H * W;
    if (idx >= total) return;
    int x = (int)(idx % W), y = (int)((idx / W) % H), c = (int)((idx / ((long long)W * H)) % C), n = (int)(idx / ((long long)W * H * C));
    dst[idx] = ld1(src + (((size_t)n * Hp + y + pt) * Wp + x + pl) * C + c);
}
bool Engine::get_activation(const std::string& name, float* out, int* shape4, cudaStream_t st) {
    const Tensor* t = nullptr;
    auto dec_idx = [&](const std::string& n) { return n == "d5" ? 0 : n == "d4" ? 1 : n == "d3" ? 2 : n == "d2" ? 3 : n == "d1" ? 4 : -1; };
    if (name == "stem") t = &stem_out_.t;
    else if (name == "e2") t = &enc_out_[0]->t;
    else if (name == "e3") t = &enc_out_[1]->t;
    else if (name == "e4") t = &enc_out_[2]->t;
    else if (name == "e5") t = &enc_out_[3]->t;
    else if (name == "center") t = &center_out_.t;
    else if (dec_idx(name) >= 0) t = &dec_[dec_idx(name)].out.t;
    else if (name == "final_raw") t = &final0_.raw;
    else if (name == "final_in") t = &final0_.P;
    else if (name.rfind("g_", 0) == 0) {           // gradient buffers: g_e2.., g_d1.., g_center, g_stem
        std::string n = name.substr(2);
        if (n == "stem") t = &stem_out_.gb->g;
        else if (n == "e2") t = &enc_out_[0]->gb->g;
        else if (n == "e3") t = &enc_out_[1]->gb->g;
        else if (n == "e4") t = &enc_out_[2]->gb->g;
        else if (n == "e5") t = &enc_out_[3]->gb->g;
        else if (n == "center") t = &center_out_.gb->g;
        else if (dec_idx(n) >= 0) t = &dec_[dec_idx(n)].out.gb->g;
    }
    if (!t) return false;
    Tensor v = view(*t);
    shape4[0] = v.B; shape4[1] = v.C; shape4[2] = v.H; shape4[3] = v.W;
    if (out) {
        long long total = (long long)v.B * v.C * v.H * v.W;
        SALT_DISPATCH(v.dt, T, (nhwc_to_nchw_f32_kernel<T><<<cdiv(total, 256), 256, 0, st>>>((const T*)v.p, out, v.B, v.H, v.W, v.C,
                                                                                          v.pt, v.pl, v.Hp(), v.Wp())));
    }
    return true;
}

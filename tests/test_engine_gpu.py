"""GPU parity tests: the CUDA engine (through the C ABI) against the CPU oracle and the golden fixtures.

Tolerances (stated per test):
  fp32 precision mode  - forward logits <= 1e-3 max-abs (BASELINE.json north_star), gradients <= 2e-3 relative
  bf16 precision mode  - looser, stated bounds + mask IoU
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import synth, unet_oracle, losses_oracle            # noqa: E402
from oracle.make_golden import sample, GRAD_KEYS, grad_keys, ARCH_NAMES     # noqa: E402


def _engine(*a, **k):
    from salt_b200.engine import UNetEngine
    return UNetEngine(*a, **k)


def report(name, got, ref, atol=0.0, rtol=0.0, l2rel=None, outlier_frac=0.0):
    """max-abs / relative-L2 comparison.  outlier_frac > 0 (gradient checks): a ReLU whose pre-activation differs from the
    oracle's by one fp32 ulp around 0 flips its mask, which changes the gradient of a handful of isolated elements by their full
    magnitude - so up to that fraction of elements may exceed the element-wise bound, the relative-L2 bound still applies."""
    got = torch.as_tensor(got).detach().float().cpu()
    ref = torch.as_tensor(ref).detach().float().cpu()
    assert got.shape == ref.shape, '%s: shape %s vs %s' % (name, tuple(got.shape), tuple(ref.shape))
    assert torch.isfinite(got).all(), '%s: non-finite values' % name
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    l2 = ((got - ref).double().norm() / (ref.double().norm() + 1e-30)).item()
    ok = err <= atol + rtol * scale
    if not ok and outlier_frac > 0:
        bad = ((got - ref).abs() > atol + rtol * scale).float().mean().item()
        ok = bad <= outlier_frac
    if l2rel is not None:
        ok = ok and l2 <= l2rel
    print('%-60s max-abs-err %.3e  ref-max %.3e  rel %.3e  l2rel %.3e  %s' % (name, err, scale, err / (scale + 1e-30), l2, 'ok' if ok else 'FAIL'))
    return ok, err


# --------------------------------------------------------------------------------------------- single ops
def _to_nhwc(t, prec):
    t = t.permute(0, 2, 3, 1).contiguous().cuda()
    return t.to(torch.bfloat16) if prec == 'bf16' else t


def _from_nhwc(t):
    return t.float().permute(0, 3, 1, 2).contiguous().cpu()


CONV_CASES = [
    # (B, Cin, Cout, H, W, k, stride, pad)
    (2, 64, 64, 16, 16, 3, 1, 1),
    (3, 64, 128, 16, 16, 3, 2, 1),
    (2, 64, 128, 16, 16, 1, 2, 0),
    (2, 4, 64, 32, 32, 7, 2, 3),
    (2, 96, 32, 10, 12, 3, 1, 0),       # valid conv on a bordered tensor, N=32, ragged sizes
    (1, 512, 64, 4, 4, 3, 1, 1),
]


@pytest.mark.parametrize('prec', ['fp32', 'bf16'])
@pytest.mark.parametrize('case', CONV_CASES)
def test_conv_ops(case, prec):
    from salt_b200 import _lib
    lib = _lib.load()
    B, Ci, Co, H, W, k, s, p = case
    g = torch.Generator().manual_seed(hash(case) % 1000)
    x = torch.randn(B, Ci, H, W, generator=g)
    w = torch.randn(Co, Ci, k, k, generator=g) * (2.0 / (Ci * k * k)) ** 0.5
    bias = torch.randn(Co, generator=g) * 0.1
    if prec == 'bf16':
        x, w = x.bfloat16().float(), w.bfloat16().float()
    y_ref = F.conv2d(x, w, bias, stride=s, padding=p)
    Ho, Wo = y_ref.shape[2:]
    d = _lib.SaltConvDesc(B, H, W, Ci, Ho, Wo, Co, k, s, p, 0 if prec == 'fp32' else 1, 0)
    tol = dict(atol=1e-4, rtol=1e-5) if prec == 'fp32' else dict(atol=2e-2, rtol=1e-2)
    xd, wd, bd = _to_nhwc(x, prec), w.cuda().contiguous(), bias.cuda()
    out = torch.empty((B, Ho, Wo, Co), dtype=xd.dtype, device='cuda')
    stats = torch.zeros(2 * Co, dtype=torch.float64, device='cuda')
    _lib.check(lib.salt_op_conv_forward(C.byref(d), xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), out.data_ptr(), stats.data_ptr(), None))
    torch.cuda.synchronize()
    ok1, _ = report('conv fwd %s %s' % (case, prec), _from_nhwc(out), y_ref, **tol)
    ok2, _ = report('conv stats sum', stats[:Co].cpu().float(), y_ref.sum((0, 2, 3)), atol=1e-2 if prec == 'fp32' else 1.0, rtol=1e-3)
    ok3, _ = report('conv stats sumsq', stats[Co:].cpu().float(), (y_ref ** 2).sum((0, 2, 3)), atol=1e-2, rtol=1e-3 if prec == 'fp32' else 2e-2)
    # dgrad / wgrad against autograd
    gy = torch.randn(B, Co, Ho, Wo, generator=g)
    if prec == 'bf16':
        gy = gy.bfloat16().float()
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    F.conv2d(xr, wr, None, stride=s, padding=p).backward(gy)
    gyd = _to_nhwc(gy, prec)
    gin = torch.full((B, H, W, Ci), 7.0, dtype=xd.dtype, device='cuda')
    _lib.check(lib.salt_op_conv_dgrad(C.byref(d), gyd.data_ptr(), wd.data_ptr(), gin.data_ptr(), 0, None))
    ok4, _ = report('conv dgrad', _from_nhwc(gin), xr.grad, **(dict(atol=1e-4, rtol=1e-5) if prec == 'fp32' else dict(atol=5e-2, rtol=1e-2)))
    _lib.check(lib.salt_op_conv_dgrad(C.byref(d), gyd.data_ptr(), wd.data_ptr(), gin.data_ptr(), 1, None))
    ok5, _ = report('conv dgrad accumulate', _from_nhwc(gin), 2 * xr.grad, **(dict(atol=2e-4, rtol=1e-5) if prec == 'fp32' else dict(atol=1e-1, rtol=2e-2)))
    dw = torch.zeros_like(wd)
    _lib.check(lib.salt_op_conv_wgrad(C.byref(d), xd.data_ptr(), gyd.data_ptr(), dw.data_ptr(), None))
    torch.cuda.synchronize()
    ok6, _ = report('conv wgrad', dw.cpu(), wr.grad, atol=1e-3, rtol=1e-4 if prec == 'fp32' else 1e-2)
    assert ok1 and ok2 and ok3 and ok4 and ok5 and ok6


def test_adam_matches_torch():
    from salt_b200 import _lib
    lib = _lib.load()
    n = 1000 + 3
    g = torch.Generator().manual_seed(3)
    p0 = torch.randn(n, generator=g)
    p_ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([{'params': [p_ref], 'weight_decay': 1e-4}], lr=1e-3)
    p = p0.cuda().clone()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 4):
        grad = torch.randn(n, generator=g)
        p_ref.grad = grad.clone()
        opt.step()
        gd = grad.cuda()
        _lib.check(lib.salt_op_adam(p.data_ptr(), gd.data_ptr(), m.data_ptr(), v.data_ptr(), n, 1e-3, 1e-4, 0.9, 0.999, 1e-8, step, 1.0, None))
    torch.cuda.synchronize()
    ok, _ = report('adam 3 steps', p.cpu(), p_ref.detach(), atol=1e-6, rtol=1e-6)
    assert ok


# --------------------------------------------------------------------------------------------- whole network
STAGES = ['e2', 'e3', 'e4', 'e5', 'center', 'd5', 'd4', 'd3', 'd2', 'd1']


def _setup(depth, b, s, wseed=0, dseed=1234, arch=None):
    sd_np = synth.synth_state_dict(depth, 2, wseed, arch)
    x = torch.from_numpy(synth.synth_inputs(b, s, dseed))
    t = torch.from_numpy(synth.synth_targets(b, s, dseed))
    return sd_np, x, t


@pytest.mark.parametrize('tc', [True, False], ids=['tcgen05-split-bf16', 'simt-fp32'])
@pytest.mark.parametrize('depth,b,s', [(18, 2, 64), (34, 3, 64), (18, 8, 128)])
def test_forward_eval_fp32(depth, b, s, tc):
    """fp32 mode, eval BatchNorm: every stage and the logits vs the oracle; logits <= 1e-3 max-abs - with the convolutions on the
    tensor cores (split-bf16 operands, kernels.h) and with the fp32-FMA implicit GEMM."""
    sd_np, x, _ = _setup(depth, b, s)
    with torch.no_grad():
        ref, stages = unet_oracle.unet_resnet_forward(unet_oracle.to_torch_state(sd_np), x, depth, False, return_stages=True)
    eng = _engine(depth, 2, b, s, precision='fp32', training=False, use_tensor_cores=tc)
    eng.load_state(sd_np)
    logits = eng.forward(x.cuda(), train=False)
    torch.cuda.synchronize()
    oks = []
    for name in STAGES:
        oks.append(report('eval %s' % name, eng.activation(name), stages[name], atol=1e-3, rtol=1e-4)[0])
    ok, err = report('eval logits', logits, ref, atol=1e-3)
    assert ok and all(oks)


@pytest.mark.parametrize('loss_name', ['lovasz', 'bcedice'])
@pytest.mark.parametrize('depth,b,s', [(18, 2, 64), (34, 2, 64)])
def test_train_step_fp32(depth, b, s, loss_name):
    """fp32 mode, training step: logits, loss, dL/dlogits, every parameter gradient, BN running stats and the
    Adam update vs the oracle (autograd on the restated network)."""
    sd_np, x, t = _setup(depth, b, s)
    sd = unet_oracle.to_torch_state(sd_np, requires_grad=True)
    ref, stages = unet_oracle.unet_resnet_forward(sd, x, depth, True, return_stages=True)
    ref.retain_grad()
    for v in stages.values():
        v.retain_grad()
    loss_ref = (losses_oracle.lovasz_hinge_per_image if loss_name == 'lovasz' else losses_oracle.bce_dice)(ref, t)
    loss_ref.backward()

    # every gradient is compared element-wise at fp32-rounding tolerances: that is a property of the fp32-FMA kernels
    # (use_tensor_cores=False); the tensor-core forward of the fp32 mode is checked in test_golden_fixtures_fp32 / test_forward_eval_fp32
    eng = _engine(depth, 2, b, s, precision='fp32', use_tensor_cores=False)
    eng.load_state(sd_np)
    logits = eng.forward(x.cuda(), train=True)
    loss, dlogits = (eng.loss_lovasz if loss_name == 'lovasz' else eng.loss_bce_dice)(logits, t.cuda())
    eng.backward(dlogits)
    torch.cuda.synchronize()
    oks = [report('train logits', logits, ref, atol=1e-3)[0]]
    oks.append(report('loss %s' % loss_name, loss.cpu()[0], loss_ref, atol=1e-5, rtol=1e-4)[0])
    oks.append(report('dlogits', dlogits, ref.grad, atol=1e-9, rtol=2e-3)[0])
    for name in ['d1', 'd2', 'd3', 'd4', 'd5', 'center']:
        oks.append(report('grad act %s' % name, eng.activation('g_' + name), stages[name].grad, atol=1e-9, rtol=2e-2, l2rel=2e-2, outlier_frac=1e-3)[0])
    bad = []
    for k, (shape, off, numel, isbuf) in eng.table.items():
        if isbuf:
            ok, _ = report('buffer %s' % k, eng.view(k), sd[k], atol=1e-5, rtol=1e-4)
        else:
            gref = sd[k].grad
            # conv biases in front of a train-mode BatchNorm have an exactly-zero gradient; autograd yields rounding noise
            if k.endswith('.conv.bias'):
                ok = gref.abs().max().item() < 1e-5 and eng.view(k, grad=True).abs().max().item() < 1e-5
            else:
                # single-element gradients are sums with heavy cancellation: absolute tolerance instead of relative L2
                ok, _ = report('grad %s' % k, eng.view(k, grad=True), gref, atol=1e-5 if numel == 1 else 1e-7, rtol=2e-2,
                               l2rel=None if numel == 1 else 5e-3, outlier_frac=1e-3)
        if not ok:
            bad.append(k)
    assert all(oks) and not bad, bad
    # Adam + L2 update of every parameter.  The first Adam step is ~ lr*sign(g): it is compared on the engine's own
    # gradients (checked above) so that rounding noise on near-zero gradients cannot flip an update.
    params = {k: v.detach().clone() for k, v in sd.items() if v.requires_grad}
    grads = {k: eng.view(k, grad=True).cpu().clone() for k in params}
    m = {k: torch.zeros_like(v) for k, v in params.items()}
    vv = {k: torch.zeros_like(v) for k, v in params.items()}
    unet_oracle.adam_l2_step(params, grads, m, vv, 1, lr=1e-4, wd=1e-4)
    eng.adam_step(lr=1e-4, weight_decay=1e-4)
    torch.cuda.synchronize()
    worst = 0.0
    for k in params:
        worst = max(worst, (eng.view(k).cpu() - params[k]).abs().max().item())
    print('adam: worst parameter deviation after one step %.3e (lr 1e-4)' % worst)
    assert worst <= 1e-6


@pytest.mark.parametrize('tag,tc', [('r18_b2_s64', True), ('r34_b2_s64', True), ('r18_b8_s128', True), ('se50_b2_s64', True),
                                    ('se101_b2_s64', True), ('r18_b8_s128', False), ('r34_b2_s64', False),
                                    ('sex50_b2_s64', True), ('sex50_b2_s64', False)])
def test_golden_fixtures_fp32(golden_dir, tag, tc):
    """Engine vs vectors produced by the unmodified reference modules (tests/golden, oracle/make_golden.py).  tc: forward
    convolutions on tcgen05 with split-bf16 operands (the default fp32 mode) or the fp32-FMA kernels (use_tensor_cores=False)."""
    g = np.load(os.path.join(golden_dir, tag + '.npz'))
    m = {k[5:]: int(g[k]) for k in g.files if k.startswith('meta_')}
    arch = ARCH_NAMES[m.get('arch', 0)]              # 'UNetSeResNetXt' for the sex50 fixture (grouped convolutions run densely)
    sd_np, x, t = _setup(m['depth'], m['batch'], m['size'], m['wseed'], m['dseed'], arch)
    eng = _engine(m['depth'], 2, m['batch'], m['size'], precision='fp32', use_tensor_cores=tc, architecture=arch)
    eng.load_state(sd_np)
    xd, td = x.cuda(), t.cuda()
    oks = [report('golden eval logits', eng.forward(xd, train=False), g['logits_eval'], atol=1e-3)[0]]
    if 'tta_masks' in g.files:
        lf = eng.forward(torch.flip(xd, dims=[3]).contiguous(), train=False)
        lo = eng.forward(xd, train=False)
        probs, mask = eng.predict(lo, lf, crop=101, threshold=0.5)
        torch.cuda.synchronize()
        oks.append(report('golden tta probs', sample(probs.cpu().numpy()), g['tta_probs'], atol=1e-4)[0])
        iou = losses_oracle.iou_masks(mask.cpu().numpy(), g['tta_masks'])
        print('golden tta mask IoU vs reference: %.6f' % iou)
        assert iou >= 1 - 1e-4
    for loss_name in ('lovasz', 'bcedice'):
        eng.load_state(sd_np)
        logits = eng.forward(xd, train=True)
        loss, dlogits = (eng.loss_lovasz if loss_name == 'lovasz' else eng.loss_bce_dice)(logits, td)
        eng.backward(dlogits)
        torch.cuda.synchronize()
        oks.append(report('golden train logits', logits, g['logits_train'], atol=1e-3)[0])
        oks.append(report('golden loss ' + loss_name, loss.cpu()[0], float(g['loss_' + loss_name]), atol=1e-5, rtol=1e-4)[0])
        oks.append(report('golden dlogits ' + loss_name, sample(dlogits.cpu().numpy()), g['dlogits_' + loss_name], atol=1e-9, rtol=2e-3)[0])
        for k in grad_keys(m['depth']):
            if k.endswith('.conv.bias'):
                continue
            # a 2-image batch through ~100 train-mode BatchNorm layers amplifies fp32 rounding (and flips isolated ReLU masks): the
            # 101-layer encoder gets 3x the element-wise bound of the 18/34/50-layer ones plus a relative-L2 bound of 1e-2 (measured 1-5e-3)
            # ... and so does the tensor-core forward of the fp32 mode (split-bf16 operands: ~5e-6 relative per convolution instead of
            # ~1e-7), whose activations feed the fp32-FMA backward
            deep = m['depth'] >= 101 or tc
            ref_g = g['grad_%s_%s' % (loss_name, k)]
            # (conv biases in front of a train-mode BatchNorm have an exactly zero gradient: the reference holds ~1e-9 noise there)
            oks.append(report('golden grad %s' % k, sample(eng.view(k, grad=True).cpu().numpy()), ref_g, atol=1e-7,
                              rtol=1.5e-2 if deep else 5e-3, l2rel=1e-2 if deep and np.abs(ref_g).max() > 1e-6 else None,
                              outlier_frac=2e-3 if deep else 0.0)[0])
    assert all(oks)


@pytest.mark.parametrize('depth,b,s,arch', [(18, 8, 128, None), (34, 4, 128, None), (50, 4, 128, None), (50, 4, 128, 'UNetSeResNetXt')])
def test_bf16_mode(depth, b, s, arch):
    """bf16 precision mode (configs 2-5): bounded deviation from the fp32 oracle.  bf16 storage rounds every
    activation to 8 mantissa bits (2^-9 relative) ~50 times along the deepest path, so the stated tolerances are:
    logits max-abs <= 3 % of the logit range + 0.05, mean-abs <= 0.03; thresholded masks identical except where
    the oracle's logit is within that max-abs bound of 0; train-mode (small-batch BatchNorm) logits <= 8 % of the range
    + 0.05; Lovasz loss within 2 %; gradient cosine >= 0.90 (the stem, behind the longest bf16 chain, is the worst)."""
    sd_np, x, t = _setup(depth, b, s, arch=arch)
    sd = unet_oracle.to_torch_state(sd_np, requires_grad=True)
    with torch.no_grad():
        ref_eval = unet_oracle.unet_resnet_forward(sd, x, depth, False, arch=arch)
    ref = unet_oracle.unet_resnet_forward(sd, x, depth, True, arch=arch)
    loss_ref = losses_oracle.lovasz_hinge_per_image(ref, t)
    loss_ref.backward()
    eng = _engine(depth, 2, b, s, precision='bf16', architecture=arch)
    eng.load_state(sd_np)
    xd, td = x.cuda(), t.cuda()
    le = eng.forward(xd, train=False)
    _, mask = eng.predict(le, None, crop=min(101, s))
    torch.cuda.synchronize()
    bound = 0.05 + 0.03 * ref_eval.abs().max().item()
    ok, err = report('bf16 eval logits', le, ref_eval, atol=bound)
    mean_err = (le.cpu() - ref_eval).abs().mean().item()
    crop = min(101, s)
    _, mask_ref = losses_oracle.predict_masks(ref_eval.numpy(), None, crop, 0.5)
    iou = losses_oracle.iou_masks(mask.cpu().numpy(), mask_ref)
    top, bottom, left, right = losses_oracle.crop_bounds(s, crop)
    margin = ref_eval[:, 1, top:s - bottom, left:s - right].abs().numpy()
    differ = mask.cpu().numpy() != mask_ref
    print('bf16 eval: mean-abs err %.4e, mask IoU vs oracle %.5f, differing pixels %d (max |ref logit| there %.4f)'
          % (mean_err, iou, differ.sum(), margin[differ].max() if differ.any() else 0.0))
    assert ok and mean_err <= 0.03
    assert not differ.any() or margin[differ].max() <= bound
    lt = eng.forward(xd, train=True)
    loss, dlogits = eng.loss_lovasz(lt, td)
    eng.backward(dlogits)
    torch.cuda.synchronize()
    assert report('bf16 train logits', lt, ref, atol=0.05 + 0.08 * ref.abs().max().item())[0]
    assert report('bf16 lovasz loss', loss.cpu()[0], loss_ref, rtol=0.02)[0]
    for k in grad_keys(depth):
        if k.endswith('.conv.bias'):
            continue
        a, r = eng.view(k, grad=True).cpu().flatten(), sd[k].grad.flatten()
        if r.norm().item() == 0.0:                      # dead ReLU in the SE bottleneck: exactly zero in both
            assert a.norm().item() < 1e-12, k
            continue
        cos = F.cosine_similarity(a, r, dim=0).item()
        print('bf16 grad cosine %-50s %.5f  (norm ratio %.4f)' % (k, cos, (a.norm() / (r.norm() + 1e-30)).item()))
        assert cos >= 0.90, k


@pytest.mark.parametrize('depth,prec', [(34, 'bf16'), (34, 'fp32'), (50, 'bf16')])
def test_eval_forward_fused_epilogue(depth, prec, monkeypatch):
    """Inference folds BatchNorm (+ residual, + ReLU, + the replicate border of the decoder inputs) into the convolution
    epilogues.  Same logits as the pass-per-operation path (SALT_ENGINE_FUSE_EVAL=0) up to the rounding the fusion removes
    (the unfused bf16 path rounds the raw convolution output to bf16 before BatchNorm), far fewer launches."""
    from salt_b200 import _lib
    b, s = 4, 128
    sd_np, x, _ = _setup(depth, b, s)
    with torch.no_grad():
        ref = unet_oracle.unet_resnet_forward(unet_oracle.to_torch_state(sd_np), x, depth, False)
    outs, launches = {}, {}
    for fuse in ('1', '0'):
        monkeypatch.setenv('SALT_ENGINE_FUSE_EVAL', fuse)
        eng = _engine(depth, 2, b, s, precision=prec, training=False)
        eng.load_state(sd_np)
        eng.forward(x.cuda(), train=False)                      # packs the weights, computes the eval coefficients
        n0 = _lib.launch_count()
        outs[fuse] = eng.forward(x.cuda(), train=False).clone()
        torch.cuda.synchronize()
        launches[fuse] = _lib.launch_count() - n0
    bound = 1e-3 if prec == 'fp32' else 0.05 + 0.03 * ref.abs().max().item()
    e1, e0 = (outs['1'].cpu() - ref).abs().max().item(), (outs['0'].cpu() - ref).abs().max().item()
    print('eval forward %s depth %d: fused err %.3e (%d launches), unfused err %.3e (%d launches)' % (prec, depth, e1, launches['1'], e0, launches['0']))
    assert e1 <= bound and e0 <= bound
    assert (outs['1'] - outs['0']).abs().max().item() <= (1e-4 if prec == 'fp32' else bound)
    assert launches['1'] < 0.75 * launches['0']


def test_loss_kernels_edge_cases():
    """Lovasz / BCE-Dice kernels alone: empty masks, full masks, single pixels, ties."""
    b, s = 6, 64
    g = torch.Generator().manual_seed(11)
    logits = torch.randn(b, 2, s, s, generator=g)
    logits[3] = torch.round(logits[3] * 2) / 2            # many exact ties
    t = torch.zeros(b, 2, s, s)
    t[1] = 1.0
    t[2, 1, 5, 7] = 1.0
    t[3, :, 10:30, 20:50] = 1.0
    t[4, 0] = 1.0
    t[5] = (torch.rand(2, s, s, generator=g) > 0.5).float()
    eng = _engine(18, 2, b, s, precision='fp32')
    lg = logits.clone().requires_grad_(True)
    ref = losses_oracle.lovasz_hinge_per_image(lg, t)
    ref.backward()
    loss, dl = eng.loss_lovasz(logits.cuda(), t.cuda())
    torch.cuda.synchronize()
    ok1 = report('lovasz loss (edge cases)', loss.cpu()[0], ref, atol=1e-5, rtol=1e-5)[0]
    # tie order may differ -> compare gradients on the tie-free images only
    sel = [0, 1, 2, 4, 5]
    ok2 = report('lovasz dlogits (tie-free images)', dl.cpu()[sel], lg.grad[sel], atol=1e-8, rtol=1e-3)[0]
    lg2 = logits.clone().requires_grad_(True)
    ref2 = losses_oracle.bce_dice(lg2, t)
    ref2.backward()
    loss2, dl2 = eng.loss_bce_dice(logits.cuda(), t.cuda())
    torch.cuda.synchronize()
    ok3 = report('bce+dice loss', loss2.cpu()[0], ref2, atol=1e-6, rtol=1e-5)[0]
    ok4 = report('bce+dice dlogits', dl2, lg2.grad, atol=1e-10, rtol=1e-3)[0]
    assert ok1 and ok2 and ok3 and ok4


def test_lovasz_large_images():
    """256x256 inputs (BASELINE config 4): 131072 logits per image do not fit one CTA's shared memory, the loss runs the
    global-memory bitonic sort.  Same oracle, same tolerances as the in-smem path; one image carries exact ties, one is empty."""
    b, s = 3, 256
    g = torch.Generator().manual_seed(5)
    logits = torch.randn(b, 2, s, s, generator=g) * 2
    logits[1] = torch.round(logits[1] * 4) / 4
    t = torch.zeros(b, 2, s, s)
    t[0, 1, 40:200, 30:180] = 1.0
    t[0, 0] = 1.0 - t[0, 1]
    t[1] = (torch.rand(2, s, s, generator=g) > 0.7).float()
    eng = _engine(18, 2, b, s, precision='fp32', training=False)
    lg = logits.clone().requires_grad_(True)
    ref = losses_oracle.lovasz_hinge_per_image(lg, t)
    ref.backward()
    loss, dl = eng.loss_lovasz(logits.cuda(), t.cuda())
    torch.cuda.synchronize()
    ok1 = report('lovasz loss (131072 logits / image)', loss.cpu()[0], ref, atol=1e-5, rtol=1e-5)[0]
    ok2 = report('lovasz dlogits (tie-free images)', dl.cpu()[[0, 2]], lg.grad[[0, 2]], atol=1e-9, rtol=1e-3)[0]
    # inside a group of tied errors the sort order (hence the per-pixel gradient) is free; the loss above is not
    assert ok1 and ok2 and torch.isfinite(dl).all()


def test_batch_independence_full_size():
    """Size-independent property at the benchmark shape (ResNet-34, 128x128, B=128, bf16): in eval mode every
    image is processed independently, so the first 8 logits of a 128-batch equal those of an 8-batch, and a
    permuted batch gives permuted logits (bit-exact: same kernels, same per-image arithmetic)."""
    depth, s = 34, 128
    sd_np = synth.synth_state_dict(depth, 2, 0)
    x = torch.from_numpy(synth.synth_inputs(128, s, 7)).cuda()
    eng = _engine(depth, 2, 128, s, precision='bf16', training=False)
    eng.load_state(sd_np)
    full = eng.forward(x, train=False).clone()
    part = eng.forward(x[:8].contiguous(), train=False).clone()
    perm = torch.randperm(128, generator=torch.Generator().manual_seed(0)).cuda()
    permuted = eng.forward(x[perm].contiguous(), train=False)
    torch.cuda.synchronize()
    assert torch.isfinite(full).all()
    assert torch.equal(full[:8], part)
    assert torch.equal(full[perm], permuted)

"""CPU: the oracle restatement replays the fixtures that were generated from the UNMODIFIED reference
modules (oracle/make_golden.py).  This is what pins the oracle (the reference has no tests of its own)."""
import os

import numpy as np
import pytest
import torch

from oracle import synth, unet_oracle, losses_oracle
from oracle.make_golden import sample, grad_keys, stem_bn, ARCH_NAMES

CASES = ['r18_b2_s64', 'r34_b2_s64', 'se50_b2_s64', 'se101_b2_s64', 'sex50_b2_s64']


def _load(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, tag + '.npz'))
    meta = {k[5:]: int(g[k]) for k in g.files if k.startswith('meta_')}
    meta['arch'] = ARCH_NAMES[meta.get('arch', 0)]          # None, or 'UNetSeResNetXt' 
    return g, meta


@pytest.mark.parametrize('tag', CASES)
def test_eval_forward_matches_reference(golden_dir, tag):
    g, m = _load(golden_dir, tag)
    sd = unet_oracle.to_torch_state(synth.synth_state_dict(m['depth'], 2, m['wseed'], m['arch']))
    x = torch.from_numpy(synth.synth_inputs(m['batch'], m['size'], m['dseed']))
    with torch.no_grad():
        logits = unet_oracle.unet_resnet_forward(sd, x, m['depth'], train=False, arch=m['arch'])
    assert np.abs(logits.numpy() - g['logits_eval']).max() <= 1e-5


@pytest.mark.parametrize('tag', CASES)
@pytest.mark.parametrize('loss_name', ['lovasz', 'bcedice'])
def test_train_step_matches_reference(golden_dir, tag, loss_name):
    g, m = _load(golden_dir, tag)
    sd = unet_oracle.to_torch_state(synth.synth_state_dict(m['depth'], 2, m['wseed'], m['arch']), requires_grad=True)
    x = torch.from_numpy(synth.synth_inputs(m['batch'], m['size'], m['dseed']))
    t = torch.from_numpy(synth.synth_targets(m['batch'], m['size'], m['dseed']))
    logits = unet_oracle.unet_resnet_forward(sd, x, m['depth'], train=True, arch=m['arch'])
    logits.retain_grad()
    fn = losses_oracle.lovasz_hinge_per_image if loss_name == 'lovasz' else losses_oracle.bce_dice
    loss = fn(logits, t)
    loss.backward()
    assert np.abs(logits.detach().numpy() - g['logits_train']).max() <= 1e-5
    assert abs(loss.item() - float(g['loss_' + loss_name])) <= 1e-5 * max(1.0, abs(loss.item()))
    ref = g['dlogits_' + loss_name]
    assert np.abs(sample(logits.grad.numpy()) - ref).max() <= 1e-7 + 1e-4 * np.abs(ref).max()
    for k in grad_keys(m['depth']):
        ref = g['grad_%s_%s' % (loss_name, k)]
        got = sample(sd[k].grad.numpy())
        assert np.abs(got - ref).max() <= 2e-3 * (np.abs(ref).max() + 1e-12), k
    # BatchNorm running statistics were updated like the reference's
    assert np.abs(sd[stem_bn(m['depth']) + '.running_mean'].numpy() - g['running_mean_stem_' + loss_name]).max() <= 1e-6
    assert np.abs(sd['final.0.batch_norm.running_var'].numpy() - g['running_var_final0_' + loss_name]).max() <= 1e-5


def test_postprocessing_matches_reference(golden_dir):
    g, m = _load(golden_dir, 'r18_b8_s128')
    probs, masks = losses_oracle.predict_masks(g['logits_eval'], g['logits_eval_flip'], 101, 0.5)
    assert np.abs(sample(probs) - g['tta_probs']).max() <= 1e-6
    assert (masks == g['tta_masks']).all()
    assert masks.shape == (m['batch'], 101, 101)


def test_lovasz_edge_cases():
    # empty image (no salt), full image, single pixel: finite and matches the closed forms
    lg = torch.randn(3, 2, 8, 8)
    t = torch.zeros(3, 2, 8, 8)
    t[1] = 1.0
    t[2, 0, 0, 0] = 1.0
    loss = losses_oracle.lovasz_hinge_per_image(lg, t)
    assert torch.isfinite(loss)
    # all-background image: jaccard gradient is (1,0,0,...) -> loss = elu(max error)
    e = 1.0 + lg[0].reshape(-1)
    one = losses_oracle.lovasz_hinge_per_image(lg[:1], t[:1])
    assert abs(one.item() - torch.nn.functional.elu(e.max()).item()) < 1e-6


def test_alias_keys_cover_reference_state_dict():
    amap = unet_oracle.alias_keys(18)
    assert amap['encoders.conv1.0.weight'] == 'encoders.encoder.conv1.weight'
    assert amap['encoders.encoder3.0.downsample.1.running_var'] == 'encoders.encoder.layer2.0.downsample.1.running_var'

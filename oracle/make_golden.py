"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules on CPU.

Build-container only (needs /root/reference):  ``python -m oracle.make_golden``.
Also checks the oracle restatement against the reference while it is at it and
refuses to write fixtures if they disagree.

Each fixture is small: inputs/targets/weights are regenerated from seeds by
``oracle/synth.py``; only the reference *outputs* are stored.
"""
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shims, synth, unet_oracle, losses_oracle  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')

# (tag, depth, batch, size, weight seed, data seed)
CASES = [
    ('r18_b2_s64', 18, 2, 64, 0, 1234),
    ('r34_b2_s64', 34, 2, 64, 1, 4321),
    ('r18_b8_s128', 18, 8, 128, 0, 1234),     # BASELINE.json configs[0] shape
    # UNetSeResNet-50 (BASELINE.json configs[3] architecture): the reference's own UNetSeResNet / SeResNetEncoders classes
    # on top of oracle/senet_restated.py (pretrainedmodels is absent - see that file's header)
    ('se50_b2_s64', 50, 2, 64, 2, 777),
    ('se101_b2_s64', 101, 2, 64, 4, 555),     # depth variant of the same class (encoders.py:54-55: se_resnet101, layers [3,4,23,3])
    # UNetSeResNetXt: the reference's UNetSeResNetXt / SeResNetXtEncoders on the restated se_resnext50_32x4d
    ('sex50_b2_s64', 50, 2, 64, 6, 999, 'UNetSeResNetXt'),
]
ARCH_IDS = {None: 0, 'UNetSeResNetXt': 2}
ARCH_NAMES = {0: None, 2: 'UNetSeResNetXt'}

GRAD_KEYS = ['encoders.encoder.conv1.weight', 'encoders.encoder.layer1.0.conv1.weight',
             'encoders.encoder.layer2.0.downsample.0.weight', 'encoders.encoder.layer4.1.bn2.weight',
             'center.1.conv.weight', 'dec5.conv1.conv.weight', 'dec3.channel_se.fc.0.weight',
             'dec2.spatial_se.fc.weight', 'dec1.conv1.conv.bias', 'final.0.conv.weight',
             'final.0.batch_norm.bias', 'final.1.weight', 'final.1.bias']


GRAD_KEYS_SE50 = ['encoders.encoder.layer0.conv1.weight', 'encoders.encoder.layer1.0.conv1.weight',
                  'encoders.encoder.layer1.0.downsample.0.weight', 'encoders.encoder.layer2.0.conv1.weight',
                  'encoders.encoder.layer2.0.downsample.0.weight', 'encoders.encoder.layer2.1.se_module.fc2.bias',
                  'encoders.encoder.layer3.2.se_module.fc1.weight', 'encoders.encoder.layer3.5.conv2.weight',
                  'encoders.encoder.layer4.1.bn3.weight', 'encoders.encoder.layer4.2.conv3.weight',
                  'center.1.conv.weight', 'dec5.conv1.conv.weight', 'dec3.channel_se.fc.0.weight',
                  'dec2.spatial_se.fc.weight', 'dec1.conv1.conv.bias', 'final.0.conv.weight',
                  'final.0.batch_norm.bias', 'final.1.weight', 'final.1.bias']


def grad_keys(depth):
    return GRAD_KEYS_SE50 if depth >= 50 else GRAD_KEYS


def stem_bn(depth):
    return 'encoders.encoder.layer0.bn1' if depth >= 50 else 'encoders.encoder.bn1'


MAX_SAMPLES = 8192


def sample(a, max_samples=MAX_SAMPLES):
    """Deterministic strided subsample of a flattened array (keeps fixtures small)."""
    flat = np.asarray(a).reshape(-1)
    stride = max(1, -(-flat.size // max_samples))
    return flat[::stride].copy()


def _ref_losses():
    ref_shims.install()
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        from common_blocks import models as ref_models
    return ref_models.lovasz_loss, ref_models.mixed_dice_bce_loss


def run_case(tag, depth, batch, size, wseed, dseed, arch=None):
    torch.manual_seed(0)
    torch.set_num_threads(8)
    sd_np = synth.synth_state_dict(depth, 2, wseed, arch)
    x = torch.from_numpy(synth.synth_inputs(batch, size, dseed))
    t = torch.from_numpy(synth.synth_targets(batch, size, dseed))
    lovasz_ref, bcedice_ref = _ref_losses()

    out = {}
    net = ref_shims.load_numpy_state(ref_shims.reference_unet(depth, 2, arch), sd_np, depth)

    # ---- eval-mode forward (running statistics)
    net.eval()
    with torch.no_grad():
        logits_eval = net(x)
    out['logits_eval'] = logits_eval.numpy()
    with torch.no_grad():
        logits_eval_flip = net(torch.flip(x, dims=[3])).numpy()     # h-flip TTA copy, same (pristine) weights
    sd_t = unet_oracle.to_torch_state(sd_np)
    with torch.no_grad():
        mine = unet_oracle.unet_resnet_forward(sd_t, x, depth, train=False, arch=arch)
    d = (mine - logits_eval).abs().max().item()
    assert d <= 1e-5, 'oracle eval forward differs from reference: %g' % d

    # ---- train-mode forward + both losses + backward (batch statistics)
    for loss_name, loss_fn, mine_fn in (('lovasz', lovasz_ref, losses_oracle.lovasz_hinge_per_image),
                                        ('bcedice', bcedice_ref, losses_oracle.bce_dice)):
        net = ref_shims.load_numpy_state(ref_shims.reference_unet(depth, 2, arch), sd_np, depth)
        net.train()
        logits = net(x)
        logits.retain_grad()
        loss = loss_fn(logits, t.clone())
        loss.backward()
        out['logits_train'] = logits.detach().numpy()
        out['loss_' + loss_name] = np.float32(loss.item())
        out['dlogits_' + loss_name] = sample(logits.grad.numpy())
        named = dict(net.named_parameters())
        for k in grad_keys(depth):
            out['grad_%s_%s' % (loss_name, k)] = sample(named[k].grad.numpy())
            out['gradl2_%s_%s' % (loss_name, k)] = np.float64(named[k].grad.double().norm().item())
        gsq = {k: float((p.grad.double() ** 2).sum()) for k, p in named.items() if p.grad is not None}
        out['gradnorm_' + loss_name] = np.float64(np.sqrt(sum(gsq.values())))
        out['running_mean_stem_' + loss_name] = net.state_dict()[stem_bn(depth) + '.running_mean'].numpy()
        out['running_var_final0_' + loss_name] = net.state_dict()['final.0.batch_norm.running_var'].numpy()

        sd_t = unet_oracle.to_torch_state(sd_np, requires_grad=True)
        lg = unet_oracle.unet_resnet_forward(sd_t, x, depth, train=True, arch=arch)
        lg.retain_grad()
        ls = mine_fn(lg, t)
        ls.backward()
        assert (lg.detach() - logits.detach()).abs().max().item() <= 1e-5, 'oracle train forward differs'
        assert abs(ls.item() - loss.item()) <= 1e-5 * max(1.0, abs(loss.item())), \
            'oracle %s loss differs: %r vs %r' % (loss_name, ls.item(), loss.item())
        dd = (lg.grad - logits.grad).abs().max().item()
        assert dd <= 1e-7 + 1e-4 * logits.grad.abs().max().item(), 'oracle dlogits differ: %g' % dd
        for k in grad_keys(depth):
            a, b = sd_t[k].grad, named[k].grad
            rel = (a - b).abs().max().item() / (b.abs().max().item() + 1e-12)
            assert rel <= 2e-3, 'oracle grad %s differs rel %g' % (k, rel)
        rm = sd_t[stem_bn(depth) + '.running_mean']
        assert (rm - net.state_dict()[stem_bn(depth) + '.running_mean']).abs().max().item() <= 1e-6

    # ---- post-processing (sigmoid, TTA h-flip mean, crop, binarize) through the reference functions
    if size == 128:
        ref_shims.install()
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            from common_blocks import postprocessing as ref_post
            from common_blocks.utils import sigmoid as ref_sigmoid
        lf = logits_eval_flip
        probs, masks = [], []
        for i in range(batch):
            p0 = ref_sigmoid(np.squeeze(out['logits_eval'][i]))
            p1 = ref_sigmoid(np.squeeze(lf[i]))
            p1 = np.stack([np.fliplr(ch) for ch in p1])            # augmentation.py:170-175
            agg = np.mean(np.stack([p0, p1], axis=-1), axis=-1)    # loaders.py:751-760
            probs.append(agg)
            masks.append(ref_post.binarize(ref_post.crop_image(agg, (101, 101)), 0.5))
        out['logits_eval_flip'] = lf
        full_probs = np.stack(probs).astype(np.float32)
        out['tta_probs'] = sample(full_probs)
        out['tta_masks'] = np.stack(masks).astype(np.uint8)
        p_mine, m_mine = losses_oracle.predict_masks(out['logits_eval'], lf, 101, 0.5)
        assert np.abs(p_mine - full_probs).max() <= 1e-6
        assert (m_mine == out['tta_masks']).all()

    meta = dict(depth=depth, batch=batch, size=size, wseed=wseed, dseed=dseed)
    if arch is not None:
        meta['arch'] = ARCH_IDS[arch]
    np.savez_compressed(os.path.join(GOLDEN_DIR, tag + '.npz'), **out,
                        **{'meta_' + k: np.int64(v) for k, v in meta.items()})
    print('wrote', tag, {k: getattr(v, 'shape', None) for k, v in out.items() if k.startswith('lo')})


def io_inputs():
    """Seeded inputs of the I/O fixture (regenerated by the tests; only reference OUTPUTS are stored)."""
    rng = np.random.default_rng(2024)
    tiles = synth.synth_tiles_u8(3, 101, 77)                                   # the competition's tile size -> 128
    tiles_odd = rng.integers(0, 256, (2, 50, 37), dtype=np.uint8)              # ragged tile -> 64 (odd pads both ways)
    masks = []
    for kind in range(8):
        m = np.zeros((101, 101), np.uint8)
        if kind == 1:
            m[:] = 1                                                           # full: one run of 10201
        elif kind == 2:
            m[100, 3] = 1; m[0, 4] = 1                                         # run crossing a column boundary
        elif kind == 3:
            m[::2, :] = 1                                                      # many runs; (101*101+1)/2 = 5101 of length 1
            m = (np.arange(101 * 101).reshape(101, 101).T % 2 == 0).astype(np.uint8)
        elif kind == 4:
            m[0, 0] = 1; m[100, 100] = 1                                       # first and last pixel
        elif kind >= 5:
            m = (rng.random((101, 101)) < (0.1, 0.5, 0.9)[kind - 5]).astype(np.uint8)
        masks.append(m)
    masks = np.stack(masks)
    logits = (rng.standard_normal((6, 2, 128, 128)) * 1.5).astype(np.float32)
    logits[:, 1] += np.linspace(-2, 2, 128, dtype=np.float32)[None, None, :]   # a salt/no-salt edge
    logits[4, 1] = -9.0                                                        # empty prediction
    y_true = np.zeros((6, 101, 101), np.uint8)
    y_true[:, :, 55:] = 1
    y_true[1] = (rng.random((101, 101)) < 0.5)
    y_true[3] = 0                                                              # empty truth, non-empty prediction
    y_true[4] = 0                                                              # both empty
    y_true[5, :, :] = 1
    return dict(tiles=tiles, tiles_odd=tiles_odd, masks=masks, logits=logits, y_true=y_true)


def make_io_golden():
    """tests/golden/io_cases.npz: outputs of the UNMODIFIED reference functions for SURVEY.md 8(f) N1-N3."""
    from PIL import Image
    from torchvision import transforms
    from oracle import io_oracle
    ref_shims.install()
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        from common_blocks import utils as ref_utils
        from common_blocks import metrics as ref_metrics
        from common_blocks import postprocessing as ref_post
    inp = io_inputs()
    out = {}

    # ---- N2 tile adapter: torchvision transforms of loaders.py:607-612 + the reference AddDepthChannels / pad split
    tf = transforms.Compose([transforms.Grayscale(num_output_channels=3), transforms.ToTensor(),
                             transforms.Normalize(mean=synth.MEAN, std=synth.STD), ref_utils.AddDepthChannels()])

    def ref_adapt(tiles, size, flip):
        res = []
        for tile in tiles:
            if flip:
                tile = np.fliplr(tile)                                         # augmentation.py:146-147
            top, right, bottom, left = ref_utils.get_crop_pad_sequence(size - tile.shape[0], size - tile.shape[1])
            padded = np.pad(tile, ((top, bottom), (left, right)), mode='edge')  # imgaug iaa.Pad(px=(t,r,b,l), pad_mode='edge')
            res.append(tf(Image.fromarray(padded)).numpy())
        return np.stack(res)
    for key, tiles, size in (('adapt', inp['tiles'], 128), ('adapt_odd', inp['tiles_odd'], 64)):
        for flip in (0, 1):
            ref = ref_adapt(tiles, size, flip)
            mine = io_oracle.adapt_tiles_hflip(tiles, size) if flip else io_oracle.adapt_tiles(tiles, size)
            assert mine.dtype == ref.dtype and (mine == ref).all(), 'tile adapter restatement is not bit-exact (%s flip %d)' % (key, flip)
            out['%s_flip%d' % (key, flip)] = ref

    # ---- N3 run-length encoding (utils.py:99-111) and its inverse (utils.py:114-132)
    flat, lens = [], []
    for m in inp['masks']:
        ref = ref_utils.run_length_encoding(m)
        assert ref == io_oracle.run_length_encoding(m)
        if ref:
            back = ref_utils.run_length_decoding(' '.join(str(v) for v in ref), m.shape)
            assert (back == m).all()
        flat.extend(ref); lens.append(len(ref))
    out['rle_flat'] = np.asarray(flat, np.int64)
    out['rle_lens'] = np.asarray(lens, np.int64)

    # ---- N1 validation sweep: callbacks.py:499-527 replayed with the reference's sigmoid / crop_image / binarize / metrics
    probs = [ref_utils.sigmoid(np.squeeze(m)) for m in inp['logits']]
    y_true = list(inp['y_true'])
    iout_best, threshold_best, seen = 0.0, 0.5, []
    all_iout = []
    for thr in np.linspace(0.5, 0.3, 21):
        y_pred = [ref_post.binarize(ref_post.crop_image(p, (101, 101)), thr) for p in probs]
        all_iout.append(ref_metrics.intersection_over_union_thresholds(y_true, y_pred))
    for thr, sc in zip(np.linspace(0.5, 0.3, 21), all_iout):
        seen.append(sc)
        if sc > iout_best:
            iout_best, threshold_best = sc, thr
        else:
            break
    y_pred = [ref_post.binarize(ref_post.crop_image(p, (101, 101)), threshold_best) for p in probs]
    out['val_threshold'] = np.float64(threshold_best)
    out['val_iout'] = np.float64(ref_metrics.intersection_over_union_thresholds(y_true, y_pred))
    out['val_iou'] = np.float64(ref_metrics.intersection_over_union(y_true, y_pred))
    out['val_iout_all'] = np.asarray(all_iout, np.float64)
    out['val_pred_best'] = np.stack(y_pred).astype(np.uint8)
    mine = io_oracle.validation_sweep(inp['logits'], y_true)
    assert mine['threshold'] == float(threshold_best) and abs(mine['iout'] - out['val_iout']) < 1e-12 \
        and abs(mine['iou'] - out['val_iou']) < 1e-12, (mine, threshold_best, out['val_iout'], out['val_iou'])
    assert np.allclose(mine['iouts_seen'], seen, atol=1e-12)
    np.savez_compressed(os.path.join(GOLDEN_DIR, 'io_cases.npz'), **out)
    print('wrote io_cases', {k: getattr(v, 'shape', None) for k, v in out.items()})


def main():
    if not ref_shims.available():
        sys.exit('reference tree not available; fixtures can only be regenerated in the build container')
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    only = set(sys.argv[1:])                   # optional: tags to (re)generate
    for case in CASES:
        if not only or case[0] in only:
            run_case(*case)
    if not only or 'io_cases' in only:
        make_io_golden()


if __name__ == '__main__':
    main()

"""SURVEY.md section 8(f) N1-N3 - the data formats either side of the network: u8 tile adapter, run-length encoding,
validation threshold sweep.

CPU (`-m "not gpu"`): the numpy oracle replays tests/golden/io_cases.npz, which oracle/make_golden.py produced with the
UNMODIFIED reference functions; the host-side selection logic (salt_b200/validation.py) is checked on oracle counts.
GPU (`-m gpu`): the CUDA kernels, called through the C ABI, against the same oracle and fixtures - bit-exact (integer / byte
work; the adapter's fp32 arithmetic is IEEE add/div and also bit-exact).
"""
import os

import numpy as np
import pytest
import torch

from oracle import io_oracle
from oracle.make_golden import io_inputs

gpu = pytest.mark.gpu


@pytest.fixture(scope='module')
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, 'io_cases.npz'))


@pytest.fixture(scope='module')
def inp():
    return io_inputs()


def _split_rle(gold):
    flat, lens = gold['rle_flat'], gold['rle_lens']
    out, o = [], 0
    for n in lens:
        out.append(flat[o:o + n].tolist())
        o += n
    return out


# ------------------------------------------------------------------------------------------------- CPU: oracle vs reference
def test_oracle_adapter_matches_reference(gold, inp):
    for key, tiles, size in (('adapt', inp['tiles'], 128), ('adapt_odd', inp['tiles_odd'], 64)):
        assert (io_oracle.adapt_tiles(tiles, size) == gold[key + '_flip0']).all()
        assert (io_oracle.adapt_tiles_hflip(tiles, size) == gold[key + '_flip1']).all()


def test_oracle_rle_matches_reference(gold, inp):
    for m, ref in zip(inp['masks'], _split_rle(gold)):
        got = io_oracle.run_length_encoding(m)
        assert got == ref
        assert (io_oracle.run_length_decoding(got, m.shape) == m).all()


def test_oracle_validation_sweep_matches_reference(gold, inp):
    r = io_oracle.validation_sweep(inp['logits'], list(inp['y_true']))
    assert r['threshold'] == float(gold['val_threshold'])
    assert abs(r['iout'] - float(gold['val_iout'])) < 1e-12 and abs(r['iou'] - float(gold['val_iou'])) < 1e-12
    n = len(r['iouts_seen'])
    assert np.allclose(r['iouts_seen'], gold['val_iout_all'][:n], atol=1e-12)


def test_host_threshold_selection_on_oracle_counts(gold, inp):
    """salt_b200.validation (host logic above the counts kernel) reproduces the reference's sweep."""
    from salt_b200 import validation
    inter, pred, gts = io_oracle.validation_counts(inp['logits'], inp['y_true'])
    r = validation.select_threshold(inter, pred, gts)
    assert r['threshold'] == float(gold['val_threshold'])
    assert abs(r['iout'] - float(gold['val_iout'])) < 1e-12 and abs(r['iou'] - float(gold['val_iou'])) < 1e-12
    assert np.allclose(r['iout_per_threshold'], gold['val_iout_all'], atol=1e-12)


def test_host_threshold_selection_edge_cases():
    from salt_b200 import validation
    # nothing ever matches: IoUT 0 at the first threshold -> the reference keeps threshold_best = 0.5 (callbacks.py:501)
    inter = np.zeros((3, 21), np.int64); pred = np.full((3, 21), 5, np.int64); gts = np.array([7, 7, 7])
    r = validation.select_threshold(inter, pred, gts)
    assert r['threshold'] == 0.5 and r['iout'] == 0.0 and r['iou'] == 0.0
    # all empty on both sides: IoU 1 everywhere, first threshold wins, sweep stops at the second (not strictly better)
    z = np.zeros((2, 21), np.int64)
    r = validation.select_threshold(z, z, np.zeros(2, np.int64))
    assert r['threshold'] == 0.5 and r['iout'] == 1.0 and r['iou'] == 1.0


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_io_ops_have_no_cpu_fallback():
    """Without a CUDA device the product entry points raise instead of computing on the host (the oracle is never used)."""
    import ctypes as C
    from salt_b200 import _lib, io_ops, validation
    lib = _lib.load()
    with pytest.raises(_lib.SaltEngineError):
        io_ops.adapt_tiles(torch.zeros((1, 101, 101), dtype=torch.uint8))
    with pytest.raises(_lib.SaltEngineError):
        io_ops.encode_rle([np.zeros((101, 101), np.uint8)])
    with pytest.raises(_lib.SaltEngineError):
        validation.validation_counts(torch.zeros(1, 2, 128, 128), torch.zeros((1, 101, 101), dtype=torch.uint8))
    buf = (C.c_ubyte * 16)()
    assert lib.salt_rle_encode(buf, 1, 4, 4, 8, buf, buf, None) != 0 and b'no CUDA device' in lib.salt_last_error()
    assert lib.salt_adapt_tiles(buf, 1, 4, 4, 8, 0.5, 0.5, 0, buf, None) != 0 and b'no CUDA device' in lib.salt_last_error()


# ------------------------------------------------------------------------------------------------- GPU: kernels vs oracle
@gpu
def test_adapt_tiles_bit_exact(gold, inp):
    from salt_b200 import io_ops
    for key, tiles, size in (('adapt', inp['tiles'], 128), ('adapt_odd', inp['tiles_odd'], 64)):
        for flip in (0, 1):
            got = io_ops.adapt_tiles(torch.from_numpy(tiles).cuda(), size, hflip=bool(flip)).cpu().numpy()
            ref = gold['%s_flip%d' % (key, flip)]
            assert got.dtype == ref.dtype and (got == ref).all(), (key, flip, np.abs(got - ref).max())
    # maximum size (tile == network input: no padding) and a 1x1 tile (everything is border)
    rng = np.random.default_rng(5)
    for th, tw, size in ((128, 128, 128), (1, 1, 32), (101, 101, 256)):
        t = rng.integers(0, 256, (2, th, tw), dtype=np.uint8)
        got = io_ops.adapt_tiles(torch.from_numpy(t).cuda(), size).cpu().numpy()
        assert (got == io_oracle.adapt_tiles(t, size)).all()
    # empty batch
    assert io_ops.adapt_tiles(torch.zeros((0, 101, 101), dtype=torch.uint8).cuda(), 128).shape == (0, 3, 128, 128)


@gpu
@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_forward_tiles_equals_forward_of_adapted_input(precision):
    """The adapter fused into the stem's im2col gives the same logits, bit for bit, as feeding the adapted fp32 tensor."""
    from oracle import synth
    from salt_b200.engine import UNetEngine
    eng = UNetEngine(18, 2, 3, 128, precision=precision)
    eng.load_state(synth.synth_state_dict(18, 2, 0))
    tiles = synth.synth_tiles_u8(3, 101, 31)
    for flip in (False, True):
        x = io_oracle.adapt_tiles_hflip(tiles, 128) if flip else io_oracle.adapt_tiles(tiles, 128)
        a = eng.forward(torch.from_numpy(x).cuda(), train=False).clone()
        b = eng.forward_tiles(torch.from_numpy(tiles).cuda(), train=False, hflip=flip)
        assert torch.equal(a, b)
    # train mode reads the same stem patches and the BatchNorm statistics are reduced in a fixed order (kernels.h SALT_STAT_SLOTS):
    # bit for bit the same logits as well, run to run and tiles-vs-tensor
    x = torch.from_numpy(io_oracle.adapt_tiles(tiles, 128)).cuda()
    state = synth.synth_state_dict(18, 2, 0)
    outs = []
    for kind in ('tensor', 'tiles', 'tensor'):
        eng.load_state(state)                         # reset the running statistics the training forward updates
        out = eng.forward(x, train=True) if kind == 'tensor' else eng.forward_tiles(torch.from_numpy(tiles).cuda(), train=True)
        outs.append(out.clone())
    assert torch.equal(outs[0], outs[2]), 'train-mode forward is not reproducible run to run'
    assert torch.equal(outs[0], outs[1])


@gpu
def test_rle_encode_bit_exact(gold, inp):
    from salt_b200 import io_ops
    got = io_ops.encode_rle(torch.from_numpy(inp['masks']).cuda())
    ref = _split_rle(gold)
    assert got == ref
    # ragged shapes, single row / single column, and a capped output buffer
    rng = np.random.default_rng(11)
    for h, w in ((1, 1), (1, 64), (64, 1), (7, 13), (202, 202)):
        m = (rng.random((5, h, w)) < 0.4).astype(np.uint8)
        m[0] = 0
        m[1] = 1
        got = io_ops.encode_rle(torch.from_numpy(m).cuda())
        assert got == [io_oracle.run_length_encoding(x) for x in m], (h, w)
    m = (np.arange(101 * 101).reshape(101, 101).T % 2 == 0).astype(np.uint8)[None]
    runs, nruns = io_ops.rle_encode_device(torch.from_numpy(m).cuda(), cap_runs=16)
    assert int(nruns[0]) == 5101 and runs.shape == (1, 16, 2)
    assert runs[0].cpu().numpy().reshape(-1).tolist() == io_oracle.run_length_encoding(m[0])[:32]
    assert io_ops.encode_rle(torch.zeros((0, 101, 101), dtype=torch.uint8)) == []
    import pandas as pd
    sub = io_ops.create_submission(pd.DataFrame({'id': ['a', 'b']}), [inp['masks'][1], inp['masks'][0]])
    assert list(sub.columns) == ['id', 'rle_mask'] and sub.values.tolist() == [['a', '1 10201'], ['b', '']]
    assert io_ops.create_submission(['a'], [inp['masks'][4]]).values.tolist() == [['a', '1 1 10201 1']]


@gpu
def test_rle_round_trip_full_size():
    """Size-independent property at the bench's batch: decode(encode(m)) == m for 512 random masks."""
    from salt_b200 import io_ops
    rng = np.random.default_rng(3)
    m = (rng.random((512, 101, 101)) < rng.random((512, 1, 1))).astype(np.uint8)
    for x, rle in zip(m, io_ops.encode_rle(torch.from_numpy(m).cuda())):
        assert (io_oracle.run_length_decoding(rle, x.shape) == x).all()
        assert sum(rle[1::2]) == int(x.sum())


@gpu
def test_validation_counts_and_sweep(gold, inp):
    from salt_b200 import validation
    lg = torch.from_numpy(inp['logits']).cuda()
    gt = torch.from_numpy(inp['y_true']).cuda()
    inter, pred, gts = validation.validation_counts(lg, gt)
    o_inter, o_pred, o_gts = io_oracle.validation_counts(inp['logits'], inp['y_true'])
    # integer counts: exact except for pixels whose probability is within 2e-7 of a threshold (expf vs numpy exp, 1 ulp)
    p = io_oracle.crop_image(io_oracle.sigmoid(inp['logits'][:, 1]), (101, 101)).astype(np.float64)
    near = np.stack([(np.abs(p - t) <= 2e-7).reshape(len(p), -1).sum(1) for t in validation.SWEEP_THRESHOLDS], 1)
    assert (np.abs(pred.cpu().numpy() - o_pred) <= near).all()
    assert (np.abs(inter.cpu().numpy() - o_inter) <= near).all()
    assert (gts.cpu().numpy() == o_gts).all()
    sc = validation.ValidationScorer()
    sc.update(lg[:4], gt[:4])
    sc.update(lg[4:], gt[4:])            # ragged second batch
    r = sc.result()
    assert r['threshold'] == float(gold['val_threshold'])
    assert abs(r['iout'] - float(gold['val_iout'])) < 1e-9 and abs(r['iou'] - float(gold['val_iou'])) < 1e-6
    # TTA variant against the oracle
    lf = torch.from_numpy(np.ascontiguousarray(inp['logits'][:, :, :, ::-1] * 0.9)).cuda()
    i2, p2, _ = validation.validation_counts(lg, gt, logits_flip=lf)
    oi2, op2, _ = io_oracle.validation_counts(inp['logits'], inp['y_true'], logits_flip=lf.cpu().numpy())
    assert np.abs(p2.cpu().numpy() - op2).max() <= 2 and np.abs(i2.cpu().numpy() - oi2).max() <= 2


def test_validation_monitor_dropin_logic(gold, inp):
    """salt_b200.validation.ValidationMonitor (drop-in for callbacks.py:455-527) on a fake transformer: the callback surface, the
    epoch_every gate, the averaged validation loss and the reference's threshold / IoU / IoUT (counts supplied by the oracle, so this
    runs without a GPU; the GPU test below runs the real kernel through the same class)."""
    from salt_b200 import validation

    def oracle_counts(logits, gt, thresholds, logits_flip):
        return tuple(torch.from_numpy(np.ascontiguousarray(a)) for a in
                     io_oracle.validation_counts(logits.numpy(), gt.numpy(), thresholds=thresholds))

    class FakeModel:
        training = True
        def eval(self): self.training = False; return self
        def train(self): self.training = True; return self
        def __call__(self, X): assert not self.training; return X        # the "network" returns its input as logits

    class FakeTransformer:
        model, optimizer, output_names, activation_func, validation_loss, dp = FakeModel(), None, ['mask'], 'sigmoid', {}, None
        loss_function = [('mask', lambda out, tgt: torch.tensor([float(out.shape[0])]), 0.5)]

    lg = torch.from_numpy(inp['logits'])
    batches = [[lg[:4], None], [lg[4:], None]]
    mon = validation.ValidationMonitor(epoch_every=2, y_true=list(inp['y_true']),
                                       scorer_factory=lambda dp: validation.ValidationScorer(dp=dp, counts_fn=oracle_counts))
    tr = FakeTransformer()
    mon.set_params(tr, validation_datagen=(batches, 2), meta_valid=None)
    mon.on_train_begin()
    mon.on_epoch_end()                       # epoch 0: 0 % 2 == 0 -> validates
    assert set(tr.validation_loss) == {0} and tr.model.training
    v = tr.validation_loss[0]
    assert abs(float(v['sum'][0]) - (4 + 2) * 0.5 / 2) < 1e-6            # sum of weighted batch losses / steps (callbacks.py:552)
    assert abs(float(v['iou'][0]) - float(gold['val_iou'])) < 1e-6 and abs(float(v['iout'][0]) - float(gold['val_iout'])) < 1e-6
    assert mon.last_result['threshold'] == float(gold['val_threshold'])
    mon.on_epoch_end()                       # epoch 1: skipped
    assert set(tr.validation_loss) == {0} and mon.epoch_id == 2
    with pytest.raises(NotImplementedError):
        validation.ValidationMonitor(loader_mode='resize')


@gpu
def test_validation_monitor_on_engine(monkeypatch):
    """The same class on the real engine: eval forward on the GPU, counts kernel, reference selection logic; compared with the
    oracle sweep over the engine's own logits."""
    from oracle import synth
    from salt_b200 import validation
    from salt_b200.models import SegmentationModel
    for k, v in dict(SALT_ENGINE_PRECISION='fp32', SALT_ENGINE_MAX_BATCH='4', SALT_ENGINE_SIZE='128', SALT_ENGINE_LOSS='lovasz').items():
        monkeypatch.setenv(k, v)
    arch = {'model_params': {'architecture': 'UNetResNet', 'encoder_depth': 18, 'in_channels': 3, 'out_channels': 2, 'activation': 'sigmoid'},
            'optimizer_params': {'lr': 1e-4}, 'regularizer_params': {'regularize': True, 'weight_decay_conv2d': 1e-4},
            'weights_init': {'function': 'he', 'pretrained': False}}
    m = SegmentationModel(arch, {'epochs': 1}, {})
    m.engine.load_state(synth.synth_state_dict(18, 2, 0))
    xs = [torch.from_numpy(synth.synth_inputs(4, 128, 50 + i)) for i in range(2)]
    ts = [torch.from_numpy(synth.synth_targets(4, 128, 50 + i)) for i in range(2)]
    y_true = [t[i, 1, 13:114, 14:115].numpy().astype(np.uint8) for t in ts for i in range(4)]
    mon = validation.ValidationMonitor(epoch_every=1, y_true=y_true)
    mon.set_params(m, validation_datagen=(list(zip(xs, ts)), 2))
    mon.on_train_begin()
    mon.on_epoch_end()
    v = m.validation_loss[0]
    logits = np.concatenate([m.engine.forward(x.cuda(), train=False).cpu().numpy() for x in xs])
    ref = io_oracle.validation_sweep(logits, y_true)
    print('validation monitor:', mon.last_result['threshold'], float(v['iou'][0]), float(v['iout'][0]), 'oracle', ref)
    assert mon.last_result['threshold'] == ref['threshold']
    # (a pixel whose probability is within 1 ulp of the threshold may flip: one pixel moves an image's IoU by ~1e-4)
    assert abs(float(v['iou'][0]) - ref['iou']) < 2e-4 and abs(float(v['iout'][0]) - ref['iout']) < 1e-6
    assert np.isfinite(float(v['sum'][0]))

#!/usr/bin/env python
"""Benchmark of the hot path: one training step of UNetResNet-34 (128x128 network input = one padded 101x101
tile, bf16, 128 images per GPU, BCE+Dice) - BASELINE.json configs[1] - through the drop-in SegmentationModel.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                     (the reference algorithm's CPU path, oracle port)

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definition of every field.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, 'open-solution-salt-identification_b200')
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

# stdout must carry exactly one JSON line: keep NCCL's banner ("NCCL version ...") off it
if os.environ.get('NCCL_DEBUG', '').upper() in ('', 'VERSION'):
    os.environ['NCCL_DEBUG'] = 'WARN'

import numpy as np  # noqa: E402
import torch        # noqa: E402

DEPTH, SIZE, BATCH_PER_GPU, CLASSES = 34, 128, 128, 2
WORKLOAD = 'UNetResNet-34 train step (fwd + BCE+Dice + bwd + Adam-L2), 3x128x128 inputs (padded 101x101 tiles), bf16, batch 128 per GPU'
# algorithmic conv FLOPs of one training step per image: fwd + dgrad + wgrad
TRAIN_GFLOP_PER_IMAGE = 58.42
CPU_REFERENCE_BATCH = 128      # --impl reference: the SAME 128-image batch as the GPU arm (one CPU step is ~5 s on 16 cores)
CPU_SAMPLE_BATCH = 32          # cpu_baseline inside the GPU arm's line: a bounded sample (3 steps, ~10 s)


def conv_source_sha():
    """sha1 over the tensor-core convolution sources: profiles/roofline_traffic.json is only valid for the kernels it was captured on."""
    import hashlib
    h = hashlib.sha1()
    for f in ('conv_tc.cu', 'conv_tc_rows.cu', 'conv_wgrad_tc.cu', 'tc_common.cuh'):
        with open(os.path.join(PKG, 'csrc', f), 'rb') as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def _peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {'tflops': float(p.get('bf16_tflops_sustained', p.get('bf16_tflops'))), 'hbm_gbs': float(p['hbm_gbs']),
                'source': 'MEASURED_PEAKS.json (bf16_tflops_sustained)'}
    return {'tflops': 1400.0, 'hbm_gbs': 6650.0, 'source': 'fallback (B200_PROFILING.md: ~1.4 PFLOP/s sustained, 6.65 TB/s)'}


# ------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi in a side process, one line every 20 ms with a timestamp; started BEFORE the warm-up (the process needs ~100 ms to
    deliver its first line, a 20-step timed region is only ~0.3 s long), the samples are then cut to the timed region [t0, t1]."""
    Q = 'timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,' \
        'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix='.csv')
        self.proc = None
        self.t0 = self.t1 = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(gpu_index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                          '-lms', '20'], stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def begin(self):
        self.t0 = time.time()

    def end(self):
        self.t1 = time.time()

    def stop(self):
        import datetime
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.proc is None:
            return out
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = []
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(',')]
                if len(f) < 9:
                    continue
                try:
                    ts = datetime.datetime.strptime(f[0], '%Y/%m/%d %H:%M:%S.%f').timestamp()
                    rows.append((ts, float(f[1]), float(f[2]), f[5:9]))
                except ValueError:
                    continue
            os.unlink(self.path)
        except Exception:
            pass
        inside = [r for r in rows if self.t0 is not None and self.t0 <= r[0] <= self.t1]
        if len(inside) < 2 and self.t0 is not None:          # clock skew between time.time() and the tool's timestamps: widen
            inside = [r for r in rows if self.t0 - 0.1 <= r[0] <= self.t1 + 0.1]
        if not inside:
            inside = rows
        sm, mx, reasons = [r[1] for r in inside], [r[2] for r in inside], set()
        for r in inside:
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------- CPU arm
def cpu_train_steps(steps, warmup, batch, threads=None):
    """The reference algorithm's CPU path (oracle port, plain PyTorch fp32 on the host cores): `steps` timed training
    steps on a bounded sample of `batch` images of the workload.  Returns (images_per_s, seconds_per_step, threads)."""
    from oracle import synth, unet_oracle, losses_oracle
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd = unet_oracle.to_torch_state(synth.synth_state_dict(DEPTH, CLASSES, 0), requires_grad=True)
    params = {k: v for k, v in sd.items() if v.requires_grad}
    m = {k: torch.zeros_like(v) for k, v in params.items()}
    vv = {k: torch.zeros_like(v) for k, v in params.items()}
    x = torch.from_numpy(synth.synth_inputs(batch, SIZE, 1234))
    t = torch.from_numpy(synth.synth_targets(batch, SIZE, 1234))
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        for p in params.values():
            p.grad = None
        loss = losses_oracle.bce_dice(unet_oracle.unet_resnet_forward(sd, x, DEPTH, train=True), t)
        loss.backward()
        with torch.no_grad():
            unet_oracle.adam_l2_step({k: v.data for k, v in params.items()}, {k: v.grad for k, v in params.items()}, m, vv, it + 1)
        loss.item()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    sec = float(np.sum(times))
    return batch * steps / sec, sec / steps, threads


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    ips, sec_step, threads = cpu_train_steps(args.steps, args.warmup, CPU_REFERENCE_BATCH)
    sample = ('the full %d-image batch per step; PORT of the reference algorithm (oracle/, plain PyTorch fp32 on the host cores, pinned to '
              'the unmodified reference modules by oracle/make_golden.py), %d threads' % (CPU_REFERENCE_BATCH, threads))
    line = {'impl': 'reference', 'metric': 'images/sec', 'value': ips, 'unit': 'images/s', 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': sec_step * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic', 'config': {'workload': WORKLOAD, 'global_batch': CPU_REFERENCE_BATCH, 'loss': 'bce_dice', 'sample': sample},
            'cpu_baseline': {'value': ips, 'unit': 'images/s', 'cores': threads, 'kind': 'port', 'sample': sample},
            'e2e': {'value': ips, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------- our arm
def make_model(precision, batch, loss):
    from salt_b200.models import SegmentationModel
    os.environ['SALT_ENGINE_PRECISION'] = precision
    os.environ['SALT_ENGINE_MAX_BATCH'] = str(batch)
    os.environ['SALT_ENGINE_SIZE'] = str(SIZE)
    os.environ['SALT_ENGINE_LOSS'] = loss
    arch = {'model_params': {'architecture': 'UNetResNet', 'encoder_depth': DEPTH, 'in_channels': 3, 'out_channels': CLASSES, 'activation': 'sigmoid'},
            'optimizer_params': {'lr': 1e-4}, 'regularizer_params': {'regularize': True, 'weight_decay_conv2d': 1e-4},
            'weights_init': {'function': 'he', 'pretrained': False}}
    return SegmentationModel(arch, {'epochs': 1}, {})


def se50_train_step(ctx, timed, batch=64, size=256):
    """BASELINE configs[3]: UNetSeResNet-50, 202x202 tiles padded to a 256x256 network input, bf16, 64 images per GPU, full
    training step (forward, Lovasz hinge with the global-memory sort, backward, gradient all-reduce, Adam).  Secondary number."""
    from salt_b200 import synthetic as synth
    from salt_b200.engine import UNetEngine
    eng = UNetEngine(architecture='UNetSeResNet', encoder_depth=50, num_classes=CLASSES, max_batch=batch, size=size,
                     precision='bf16', device=ctx.device)
    eng.load_state(synth.synth_state_dict(50, CLASSES, 0))
    x = torch.from_numpy(synth.synth_inputs(batch, size, 99 + ctx.rank)).to(eng.device)
    t = torch.from_numpy(synth.synth_targets(batch, size, 99 + ctx.rank)).to(eng.device)

    def step():
        logits = eng.forward(x, train=True)
        _, dl = eng.loss_lovasz(logits, t)
        eng.backward(dl)
        eng.adam_step(grad_scale=ctx.allreduce_grads(eng.grads))
    for _ in range(2):
        step()
    ms = timed(step, 3) / 3
    eng.profile(True)
    step()
    prof = eng.profile_read()
    eng.profile(False)
    tot_ms = sum(v[0] for v in prof.values())
    tot_fl = sum(v[1] for v in prof.values())
    out = {'value': batch * ctx.world / (ms / 1e3), 'unit': 'images/s', 'ms_per_step': ms,
           'what': 'UNetSeResNet-50, %dx%d input, bf16, %d images/GPU, Lovasz hinge, full training step' % (size, size, batch),
           'conv_tflops': tot_fl / (tot_ms * 1e-3) / 1e12 if tot_ms > 0 else 0.0, 'conv_ms_per_step': tot_ms,
           'train_gflop_per_image': tot_fl / batch / 1e9}
    del eng, x, t
    torch.cuda.empty_cache()
    return out


def tta_parity(eng, synth, dev, tiles=32, train_steps=40):
    """Accuracy side of BASELINE configs[4] ("throughput + IoU vs ref"): train the engine briefly on learnable synthetic scenes
    so that the masks are non-trivial, then compare TTA masks / logits of `tiles` tiles with the oracle (CPU, fp32)."""
    from oracle import unet_oracle, losses_oracle
    for it in range(train_steps):
        x, t = synth.synth_salt_scenes(64, SIZE, 1000 + it)
        logits = eng.forward(torch.from_numpy(x).to(dev), train=True)
        _, dl = eng.loss_bce_dice(logits, torch.from_numpy(t).to(dev))
        eng.backward(dl)
        eng.adam_step(lr=3e-4)
    sd_np = {k: eng.view(k).cpu().numpy().copy() for k in eng.table}
    x = torch.from_numpy(synth.synth_salt_scenes(tiles, SIZE, 77)[0])
    sd = unet_oracle.to_torch_state(sd_np)
    with torch.no_grad():
        ro = unet_oracle.unet_resnet_forward(sd, x, DEPTH, train=False)
        rf = unet_oracle.unet_resnet_forward(sd, torch.flip(x, dims=[3]), DEPTH, train=False)
    _, mask_ref = losses_oracle.predict_masks(ro.numpy(), rf.numpy(), 101, 0.5)
    xd = x.to(dev)
    lo = eng.forward(xd, train=False).clone()
    lf = eng.forward(torch.flip(xd, dims=[3]).contiguous(), train=False)
    _, mask = eng.predict(lo, lf, crop=101, threshold=0.5, want_probs=False)
    return {'iou_vs_oracle': losses_oracle.iou_masks(mask.cpu().numpy(), mask_ref), 'logits_max_abs': float((lo.cpu() - ro).abs().max()),
            'oracle_logit_range': float(ro.abs().max()), 'salt_fraction': float(mask_ref.mean()),
            'parity_sample': '%d tiles (x2 flips) after %d training steps on synthetic scenes, bf16 engine vs fp32 CPU oracle' % (tiles, train_steps)}


def run_ours(args):
    from salt_b200 import synthetic as synth
    from salt_b200 import _lib
    model = make_model(args.precision, args.batch, args.loss)
    ctx, eng = model.dp, model.engine
    dev = eng.device
    eng.load_state(synth.synth_state_dict(DEPTH, CLASSES, 0))
    B = args.batch
    x_h = torch.from_numpy(synth.synth_inputs(B, SIZE, 1234 + ctx.rank)).pin_memory()
    t_h = torch.from_numpy(synth.synth_targets(B, SIZE, 1234 + ctx.rank)).pin_memory()
    x_d, t_d = x_h.to(dev), t_h.to(dev)

    def timed(fn, steps):
        ctx.barrier(); torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev); ctx.barrier()
        return ctx.max_over_ranks(e0.elapsed_time(e1))

    def step_device():
        return model.train_step_device(x_d, [t_d])

    class LossReader:
        """Minimal callback list for the end-to-end run: reads every step's loss on the host, as the reference's monitors do
        (callbacks.py:115-135 TrainingMonitor)."""
        losses = []

        def on_batch_end(self, metrics=None, *a, **k):
            self.losses.append(float(metrics['sum'].cpu()[0]))       # D2H read of the step's loss

        def training_break(self, *a, **k):
            return False

        def __getattr__(self, name):
            return lambda *a, **k: None

    def run_e2e(k):
        """k training steps through the public entry point, SegmentationModel.fit(datagen=(batches, steps)): every step copies its
        X / target from pinned host memory (prefetched one step ahead on a side stream) and its loss is read back."""
        model.callbacks = LossReader()
        model.fit(datagen=([[x_h, t_h]] * k, k))

    sampler = ClockSampler(ctx.local_rank) if ctx.rank == 0 else None
    for _ in range(args.warmup):
        step_device()
    l0 = _lib.launch_count()
    if sampler: sampler.begin()
    ms = timed(step_device, args.steps)
    if sampler: sampler.end()
    launches = _lib.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    run_e2e(3)
    ms_e2e = timed(lambda: run_e2e(args.steps), 1)

    # per-kernel-class device time of the convolution kernels (CUDA events on the launch stream), 2 extra steps
    eng.profile(True)
    for _ in range(2):
        step_device()
    prof = eng.profile_read()
    prof_groups = eng.profile_read_groups()
    prof_passes = eng.profile_read_passes()
    eng.profile(False)

    # ---- secondary measurements (not the headline): Lovasz-hinge training step (BASELINE configs[2] loss) and fused
    #      sigmoid + h-flip TTA + crop + threshold inference (configs[4]) on the same engine
    extra = {}
    if not args.no_extra:
        def step_lovasz():
            logits = eng.forward(x_d, train=True)
            _, dl = eng.loss_lovasz(logits, t_d)
            eng.backward(dl)
            eng.adam_step(grad_scale=ctx.allreduce_grads(eng.grads))
        x_flip = torch.flip(x_d, dims=[3]).contiguous()

        # BASELINE configs[4]: 512 network inputs = 256 tiles x {orig, h-flip}; the engine takes them as 2 x (128 + 128)
        x2_d = torch.from_numpy(synth.synth_inputs(B, SIZE, 777 + ctx.rank)).to(dev)
        x2_flip = torch.flip(x2_d, dims=[3]).contiguous()

        def infer_tta():
            out = []
            for a, b in ((x_d, x_flip), (x2_d, x2_flip)):
                lo = eng.forward(a, train=False).clone()
                lf = eng.forward(b, train=False)
                out.append(eng.predict(lo, lf, crop=101, threshold=0.5, want_probs=False)[1])
            return out
        # SURVEY.md 8(f) N2 + N3 end to end: raw u8 101x101 tiles in pinned host memory -> H2D (1.3 MB instead of 25 MB) ->
        # adapter fused into the stem (orig + h-flipped tile) -> fused sigmoid/un-flip/mean/crop/threshold -> column-major RLE
        # on the device -> D2H of the run table; and N1: the 21-threshold validation sweep counts for the same batch
        from salt_b200 import io_ops, validation
        tiles_h = torch.from_numpy(synth.synth_tiles_u8(B, 101, 4321 + ctx.rank)).pin_memory()
        gt_d = (t_d[:, 1, 13:114, 14:115] > 0.5).to(torch.uint8).contiguous()
        d2h = {}

        def infer_tiles_rle():
            tl = tiles_h.to(dev, non_blocking=True)
            lo = eng.forward_tiles(tl, train=False)
            lf = eng.forward_tiles(tl, train=False, hflip=True)
            _, mask = eng.predict(lo, lf, crop=101, threshold=0.5, want_probs=False)
            runs, nruns = io_ops.rle_encode_device(mask, cap_runs=256)
            r, n = runs.cpu(), nruns.cpu()
            d2h['bytes'] = r.numel() * 4 + n.numel() * 4
            return r, n

        def validation_sweep():
            lo = eng.forward(x_d, train=False)
            return validation.select_threshold(*[a.cpu().numpy() for a in validation.validation_counts(lo, gt_d)])
        for fn in (step_lovasz, infer_tta, infer_tiles_rle, validation_sweep):
            for _ in range(2):
                fn()
        ms_lv = timed(step_lovasz, 5)
        ms_inf = timed(infer_tta, 5)
        ms_rle = timed(infer_tiles_rle, 5)
        ms_val = timed(validation_sweep, 5)
        se50 = None
        if not args.no_se50:
            se50 = se50_train_step(ctx, timed)
        extra = {'lovasz_train_step': {'value': B * ctx.world * 5 / (ms_lv / 1e3), 'unit': 'images/s', 'ms_per_step': ms_lv / 5},
                 'inference_tta_hflip': {'value': 2 * B * ctx.world * 5 / (ms_inf / 1e3), 'unit': 'tiles/s', 'ms_per_batch': ms_inf / 5,
                                         'network_inputs_per_batch': 4 * B,
                                         'what': '%d tiles per GPU per pass = %d network inputs (orig + h-flip), eval forward with BatchNorm/ReLU/residual folded into the convolution epilogues, fused sigmoid/un-flip/mean/crop/threshold -> u8 masks' % (2 * B, 4 * B)}}

        extra['inference_u8_tiles_to_rle_e2e'] = {
            'value': B * ctx.world * 5 / (ms_rle / 1e3), 'unit': 'tiles/s', 'ms_per_batch': ms_rle / 5,
            'h2d_bytes_per_batch': int(tiles_h.numel()), 'd2h_bytes_per_batch': int(d2h.get('bytes', 0)),
            'what': 'pinned u8 101x101 tiles -> fused adapter+stem (orig + h-flip) -> TTA mean/crop/threshold -> device RLE -> host run table'}
        extra['validation_sweep'] = {
            'value': B * ctx.world * 5 / (ms_val / 1e3), 'unit': 'tiles/s', 'ms_per_batch': ms_val / 5,
            'what': 'eval forward + 21-threshold intersection/prediction counts in one kernel + host IoU/IoUT selection (callbacks.py:499-527)'}
        if se50:
            extra['seresnet50_256_train_step'] = se50
        if ctx.rank == 0 and ctx.world == 1 and not args.no_cpu_baseline:
            extra['inference_tta_hflip'].update(tta_parity(eng, synth, dev))

    if ctx.rank != 0:
        return
    peaks = _peaks()
    n = ctx.world
    value = B * n * args.steps / (ms / 1e3)
    e2e = B * n * args.steps / (ms_e2e / 1e3)
    kern = {k: {'ms_per_step': v[0] / 2, 'tflops': (v[1] / (v[0] * 1e-3) / 1e12 if v[0] > 0 else 0.0), 'launches_per_step': v[2] // 2}
            for k, v in prof.items()}
    tot_ms = sum(v[0] for v in prof.values())
    tot_fl = sum(v[1] for v in prof.values())
    achieved = tot_fl / (tot_ms * 1e-3) / 1e12 if tot_ms > 0 else 0.0
    traffic, traffic_note = None, 'no ncu capture committed'
    tp = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
    if os.path.exists(tp):
        tj = json.load(open(tp))
        if tj.get('conv_source_sha') == conv_source_sha():
            traffic, traffic_note = tj.get('traffic'), tj.get('source')
        else:
            traffic_note = 'stale: profiles/roofline_traffic.json was captured on other kernel sources (%s != %s)' % (tj.get('conv_source_sha'), conv_source_sha())
    per_group = {}
    for gname, d in prof_groups.items():
        g_ms = sum(v[0] for v in d.values()); g_fl = sum(v[1] for v in d.values())
        if g_ms > 0:
            per_group[gname] = {'ms_per_step': g_ms / 2, 'tflops': g_fl / (g_ms * 1e-3) / 1e12, 'frac': g_fl / (g_ms * 1e-3) / 1e12 / peaks['tflops'],
                                'fwd_tflops': d['conv_fwd'][1] / (d['conv_fwd'][0] * 1e-3) / 1e12 if d['conv_fwd'][0] > 0 else None}
    line = {
        'metric': 'images/sec', 'value': value, 'unit': 'images/s', 'n_gpus': n, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': args.precision, 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'global_batch': B * n, 'loss': args.loss, 'parallelism': 'dp%d' % n,
                   'cuda_graphs': os.environ.get('SALT_ENGINE_GRAPH', '1') != '0',
                   'l2': 'per-step working set (saved activations of %d images, several GB) is far larger than the 126 MB L2; no explicit flush' % B},
        'clocks': clocks,
        'e2e': {'value': e2e, 'unit': 'images/s', 'ms_per_step': ms_e2e / args.steps,
                'h2d_bytes_per_step': int(x_h.numel() * 4 + t_h.numel() * 4), 'd2h_bytes_per_step': 4,
                'api': 'salt_b200.models.SegmentationModel.fit(datagen=(batches, steps)) - pinned host tensors in, H2D of every step inside the timed region, loss read back every step'},
        'gpu_launches': int(launches),
        'roofline': {'bound': 'tensor', 'achieved': achieved, 'peak': peaks['tflops'], 'unit': 'TFLOP/s',
                     'frac': achieved / peaks['tflops'], 'traffic': traffic, 'traffic_source': traffic_note, 'peak_source': peaks['source'],
                     'per_group': per_group,
                     'memory_bound_passes': {k: {'ms_per_step': v[0] / 2, 'algorithmic_GBps': (v[1] / (v[0] * 1e-3) / 1e9 if v[0] > 0 else 0.0),
                                                 'launches_per_step': v[2] // 2} for k, v in prof_passes.items()},
                     'kernel': 'implicit-GEMM convolution (forward + dgrad + wgrad launches of one step, algorithmic FLOPs / CUDA-event time)',
                     'conv_share_of_step': (tot_ms / 2) / (ms / args.steps), 'per_class': kern,
                     'step_tflops': value / n * TRAIN_GFLOP_PER_IMAGE / 1e3,
                     'step_frac_of_peak': value / n * TRAIN_GFLOP_PER_IMAGE / 1e3 / peaks['tflops']},
    }
    if extra:
        line['extra'] = extra
    if n == 1 and not args.no_cpu_baseline:
        ips, sec_step, threads = cpu_train_steps(2, 1, CPU_SAMPLE_BATCH)
        line['cpu_baseline'] = {'value': ips, 'unit': 'images/s', 'cores': threads, 'kind': 'port',
                                'sample': '%d-image sample of the batch, 2 timed training steps of the oracle PORT (PyTorch fp32 CPU); '
                                          '--impl reference times the full 128-image batch' % CPU_SAMPLE_BATCH}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=BATCH_PER_GPU, help='images per GPU')
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--loss', default='bce_dice', choices=['bce_dice', 'lovasz'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-se50', action='store_true', help='skip the secondary UNetSeResNet-50 256x256 training-step measurement')
    ap.add_argument('--no-extra', action='store_true', help='skip the secondary Lovasz / TTA-inference measurements')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()

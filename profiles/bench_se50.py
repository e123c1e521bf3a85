"""UNetSeResNet-50 training step at BASELINE config 4 shape (256x256 network input, bf16, 64 images): step time and
per-class convolution throughput.  Usage: python profiles/bench_se50.py [batch] [size] [loss]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'open-solution-salt-identification_b200'))
from salt_b200 import synthetic as synth          # noqa: E402
from salt_b200.engine import UNetEngine           # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
S = int(sys.argv[2]) if len(sys.argv) > 2 else 256
LOSS = sys.argv[3] if len(sys.argv) > 3 else 'lovasz'
t0 = time.time()
eng = UNetEngine(architecture='UNetSeResNet', encoder_depth=50, num_classes=2, max_batch=B, size=S, precision='bf16')
eng.load_state(synth.synth_state_dict(50, 2, 0))
x = torch.from_numpy(synth.synth_inputs(B, S, 1)).cuda()
t = torch.from_numpy(synth.synth_targets(B, S, 1)).cuda()
print('setup %.1f s, params %.1f M, workspace %.1f GB' % (time.time() - t0, eng.params.numel() / 1e6, eng.workspace.numel() / 2**30))


def step():
    logits = eng.forward(x, train=True)
    loss, dl = (eng.loss_lovasz if LOSS == 'lovasz' else eng.loss_bce_dice)(logits, t)
    eng.backward(dl)
    eng.adam_step()
    return loss


if os.environ.get('PROFILE_ONE'):       # ncu --profile-from-start off: capture exactly one warm step
    step()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    sys.exit(0)
for _ in range(3):
    loss = step()
torch.cuda.synchronize()
print('loss', float(loss.cpu()[0]))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
N = 5
for _ in range(N):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / N
print('step %.2f ms  = %.1f images/s' % (ms, B / ms * 1e3))
eng.profile(True)
step()
prof = eng.profile_read()
eng.profile(False)
tot_ms = sum(v[0] for v in prof.values())
tot_fl = sum(v[1] for v in prof.values())
for k, v in prof.items():
    print('%-12s %8.2f ms  %7.1f TFLOP/s  %d launches' % (k, v[0], v[1] / (v[0] * 1e-3) / 1e12 if v[0] else 0, v[2]))
print('conv total %.2f ms (%.0f %% of step), %.1f TFLOP/s; %.1f GFLOP/image' % (tot_ms, 100 * tot_ms / ms, tot_fl / (tot_ms * 1e-3) / 1e12, tot_fl / B / 1e9))

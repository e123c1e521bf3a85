"""Per-launch timeline of one eager training step (BASELINE configs[1]: UNetResNet-34, 128x128, bf16, B=128, BCE+Dice): every
convolution / memory-bound pass bracketed by CUDA events on the launch stream (salt_profile_records), in launch order, with
TFLOP/s or algorithmic GB/s per launch.  usage: python profiles/step_timeline.py [> profiles/r2_step_timeline.txt]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'open-solution-salt-identification_b200'))
import collections
import torch
from salt_b200 import synthetic as synth
from salt_b200.engine import UNetEngine
B, S = 128, 128
eng = UNetEngine(34, 2, B, S, precision='bf16')
eng.load_state(synth.synth_state_dict(34, 2, 0))
x = torch.from_numpy(synth.synth_inputs(B, S, 1)).cuda()
t = torch.from_numpy(synth.synth_targets(B, S, 1)).cuda()
def step():
    logits = eng.forward(x, train=True)
    loss, dl = eng.loss_bce_dice(logits, t)
    eng.backward(dl)
    eng.adam_step()
for _ in range(3): step()
torch.cuda.synchronize()
eng.profile(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record()
torch.cuda.synchronize()
recs = eng.profile_records()
eng.profile(False)
tot = sum(r[3] for r in recs)
print('# step (eager, with events) %.3f ms; bracketed launches %d, sum of bracketed times %.3f ms' % (e0.elapsed_time(e1), len(recs), tot))
agg = collections.OrderedDict()
for i, (cls, grp, work, ms) in enumerate(recs):
    rate = work / ms / 1e9 if ms > 0 else 0.0           # conv: TFLOP/s * 1e3 ... printed per unit below
    unit = 'TFLOP/s' if cls.startswith('conv') else 'GB/s'
    val = work / ms / 1e9 if cls.startswith('conv') else work / ms / 1e6
    print('%4d %-14s %-7s %9.1f us  %9.1f %s  (work %.3e)' % (i, cls, grp, ms * 1e3, val, unit, work))
    a = agg.setdefault((cls, grp), [0, 0.0, 0.0]); a[0] += 1; a[1] += ms; a[2] += work
print('# per (class, group)')
for (cls, grp), (n, ms, work) in agg.items():
    val = work / ms / 1e9 if cls.startswith('conv') else work / ms / 1e6
    print('# %-14s %-7s n=%3d %8.3f ms  %9.1f %s' % (cls, grp, n, ms, val, 'TFLOP/s' if cls.startswith('conv') else 'GB/s'))

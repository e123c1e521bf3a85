"""Two eager training steps of the bench workload (UNetResNet-34, 128x128, bf16, B=128, BCE+Dice) - the process ncu attaches to for
the launch list (profiles/r2_launches_final.md) and the --set full captures.  usage: python profiles/one_step.py [steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'open-solution-salt-identification_b200'))
import torch
from salt_b200 import synthetic as synth
from salt_b200.engine import UNetEngine
B, S = 128, 128
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
eng = UNetEngine(34, 2, B, S, precision='bf16')
eng.load_state(synth.synth_state_dict(34, 2, 0))
x = torch.from_numpy(synth.synth_inputs(B, S, 1)).cuda()
t = torch.from_numpy(synth.synth_targets(B, S, 1)).cuda()
for _ in range(steps):
    logits = eng.forward(x, train=True)
    loss, dl = eng.loss_bce_dice(logits, t)
    eng.backward(dl)
    eng.adam_step()
torch.cuda.synchronize()
print('loss', float(loss))

"""One bf16 training step + TTA prediction of UNetResNet-18 (2x3x64x64) - small enough to run under compute-sanitizer."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'open-solution-salt-identification_b200'))
import torch
from salt_b200 import synthetic as synth
from salt_b200.engine import UNetEngine
for prec in ('bf16', 'fp32'):
    eng = UNetEngine(18, 2, 3, 64, precision=prec)
    eng.load_state(synth.synth_state_dict(18, 2, 0))
    x = torch.from_numpy(synth.synth_inputs(3, 64, 1)).cuda(); t = torch.from_numpy(synth.synth_targets(3, 64, 1)).cuda()
    for loss in ('lovasz', 'bce_dice'):
        lg = eng.forward(x, train=True)
        l, dl = (eng.loss_lovasz if loss == 'lovasz' else eng.loss_bce_dice)(lg, t)
        eng.backward(dl); eng.adam_step()
    le = eng.forward(x, train=False); lf = eng.forward(torch.flip(x, dims=[3]).contiguous(), train=False)
    p, m = eng.predict(le, lf, crop=50)
    torch.cuda.synchronize()
    print(prec, 'ok', float(l), int(m.sum()))

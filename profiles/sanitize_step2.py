"""More shapes for compute-sanitizer: UNetResNet-34 at 128x128 (every layer shape of the benchmark: N = 32 tiles, 8x8 maps, f = 16
upsample adjoints, halo weight gradients of both forms) and UNetSeResNetXt-50 at 64x64 (grouped convolutions), one bf16 training
step each + an eval forward."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'open-solution-salt-identification_b200'))
import torch
from salt_b200 import synthetic as synth
from salt_b200.engine import UNetEngine
for depth, arch, b, s in ((34, None, 16, 128), (50, 'UNetSeResNetXt', 2, 64)):
    eng = UNetEngine(depth, 2, b, s, precision='bf16', architecture=arch)
    eng.load_state(synth.synth_state_dict(depth, 2, 0, arch))
    x = torch.from_numpy(synth.synth_inputs(b, s, 1)).cuda(); t = torch.from_numpy(synth.synth_targets(b, s, 1)).cuda()
    lg = eng.forward(x, train=True)
    l, dl = eng.loss_bce_dice(lg, t)
    eng.backward(dl); eng.adam_step()
    le = eng.forward(x, train=False)
    torch.cuda.synchronize()
    print(depth, arch, 'ok', float(l), float(le.abs().max()))

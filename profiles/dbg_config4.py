"""Gradient cosines of every parameter, UNetSeResNet-50 / UNetResNet-34 fp32 vs the CPU oracle (debug aid).
usage: python profiles/dbg_config4.py [depth] [size]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'open-solution-salt-identification_b200'))
import torch
from oracle import synth, unet_oracle, losses_oracle
from salt_b200.engine import UNetEngine
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 50
s = int(sys.argv[2]) if len(sys.argv) > 2 else 256
b = 2
sd_np = synth.synth_state_dict(depth, 2, 0)
x = torch.from_numpy(synth.synth_inputs(b, s, 11)); t = torch.from_numpy(synth.synth_targets(b, s, 11))
sd = unet_oracle.to_torch_state(sd_np, requires_grad=True)
ref = unet_oracle.unet_resnet_forward(sd, x, depth, train=True)
loss_ref = losses_oracle.lovasz_hinge_per_image(ref, t)
loss_ref.backward()
eng = UNetEngine(depth, 2, b, s, precision='fp32')
eng.load_state(sd_np)
logits = eng.forward(x.cuda(), train=True)
loss, dl = eng.loss_lovasz(logits, t.cuda())
eng.backward(dl)
torch.cuda.synchronize()
print('logits err %.3e' % (logits.cpu() - ref.detach()).abs().max().item())
bad = 0
for k, v in sd.items():
    if v.grad is None or v.grad.norm().item() == 0: continue
    a = eng.view(k, grad=True).cpu().flatten().double(); r = v.grad.flatten().double()
    c = (a @ r / (a.norm() * r.norm() + 1e-300)).item()
    if c < 0.9995:
        bad += 1
        print('%-60s cos %.5f  |a|/|r| %.4f' % (k, c, (a.norm() / r.norm()).item()))
print('keys below 0.9995:', bad)

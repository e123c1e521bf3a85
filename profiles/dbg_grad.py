import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/open-solution-salt-identification_b200')
import torch, numpy as np
from oracle import synth, unet_oracle, losses_oracle
from salt_b200.engine import UNetEngine
depth,b,s=18,2,64
sd_np=synth.synth_state_dict(depth,2,0)
x=torch.from_numpy(synth.synth_inputs(b,s,1234)); t=torch.from_numpy(synth.synth_targets(b,s,1234))
sd=unet_oracle.to_torch_state(sd_np, requires_grad=True)
ref,stages=unet_oracle.unet_resnet_forward(sd,x,depth,True,return_stages=True)
for v in stages.values(): v.retain_grad()
loss=losses_oracle.bce_dice(ref,t); loss.backward()
eng=UNetEngine(depth,2,b,s,precision='fp32'); eng.load_state(sd_np)
lg=eng.forward(x.cuda(),train=True); l,dl=eng.loss_bce_dice(lg,t.cuda()); eng.backward(dl); torch.cuda.synchronize()
for name in ['d1','d2','d5']:
    g=eng.activation('g_'+name).cpu(); r=stages[name].grad
    e=(g-r)
    print(name,'max err',e.abs().max().item(),'mean err',e.mean().item(),'mean abs err',e.abs().mean().item(),'ref mean abs',r.abs().mean().item())
    # per-channel mean error
    print('  per-channel mean err (first 8):',[float('%.2e'%v) for v in e.mean(dim=(0,2,3))[:8]])
    print('  border rows err', e[:,:,0,:].abs().mean().item(), e[:,:,-1,:].abs().mean().item(), 'cols', e[:,:,:,0].abs().mean().item(), e[:,:,:,-1].abs().mean().item(), 'interior', e[:,:,8:-8,8:-8].abs().mean().item())
st_e=eng.activation('stem').cpu()
import torch.nn.functional as F
with torch.no_grad():
    sdd={k:v.detach() for k,v in unet_oracle.to_torch_state(sd_np).items()}
    y=F.conv2d(x, sdd['encoders.encoder.conv1.weight'], None, stride=2, padding=3)
    y=F.relu(F.batch_norm(y, None, None, sdd['encoders.encoder.bn1.weight'], sdd['encoders.encoder.bn1.bias'], True, 0.1, 1e-5))
print('stem act err', (st_e-y).abs().max().item())
for k in ['final.0.conv.weight','final.0.batch_norm.bias','encoders.encoder.conv1.weight']:
    a=eng.view(k,grad=True).cpu(); r=sd[k].grad
    print(k,'grad err',(a-r).abs().max().item(),'ref max',r.abs().max().item())

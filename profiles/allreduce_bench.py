"""All-reduce of the flat fp32 gradient buffer of UNetResNet-34 (30 178 700 floats = 120.7 MB), NCCL over NVLink / NVSwitch, timed with
CUDA events on the device (max over ranks).  usage: torchrun --nproc-per-node N profiles/allreduce_bench.py  (env NCCL_* selects variants)"""
import os
import torch
import torch.distributed as dist
rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE']); local = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
n = 30178700
g = torch.randn(n, device='cuda')
for _ in range(5): dist.all_reduce(g)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = 30
e0.record()
for _ in range(iters): dist.all_reduce(g)
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / iters], device='cuda')
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    t = ms.item()
    print('world %d  %s  all-reduce of %.1f MB: %.3f ms  (algbw %.0f GB/s, busbw %.0f GB/s)' % (
        world, ' '.join('%s=%s' % (k, v) for k, v in sorted(os.environ.items()) if k.startswith('NCCL_')) or 'defaults',
        n * 4 / 1e6, t, n * 4 / t / 1e6, n * 4 / t / 1e6 * 2 * (world - 1) / world))
dist.destroy_process_group()

"""Data-parallel plumbing: one process per GPU, full replica, local minibatch shard, ONE all-reduce (SUM) of the
flat fp32 gradient buffer per step, scaled by 1/world inside the fused Adam kernel.

Replaces the reference's single-process ``nn.DataParallel`` (common_blocks/models.py:81-85): like it, BatchNorm
statistics stay per replica and nothing but gradients (and the 3*C+1 Dice/BCE partial sums) is communicated.
``torch.distributed`` (NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU tests) is used for the collective.
"""
import os

import torch


class DataParallelContext:
    def __init__(self, rank=0, world=1, local_rank=0, group=None, device=None):
        self.rank, self.world, self.local_rank, self.group, self.device = rank, world, local_rank, group, device

    @classmethod
    def from_env(cls):
        """Single process unless launched by torchrun (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* in the env)."""
        world = int(os.environ.get('WORLD_SIZE', '1'))
        rank = int(os.environ.get('RANK', '0'))
        local_rank = int(os.environ.get('LOCAL_RANK', '0'))
        device = None
        if torch.cuda.is_available():
            torch.cuda.set_device(local_rank % torch.cuda.device_count())
            device = torch.device('cuda', local_rank % torch.cuda.device_count())
        group = None
        if world > 1:
            import torch.distributed as dist
            if not dist.is_initialized():
                os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
                os.environ.setdefault('MASTER_PORT', '29500')
                kw = {}
                if torch.cuda.is_available():
                    kw['device_id'] = device
                dist.init_process_group(backend='nccl' if torch.cuda.is_available() else 'gloo', rank=rank,
                                        world_size=world, **kw)
            group = dist.group.WORLD
        return cls(rank, world, local_rank, group, device)

    def shard(self, n):
        """[lo, hi) slice of a global batch of n images owned by this rank (n must divide evenly: every rank
        contributes the same number of images, as the reference's scatter does for full batches)."""
        if n % self.world:
            raise ValueError('global batch %d is not divisible by world size %d' % (n, self.world))
        per = n // self.world
        return self.rank * per, (self.rank + 1) * per

    def broadcast(self, *tensors):
        if self.world > 1:
            import torch.distributed as dist
            for t in tensors:
                dist.broadcast(t, src=0, group=self.group)

    def allreduce_grads(self, flat_grads):
        """In-place SUM all-reduce of the flat gradient buffer.  Returns the factor the optimiser must apply
        (1/world) - fused into the Adam kernel instead of a separate scaling pass."""
        if self.world == 1:
            return 1.0
        import torch.distributed as dist
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=self.group)
        return 1.0 / self.world

    def allreduce_async(self, tensor):
        """Start an in-place SUM all-reduce of one gradient bucket and return its handle (``.wait()`` makes the current stream wait).
        NCCL runs it on the process group's own stream, ordered after the work already queued on the current stream, so the
        kernels launched next (the following backward segment) overlap it."""
        if self.world == 1:
            return None
        import torch.distributed as dist
        return dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def gather_rows(self, *arrays):
        """Concatenate per-rank numpy arrays along axis 0 in rank order on every rank (validation counts: the images of a
        validation epoch are sharded across ranks, the threshold sweep needs all of them; a few KB, once per epoch)."""
        if self.world == 1:
            return arrays
        import numpy as np
        import torch.distributed as dist
        parts = [None] * self.world
        dist.all_gather_object(parts, arrays, group=self.group)
        return tuple(np.concatenate([p[i] for p in parts], axis=0) for i in range(len(arrays)))

    def max_over_ranks(self, value):
        if self.world == 1:
            return float(value)
        import torch.distributed as dist
        t = torch.tensor([float(value)], dtype=torch.float64, device=self.device if self.device is not None else 'cpu')
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return float(t.item())

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier(group=self.group)

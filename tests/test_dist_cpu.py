"""CPU (gloo, world_size 2): the data-parallel plumbing of salt_b200/dist.py - rendezvous from the torchrun
environment, batch sharding, the single SUM all-reduce of the flat gradient buffer with the 1/world factor handed
to the optimiser, parameter broadcast, max-over-ranks timing reduction."""
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import os, sys, json
    sys.path.insert(0, os.path.join(%r, 'open-solution-salt-identification_b200'))
    import torch
    from salt_b200.dist import DataParallelContext
    ctx = DataParallelContext.from_env()
    assert ctx.world == 2 and ctx.rank == int(os.environ['RANK'])
    lo, hi = ctx.shard(256)
    grads = torch.full((1000,), float(ctx.rank + 1))          # rank0: 1, rank1: 2
    scale = ctx.allreduce_grads(grads)                        # SUM -> 3, scale 1/2
    params = torch.arange(10, dtype=torch.float32) * (1 if ctx.rank == 0 else -1)
    ctx.broadcast(params)
    mx = ctx.max_over_ranks(10.0 + ctx.rank)
    ctx.barrier()
    try:
        ctx.shard(255)
        bad = False
    except ValueError:
        bad = True
    # validation epoch sharded over the ranks: each scores its half, the sweep sees all images (rank order)
    import numpy as np
    from salt_b200 import validation
    sys.path.insert(0, %r)
    from oracle import io_oracle
    from oracle.make_golden import io_inputs
    inp = io_inputs()
    lo6, hi6 = ctx.shard(6)
    counts = io_oracle.validation_counts(inp['logits'][lo6:hi6], inp['y_true'][lo6:hi6])
    sc = validation.ValidationScorer(dp=ctx)
    sc._parts.append(tuple(torch.from_numpy(np.ascontiguousarray(c)) for c in counts))
    res = sc.result()
    print(json.dumps(dict(rank=ctx.rank, shard=[lo, hi], gsum=float(grads[0]), gmean=float(grads[0] * scale),
                          p1=float(params[1]), mx=mx, bad=bad, thr=res['threshold'], iout=res['iout'], iou=res['iou'])))
    torch.distributed.destroy_process_group()
''') % (ROOT, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_data_parallel_context_gloo_world2(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(WORKER)
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE='2', MASTER_ADDR='127.0.0.1',
                   MASTER_PORT=str(port), CUDA_VISIBLE_DEVICES='')
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = []
    for p in procs:
        out, err = p.communicate(timeout=180)
        assert p.returncode == 0, err[-2000:]
        outs.append(__import__('json').loads(out.strip().splitlines()[-1]))
    outs.sort(key=lambda d: d['rank'])
    assert outs[0]['shard'] == [0, 128] and outs[1]['shard'] == [128, 256]
    for d in outs:
        assert d['gsum'] == 3.0 and d['gmean'] == 1.5        # SUM all-reduce, 1/world applied by the optimiser
        assert d['p1'] == 1.0                                 # rank 0's parameters everywhere
        assert d['mx'] == 11.0                                # max over ranks (timing reduction)
        assert d['bad']                                       # uneven global batches are rejected
    import numpy as np
    gold = np.load(os.path.join(ROOT, 'tests', 'golden', 'io_cases.npz'))
    for d in outs:                                            # both ranks select the reference's threshold / scores
        assert d['thr'] == float(gold['val_threshold'])
        assert abs(d['iout'] - float(gold['val_iout'])) < 1e-12 and abs(d['iou'] - float(gold['val_iou'])) < 1e-12


def test_single_process_context_is_identity():
    sys.path.insert(0, os.path.join(ROOT, 'open-solution-salt-identification_b200'))
    import torch
    from salt_b200.dist import DataParallelContext
    ctx = DataParallelContext(0, 1, 0, None, None)
    g = torch.ones(4)
    assert ctx.allreduce_grads(g) == 1.0 and ctx.shard(128) == (0, 128) and ctx.max_over_ranks(3.5) == 3.5


def test_fit_prefetch_order_and_termination():
    """Host logic of SegmentationModel._prefetch (no CUDA): batch i+1 is staged BEFORE batch i is handed to the training step,
    every batch is staged exactly once and in order, and the generator ends with the loader."""
    sys.path.insert(0, os.path.join(ROOT, 'open-solution-salt-identification_b200'))
    from salt_b200.models import SegmentationModel
    log = []

    class Fake:
        def _stage(self, data):
            if data is not None:
                log.append(('stage', data))
            return data
    fake = Fake()
    out = []
    for item in SegmentationModel._prefetch(fake, iter(range(4))):
        log.append(('step', item))
        out.append(item)
    assert out == [0, 1, 2, 3]
    assert log == [('stage', 0), ('stage', 1), ('step', 0), ('stage', 2), ('step', 1), ('stage', 3), ('step', 2), ('step', 3)]
    assert list(SegmentationModel._prefetch(fake, iter(()))) == []

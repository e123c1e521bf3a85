"""Micro-benchmark of single convolution launches through the C ABI (CUDA-event timing, 20 reps after 3 warm-ups).
usage: python profiles/microbench_conv.py   (env SALT_TC_ROWS / SALT_TC_DEBUG select kernel variants / timing experiments)"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'open-solution-salt-identification_b200'))
import torch
from salt_b200 import _lib
lib = _lib.load()
CASES = [('layer1 64->64 @64', 128, 64, 64, 64, 64, 3, 1, 1), ('final.0 320->64 @128(+2)', 128, 320, 64, 130, 130, 3, 1, 0),
         ('layer2 128->128 @32', 128, 128, 128, 32, 32, 3, 1, 1), ('layer3 256->256 @16', 128, 256, 256, 16, 16, 3, 1, 1),
         ('dec2.c1 128->64 @64(+2)', 128, 128, 64, 66, 66, 3, 1, 0)]
for name, B, Ci, Co, H, W, k, s, p in CASES:
    Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    x = torch.randn(B, H, W, Ci, device='cuda').bfloat16()
    w = torch.randn(Co, Ci, k, k, device='cuda') * 0.05
    out = torch.empty(B, Ho, Wo, Co, device='cuda', dtype=torch.bfloat16)
    stats = torch.zeros(2 * Co, dtype=torch.float64, device='cuda')
    d = _lib.SaltConvDesc(B, H, W, Ci, Ho, Wo, Co, k, s, p, 1, 1)
    def run():
        _lib.check(lib.salt_op_conv_forward(C.byref(d), x.data_ptr(), w.data_ptr(), None, out.data_ptr(), stats.data_ptr(), None))
    for _ in range(1): run()
    # the op entry point packs weights and synchronises; time the kernel itself with the profiler-free event pair around many calls
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(1):
        torch.cuda.synchronize(); e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    t = min(ts)
    fl = 2.0 * B * Ho * Wo * Co * Ci * k * k
    print('%-28s %8.3f ms (incl. weight pack)  %7.0f TFLOP/s' % (name, t, fl / t / 1e9))

"""Pipeline-stall accounting of conv_tc_rows_kernel (library built with `python open-solution-salt-identification_b200/build.py --timing`):
for the layer shapes of the benchmark, per CTA averages of
  issuer   total cycles, cycles waiting for TMA (full barriers), cycles waiting for the epilogue (tmem_empty), stages issued
  producer cycles waiting for a free stage (empty barriers)
  epilogue cycles waiting for a finished accumulator
usage: SALT_LIB_PATH=open-solution-salt-identification_b200/libsaltunet_timing.so python profiles/rows_timing.py"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'open-solution-salt-identification_b200'))
import numpy as np
import torch
from salt_b200 import _lib
lib = _lib.load()
lib.salt_debug_rows_timing.argtypes = [C.c_void_p, C.c_int]
CASES = [('layer1 64->64 @64', 128, 64, 64, 64, 64), ('final.0 320->64 @128(+2)', 128, 320, 64, 130, 130),
         ('layer2 128->128 @32', 128, 128, 128, 32, 32), ('layer3 256->256 @16', 128, 256, 256, 16, 16),
         ('dec2.c1 128->64 @64(+2)', 128, 128, 64, 66, 66), ('dgrad final.0 64->320 @128', 128, 64, 320, 128, 128)]
buf = np.zeros((148, 8), dtype=np.uint64)
for name, B, Ci, Co, H, W in CASES:
    p = 1 if H in (64, 32, 16, 128) else 0
    Ho, Wo = H + 2 * p - 2, W + 2 * p - 2
    x = torch.randn(B, H, W, Ci, device='cuda').bfloat16()
    w = torch.randn(Co, Ci, 3, 3, device='cuda') * 0.05
    out = torch.empty(B, Ho, Wo, Co, device='cuda', dtype=torch.bfloat16)
    stats = torch.zeros(2 * Co, dtype=torch.float64, device='cuda')
    d = _lib.SaltConvDesc(B, H, W, Ci, Ho, Wo, Co, 3, 1, p, 1, 1)
    for rep in range(3):
        lib.salt_debug_rows_timing(None, 1)
        _lib.check(lib.salt_op_conv_forward(C.byref(d), x.data_ptr(), w.data_ptr(), None, out.data_ptr(), stats.data_ptr(), None))
        lib.salt_debug_rows_timing(buf.ctypes.data, 0)
    a = buf[buf[:, 0] > 0].astype(np.float64)
    tot, full, tempty, stages, pwait, ptot, ewait = [a[:, i].mean() for i in range(7)]
    mmas = stages * 12
    print('%-28s CTAs %3d | issuer %8.0f clk = %5.1f clk/MMA | wait TMA %4.1f %% | wait epilogue %4.1f %% | producer waits for a free stage %4.1f %% of %8.0f clk'
          ' | epilogue waits for an accumulator %4.1f %%' % (name, len(a), tot, tot / mmas, 100 * full / tot, 100 * tempty / tot, 100 * pwait / ptot, ptot,
                                                         100 * ewait / tot))

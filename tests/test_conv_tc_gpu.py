"""GPU: the tcgen05 implicit-GEMM convolution (forward + stride-1 dgrad) against torch.nn.functional.conv2d
(fp32, on bf16-rounded operands).  Tolerance: bf16 output rounding (2^-9 relative) + fp32 accumulation order:
max-abs <= 2e-2 + 1e-2 * max|ref|.  BatchNorm statistics (fp32 accumulators, pre-rounding): 2e-3 relative."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from test_engine_gpu import report, _to_nhwc, _from_nhwc   # noqa: E402

# (B, Cin, Cout, H, W, k, stride, pad)   H, W are the physical input dims
TC_CASES = [
    (2, 64, 64, 16, 16, 3, 1, 1),        # layer1-like, one tile per image pair
    (4, 64, 64, 64, 64, 3, 1, 1),        # many tiles, persistent loop + double-buffered accumulators
    (3, 64, 128, 32, 32, 3, 2, 1),       # stride 2 via TMA element strides, odd batch
    (2, 64, 128, 32, 32, 1, 2, 0),       # 1x1 stride-2 downsample
    (2, 128, 256, 16, 16, 3, 1, 1),      # BN = 256 -> or halved by the occupancy heuristic
    (5, 512, 512, 8, 8, 3, 1, 1),        # 8x8 maps: two images per tile, odd batch, 2 channel tiles
    (2, 96, 32, 18, 34, 3, 1, 0),        # valid conv on a bordered tensor, Cin = 96 (BK = 32), N = 32
    (2, 32, 64, 34, 34, 3, 1, 0),        # dec1.conv2-like: Cin = 32
    (2, 320, 64, 18, 18, 3, 1, 0),       # final.0-like: Cin = 320
    (1, 768, 512, 10, 10, 3, 1, 0),      # dec5.conv1-like
    # enough pixel tiles for the thread-block-cluster (weight multicast) variant of the row-halo kernel
    (7, 64, 64, 32, 24, 3, 1, 1),        # 42 pixel tiles: the last cluster group is short (dropped dummy tiles), ragged width
    (6, 128, 128, 32, 32, 3, 1, 1),      # BN = 128, two 64-channel blocks per tap row
    (5, 128, 256, 32, 32, 3, 1, 1),      # two channel tiles per pixel tile
    (5, 320, 64, 34, 66, 3, 1, 0),       # final.0-like at cluster size: forward 5 channel blocks, dgrad 5 channel tiles of 64
    (4, 64, 32, 66, 66, 3, 1, 0),        # N = 32 (8-row multicast parts), bordered input
    # the tap-table kernel at sizes with many position tiles
    (8, 64, 128, 64, 64, 3, 2, 1),       # stride 2 forward + 4-phase stride-2 dgrad, 64 position tiles
    (8, 64, 128, 64, 64, 1, 2, 0),       # 1x1 stride-2
    (70, 512, 512, 8, 8, 3, 1, 1),       # 8x8 maps, two images per tile: 35 position tiles (short last group), 2-4 channel tiles
    (4, 32, 64, 66, 66, 3, 1, 0),        # Cin = 32: 64-byte swizzle, 512-byte multicast parts
    (4, 192, 128, 34, 34, 3, 1, 0),      # dec3.conv1-like: dgrad with 192 output channels = ONE wide N = 192 tile (SALT_TC_WIDE=0: 3 x 64)
    (40, 64, 64, 64, 64, 3, 1, 1),       # layer1 shape with 1280 pixel tiles: > 2 groups per SM
]


@pytest.mark.parametrize('case', TC_CASES)
def test_conv_tc_forward_and_dgrad(case):
    from salt_b200 import _lib
    lib = _lib.load()
    B, Ci, Co, H, W, k, s, p = case
    g = torch.Generator().manual_seed(sum(case))
    x = (torch.randn(B, Ci, H, W, generator=g)).bfloat16().float()
    w = (torch.randn(Co, Ci, k, k, generator=g) * (2.0 / (Ci * k * k)) ** 0.5).bfloat16().float()
    bias = torch.randn(Co, generator=g) * 0.1
    y_ref = F.conv2d(x, w, bias, stride=s, padding=p)
    Ho, Wo = y_ref.shape[2:]
    d = _lib.SaltConvDesc(B, H, W, Ci, Ho, Wo, Co, k, s, p, 1, 1)
    xd, wd, bd = _to_nhwc(x, 'bf16'), w.cuda().contiguous(), bias.cuda()
    out = torch.full((B, Ho, Wo, Co), float('nan'), dtype=torch.bfloat16, device='cuda')
    stats = torch.zeros(2 * Co, dtype=torch.float64, device='cuda')
    cl0 = lib.salt_cluster_launch_count()
    _lib.check(lib.salt_op_conv_forward(C.byref(d), xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), out.data_ptr(), stats.data_ptr(), None))
    torch.cuda.synchronize()
    import os
    m_tiles = B * ((Ho + 15) // 16) * ((Wo + 7) // 8)
    resident = Ci <= 64 and Co <= 64 and os.environ.get('SALT_TC_RESW', '1') != '0'      # weights stay in shared memory: one CTA per tile
    if (k == 3 and s == 1 and Ci % 64 == 0 and Co % 32 == 0 and Ho >= 16 and m_tiles >= 32 and not resident
            and os.environ.get('SALT_TC_CLUSTER', '2') != '1'):
        # large enough for the thread-block-cluster variant: make sure THAT kernel (TMA-multicast weights) produced `out`
        assert lib.salt_cluster_launch_count() > cl0, 'cluster variant of the row-halo convolution was not launched'
    oks = [report('tc conv fwd %s' % (case,), _from_nhwc(out), y_ref, atol=2e-2, rtol=1e-2)[0]]
    oks.append(report('tc conv stats sum', stats[:Co].cpu().float(), y_ref.sum((0, 2, 3)), atol=5e-2, rtol=2e-3)[0])
    oks.append(report('tc conv stats sumsq', stats[Co:].cpu().float(), (y_ref ** 2).sum((0, 2, 3)), atol=5e-2, rtol=2e-3)[0])
    gy = torch.randn(B, Co, Ho, Wo, generator=g).bfloat16().float()
    xr = x.clone().requires_grad_(True)
    F.conv2d(xr, w, None, stride=s, padding=p).backward(gy)
    gyd = _to_nhwc(gy, 'bf16')
    if s == 1 or k > 1:
        gin = torch.full((B, H, W, Ci), float('nan'), dtype=torch.bfloat16, device='cuda')
        _lib.check(lib.salt_op_conv_dgrad(C.byref(d), gyd.data_ptr(), wd.data_ptr(), gin.data_ptr(), 0, None))
        oks.append(report('tc conv dgrad (stride %d)' % s, _from_nhwc(gin), xr.grad, atol=2e-2, rtol=1e-2)[0])
        _lib.check(lib.salt_op_conv_dgrad(C.byref(d), gyd.data_ptr(), wd.data_ptr(), gin.data_ptr(), 1, None))
        oks.append(report('tc conv dgrad accumulate', _from_nhwc(gin), 2 * xr.grad, atol=4e-2, rtol=2e-2)[0])
    else:
        # 1x1 stride-2: only the even/even lattice receives gradient -> the kernel is accumulate-only
        base = torch.randn(B, Ci, H, W, generator=g).bfloat16().float()
        gin = _to_nhwc(base, 'bf16').clone()
        _lib.check(lib.salt_op_conv_dgrad(C.byref(d), gyd.data_ptr(), wd.data_ptr(), gin.data_ptr(), 1, None))
        oks.append(report('tc conv dgrad 1x1 s2 accumulate', _from_nhwc(gin), base + xr.grad, atol=4e-2, rtol=2e-2)[0])
    assert all(oks)


@pytest.mark.parametrize('case', [TC_CASES[i] for i in (0, 1, 2, 3, 4, 5, 6, 7, 8, 12, 13, 14, 18)])
def test_conv_tc_split_fp32(case):
    """fp32 parity mode ON THE TENSOR CORES: fp32 operands as three bf16 terms each, six bf16 x bf16 products accumulated in
    fp32 by tcgen05 (kernels.h k_split6_*), fp32 output - vs F.conv2d on the UNROUNDED fp32 operands.  Tolerance: 2e-5 of the
    largest output (the dropped product terms are < 2^-24 relative; fp32 accumulation order)."""
    from salt_b200 import _lib
    lib = _lib.load()
    B, Ci, Co, H, W, k, s, p = case
    g = torch.Generator().manual_seed(sum(case) + 7)
    x = torch.randn(B, Ci, H, W, generator=g)
    w = torch.randn(Co, Ci, k, k, generator=g) * (2.0 / (Ci * k * k)) ** 0.5
    bias = torch.randn(Co, generator=g) * 0.1
    y_ref = F.conv2d(x.double(), w.double(), bias.double(), stride=s, padding=p).float()
    Ho, Wo = y_ref.shape[2:]
    d = _lib.SaltConvDesc(B, H, W, Ci, Ho, Wo, Co, k, s, p, 0, 1)           # precision fp32, use_tensor_cores
    xd, wd, bd = _to_nhwc(x, 'fp32'), w.cuda().contiguous(), bias.cuda()
    out = torch.full((B, Ho, Wo, Co), float('nan'), dtype=torch.float32, device='cuda')
    stats = torch.zeros(2 * Co, dtype=torch.float64, device='cuda')
    n0 = lib.salt_launch_count()
    _lib.check(lib.salt_op_conv_forward(C.byref(d), xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), out.data_ptr(), stats.data_ptr(), None))
    torch.cuda.synchronize()
    ok1 = report('split-bf16 tc conv fwd %s' % (case,), _from_nhwc(out), y_ref, atol=1e-6, rtol=2e-5)[0]
    ok2 = report('split-bf16 tc conv stats sum', stats[:Co].cpu().float(), y_ref.sum((0, 2, 3)), atol=1e-2, rtol=1e-4)[0]
    ok3 = report('split-bf16 tc conv stats sumsq', stats[Co:].cpu().float(), (y_ref ** 2).sum((0, 2, 3)), atol=1e-2, rtol=1e-4)[0]
    assert ok1 and ok2 and ok3


WG_CASES = [c for c in TC_CASES if c[4] != 24] + [      # the wgrad kernel takes 32-pixel row chunks: widths that are multiples of 8 only via padding
    (2, 64, 32, 34, 34, 3, 1, 0),        # dec1.conv1-like: Cout = 32 (zero-filled channel box)
    (4, 64, 64, 128, 128, 3, 1, 1),      # many pixel chunks per CTA
    (2, 256, 512, 16, 16, 3, 2, 1),      # layer4.0.conv1-like, stride 2
]


@pytest.mark.parametrize('case', WG_CASES)
def test_conv_tc_wgrad(case):
    """tcgen05 weight gradient (MN-major operands straight from NHWC TMA boxes, slot pairs stacked in M) vs autograd.
    Tolerance: fp32 accumulation of bf16 products, different summation order: 1e-3 relative to the largest entry."""
    from salt_b200 import _lib
    lib = _lib.load()
    B, Ci, Co, H, W, k, s, p = case
    g = torch.Generator().manual_seed(sum(case) + 1)
    x = torch.randn(B, Ci, H, W, generator=g).bfloat16().float()
    w = (torch.randn(Co, Ci, k, k, generator=g) * 0.05).requires_grad_(True)
    y = F.conv2d(x, w, None, stride=s, padding=p)
    Ho, Wo = y.shape[2:]
    gy = torch.randn(B, Co, Ho, Wo, generator=g).bfloat16().float()
    y.backward(gy)
    d = _lib.SaltConvDesc(B, H, W, Ci, Ho, Wo, Co, k, s, p, 1, 1)
    xd, gyd = _to_nhwc(x, 'bf16'), _to_nhwc(gy, 'bf16')
    dw = torch.zeros((Co, Ci, k, k), dtype=torch.float32, device='cuda')
    _lib.check(lib.salt_op_conv_wgrad(C.byref(d), xd.data_ptr(), gyd.data_ptr(), dw.data_ptr(), None))
    torch.cuda.synchronize()
    ok1 = report('tc wgrad %s' % (case,), dw.cpu(), w.grad, atol=1e-3, rtol=1e-3)[0]
    _lib.check(lib.salt_op_conv_wgrad(C.byref(d), xd.data_ptr(), gyd.data_ptr(), dw.data_ptr(), None))
    torch.cuda.synchronize()
    ok2 = report('tc wgrad accumulates into dw', dw.cpu(), 2 * w.grad, atol=2e-3, rtol=1e-3)[0]
    assert ok1 and ok2

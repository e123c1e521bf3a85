"""Functional CPU restatement of the reference UNetResNet / UNetSeResNet forward (fp32, PyTorch).

TEST INFRASTRUCTURE (see oracle/__init__.py).  The network is expressed as pure
functions over a ``state_dict`` (canonical ``encoders.encoder.*`` keys), so the
same weights drive the oracle, the reference modules and the CUDA engine.

Reference followed:
  common_blocks/architectures/unet.py:44-109     UNetResNet.__init__/forward
  common_blocks/architectures/base.py:7-37       Conv2dBnRelu (replication pad (0,2,2,0))
  common_blocks/architectures/base.py:65-117     DecoderBlock, ChannelSELayer, SpatialSELayer
  common_blocks/architectures/encoders.py:6-45   ResNetEncoders (pool0=False)
  torchvision/models/resnet.py:59-105            BasicBlock
  common_blocks/architectures/unet.py:112-172    UNetSeResNet (depth 50; bottom_channel_nr 2048)
  common_blocks/architectures/encoders.py:48-83  SeResNetEncoders (pretrainedmodels se_resnet50, see senet_restated.py)
  common_blocks/architectures/unet.py:175-236    UNetSeResNetXt (same decoder; arch='UNetSeResNetXt')
  common_blocks/architectures/encoders.py:86-118 SeResNetXtEncoders (pretrainedmodels se_resnext50_32x4d, see senet_restated.py)
"""
import numpy as np
import torch
import torch.nn.functional as F

from .synth import param_specs, resnet_block_counts

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


def to_torch_state(sd_np, requires_grad=False):
    """numpy state -> dict of torch tensors (leaf tensors if requires_grad)."""
    out = {}
    for k, v in sd_np.items():
        t = torch.from_numpy(np.ascontiguousarray(v)).clone()
        if requires_grad and not (k.endswith('running_mean') or k.endswith('running_var')):
            t.requires_grad_(True)
        out[k] = t
    return out


def alias_keys(depth=34, arch=None):
    """alias key -> canonical key, for the duplicated registrations the reference
    creates in ResNetEncoders (encoders.py:21-36)."""
    amap = {}
    stem = 'encoders.encoder.layer0.' if depth >= 50 else 'encoders.encoder.'       # encoders.py:59-66 vs :21-28
    for s in ('weight',):
        amap['encoders.conv1.0.' + s] = stem + 'conv1.' + s
    for s in ('weight', 'bias', 'running_mean', 'running_var', 'num_batches_tracked'):
        amap['encoders.conv1.1.' + s] = stem + 'bn1.' + s
    for name, _, _ in param_specs(depth, 2, arch):
        for li in (1, 2, 3, 4):
            pre = 'encoders.encoder.layer%d.' % li
            if name.startswith(pre):
                amap['encoders.encoder%d.%s' % (li + 1, name[len(pre):])] = name
    return amap


def _bn(sd, prefix, x, train):
    # nn.BatchNorm2d defaults: eps 1e-5, momentum 0.1; train => batch stats (biased var to
    # normalise, unbiased var into running_var), eval => running stats.
    return F.batch_norm(x, sd[prefix + '.running_mean'], sd[prefix + '.running_var'],
                        sd[prefix + '.weight'], sd[prefix + '.bias'],
                        training=train, momentum=BN_MOMENTUM, eps=BN_EPS)


def _basic_block(sd, p, x, stride, train):
    # torchvision resnet.py:89-105
    out = F.conv2d(x, sd[p + 'conv1.weight'], None, stride=stride, padding=1)
    out = F.relu(_bn(sd, p + 'bn1', out, train))
    out = F.conv2d(out, sd[p + 'conv2.weight'], None, stride=1, padding=1)
    out = _bn(sd, p + 'bn2', out, train)
    if (p + 'downsample.0.weight') in sd:
        idn = F.conv2d(x, sd[p + 'downsample.0.weight'], None, stride=stride)
        idn = _bn(sd, p + 'downsample.1', idn, train)
    else:
        idn = x
    return F.relu(out + idn)


def _se_bottleneck(sd, p, x, stride, train):
    # pretrainedmodels senet.py SEResNetBottleneck / SEModule (restated: oracle/senet_restated.py): stride on conv1
    out = F.conv2d(x, sd[p + 'conv1.weight'], None, stride=stride)
    out = F.relu(_bn(sd, p + 'bn1', out, train))
    out = F.conv2d(out, sd[p + 'conv2.weight'], None, stride=1, padding=1)
    out = F.relu(_bn(sd, p + 'bn2', out, train))
    out = F.conv2d(out, sd[p + 'conv3.weight'], None)
    out = _bn(sd, p + 'bn3', out, train)
    if (p + 'downsample.0.weight') in sd:
        idn = F.conv2d(x, sd[p + 'downsample.0.weight'], None, stride=stride)
        idn = _bn(sd, p + 'downsample.1', idn, train)
    else:
        idn = x
    g = out.mean(dim=(2, 3), keepdim=True)
    g = F.relu(F.conv2d(g, sd[p + 'se_module.fc1.weight'], sd[p + 'se_module.fc1.bias']))
    g = torch.sigmoid(F.conv2d(g, sd[p + 'se_module.fc2.weight'], sd[p + 'se_module.fc2.bias']))
    return F.relu(out * g + idn)


def _sex_bottleneck(sd, p, x, stride, train):
    # pretrainedmodels senet.py SEResNeXtBottleneck (restated: oracle/senet_restated.py): 1x1 -> 3x3 (stride, 32 groups) -> 1x1 + SE
    out = F.conv2d(x, sd[p + 'conv1.weight'], None)
    out = F.relu(_bn(sd, p + 'bn1', out, train))
    out = F.conv2d(out, sd[p + 'conv2.weight'], None, stride=stride, padding=1, groups=32)
    out = F.relu(_bn(sd, p + 'bn2', out, train))
    out = F.conv2d(out, sd[p + 'conv3.weight'], None)
    out = _bn(sd, p + 'bn3', out, train)
    if (p + 'downsample.0.weight') in sd:
        idn = F.conv2d(x, sd[p + 'downsample.0.weight'], None, stride=stride)
        idn = _bn(sd, p + 'downsample.1', idn, train)
    else:
        idn = x
    g = out.mean(dim=(2, 3), keepdim=True)
    g = F.relu(F.conv2d(g, sd[p + 'se_module.fc1.weight'], sd[p + 'se_module.fc1.bias']))
    g = torch.sigmoid(F.conv2d(g, sd[p + 'se_module.fc2.weight'], sd[p + 'se_module.fc2.bias']))
    return F.relu(out * g + idn)


def encoder_forward(sd, x, depth, train, arch=None):
    # encoders.py:38-45 / :76-83 with pool0=False (no maxpool)
    e = 'encoders.encoder.'
    stem = e + ('layer0.' if depth >= 50 else '')
    block = _sex_bottleneck if arch == 'UNetSeResNetXt' else (_se_bottleneck if depth >= 50 else _basic_block)
    y = F.conv2d(x, sd[stem + 'conv1.weight'], None, stride=2, padding=3)
    y = F.relu(_bn(sd, stem + 'bn1', y, train))
    feats = []
    for li, nblk in enumerate(resnet_block_counts(depth), start=1):
        for b in range(nblk):
            stride = 2 if (b == 0 and li > 1) else 1
            y = block(sd, '%slayer%d.%d.' % (e, li, b), y, stride, train)
        feats.append(y)
    return feats  # encoder2..encoder5


def conv_bn_relu(sd, prefix, x, train):
    # base.py:29-37: ReplicationPad2d((left 0, right 2, top 2, bottom 0)) -> valid 3x3 conv(+bias) -> BN -> ReLU
    x = F.pad(x, (0, 2, 2, 0), mode='replicate')
    x = F.conv2d(x, sd[prefix + '.conv.weight'], sd[prefix + '.conv.bias'])
    return F.relu(_bn(sd, prefix + '.batch_norm', x, train))


def _up(x, scale):
    # nn.Upsample / F.upsample(mode='bilinear') -> align_corners=False under torch>=0.4
    return F.interpolate(x, scale_factor=scale, mode='bilinear', align_corners=False)


def decoder_block(sd, name, x, skip, train):
    # base.py:75-86
    x = _up(x, 2)
    if skip is not None:
        x = torch.cat([x, skip], 1)
    x = conv_bn_relu(sd, name + '.conv1', x, train)
    x = conv_bn_relu(sd, name + '.conv2', x, train)
    b, c = x.shape[:2]
    g = x.mean(dim=(2, 3))                                            # base.py:100-104
    g = F.relu(F.linear(g, sd[name + '.channel_se.fc.0.weight'], sd[name + '.channel_se.fc.0.bias']))
    g = torch.sigmoid(F.linear(g, sd[name + '.channel_se.fc.2.weight'], sd[name + '.channel_se.fc.2.bias']))
    cse = x * g.view(b, c, 1, 1)
    s = torch.sigmoid(F.conv2d(x, sd[name + '.spatial_se.fc.weight'], sd[name + '.spatial_se.fc.bias']))
    sse = x * s                                                      # base.py:113-117
    return F.relu(cse + sse)


def unet_resnet_forward(sd, x, depth=34, train=False, return_stages=False, arch=None):
    """[B,3,H,W] fp32 -> logits [B,num_classes,H,W] (unet.py:89-109, hypercolumn on,
    dropout_2d p=0 is the identity).  depth 50 = UNetSeResNet (unet.py:152-172): same graph, SE-ResNet-50 encoder,
    decoder widths derived from the state (bottom_channel_nr 2048)."""
    e2, e3, e4, e5 = encoder_forward(sd, x, depth, train, arch)
    c = conv_bn_relu(sd, 'center.0', e5, train)
    c = conv_bn_relu(sd, 'center.1', c, train)
    c = F.avg_pool2d(c, 2, 2)
    d5 = decoder_block(sd, 'dec5', c, e5, train)
    d4 = decoder_block(sd, 'dec4', d5, e4, train)
    d3 = decoder_block(sd, 'dec3', d4, e3, train)
    d2 = decoder_block(sd, 'dec2', d3, e2, train)
    d1 = decoder_block(sd, 'dec1', d2, None, train)
    hyper = torch.cat([d1, _up(d2, 2), _up(d3, 4), _up(d4, 8), _up(d5, 16)], 1)
    f = conv_bn_relu(sd, 'final.0', hyper, train)
    logits = F.conv2d(f, sd['final.1.weight'], sd['final.1.bias'])
    if return_stages:
        return logits, dict(e2=e2, e3=e3, e4=e4, e5=e5, center=c, d5=d5, d4=d4, d3=d3, d2=d2, d1=d1, f=f)
    return logits


# ---------------------------------------------------------------- optimiser
def adam_l2_step(params, grads, m, v, step, lr=1e-4, wd=1e-4, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam with (non-decoupled) L2 `grad += wd*p`, the optimiser the
    reference builds (models.py:74-75, 289-297).  In-place on dicts of tensors."""
    for k in params:
        g = grads[k] + wd * params[k]
        m[k].mul_(b1).add_(g, alpha=1 - b1)
        v[k].mul_(b2).addcmul_(g, g, value=1 - b2)
        bc1 = 1 - b1 ** step
        bc2 = 1 - b2 ** step
        denom = (v[k].sqrt() / (bc2 ** 0.5)).add_(eps)
        params[k].addcdiv_(m[k], denom, value=-lr / bc1)

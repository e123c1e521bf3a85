"""Deterministic synthetic weights / inputs / targets.

Pure numpy so the generated
tensors do not depend on torch's RNG or initialisers.
"""
import numpy as np

MEAN = (0.485, 0.456, 0.406)   # reference main.py:55
STD = (0.229, 0.224, 0.225)    # reference main.py:56


def resnet_block_counts(depth):
    return {18: (2, 2, 2, 2), 34: (3, 4, 6, 3), 50: (3, 4, 6, 3), 101: (3, 4, 23, 3), 152: (3, 8, 36, 3)}[depth]


def param_specs(depth=34, num_classes=2, arch=None):
    """Canonical (name, shape, kind) list of every tensor the network owns.

    depth 18/34 -> UNetResNet (unet.py:22-109), depth 50/101/152 -> UNetSeResNet
    (unet.py:112-172, SE-ResNet-50 encoder).  Names are the reference's
    ``state_dict`` keys under ``encoders.encoder.*`` (the aliases
    ``encoders.conv1.*`` / ``encoders.encoderN.*`` share storage, reference
    encoders.py:21-36 / :59-74).  kind in {conv_w, bias, bn_w, bn_b, bn_rm, bn_rv, lin_w}.
    """
    specs = []

    def bn(prefix, c):
        specs.append((prefix + '.weight', (c,), 'bn_w'))
        specs.append((prefix + '.bias', (c,), 'bn_b'))
        specs.append((prefix + '.running_mean', (c,), 'bn_rm'))
        specs.append((prefix + '.running_var', (c,), 'bn_rv'))

    e = 'encoders.encoder.'
    se50 = depth >= 50          # SE-ResNet-50 / 101 / 152 (reference encoders.py:52-57)
    resnext = arch == 'UNetSeResNetXt'      # SE-ResNeXt 32x4d (encoders.py:90-95): width = 2 * planes, conv2 has 32 groups
    specs.append((e + ('layer0.conv1.weight' if se50 else 'conv1.weight'), (64, 3, 7, 7), 'conv_w'))
    bn(e + ('layer0.bn1' if se50 else 'bn1'), 64)
    cin = 64
    for li, (nblk, cout) in enumerate(zip(resnet_block_counts(depth), (64, 128, 256, 512)), start=1):
        for b in range(nblk):
            p = '%slayer%d.%d.' % (e, li, b)
            if se50:        # pretrainedmodels SEResNetBottleneck: registration order conv1..bn3, se_module, downsample
                co = 4 * cout
                wd = 2 * cout if resnext else cout
                specs.append((p + 'conv1.weight', (wd, cin, 1, 1), 'conv_w'))
                bn(p + 'bn1', wd)
                specs.append((p + 'conv2.weight', (wd, wd // 32 if resnext else wd, 3, 3), 'conv_w'))
                bn(p + 'bn2', wd)
                specs.append((p + 'conv3.weight', (co, wd, 1, 1), 'conv_w'))
                bn(p + 'bn3', co)
                specs.append((p + 'se_module.fc1.weight', (co // 16, co, 1, 1), 'conv_w'))
                specs.append((p + 'se_module.fc1.bias', (co // 16,), 'bias'))
                specs.append((p + 'se_module.fc2.weight', (co, co // 16, 1, 1), 'conv_w'))
                specs.append((p + 'se_module.fc2.bias', (co,), 'bias'))
                if b == 0:
                    specs.append((p + 'downsample.0.weight', (co, cin, 1, 1), 'conv_w'))
                    bn(p + 'downsample.1', co)
                cin = co
                continue
            specs.append((p + 'conv1.weight', (cout, cin, 3, 3), 'conv_w'))
            bn(p + 'bn1', cout)
            specs.append((p + 'conv2.weight', (cout, cout, 3, 3), 'conv_w'))
            bn(p + 'bn2', cout)
            if b == 0 and li > 1:
                specs.append((p + 'downsample.0.weight', (cout, cin, 1, 1), 'conv_w'))
                bn(p + 'downsample.1', cout)
            cin = cout

    def cbr(prefix, ci, co):
        bn(prefix + '.batch_norm', co)
        specs.append((prefix + '.conv.weight', (co, ci, 3, 3), 'conv_w'))
        specs.append((prefix + '.conv.bias', (co,), 'bias'))

    bc = 2048 if se50 else 512
    cbr('center.0', bc, bc)
    cbr('center.1', bc, bc // 2)
    dec = {'dec5': (bc + bc // 2, bc, bc // 8), 'dec4': (bc // 2 + bc // 8, bc // 2, bc // 8),
           'dec3': (bc // 4 + bc // 8, bc // 4, bc // 8), 'dec2': (bc // 8 + bc // 8, bc // 8, bc // 8),
           'dec1': (bc // 8, bc // 16, bc // 8)}
    for name in ('dec5', 'dec4', 'dec3', 'dec2', 'dec1'):
        ci, cm, co = dec[name]
        cbr(name + '.conv1', ci, cm)
        cbr(name + '.conv2', cm, co)
        specs.append((name + '.channel_se.fc.0.weight', (co // 16, co), 'lin_w'))
        specs.append((name + '.channel_se.fc.0.bias', (co // 16,), 'bias'))
        specs.append((name + '.channel_se.fc.2.weight', (co, co // 16), 'lin_w'))
        specs.append((name + '.channel_se.fc.2.bias', (co,), 'bias'))
        specs.append((name + '.spatial_se.fc.weight', (1, co, 1, 1), 'conv_w'))
        specs.append((name + '.spatial_se.fc.bias', (1,), 'bias'))
    cbr('final.0', 5 * bc // 8, bc // 8)
    specs.append(('final.1.weight', (num_classes, bc // 8, 1, 1), 'conv_w'))
    specs.append(('final.1.bias', (num_classes,), 'bias'))
    return specs


def synth_state_dict(depth=34, num_classes=2, seed=0, arch=None):
    """name -> float32 ndarray; He-scaled conv weights, non-trivial BN statistics."""
    rng = np.random.default_rng(seed)
    sd = {}
    for name, shape, kind in param_specs(depth, num_classes, arch):
        if kind in ('conv_w', 'lin_w'):
            fan_in = int(np.prod(shape[1:]))
            a = rng.standard_normal(shape) * np.sqrt(2.0 / fan_in)
        elif kind == 'bias':
            a = rng.standard_normal(shape) * 0.1
        elif kind == 'bn_w':
            # the last BN of a residual branch gets a small gain so that eval-mode activations (running
            # statistics, no renormalisation) stay O(1) through 16 residual blocks
            last = 'bn3.weight' if depth >= 50 else 'bn2.weight'
            a = rng.uniform(0.1, 0.5, shape) if name.endswith(last) else rng.uniform(0.5, 1.5, shape)
        elif kind == 'bn_b':
            a = rng.standard_normal(shape) * 0.1
        elif kind == 'bn_rm':
            a = rng.standard_normal(shape) * 0.1
        elif kind == 'bn_rv':
            a = rng.uniform(0.5, 1.5, shape)
        else:
            raise ValueError(kind)
        sd[name] = a.astype(np.float32)
    return sd


def synth_tiles_u8(batch, size=101, seed=1234):
    """Raw single-channel u8 tiles, the data the pipeline starts from."""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (batch, size, size), dtype=np.uint8)
    # smooth a little so the tiles are not pure noise (3x3 box blur, integer math)
    p = np.pad(base.astype(np.int32), ((0, 0), (1, 1), (1, 1)), mode='edge')
    acc = sum(p[:, dy:dy + size, dx:dx + size] for dy in range(3) for dx in range(3))
    return (acc // 9).astype(np.uint8)


def adapt_tiles(tiles_u8, out_size=128):
    """u8 [B,h,w] -> fp32 NCHW [B,3,S,S] the way the reference loader does for
    inference (loaders.py:607-612, augmentation.py:272-281, utils.py:494-500):
    edge-pad to S (top/left = floor(d/2), rest bottom/right), /255, ImageNet
    normalise each of the 3 replicated grey channels, then ch1 := linspace(0,1,S)
    per row and ch2 := ch0*ch1."""
    b, h, w = tiles_u8.shape
    dv, dh = out_size - h, out_size - w
    top, left = dv // 2, dh - dh // 2
    x = np.pad(tiles_u8, ((0, 0), (top, dv - top), (left, dh - left)), mode='edge').astype(np.float32) / 255.0
    out = np.empty((b, 3, out_size, out_size), np.float32)
    for c in range(3):
        out[:, c] = (x - MEAN[c]) / STD[c]
    out[:, 1] = np.linspace(0, 1, out_size, dtype=np.float64).astype(np.float32)[None, :, None]
    out[:, 2] = out[:, 0] * out[:, 1]
    return out


def synth_inputs(batch, size=128, seed=1234):
    tile = size - 27 if size >= 64 else size
    return adapt_tiles(synth_tiles_u8(batch, tile, seed), size)


def synth_targets(batch, size=128, seed=1234):
    """[B,2,S,S] fp32: ch1 = salt (union of 0-3 rectangles, ~40% empty), ch0 = 1-salt
    (loaders.py:186-190 emits [background, salt])."""
    rng = np.random.default_rng(seed + 1)
    m = np.zeros((batch, size, size), np.float32)
    for i in range(batch):
        if rng.random() < 0.4:
            continue
        for _ in range(int(rng.integers(1, 4))):
            y0, x0 = rng.integers(0, size - 4, 2)
            hh, ww = rng.integers(4, size // 2 + 4, 2)
            m[i, y0:min(size, y0 + hh), x0:min(size, x0 + ww)] = 1.0
    return np.stack([1.0 - m, m], axis=1).astype(np.float32)


def synth_salt_scenes(batch, size=128, seed=1234):
    """(x fp32 [B,3,S,S], target fp32 [B,2,S,S]) where the salt mask can be LEARNED from the image: inside the salt rectangles of
    ``synth_targets`` the grey tile is brighter and smoother than outside.  Used to train a network for a few dozen steps so that the
    inference parity checks (BASELINE config 5: masks / IoU vs the reference path) run on confident, non-trivial masks instead of
    the all-background output of a random initialisation."""
    tile = size - 27 if size >= 64 else size
    t = synth_targets(batch, size, seed)
    rng = np.random.default_rng(seed + 2)
    dv = size - tile
    top, left = dv // 2, dv - dv // 2
    salt = t[:, 1, top:top + tile, left:left + tile]
    noise = rng.integers(0, 256, (batch, tile, tile)).astype(np.float32)
    p = np.pad(noise, ((0, 0), (1, 1), (1, 1)), mode='edge')
    smooth = sum(p[:, dy:dy + tile, dx:dx + tile] for dy in range(3) for dx in range(3)) / 9.0
    img = np.where(salt > 0, 150.0 + 0.35 * (smooth - 128.0), 70.0 + 0.8 * (noise - 128.0) * 0.5)
    tiles = np.clip(img, 0, 255).astype(np.uint8)
    return adapt_tiles(tiles, size), t

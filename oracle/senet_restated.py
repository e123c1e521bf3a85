"""nn.Module restatement of ``pretrainedmodels.se_resnet50`` (Cadene pretrained-models.pytorch, senet.py).

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference pins ``pretrainedmodels==0.7.0``
(environment.yml:19) and calls ``pretrainedmodels.__dict__['se_resnet50'](num_classes=1000, pretrained=...)``
(common_blocks/architectures/encoders.py:52-53).  That package is NOT in /root/reference and not installed
here, so the SE-ResNet-50 encoder is restated from its published architecture (Hu et al., "Squeeze-and-
Excitation Networks"; Caffe-style bottleneck) with the attribute names of that package, so the reference's
own ``SeResNetEncoders`` / ``UNetSeResNet`` wrap it unchanged and produce the real ``state_dict`` keys:

  layer0 = Sequential(conv1 7x7/2 pad 3 (no bias), bn1, relu1, pool0 3x3/2 ceil)      (pool0 is skipped by the U-Net)
  layerN = Sequential of SEResNetBottleneck: conv1 1x1 (STRIDE HERE) - bn1 - relu - conv2 3x3 pad 1 - bn2 - relu -
           conv3 1x1 (x4) - bn3 - se_module - (+ downsample(x) = 1x1 conv stride s + BN on the first block) - relu
  se_module = global average pool - fc1 (1x1 conv C -> C/16, bias) - relu - fc2 (1x1 conv C/16 -> C, bias) - sigmoid - scale
  blocks per layer [3, 4, 6, 3], planes [64, 128, 256, 512], avg_pool 7, last_linear 2048 -> num_classes

SE-ResNeXt 32x4d (``se_resnext50_32x4d`` / ``se_resnext101_32x4d``, reference encoders.py:90-95; Xie et al., "Aggregated Residual
Transformations"): the same SENet skeleton with SEResNeXtBottleneck - conv1 1x1 (stride 1) to width = floor(planes * 4 / 64) * 32
= 2 * planes, conv2 3x3 pad 1 with the STRIDE and 32 groups, conv3 1x1 to 4 * planes, then bn3, se_module, shortcut, relu.

PARITY UNPINNED for these encoders: there is no copy of the original to run against; what IS pinned by
oracle/make_golden.py is everything the reference owns on top of it (UNetSeResNet wiring, decoder, key aliasing).
"""
from collections import OrderedDict

from torch import nn


class SEModule(nn.Module):
    def __init__(self, channels, reduction):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.fc1 = nn.Conv2d(channels, channels // reduction, kernel_size=1, padding=0)
        self.relu = nn.ReLU(inplace=True)
        self.fc2 = nn.Conv2d(channels // reduction, channels, kernel_size=1, padding=0)
        self.sigmoid = nn.Sigmoid()

    def forward(self, x):
        gate = self.sigmoid(self.fc2(self.relu(self.fc1(self.avg_pool(x)))))
        return x * gate


class SEResNetBottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, groups, reduction, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, kernel_size=1, bias=False, stride=stride)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, padding=1, groups=groups, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, kernel_size=1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.se_module = SEModule(planes * 4, reduction=reduction)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        shortcut = x if self.downsample is None else self.downsample(x)
        y = self.relu(self.bn1(self.conv1(x)))
        y = self.relu(self.bn2(self.conv2(y)))
        y = self.bn3(self.conv3(y))
        return self.relu(self.se_module(y) + shortcut)


class SEResNeXtBottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, groups, reduction, stride=1, downsample=None, base_width=4):
        super().__init__()
        width = (planes * base_width // 64) * groups
        self.conv1 = nn.Conv2d(inplanes, width, kernel_size=1, bias=False, stride=1)
        self.bn1 = nn.BatchNorm2d(width)
        self.conv2 = nn.Conv2d(width, width, kernel_size=3, stride=stride, padding=1, groups=groups, bias=False)
        self.bn2 = nn.BatchNorm2d(width)
        self.conv3 = nn.Conv2d(width, planes * 4, kernel_size=1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.se_module = SEModule(planes * 4, reduction=reduction)
        self.downsample = downsample
        self.stride = stride

    forward = SEResNetBottleneck.forward


class SENet(nn.Module):
    def __init__(self, layers=(3, 4, 6, 3), reduction=16, num_classes=1000, block=SEResNetBottleneck, groups=1):
        super().__init__()
        self.block, self.groups = block, groups
        self.inplanes = 64
        self.layer0 = nn.Sequential(OrderedDict([
            ('conv1', nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)),
            ('bn1', nn.BatchNorm2d(64)),
            ('relu1', nn.ReLU(inplace=True)),
            ('pool0', nn.MaxPool2d(3, stride=2, ceil_mode=True)),
        ]))
        self.layer1 = self._make_layer(64, layers[0], reduction, stride=1)
        self.layer2 = self._make_layer(128, layers[1], reduction, stride=2)
        self.layer3 = self._make_layer(256, layers[2], reduction, stride=2)
        self.layer4 = self._make_layer(512, layers[3], reduction, stride=2)
        self.avg_pool = nn.AvgPool2d(7, stride=1)
        self.dropout = None
        self.last_linear = nn.Linear(512 * SEResNetBottleneck.expansion, num_classes)

    def _make_layer(self, planes, blocks, reduction, stride):
        out_planes = planes * SEResNetBottleneck.expansion
        downsample = None
        if stride != 1 or self.inplanes != out_planes:
            downsample = nn.Sequential(nn.Conv2d(self.inplanes, out_planes, kernel_size=1, stride=stride, padding=0, bias=False),
                                       nn.BatchNorm2d(out_planes))
        mods = [self.block(self.inplanes, planes, self.groups, reduction, stride, downsample)]
        self.inplanes = out_planes
        for _ in range(1, blocks):
            mods.append(self.block(self.inplanes, planes, self.groups, reduction))
        return nn.Sequential(*mods)

    def forward(self, x):
        x = self.layer4(self.layer3(self.layer2(self.layer1(self.layer0(x)))))
        x = self.avg_pool(x)
        return self.last_linear(x.view(x.size(0), -1))


def se_resnet50(num_classes=1000, pretrained=None):
    """Same call signature as the pinned package; there is no network here, so pretrained weights are never loaded."""
    return SENet((3, 4, 6, 3), reduction=16, num_classes=num_classes)


def se_resnet101(num_classes=1000, pretrained=None):
    """pretrainedmodels se_resnet101: the same SEResNetBottleneck, layers [3, 4, 23, 3] (reference encoders.py:54-55)."""
    return SENet((3, 4, 23, 3), reduction=16, num_classes=num_classes)


def se_resnet152(num_classes=1000, pretrained=None):
    """pretrainedmodels se_resnet152: layers [3, 8, 36, 3] (reference encoders.py:56-57)."""
    return SENet((3, 8, 36, 3), reduction=16, num_classes=num_classes)


def se_resnext50_32x4d(num_classes=1000, pretrained=None):
    """pretrainedmodels se_resnext50_32x4d: SEResNeXtBottleneck, layers [3, 4, 6, 3], 32 groups (reference encoders.py:90-91)."""
    return SENet((3, 4, 6, 3), reduction=16, num_classes=num_classes, block=SEResNeXtBottleneck, groups=32)


def se_resnext101_32x4d(num_classes=1000, pretrained=None):
    """pretrainedmodels se_resnext101_32x4d: layers [3, 4, 23, 3] (reference encoders.py:92-93)."""
    return SENet((3, 4, 23, 3), reduction=16, num_classes=num_classes, block=SEResNeXtBottleneck, groups=32)

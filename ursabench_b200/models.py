"""Network definitions the hot path is measured on.

These are the *forward targets* of the BMA kernels and the shapes the flat
weight vector is laid out for; they are written from scratch but keep the
reference's module attribute names, so ``state_dict`` keys and
``model.parameters()`` order (= the flat layout, reference ``util.flatten``
util.py:163-169) are identical:

* ``MLP``         reference models/mlp.py:8-23     (fc1 -> relu -> fc2 -> relu -> fc3)
* ``PreResNet``   reference models/preresnet.py:90-151 (pre-activation BasicBlock, depth < 44)
* ``WideResNet``  reference models/wideresnet.py:78-120 (WRN-28-10: stride on conv2, biased convs)

Config holders (``MLP400MNIST`` ...) expose ``base / args / kwargs`` like the
reference so ``model_cfg.base(*model_cfg.args, num_classes=C, **model_cfg.kwargs)``
(experiment.py:74-76) keeps working.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

__all__ = ["MLP", "PreResNet", "WideResNet", "MLP200MNIST", "MLP400MNIST", "MLP600MNIST",
           "PreResNet8", "PreResNet20", "WideResNet28x10"]


class MLP(nn.Module):
    def __init__(self, hidden_size, input_dim, num_classes):
        super().__init__()
        self.input_dim, self.hidden_size, self.num_classes = input_dim, hidden_size, num_classes
        self.fc1 = nn.Linear(input_dim, hidden_size)
        self.fc2 = nn.Linear(hidden_size, hidden_size)
        self.fc3 = nn.Linear(hidden_size, num_classes)

    def forward(self, x):
        h = self.fc1(x.view(-1, self.input_dim))
        h = self.fc2(F.relu(h))
        return self.fc3(F.relu(h))


class BasicBlock(nn.Module):
    """bn1-relu-conv1(stride)-bn2-relu-conv2 + shortcut on the raw block input."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.bn1 = nn.BatchNorm2d(inplanes)
        self.relu = nn.ReLU(inplace=True)
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=1, padding=1, bias=False)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        y = self.conv1(self.relu(self.bn1(x)))
        y = self.conv2(self.relu(self.bn2(y)))
        skip = x if self.downsample is None else self.downsample(x)
        return y + skip


class PreResNet(nn.Module):
    """CIFAR pre-activation ResNet, BasicBlock variant only (depth = 6n + 2 < 44)."""

    def __init__(self, num_classes=10, depth=20):
        super().__init__()
        if depth >= 44 or (depth - 2) % 6 != 0:
            raise ValueError("this engine covers the BasicBlock PreResNets: depth = 6n+2 < 44")
        n = (depth - 2) // 6
        self.depth = depth
        self.num_classes = num_classes
        self.conv1 = nn.Conv2d(3, 16, 3, padding=1, bias=False)
        widths, strides, inplanes = (16, 32, 64), (1, 2, 2), 16
        for i, (w, s) in enumerate(zip(widths, strides), start=1):
            blocks = []
            for b in range(n):
                stride = s if b == 0 else 1
                down = None
                if stride != 1 or inplanes != w:
                    down = nn.Sequential(nn.Conv2d(inplanes, w, 1, stride=stride, bias=False))
                blocks.append(BasicBlock(inplanes, w, stride, down))
                inplanes = w
            setattr(self, "layer%d" % i, nn.Sequential(*blocks))
        self.bn = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.avgpool = nn.AvgPool2d(8)
        self.fc = nn.Linear(64, num_classes)
        for m in self.modules():                       # reference preresnet.py:114-120
            if isinstance(m, nn.Conv2d):
                fan = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2.0 / fan))
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    def forward(self, x):
        x = self.layer3(self.layer2(self.layer1(self.conv1(x))))
        x = self.avgpool(self.relu(self.bn(x)))
        return self.fc(x.view(x.size(0), -1))


class WideBasic(nn.Module):
    def __init__(self, in_planes, planes, dropout_rate, stride=1):
        super().__init__()
        self.bn1 = nn.BatchNorm2d(in_planes)
        self.conv1 = nn.Conv2d(in_planes, planes, 3, padding=1, bias=True)
        self.dropout = nn.Dropout(p=dropout_rate)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=stride, padding=1, bias=True)
        self.shortcut = nn.Sequential()
        if stride != 1 or in_planes != planes:
            self.shortcut = nn.Sequential(nn.Conv2d(in_planes, planes, 1, stride=stride, bias=True))

    def forward(self, x):
        y = self.dropout(self.conv1(F.relu(self.bn1(x))))
        y = self.conv2(F.relu(self.bn2(y)))
        return y + self.shortcut(x)


class WideResNet(nn.Module):
    def __init__(self, num_classes=10, depth=28, widen_factor=10, dropout_rate=0.0):
        super().__init__()
        if (depth - 4) % 6 != 0:
            raise ValueError("Wide-resnet depth should be 6n+4")
        n = (depth - 4) // 6
        widths = [16, 16 * widen_factor, 32 * widen_factor, 64 * widen_factor]
        self.num_classes, self.depth, self.widen_factor = num_classes, depth, widen_factor
        self.conv1 = nn.Conv2d(3, widths[0], 3, padding=1, bias=True)
        in_planes = widths[0]
        for i, stride in enumerate((1, 2, 2), start=1):
            blocks = []
            for b in range(n):
                blocks.append(WideBasic(in_planes, widths[i], dropout_rate, stride if b == 0 else 1))
                in_planes = widths[i]
            setattr(self, "layer%d" % i, nn.Sequential(*blocks))
        self.bn1 = nn.BatchNorm2d(widths[3], momentum=0.9)
        self.linear = nn.Linear(widths[3], num_classes)

    def forward(self, x):
        y = self.layer3(self.layer2(self.layer1(self.conv1(x))))
        y = F.avg_pool2d(F.relu(self.bn1(y)), 8)
        return self.linear(y.view(y.size(0), -1))


class _Cfg:
    args = list()
    kwargs = dict()
    transform_train = None
    transform_test = None


class MLP200MNIST(_Cfg):
    base = MLP
    kwargs = {"hidden_size": 200, "input_dim": 784}


class MLP400MNIST(_Cfg):
    base = MLP
    kwargs = {"hidden_size": 400, "input_dim": 784}


class MLP600MNIST(_Cfg):
    base = MLP
    kwargs = {"hidden_size": 600, "input_dim": 784}


class PreResNet20(_Cfg):
    """Config 2/5 of BASELINE.json (the reference has no such holder; it is PreResNet(depth=20))."""
    base = PreResNet
    kwargs = {"depth": 20}


class PreResNet8(_Cfg):
    base = PreResNet
    kwargs = {"depth": 8}


class WideResNet28x10(_Cfg):
    base = WideResNet
    kwargs = {"depth": 28, "widen_factor": 10}

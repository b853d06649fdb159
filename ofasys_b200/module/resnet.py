"""ResNet feature extractor (conv1 .. layer3, stride 16, 1024 channels) with the reference's module tree and
parameter names (ofasys/module/resnet.py:139-261) executed channel-last on the sm_100a kernels:
1x1 convs -> tcgen05 GEMM, 3x3 / 7x7 convs -> im2col + GEMM, BatchNorm (training statistics) + residual + ReLU
fused (csrc/resnet.cu); eval-mode / frozen BatchNorm uses the running statistics (ops.batch_norm_eval)."""
import torch
import torch.nn as nn

from .. import ops


def conv3x3(in_planes, out_planes, stride=1):
    return nn.Conv2d(in_planes, out_planes, kernel_size=3, stride=stride, padding=1, bias=False)


def conv1x1(in_planes, out_planes, stride=1):
    return nn.Conv2d(in_planes, out_planes, kernel_size=1, stride=stride, bias=False)


def _bn(bn: nn.BatchNorm2d, x, residual=None, relu=False):
    """nn.BatchNorm2d.forward: batch statistics + running-buffer update in training mode, running statistics in eval mode
    (model.eval(), or the frozen backbone of `freeze_resnet`, adaptor/image_resnet.py:107-114)."""
    if not bn.training:
        return ops.batch_norm_eval(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, residual, relu, bn.eps)
    return ops.batch_norm_train(x, bn.weight, bn.bias, residual, relu, bn.eps, bn.momentum, bn.running_mean, bn.running_var)


def _conv1x1(conv: nn.Conv2d, x):
    """x: [B, H, W, Cin] bf16 -> [B, H', W', Cout]"""
    if conv.stride[0] == 2:
        x = ops.subsample2(x)
    B, H, W, C = x.shape
    y = ops.linear(x.reshape(B * H * W, C), conv.weight.reshape(conv.out_channels, C), None)
    return y.view(B, H, W, conv.out_channels)


def _conv3x3(conv: nn.Conv2d, x):
    B, H, W, C = x.shape
    s = conv.stride[0]
    cols = ops.im2col_nhwc(x, 3, s, 1)
    w = ops.transpose_last2(conv.weight.reshape(conv.out_channels, C, 9)).reshape(conv.out_channels, 9 * C)
    Ho, Wo = (H + 2 - 3) // s + 1, (W + 2 - 3) // s + 1
    return ops.linear(cols, w, None).view(B, Ho, Wo, conv.out_channels)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None, drop_path_rate=0.0):
        super().__init__()
        width = planes
        self.conv1 = conv1x1(inplanes, width)
        self.bn1 = nn.BatchNorm2d(width)
        self.conv2 = conv3x3(width, width, stride)
        self.bn2 = nn.BatchNorm2d(width)
        self.conv3 = conv1x1(width, planes * self.expansion)
        self.bn3 = nn.BatchNorm2d(planes * self.expansion)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride
        assert drop_path_rate == 0.0, "resnet drop-path is 0 in every BASELINE config"

    def forward(self, x):
        identity = x
        out = _bn(self.bn1, _conv1x1(self.conv1, x), relu=True)
        out = _bn(self.bn2, _conv3x3(self.conv2, out), relu=True)
        out = _conv1x1(self.conv3, out)
        if self.downsample is not None:
            identity = _bn(self.downsample[1], _conv1x1(self.downsample[0], x))
        return _bn(self.bn3, out, residual=identity, relu=True)  # relu(identity + bn3(conv3))


class ResNet(nn.Module):
    def __init__(self, layers, norm_layer=None, drop_path_rate=0.0):
        super().__init__()
        assert norm_layer is None, "sync_bn is off in every BASELINE config"
        self.inplanes = 64
        self.conv1 = nn.Conv2d(3, self.inplanes, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(self.inplanes)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.layer1 = self._make_layer(64, layers[0])
        self.layer2 = self._make_layer(128, layers[1], stride=2)
        self.layer3 = self._make_layer(256, layers[2], stride=2)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def _make_layer(self, planes, blocks, stride=1):
        downsample = None
        if stride != 1 or self.inplanes != planes * Bottleneck.expansion:
            downsample = nn.Sequential(conv1x1(self.inplanes, planes * Bottleneck.expansion, stride), nn.BatchNorm2d(planes * Bottleneck.expansion))
        layers = [Bottleneck(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes * Bottleneck.expansion
        for _ in range(1, blocks):
            layers.append(Bottleneck(self.inplanes, planes))
        return nn.Sequential(*layers)

    def forward(self, x):
        """x: [B, 3, H, W] (fp32 or bf16) -> channel-last features bf16 [B, H/16, W/16, 1024]."""
        B = x.shape[0]
        if self.training:  # torch's _BatchNorm.forward bumps num_batches_tracked (checkpoints carry it): one multi-tensor add
            nbt = [m.num_batches_tracked for m in self.modules() if isinstance(m, nn.BatchNorm2d) and m.training and m.num_batches_tracked is not None]
            if nbt:
                torch._foreach_add_(nbt, 1)
        k = 3 * 49
        kp = (k + 7) // 8 * 8
        cols, Ho, Wo = ops.im2col_nchw(x, 7, 2, 3, kp)
        w = self.conv1.weight.reshape(64, k)
        if kp != k:
            w = torch.nn.functional.pad(w, (0, kp - k))
        y = ops.linear(cols, w, None).view(B, Ho, Wo, 64)
        y = _bn(self.bn1, y, relu=True)
        y = ops.maxpool3x3s2(y)
        y = self.layer1(y)
        y = self.layer2(y)
        return self.layer3(y)


def resnet50_backbone(norm_layer=None, drop_path_rate=0.0):
    return ResNet([3, 4, 6], norm_layer, drop_path_rate)


def resnet101_backbone(norm_layer=None, drop_path_rate=0.0):
    return ResNet([3, 4, 23], norm_layer, drop_path_rate)


def resnet152_backbone(norm_layer=None, drop_path_rate=0.0):
    return ResNet([3, 8, 36], norm_layer, drop_path_rate)

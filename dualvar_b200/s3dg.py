"""Drop-in S3D / S3D-G (backbone/s3dg.py:8-217): same module tree and state_dict keys, running on the
sm_100a kernels. Inception branches write straight into channel slices of one concat tensor (no
torch.cat copy) and S3D-G's self-gating scales those slices in place.

As in backbones.py the nn.Conv3d / nn.BatchNorm3d / nn.Linear objects are parameter containers only.
"""
import torch
import torch.nn as nn

from . import engine as E
from .backbones import _Encoder


class BasicConv3d(nn.Module):
    """conv (no bias, N(0, 0.01) init) -> BN -> ReLU (backbone/s3dg.py:8-28)."""

    def __init__(self, in_planes, out_planes, kernel_size, stride, padding=0):
        super().__init__()
        self.conv = nn.Conv3d(in_planes, out_planes, kernel_size=kernel_size, stride=stride, padding=padding,
                              bias=False)
        self.bn = nn.BatchNorm3d(out_planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv.weight.data.normal_(mean=0, std=0.01)
        self.bn.weight.data.fill_(1)
        self.bn.bias.data.zero_()

    def run(self, ctx, x, out=None, coff=0):
        raw = E.conv_stats(ctx, x, self.conv, self.bn)
        return E.activate(ctx, raw, out=out, out_coff=coff), raw


class STConv3d(nn.Module):
    """(1,k,k) conv-BN-ReLU then (k,1,1) conv-BN-ReLU (backbone/s3dg.py:30-65)."""

    def __init__(self, in_planes, out_planes, kernel_size, stride, padding=0):
        super().__init__()
        if isinstance(stride, tuple):
            t_stride, stride = stride[0], stride[-1]
        else:
            t_stride = stride
        self.conv1 = nn.Conv3d(in_planes, out_planes, kernel_size=(1, kernel_size, kernel_size),
                               stride=(1, stride, stride), padding=(0, padding, padding), bias=False)
        self.conv2 = nn.Conv3d(out_planes, out_planes, kernel_size=(kernel_size, 1, 1), stride=(t_stride, 1, 1),
                               padding=(padding, 0, 0), bias=False)
        self.bn1 = nn.BatchNorm3d(out_planes)
        self.bn2 = nn.BatchNorm3d(out_planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv1.weight.data.normal_(mean=0, std=0.01)
        self.conv2.weight.data.normal_(mean=0, std=0.01)
        for bn in (self.bn1, self.bn2):
            bn.weight.data.fill_(1)
            bn.bias.data.zero_()

    def run(self, ctx, x, out=None, coff=0):
        h = E.activate(ctx, E.conv_stats(ctx, x, self.conv1, self.bn1), conv_only=True)   # feeds conv2 only: applied inside it
        raw = E.conv_stats(ctx, h, self.conv2, self.bn2)
        return E.activate(ctx, raw, out=out, out_coff=coff), raw


class SelfGating(nn.Module):
    """backbone/s3dg.py:68-78 (parameter container; applied by engine.self_gate)."""

    def __init__(self, input_dim):
        super().__init__()
        self.fc = nn.Linear(input_dim, input_dim)


class SepInception(nn.Module):
    """backbone/s3dg.py:81-132."""

    def __init__(self, in_planes, out_planes, gating=False):
        super().__init__()
        assert len(out_planes) == 6 and isinstance(out_planes, list)
        o0, o1a, o1b, o2a, o2b, o3 = out_planes
        self.branch0 = nn.Sequential(BasicConv3d(in_planes, o0, kernel_size=1, stride=1))
        self.branch1 = nn.Sequential(BasicConv3d(in_planes, o1a, kernel_size=1, stride=1),
                                     STConv3d(o1a, o1b, kernel_size=3, stride=1, padding=1))
        self.branch2 = nn.Sequential(BasicConv3d(in_planes, o2a, kernel_size=1, stride=1),
                                     STConv3d(o2a, o2b, kernel_size=3, stride=1, padding=1))
        self.branch3 = nn.Sequential(nn.MaxPool3d(kernel_size=(3, 3, 3), stride=1, padding=1),
                                     BasicConv3d(in_planes, o3, kernel_size=1, stride=1))
        self.widths = (o0, o1b, o2b, o3)
        self.out_channels = sum(self.widths)
        self.gating = gating
        if gating:
            self.gating_b0 = SelfGating(o0)
            self.gating_b1 = SelfGating(o1b)
            self.gating_b2 = SelfGating(o2b)
            self.gating_b3 = SelfGating(o3)

    def run(self, ctx, x):
        cat = E.new_concat(x, self.out_channels)
        offs = [0, self.widths[0], self.widths[0] + self.widths[1], self.widths[0] + self.widths[1] + self.widths[2]]
        raws = []
        _, r = self.branch0[0].run(ctx, x, out=cat, coff=offs[0]); raws.append(r)
        h, _ = self.branch1[0].run(ctx, x)
        _, r = self.branch1[1].run(ctx, h, out=cat, coff=offs[1]); raws.append(r)
        h, _ = self.branch2[0].run(ctx, x)
        _, r = self.branch2[1].run(ctx, h, out=cat, coff=offs[2]); raws.append(r)
        p = E.max_pool(ctx, x, (3, 3, 3), (1, 1, 1), (1, 1, 1))
        _, r = self.branch3[1].run(ctx, p, out=cat, coff=offs[3]); raws.append(r)
        if self.gating:
            for off, r, g in zip(offs, raws, (self.gating_b0, self.gating_b1, self.gating_b2, self.gating_b3)):
                E.self_gate(ctx, cat, off, r, g.fc)
        return cat


_MIXED = [("Mixed_3b", 192, [64, 96, 128, 16, 32, 32]), ("Mixed_3c", 256, [128, 128, 192, 32, 96, 64]),
          ("Mixed_4b", 480, [192, 96, 208, 16, 48, 64]), ("Mixed_4c", 512, [160, 112, 224, 24, 64, 64]),
          ("Mixed_4d", 512, [128, 128, 256, 24, 64, 64]), ("Mixed_4e", 512, [112, 144, 288, 32, 64, 64]),
          ("Mixed_4f", 528, [256, 160, 320, 32, 128, 128]), ("Mixed_5b", 832, [256, 160, 320, 32, 128, 128]),
          ("Mixed_5c", 832, [384, 192, 384, 48, 128, 128])]


class S3D(_Encoder):
    """backbone/s3dg.py:135-217 (modules registered both by name and inside blockN, like the reference)."""

    def __init__(self, input_channel=3, gating=False, slow=False):
        super().__init__()
        self.gating, self.slow = gating, slow
        self.Conv_1a = STConv3d(input_channel, 64, kernel_size=7, stride=(1, 2, 2) if slow else 2, padding=3)
        self.block1 = nn.Sequential(self.Conv_1a)
        self.MaxPool_2a = nn.MaxPool3d(kernel_size=(1, 3, 3), stride=(1, 2, 2), padding=(0, 1, 1))
        self.Conv_2b = BasicConv3d(64, 64, kernel_size=1, stride=1)
        self.Conv_2c = STConv3d(64, 192, kernel_size=3, stride=1, padding=1)
        self.block2 = nn.Sequential(self.MaxPool_2a, self.Conv_2b, self.Conv_2c)
        self.MaxPool_3a = nn.MaxPool3d(kernel_size=(1, 3, 3), stride=(1, 2, 2), padding=(0, 1, 1))
        mixed = {}
        for name, cin, planes in _MIXED[:2]:
            mixed[name] = SepInception(cin, planes, gating=gating)
            setattr(self, name, mixed[name])
        self.block3 = nn.Sequential(self.MaxPool_3a, mixed["Mixed_3b"], mixed["Mixed_3c"])
        self.MaxPool_4a = nn.MaxPool3d(kernel_size=(3, 3, 3), stride=(2, 2, 2), padding=(1, 1, 1))
        for name, cin, planes in _MIXED[2:7]:
            mixed[name] = SepInception(cin, planes, gating=gating)
            setattr(self, name, mixed[name])
        self.block4 = nn.Sequential(self.MaxPool_4a, *[mixed[n] for n, _, _ in _MIXED[2:7]])
        self.MaxPool_5a = nn.MaxPool3d(kernel_size=(2, 2, 2), stride=(2, 2, 2), padding=(0, 0, 0))
        for name, cin, planes in _MIXED[7:]:
            mixed[name] = SepInception(cin, planes, gating=gating)
            setattr(self, name, mixed[name])
        self.block5 = nn.Sequential(self.MaxPool_5a, mixed["Mixed_5b"], mixed["Mixed_5c"])

    def first_conv(self):
        return self.Conv_1a.conv1

    @staticmethod
    def _pool(ctx, x, mod):
        t3 = lambda v: (v, v, v) if isinstance(v, int) else tuple(v)  # noqa: E731
        return E.max_pool(ctx, x, t3(mod.kernel_size), t3(mod.stride), t3(mod.padding))

    def program(self, ctx, x):
        x, _ = self.Conv_1a.run(ctx, x)
        x = self._pool(ctx, x, self.MaxPool_2a)
        x, _ = self.Conv_2b.run(ctx, x)
        x, _ = self.Conv_2c.run(ctx, x)
        x = self._pool(ctx, x, self.MaxPool_3a)
        x = self.Mixed_3b.run(ctx, x)
        x = self.Mixed_3c.run(ctx, x)
        x = self._pool(ctx, x, self.MaxPool_4a)
        for name in ("Mixed_4b", "Mixed_4c", "Mixed_4d", "Mixed_4e", "Mixed_4f"):
            x = getattr(self, name).run(ctx, x)
        x = self._pool(ctx, x, self.MaxPool_5a)
        x = self.Mixed_5b.run(ctx, x)
        return self.Mixed_5c.run(ctx, x)

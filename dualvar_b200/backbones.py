"""Drop-in clip encoders: same constructors, attribute names and ``state_dict`` keys as the
reference's backbone/*.py, but ``forward`` runs the sm_100a kernels through dualvar_b200.engine.

The nn.Conv3d / nn.BatchNorm3d objects below are PARAMETER CONTAINERS only (their own forward is
never called): using them keeps initialisation, state_dict layout, ``.cuda()``, DDP and
``nn.SyncBatchNorm.convert_sync_batchnorm`` behaving exactly as with the reference modules
(pretrain.py:244-248). Module contract (backbone/select_backbone.py:30-31): input fp32
(B, 3, T, H, W), output fp32 (B, feature_size, T', H', W') after ReLU.

Reference: backbone/select_backbone.py:7-32, backbone/r21d.py, backbone/r3d.py, backbone/c3d.py,
backbone/s3dg.py.
"""
import math

import torch
import torch.nn as nn

from . import engine as E


def _t3(v):
    return (v, v, v) if isinstance(v, int) else tuple(v)


class _Encoder(nn.Module):
    """Shared plumbing: fp32 NCDHW module contract on top of an engine program."""

    def program(self, ctx, x):  # pragma: no cover - abstract
        raise NotImplementedError

    def first_conv(self):
        """The nn.Conv3d applied to the raw frames (decides the ingest layout)."""
        return None

    def wants_s2d(self, src):
        shape = src.block_shape if isinstance(src, E.RawClips) else tuple(src.shape)
        conv = self.first_conv()
        return conv is not None and E.stem_eligible(conv, shape[-2], shape[-1])

    def encode(self, src, pooled, make_input=None, **ingest_kw):
        """Ingest ``src`` (reference block, clip batch or RawClips) and run the backbone.
        pooled=True returns the (N, C) global average instead of the fp32 NCDHW feature map.
        ``make_input`` (optional) builds the input Act itself (e.g. several ingests into one batch)."""
        if make_input is None:
            s2d = self.wants_s2d(src)
            make_input = lambda: E.ingest(src, s2d=s2d, **ingest_kw)  # noqa: E731
        return E.run_backbone(self, self.program, make_input, pooled)

    def encode_pair(self, src, kw_a, kw_b):
        """Pooled features of two independent passes over ``src`` (different ingest arguments) issued on two streams
        from one autograd node (engine.BackbonePairFunction)."""
        s2d = self.wants_s2d(src)
        mk = [lambda kw=kw: E.ingest(src, s2d=s2d, **kw) for kw in (kw_a, kw_b)]
        return E.run_backbone_pair(self, self.program, mk)

    def forward(self, x):
        if not x.is_cuda:
            raise E._lib.DualVarNativeError("dualvar_b200 backbones run on a B200 only (no CPU fallback)")
        return self.encode(x.contiguous().float(), pooled=False)


# ------------------------------------------------------------------------------------ R(2+1)D / R3D
class SpatioTemporalConv(nn.Module):
    """Factorised (2+1)D conv container (backbone/r21d.py:25-70)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=False, first_conv=False):
        super().__init__()
        k, s, p = _t3(kernel_size), _t3(stride), _t3(padding)
        mid = int(math.floor((k[0] * k[1] * k[2] * in_channels * out_channels) /
                             (k[1] * k[2] * in_channels + k[0] * out_channels)))
        self.spatial_conv = nn.Conv3d(in_channels, mid, (1, k[1], k[2]), stride=(1, s[1], s[2]),
                                      padding=(0, p[1], p[2]), bias=bias)
        self.bn = nn.BatchNorm3d(mid)
        self.relu = nn.ReLU()
        self.temporal_conv = nn.Conv3d(mid, out_channels, (k[0], 1, 1), stride=(s[0], 1, 1),
                                       padding=(p[0], 0, 0), bias=bias)

    def run(self, ctx, x, out_bn):
        """spatial conv -> BN -> ReLU -> temporal conv; returns the raw output paired with out_bn."""
        # the BatchNorm + ReLU between the two convs feeds the temporal conv only: applied inside it (consumer-side)
        h = E.activate(ctx, E.conv_stats(ctx, x, self.spatial_conv, self.bn), conv_only=True)
        return E.conv_stats(ctx, h, self.temporal_conv, out_bn)


class FullSpatioTemporalConv(nn.Module):
    """Single full 3-D conv container (backbone/r3d.py:24-38)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=False):
        super().__init__()
        self.temporal_spatial_conv = nn.Conv3d(in_channels, out_channels, _t3(kernel_size), stride=_t3(stride),
                                               padding=_t3(padding), bias=bias)

    def run(self, ctx, x, out_bn):
        return E.conv_stats(ctx, x, self.temporal_spatial_conv, out_bn)


class SpatioTemporalResBlock(nn.Module):
    """Residual block (backbone/r21d.py:83-122, backbone/r3d.py:51-89)."""

    conv_cls = SpatioTemporalConv

    def __init__(self, in_channels, out_channels, kernel_size, downsample=False):
        super().__init__()
        self.downsample = downsample
        padding = kernel_size // 2
        cc = self.conv_cls
        if self.downsample:
            self.downsampleconv = cc(in_channels, out_channels, 1, stride=2)
            self.downsamplebn = nn.BatchNorm3d(out_channels)
            self.conv1 = cc(in_channels, out_channels, kernel_size, padding=padding, stride=2)
        else:
            self.conv1 = cc(in_channels, out_channels, kernel_size, padding=padding)
        self.bn1 = nn.BatchNorm3d(out_channels)
        self.relu1 = nn.ReLU()
        self.conv2 = cc(out_channels, out_channels, kernel_size, padding=padding)
        self.bn2 = nn.BatchNorm3d(out_channels)
        self.outrelu = nn.ReLU()

    def run(self, ctx, x):
        res = E.activate(ctx, self.conv1.run(ctx, x, self.bn1))
        main = self.conv2.run(ctx, res, self.bn2)
        if self.downsample:
            short = self.downsampleconv.run(ctx, x, self.downsamplebn)
            return E.activate(ctx, main, r2=short)          # relu(bn2(.) + downsamplebn(.))
        return E.activate(ctx, main, res=x)                  # relu(bn2(.) + x)


class R3DResBlock(SpatioTemporalResBlock):
    conv_cls = FullSpatioTemporalConv


class SpatioTemporalResLayer(nn.Module):
    """block1 + (layer_size-1) identity blocks (backbone/r21d.py:188-206)."""

    def __init__(self, in_channels, out_channels, kernel_size, layer_size, block_type=SpatioTemporalResBlock,
                 downsample=False):
        super().__init__()
        self.block1 = block_type(in_channels, out_channels, kernel_size, downsample)
        self.blocks = nn.ModuleList([])
        for _ in range(layer_size - 1):
            self.blocks += [block_type(out_channels, out_channels, kernel_size)]

    def run(self, ctx, x):
        x = self.block1.run(ctx, x)
        for b in self.blocks:
            x = b.run(ctx, x)
        return x


class R2Plus1DNet(_Encoder):
    """backbone/r21d.py:214-266. ``layer_sizes=(1,1,1,1)`` is what select_backbone('r21d') builds
    (14.4 M parameters); (2,2,2,2) is the 18-layer network."""

    conv_cls = SpatioTemporalConv

    def __init__(self, layer_sizes=(1, 1, 1, 1), block_type=None):
        super().__init__()
        block_type = block_type or (SpatioTemporalResBlock if self.conv_cls is SpatioTemporalConv else R3DResBlock)
        self.conv1 = self.conv_cls(3, 64, (3, 7, 7), stride=(1, 2, 2), padding=(1, 3, 3))
        self.bn1 = nn.BatchNorm3d(64)
        self.relu1 = nn.ReLU()
        self.conv2 = SpatioTemporalResLayer(64, 64, 3, layer_sizes[0], block_type=block_type)
        self.conv3 = SpatioTemporalResLayer(64, 128, 3, layer_sizes[1], block_type=block_type, downsample=True)
        self.conv4 = SpatioTemporalResLayer(128, 256, 3, layer_sizes[2], block_type=block_type, downsample=True)
        self.conv5 = SpatioTemporalResLayer(256, 512, 3, layer_sizes[3], block_type=block_type, downsample=True)

    def first_conv(self):
        c = self.conv1
        return c.spatial_conv if hasattr(c, "spatial_conv") else c.temporal_spatial_conv

    def program(self, ctx, x):
        x = E.activate(ctx, self.conv1.run(ctx, x, self.bn1))
        for layer in (self.conv2, self.conv3, self.conv4, self.conv5):
            x = layer.run(ctx, x)
        return x

    def forward(self, x, ret_frame_feature=False, multi_level=False, aug_feature_lvls=[], aug_prob=0.5,
                aug_range=-1):
        if ret_frame_feature:
            raise NotImplementedError(
                "ret_frame_feature is only used by SimCLR_Naked.get_features (model/simclr.py:123-127), "
                "a visualisation helper outside the pretraining hot path")
        return super().forward(x)


class R3DNet(R2Plus1DNet):
    """backbone/r3d.py:126-157."""

    conv_cls = FullSpatioTemporalConv

    def forward(self, x):
        return _Encoder.forward(self, x)


# ------------------------------------------------------------------------------------ C3D
class C3D(_Encoder):
    """C3D with BN (backbone/c3d.py:9-83): conv(bias) -> BN -> ReLU x8, four max-pools."""

    _PLAN = [("1", 3, 64, (1, 2, 2)), ("2", 64, 128, (2, 2, 2)), ("3a", 128, 256, None),
             ("3b", 256, 256, (2, 2, 2)), ("4a", 256, 512, None), ("4b", 512, 512, (2, 2, 2)),
             ("5a", 512, 512, None), ("5b", 512, 512, None)]

    def __init__(self):
        super().__init__()
        for tag, cin, cout, pool in self._PLAN:
            setattr(self, "conv" + tag, nn.Conv3d(cin, cout, kernel_size=(3, 3, 3), padding=(1, 1, 1)))
            setattr(self, "bn" + tag, nn.BatchNorm3d(cout))
            setattr(self, "relu" + tag, nn.ReLU())
            if pool is not None:
                setattr(self, "pool" + tag.rstrip("ab"), nn.MaxPool3d(kernel_size=pool, stride=pool))

    def program(self, ctx, x):
        for tag, _, _, pool in self._PLAN:
            x = E.activate(ctx, E.conv_stats(ctx, x, getattr(self, "conv" + tag), getattr(self, "bn" + tag)))
            if pool is not None:
                x = E.max_pool(ctx, x, pool, pool, (0, 0, 0))
        return x


# ------------------------------------------------------------------------------------ r2d3d18
class BasicBlock2d(nn.Module):
    """2-D basic block container (backbone/resnet_2d3d.py:45-78)."""

    def __init__(self, inplanes, planes, stride=1, downsample=None, use_final_relu=True):
        super().__init__()
        self.use_final_relu = use_final_relu
        self.conv1 = nn.Conv3d(inplanes, planes, (1, 3, 3), stride=(1, stride, stride), padding=(0, 1, 1), bias=False)
        self.bn1 = nn.BatchNorm3d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv3d(planes, planes, (1, 3, 3), stride=1, padding=(0, 1, 1), bias=False)
        self.bn2 = nn.BatchNorm3d(planes)
        self.downsample = downsample

    def run(self, ctx, x):
        h = E.activate(ctx, E.conv_stats(ctx, x, self.conv1, self.bn1))
        main = E.conv_stats(ctx, h, self.conv2, self.bn2)
        if self.downsample is not None:
            short = E.conv_stats(ctx, x, self.downsample[0], self.downsample[1])
            return E.activate(ctx, main, r2=short, relu=self.use_final_relu)
        return E.activate(ctx, main, res=x, relu=self.use_final_relu)


class ResNet2d3dFull(_Encoder):
    """select_backbone('r2d3d18') (backbone/resnet_2d3d.py:193-271,352-356): stem (1,7,7)/(1,2,2) + max-pool,
    four stages of two 2-D basic blocks, no ReLU after the last block."""

    def __init__(self, layers=(2, 2, 2, 2)):
        super().__init__()
        self.inplanes = 64
        self.conv1 = nn.Conv3d(3, 64, (1, 7, 7), stride=(1, 2, 2), padding=(0, 3, 3), bias=False)
        self.bn1 = nn.BatchNorm3d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool3d((1, 3, 3), stride=(1, 2, 2), padding=(0, 1, 1))
        self.layer1 = self._make_layer(64, layers[0])
        self.layer2 = self._make_layer(128, layers[1], stride=2)
        self.layer3 = self._make_layer(256, layers[2], stride=2)
        self.layer4 = self._make_layer(256, layers[3], stride=2, is_final=True)
        for m in self.modules():                      # backbone/resnet_2d3d.py:214-220
            if isinstance(m, nn.Conv3d):
                m.weight = nn.init.kaiming_normal_(m.weight, mode="fan_out")
            elif isinstance(m, nn.BatchNorm3d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    def _make_layer(self, planes, blocks, stride=1, is_final=False):
        downsample = None
        if stride != 1 or self.inplanes != planes:
            downsample = nn.Sequential(
                nn.Conv3d(self.inplanes, planes, 1, stride=(1, stride, stride), bias=False), nn.BatchNorm3d(planes))
        layers = [BasicBlock2d(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes
        for i in range(1, blocks):
            layers.append(BasicBlock2d(planes, planes, use_final_relu=not (is_final and i == blocks - 1)))
        return nn.Sequential(*layers)

    def first_conv(self):
        return self.conv1

    def program(self, ctx, x):
        x = E.activate(ctx, E.conv_stats(ctx, x, self.conv1, self.bn1))
        x = E.max_pool(ctx, x, (1, 3, 3), (1, 2, 2), (0, 1, 1))
        for layer in (self.layer1, self.layer2, self.layer3, self.layer4):
            for block in layer:
                x = block.run(ctx, x)
        return x


def select_backbone(network, first_channel=3):
    """backbone/select_backbone.py:7-32 — same names, same return value."""
    param = {'feature_size': 1024}
    if network == 'c3d':
        model = C3D()
        param['feature_size'] = 512
    elif network == 'r21d':
        param['feature_size'] = 512
        model = R2Plus1DNet()
    elif network == 'r3d':
        param['feature_size'] = 512
        model = R3DNet()
    elif network in ('s3d', 's3dg'):
        from .s3dg import S3D
        model = S3D(input_channel=first_channel, gating=(network == 's3dg'))
    elif network == 'r2d3d18':
        param['feature_size'] = 256
        model = ResNet2d3dFull()
    else:
        # 'r50' raises TypeError at construction in the reference itself (backbone/select_backbone.py:18 passes
        # input_channel to a constructor that does not take it)
        raise NotImplementedError
    return model, param

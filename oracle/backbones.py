"""ORACLE (test infrastructure): plain-torch fp32 restatement of the reference clip encoders.

Module/attribute names reproduce the reference's ``state_dict`` keys so weights interchange with
the reference and with ``dualvar_b200`` both ways; submodules are created in the reference's order
so that construction under the same ``torch.manual_seed`` draws the same initial weights.

Reference: backbone/select_backbone.py:7-32, backbone/r21d.py, backbone/r3d.py, backbone/c3d.py,
backbone/s3dg.py (paths under /root/reference).
"""
import torch
import torch.nn as nn


def _t3(v):
    return (v, v, v) if isinstance(v, int) else tuple(v)


def r21d_mid_channels(cin, cout, k):
    """Intermediate width of a factorised conv (backbone/r21d.py:47-49)."""
    kt, kh, kw = k
    return int((kt * kh * kw * cin * cout) // (kh * kw * cin + kt * cout))


class FactoredConv(nn.Module):
    """(1,k,k) conv -> BN -> ReLU -> (k,1,1) conv, no bias (backbone/r21d.py:25-70)."""

    def __init__(self, cin, cout, kernel, stride=1, padding=0):
        super().__init__()
        k, s, p = _t3(kernel), _t3(stride), _t3(padding)
        mid = r21d_mid_channels(cin, cout, k)
        self.spatial_conv = nn.Conv3d(cin, mid, (1, k[1], k[2]), stride=(1, s[1], s[2]),
                                      padding=(0, p[1], p[2]), bias=False)
        self.bn = nn.BatchNorm3d(mid)
        self.relu = nn.ReLU()
        self.temporal_conv = nn.Conv3d(mid, cout, (k[0], 1, 1), stride=(s[0], 1, 1),
                                       padding=(p[0], 0, 0), bias=False)

    def forward(self, x):
        return self.temporal_conv(self.relu(self.bn(self.spatial_conv(x))))


class FullConv(nn.Module):
    """Single full 3-D conv, no bias (backbone/r3d.py:24-38)."""

    def __init__(self, cin, cout, kernel, stride=1, padding=0):
        super().__init__()
        self.temporal_spatial_conv = nn.Conv3d(cin, cout, _t3(kernel), stride=_t3(stride),
                                               padding=_t3(padding), bias=False)

    def forward(self, x):
        return self.temporal_spatial_conv(x)


class ResBlock(nn.Module):
    """conv1-bn1-relu-conv2-bn2 (+ 1x1x1 stride-2 conv+bn shortcut) -> add -> relu.
    backbone/r21d.py:83-122 and backbone/r3d.py:51-89 share this skeleton."""

    def __init__(self, conv_cls, cin, cout, kernel, downsample=False):
        super().__init__()
        self.downsample = downsample
        pad = kernel // 2
        if downsample:
            self.downsampleconv = conv_cls(cin, cout, 1, stride=2)
            self.downsamplebn = nn.BatchNorm3d(cout)
            self.conv1 = conv_cls(cin, cout, kernel, padding=pad, stride=2)
        else:
            self.conv1 = conv_cls(cin, cout, kernel, padding=pad)
        self.bn1 = nn.BatchNorm3d(cout)
        self.relu1 = nn.ReLU()
        self.conv2 = conv_cls(cout, cout, kernel, padding=pad)
        self.bn2 = nn.BatchNorm3d(cout)
        self.outrelu = nn.ReLU()

    def forward(self, x):
        res = self.relu1(self.bn1(self.conv1(x)))
        res = self.bn2(self.conv2(res))
        if self.downsample:
            x = self.downsamplebn(self.downsampleconv(x))
        return self.outrelu(x + res)


class ResLayer(nn.Module):
    """block1 + (layer_size - 1) identity blocks (backbone/r21d.py:188-206)."""

    def __init__(self, conv_cls, cin, cout, kernel, layer_size, downsample=False):
        super().__init__()
        self.block1 = ResBlock(conv_cls, cin, cout, kernel, downsample)
        self.blocks = nn.ModuleList([ResBlock(conv_cls, cout, cout, kernel) for _ in range(layer_size - 1)])

    def forward(self, x):
        x = self.block1(x)
        for b in self.blocks:
            x = b(x)
        return x


class _ResNet3D(nn.Module):
    def __init__(self, conv_cls, layer_sizes):
        super().__init__()
        self.conv1 = conv_cls(3, 64, (3, 7, 7), stride=(1, 2, 2), padding=(1, 3, 3))
        self.bn1 = nn.BatchNorm3d(64)
        self.relu1 = nn.ReLU()
        self.conv2 = ResLayer(conv_cls, 64, 64, 3, layer_sizes[0])
        self.conv3 = ResLayer(conv_cls, 64, 128, 3, layer_sizes[1], downsample=True)
        self.conv4 = ResLayer(conv_cls, 128, 256, 3, layer_sizes[2], downsample=True)
        self.conv5 = ResLayer(conv_cls, 256, 512, 3, layer_sizes[3], downsample=True)

    def stages(self, x):
        x = self.relu1(self.bn1(self.conv1(x)))
        feats = []
        for layer in (self.conv2, self.conv3, self.conv4, self.conv5):
            x = layer(x)
            feats.append(x)
        return x, feats


class R2Plus1DNet(_ResNet3D):
    """backbone/r21d.py:214-266; default (1,1,1,1) is what select_backbone('r21d') builds."""

    def __init__(self, layer_sizes=(1, 1, 1, 1)):
        super().__init__(FactoredConv, layer_sizes)

    def forward(self, x, ret_frame_feature=False, multi_level=False, aug_feature_lvls=[], aug_prob=0.5,
                aug_range=-1):
        x, feats = self.stages(x)
        if not ret_frame_feature:
            return x
        return (x, feats) if multi_level else (x, feats[0])


class R3DNet(_ResNet3D):
    """backbone/r3d.py:126-157."""

    def __init__(self, layer_sizes=(1, 1, 1, 1)):
        super().__init__(FullConv, layer_sizes)

    def forward(self, x):
        return self.stages(x)[0]


class C3D(nn.Module):
    """Eight 3x3x3 convs (bias=True) each followed by BN+ReLU, four max-pools (backbone/c3d.py:12-83)."""

    PLAN = [("1", 3, 64, (1, 2, 2)), ("2", 64, 128, (2, 2, 2)), ("3a", 128, 256, None),
            ("3b", 256, 256, (2, 2, 2)), ("4a", 256, 512, None), ("4b", 512, 512, (2, 2, 2)),
            ("5a", 512, 512, None), ("5b", 512, 512, None)]

    def __init__(self):
        super().__init__()
        for tag, cin, cout, pool in self.PLAN:
            setattr(self, "conv" + tag, nn.Conv3d(cin, cout, kernel_size=(3, 3, 3), padding=(1, 1, 1)))
            setattr(self, "bn" + tag, nn.BatchNorm3d(cout))
            setattr(self, "relu" + tag, nn.ReLU())
            if pool is not None:
                setattr(self, "pool" + tag.rstrip("ab"), nn.MaxPool3d(kernel_size=pool, stride=pool))

    def forward(self, x):
        for tag, _, _, pool in self.PLAN:
            x = getattr(self, "relu" + tag)(getattr(self, "bn" + tag)(getattr(self, "conv" + tag)(x)))
            if pool is not None:
                x = getattr(self, "pool" + tag.rstrip("ab"))(x)
        return x


class BasicConv3d(nn.Module):
    """conv(no bias, N(0,0.01) init) -> BN -> ReLU (backbone/s3dg.py:8-28)."""

    def __init__(self, cin, cout, kernel_size, stride, padding=0):
        super().__init__()
        self.conv = nn.Conv3d(cin, cout, kernel_size=kernel_size, stride=stride, padding=padding, bias=False)
        self.bn = nn.BatchNorm3d(cout)
        self.relu = nn.ReLU(inplace=True)
        self.conv.weight.data.normal_(mean=0, std=0.01)
        self.bn.weight.data.fill_(1)
        self.bn.bias.data.zero_()

    def forward(self, x):
        return self.relu(self.bn(self.conv(x)))


class STConv3d(nn.Module):
    """(1,k,k) conv-BN-ReLU then (k,1,1) conv-BN-ReLU (backbone/s3dg.py:30-65)."""

    def __init__(self, cin, cout, kernel_size, stride, padding=0):
        super().__init__()
        if isinstance(stride, tuple):
            t_stride, stride = stride[0], stride[-1]
        else:
            t_stride = stride
        self.conv1 = nn.Conv3d(cin, cout, kernel_size=(1, kernel_size, kernel_size),
                               stride=(1, stride, stride), padding=(0, padding, padding), bias=False)
        self.conv2 = nn.Conv3d(cout, cout, kernel_size=(kernel_size, 1, 1), stride=(t_stride, 1, 1),
                               padding=(padding, 0, 0), bias=False)
        self.bn1 = nn.BatchNorm3d(cout)
        self.bn2 = nn.BatchNorm3d(cout)
        self.relu = nn.ReLU(inplace=True)
        self.conv1.weight.data.normal_(mean=0, std=0.01)
        self.conv2.weight.data.normal_(mean=0, std=0.01)
        for bn in (self.bn1, self.bn2):
            bn.weight.data.fill_(1)
            bn.bias.data.zero_()

    def forward(self, x):
        x = self.relu(self.bn1(self.conv1(x)))
        return self.relu(self.bn2(self.conv2(x)))


class SelfGating(nn.Module):
    """channel gate = sigmoid(Linear(mean over T,H,W)) (backbone/s3dg.py:68-78)."""

    def __init__(self, dim):
        super().__init__()
        self.fc = nn.Linear(dim, dim)

    def forward(self, x):
        w = torch.sigmoid(self.fc(x.mean(dim=[2, 3, 4])))
        return w[:, :, None, None, None] * x


class SepInception(nn.Module):
    """Four-branch separable Inception block with optional gating (backbone/s3dg.py:81-132)."""

    def __init__(self, cin, out_planes, gating=False):
        super().__init__()
        o0, o1a, o1b, o2a, o2b, o3 = out_planes
        self.branch0 = nn.Sequential(BasicConv3d(cin, o0, kernel_size=1, stride=1))
        self.branch1 = nn.Sequential(BasicConv3d(cin, o1a, kernel_size=1, stride=1),
                                     STConv3d(o1a, o1b, kernel_size=3, stride=1, padding=1))
        self.branch2 = nn.Sequential(BasicConv3d(cin, o2a, kernel_size=1, stride=1),
                                     STConv3d(o2a, o2b, kernel_size=3, stride=1, padding=1))
        self.branch3 = nn.Sequential(nn.MaxPool3d(kernel_size=(3, 3, 3), stride=1, padding=1),
                                     BasicConv3d(cin, o3, kernel_size=1, stride=1))
        self.out_channels = o0 + o1b + o2b + o3
        self.gating = gating
        if gating:
            self.gating_b0 = SelfGating(o0)
            self.gating_b1 = SelfGating(o1b)
            self.gating_b2 = SelfGating(o2b)
            self.gating_b3 = SelfGating(o3)

    def forward(self, x):
        outs = [self.branch0(x), self.branch1(x), self.branch2(x), self.branch3(x)]
        if self.gating:
            gates = (self.gating_b0, self.gating_b1, self.gating_b2, self.gating_b3)
            outs = [g(o) for g, o in zip(gates, outs)]
        return torch.cat(outs, 1)


S3D_MIXED = [  # name, in_planes, out_planes  (backbone/s3dg.py:163-193)
    ("Mixed_3b", 192, [64, 96, 128, 16, 32, 32]), ("Mixed_3c", 256, [128, 128, 192, 32, 96, 64]),
    ("Mixed_4b", 480, [192, 96, 208, 16, 48, 64]), ("Mixed_4c", 512, [160, 112, 224, 24, 64, 64]),
    ("Mixed_4d", 512, [128, 128, 256, 24, 64, 64]), ("Mixed_4e", 512, [112, 144, 288, 32, 64, 64]),
    ("Mixed_4f", 528, [256, 160, 320, 32, 128, 128]), ("Mixed_5b", 832, [256, 160, 320, 32, 128, 128]),
    ("Mixed_5c", 832, [384, 192, 384, 48, 128, 128]),
]


class S3D(nn.Module):
    """S3D / S3D-G (backbone/s3dg.py:135-217). Registered twice like the reference: once by its
    own name and once inside the ``blockN`` Sequentials, so state_dict carries both key sets."""

    def __init__(self, input_channel=3, gating=False, slow=False):
        super().__init__()
        self.gating = gating
        self.slow = slow
        self.Conv_1a = STConv3d(input_channel, 64, kernel_size=7, stride=(1, 2, 2) if slow else 2, padding=3)
        self.block1 = nn.Sequential(self.Conv_1a)
        self.MaxPool_2a = nn.MaxPool3d(kernel_size=(1, 3, 3), stride=(1, 2, 2), padding=(0, 1, 1))
        self.Conv_2b = BasicConv3d(64, 64, kernel_size=1, stride=1)
        self.Conv_2c = STConv3d(64, 192, kernel_size=3, stride=1, padding=1)
        self.block2 = nn.Sequential(self.MaxPool_2a, self.Conv_2b, self.Conv_2c)
        self.MaxPool_3a = nn.MaxPool3d(kernel_size=(1, 3, 3), stride=(1, 2, 2), padding=(0, 1, 1))
        mixed = {}
        for name, cin, planes in S3D_MIXED[:2]:
            mixed[name] = SepInception(cin, planes, gating=gating)
            setattr(self, name, mixed[name])
        self.block3 = nn.Sequential(self.MaxPool_3a, mixed["Mixed_3b"], mixed["Mixed_3c"])
        self.MaxPool_4a = nn.MaxPool3d(kernel_size=(3, 3, 3), stride=(2, 2, 2), padding=(1, 1, 1))
        for name, cin, planes in S3D_MIXED[2:7]:
            mixed[name] = SepInception(cin, planes, gating=gating)
            setattr(self, name, mixed[name])
        self.block4 = nn.Sequential(self.MaxPool_4a, *[mixed[n] for n, _, _ in S3D_MIXED[2:7]])
        self.MaxPool_5a = nn.MaxPool3d(kernel_size=(2, 2, 2), stride=(2, 2, 2), padding=(0, 0, 0))
        for name, cin, planes in S3D_MIXED[7:]:
            mixed[name] = SepInception(cin, planes, gating=gating)
            setattr(self, name, mixed[name])
        self.block5 = nn.Sequential(self.MaxPool_5a, mixed["Mixed_5b"], mixed["Mixed_5c"])

    def forward(self, x):
        for blk in (self.block1, self.block2, self.block3, self.block4, self.block5):
            x = blk(x)
        return x


class BasicBlock2d(nn.Module):
    """(1,3,3) conv -> BN -> ReLU -> (1,3,3) conv -> BN (+ 1x1x1 strided conv + BN shortcut) -> add -> ReLU?
    (backbone/resnet_2d3d.py:45-78)."""

    def __init__(self, inplanes, planes, stride=1, downsample=None, use_final_relu=True):
        super().__init__()
        self.use_final_relu = use_final_relu
        self.conv1 = nn.Conv3d(inplanes, planes, (1, 3, 3), stride=(1, stride, stride), padding=(0, 1, 1), bias=False)
        self.bn1 = nn.BatchNorm3d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv3d(planes, planes, (1, 3, 3), stride=1, padding=(0, 1, 1), bias=False)
        self.bn2 = nn.BatchNorm3d(planes)
        self.downsample = downsample

    def forward(self, x):
        out = self.bn2(self.conv2(self.relu(self.bn1(self.conv1(x)))))
        out = out + (x if self.downsample is None else self.downsample(x))
        return self.relu(out) if self.use_final_relu else out


class ResNet2d3dFull(nn.Module):
    """CVRL-style ResNet of 2-D basic blocks = select_backbone('r2d3d18') (backbone/resnet_2d3d.py:193-271,352-356):
    (1,7,7)/(1,2,2) stem -> BN -> ReLU -> max-pool (1,3,3)/(1,2,2) -> 4 stages of 2 blocks (64,128,256,256 planes,
    stages 2-4 stride 2 in H,W only); the last block has no output ReLU; kaiming-normal(fan_out) conv init."""

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

    def forward(self, x):
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        return self.layer4(self.layer3(self.layer2(self.layer1(x))))


def select_backbone(network, first_channel=3):
    """Name -> (module, {'feature_size'}) (backbone/select_backbone.py:7-32). 'r50' is broken in the reference
    itself (TypeError at construction) and raises like an unknown name."""
    param = {"feature_size": 1024}
    if network == "s3d":
        model = S3D(input_channel=first_channel)
    elif network == "c3d":
        model = C3D()
        param["feature_size"] = 512
    elif network == "s3dg":
        model = S3D(input_channel=first_channel, gating=True)
    elif network == "r21d":
        param["feature_size"] = 512
        model = R2Plus1DNet()
    elif network == "r3d":
        param["feature_size"] = 512
        model = R3DNet()
    elif network == "r2d3d18":
        param["feature_size"] = 256
        model = ResNet2d3dFull()
    else:
        raise NotImplementedError
    return model, param

"""torch-fp32 restatement of the network the reference runs (TEST INFRASTRUCTURE).

Reference: inference/inference.py:190-197 builds
``BasicUNet(spatial_dims=3, in_channels=1, out_channels=1,
features=(32, 32, 64, 128, 256, 32), dropout=0.1, act="mish")`` from
MONAI 1.2.0 (requirements.txt:21, not vendored).  MONAI's published block
structure, restated:

* ``TwoConv``  = 2 x [Conv3d(k3, p1, bias) -> InstanceNorm3d(affine, eps 1e-5)
                 -> Dropout3d(0.1) -> Mish]                (ADN order "NDA")
* ``Down``     = MaxPool3d(2) -> TwoConv
* ``UpCat``    = ConvTranspose3d(k2, s2, bias) -> cat([skip, up], 1) -> TwoConv
* ``final``    = Conv3d(32 -> 1, k1)

Parameter names mirror MONAI's so that the shipped checkpoint
(models/inference_weights.tar, 82 tensors, ``module.`` prefix from
DataParallel, inference.py:217-222) loads with ``strict=True`` - that strict
load is the structural pin for this restatement.
"""
from collections import OrderedDict

import torch
import torch.nn as nn

FEATURES = (32, 32, 64, 128, 256, 32)


class _ADN(nn.Sequential):
    def __init__(self, ch, dropout):
        super().__init__(OrderedDict([
            ("N", nn.InstanceNorm3d(ch, eps=1e-5, affine=True)),
            ("D", nn.Dropout3d(dropout)),
            ("A", nn.Mish()),
        ]))


class _ConvBlock(nn.Sequential):
    def __init__(self, cin, cout, dropout):
        super().__init__(OrderedDict([
            ("conv", nn.Conv3d(cin, cout, kernel_size=3, stride=1, padding=1, bias=True)),
            ("adn", _ADN(cout, dropout)),
        ]))


class TwoConv(nn.Sequential):
    def __init__(self, cin, cout, dropout):
        super().__init__(OrderedDict([
            ("conv_0", _ConvBlock(cin, cout, dropout)),
            ("conv_1", _ConvBlock(cout, cout, dropout)),
        ]))


class Down(nn.Sequential):
    def __init__(self, cin, cout, dropout):
        super().__init__(OrderedDict([
            ("max_pooling", nn.MaxPool3d(kernel_size=2)),
            ("convs", TwoConv(cin, cout, dropout)),
        ]))


class _UpSample(nn.Sequential):
    def __init__(self, cin, cout):
        super().__init__(OrderedDict([
            ("deconv", nn.ConvTranspose3d(cin, cout, kernel_size=2, stride=2, bias=True)),
        ]))


class UpCat(nn.Module):
    def __init__(self, in_chns, cat_chns, out_chns, dropout, halves=True):
        super().__init__()
        up_chns = in_chns // 2 if halves else in_chns
        self.upsample = _UpSample(in_chns, up_chns)
        self.convs = TwoConv(cat_chns + up_chns, out_chns, dropout)

    def forward(self, x, x_e):
        x_0 = self.upsample(x)
        # MONAI replicate-pads odd sizes; never hit when window dims % 16 == 0.
        pads = []
        for i in range(3):
            if x_e.shape[-i - 1] != x_0.shape[-i - 1]:
                pads += [0, 1]
            else:
                pads += [0, 0]
        if any(pads):
            x_0 = nn.functional.pad(x_0, pads, "replicate")
        return self.convs(torch.cat([x_e, x_0], dim=1))


class BasicUNet(nn.Module):
    """Signature follows monai.networks.nets.BasicUNet for the args the reference passes."""

    def __init__(self, spatial_dims=3, in_channels=1, out_channels=1, features=FEATURES,
                 act="mish", norm=("instance", {"affine": True}), bias=True, dropout=0.0,
                 upsample="deconv"):
        super().__init__()
        if spatial_dims != 3 or str(act).lower() != "mish" or upsample != "deconv":
            raise NotImplementedError("oracle restates only the configuration the reference uses")
        f = tuple(features)
        self.conv_0 = TwoConv(in_channels, f[0], dropout)
        self.down_1 = Down(f[0], f[1], dropout)
        self.down_2 = Down(f[1], f[2], dropout)
        self.down_3 = Down(f[2], f[3], dropout)
        self.down_4 = Down(f[3], f[4], dropout)
        self.upcat_4 = UpCat(f[4], f[3], f[3], dropout)
        self.upcat_3 = UpCat(f[3], f[2], f[2], dropout)
        self.upcat_2 = UpCat(f[2], f[1], f[1], dropout)
        self.upcat_1 = UpCat(f[1], f[0], f[5], dropout, halves=False)
        self.final_conv = nn.Conv3d(f[5], out_channels, kernel_size=1)

    def forward(self, x):
        x0 = self.conv_0(x)
        x1 = self.down_1(x0)
        x2 = self.down_2(x1)
        x3 = self.down_3(x2)
        x4 = self.down_4(x3)
        u4 = self.upcat_4(x4, x3)
        u3 = self.upcat_3(u4, x2)
        u2 = self.upcat_2(u3, x1)
        u1 = self.upcat_1(u2, x0)
        return self.final_conv(u1)


def strip_module_prefix(state_dict):
    """DataParallel checkpoint keys -> bare keys (inference.py:217-222)."""
    return OrderedDict((k[len("module."):] if k.startswith("module.") else k, v)
                       for k, v in state_dict.items())


def load_reference_net(weights_path):
    """Strict-load the shipped checkpoint into the restated net (eval mode, fp32)."""
    ck = torch.load(weights_path, map_location="cpu", weights_only=True)
    net = BasicUNet(dropout=0.1)
    net.load_state_dict(strip_module_prefix(ck["state_dict"]), strict=True)
    return net.eval()


def random_state_dict(seed=0):
    """Seeded random weights with the checkpoint's names/shapes and realistic scales.

    Used when the shipped checkpoint is not available (the GPU box has no
    /root/reference): same architecture, synthetic parameters.
    """
    g = torch.Generator().manual_seed(seed)
    net = BasicUNet(dropout=0.1)
    sd = OrderedDict()
    for k, v in net.state_dict().items():
        if k.endswith("adn.N.weight"):
            t = 1.0 + 0.05 * torch.randn(v.shape, generator=g)
        elif k.endswith("adn.N.bias"):
            t = 0.05 * torch.randn(v.shape, generator=g)
        elif k.endswith("bias"):
            t = 0.1 * torch.randn(v.shape, generator=g)
        else:
            fan_in = v[0].numel() if "deconv" not in k else v.shape[0]
            t = torch.randn(v.shape, generator=g) * (1.6 / fan_in) ** 0.5
        sd["module." + k] = t.to(torch.float32)
    return sd

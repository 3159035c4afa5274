"""tcgen05 convolution / transposed convolution kernels vs torch fp32 on the same bf16-rounded operands."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from gpu_common import bf16_round, ctx_with

pytestmark = pytest.mark.gpu

CONV_CASES = [
    # layer, n, D, H, W
    ("conv_0.conv_1", 2, 16, 16, 16),
    ("conv_0.conv_1", 1, 8, 24, 40),
    ("down_1.convs.conv_0", 3, 8, 8, 8),
    ("down_2.convs.conv_0", 2, 8, 8, 4),
    ("down_2.convs.conv_1", 2, 8, 8, 4),
    ("down_3.convs.conv_0", 2, 4, 4, 2),
    ("down_3.convs.conv_1", 5, 4, 6, 4),
    ("down_4.convs.conv_0", 3, 2, 2, 2),
    ("down_4.convs.conv_1", 7, 6, 6, 4),
    ("upcat_4.convs.conv_0", 2, 4, 4, 2),
    ("upcat_3.convs.conv_0", 2, 8, 8, 4),
    ("upcat_2.convs.conv_0", 2, 16, 16, 8),
    ("upcat_1.convs.conv_0", 1, 32, 32, 16),
    ("upcat_1.convs.conv_1", 1, 96, 96, 64),
    ("upcat_1.convs.conv_0", 2, 96, 96, 64),
]


@pytest.mark.parametrize("layer,n,D,H,W", CONV_CASES)
def test_conv3d_matches_torch(layer, n, D, H, W):
    ctx, sd, _ = ctx_with("random")
    w = sd["module." + layer + ".conv.weight"].cuda()
    cin = w.shape[1]
    g = torch.Generator(device="cuda").manual_seed(hash((layer, n, D)) % 1000)
    x = torch.randn(n, cin, D, H, W, device="cuda", generator=g) * 1.5 + 0.3
    y = torch.empty(n, w.shape[0], D, H, W, device="cuda")
    stats = torch.zeros(n, w.shape[0], 2, device="cuda", dtype=torch.float64)
    ctx.op_conv3d(layer, x, y, stats)
    ref = F.conv3d(bf16_round(x), bf16_round(w), None, padding=1)
    # output is stored bf16-rounded (rel 2^-9); accumulation order differs (fp32)
    tol = 2.0 ** -8 * ref.abs() + 2e-3
    err = (y - ref).abs()
    assert bool((err <= tol).all()), f"max err {err.max().item()} at ref {ref.flatten()[err.argmax()].item()}"
    s_ref = ref.double().sum(dim=(2, 3, 4))
    q_ref = (ref.double() ** 2).sum(dim=(2, 3, 4))
    nv = D * H * W
    assert torch.allclose(stats[..., 0], s_ref, rtol=1e-4, atol=1e-3 * nv ** 0.5)
    assert torch.allclose(stats[..., 1], q_ref, rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("name,n,D,H,W", [("upcat_4", 2, 2, 2, 2), ("upcat_4", 3, 6, 6, 4), ("upcat_3", 2, 4, 4, 2),
                                          ("upcat_2", 2, 8, 8, 4), ("upcat_1", 1, 16, 16, 8), ("upcat_1", 1, 48, 48, 32)])
def test_deconv_matches_torch(name, n, D, H, W):
    ctx, sd, _ = ctx_with("random")
    w = sd["module." + name + ".upsample.deconv.weight"].cuda()
    b = sd["module." + name + ".upsample.deconv.bias"].cuda()
    x = torch.randn(n, w.shape[0], D, H, W, device="cuda")
    y = torch.empty(n, w.shape[1], 2 * D, 2 * H, 2 * W, device="cuda")
    ctx.op_deconv(name, x, y)
    ref = F.conv_transpose3d(bf16_round(x), bf16_round(w), b, stride=2)
    tol = 2.0 ** -8 * ref.abs() + 2e-3
    err = (y - ref).abs()
    assert bool((err <= tol).all()), f"max err {err.max().item()}"

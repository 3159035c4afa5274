"""Whole-network parity: dlv_unet_forward vs the torch-fp32 oracle net (oracle/unet_ref.py).

Tolerance (stated per north_star): bf16 operands / fp32 accumulation ->
|logit - ref| <= 1.0 + 0.02*|ref| everywhere and >= 99.9 % agreement of (logit >= 0).
"""
import numpy as np
import pytest
import torch

from conftest import weights_path
from gpu_common import ctx_with
from oracle import pipeline_ref as P

pytestmark = pytest.mark.gpu


def _windows(roi, n, seed):
    vol = P.synth_volume((roi[0], roi[1] * n, roi[2]), seed)
    vol = np.where(vol == 0, 300, vol).astype(np.uint16)          # fully bright windows ...
    w = np.stack([vol[:, i * roi[1]:(i + 1) * roi[1], :] for i in range(n)])
    w[0, : roi[0] // 2] = 0                                       # ... except a half-empty one
    return np.ascontiguousarray(w)


def _check(ctx, net, roi, n, seed, min_agree=0.999):
    w = _windows(roi, n, seed)
    wd = torch.from_numpy(w).cuda()                                # uint16 on the device
    out = torch.empty((n,) + tuple(roi), device="cuda", dtype=torch.float32)
    ctx.unet_forward(wd, roi, out)
    with torch.no_grad():
        ref = torch.cat([net(torch.from_numpy(w[i:i + 4].astype(np.float32)).cuda()[:, None])[:, 0] for i in range(0, n, 4)])
    err = (out - ref).abs()
    tol = 1.0 + 0.02 * ref.abs()
    agree = ((out >= 0) == (ref >= 0)).float().mean().item()
    print(f"roi {roi} n {n}: max|err| {err.max().item():.4f}  p99.9 {err.flatten().kthvalue(int(err.numel() * 0.999)).values.item():.4f} "
          f"median {err.median().item():.5f}  agreement {agree:.6f}  ref range [{ref.min().item():.1f},{ref.max().item():.1f}]")
    assert bool((err <= tol).all()), f"max err {err.max().item()}"
    # sign flips may only happen where the reference logit itself is within the error band of zero
    flips = (out >= 0) != (ref >= 0)
    assert not bool(flips.any()) or ref[flips].abs().max().item() <= 0.25
    assert agree >= min_agree


@pytest.mark.parametrize("roi,n", [((32, 32, 32), 3), ((16, 48, 32), 5), ((96, 96, 64), 2), ((64, 64, 32), 33)])
def test_unet_random_weights(roi, n):
    ctx, _, net = ctx_with("random")
    # random weights put most logits within a few units of 0, so the binarised-agreement bar of the
    # shipped network (99.9 %) does not apply; every flip must still sit inside the error band
    _check(ctx, net, roi, n, seed=5, min_agree=0.99)


def test_unet_shipped_weights():
    wp = weights_path()
    if wp is None:
        pytest.skip("shipped checkpoint not staged")
    ctx, _, net = ctx_with(wp)
    _check(ctx, net, (96, 96, 64), 3, seed=7)
    _check(ctx, net, (64, 64, 32), 4, seed=8)

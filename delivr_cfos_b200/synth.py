"""Seeded synthetic cFos-like uint16 volumes generated on the GPU (bench / smoke inputs).

Ellipsoid "brain" with semi-axes 0.45*dims (0 outside, >= 1 inside - exercises the skip rule and the mask
erosion), lognormal(7.4, 0.35) background and ~370 Gaussian blobs per 10^6 voxels (sigma 2 voxels, amplitude
U(2000, 30000)).  Built slab-by-slab with torch ops so that a 256x2048x2048 volume takes seconds, not minutes.
"""
import math

import torch


def _blur_axis(x, sigma, axis):
    r = int(3 * sigma + 0.5)
    k = torch.exp(-0.5 * (torch.arange(-r, r + 1, device=x.device, dtype=torch.float32) / sigma) ** 2)
    x = x.movedim(axis, -1)
    shp = x.shape
    y = torch.nn.functional.conv1d(x.reshape(-1, 1, shp[-1]), k.view(1, 1, -1), padding=r)
    return y.reshape(shp).movedim(-1, axis)


def synth_volume_cuda(shape, seed, roi=None, device="cuda", slab=64, blobs_per_mvox=370.0):
    """-> uint16 CUDA tensor (Z,Y,X), or zero-padded to multiples of ``roi`` at the high end
    (like masked_nifti.npy, downsample_and_mask.py:391-396)."""
    Z, Y, X = (int(s) for s in shape)
    if roi is not None:
        PZ, PY, PX = (int(math.ceil(d / r) * r) for d, r in zip(shape, roi))
    else:
        PZ, PY, PX = Z, Y, X
    out = torch.zeros((PZ, PY, PX), dtype=torch.uint16, device=device)
    g = torch.Generator(device=device).manual_seed(int(seed))
    yy = ((torch.arange(Y, device=device, dtype=torch.float32) - (Y - 1) / 2) / (0.45 * Y)) ** 2
    xx = ((torch.arange(X, device=device, dtype=torch.float32) - (X - 1) / 2) / (0.45 * X)) ** 2
    sigma, halo = 2.0, 6
    for z0 in range(0, Z, slab):
        z1 = min(Z, z0 + slab)
        a0, a1 = max(0, z0 - halo), min(Z, z1 + halo)
        n = a1 - a0
        imp = torch.zeros((n, Y, X), dtype=torch.float32, device=device)
        # blob centres of this slab (+halo) are drawn from a generator keyed by plane so that slabs agree
        for z in range(a0, a1):
            gz = torch.Generator(device=device).manual_seed(int(seed) * 1000003 + z)
            nb = max(1, int(blobs_per_mvox * Y * X / 1e6))
            cy = torch.randint(0, Y, (nb,), generator=gz, device=device)
            cx = torch.randint(0, X, (nb,), generator=gz, device=device)
            amp = torch.rand((nb,), generator=gz, device=device) * 28000 + 2000
            imp[z - a0].index_put_((cy, cx), amp, accumulate=True)
        for ax in range(3):
            imp = _blur_axis(imp, sigma, ax)
        imp = imp[z0 - a0: z0 - a0 + (z1 - z0)]
        bg = torch.exp(torch.randn((z1 - z0, Y, X), generator=g, device=device) * 0.35 + 7.4)
        vol = (bg + imp).clamp_(1, 65535)
        zz = ((torch.arange(z0, z1, device=device, dtype=torch.float32) - (Z - 1) / 2) / (0.45 * Z)) ** 2
        inside = (zz[:, None, None] + yy[None, :, None] + xx[None, None, :]) <= 1.0
        vol = torch.where(inside, vol, torch.zeros_like(vol))
        out[z0:z1, :Y, :X] = vol.to(torch.int32).to(torch.uint16)
    return out

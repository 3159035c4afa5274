"""Seeded synthetic cFos-like uint16 volumes generated on the GPU (bench / smoke inputs).

Ellipsoid "brain" with semi-axes 0.45*dims (0 outside, >= 1 inside - exercises the skip rule and the mask
erosion), lognormal(7.4, 0.35) background and ~370 Gaussian blobs per 10^6 voxels (sigma 2 voxels, amplitude
U(2000, 30000)).  Every plane draws from its own generator keyed by (seed, z), so any z-range of the volume can
be produced independently (z-slab sharded runs generate only their own planes) and is identical to the same
planes of the full volume.
"""
import math

import torch

_SIGMA, _HALO = 2.0, 6


def _blur_axis(x, sigma, axis):
    r = int(3 * sigma + 0.5)
    k = torch.exp(-0.5 * (torch.arange(-r, r + 1, device=x.device, dtype=torch.float32) / sigma) ** 2)
    x = x.movedim(axis, -1)
    shp = x.shape
    y = torch.nn.functional.conv1d(x.reshape(-1, 1, shp[-1]), k.view(1, 1, -1), padding=r)
    return y.reshape(shp).movedim(-1, axis)


def _plane(seed, z, Y, X, device, blobs_per_mvox):
    """(impulses, background) of plane z - a pure function of (seed, z)."""
    g = torch.Generator(device=device).manual_seed(int(seed) * 1000003 + int(z))
    nb = max(1, int(blobs_per_mvox * Y * X / 1e6))
    cy = torch.randint(0, Y, (nb,), generator=g, device=device)
    cx = torch.randint(0, X, (nb,), generator=g, device=device)
    amp = torch.rand((nb,), generator=g, device=device) * 28000 + 2000
    imp = torch.zeros((Y, X), dtype=torch.float32, device=device)
    imp.index_put_((cy, cx), amp, accumulate=True)
    bg = torch.exp(torch.randn((Y, X), generator=g, device=device) * 0.35 + 7.4)
    return imp, bg


def synth_volume_cuda(shape, seed, roi=None, device="cuda", slab=32, blobs_per_mvox=370.0, z_range=None):
    """-> uint16 CUDA tensor.  Whole volume (Z,Y,X) - zero-padded at the high end to multiples of ``roi`` if given,
    like masked_nifti.npy (downsample_and_mask.py:391-396) - or only planes ``z_range=(z0,z1)`` (unpadded in-plane)."""
    Z, Y, X = (int(s) for s in shape)
    zr0, zr1 = (0, Z) if z_range is None else (int(z_range[0]), int(z_range[1]))
    if roi is not None and z_range is None:
        PZ, PY, PX = (int(math.ceil(d / r) * r) for d, r in zip(shape, roi))
    else:
        PZ, PY, PX = zr1 - zr0, Y, X
    out = torch.zeros((PZ, PY, PX), dtype=torch.uint16, device=device)
    yy = ((torch.arange(Y, device=device, dtype=torch.float32) - (Y - 1) / 2) / (0.45 * Y)) ** 2
    xx = ((torch.arange(X, device=device, dtype=torch.float32) - (X - 1) / 2) / (0.45 * X)) ** 2
    for z0 in range(zr0, zr1, slab):
        z1 = min(zr1, z0 + slab)
        a0, a1 = max(0, z0 - _HALO), min(Z, z1 + _HALO)
        planes = [_plane(seed, z, Y, X, device, blobs_per_mvox) for z in range(a0, a1)]
        imp = torch.stack([p[0] for p in planes])
        for ax in range(3):
            imp = _blur_axis(imp, _SIGMA, ax)
        imp = imp[z0 - a0: z0 - a0 + (z1 - z0)]
        bg = torch.stack([p[1] for p in planes[z0 - a0: z0 - a0 + (z1 - z0)]])
        vol = (bg + imp).clamp_(1, 65535)
        zz = ((torch.arange(z0, z1, device=device, dtype=torch.float32) - (Z - 1) / 2) / (0.45 * Z)) ** 2
        inside = (zz[:, None, None] + yy[None, :, None] + xx[None, None, :]) <= 1.0
        vol = torch.where(inside, vol, torch.zeros_like(vol))
        out[z0 - zr0: z1 - zr0, :Y, :X] = vol.to(torch.int32).to(torch.uint16)
    return out


def synth_mask_cuda(shape, seed, device="cuda", kind="blobs", p=0.08, blobs_per_mvox=370.0, chunk=50):
    """Seeded uint8 binary mask for config 3, generated on the GPU in z-chunks.

    ``blobs``: ball-shaped blobs (offsets with dz^2+dy^2+dx^2 <= 5) around ~370 uniformly drawn centres per 10^6
    voxels (foreground ~2 %); ``bernoulli``: independent voxels with probability ``p`` (stress case).
    Every chunk draws from its own generator keyed by (seed, chunk), so the mask is a pure function of the arguments."""
    Z, Y, X = (int(s) for s in shape)
    out = torch.zeros((Z, Y, X), dtype=torch.uint8, device=device)
    offs = [(a, b, c) for a in range(-2, 3) for b in range(-2, 3) for c in range(-2, 3) if a * a + b * b + c * c <= 5]
    for ci, z0 in enumerate(range(0, Z, chunk)):
        z1 = min(Z, z0 + chunk)
        g = torch.Generator(device=device).manual_seed(int(seed) * 7919 + ci)
        if kind == "bernoulli":
            out[z0:z1] = (torch.rand((z1 - z0, Y, X), generator=g, device=device) < p).to(torch.uint8)
            continue
        n = max(1, int(blobs_per_mvox * (z1 - z0) * Y * X / 1e6))
        cz = torch.randint(z0, z1, (n,), generator=g, device=device)
        cy = torch.randint(0, Y, (n,), generator=g, device=device)
        cx = torch.randint(0, X, (n,), generator=g, device=device)
        for a, b, c in offs:
            z = (cz + a).clamp_(0, Z - 1)
            y = (cy + b).clamp_(0, Y - 1)
            x = (cx + c).clamp_(0, X - 1)
            out[z, y, x] = 1
    return out

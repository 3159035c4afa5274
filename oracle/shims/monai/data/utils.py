"""monai.data.utils names used at inference/sliding_window_inferer.py:19,143,148.

Restated from MONAI 1.2.0's published behaviour.
"""
import itertools
import math

import torch


def get_valid_patch_size(image_size, patch_size):
    nd = len(image_size)
    if isinstance(patch_size, int):
        patch_size = (patch_size,) * nd
    return tuple(min(ms, ps or ms) for ms, ps in zip(image_size, patch_size))


def dense_patch_slices(image_size, patch_size, scan_interval):
    num_spatial_dims = len(image_size)
    patch_size = get_valid_patch_size(image_size, patch_size)
    scan_num = []
    for i in range(num_spatial_dims):
        if scan_interval[i] == 0:
            scan_num.append(1)
        else:
            num = int(math.ceil(float(image_size[i]) / scan_interval[i]))
            scan_dim = next((d for d in range(num) if d * scan_interval[i] + patch_size[i] >= image_size[i]), None)
            scan_num.append(scan_dim + 1 if scan_dim is not None else 1)
    starts = []
    for dim in range(num_spatial_dims):
        dim_starts = []
        for idx in range(scan_num[dim]):
            start_idx = idx * scan_interval[dim]
            start_idx -= max(start_idx + patch_size[dim] - image_size[dim], 0)
            dim_starts.append(start_idx)
        starts.append(dim_starts)
    out = []
    for s in itertools.product(*starts):  # first dim slowest == meshgrid(indexing="ij")
        out.append(tuple(slice(a, a + p) for a, p in zip(s, patch_size)))
    return out


def compute_importance_map(patch_size, mode="constant", sigma_scale=0.125, device="cpu"):
    mode = str(getattr(mode, "value", mode)).lower()
    if mode == "constant":
        return torch.ones(tuple(patch_size), device=device, dtype=torch.float)
    if mode == "gaussian":
        if not isinstance(sigma_scale, (tuple, list)):
            sigma_scale = (sigma_scale,) * len(patch_size)
        w = None
        for n, s in zip(patch_size, sigma_scale):
            x = torch.arange(-(n - 1) / 2.0, (n - 1) / 2.0 + 1, dtype=torch.float, device=device)
            g = torch.exp(x ** 2 / (-2 * (s * n) ** 2))
            w = g if w is None else w.unsqueeze(-1) * g
        w = w / w.max()
        mn = w[w != 0].min()
        return torch.clamp(w, min=float(mn))
    raise ValueError(mode)

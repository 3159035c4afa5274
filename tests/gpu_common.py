"""Shared helpers for the -m gpu parity tests (everything goes through the C ABI)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import unet_ref  # noqa: E402  (tests may use the oracle)

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

_CTX = {}


def ctx_with(weights="random"):
    """Context with either seeded random weights or the shipped checkpoint (if staged)."""
    from delivr_cfos_b200 import Context
    if weights in _CTX:
        return _CTX[weights]
    c = Context(0)
    if weights == "random":
        sd = unet_ref.random_state_dict(0)
    else:
        sd = torch.load(weights, map_location="cpu", weights_only=True)["state_dict"]
    c.load_weights(sd)
    net = unet_ref.BasicUNet(dropout=0.1)
    net.load_state_dict(unet_ref.strip_module_prefix(sd), strict=True)
    net = net.eval().cuda()
    _CTX[weights] = (c, sd, net)
    return _CTX[weights]


def bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)

"""Profiling target: one U-Net batch (dlv_unet_forward) on N windows of 96x96x64.  Usage: python tools/run_unet_once.py [nwin]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import state_dict
from delivr_cfos_b200 import Context
from delivr_cfos_b200.synth import synth_volume_cuda

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
roi = (96, 96, 64)
ctx = Context(0)
ctx.load_weights(state_dict()[0])
vol = synth_volume_cuda((96 * n, 96, 64), 5)
v32 = vol.to(torch.int32)
vol = torch.where(v32 == 0, torch.full_like(v32, 300), v32).to(torch.uint16).contiguous().view(n, 96, 96, 64)
out = torch.empty((n,) + roi, dtype=torch.float32, device="cuda")
ctx.unet_forward(vol, roi, out)
torch.cuda.synchronize()
print("ok", float(out.mean()))

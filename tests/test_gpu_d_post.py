"""dlv_op_finalise (sigmoid/threshold + block-aware erosion) vs the oracle's create_binaries - bit-exact."""
import numpy as np
import pytest
import torch

from oracle import pipeline_ref as P

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,roi,block", [((40, 70, 90), (16, 16, 16), 0), ((70, 64, 96), (32, 32, 32), 17),
                                             ((33, 50, 130), (16, 16, 16), 62), ((20, 40, 33), (16, 16, 16), 5)])
def test_finalise_bit_exact(shape, roi, block, monkeypatch):
    from delivr_cfos_b200 import Context
    ctx = Context(0)
    rng = np.random.default_rng(3)
    vol = P.synth_volume(shape, 21, roi=roi)
    Z, Y, X = shape
    zz, yy, xx = np.ogrid[:Z, :Y, :X]
    vol[:Z, :Y, :X][((zz - Z * 0.5) ** 2 + (yy - Y * 0.4) ** 2 + (xx - X * 0.6) ** 2) < 30] = 0
    avg = (rng.normal(0, 3, size=vol.shape)).astype(np.float32)
    avg16 = avg.astype(np.float16)
    if block:
        monkeypatch.setattr(P, "arrayterator_blocks", lambda s, buf=0: [(z0, min(s[0], z0 + block)) for z0 in range(0, s[0], block)])
    ref, ref_sig = P.create_binaries(avg16, vol, shape, 0.5, return_sigmoid=True)
    out = torch.empty(shape, dtype=torch.uint8, device="cuda")
    sig = torch.empty(shape, dtype=torch.float32, device="cuda")
    ctx.op_finalise(torch.from_numpy(avg).cuda(), torch.from_numpy(vol).cuda(),
                    vol.shape, shape, out, 0.5, 30, block, sig)
    assert np.array_equal(out.cpu().numpy(), ref)
    assert np.abs(sig.cpu().numpy() - ref_sig).max() < 1e-6

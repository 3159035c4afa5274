"""Oracle-backed stand-ins for the CUDA engine / slab worker (tests only): the host logic of the drop-ins - planning,
out-of-core chunking, rank processes over gloo, file layout - runs here without a GPU.  The 'network' is a
deterministic function of the voxel value; blending is the library's integer sum; erosion / labelling come from the
C oracle."""
import os

import numpy as np
import torch

from delivr_cfos_b200 import slabs
from oracle import ccl_ref

ACC_SCALE = 4096.0           # the library's fixed-point logit unit (2^-12)


def fake_logits(w):
    """Fixed-point logits of a window (int64 voxel values): ~11 % foreground."""
    return np.where(w % 9 == 0, 4096, -4096).astype(np.int32)


def boundary_pairs_ref(lo_plane, hi_plane):
    out = set()
    Y, X = hi_plane.shape
    for y, x in zip(*np.nonzero(hi_plane)):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                yy, xx = y + dy, x + dx
                if 0 <= yy < Y and 0 <= xx < X and lo_plane[yy, xx]:
                    out.add((int(lo_plane[yy, xx]), int(hi_plane[y, x])))
    return np.array(sorted(out), dtype=np.uint32).reshape(-1, 2)


class OracleLabelOps:
    def ccl(self):
        b = self.binaries.numpy() if hasattr(self.binaries, "numpy") else np.asarray(self.binaries)
        if b.shape[0] == 0:
            self.labels, self.table = torch.zeros((0,) + tuple(b.shape[1:]), dtype=torch.int32), None
            return 0
        lab, n = ccl_ref.connected_components26(np.ascontiguousarray(b))
        self.table = {"n": n, **ccl_ref.statistics(lab, n)}
        self.labels = torch.from_numpy(lab.astype(np.int32))
        return n

    def last_plane(self):
        return self.labels[-1].contiguous()

    def empty_plane(self):
        return torch.empty(tuple(self.binaries.shape[1:]), dtype=torch.int32)

    def boundary_pairs(self, lo):
        return boundary_pairs_ref(lo.numpy(), self.labels[0].numpy())

    def relabel(self, lut):
        self.labels = torch.from_numpy(lut.astype(np.int64)[self.labels.numpy()].astype(np.int32))


class OracleLabelSlab(OracleLabelOps):
    def __init__(self, binaries):
        self.binaries = torch.from_numpy(np.ascontiguousarray(binaries))
        self.labels = self.table = None


class OracleWorker(OracleLabelOps):
    """CPU stand-in for CudaSlabWorker (same interface)."""

    def __init__(self, plan, rank, planes_fn, threshold=0.5, tta=False, erosion_block_planes=0, blend_mode=0,
                 want_sigmoid=False, keep_avg=False, window_batch=0):
        self.plan, self.r = plan, rank
        self.info = plan.rank(rank)
        z0, z1 = self.info["slab"]
        self.slab = np.asarray(planes_fn(z0, z1)) if z1 > z0 else None
        self.want_sigmoid, self.keep_avg, self.tta = want_sigmoid, keep_avg, tta
        self.sigmoid = self.avg_own = None

    def accumulate(self):
        if self.slab is None:
            return np.zeros(0, dtype=np.int32)
        z0 = self.info["slab"][0]
        rz, ry, rx = self.plan.roi
        self.acc = torch.zeros(self.slab.shape, dtype=torch.int32)
        act = []
        rep = 13 if self.tta else 1
        for (z, y, x) in self.plan.windows_of(self.r):
            w = self.slab[z - z0:z - z0 + rz, y:y + ry, x:x + rx].astype(np.int64)
            a = int(w.max() > 0)
            act.append(a)
            if a:
                self.acc[z - z0:z - z0 + rz, y:y + ry, x:x + rx] += torch.from_numpy(fake_logits(w) * rep)
        return np.array(act, dtype=np.int32)

    def acc_planes(self, g0, g1):
        z0 = self.info["slab"][0]
        return self.acc[g0 - z0:g1 - z0]

    def add_planes(self, g0, g1, t):
        z0 = self.info["slab"][0]
        self.acc[g0 - z0:g1 - z0] += t

    def _average(self, active_global):
        """(sum of active logits + (-1000) per skipped covering window) / covering windows, like average_kernel."""
        plan = self.plan
        z0, z1 = self.info["slab"]
        PZ, PY, PX = plan.shape_pad
        rz, ry, rx = plan.roi
        act = np.asarray(active_global).reshape(len(plan.sz), len(plan.sy), len(plan.sx))
        cnt = np.zeros((z1 - z0, PY, PX), dtype=np.float32)
        skipped = np.zeros_like(cnt)
        for iz, sz in enumerate(plan.sz):
            a, b = max(sz, z0), min(sz + rz, z1)
            if b <= a:
                continue
            for iy, sy in enumerate(plan.sy):
                for ix, sx in enumerate(plan.sx):
                    cnt[a - z0:b - z0, sy:sy + ry, sx:sx + rx] += 1
                    if not act[iz, iy, ix]:
                        skipped[a - z0:b - z0, sy:sy + ry, sx:sx + rx] += 1
        passes = 13 if self.tta else 1
        s = self.acc.numpy().astype(np.float32) * np.float32(1.0 / ACC_SCALE) + np.float32(-1000.0) * skipped * passes
        return s / (cnt * passes)

    def finalise(self, active_global):
        o0, o1 = self.info["own_real"]
        Z, Y, X = self.plan.shape_real
        if self.slab is None or o1 <= o0:
            self.binaries = torch.zeros((0, Y, X), dtype=torch.uint8)
            return self.binaries
        z0, z1 = self.info["slab"]
        avg = self._average(active_global)
        # the slab holds the erosion halo of the planes it owns (planes beyond the volume are zero padding)
        vol = self.slab[:max(0, min(z1, Z) - z0), :Y, :X]
        mask = ccl_ref.erode6((vol > 0).astype(np.uint8), self.plan.iters)
        a = avg[o0 - z0:o1 - z0, :Y, :X].astype(np.float16).astype(np.float32)     # fp16 rounding kept (inference.py:242)
        with np.errstate(over="ignore"):
            sig = 1.0 / (1.0 + np.exp(-a))
        self.binaries = torch.from_numpy(((sig >= 0.5) & (mask[o0 - z0:o1 - z0] > 0)).astype(np.uint8))
        if self.want_sigmoid:
            self.sigmoid = torch.from_numpy(sig.astype(np.float32))
        if self.keep_avg:
            p0, p1 = self.info["own"]
            self.avg_own = torch.from_numpy(avg[p0 - z0:p1 - z0])
        self.acc = None
        return self.binaries


class OracleEngine:
    """Stand-in for inference.CudaEngine (DLV_ENGINE=tests.cpu_engine:OracleEngine in rank processes)."""

    backend = "gloo"

    def __init__(self, model_weights=None, device=0):
        self.device = device

    def free_bytes(self):
        return int(os.environ.get("DLV_TEST_FREE_BYTES", 1 << 40))

    def comm(self):
        return slabs.TorchComm(None)

    def planes_fn(self, source):
        return lambda z0, z1: np.array(source[z0:z1])

    def windows_active(self, slab, local_windows, roi):
        rz, ry, rx = roi
        return np.array([int(slab[z:z + rz, y:y + ry, x:x + rx].max() > 0) for z, y, x in local_windows], dtype=np.int32)

    def make_worker(self, plan, r, planes_fn, **kw):
        return OracleWorker(plan, r, planes_fn, **kw)

    def to_host(self, t):
        return t.numpy() if hasattr(t, "numpy") else np.asarray(t)

    def segment_incore(self, volume, shape_pad, shape_real, roi, binarized, overlap=0.5, erosion_iters=30,
                       avg_logits_out=None, sigmoid_out=None, **kw):
        plan = slabs.SlabPlan(shape_real, roi, overlap, 1, erosion_iters=erosion_iters)
        w = OracleWorker(plan, 0, lambda z0, z1: np.asarray(volume)[z0:z1], want_sigmoid=sigmoid_out is not None,
                         keep_avg=avg_logits_out is not None, **kw)
        act = w.accumulate()
        w.finalise(act)
        binarized[...] = w.binaries.numpy()
        if sigmoid_out is not None:
            sigmoid_out[...] = w.sigmoid.numpy()
        if avg_logits_out is not None:
            avg_logits_out[...] = w.avg_own.numpy()
        return {"windows_active": int(act.sum()), "windows_total": len(act), "passes": 13 if kw.get("tta") else 1,
                "ms_unet": 0.0, "ms_finalise": 0.0}

"""Runs the reference's OWN hot-path files, unmodified, on the CPU (TEST / BASELINE INFRASTRUCTURE - never part of
the product path).  Used by ``bench.py --impl reference`` and the ``cpu_baseline`` leg so that the reported CPU number
comes from the reference's code (``inference/inference.py::run_inference`` + ``count_blobs.py::count_blobs``), not
from a port.

The files are imported from ``$DLV_REFERENCE``, ``/root/reference`` (authoring container) or the copy that
``__graft_entry__.build()`` stages under the git-ignored ``baseline/_ref/reference/`` (what travels to the GPU box),
with ``oracle/shims`` standing in for the third-party modules that are not installed (monai, cc3d, path, nibabel,
skimage - SURVEY.md appendix A).  CPU-only environment patches, the same as ``oracle/make_golden.py``:
``Tensor.cuda`` -> identity, ``torch.cuda.device_count`` -> 1 and a ``mem_get_info`` chosen so that
inference.py:172-186 derives the requested sliding-window batch size.
"""
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
FILES = ("__main__.py", "count_blobs.py", "filehandling.py", "inference/inference.py", "inference/sliding_window_inferer.py",
         "automate_mBrainaligner.py")


def reference_root():
    for p in (os.environ.get("DLV_REFERENCE"), "/root/reference", os.path.join(ROOT, "baseline", "_ref", "reference")):
        if p and os.path.exists(os.path.join(p, "inference", "inference.py")):
            return p
    return None


def stage(dst=None):
    """Copy the reference files this runner (and the __main__ integration test) executes into baseline/_ref/reference/
    - git-ignored, so the sources never enter the repository's history, but shipped to the GPU box with the snapshot."""
    import shutil
    src = "/root/reference"
    dst = dst or os.path.join(ROOT, "baseline", "_ref", "reference")
    if not os.path.exists(os.path.join(src, "inference", "inference.py")):
        return False
    for f in FILES:
        os.makedirs(os.path.dirname(os.path.join(dst, f)), exist_ok=True)
        shutil.copyfile(os.path.join(src, f), os.path.join(dst, f))
    return True


_MODS = None


def load():
    """-> (reference inference module, reference count_blobs module, Path class), CPU patches applied."""
    global _MODS
    if _MODS is not None:
        return _MODS
    ref = reference_root()
    if ref is None:
        raise FileNotFoundError("reference files neither at /root/reference nor staged under baseline/_ref/reference")
    for p in (ref, os.path.join(HERE, "shims")):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    for name in ("inference", "inference.inference", "inference.sliding_window_inferer", "count_blobs"):
        sys.modules.pop(name, None)
    import torch
    import inference.inference as ref_inf
    import count_blobs as ref_cb
    assert os.path.abspath(ref_inf.__file__).startswith(os.path.abspath(ref)), ref_inf.__file__
    assert os.path.abspath(ref_cb.__file__).startswith(os.path.abspath(ref)), ref_cb.__file__
    from path import Path
    torch.Tensor.cuda = lambda self, *a, **k: self             # sliding_window_inferer.py:208
    torch.cuda.empty_cache = lambda: None
    torch.cuda.device_count = lambda: 1
    _MODS = (ref_inf, ref_cb, Path)
    return _MODS


class _cpu_only:
    """The reference picks `cuda` whenever torch sees a GPU (inference.py:156,209) and wraps the model in DataParallel
    (:217-219), which would move every window batch to the GPU: on a GPU box the "CPU baseline" silently ran its U-Net
    through cuDNN.  While the reference runs, torch.cuda.is_available() answers False - device = cpu, DataParallel
    degenerates to a plain call of the module - so the whole path really runs on the host cores."""

    def __enter__(self):
        import torch
        self._orig = torch.cuda.is_available
        torch.cuda.is_available = lambda: False

    def __exit__(self, *exc):
        import torch
        torch.cuda.is_available = self._orig
        return False


def run(volume_pad, shape_real, roi, weights, tta=False, sw_batch=4, load_all_ram=True, quiet=True):
    """One full pass of the reference's two stages over one padded uint16 volume.
    -> dict(t_inference, t_count (seconds), binaries, csv, n)."""
    import contextlib
    import io
    import torch
    ref_inf, ref_cb, Path = load()
    per_win_mb = roi[0] * roi[1] * roi[2] * 32 * 45 / (1024 ** 2)
    fb = int(sw_batch * per_win_mb * (1024 ** 2) / 0.95) + 1024
    torch.cuda.mem_get_info = lambda i=0: (fb, fb)
    cwd = os.getcwd()
    sink = io.StringIO()
    with tempfile.TemporaryDirectory() as tmp, (contextlib.redirect_stdout(sink) if quiet else contextlib.nullcontext()), \
            (contextlib.redirect_stderr(sink) if quiet else contextlib.nullcontext()):
        brain = "brainA"
        nif_dir = os.path.join(tmp, "in", brain, "masked_niftis")
        os.makedirs(nif_dir)
        mm = np.lib.format.open_memmap(os.path.join(nif_dir, "masked_nifti.npy"), mode="w+", dtype=np.uint16, shape=(1, 1) + tuple(volume_pad.shape))
        mm[0, 0] = volume_pad
        mm.flush()
        del mm
        out_dir, post_dir = os.path.join(tmp, "out02"), os.path.join(tmp, "out03") + "/"
        os.makedirs(out_dir)
        settings = {"blob_detection": {"window_dimensions": {"window_dim_0": roi[0], "window_dim_1": roi[1], "window_dim_2": roi[2]}},
                    "postprocessing": {"output_location": post_dir}, "FLAGS": {"SAVE_ACTIVATED_OUTPUT": False, "LOAD_ALL_RAM": load_all_ram}}
        t0 = time.perf_counter()
        try:
            with _cpu_only():
                session = ref_inf.run_inference(niftis=[Path(os.path.join(nif_dir, "masked_nifti.npy"))], output_folder=out_dir,
                                                stack_shape=(1, 1, *shape_real), model_weights=weights, tta=tta, comment=brain,
                                                load_all_ram=load_all_ram, settings=settings)
        finally:
            os.chdir(cwd)
        t1 = time.perf_counter()
        ref_cb.count_blobs(settings, out_dir, 0, brain, (1, 1, *shape_real))
        t2 = time.perf_counter()
        binaries = np.load(os.path.join(session, "binary_segmentations", "binaries.npy"))
        csv_file = [f for f in os.listdir(post_dir) if f.endswith(".csv")][0]
        csv = open(os.path.join(post_dir, csv_file)).read()
        n = int([f for f in os.listdir(post_dir) if f.endswith("-cc3d.npy")][0].split("-")[1])
    return {"t_inference": t1 - t0, "t_count": t2 - t1, "binaries": binaries, "csv": csv, "n": n}

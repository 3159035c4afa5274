"""CPU suite: the C-ABI library loads and exports exactly what include/delivr_b200.h declares; host logic."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT


def _header_functions():
    txt = open(os.path.join(ROOT, "include", "delivr_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(dlv_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    import delivr_cfos_b200
    from delivr_cfos_b200._lib import EXPORTS, LIB_PATH
    lib = delivr_cfos_b200.load_library()
    declared = _header_functions()
    assert declared == sorted(EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None
    out = subprocess.run(["nm", "-D", "--defined-only", LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (dlv_[a-z0-9_]+)", out)))
    assert exported == declared
    assert lib.dlv_abi_version() == 4


def test_no_cpu_fallback_without_gpu():
    """On a box without a CUDA device the product fails loudly instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from delivr_cfos_b200 import Context, DlvError
    with pytest.raises(DlvError):
        Context(0)
    from delivr_cfos_b200.inference.sliding_window_inferer import DelivrNet
    with pytest.raises(DlvError):
        DelivrNet(state_dict={})


def test_default_engine_is_the_cuda_library(monkeypatch):
    """The engine seam of the drop-in (used by the CPU tests of the host logic) is never set by the package itself:
    without DLV_ENGINE the factory is the CUDA engine, and building it on a box without a GPU raises."""
    import torch
    from delivr_cfos_b200.inference import inference as inf
    monkeypatch.delenv("DLV_ENGINE", raising=False)
    assert inf._engine_factory() is inf.CudaEngine
    if not torch.cuda.is_available():
        from delivr_cfos_b200 import DlvError
        with pytest.raises(DlvError):
            inf.CudaEngine("no-such-checkpoint.tar", 0)
    src = "".join(open(os.path.join(dp, f)).read() for dp, _, fs in os.walk(os.path.join(ROOT, "delivr_cfos_b200")) for f in fs if f.endswith(".py"))
    assert src.count('"DLV_ENGINE"') == 1 and "_ENGINE_FACTORY =" in src and src.count("_ENGINE_FACTORY = ") == 1


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under delivr_cfos_b200/ may reference it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "delivr_cfos_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dirpath, f)
                assert "scipy" not in src, os.path.join(dirpath, f)


def test_struct_layouts_match_header():
    from delivr_cfos_b200._lib import SegParams, SegStats, Table
    assert ctypes.sizeof(SegParams) == 3 * 8 + 3 * 8 + 3 * 4 + 4 + 4 + 4 + 4 + 4 + 8 + 4 + 4 + 4 + 4   # incl. alignment padding
    assert ctypes.sizeof(SegStats) == 4 * 8 + 3 * 8
    assert ctypes.sizeof(Table) == 8 + 4 * 8


def test_update_idx_and_memmap_helpers(tmp_path):
    from delivr_cfos_b200.inference.inference import create_empty_memmap, update_idx
    old, new = update_idx([0, 0, 0], [10, 20, 30], [25, 20, 30])
    assert (old, new) == ([0, 0, 0], [10, 20, 30])
    old, new = update_idx([10, 20, 30], [10, 20, 30], [25, 20, 30])
    assert (old, new) == ([10, 0, 0], [20, 20, 30])
    p = str(tmp_path / "x.npy")
    mm = create_empty_memmap(p, (1, 1, 4, 5, 6), dtype=np.float16, return_torch=False)
    assert mm.offset == 128 and mm.shape == (1, 1, 4, 5, 6) and not mm.any()
    assert np.memmap(p, dtype=np.float16, mode="r", shape=(1, 1, 4, 5, 6), offset=128).sum() == 0


def test_cover_count_matches_oracle_pass():
    from delivr_cfos_b200.inference.sliding_window_inferer import cover_count
    from oracle import pipeline_ref as P
    import torch
    shape, roi = (64, 96, 64), (32, 32, 32)
    out = np.zeros(shape, np.float16)
    cnt = np.zeros(shape, np.uint8)
    vol = np.ones(shape, np.uint16)
    P.sliding_window_pass(vol, roi, 0.5, lambda t: torch.zeros_like(t), 4, out, cnt)
    assert np.array_equal(cover_count(shape, roi, 0.5), cnt)

"""Row f2: the SWC writer / size re-attach mirrors against files the UNMODIFIED reference
automate_mBrainaligner.py wrote (oracle/make_golden_swc.py -> tests/golden/s1_swc.json), byte for byte."""
import json
import os

import numpy as np
import pytest

from delivr_cfos_b200 import automate_mBrainaligner as A
from delivr_cfos_b200.count_blobs import csv_text

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "s1_swc.json")))


@pytest.mark.parametrize("case", sorted(GOLD))
def test_rewrite_swc_matches_reference_files(case, tmp_path, monkeypatch):
    g = GOLD[case]
    xyz, par = "_xyz1" in case, "_par1" in case
    monkeypatch.setattr(os, "cpu_count", lambda: 8)              # the generator's setting (chunks = cpu_count - 1)
    csv_path = os.path.join(str(tmp_path), g["csv_name"])
    with open(csv_path, "w") as f:
        f.write(g["csv"])
    od = os.path.join(str(tmp_path), "o")
    os.makedirs(od)
    files = A.rewrite_swc(csv_path, od, XYZ=xyz, parallel_processing=par)
    assert [os.path.relpath(p, od) for p in files] == [n for n, _ in g["files"]]
    for p, (_, text) in zip(files, g["files"]):
        assert open(p).read() == text
    assert A.split_parameters(csv_path) == g["split_parameters"]
    if "registered_swc" in g:
        swc = os.path.join(str(tmp_path), "reg.swc")
        with open(swc, "w") as f:
            f.write(g["registered_swc"])
        coll = os.path.join(str(tmp_path), "coll")
        os.makedirs(coll)
        A.reattach_size_and_copy(csv_path, swc, "mouseA", od, coll)
        for d in (od, coll):
            assert open(os.path.join(d, "mouseA_local_registered_with_original_size.csv")).read() == g["reattached"]


def test_swc_from_table_equals_csv_route(tmp_path):
    """Emitting from the statistics table directly gives the files the CSV route gives."""
    rng = np.random.default_rng(5)
    n = 57
    cnt = rng.integers(1, 900, size=n + 1).astype(np.uint64)
    sums = (rng.random((n + 1, 3)) * [1500, 4000, 4000] * cnt[:, None].astype(np.float64)).astype(np.uint64)
    stats = {"voxel_counts": cnt, "centroids": sums.astype(np.float64) / cnt[:, None].astype(np.float64)}
    name = "(1500, 4000, 4000)_brain B.csv"
    a, b = os.path.join(str(tmp_path), "a"), os.path.join(str(tmp_path), "b")
    os.makedirs(a), os.makedirs(b)
    with open(os.path.join(str(tmp_path), name), "w") as f:
        f.write(csv_text(stats, n))
    fa = A.rewrite_swc(os.path.join(str(tmp_path), name), a)
    fb = A.swc_from_table(stats, n, name, b)
    assert [os.path.basename(p) for p in fa] == [os.path.basename(p) for p in fb]
    assert open(fa[0]).read() == open(fb[0]).read() and len(open(fa[0]).read().splitlines()) == n

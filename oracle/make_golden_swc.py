"""Golden vectors for SURVEY.md section 8 row f2 (the consumer of count_blobs' CSV): runs the UNMODIFIED
/root/reference/automate_mBrainaligner.py::rewrite_swc (:75-197) and ::reattach_size_and_copy (:237-253) in this
container on the CSV text the reference's own count_blobs wrote for the g1 fixture (tests/golden/g1_notta.npz) and on a
CSV with awkward floats, and stores every file they write in tests/golden/s1_swc.json.

Environment shims (test infrastructure, nothing is shipped): oracle/shims for tifffile / more_itertools; the reference
pins pandas 1.4.3, whose ``Series.str.replace`` treats the pattern as a regular expression by default - pandas >= 2
flipped that default, so the generator restores ``regex=True`` (the file calls it with re.escape'd patterns);
os.cpu_count is fixed at 8 so that the "parallel" split (cpu_count - 1 chunks, :150) is reproducible.

    python oracle/make_golden_swc.py
"""
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "shims"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import pandas as pd  # noqa: E402


def _pandas_1_4_replace():
    orig = pd.core.strings.accessor.StringMethods.replace

    def replace(self, pat, repl, n=-1, case=None, flags=0, regex=True):
        return orig(self, pat, repl, n=n, case=case, flags=flags, regex=regex)
    pd.core.strings.accessor.StringMethods.replace = replace


def awkward_csv():
    """count_blobs-style CSV text with floats that exercise rounding / repr: exponents, halves, many digits."""
    rng = np.random.default_rng(77)
    rows = [",Blob,Coords,Size\n"]
    vals = [[0.0005, 1e-05, 123456.7895], [2.5, 0.125, 1234.0005], [1 / 3, 2 / 3, 1e3], [99.9995, 7.0, 0.30000000000000004]]
    vals += rng.random((40, 3)).__mul__([1500.0, 4000.0, 4000.0]).tolist()
    for i, c in enumerate(vals, 1):
        rows.append(f'0,{i},"{[float(v) for v in c]}",{int(rng.integers(1, 5000))}\n')
    return "".join(rows)


def main():
    import importlib.util
    from helpers import load_golden
    _pandas_1_4_replace()
    os.cpu_count = lambda: 8
    spec = importlib.util.spec_from_file_location("ref_automate", "/root/reference/automate_mBrainaligner.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    cases = {"g1": ("(64, 160, 128)_brainA.csv", load_golden("g1_notta")["csv"]), "awkward": ("(1500, 4000, 4000)_mouse 7.csv", awkward_csv())}
    out = {}
    for name, (fname, text) in cases.items():
        for xyz in (False, True):
            for par in (False, True):
                with tempfile.TemporaryDirectory() as tmp:
                    csv_path = os.path.join(tmp, fname)
                    with open(csv_path, "w") as f:
                        f.write(text)
                    od = os.path.join(tmp, "o")
                    os.makedirs(od)
                    files = ref.rewrite_swc(csv_path, od, XYZ=xyz, parallel_processing=par)
                    rec = {"csv_name": fname, "csv": text, "files": [[os.path.relpath(p, od), open(p).read()] for p in files]}
                    if not par and not xyz:
                        # a "registered" swc as mBrainAligner leaves it (same rows, shifted coordinates) -> size re-attach
                        swc = os.path.join(tmp, "local_registered_data.swc")
                        body = open(files[0]).read().splitlines()[1:]
                        with open(swc, "w") as f:
                            f.write("##n type x y z radius parent\n")
                            for ln in body:
                                n, t, x, y, z, r, p = ln.split(" ")
                                f.write(f"{n} {t} {float(x) * 0.5 + 1.25} {float(y) * 0.25} {float(z) + 3.0} {r} {p}\n")
                        coll = os.path.join(tmp, "coll")
                        os.makedirs(coll)
                        ref.reattach_size_and_copy(csv_path, swc, "mouseA", od, coll)
                        name_out = "mouseA_local_registered_with_original_size.csv"
                        rec["registered_swc"] = open(swc).read()
                        rec["reattached"] = open(os.path.join(od, name_out)).read()
                        assert rec["reattached"] == open(os.path.join(coll, name_out)).read()
                    rec["split_parameters"] = ref.split_parameters(csv_path)
                    out[f"{name}_xyz{int(xyz)}_par{int(par)}"] = rec
    with open(os.path.join(ROOT, "tests", "golden", "s1_swc.json"), "w") as f:
        json.dump(out, f)
    print({k: [len(v["files"]), sum(len(t) for _, t in v["files"])] for k, v in out.items()})


if __name__ == "__main__":
    main()

"""Drop-in for the part of the reference's ``automate_mBrainaligner.py`` that consumes count_blobs' table
(SURVEY.md section 8, row f2): ``rewrite_swc`` (:75-197), ``split_parameters`` (:199-213) and
``reattach_size_and_copy`` (:237-253).  Same arguments, same file names, byte-identical files - written as plain
text instead of through pandas DataFrames (the reference parses the ``Coords`` strings back with six regex passes
and re-serialises row by row).  ``swc_from_table`` emits the same SWC straight from the statistics table of
``dlv_ccl`` without the CSV round trip.  Host-side text only; the registration itself (mBrainAligner binaries) is out
of scope."""
import os
import re

import numpy as np

SWC_HEADER = "##n type x y z radius parent\n"


def _read_count_csv(csv_path):
    """The CSV count_blobs writes (count_blobs.py:101-114): ``,Blob,Coords,Size`` then ``0,i,"[z, y, x]",size``.
    -> (float64 [n,3] in file order, int64 [n])."""
    coords, sizes = [], []
    with open(csv_path) as f:
        header = f.readline()
        if header.strip() != ",Blob,Coords,Size":
            raise ValueError(f"{csv_path}: not a count_blobs table (header {header!r})")
        for line in f:
            if not line.strip():
                continue
            a, b = line.index('"['), line.index(']"')
            coords.append([float(t) for t in line[a + 2:b].replace(",", " ").split()])
            sizes.append(int(line[b + 3:]))
    return np.asarray(coords, dtype=np.float64).reshape(-1, 3), np.asarray(sizes, dtype=np.int64)


def _swc_lines(coords, sizes, XYZ=False):
    """One SWC row per cell: ``index 1 x y z Size -1`` with coordinates rounded to 3 decimals the way
    ``Series.round(3)`` does (numpy: scale, round half to even, unscale) and printed like ``DataFrame.to_csv``
    (shortest repr)."""
    c = np.round(np.asarray(coords, dtype=np.float64).reshape(-1, 3), 3)
    first, second, third = c[:, 0].tolist(), c[:, 1].tolist(), c[:, 2].tolist()
    # Coords are (z, y, x) unless XYZ (automate_mBrainaligner.py:100-108); columns go out as x y z
    xs, ys, zs = (first, second, third) if XYZ else (third, second, first)
    return [f"{i} 1 {x!r} {y!r} {z!r} {s} -1\n" for i, (x, y, z, s) in enumerate(zip(xs, ys, zs, np.asarray(sizes).tolist()))]


def _swc_name(output_dir, csv_name, suffix):
    name = os.path.join(output_dir, csv_name + suffix)
    return name.replace(" ", "").replace("(", "").replace(")", "")          # :163-166, :185-188


def _write_swcs(lines, csv_name, output_dir, parallel_processing):
    if not parallel_processing:
        target = _swc_name(output_dir, csv_name, ".swc")
        with open(target, "w") as f:
            f.write(SWC_HEADER)
            f.writelines(lines)
        print("successfully wrote " + str(target))
        return [target]
    n_chunks = os.cpu_count() - 1                                           # :150
    chunk_length = round(np.ceil(len(lines) / n_chunks))
    out = []
    for first in range(0, len(lines), chunk_length):
        target = _swc_name(output_dir, csv_name, "chunk_" + str(first).zfill(7) + ".swc")
        with open(target, "w") as f:
            f.write(SWC_HEADER)
            f.writelines(lines[first:first + chunk_length])
        print("successfully wrote " + str(target))
        out.append(target)
    return out


def rewrite_swc(csv_path, output_dir, XYZ=False, parallel_processing=False):
    """automate_mBrainaligner.py:75-197: count_blobs CSV -> SWC file(s) for the atlas registration.
    -> list of written paths (one, or cpu_count-1 chunks with ``parallel_processing``)."""
    coords, sizes = _read_count_csv(csv_path)
    return _write_swcs(_swc_lines(coords, sizes, XYZ), os.path.split(csv_path)[1], output_dir, parallel_processing)


def swc_from_table(stats, N, csv_name, output_dir, XYZ=False, parallel_processing=False):
    """The same files straight from the statistics table (rows 1..N-1, the rows count_blobs writes, count_blobs.py:104)
    - no CSV parse.  ``csv_name``: the name count_blobs gives its CSV, e.g. ``"(Z, Y, X)_brain.csv"``."""
    cent = np.asarray(stats["centroids"], dtype=np.float64)[1:N]
    cnt = np.asarray(stats["voxel_counts"])[1:N].astype(np.int64)
    return _write_swcs(_swc_lines(cent, cnt, XYZ), csv_name, output_dir, parallel_processing)


def split_parameters(file_path):
    """automate_mBrainaligner.py:199-213: the (Z, Y, X) in a file name -> [Z, Y, X]."""
    filename = os.path.split(file_path)[1]
    parameters = re.findall(r"\(([^)]+)", filename)
    return list(map(int, str(parameters[0]).replace(" ", "").split(sep=",")))


def _num(tok):
    try:
        return int(tok)
    except ValueError:
        return float(tok)


def reattach_size_and_copy(csv_path, swc_local, mouse_name, output_dir, aligned_results_folder):
    """automate_mBrainaligner.py:237-253: the registered SWC's rows with the ``Size`` column of the original CSV in
    place of radius / parent, written to ``<mouse>_local_registered_with_original_size.csv`` in both folders."""
    _, sizes = _read_count_csv(csv_path)
    rows = []
    with open(swc_local) as f:
        f.readline()
        for line in f:
            if line.strip():
                rows.append(line.rstrip("\n").split(" ")[:5])
    if len(rows) > len(sizes):
        raise ValueError(f"{swc_local} has {len(rows)} cells, {csv_path} only {len(sizes)} sizes")
    cols = list(zip(*rows)) if rows else [[]] * 5
    # pandas infers one dtype per column: integers stay integers, anything else becomes float (printed as repr)
    typed = []
    for col in cols:
        vals = [_num(t) for t in col]
        if any(isinstance(v, float) for v in vals):
            vals = [float(v) for v in vals]
        typed.append(vals)
    text = ["n type x y z Size\n"]
    for i in range(len(rows)):
        text.append(" ".join(repr(typed[k][i]) for k in range(5)) + f" {int(sizes[i])}\n")
    output_file_name = mouse_name + "_local_registered_with_original_size.csv"
    for folder in (output_dir, aligned_results_folder):
        with open(os.path.join(folder, output_file_name), "w") as f:
            f.writelines(text)

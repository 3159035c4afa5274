"""The union-pruning rule of ccl_merge_kernel (csrc/dlv_ccl.cu), restated in Python and checked against scipy.

The kernel unites every x-run with the runs it touches in its four raster-predecessor rows, but skips a union when
the neighbour run touches (x +- 1) another neighbour run that is already tied to the current run and whose row is
adjacent to its own - that union is made by the word owning the later of the two runs.  The rule is pure bit logic, so
it is pinned here on the CPU: the pairs the kernel would enumerate must give scipy's 26-connected partition
(cc3d.connected_components(connectivity=26), count_blobs.py:61), with and without the pruning.
"""
import numpy as np
import pytest
from scipy import ndimage


def runs(m):
    """(start, length) of the maximal runs of set bits of a non-negative int"""
    while m:
        a = (m & -m).bit_length() - 1
        ln = 0
        while (m >> (a + ln)) & 1:
            ln += 1
        yield a, ln
        m &= ~(((1 << ln) - 1) << a)


def kernel_pairs(mask, prune):
    """(label, label) unions in the kernel's enumeration; label = 1 + linear voxel index"""
    Z, Y, X = mask.shape
    W = (X + 31) // 32
    flat = mask.reshape(Z * Y, X)
    bits = [[0] * W for _ in range(Z * Y)]
    for r in range(Z * Y):
        for x in np.nonzero(flat[r])[0]:
            bits[r][x // 32] |= 1 << (int(x) % 32)
    pairs = []
    for r in range(Z * Y):
        z, y = divmod(r, Y)
        for w in range(W):
            cur = bits[r][w]
            if not cur:
                continue
            vbase = r * X + w * 32 + 1
            nrows = [r - 1, r - Y - 1, r - Y, r - Y + 1]
            ok = [y > 0, z > 0 and y > 0, z > 0, z > 0 and y + 1 < Y]
            comb = [0] * 4
            for k in range(4):
                if ok[k]:
                    nb = bits[nrows[k]]
                    comb[k] = (nb[w] << 1) | ((nb[w - 1] >> 31) if w > 0 else 0) | (((nb[w + 1] & 1) << 33) if w + 1 < W else 0)
            if w > 0 and (cur & 1) and (bits[r][w - 1] >> 31):
                pairs.append((vbase, vbase - 1))
            for a, ln in runs(cur):
                span = ((1 << (ln + 2)) - 1) << a
                c = [comb[k] & span for k in range(4)]
                cover = {2: 0, 0: c[2], 1: c[2] | c[0], 3: c[2]}
                for k in (2, 0, 1, 3):
                    for i, iln in runs(c[k]):
                        rb = ((1 << iln) - 1) << i
                        if prune and k != 2 and ((rb | (rb << 1) | (rb >> 1)) & cover[k]):
                            continue
                        pairs.append((vbase + a, nrows[k] * X + w * 32 + i))
    return pairs


def partition(mask, pairs):
    n = mask.size
    parent = list(range(n + 1))

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    def union(a, b):
        a, b = find(a), find(b)
        if a != b:
            parent[max(a, b)] = min(a, b)

    X = mask.shape[2]
    flat = mask.reshape(-1)
    fg = np.nonzero(flat)[0]
    for v in fg:                      # ccl_init: the voxels of an x-run inside one 32-voxel word share a label
        if (v % X) % 32 and flat[v - 1]:
            union(int(v) + 1, int(v))
    for a, b in pairs:
        union(a, b)
    lab = np.zeros(n, dtype=np.int64)
    for v in fg:
        lab[v] = find(int(v) + 1)
    return lab.reshape(mask.shape)


def same_partition(a, b):
    fg = a > 0
    if not np.array_equal(fg, b > 0):
        return False
    fwd, bwd = {}, {}
    for x, y in zip(a[fg].tolist(), b[fg].tolist()):
        if fwd.setdefault(x, y) != y or bwd.setdefault(y, x) != x:
            return False
    return True


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_pruned_unions_give_the_26_connected_partition(seed):
    rng = np.random.default_rng(seed)
    s26 = np.ones((3, 3, 3), dtype=int)
    full = pruned = 0
    for trial in range(24):
        Z, Y = int(rng.integers(1, 6)), int(rng.integers(1, 8))
        X = int(rng.choice([1, 5, 31, 32, 33, 40, 64, 70, 96]))
        dens = float(rng.choice([0.02, 0.1, 0.3, 0.5, 0.7, 0.9]))
        mask = (rng.random((Z, Y, X)) < dens).astype(np.uint8)
        if trial % 3 == 0:            # blob-like
            mask = ndimage.binary_dilation(rng.random((Z, Y, X)) < dens * 0.1, iterations=int(rng.integers(1, 3))).astype(np.uint8)
        ref, _ = ndimage.label(mask, structure=s26)
        pf, pp = kernel_pairs(mask, False), kernel_pairs(mask, True)
        assert same_partition(partition(mask, pf), ref), ("all unions", seed, trial)
        assert same_partition(partition(mask, pp), ref), ("pruned", seed, trial, (Z, Y, X), dens)
        assert set(pp) <= set(pf)
        full += len(pf)
        pruned += len(pp)
    assert pruned < full


def test_pruning_on_the_blob_field_of_the_cc_workload():
    from oracle import pipeline_ref as P
    mask = P.synth_mask((12, 64, 96), 1003)
    ref, _ = ndimage.label(mask, structure=np.ones((3, 3, 3), dtype=int))
    pf, pp = kernel_pairs(mask, False), kernel_pairs(mask, True)
    assert same_partition(partition(mask, pp), ref)
    assert len(pp) < 0.6 * len(pf)        # blobs: most unions are implied (cfg3-like field: about a third remain)

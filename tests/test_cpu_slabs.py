"""Host logic of the z-slab sharding (no GPU): planning, global label resolution, table merge, and the
distributed driver over gloo with world_size 2 (an oracle-backed worker stands in for the CUDA worker)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from delivr_cfos_b200 import slabs
from oracle import ccl_ref, pipeline_ref as P
from cpu_engine import OracleWorker as _OracleWorker, boundary_pairs_ref


def _starts(shape_pad, roi, overlap):
    return P.window_starts(shape_pad, roi, P.scan_interval(shape_pad, roi, overlap))


def test_plan_covers_volume_once():
    for shape, roi, world in [((1500, 400, 400), (96, 96, 64), 8), ((256, 64, 64), (96, 96, 64), 2), ((100, 80, 70), (32, 32, 32), 3),
                              ((64, 64, 64), (32, 32, 32), 5), ((40, 24, 24), (16, 16, 16), 7), ((20, 40, 40), (16, 16, 16), 6)]:
        pad = P.padded_shape(shape, roi)
        plan = slabs.SlabPlan(shape, roi, 0.5, world, starts=_starts(pad, roi, 0.5))
        owned = np.zeros(pad[0], int)
        wins = []
        nwin = len(plan.sz) * len(plan.sy) * len(plan.sx)
        for r in range(world):
            info = plan.rank(r)
            c0, c1 = plan.wrange[r]
            wins += list(range(c0, c1))
            owned[info["own"][0]:info["own"][1]] += 1
            assert len(plan.windows_of(r)) == c1 - c0
            if c1 > c0:
                w = plan.windows_of(r)
                assert info["win"] == (int(w[:, 0].min()), int(w[:, 0].max()) + roi[0])
                assert info["slab"][0] <= min(info["own"][0], info["win"][0]) and info["slab"][1] >= info["win"][1]
                assert info["slab"][0] <= max(0, info["own"][0] - 31) and info["slab"][1] >= min(pad[0], info["own"][1] + 31)
                for rng in (info["send"], info["recv"]):
                    if rng:
                        assert info["slab"][0] <= rng[0] < rng[1] <= info["slab"][1]
                if info["send"]:
                    q = plan._next_nonempty(r)
                    assert plan.rank(q)["recv"] == info["send"] and info["send"][0] == info["own"][1]
                # everything a rank touches beyond its own planes is handed on
                touched = max(info["win"][1], info["recv"][1] if info["recv"] else 0)
                assert (info["send"] is None) == (plan._next_nonempty(r) is None or touched <= info["own"][1])
        assert wins == list(range(nwin))                      # every window exactly once, in dense_patch_slices order
        assert (owned == 1).all()
        # balance: window counts differ by at most one between the ranks that have work
        sizes = [c1 - c0 for c0, c1 in plan.wrange if c1 > c0]
        assert max(sizes) - min(sizes) <= 1


def test_plan_balances_active_windows():
    shape, roi, world = (300, 200, 200), (32, 32, 32), 4
    pad = P.padded_shape(shape, roi)
    st = _starts(pad, roi, 0.5)
    n = len(st[0]) * len(st[1]) * len(st[2])
    act = (np.random.default_rng(0).random(n) < 0.6).astype(np.int32)
    act[: n // 5] = 0
    plan = slabs.SlabPlan(shape, roi, 0.5, world, starts=st, window_weights=act)
    per_rank = [int(act[c0:c1].sum()) for c0, c1 in plan.wrange]
    assert max(per_rank) - min(per_rank) <= 2 and sum(per_rank) == int(act.sum())


def _slab_tables(mask, cuts):
    tabs, labs = [], []
    for a, b in zip(cuts[:-1], cuts[1:]):
        lab, n = ccl_ref.connected_components26(mask[a:b])
        st = ccl_ref.statistics(lab, n)
        tabs.append({"n": n, **st})
        labs.append(lab)
    return tabs, labs


_pairs = boundary_pairs_ref


@pytest.mark.parametrize("seed,cuts", [(0, [0, 7, 20]), (1, [0, 5, 6, 13, 20]), (2, [0, 10, 11, 12, 20])])
def test_label_resolution_and_table_merge_exact(seed, cuts):
    rng = np.random.default_rng(seed)
    mask = (rng.random((20, 24, 28)) < 0.12).astype(np.uint8)
    mask[3:18, 5, 5] = 1                      # a component threading through every slab
    ref_lab, ref_n = ccl_ref.connected_components26(mask)
    ref = ccl_ref.statistics(ref_lab, ref_n)
    tabs, labs = _slab_tables(mask, cuts)
    pairs = [None] + [_pairs(labs[i - 1][-1], labs[i][0]) for i in range(1, len(labs))]
    luts, n = slabs.resolve_global_labels([t["n"] for t in tabs], pairs)
    assert n == ref_n
    merged = np.concatenate([lut[lab] for lut, lab in zip(luts, labs)])
    assert np.array_equal(merged, ref_lab)
    table = slabs.merge_tables(tabs, luts, cuts[:-1], n, mask.shape)
    assert np.array_equal(table["voxel_counts"], ref["voxel_counts"])
    assert np.array_equal(table["sums"], ref["sums"])
    assert np.array_equal(table["bounding_boxes"], ref["bounding_boxes"])
    assert np.array_equal(table["centroids"], ref["centroids"], equal_nan=True)


class OracleWorker(_OracleWorker):
    """The shared CPU stand-in (tests/cpu_engine.py) fed from a whole in-memory volume."""

    def __init__(self, plan, rank, volume):
        super().__init__(plan, rank, lambda z0, z1: volume[z0:z1])


def _make_volume(shape, roi):
    rng = np.random.default_rng(5)
    pad = P.padded_shape(shape, roi)
    vol = np.zeros(pad, dtype=np.uint16)
    vol[:shape[0], :shape[1], :shape[2]] = rng.integers(1, 60000, size=shape)
    vol[:2] = 0
    return vol


def _single(volume, shape, roi):
    pad = volume.shape
    plan = slabs.SlabPlan(shape, roi, 0.5, 1, starts=_starts(pad, roi, 0.5), erosion_iters=3)
    w = OracleWorker(plan, 0, volume)
    table = slabs.run_virtual([w], plan)
    return w, table


def _gloo_worker(rank, world, port, shape, roi, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    volume = _make_volume(shape, roi)
    plan = slabs.SlabPlan(shape, roi, 0.5, world, starts=_starts(volume.shape, roi, 0.5), erosion_iters=3)
    w = OracleWorker(plan, rank, volume)
    table = slabs.run_distributed(w, plan, slabs.TorchComm())
    q.put((rank, w.binaries.numpy(), w.labels.numpy(), table))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 7])      # 7: ranks cut inside window layers, one of them owning no plane
def test_distributed_driver_gloo(world):
    shape, roi = (40, 24, 24), (16, 16, 16)
    volume = _make_volume(shape, roi)
    w1, t1 = _single(volume, shape, roi)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, shape, roi, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    b = np.concatenate([r[1] for r in res])
    lab = np.concatenate([r[2] for r in res])
    assert np.array_equal(b, w1.binaries.numpy())
    assert np.array_equal(lab, w1.labels.numpy())
    for r in res:
        t = r[3]
        assert t["n"] == t1["n"] and t["n"] > 2
        assert np.array_equal(t["voxel_counts"], t1["voxel_counts"])
        assert np.array_equal(t["sums"], t1["sums"])
        assert np.array_equal(t["bounding_boxes"], t1["bounding_boxes"])


def test_virtual_slabs_equal_single_oracle():
    shape, roi = (50, 24, 24), (16, 16, 16)
    volume = _make_volume(shape, roi)
    w1, t1 = _single(volume, shape, roi)
    for world in (2, 3, 6, 12, 20):          # 12, 20: several ranks inside one window layer (ranks that own no plane, forwarding)
        plan = slabs.SlabPlan(shape, roi, 0.5, world, starts=_starts(volume.shape, roi, 0.5), erosion_iters=3)
        ws = [OracleWorker(plan, r, volume) for r in range(world)]
        t = slabs.run_virtual(ws, plan)
        assert np.array_equal(np.concatenate([w.binaries.numpy() for w in ws]), w1.binaries.numpy())
        assert np.array_equal(np.concatenate([w.labels.numpy() for w in ws]), w1.labels.numpy())
        assert t["n"] == t1["n"] and np.array_equal(t["sums"], t1["sums"])


@pytest.mark.parametrize("overlap,shape,roi", [(0.25, (50, 24, 24), (16, 16, 16)), (0.75, (40, 24, 24), (16, 16, 16)),
                                               (0.5, (33, 40, 20), (16, 32, 16))])
def test_virtual_slabs_other_overlaps(overlap, shape, roi):
    """BASELINE.json configs[4] sweeps the overlap: with 0.75 a plane is covered by four window layers, so the sums a
    rank touches beyond its own planes reach further down the chain; with 0.25 layers barely overlap."""
    volume = _make_volume(shape, roi)
    st = _starts(volume.shape, roi, overlap)
    plan1 = slabs.SlabPlan(shape, roi, overlap, 1, starts=st, erosion_iters=3)
    w1 = OracleWorker(plan1, 0, volume)
    t1 = slabs.run_virtual([w1], plan1)
    nwin = len(st[0]) * len(st[1]) * len(st[2])
    for world in (2, 3, 5, 9, nwin + 2):
        plan = slabs.SlabPlan(shape, roi, overlap, world, starts=st, erosion_iters=3)
        ws = [OracleWorker(plan, r, volume) for r in range(world)]
        t = slabs.run_virtual(ws, plan)
        assert np.array_equal(np.concatenate([w.binaries.numpy() for w in ws]), w1.binaries.numpy()), world
        assert np.array_equal(np.concatenate([w.labels.numpy() for w in ws]), w1.labels.numpy()), world
        assert t["n"] == t1["n"] and np.array_equal(t["sums"], t1["sums"]) and np.array_equal(t["bounding_boxes"], t1["bounding_boxes"])


def test_plan_invariants_random_configs():
    """Seeded sweep over window shapes, overlaps (0.25 / 0.5 / 0.75) and rank counts - the plan must hand every window
    to exactly one rank, every plane to exactly one owner, keep everything a rank touches inside its slab, and pass
    on whatever it touches beyond its own planes."""
    rng = np.random.default_rng(11)
    for _ in range(120):
        roi = tuple(int(16 * rng.integers(1, 5)) for _ in range(3))
        shape = tuple(int(rng.integers(r // 2 + 1, 5 * r)) for r in roi)
        overlap = float(rng.choice([0.25, 0.5, 0.75]))
        world = int(rng.integers(1, 12))
        pad = P.padded_shape(shape, roi)
        st = _starts(pad, roi, overlap)
        nwin = len(st[0]) * len(st[1]) * len(st[2])
        act = (rng.random(nwin) < 0.7).astype(np.int32)
        plan = slabs.SlabPlan(shape, roi, overlap, world, starts=st, window_weights=act if rng.random() < 0.5 else None)
        owned = np.zeros(pad[0], int)
        wins = []
        for r in range(world):
            info = plan.rank(r)
            c0, c1 = plan.wrange[r]
            wins += list(range(c0, c1))
            owned[info["own"][0]:info["own"][1]] += 1
            if c1 == c0:
                continue
            w = plan.windows_of(r)
            assert info["win"] == (int(w[:, 0].min()), int(w[:, 0].max()) + roi[0])
            touched = max(info["win"][1], info["recv"][1] if info["recv"] else 0)
            assert info["slab"][0] <= min(info["own"][0], info["win"][0]) and info["slab"][1] >= touched
            assert info["slab"][0] <= max(0, info["own"][0] - 31) and info["slab"][1] >= min(pad[0], info["own"][1] + 31)
            nxt = plan._next_nonempty(r)
            if nxt is None:
                assert info["send"] is None and info["own"][1] == pad[0]
            else:
                assert (info["send"] is None) == (touched <= info["own"][1])
                if info["send"]:
                    assert info["send"] == (info["own"][1], touched) == plan.rank(nxt)["recv"]
            # nothing a rank computes lands on a plane below its own planes
            assert info["win"][0] >= info["own"][0]
        assert wins == list(range(nwin)) and (owned == 1).all(), (shape, roi, overlap, world)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_virtual_slabs_random_configs(seed):
    rng = np.random.default_rng(seed)
    roi = tuple(int(16 * rng.integers(1, 3)) for _ in range(3))
    shape = tuple(int(rng.integers(r, 3 * r)) for r in roi)
    overlap = float(rng.choice([0.25, 0.5, 0.75]))
    volume = _make_volume(shape, roi)
    st = _starts(volume.shape, roi, overlap)
    plan1 = slabs.SlabPlan(shape, roi, overlap, 1, starts=st, erosion_iters=3)
    w1 = OracleWorker(plan1, 0, volume)
    t1 = slabs.run_virtual([w1], plan1)
    for world in (2, 4, 7):
        act = (rng.random(len(st[0]) * len(st[1]) * len(st[2])) < 0.8).astype(np.int32)
        plan = slabs.SlabPlan(shape, roi, overlap, world, starts=st, erosion_iters=3, window_weights=act)
        ws = [OracleWorker(plan, r, volume) for r in range(world)]
        t = slabs.run_virtual(ws, plan)
        assert np.array_equal(np.concatenate([w.binaries.numpy() for w in ws]), w1.binaries.numpy()), (shape, roi, overlap, world)
        assert np.array_equal(np.concatenate([w.labels.numpy() for w in ws]), w1.labels.numpy())
        assert t["n"] == t1["n"] and np.array_equal(t["voxel_counts"], t1["voxel_counts"])


def test_table_merge_threaded_path_matches_numpy():
    """dlv_table_merge splits the global rows over host threads above 200 000 rows: every field against a plain numpy
    merge (scatter adds / min / max), with merged components, empty rows (neutral boxes) and a rank without planes."""
    rng = np.random.default_rng(11)
    world, n_per = 4, 90000
    counts = [n_per, 0, n_per, n_per]
    pairs = [np.zeros((0, 2), np.uint32), np.zeros((0, 2), np.uint32),
             np.stack([rng.integers(1, n_per + 1, 5000), rng.integers(1, n_per + 1, 5000)], 1).astype(np.uint32),
             np.stack([rng.integers(1, n_per + 1, 5000), rng.integers(1, n_per + 1, 5000)], 1).astype(np.uint32)]
    order = [0, 2, 3]
    luts_o, n = slabs.resolve_global_labels([counts[q] for q in order], [pairs[0], pairs[2], pairs[3]])
    assert n + 1 > 200000
    luts = [luts_o[0], np.zeros(1, np.uint32), luts_o[1], luts_o[2]]

    def table(k):
        vc = rng.integers(0, 50, k + 1).astype(np.uint64)
        sums = (rng.integers(0, 10 ** 6, (k + 1, 3)).astype(np.uint64)) * vc[:, None]
        bb = np.zeros((k + 1, 6), np.int64)
        bb[:, 0::2] = rng.integers(0, 100, (k + 1, 3))
        bb[:, 1::2] = bb[:, 0::2] + rng.integers(0, 9, (k + 1, 3))
        empty = vc == 0
        bb[empty] = [100, -1, 100, -1, 100, -1]              # neutral box of a row without voxels
        return {"n": k, "voxel_counts": vc, "sums": sums, "bounding_boxes": bb}

    tabs = [table(n_per), None, table(n_per), table(n_per)]
    zoff = [0, 40, 40, 90]
    shape = (150, 100, 100)
    got = slabs.merge_tables(tabs, luts, zoff, n, shape)
    vc = np.zeros(n + 1, np.uint64)
    sm = np.zeros((n + 1, 3), np.uint64)
    bb = np.tile(np.array([shape[0], -1, shape[1], -1, shape[2], -1], np.int64), (n + 1, 1))
    for t, lut, z0 in zip(tabs, luts, zoff):
        if t is None:
            continue
        g = lut.astype(np.int64)
        np.add.at(vc, g, t["voxel_counts"])
        s = t["sums"].copy()
        s[:, 0] += t["voxel_counts"] * np.uint64(z0)
        for k in range(3):
            np.add.at(sm[:, k], g, s[:, k])
        ok = t["bounding_boxes"][:, 1] >= 0
        b = t["bounding_boxes"][ok].copy()
        b[:, 0:2] += z0
        for k in (0, 2, 4):
            np.minimum.at(bb[:, k], g[ok], b[:, k])
            np.maximum.at(bb[:, k + 1], g[ok], b[:, k + 1])
    assert got["n"] == n
    assert np.array_equal(got["voxel_counts"], vc)
    assert np.array_equal(got["sums"], sm)
    assert np.array_equal(got["bounding_boxes"], bb)
    with np.errstate(invalid="ignore", divide="ignore"):
        assert np.array_equal(got["centroids"], sm.astype(np.float64) / vc.astype(np.float64)[:, None], equal_nan=True)
    # the column-block wire format round-trips without copies
    vec = slabs.pack_table(tabs[2])
    back = slabs.unpack_table(vec)
    for k in ("voxel_counts", "sums", "bounding_boxes"):
        assert np.array_equal(back[k], tabs[2][k]) and np.shares_memory(back[k], vec)
    assert slabs.unpack_table(slabs.pack_table(None)) is None

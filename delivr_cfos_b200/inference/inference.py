"""Drop-in for the reference's ``inference/inference.py`` (:21-332).

``run_inference`` keeps the reference signature, reads the same config keys
(``blob_detection.window_dimensions``, ``FLAGS.SAVE_ACTIVATED_OUTPUT``,
``FLAGS.LOAD_ALL_RAM`` via ``load_all_ram``) and writes the same files:
``<out>/<brain>/binary_segmentations/binaries.npy`` (uint8 (Z,Y,X), 128-byte
header), optionally ``network_output.npy`` + the empty ``network_outputs/``
directory, and - when ``load_all_ram`` is false - ``inference_output.npy``
(fp16 averaged logits, padded shape).  All passes, the averaging and the
binarisation run on the GPU inside one ``dlv_segment`` call.
"""
import datetime
import os

import numpy as np

from .sliding_window_inferer import DelivrNet, SlidingWindowInferer  # noqa: F401

ARRAYTERATOR_BUF = 1000 ** 3     # inference.py:53,285


def update_idx(old_idx, new_idx, total_size):
    """inference.py:21-29."""
    for i in range(len(old_idx)):
        if new_idx[i] < total_size[i]:
            new_idx[i] += old_idx[i]
        if old_idx[i] == total_size[i]:
            old_idx[i] = 0
    return old_idx, new_idx


def erosion_block_planes(shape_real, buf=ARRAYTERATOR_BUF):
    """z-extent of the blocks numpy.lib.Arrayterator(arr[:Z,:Y,:X], buf) yields (inference.py:53); 0 = one block."""
    Z, Y, X = (int(s) for s in shape_real)
    count = buf // X
    if count <= Y:
        raise NotImplementedError("planes larger than the 1e9-element Arrayterator buffer (y-split blocks)")
    count //= Y
    return 0 if count >= Z else int(count)


def create_empty_memmap(file_location, shape, dtype=np.uint16, return_torch=True, torch_dtype=None):
    """inference.py:98-109: zeroed .npy on disk (128-byte header), optionally as a torch tensor copy."""
    try:
        os.remove(file_location)
    except OSError:
        pass
    empty_memmap = np.lib.format.open_memmap(file_location, mode="w+", dtype=dtype, shape=tuple(shape))
    if return_torch:
        import torch
        empty_memmap = torch.as_tensor(empty_memmap, dtype=torch_dtype or torch.float16)
    return empty_memmap


def create_nifti_seg(threshold, model_output, output_file, network_output_file, dataset, original_stack_shape,
                     device=0, _ctx=None):
    """inference.py:31-95 on the GPU (dlv_op_finalise): sigmoid >= threshold AND erode30(dataset > 0) per block."""
    import torch
    from .._lib import Context
    ctx = _ctx or Context(device)
    shape_real = tuple(int(s) for s in original_stack_shape[2:])
    mo = model_output.numpy() if hasattr(model_output, "numpy") else np.asarray(model_output)
    shape_pad = tuple(int(s) for s in mo.shape[2:])
    binarized = np.lib.format.open_memmap(output_file, mode="w+", dtype=np.uint8, shape=shape_real)
    avg = torch.from_numpy(np.ascontiguousarray(mo[0, 0]).astype(np.float32)).cuda(device)
    vol = torch.from_numpy(np.ascontiguousarray(np.asarray(dataset)[0, 0])).cuda(device)
    out = torch.empty(shape_real, dtype=torch.uint8, device=f"cuda:{device}")
    sig = torch.empty(shape_real, dtype=torch.float32, device=f"cuda:{device}") if network_output_file is not None else None
    ctx.op_finalise(avg, vol, shape_pad, shape_real, out, threshold, 30, erosion_block_planes(shape_real), sig)
    binarized[...] = out.cpu().numpy()
    binarized.flush()
    if network_output_file is not None:
        act = np.lib.format.open_memmap(network_output_file, mode="w+", dtype=np.float32, shape=shape_real)
        act[...] = sig.cpu().numpy()
        act.flush()


# ------------------------------------------------------------------------------------------- execution engines
class CudaEngine:
    """What run_inference needs from the device side: the loaded network, free memory, the in-core call, slab workers
    and host copies.  The only implementation in the product is this one (libdelivr_b200.so on a B200); there is no CPU
    engine - a missing library or device raises DlvError.  (tests/ substitutes an oracle-backed engine to exercise the
    host logic of the out-of-core and multi-rank modes without a GPU.)"""

    backend = "nccl"

    def __init__(self, model_weights, device):
        import torch
        self.torch = torch
        self.device = int(device)
        self.net = DelivrNet(checkpoint_path=os.path.abspath(str(model_weights)), device=self.device)
        self.ctx = self.net.ctx
        self.dev = torch.device("cuda", self.device)
        self.stream = torch.cuda.ExternalStream(self.ctx._L.dlv_stream(self.ctx._h), device=self.dev)

    def free_bytes(self):
        return int(self.torch.cuda.mem_get_info(self.device)[0])

    def comm(self):
        from ..slabs import TorchComm
        return TorchComm(self.dev)

    def planes_fn(self, source):
        """source: the padded uint16 volume as a (Zp,Yp,Xp) numpy array / memmap or a device tensor.
        -> f(z0, z1) = device tensor of those planes (read from the file only for the planes asked for)."""
        torch = self.torch
        if hasattr(source, "is_cuda"):
            return lambda z0, z1: source[z0:z1]

        plane = int(source.shape[1]) * int(source.shape[2])
        group = max(1, (256 << 20) // (2 * plane))                  # planes per staging buffer (~256 MB)
        stage = [None, None]                                        # two pinned buffers: the file read of one group
        done = [None, None]                                         # overlaps the H2D copy of the previous one

        def load(z0, z1):
            with torch.cuda.stream(self.stream):
                dst = torch.empty((z1 - z0,) + tuple(source.shape[1:]), dtype=torch.uint16, device=self.dev)
                for i, a in enumerate(range(z0, z1, group)):
                    b, k = min(z1, a + group), i & 1
                    if stage[k] is None:
                        stage[k] = torch.empty((group,) + tuple(source.shape[1:]), dtype=torch.uint16).pin_memory()
                    elif done[k] is not None:
                        done[k].synchronize()
                    np.copyto(stage[k].numpy()[:b - a], source[a:b])
                    dst[a - z0:b - z0].copy_(stage[k][:b - a], non_blocking=True)
                    done[k] = torch.cuda.Event()
                    done[k].record(self.stream)
            return dst
        return load

    def segment_incore(self, volume, shape_pad, shape_real, roi, binarized, **kw):
        return self.ctx.segment(volume, shape_pad, shape_real, roi, binarized, **kw)

    def windows_active(self, slab, local_windows, roi):
        return self.ctx.windows_active(slab, local_windows, roi)

    def make_worker(self, plan, r, planes_fn, **kw):
        from ..slabs import CudaSlabWorker
        return CudaSlabWorker(self.ctx, plan, r, planes_fn, **kw)

    def to_host(self, t):
        self.ctx.synchronize()
        return t.cpu().numpy()


# Test seam, not a fallback: the CPU tests of the host logic (chunking, rank start-up, file contract) install their own
# engine here - or name its factory in DLV_ENGINE for the rank processes run_inference spawns - so that this module can be
# exercised without a GPU.  Nothing in the package ever sets either; with neither set the engine is the CUDA library, and
# a missing .so / non-sm_100 device raises DlvError at the first call (tests/test_cpu_library.py checks both).
_ENGINE_FACTORY = CudaEngine


def _engine_factory():
    name = os.environ.get("DLV_ENGINE")
    if name:
        import importlib
        mod, attr = name.split(":")
        return getattr(importlib.import_module(mod), attr)
    return _ENGINE_FACTORY


ACTIVATION_RESERVE = 44 << 30       # window-batch activations (<= ~36 GB at the default batch) + CCL / erosion scratch


def incore_bytes(shape_pad, shape_real, save_act, keep_avg):
    """Device bytes dlv_segment holds at its peak: padded volume (2 B) + int32 blend sums (4 B) per padded voxel,
    binaries + erosion scratch (2 B) per real voxel, fp32 sigmoid (4 B) if saved (csrc/dlv_segment.cu)."""
    vp = int(np.prod(shape_pad))
    vr = int(np.prod(shape_real))
    return vp * 6 + vr * (2 + (4 if save_act else 0))


def chunks_needed(shape_pad, shape_real, roi, free_bytes, save_act, keep_avg):
    """Number of z-chunks (virtual slabs) for the out-of-core mode of one GPU: the smallest k whose slab (own planes +
    one window depth + the erosion halo either side) fits next to the activation reserve.  1 = in-core."""
    budget = free_bytes - ACTIVATION_RESERVE
    if incore_bytes(shape_pad, shape_real, save_act, keep_avg) <= budget:
        return 1
    PZ, PY, PX = (int(v) for v in shape_pad)
    per_plane = PY * PX * (6 + 2 + (4 if save_act else 0))
    nlayers = max(1, 2 * PZ // int(roi[0]) - 1)
    for k in range(2, nlayers + 1):
        planes = -(-PZ // k) + 2 * int(roi[0]) + 64
        if planes * per_plane <= budget:
            return k
    raise MemoryError(f"a single window layer of {PY}x{PX} planes does not fit the device ({free_bytes >> 30} GiB free)")


def _distributed_env():
    """(rank, world, local_rank) of a torchrun-style launch (RANK / WORLD_SIZE / LOCAL_RANK), else (0, 1, 0)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        return int(os.environ.get("RANK", 0)), world, int(os.environ.get("LOCAL_RANK", os.environ.get("RANK", 0)))
    return 0, 1, 0


def _ensure_process_group(backend, device):
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        if backend == "nccl":
            torch.cuda.set_device(device)
            dist.init_process_group("nccl", device_id=torch.device("cuda", device))
        else:
            dist.init_process_group(backend)
    return dist


class _Outputs:
    """The reference's output files as .npy memmaps, written plane range by plane range (rank 0 creates them)."""

    def __init__(self, output_file, network_output_file, avg_file, shape_real, shape_pad5, create):
        self.paths = {"bin": (output_file, np.uint8, tuple(shape_real)),
                      "sig": (network_output_file, np.float32, tuple(shape_real)),
                      "avg": (avg_file, np.float16, tuple(shape_pad5))}
        if create:
            for path, dtype, shape in self.paths.values():
                if path:
                    m = np.lib.format.open_memmap(path, mode="w+", dtype=dtype, shape=shape)
                    del m
        self.maps = {}

    def open(self):
        for k, (path, _, _) in self.paths.items():
            if path:
                self.maps[k] = np.load(path, mmap_mode="r+")

    def write(self, engine, w):
        """Copy slab worker w's final planes (binaries / sigmoid: own real planes; averaged logits: own padded planes)."""
        o0, o1 = w.info["own_real"]
        if o1 > o0:
            self.maps["bin"][o0:o1] = engine.to_host(w.binaries)
            if "sig" in self.maps:
                self.maps["sig"][o0:o1] = engine.to_host(w.sigmoid)
        if "avg" in self.maps and w.avg_own is not None:
            a0, a1 = w.info["own"]
            with np.errstate(over="ignore"):
                self.maps["avg"][0, 0, a0:a1] = engine.to_host(w.avg_own).astype(np.float16)

    def close(self):
        for m in self.maps.values():
            m.flush()
        self.maps = {}


def _spawn_ranks(n, kwargs):
    """``python __main__.py`` on a box with several GPUs: run_inference re-enters itself in n child processes, one per
    GPU (the reference spreads its batches over the visible GPUs inside one process, inference/inference.py:217-219)."""
    import json
    import socket
    import subprocess
    import sys
    import tempfile
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as f:
        json.dump(kwargs, f)
        payload = f.name
    procs = []
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    for r in range(n):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(n), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))
        env.pop("DLV_GPUS", None)
        procs.append(subprocess.Popen([sys.executable, "-m", "delivr_cfos_b200.inference.rank_main", payload], env=env))
    codes = [p.wait() for p in procs]
    os.unlink(payload)
    if any(codes):
        raise RuntimeError(f"run_inference: rank processes exited with {codes}")


def run_inference(niftis, output_folder, stack_shape, comment="none", model_weights="weights/inference_weights.tar",
                  tta=False, threshold=0.5, cuda_devices="0,1", crop_size=(64, 64, 32), workers=0, sw_batch_size=100,
                  overlap=0.5, verbosity=True, load_all_ram=False, settings=None, blend="constant", device=None, _net=None,
                  volume=None):
    """Sliding-window U-Net inference + binarisation; same contract as the reference (inference.py:113-332).

    Three execution modes, chosen from what is there (all write the same files, bit-identical binaries):

    * in-core - the padded volume, its int32 blend sums and the outputs fit one GPU: ONE ``dlv_segment`` call;
    * out-of-core - they do not (a 1500x4000x4000 brain needs ~200 GB): the volume is cut into z-chunks at window
      granularity that are read from the memmap, run and written one after the other (``slabs.run_streamed``; the
      reference streams from memmaps too, inference.py:234,244-247);
    * multi-GPU - launched under ``torchrun`` (one rank per GPU; RANK / WORLD_SIZE set) every rank reads only its
      z-slab and writes only its planes (``slabs.distributed_segment``, NCCL halo exchange).  Without torchrun,
      ``DLV_GPUS=<n>|all`` makes run_inference start the n rank processes itself; with ``DLV_GPUS`` unset this happens
      automatically when several GPUs are visible and the volume does not fit one of them.

    ``volume`` (extension): a device-resident uint16 ``(Zp, Yp, Xp)`` tensor from
    ``tiff_planes.load_masked_volume`` - then ``niftis`` is not read (no masked_nifti.npy round trip).

    ``cuda_devices``, ``workers`` and ``sw_batch_size`` are accepted for signature compatibility; the window batch is
    chosen by the library (the reference derives it from free VRAM, inference.py:171-187).

    Deviations from the reference's numerics: bf16 tensor-core arithmetic and an exact
    integer blend instead of a running fp16 sum (binaries agree >= 99.9 %, not bit for bit); the skip rule is applied
    per window, not per VRAM-sized batch, and the TTA noise (sigma <= 1e-3 on integer intensities) is dropped - so
    ``network_output.npy`` / ``inference_output.npy`` differ from the reference's where the eroded mask is 0.
    """
    print(f"{datetime.datetime.now()} : Setting up inference parameters ")
    if settings is not None:
        wd = settings["blob_detection"]["window_dimensions"]
        crop_size = (wd["window_dim_0"], wd["window_dim_1"], wd["window_dim_2"])
    crop_size = tuple(int(c) for c in crop_size)
    print("using crop size:  ", crop_size)

    rank, world, local_rank = _distributed_env()
    if device is None:
        device = local_rank
    stack_shape_pad = list(stack_shape)
    for idx, dim in enumerate(stack_shape_pad[2:]):
        stack_shape_pad[idx + 2] = int(np.ceil(dim / crop_size[idx]) * crop_size[idx])
    shape_pad = tuple(int(v) for v in stack_shape_pad[2:])
    shape_real = tuple(int(s) for s in stack_shape[2:])
    save_act = bool(settings["FLAGS"]["SAVE_ACTIVATED_OUTPUT"]) if settings is not None else False
    keep_avg = not load_all_ram
    blend_mode = {"constant": 0, "gaussian": 1}[blend]

    # ---- several GPUs without torchrun: start the rank processes and let them do the work
    want = os.environ.get("DLV_GPUS", "")
    if world == 1 and volume is None and _net is None and want not in ("", "1"):
        import torch
        ngpu = torch.cuda.device_count() if want == "all" else int(want)
        if ngpu > 1:
            _spawn_ranks(ngpu, dict(niftis=[str(n) for n in niftis], output_folder=str(output_folder), stack_shape=[int(v) for v in stack_shape],
                                    comment=comment, model_weights=str(model_weights), tta=bool(tta), threshold=threshold,
                                    crop_size=list(crop_size), overlap=overlap, load_all_ram=bool(load_all_ram),
                                    settings=settings, blend=blend))
            print(f"{datetime.datetime.now()} : Blob Detection finished")
            return os.path.abspath(output_folder + "/" + comment)

    engine = _net if _net is not None and hasattr(_net, "segment_incore") else None
    if engine is None:
        if _net is not None:                      # a DelivrNet handed in by the caller (tests): wrap it
            engine = CudaEngine.__new__(CudaEngine)
            import torch
            engine.torch, engine.device, engine.net, engine.ctx = torch, _net.ctx.device, _net, _net.ctx
            engine.dev = torch.device("cuda", engine.device)
            engine.stream = torch.cuda.ExternalStream(engine.ctx._L.dlv_stream(engine.ctx._h), device=engine.dev)
        else:
            engine = _engine_factory()(model_weights, device)

    print(f"{datetime.datetime.now()} : Loading Data")
    if volume is None:
        dataset = np.memmap(str(niftis[0]), dtype=np.uint16, mode="r", shape=tuple(stack_shape_pad), offset=128)
        source = dataset[0, 0]
    elif tuple(volume.shape) != shape_pad:
        raise ValueError(f"volume has shape {tuple(volume.shape)}, expected the padded stack shape {shape_pad}")
    else:
        source = volume

    testing_session_path = os.path.abspath(output_folder + "/" + comment)
    binaries_path = testing_session_path + "/binary_segmentations/"
    output_file = os.path.join(binaries_path, "binaries.npy")
    network_output_file = os.path.join(binaries_path, "network_output.npy") if save_act else None
    avg_file = os.path.join(output_folder, comment, "inference_output.npy") if keep_avg else None
    if rank == 0:
        os.makedirs(binaries_path, exist_ok=True)
        if save_act:
            os.makedirs(testing_session_path + "/network_outputs/", exist_ok=True)

    print(f"{datetime.datetime.now()} : Starting inference")
    ebp = erosion_block_planes(shape_real)
    kw = dict(threshold=threshold, tta=bool(tta), erosion_block_planes=ebp, blend_mode=blend_mode)
    if world > 1:
        # ------------------------------------------------------------ one rank per GPU
        from .. import slabs
        if volume is not None:
            raise ValueError("volume= (device tensor) is a single-process input; ranks read their planes from the .npy")
        _ensure_process_group(engine.backend, device)
        comm = engine.comm()
        planes = engine.planes_fn(source)
        if os.environ.get("DLV_BALANCE", "1") != "0":
            plan, _ = slabs.balanced_plan(engine, comm, shape_real, crop_size, overlap, planes)
        else:
            plan = slabs.SlabPlan(shape_real, crop_size, overlap, world)
        outs = _Outputs(output_file, network_output_file, avg_file, shape_real, stack_shape_pad, create=rank == 0)
        comm.barrier()
        outs.open()
        w = engine.make_worker(plan, rank, planes, want_sigmoid=save_act, keep_avg=keep_avg, **kw)
        with slabs._stream_ctx(w):
            active = slabs.distributed_segment(w, plan, comm)
        outs.write(engine, w)
        outs.close()
        comm.barrier()
        nact, ntot = int(np.sum(active)), len(active)
        print(f"{datetime.datetime.now()} : Inference done (rank {rank}/{world}: windows {plan.wrange[rank]}, "
              f"{nact}/{ntot} active windows x {13 if tta else 1} passes)")
    else:
        k = 1 if hasattr(source, "is_cuda") else chunks_needed(shape_pad, shape_real, crop_size, engine.free_bytes(), save_act, keep_avg)
        want_gpus = os.environ.get("DLV_GPUS", "")
        if k > 1 and want_gpus == "" and _net is None:
            import torch
            if torch.cuda.device_count() > 1:     # does not fit one GPU and there are more: use them
                os.environ["DLV_GPUS"] = "all"
                try:
                    return run_inference(niftis, output_folder, stack_shape, comment, model_weights, tta, threshold, cuda_devices,
                                         crop_size, workers, sw_batch_size, overlap, verbosity, load_all_ram, settings, blend)
                finally:
                    os.environ.pop("DLV_GPUS", None)
        if k == 1:
            # -------------------------------------------------------- in-core: one dlv_segment call
            binarized = np.lib.format.open_memmap(output_file, mode="w+", dtype=np.uint8, shape=shape_real)
            activated = (np.lib.format.open_memmap(network_output_file, mode="w+", dtype=np.float32, shape=shape_real)
                         if network_output_file else None)
            avg = np.empty(shape_pad, dtype=np.float32) if keep_avg else None
            vol_in = source if hasattr(source, "is_cuda") else np.ascontiguousarray(source)
            st = engine.segment_incore(vol_in, shape_pad, shape_real, crop_size, binarized, overlap=overlap, erosion_iters=30,
                                       avg_logits_out=avg, sigmoid_out=activated, **kw)
            print(f"{datetime.datetime.now()} : Inference done ({st['windows_active']}/{st['windows_total']} windows x "
                  f"{st['passes']} passes, {st['ms_unet']:.1f} ms network, {st['ms_finalise']:.1f} ms binarisation)")
            binarized.flush()
            if activated is not None:
                activated.flush()
            if avg is not None:
                # the reference leaves the averaged fp16 logits in inference_output.npy when not LOAD_ALL_RAM (inference.py:246)
                out = np.lib.format.open_memmap(avg_file, mode="w+", dtype=np.float16, shape=tuple(stack_shape_pad))
                with np.errstate(over="ignore"):
                    out[0, 0] = avg.astype(np.float16)
                out.flush()
        else:
            # -------------------------------------------------------- out-of-core: z-chunks one after the other
            from .. import slabs
            print(f"{datetime.datetime.now()} : volume exceeds device memory, running {k} z-chunks")
            plan = slabs.SlabPlan(shape_real, crop_size, overlap, k)
            outs = _Outputs(output_file, network_output_file, avg_file, shape_real, stack_shape_pad, create=True)
            outs.open()
            planes = engine.planes_fn(source)
            slabs.run_streamed(lambda r: engine.make_worker(plan, r, planes, want_sigmoid=save_act, keep_avg=keep_avg, **kw),
                               plan, sink=lambda r, w: outs.write(engine, w), label=False)
            outs.close()
            print(f"{datetime.datetime.now()} : Inference done ({k} z-chunks)")
    print(f"{datetime.datetime.now()} : Blob Detection finished")
    return testing_session_path

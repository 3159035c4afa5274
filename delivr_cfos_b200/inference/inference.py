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


def run_inference(niftis, output_folder, stack_shape, comment="none", model_weights="weights/inference_weights.tar",
                  tta=False, threshold=0.5, cuda_devices="0,1", crop_size=(64, 64, 32), workers=0, sw_batch_size=100,
                  overlap=0.5, verbosity=True, load_all_ram=False, settings=None, blend="constant", device=0, _net=None,
                  volume=None):
    """Sliding-window U-Net inference + binarisation; same contract as the reference (inference.py:113-332).

    ``volume`` (extension): a device-resident uint16 ``(Zp, Yp, Xp)`` tensor from
    ``tiff_planes.load_masked_volume`` - then ``niftis`` is not read (no masked_nifti.npy round trip).

    ``cuda_devices``, ``workers`` and ``sw_batch_size`` are accepted for signature compatibility; the window batch is
    chosen by the library (the reference derives it from free VRAM, inference.py:171-187).
    """
    print(f"{datetime.datetime.now()} : Setting up inference parameters ")
    if settings is not None:
        wd = settings["blob_detection"]["window_dimensions"]
        crop_size = (wd["window_dim_0"], wd["window_dim_1"], wd["window_dim_2"])
    crop_size = tuple(int(c) for c in crop_size)
    print("using crop size:  ", crop_size)

    net = _net or DelivrNet(checkpoint_path=os.path.abspath(str(model_weights)), device=device)

    print(f"{datetime.datetime.now()} : Loading Data")
    stack_shape_pad = list(stack_shape)
    for idx, dim in enumerate(stack_shape_pad[2:]):
        stack_shape_pad[idx + 2] = int(np.ceil(dim / crop_size[idx]) * crop_size[idx])
    if volume is None:
        dataset = np.memmap(str(niftis[0]), dtype=np.uint16, mode="r", shape=tuple(stack_shape_pad), offset=128)
    elif tuple(volume.shape) != tuple(stack_shape_pad[2:]):
        raise ValueError(f"volume has shape {tuple(volume.shape)}, expected the padded stack shape {tuple(stack_shape_pad[2:])}")
    shape_pad = tuple(stack_shape_pad[2:])
    shape_real = tuple(int(s) for s in stack_shape[2:])

    os.makedirs(os.path.join(output_folder, comment), exist_ok=True)
    testing_session_path = os.path.abspath(output_folder + "/" + comment)
    binaries_path = testing_session_path + "/binary_segmentations/"
    os.makedirs(binaries_path, exist_ok=True)
    output_file = os.path.join(binaries_path, "binaries.npy")
    save_act = bool(settings["FLAGS"]["SAVE_ACTIVATED_OUTPUT"]) if settings is not None else False
    network_output_file = None
    if save_act:
        os.makedirs(testing_session_path + "/network_outputs/", exist_ok=True)
        network_output_file = os.path.join(binaries_path, "network_output.npy")

    print(f"{datetime.datetime.now()} : Starting inference")
    binarized = np.lib.format.open_memmap(output_file, mode="w+", dtype=np.uint8, shape=shape_real)
    activated = (np.lib.format.open_memmap(network_output_file, mode="w+", dtype=np.float32, shape=shape_real)
                 if network_output_file else None)
    avg = None if load_all_ram else np.empty(shape_pad, dtype=np.float32)
    if volume is None:
        volume = np.ascontiguousarray(dataset[0, 0])
    st = net.ctx.segment(volume, shape_pad, shape_real, crop_size, binarized, overlap=overlap, tta=bool(tta),
                         threshold=threshold, erosion_iters=30, erosion_block_planes=erosion_block_planes(shape_real),
                         blend_mode={"constant": 0, "gaussian": 1}[blend], avg_logits_out=avg, sigmoid_out=activated)
    print(f"{datetime.datetime.now()} : Inference done ({st['windows_active']}/{st['windows_total']} windows x "
          f"{st['passes']} passes, {st['ms_unet']:.1f} ms network, {st['ms_finalise']:.1f} ms binarisation)")
    binarized.flush()
    if activated is not None:
        activated.flush()
    if avg is not None:
        # the reference leaves the averaged fp16 logits in inference_output.npy when not LOAD_ALL_RAM (inference.py:246)
        out = np.lib.format.open_memmap(os.path.join(output_folder, comment, "inference_output.npy"), mode="w+",
                                        dtype=np.float16, shape=tuple(stack_shape_pad))
        with np.errstate(over="ignore"):
            out[0, 0] = avg.astype(np.float16)
        out.flush()
    print(f"{datetime.datetime.now()} : Blob Detection finished")
    return testing_session_path

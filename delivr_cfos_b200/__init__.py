"""delivr_cfos_b200 - B200 (sm_100a) implementation of DELiVR's blob_detection hot path.

Public surface mirrors the reference's modules for this path:

* ``delivr_cfos_b200.inference.inference``            run_inference, create_nifti_seg, create_empty_memmap, update_idx
* ``delivr_cfos_b200.inference.sliding_window_inferer`` SlidingWindowInferer, sliding_window_inference
* ``delivr_cfos_b200.count_blobs``                     count_blobs, load_cached_brain, load_cached_stats

All arithmetic runs in ``libdelivr_b200.so`` (hand-written CUDA, C ABI in
``include/delivr_b200.h``); importing this package does not load the library,
calling any compute entry point does and raises if it is missing.
"""
from ._lib import Context, DlvError, load_library  # noqa: F401

__version__ = "0.1.0"

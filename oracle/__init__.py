"""CPU oracle for the DELiVR blob_detection hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``delivr_cfos_b200`` imports this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may.

Contents
--------
``unet_ref``      torch-fp32 restatement of MONAI 1.2.0 ``BasicUNet`` as the
                  reference instantiates it (inference/inference.py:190-197).
``pipeline_ref``  numpy restatement of sliding_window_inference,
                  run_inference's averaging, create_nifti_seg and count_blobs
                  (inference/sliding_window_inferer.py:33-276,
                  inference/inference.py:21-95,229-299, count_blobs.py:36-118).
``ccl_ref``       ctypes wrapper around ``ccl_ref.c`` - plain-C 26-connected
                  labelling, statistics and L1 erosion (what the reference gets
                  from cc3d 3.12.3 and scipy.ndimage.binary_erosion).
``shims/``        minimal stand-ins for monai / cc3d / path / nibabel / skimage
                  so that the reference's own three files import UNMODIFIED in
                  the authoring container (``make_golden.py``).

Parity pin status (see DESIGN.md "Oracle"): the reference ships no tests and
no golden vectors.  The restatements are pinned against the reference's own
files run here through the shims (``make_golden.py`` -> ``tests/golden``);
the third-party arithmetic itself (MONAI / cc3d, absent from the image) is
restated from its published behaviour and cross-checked against torch and
scipy.ndimage - for those two libraries parity is "unpinned".
"""

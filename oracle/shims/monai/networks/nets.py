"""monai.networks.nets.BasicUNet -> oracle restatement (inference/inference.py:15,190-197)."""
from oracle.unet_ref import BasicUNet  # noqa: F401

"""Shim for monai 1.2.0 (requirements.txt:21): only what the reference's hot path imports."""
__version__ = "1.2.0-shim"

"""Import-only stub (inference/inference.py:7)."""

"""Import-only stub (inference/inference.py:4; filehandling.py:4)."""

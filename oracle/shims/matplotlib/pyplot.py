"""import-only stub (blob_depthmap.py:13)."""

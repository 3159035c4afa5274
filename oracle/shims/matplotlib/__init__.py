"""import-only stub (blob_depthmap.py:13 imports matplotlib.pyplot; nothing on the painter path plots)."""

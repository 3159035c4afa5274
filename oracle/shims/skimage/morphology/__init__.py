"""import-only stub (blob_highlighter.py:13 imports binary_dilation but never calls it)."""


def binary_dilation(*a, **k):
    raise NotImplementedError("skimage.morphology stub: not used on the hot path")

"""import-only stub (blob_highlighter.py:14 imports ellipsoid but never calls it)."""


def ellipsoid(*a, **k):
    raise NotImplementedError("skimage.draw stub: not used on the hot path")

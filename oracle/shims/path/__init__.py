"""Shim for path.py (inference/inference.py:5,199,307; __main__.py:22-23,111-132,171)."""
import fnmatch
import os


class Path(str):
    def __add__(self, other):
        return Path(str.__add__(self, other))

    def __truediv__(self, other):
        return Path(os.path.join(self, other))

    @property
    def name(self):
        return Path(os.path.basename(self))

    @property
    def parent(self):
        return Path(os.path.dirname(self))

    def dirs(self):
        return [Path(os.path.join(self, d)) for d in os.listdir(self) if os.path.isdir(os.path.join(self, d))]

    def files(self, pattern=None):
        out = [Path(os.path.join(self, f)) for f in os.listdir(self) if os.path.isfile(os.path.join(self, f))]
        return [f for f in out if pattern is None or fnmatch.fnmatch(os.path.basename(f), pattern)]

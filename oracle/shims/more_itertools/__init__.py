"""Shim for more_itertools (automate_mBrainaligner.py:17,153): only ``sliced``.

Published behaviour: ``sliced(seq, n)`` yields ``seq[0:n], seq[n:2n], ...`` until an empty slice; the reference calls
it on a pandas DataFrame (row slices)."""


def sliced(seq, n):
    i = 0
    while i < len(seq):
        yield seq[i:i + n]
        i += n

"""monai.inferers.inferer.Inferer base (inference/sliding_window_inferer.py:21,278,332)."""


class Inferer:
    def __init__(self):
        pass

    def __call__(self, inputs, network, *args, **kwargs):
        raise NotImplementedError

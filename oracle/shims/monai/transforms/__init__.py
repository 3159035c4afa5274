"""monai.transforms.RandGaussianNoise (inference/sliding_window_inferer.py:22,215).

MONAI 1.2.0: ``img + N(mean, U(0, std))`` with the sigma drawn once per call
from an unseeded numpy RandomState, noise generated in float32 on the host.
"""
import numpy as np
import torch


class RandGaussianNoise:
    def __init__(self, prob=0.1, mean=0.0, std=0.1, dtype=np.float32):
        self.prob, self.mean, self.std, self.dtype = prob, mean, std, dtype
        self.R = np.random.RandomState()

    def __call__(self, img):
        if self.R.rand() >= self.prob:
            return img
        sigma = self.R.uniform(0, self.std)
        noise = self.R.normal(self.mean, sigma, size=tuple(img.shape)).astype(self.dtype)
        return img + torch.as_tensor(noise, device=img.device)

"""tfp.math.psd_kernels (tensorflow-probability 0.8.0) restated -- TEST INFRASTRUCTURE.

The kernel arithmetic of the reference is not in its tree: it is the un-vendored pip
dependency ``tensorflow-probability==0.8.0`` (requirements.txt:9).  Reference call
sites: SVGPVAE_model.py:60 (ball RBF), :416-417 (MNIST ExpSinSquared x Linear),
:542-548 (SPRITES SE / Linear); ``.matrix`` / ``.apply`` at :81-86, :152-157,
:459-469, :573-587.

Published definitions (feature_ndims = 1, i.e. the last axis is the feature axis):

  ExponentiatedQuadratic(amplitude s, length_scale l):
      k(x, y) = s^2 exp(-||x - y||^2 / (2 l^2));   amplitude=None  ->  s = 1
  ExpSinSquared(amplitude s, length_scale l, period T):
      k(x, y) = s^2 exp(-2 sum_f sin^2(pi |x_f - y_f| / T) / l^2)
  Linear() with every parameter None:
      k(x, y) = sum_f x_f y_f

``matrix(x1 (..., e1, f), x2 (..., e2, f)) -> (..., e1, e2)``;
``apply(x1 (..., f), x2 (..., f)) -> (...)`` with broadcasting.

PARITY UNPINNED by the reference (no tests there); cross-checked against
scikit-learn's independent RBF / ExpSineSquared / DotProduct in tests/test_oracle.py.
"""
import math

import torch


class ExponentiatedQuadratic:
    def __init__(self, amplitude=None, length_scale=None):
        self.amplitude = amplitude
        self.length_scale = length_scale

    def _finish(self, sqdist):
        e = -0.5 * sqdist
        if self.length_scale is not None:
            e = e / (self.length_scale ** 2)
        k = torch.exp(e)
        if self.amplitude is not None:
            k = k * self.amplitude ** 2
        return k

    def apply(self, x1, x2):
        return self._finish(((x1 - x2) ** 2).sum(-1))

    def matrix(self, x1, x2):
        d = x1.unsqueeze(-2) - x2.unsqueeze(-3)          # (..., e1, e2, f)
        return self._finish((d ** 2).sum(-1))


class ExpSinSquared:
    def __init__(self, amplitude=None, length_scale=None, period=None):
        self.amplitude = amplitude
        self.length_scale = length_scale
        self.period = period

    def _finish(self, absdiff):
        arg = math.pi * absdiff
        if self.period is not None:
            arg = arg / self.period
        e = -2.0 * (torch.sin(arg) ** 2).sum(-1)
        if self.length_scale is not None:
            e = e / (self.length_scale ** 2)
        k = torch.exp(e)
        if self.amplitude is not None:
            k = k * self.amplitude ** 2
        return k

    def apply(self, x1, x2):
        return self._finish((x1 - x2).abs())

    def matrix(self, x1, x2):
        return self._finish((x1.unsqueeze(-2) - x2.unsqueeze(-3)).abs())


class Linear:
    """tfk.Linear() with bias_variance = slope_variance = shift = None."""

    def apply(self, x1, x2):
        return (x1 * x2).sum(-1)

    def matrix(self, x1, x2):
        return (x1.unsqueeze(-2) * x2.unsqueeze(-3)).sum(-1)

"""gpflow.likelihoods.Gaussian (the only likelihood on the reference path: GPR's default,
``noise_variance=`` at Multi-Input_GPR/main.py:422, ``Gaussian(variance=1e-4)`` at
test_scripts/SVGP.py:517)."""
from __future__ import annotations

from .base import Module, Parameter, positive

DEFAULT_VARIANCE_LOWER_BOUND = 1e-6


class Likelihood(Module):
    pass


class Gaussian(Likelihood):
    def __init__(self, variance=None, *, scale=None, variance_lower_bound: float = DEFAULT_VARIANCE_LOWER_BOUND):
        if scale is not None:
            raise NotImplementedError("Gaussian(scale=...) is not used on the reference path")
        if variance is None:
            variance = 1.0
        if float(variance) <= variance_lower_bound:
            raise ValueError(f"The variance of the Gaussian likelihood must be strictly greater than {variance_lower_bound}")
        self.variance_lower_bound = variance_lower_bound
        self.variance = Parameter(variance, transform=positive(lower=variance_lower_bound), name="variance")

    def _children(self):
        for key, val in super()._children():
            if key != "variance_lower_bound":
                yield key, val

    def predict_mean_and_var(self, X, Fmu, Fvar):
        """gpflow Gaussian._predict_mean_and_var: (Fmu, Fvar + variance)."""
        return Fmu, Fvar + float(self.variance.numpy())

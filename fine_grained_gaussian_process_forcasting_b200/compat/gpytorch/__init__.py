"""Minimal ``gpytorch`` import shim: exactly the names the reference's GP blur path imports, backed by
fine_grained_gaussian_process_forcasting_b200.gpcompat (CUDA kernels, no real gpytorch)."""
from fine_grained_gaussian_process_forcasting_b200 import gpcompat as _gp

from . import distributions, kernels, likelihoods, means, mlls, models, settings, variational  # noqa: F401

__version__ = "0.0+gpblur"
IS_GPBLUR_SHIM = True

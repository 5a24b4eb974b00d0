"""B200-native GP blur / corruption hot path of SepKfr/Fine_grained_Gaussian_Process_Forcasting.

Public surface (mirrors the reference's modules for this path):
  DeepGP.DeepGPp, DeepGP.ToyDeepGPHiddenLayer, DeepGP.DeepGP2, GPModel.ExactGPModel,
  gpcompat.{VariationalELBO, DeepApproximateMLL, GaussianLikelihood, num_likelihood_samples, ...},
  ops.* (torch-facing wrappers of the C ABI in include/gpblur.h), distributed.* (batch-sharded training).
Importing the package does not need a GPU; calling any op does (no CPU fallback).
"""
from . import _cabi  # noqa: F401

__all__ = ["DeepGP", "GPModel", "gpcompat", "ops", "distributed", "build"]
__version__ = "0.1.0"


def install_gpytorch_shim():
    """Put the minimal ``gpytorch`` import shim (compat/gpytorch) on sys.path so that the reference's
    unchanged callers (``import gpytorch`` in forecast_denoising.py:5, train.py:1, denoise_model_2.py:1)
    resolve to this package.  No-op if a real gpytorch is already imported."""
    import sys
    from pathlib import Path
    if "gpytorch" in sys.modules:
        return sys.modules["gpytorch"]
    shim = str(Path(__file__).resolve().parent / "compat")
    if shim not in sys.path:
        sys.path.insert(0, shim)
    import gpytorch  # noqa: F401
    return gpytorch

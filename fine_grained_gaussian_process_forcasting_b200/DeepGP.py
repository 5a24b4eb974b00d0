"""GP blur / corruption model with the reference's module API, on the gpblur CUDA kernels.

Drop-in for /root/reference/denoising_model/DeepGP.py: same class names, constructor arguments,
``forward`` / ``predict`` contract, parameter names and RNG side effects -

* ``ToyDeepGPHiddenLayer(input_dims, output_dims, seed, num_inducing=256, mean_type='constant')``
  (DeepGP.py:14-73): seeds numpy / random / torch, then draws the inducing points BEFORE the mean
  parameters, so identically seeded models hold identical parameters.
* ``DeepGPp(num_hidden_dims, seed)`` (DeepGP.py:76-99): one hidden layer (``output_dims=None``, linear
  mean) + ``GaussianLikelihood``; ``predict(x[B, L, D]) -> (mean[S, B, L], dist)``.

Superset asked for by the north star: ``dist`` also carries ``variance``, the fused Philox
reparameterised sample (``dist.sample_value`` / ``dist.rsample()``) and ``dist.kl``; ``blur(x, y)``
returns everything (mean, variance, sample, ELBO) in one call.  ``DeepGP2`` is the two-layer stack.
"""
from __future__ import annotations

import random
from typing import NamedTuple, Optional

import numpy as np
import torch

from . import gpcompat as gp
from . import ops


class ToyDeepGPHiddenLayer(gp.DeepGPLayer):
    def __init__(self, input_dims, output_dims, seed, num_inducing=256, mean_type='constant'):
        np.random.seed(seed)
        random.seed(seed)
        torch.manual_seed(seed)

        batch_shape = torch.Size([]) if output_dims is None else torch.Size([output_dims])
        inducing_points = torch.randn(*batch_shape, num_inducing, input_dims)
        q_u = gp.MeanFieldVariationalDistribution(num_inducing_points=num_inducing, batch_shape=batch_shape)
        strategy = gp.VariationalStrategy(self, inducing_points, q_u, learn_inducing_locations=True)
        super().__init__(strategy, input_dims, output_dims)

        if mean_type == 'constant':
            self.mean_module = gp.ConstantMean(batch_shape=batch_shape)
        else:
            self.mean_module = gp.LinearMean(input_dims)
        self.covar_module = gp.ScaleKernel(gp.RBFKernel(batch_shape=batch_shape, ard_num_dims=input_dims),
                                           batch_shape=batch_shape, ard_num_dims=None)
        self.set_rng(seed)

    def forward(self, x):
        """Prior N(mean(x), K(x, x)) (DeepGP.py:51-54); dense covariance built lazily by the CUDA kernel."""
        return gp.MultivariateNormal(self.mean_module(x), lambda: self.covar_module(x))

    def __call__(self, x, *other_inputs, **kwargs):
        # concatenation skip connections (DeepGP.py:62-71)
        if len(other_inputs):
            if isinstance(x, gp.MultitaskMultivariateNormal):
                x = x.rsample()
            extra = [inp.unsqueeze(0).expand(gp.num_likelihood_samples.value(), *inp.shape) for inp in other_inputs]
            x = torch.cat([x] + extra, dim=-1)
        return super().__call__(x, are_samples=bool(len(other_inputs)))


class BlurOutput(NamedTuple):
    mean: torch.Tensor                 # [S, B, L] predictive (blur) mean
    variance: torch.Tensor             # [S, B, L]
    sample: Optional[torch.Tensor]     # [S, B, L] mean + sqrt(var) * eps
    elbo: Optional[torch.Tensor]       # [S, B] per-window ELBO (None without targets)
    kl: torch.Tensor                   # [] KL(q(u) || p(u))
    dist: gp.MultivariateNormal


class DeepGPp(gp.DeepGP):
    def __init__(self, num_hidden_dims, seed, *, num_inducing=256):
        hidden_layer = ToyDeepGPHiddenLayer(input_dims=num_hidden_dims, output_dims=None, mean_type='linear',
                                            seed=seed, num_inducing=num_inducing)
        super().__init__()
        self.hidden_layer = hidden_layer
        self.likelihood = gp.GaussianLikelihood()

    def forward(self, inputs):
        return self.hidden_layer(inputs)

    def predict(self, x):
        dist = self(x)
        preds = self.likelihood(dist)
        return preds.mean, dist

    def blur(self, x, y=None, num_data=None) -> BlurOutput:
        """One call for the whole hot path: predictive mean / variance, fused reparameterised sample and,
        when targets ``y [B, L]`` (or ``[1, B, L]``) are given, the per-window ELBO with
        ``num_data`` defaulting to the input width (forecast_denoising.py:88 passes d_model)."""
        return _blur(self, x, y, num_data)

    def blur_segments(self, x_flat, seg_shapes, y_last=None, num_data=None):
        """The blur of SEVERAL activations of one step in one fused evaluation: ``x_flat [N, D]`` holds their points
        back to back (e.g. the encoder-side [B, 192, D] then the decoder-side [B, 24, D] activations of
        denoise_model_2.py:50-51), ``seg_shapes`` their output shapes ((B, 192), (B, 24)).  Returns one ``BlurOutput``
        per segment - the same values as separate ``blur`` calls (same Philox counters) from one launch per kernel
        instead of one per activation; ``y_last`` gives the ELBO of the last segment."""
        dists = self.hidden_layer.call_segments(x_flat, seg_shapes)
        outs = []
        for i, dist in enumerate(dists):
            elbo = None
            if y_last is not None and i == len(dists) - 1:
                nd = float(num_data if num_data is not None else x_flat.shape[-1])
                tgt = y_last if y_last.dim() == dist.mean.dim() else y_last.unsqueeze(0)
                kl = self.variational_strategy.kl_divergence()
                elbo = ops.variational_elbo(dist.mean, dist.variance, tgt.expand(dist.mean.shape),
                                            self.likelihood.raw_noise, kl, nd)
            outs.append(BlurOutput(dist.mean, dist.variance, dist.sample_value, elbo, dist.kl, dist))
        return outs


def _blur(model, x, y=None, num_data=None) -> BlurOutput:
    dist = model(x)
    elbo = None
    if y is not None:
        nd = float(num_data if num_data is not None else x.shape[-1])
        tgt = y if y.dim() == dist.mean.dim() else y.unsqueeze(0)
        kl = model.variational_strategy.kl_divergence()      # summed over every GP of the stack
        elbo = ops.variational_elbo(dist.mean, dist.variance, tgt.expand(dist.mean.shape),
                                    model.likelihood.raw_noise, kl, nd)
    else:
        kl = dist.kl
    return BlurOutput(dist.mean, dist.variance, dist.sample_value, elbo, kl, dist)


class DeepGP2(gp.DeepGP):
    """Two-layer deep GP blur model: ``D -> hidden_dims`` independent whitened SVGPs,
    elementwise reparameterised sample, ``hidden_dims -> 1`` SVGP.  Built from the same
    ``ToyDeepGPHiddenLayer`` the reference defines (DeepGP.py:21-26, 42-49 already support
    ``output_dims=H``); layer outputs are chained on device without materialising any covariance."""

    def __init__(self, num_hidden_dims, seed, *, hidden_dims=10, num_inducing=256, skip_connection=False):
        layer1 = ToyDeepGPHiddenLayer(input_dims=num_hidden_dims, output_dims=hidden_dims, mean_type='linear',
                                      seed=seed, num_inducing=num_inducing)
        in2 = hidden_dims + (num_hidden_dims if skip_connection else 0)
        layer2 = ToyDeepGPHiddenLayer(input_dims=in2, output_dims=None, mean_type='linear', seed=seed + 1,
                                      num_inducing=num_inducing)
        super().__init__()
        self.hidden_layer = layer1
        self.last_layer = layer2
        self.skip_connection = skip_connection
        self.likelihood = gp.GaussianLikelihood()
        layer1.set_rng(seed, 0, stream=1)
        layer2.set_rng(seed, 0, stream=2)

    def forward(self, inputs):
        h = self.hidden_layer(inputs)
        if self.skip_connection:
            return self.last_layer(h, inputs)
        return self.last_layer(h)

    def predict(self, x):
        dist = self(x)
        preds = self.likelihood(dist)
        return preds.mean, dist

    def blur(self, x, y=None, num_data=None) -> BlurOutput:
        """Same contract as ``DeepGPp.blur``; ``kl`` is the sum over both layers (all H + 1 GPs)."""
        return _blur(self, x, y, num_data)

"""The two steps on either side of the GP blur in one training step, as fused CUDA
ops behind the reference's own call shapes:

* ``add_gp_noise(deep_gp, proj_up, x)`` - ``denoise_model_2.add_gp_noise``
  (/root/reference/denoising_model/denoise_model_2.py:32-40): ``x + proj_up(mean.permute(1, 2, 0))`` with
  ``proj_up = nn.Linear(1, d)`` as ONE pass over x (the reference's line 21 leaves ``proj_up`` commented out and line 37
  still calls it; the evident intent - a learned [1 -> d] lift of the blur mean - is what is built here);
* ``forecast_loss(final_projection, h, y_true, elbo, lam)`` - ``final_projection`` of the denoised decoder states, the
  MSE against ``y_true`` and ``loss = mse + clip(lam, 0, 0.005) * mll_error`` with ``mll_error = -mean(elbo)``
  (/root/reference/forecast_denoising.py:84, 87-89, 102-104) as ONE pass forward and ONE backward.

No CPU fallback: the ops raise on non-CUDA tensors (the CPU restatement lives in oracle/gp_oracle.py)."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
from torch import Tensor

from . import _cabi
from .ops import _f32c, _need_cuda, _next_ticket, _ptr, _stream


def _scratch(dev, D):
    return torch.empty(_cabi.lib().gpblur_step_scratch_floats(int(D)), device=dev, dtype=torch.float32)


class _BlurApplyFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mean, weight, bias):
        _need_cuda(x, mean, weight, bias)
        D = x.shape[-1]
        x2 = _f32c(x).reshape(-1, D)
        N = x2.shape[0]
        m1 = _f32c(mean).reshape(-1)
        if m1.numel() != N:
            raise ValueError(f"blur_apply: mean has {m1.numel()} elements for {N} rows of x")
        w1, b1 = _f32c(weight).reshape(-1), _f32c(bias).reshape(-1)
        if w1.numel() != D or b1.numel() != D:
            raise ValueError("blur_apply: proj_up must be nn.Linear(1, D)")
        out = torch.empty_like(x2)
        with torch.cuda.device(x2.device):
            rc = _cabi.lib().gpblur_blur_apply_forward(_ptr(x2), _ptr(m1), _ptr(w1), _ptr(b1), N, D, _ptr(out), _stream())
        _cabi.check(rc, "gpblur_blur_apply_forward")
        ctx.save_for_backward(m1, w1)
        ctx.shapes = (x.shape, mean.shape, weight.shape, bias.shape)
        return out.reshape(x.shape)

    @staticmethod
    def backward(ctx, g_out):
        m1, w1 = ctx.saved_tensors
        xs, ms, ws, bs = ctx.shapes
        D = xs[-1]
        g2 = _f32c(g_out).reshape(-1, D)
        N = g2.shape[0]
        dev = g2.device
        need = ctx.needs_input_grad
        g_mean = torch.empty(N, device=dev, dtype=torch.float32)
        g_w = torch.empty(D, device=dev, dtype=torch.float32)
        g_b = torch.empty(D, device=dev, dtype=torch.float32)
        if N > 0 and (need[1] or need[2] or need[3]):
            scratch, ticket = _scratch(dev, D), _next_ticket(dev)      # (named: alive until the launch is issued)
            with torch.cuda.device(dev):
                rc = _cabi.lib().gpblur_blur_apply_backward(_ptr(g2), _ptr(m1), _ptr(w1), N, D, _ptr(g_mean), _ptr(g_w),
                                                            _ptr(g_b), _ptr(scratch), ticket, _stream())
            _cabi.check(rc, "gpblur_blur_apply_backward")
        return (g_out if need[0] else None, g_mean.reshape(ms) if need[1] else None,
                g_w.reshape(ws) if need[2] else None, g_b.reshape(bs) if need[3] else None)


def blur_apply(x: Tensor, mean: Tensor, weight: Tensor, bias: Tensor) -> Tensor:
    """x [..., D] + mean [...] * weight [D(, 1)] + bias [D]: ``x + nn.Linear(1, D)(mean.unsqueeze(-1))``."""
    return _BlurApplyFunction.apply(x, mean, weight, bias)


def add_gp_noise(deep_gp, proj_up: torch.nn.Linear, x: Tensor):
    """``denoise_model_2.add_gp_noise`` (denoise_model_2.py:32-40): -> (x_noisy [B, L, D], dist).  Needs
    ``num_likelihood_samples(1)`` like the reference (``eps_gp.permute(1, 2, 0)`` feeds a Linear with in_features 1)."""
    eps_gp, dist = deep_gp.predict(x)
    if eps_gp.shape[0] != 1:
        raise RuntimeError("add_gp_noise: run under num_likelihood_samples(1) (train.py:20)")
    return blur_apply(x, eps_gp[0], proj_up.weight, proj_up.bias), dist


class _ForecastLossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, weight, bias, y, elbo, lam):
        _need_cuda(h, weight, bias)
        Bp, P, D = h.shape
        hv = h if (h.dtype == torch.float32 and h.stride(2) == 1 and h.stride(1) == D and h.stride(0) >= P * D) \
            else _f32c(h)                                    # a [:, -P:, :] slice is read in place
        N = Bp * P
        dev = h.device
        w1, b1 = _f32c(weight).reshape(-1), _f32c(bias).reshape(-1)
        y1 = None if y is None else _f32c(y).reshape(-1)
        e1 = None if elbo is None else _f32c(elbo).reshape(-1)
        l1 = None if lam is None else _f32c(lam).reshape(-1)
        if y1 is not None and y1.numel() != N:
            raise ValueError("forecast_loss: y_true does not match the rows of h")
        final = torch.empty(N, device=dev, dtype=torch.float32)
        scalars = torch.empty(3, device=dev, dtype=torch.float32)
        B = 0 if e1 is None else e1.numel()
        scratch, ticket = _scratch(dev, D), _next_ticket(dev)
        with torch.cuda.device(dev):
            rc = _cabi.lib().gpblur_loss_forward(_ptr(hv), hv.stride(0), P, _ptr(w1), _ptr(b1), _ptr(y1), _ptr(e1), B,
                                                 _ptr(l1), N, D, _ptr(final), _ptr(scalars), _ptr(scratch), ticket,
                                                 _stream())
        _cabi.check(rc, "gpblur_loss_forward")
        ctx.save_for_backward(hv, w1, final, scalars, *(t for t in (y1, l1) if t is not None))
        ctx.has = (y1 is not None, l1 is not None, B)
        ctx.shapes = (h.shape, weight.shape, bias.shape, None if elbo is None else elbo.shape,
                      None if lam is None else lam.shape)
        ctx.set_materialize_grads(False)
        return final.reshape(Bp, P, 1), scalars[0], scalars[1]

    @staticmethod
    def backward(ctx, g_final, g_loss, g_mse):
        saved = list(ctx.saved_tensors)
        hv, w1, final, scalars = saved[:4]
        rest = saved[4:]
        has_y, has_lam, B = ctx.has
        y1 = rest.pop(0) if has_y else None
        l1 = rest.pop(0) if has_lam else None
        hs, ws, bs, es, ls = ctx.shapes
        Bp, P, D = hs
        N = Bp * P
        dev = hv.device
        need = ctx.needs_input_grad
        gf = None if g_final is None else _f32c(g_final).reshape(-1)
        gl = None if g_loss is None else _f32c(g_loss).reshape(1)
        gm = None if g_mse is None else _f32c(g_mse).reshape(1)
        g_h = torch.empty(N, D, device=dev, dtype=torch.float32) if need[0] else None
        g_w = torch.empty(D, device=dev, dtype=torch.float32)
        g_b = torch.empty(1, device=dev, dtype=torch.float32)
        g_e = torch.empty(B, device=dev, dtype=torch.float32) if (B and need[4]) else None
        g_l = torch.empty(1, device=dev, dtype=torch.float32) if (has_lam and need[5]) else None
        scratch, ticket = _scratch(dev, D), _next_ticket(dev)
        with torch.cuda.device(dev):
            rc = _cabi.lib().gpblur_loss_backward(_ptr(hv), hv.stride(0), P, _ptr(w1), _ptr(y1), _ptr(final), _ptr(scalars),
                                                  _ptr(l1), _ptr(gf), _ptr(gl), _ptr(gm), B, N, D, _ptr(g_h), _ptr(g_w),
                                                  _ptr(g_b), _ptr(g_e), _ptr(g_l), _ptr(scratch), ticket, _stream())
        _cabi.check(rc, "gpblur_loss_backward")
        return (g_h.reshape(hs) if g_h is not None else None, g_w.reshape(ws) if need[1] else None,
                g_b.reshape(bs) if need[2] else None, None,
                g_e.reshape(es) if g_e is not None else None, g_l.reshape(ls) if g_l is not None else None)


def forecast_loss(final_projection: torch.nn.Linear, h: Tensor, y_true: Optional[Tensor] = None,
                  elbo: Optional[Tensor] = None, lam: Optional[Tensor] = None):
    """h [B, P, D] (may be the ``[:, -pred_len:, :]`` slice of the decoder states) -> (final_outputs [B, P, 1], loss,
    mse_loss) as forecast_denoising.py:84, 102-104 computes them; ``elbo`` [B] (or [1, B]) is the per-window ELBO the GP
    path returns (``mll_error = -elbo.mean()``, :87-89), ``lam`` the model's scalar parameter (:31)."""
    return _ForecastLossFunction.apply(h, final_projection.weight, final_projection.bias, y_true, elbo, lam)

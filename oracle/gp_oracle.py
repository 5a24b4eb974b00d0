"""CPU oracle for the GP blur / corruption hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this module.  The product path (the CUDA extension behind
``fine_grained_gaussian_process_forcasting_b200``) never routes through it.

PARITY UNPINNED: the arithmetic of this path lives in the third-party dependency ``gpytorch``
(pin: ``gpytorch >= 1.9.0``, /root/reference/README.md:15) and its dependency ``linear_operator``;
neither is vendored under /root/reference nor installable here (no wheel in /opt/wheelhouse, no
network), and the reference ships no tests, fixtures or golden vectors for this path.  This file
therefore *restates the published gpytorch algorithm* at the reference's own call sites:

* model structure / init order ......... /root/reference/denoising_model/DeepGP.py:14-49, 76-88
* prior mean + kernel ................... /root/reference/denoising_model/DeepGP.py:51-54
* layer call / predict .................. /root/reference/denoising_model/DeepGP.py:56-73, 90-99
* ELBO wiring (num_data = d_model) ...... /root/reference/forecast_denoising.py:86-89
* output layout [S=1, B, L] ............. /root/reference/denoising_model/denoise_model_2.py:36-37,
                                          /root/reference/train.py:20
* exact GP model ........................ /root/reference/denoising_model/GPModel.py:4-13

and pins itself with closed-form known-answer tests (tests/test_oracle.py) and fp64 gradcheck.  One EXTERNAL anchor
exists: tests/test_pin_sklearn.py checks both predictive restatements (and the exact GP) against scikit-learn's
independent GaussianProcessRegressor - with m = L^-1 y_z, s = 0 the whitened SVGP predictive is an exact GP posterior -
which fixes the ARD kernel, the outputscale, where the 1e-4 jitter goes, the whitening and the variance formula; the
gpytorch-only conventions (variance clamp, ELBO scaling, first-call initialisation) remain restated from the source.

gpytorch semantics restated (gpytorch >= 1.9, ``VariationalStrategy.forward`` whitened form):
  Kzz = os * exp(-1/2 |(z - z')/l|^2) + 1e-4 I         (settings.variational_cholesky_jitter, fp32)
  L   = chol(Kzz)  in float64                            (settings._linalg_dtype_cholesky)
  A   = L^-1 Kzx   in float64, cast back to float32     ("interp_term")
  mu  = A^T m + x @ w + b
  var = diag(Kxx) + 1e-4 + sum_m (s_m^2 - 1) A_m^2 ,  clamped at 1e-6   (MultivariateNormal.variance)
  diag(Kxx) = os exactly (Kernel.covar_dist(diag=True) returns zeros when x1 is x2)
  ELBO_b = mean_l E_q[log N(y | f, noise)] - KL / num_data ;  noise = softplus(raw_noise) + 1e-4
  KL( N(m, diag s^2) || N(0, I) ) = 1/2 [ sum s^2 + sum m^2 - M - sum log s^2 ]
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np
import torch

JITTER = 1e-4          # gpytorch.settings.variational_cholesky_jitter for float32
MIN_VARIANCE = 1e-6    # gpytorch.settings.min_variance for float32
NOISE_LOWER = 1e-4     # GaussianLikelihood noise constraint GreaterThan(1e-4)
LOG_2PI = math.log(2.0 * math.pi)

PARAM_NAMES = (
    "inducing_points",      # [M, D]
    "raw_lengthscale",      # [1, D]
    "raw_outputscale",      # []
    "variational_mean",     # [M]
    "variational_stddev",   # [M]
    "weights",              # [D, 1]
    "bias",                 # [1]
    "raw_noise",            # [1]
)


def softplus(x: torch.Tensor) -> torch.Tensor:
    return torch.nn.functional.softplus(x)


# ----------------------------------------------------------------------------------------------
# parameter construction
# ----------------------------------------------------------------------------------------------
def init_params_reference(D: int, seed: int, M: int = 256, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Parameters exactly as ``DeepGPp(D, seed)`` creates them (DeepGP.py:17-49, 77-88).

    RNG order after ``torch.manual_seed(seed)``: inducing_points = randn(M, D) first, then the
    LinearMean ``weights = randn(D, 1)`` and ``bias = randn(1)``; all other parameters are constants
    (variational mean 0, stddev 1, raw lengthscale / outputscale / noise 0).
    """
    import random

    np.random.seed(seed)
    random.seed(seed)
    torch.manual_seed(seed)
    Z = torch.randn(M, D)
    w = torch.randn(D, 1)
    b = torch.randn(1)
    p = {
        "inducing_points": Z,
        "raw_lengthscale": torch.zeros(1, D),
        "raw_outputscale": torch.zeros(()),
        "variational_mean": torch.zeros(M),
        "variational_stddev": torch.ones(M),
        "weights": w,
        "bias": b,
        "raw_noise": torch.zeros(1),
    }
    return {k: v.to(dtype) for k, v in p.items()}


def init_params_exercise(D: int, M: int, seed: int, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """The "R-exercise" parameter regime of SURVEY.md section 8(d): lengthscales ~ sqrt(D) so that
    K(x, Z) is O(0.1..1) instead of underflowing, non-trivial q(u)."""
    g = torch.Generator().manual_seed(seed)
    Z = torch.randn(M, D, generator=g)
    ell = math.sqrt(D) * (0.75 + 0.5 * torch.rand(1, D, generator=g))
    raw_ell = torch.log(torch.expm1(ell))            # inverse softplus
    m = 0.5 * torch.randn(M, generator=g)
    s = 0.5 + torch.rand(M, generator=g)
    w = torch.randn(D, 1, generator=g) / math.sqrt(D)
    b = torch.randn(1, generator=g)
    p = {
        "inducing_points": Z,
        "raw_lengthscale": raw_ell,
        "raw_outputscale": torch.zeros(()),
        "variational_mean": m,
        "variational_stddev": s,
        "weights": w,
        "bias": b,
        "raw_noise": torch.zeros(1),
    }
    return {k: v.to(dtype) for k, v in p.items()}


def clone_params(p, dtype=None, requires_grad=False):
    out = {}
    for k, v in p.items():
        t = v.detach().clone()
        if dtype is not None:
            t = t.to(dtype)
        t.requires_grad_(requires_grad)
        out[k] = t
    return out


# ----------------------------------------------------------------------------------------------
# kernel evaluation, reference order (gpytorch Kernel.covar_dist / sq_dist)
# ----------------------------------------------------------------------------------------------
def sq_dist_reference_order(x1: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
    """gpytorch ``sq_dist``: mean-centre on x1, norm expansion as ONE matmul with inner dim D+2,
    clamp_min(0).  x1 [..., n, D], x2 [..., m, D] (already divided by the lengthscale)."""
    adjustment = x1.mean(-2, keepdim=True)
    x1 = x1 - adjustment
    x2 = x2 - adjustment
    x1_norm = x1.pow(2).sum(dim=-1, keepdim=True)
    x1_pad = torch.ones_like(x1_norm)
    x2_norm = x2.pow(2).sum(dim=-1, keepdim=True)
    x2_pad = torch.ones_like(x2_norm)
    x1_ = torch.cat([-2.0 * x1, x1_norm, x1_pad], dim=-1)
    x2_ = torch.cat([x2, x2_pad, x2_norm], dim=-1)
    res = x1_.matmul(x2_.transpose(-2, -1))
    return res.clamp_min(0)


def rbf_scale_reference_order(x1, x2, ell, os):
    """ScaleKernel(RBFKernel(ard)) dense evaluation (DeepGP.py:46-49)."""
    d2 = sq_dist_reference_order(x1 / ell, x2 / ell)
    return os * torch.exp(-0.5 * d2)


def rbf_scale_direct(x1, x2, ell, os):
    """Same kernel by direct differences (no cancellation) - the mathematically exact form."""
    diff = (x1.unsqueeze(-2) - x2.unsqueeze(-3)) / ell
    return os * torch.exp(-0.5 * diff.pow(2).sum(-1))


# ----------------------------------------------------------------------------------------------
# forward, reference order (what gpytorch launches: batch-expanded Z, fp32 kernels, fp64 linalg)
# ----------------------------------------------------------------------------------------------
def svgp_predict_reference_order(p: Dict[str, torch.Tensor], x: torch.Tensor,
                                 linalg_dtype=torch.float64) -> Tuple[torch.Tensor, torch.Tensor]:
    """``ToyDeepGPHiddenLayer.__call__`` -> whitened ``VariationalStrategy.forward`` for
    output_dims=None, linear mean (DeepGP.py:56-73).  x [B, L, D] -> mean, var [B, L].

    Op order follows gpytorch: inducing points expanded over the batch, ONE kernel evaluation on
    cat([Z, x]) per batch element, jitter on Kzz, Cholesky + solve in ``linalg_dtype``, cast back.
    Differentiable w.r.t. x and every entry of ``p`` by torch autograd.
    """
    Z = p["inducing_points"]
    B, L, D = x.shape
    M = Z.shape[0]
    ell = softplus(p["raw_lengthscale"])               # [1, D]
    os = softplus(p["raw_outputscale"])
    Zb = Z.unsqueeze(0).expand(B, M, D)
    full = torch.cat([Zb, x], dim=-2)                  # [B, M+L, D]
    full_mean = (full @ p["weights"]).squeeze(-1) + p["bias"]      # LinearMean
    full_covar = rbf_scale_reference_order(full, full, ell, os)    # [B, M+L, M+L]
    test_mean = full_mean[..., M:]
    eye = torch.eye(M, dtype=x.dtype)
    Kzz = full_covar[..., :M, :M] + JITTER * eye
    Kzx = full_covar[..., :M, M:]
    Lc = torch.linalg.cholesky(Kzz.to(linalg_dtype))
    A = torch.linalg.solve_triangular(Lc, Kzx.to(linalg_dtype), upper=False).to(x.dtype)   # [B, M, L]
    m = p["variational_mean"]
    s = p["variational_stddev"]
    mean = (A.transpose(-1, -2) @ m.unsqueeze(-1)).squeeze(-1) + test_mean
    # diag(Kxx): gpytorch evaluates the diagonal with diag=True on identical inputs -> d2 = 0 exactly
    kxx_diag = os * torch.ones(B, L, dtype=x.dtype)
    var = kxx_diag + JITTER + ((s.pow(2) - 1.0).unsqueeze(-1) * A.pow(2)).sum(-2)
    var = var.clamp_min(MIN_VARIANCE)
    return mean, var


def svgp_predict_closed_form(p: Dict[str, torch.Tensor], x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Same predictive in the shared-Z closed form of SURVEY Appendix A (ONE MxM Cholesky, direct
    squared distances).  Intended for float64 "truth" runs."""
    Z = p["inducing_points"]
    B, L, D = x.shape
    M = Z.shape[0]
    ell = softplus(p["raw_lengthscale"]).reshape(D)
    os = softplus(p["raw_outputscale"])
    Kzz = rbf_scale_direct(Z, Z, ell, os) + JITTER * torch.eye(M, dtype=x.dtype)
    Lc = torch.linalg.cholesky(Kzz)
    X = x.reshape(B * L, D)
    Kzx = rbf_scale_direct(Z, X, ell, os)              # [M, N]
    A = torch.linalg.solve_triangular(Lc, Kzx, upper=False)
    m = p["variational_mean"]
    s = p["variational_stddev"]
    mean = A.t() @ m + (X @ p["weights"]).squeeze(-1) + p["bias"]
    var = os + JITTER + ((s.pow(2) - 1.0).unsqueeze(-1) * A.pow(2)).sum(0)
    var = var.clamp_min(MIN_VARIANCE)
    return mean.reshape(B, L), var.reshape(B, L)


# ----------------------------------------------------------------------------------------------
# ELBO (forecast_denoising.py:86-89 -> DeepApproximateMLL(VariationalELBO(...)))
# ----------------------------------------------------------------------------------------------
def kl_meanfield(p) -> torch.Tensor:
    m = p["variational_mean"]
    s2 = p["variational_stddev"].pow(2)
    return 0.5 * (s2.sum() + m.pow(2).sum() - m.numel() - torch.log(s2).sum())


def noise_variance(p) -> torch.Tensor:
    return softplus(p["raw_noise"]).reshape(()) + NOISE_LOWER


def elbo_per_window(mean, var, y, noise, kl, num_data) -> torch.Tensor:
    """VariationalELBO.forward for one sample (num_likelihood_samples=1): [B, L] -> [B]."""
    L = mean.shape[-1]
    ll = -0.5 * (((y - mean).pow(2) + var) / noise + torch.log(noise) + LOG_2PI)
    return ll.sum(-1) / L - kl / num_data


def mll_error(p, x, y, num_data, reference_order=True) -> torch.Tensor:
    """``-mll(dist, y_true.permute(2,0,1)).mean()`` of forecast_denoising.py:89."""
    fn = svgp_predict_reference_order if reference_order else svgp_predict_closed_form
    mean, var = fn(p, x)
    e = elbo_per_window(mean, var, y, noise_variance(p), kl_meanfield(p), num_data)
    return -e.mean()


# ----------------------------------------------------------------------------------------------
# Philox4x32-10 + Box-Muller (counter layout shared with the CUDA sampler)
# ----------------------------------------------------------------------------------------------
PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = np.uint32(0x9E3779B9)
PHILOX_W1 = np.uint32(0xBB67AE85)


def philox4x32_10(counter: np.ndarray, key: np.ndarray) -> np.ndarray:
    """Random123 Philox4x32 with 10 rounds.  counter [..., 4] uint32, key [..., 2] uint32."""
    c = np.array(counter, dtype=np.uint32, copy=True)
    k = np.array(np.broadcast_to(key, c.shape[:-1] + (2,)), dtype=np.uint32, copy=True)
    c0, c1, c2, c3 = (c[..., i].astype(np.uint64) for i in range(4))
    k0 = k[..., 0].copy()
    k1 = k[..., 1].copy()
    mask = np.uint64(0xFFFFFFFF)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = PHILOX_M0 * c0
            p1 = PHILOX_M1 * c2
            hi0, lo0 = p0 >> np.uint64(32), p0 & mask
            hi1, lo1 = p1 >> np.uint64(32), p1 & mask
            n0 = hi1 ^ c1 ^ k0.astype(np.uint64)
            n1 = lo1
            n2 = hi0 ^ c3 ^ k1.astype(np.uint64)
            n3 = lo0
            c0, c1, c2, c3 = n0 & mask, n1, n2 & mask, n3
            k0 = (k0 + PHILOX_W0).astype(np.uint32)
            k1 = (k1 + PHILOX_W1).astype(np.uint32)
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def philox_bits(seed: int, offset: int, n: int, stream: int = 0) -> np.ndarray:
    """Raw Philox words for elements offset .. offset+n-1.

    Counter layout (independent of launch geometry and GPU count):
      counter = (lo32(e), hi32(e), stream, 0),  e = offset + element index (uint64)
      key     = (lo32(seed), hi32(seed))
    Returns uint32 [n, 4]."""
    e = np.uint64(offset) + np.arange(n, dtype=np.uint64)
    ctr = np.zeros((n, 4), dtype=np.uint32)
    ctr[:, 0] = (e & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    ctr[:, 1] = (e >> np.uint64(32)).astype(np.uint32)
    ctr[:, 2] = np.uint32(stream & 0xFFFFFFFF)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    return philox4x32_10(ctr, key)


def philox_normal(seed: int, offset: int, n: int, stream: int = 0) -> np.ndarray:
    """Standard normals (float32) from words 0 and 1 of each Philox block via Box-Muller:
      u1 = ((r0 >> 9) + 0.5) * 2^-23   in (0, 1)   (exact in fp32)
      u2 = ((r1 >> 9) + 0.5) * 2^-23
      eps = sqrt(-2 ln u1) * cos(2 pi u2)."""
    r = philox_bits(seed, offset, n, stream)
    u1 = ((r[:, 0] >> np.uint32(9)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -23)
    u2 = ((r[:, 1] >> np.uint32(9)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -23)
    rad = np.sqrt(np.float32(-2.0) * np.log(u1)).astype(np.float32)
    # cos(2 pi u2) evaluated in float64 then rounded: matches CUDA cospif(2*u2) to < 1 ulp
    ang = np.cos(2.0 * np.pi * u2.astype(np.float64)).astype(np.float32)
    return (rad * ang).astype(np.float32)


def rsample(mean: torch.Tensor, var: torch.Tensor, eps: torch.Tensor) -> torch.Tensor:
    """DeepGP-layer sampling convention: Normal(mean, var.sqrt()).rsample() with explicit eps."""
    return mean + var.sqrt() * eps


# ----------------------------------------------------------------------------------------------
# kernel-order analytic backward - the algorithm the CUDA backward implements.
# Kept here so that tests can check the *math* of the hand-written backward against autograd on CPU.
# ----------------------------------------------------------------------------------------------
@dataclass
class SvgpGrads:
    dx: torch.Tensor
    inducing_points: torch.Tensor
    raw_lengthscale: torch.Tensor
    raw_outputscale: torch.Tensor
    variational_mean: torch.Tensor
    variational_stddev: torch.Tensor
    weights: torch.Tensor
    bias: torch.Tensor


def svgp_backward_kernel_order(p, x, g_mean, g_var, g_kl=0.0) -> SvgpGrads:
    """Analytic gradient of  sum(g_mean*mean) + sum(g_var*var) + g_kl*KL  in the op order of the
    CUDA backward: whitening through the explicit inverse Linv, Gram S = sum_n g_var a a^T,
    E = m u^T + 2 diag(c) S, Gbar = Linv^T E, Lbar = -tril(Gbar), Cholesky backward via
    Kbar = sym(Linv^T Phi(L^T Lbar) Linv).  All tensors float64."""
    Z = p["inducing_points"]
    B, L_, D = x.shape
    M = Z.shape[0]
    N = B * L_
    X = x.reshape(N, D)
    gm = g_mean.reshape(N)
    gv = g_var.reshape(N)
    ell = softplus(p["raw_lengthscale"]).reshape(D)
    os = softplus(p["raw_outputscale"])
    m = p["variational_mean"]
    s = p["variational_stddev"]
    w = p["weights"].reshape(D)
    cvec = s.pow(2) - 1.0

    Kzz0 = rbf_scale_direct(Z, Z, ell, os)
    Lc = torch.linalg.cholesky(Kzz0 + JITTER * torch.eye(M, dtype=X.dtype))
    Linv = torch.linalg.solve_triangular(Lc, torch.eye(M, dtype=X.dtype), upper=False)
    K = rbf_scale_direct(X, Z, ell, os)                # [N, M]
    A = K @ Linv.t()                                   # [N, M]
    var_raw = os + JITTER + (A.pow(2) * cvec).sum(-1)
    gv = torch.where(var_raw < MIN_VARIANCE, torch.zeros_like(gv), gv)   # clamp kills the gradient

    Abar = gm[:, None] * m[None, :] + 2.0 * gv[:, None] * cvec[None, :] * A
    Kbar = Abar @ Linv                                 # k_bar_n = Linv^T a_bar_n
    W = Kbar * K
    r = W.sum(1)
    csum = W.sum(0)
    Xt = X / ell
    Zt = Z / ell
    dx = (W @ Zt - r[:, None] * Xt) / ell + gm[:, None] * w[None, :]
    WX = W.t() @ Xt                                    # [M, D]
    dZ = (WX - csum[:, None] * Zt) / ell
    q = (r[:, None] * Xt.pow(2)).sum(0)
    dell = (q - 2.0 * (Zt * WX).sum(0) + (csum[:, None] * Zt.pow(2)).sum(0)) / ell
    dos = r.sum() / os + gv.sum()
    u = (gm[:, None] * A).sum(0)                       # sum_n g_mu a
    S = A.t() @ (gv[:, None] * A)                      # Gram
    dm = u + g_kl * m
    ds = 2.0 * s * torch.diagonal(S) + g_kl * (s - 1.0 / s)
    dw = (gm[:, None] * X).sum(0)
    db = gm.sum()

    # Cholesky backward
    E = m[:, None] * u[None, :] + 2.0 * cvec[:, None] * S
    Gbar = Linv.t() @ E
    Lbar = -torch.tril(Gbar)
    P = Lc.t() @ Lbar
    Phi = torch.tril(P)
    Phi = Phi - 0.5 * torch.diag(torch.diagonal(P))
    Kb = Linv.t() @ Phi @ Linv
    Kb = 0.5 * (Kb + Kb.t())
    Wzz = Kb * Kzz0
    rz = Wzz.sum(1)
    dZ = dZ + 2.0 * (Wzz @ Zt - rz[:, None] * Zt) / ell
    diff2 = (Zt.unsqueeze(1) - Zt.unsqueeze(0)).pow(2)   # [M, M, D]
    dell = dell + (Wzz.unsqueeze(-1) * diff2).sum((0, 1)) / ell
    dos = dos + Wzz.sum() / os

    sig = torch.sigmoid
    return SvgpGrads(
        dx=dx.reshape(B, L_, D),
        inducing_points=dZ,
        raw_lengthscale=(dell * sig(p["raw_lengthscale"]).reshape(D)).reshape(1, D),
        raw_outputscale=dos * sig(p["raw_outputscale"]),
        variational_mean=dm,
        variational_stddev=ds,
        weights=dw.reshape(D, 1),
        bias=db.reshape(1),
    )


# ----------------------------------------------------------------------------------------------
# two-layer DeepGP and exact GP (GPModel.py)
# ----------------------------------------------------------------------------------------------
def init_params_hidden_layer(D: int, H: int, M: int, seed: int, exercise: bool = True,
                             dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Layer-1 parameters for ``ToyDeepGPHiddenLayer(input_dims=D, output_dims=H,
    mean_type='linear')`` (DeepGP.py:24-26, 42-49): H independent GPs sharing one LinearMean."""
    g = torch.Generator().manual_seed(seed)
    Z = torch.randn(H, M, D, generator=g)
    if exercise:
        ell = math.sqrt(D) * (0.75 + 0.5 * torch.rand(H, 1, D, generator=g))
        raw_ell = torch.log(torch.expm1(ell))
        m = 0.5 * torch.randn(H, M, generator=g)
        s = 0.5 + torch.rand(H, M, generator=g)
        w = torch.randn(D, 1, generator=g) / math.sqrt(D)
    else:
        raw_ell = torch.zeros(H, 1, D)
        m = torch.zeros(H, M)
        s = torch.ones(H, M)
        w = torch.randn(D, 1, generator=g)
    b = torch.randn(1, generator=g)
    p = {
        "inducing_points": Z,
        "raw_lengthscale": raw_ell,
        "raw_outputscale": torch.zeros(H),
        "variational_mean": m,
        "variational_stddev": s,
        "weights": w,
        "bias": b,
    }
    return {k: v.to(dtype) for k, v in p.items()}


def hidden_layer_predict(p1, x, closed_form=True):
    """Layer with output_dims=H: DeepGPLayer expands the input over H and runs H independent
    whitened SVGPs.  Returns mean, var [B, L, H]."""
    H = p1["inducing_points"].shape[0]
    means, vars_ = [], []
    fn = svgp_predict_closed_form if closed_form else svgp_predict_reference_order
    for h in range(H):
        ph = {
            "inducing_points": p1["inducing_points"][h],
            "raw_lengthscale": p1["raw_lengthscale"][h],
            "raw_outputscale": p1["raw_outputscale"][h],
            "variational_mean": p1["variational_mean"][h],
            "variational_stddev": p1["variational_stddev"][h],
            "weights": p1["weights"],
            "bias": p1["bias"],
        }
        mu, v = fn(ph, x)
        means.append(mu)
        vars_.append(v)
    return torch.stack(means, -1), torch.stack(vars_, -1)


def kl_hidden_layer(p1) -> torch.Tensor:
    m = p1["variational_mean"]
    s2 = p1["variational_stddev"].pow(2)
    return 0.5 * (s2.sum() + m.pow(2).sum() - m.numel() - torch.log(s2).sum())


def deepgp2_predict(p1, p2, x, eps, closed_form=True):
    """Two-layer stack: layer 1 (D -> H), elementwise Normal(mean, sqrt(var)).rsample() with the
    given eps [B, L, H], layer 2 (H -> scalar).  Returns (mean, var, hidden_sample)."""
    mu1, v1 = hidden_layer_predict(p1, x, closed_form)
    h = rsample(mu1, v1, eps)
    fn = svgp_predict_closed_form if closed_form else svgp_predict_reference_order
    mu2, v2 = fn(p2, h)
    return mu2, v2, h


def exact_gp_prior(x, constant, raw_lengthscale, raw_outputscale):
    """``ExactGPModel.forward`` (GPModel.py:10-13): ConstantMean + ScaleKernel(RBF, no ARD).
    x [n, D] -> (mean [n], covar [n, n])."""
    ell = softplus(raw_lengthscale).reshape(())
    os = softplus(raw_outputscale).reshape(())
    mean = constant.reshape(()).expand(x.shape[-2])
    covar = rbf_scale_direct(x, x, ell, os)
    return mean, covar


def exact_gp_posterior(train_x, train_y, test_x, constant, raw_lengthscale, raw_outputscale, raw_noise):
    """gpytorch ``ExactGP.__call__`` in eval mode: condition on (train_x, train_y) through the
    Cholesky of K + noise I.  Returns (mean [n*], covar [n*, n*])."""
    ell = softplus(raw_lengthscale).reshape(())
    os = softplus(raw_outputscale).reshape(())
    noise = softplus(raw_noise).reshape(()) + NOISE_LOWER
    c = constant.reshape(())
    Ktt = rbf_scale_direct(train_x, train_x, ell, os) + noise * torch.eye(train_x.shape[0], dtype=train_x.dtype)
    Kst = rbf_scale_direct(test_x, train_x, ell, os)
    Kss = rbf_scale_direct(test_x, test_x, ell, os)
    Lc = torch.linalg.cholesky(Ktt)
    alpha = torch.cholesky_solve((train_y - c).unsqueeze(-1), Lc).squeeze(-1)
    V = torch.linalg.solve_triangular(Lc, Kst.t(), upper=False)
    return c + Kst @ alpha, Kss - V.t() @ V


# ----------------------------------------------------------------------------------------------
# synthetic inputs of SURVEY 8(d)
# ----------------------------------------------------------------------------------------------
def make_inputs(B: int, L: int, D: int, seed: int, dtype=torch.float32, layernorm=False):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, L, D, generator=g)
    if layernorm:
        x = torch.nn.functional.layer_norm(x, (D,))
    y = torch.randn(B, L, generator=g)
    g_mean = torch.randn(B, L, generator=g)
    g_var = torch.randn(B, L, generator=g)
    return x.to(dtype), y.to(dtype), g_mean.to(dtype), g_var.to(dtype)


# ---------------------------------------------------------------------------------------------------------------------
# The steps on either side of the GP blur - plain torch, the reference's own ops.
# ---------------------------------------------------------------------------------------------------------------------
def blur_apply_reference(x: torch.Tensor, eps_gp: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """denoise_model_2.add_gp_noise, /root/reference/denoising_model/denoise_model_2.py:36-38:
    ``eps_gp = self.proj_up(eps_gp.permute(1, 2, 0)); x_noisy = x + eps_gp`` with ``eps_gp`` [1, B, L] the blur mean
    and ``proj_up = nn.Linear(1, d)`` (weight [d, 1], bias [d])."""
    return x + torch.nn.functional.linear(eps_gp.permute(1, 2, 0), weight, bias)


def forecast_loss_reference(h, weight, bias, y_true, elbo, lam):
    """/root/reference/forecast_denoising.py:84, 87-89, 102-104: ``final_outputs = final_projection(h)`` (nn.Linear(d, 1)
    on the last pred_len decoder states), ``mll_error = -mll(dist, y).mean()`` (``elbo`` = the [1, B] / [B] output of
    DeepApproximateMLL(VariationalELBO(...)), see elbo_per_window), ``mse_loss = nn.MSELoss()(y_true, final_outputs)``,
    ``loss = mse_loss + torch.clip(lam, min=0, max=0.005) * mll_error``.  -> (final_outputs, loss, mse_loss)."""
    final = torch.nn.functional.linear(h, weight, bias)
    mll_err = -elbo.mean() if elbo is not None else torch.zeros((), dtype=h.dtype)
    mse = torch.nn.MSELoss()(y_true, final)
    loss = mse + torch.clip(lam, min=0, max=0.005) * mll_err
    return final, loss.reshape(()), mse

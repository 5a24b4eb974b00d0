"""CPU tests that pin the oracle: Random123 Philox known-answer vectors, closed-form known-answer cases
of the whitened SVGP predictive, agreement of the reference-order (gpytorch op sequence) and closed-form
restatements, the analytic kernel-order backward against autograd, and the committed golden fixtures."""
import glob
import math
import os

import numpy as np
import pytest
import torch

from oracle import gp_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


# Random123 kat_vectors, philox4x32 with 10 rounds
PHILOX_KAT = [
    ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
     [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
]


@pytest.mark.parametrize("ctr,key,want", PHILOX_KAT)
def test_philox_random123_kat(ctr, key, want):
    got = O.philox4x32_10(np.array(ctr, dtype=np.uint32), np.array(key, dtype=np.uint32))
    assert [int(v) for v in got] == want


def test_philox_counter_layout_and_golden():
    r = O.philox_bits(1234, (1 << 33) + 5, 64, 3)
    want = np.load(os.path.join(GOLDEN, "philox_bits_seed1234.npy"))
    assert np.array_equal(r, want)
    # element e of a longer stream equals a fresh stream started at offset e (geometry independence)
    a = O.philox_bits(7, 100, 50)
    b = O.philox_bits(7, 120, 30)
    assert np.array_equal(a[20:], b)
    # stream id and seed change the words
    assert not np.array_equal(O.philox_bits(7, 100, 4, 0), O.philox_bits(7, 100, 4, 1))
    assert not np.array_equal(O.philox_bits(7, 100, 4), O.philox_bits(8, 100, 4))


def test_philox_normal_moments():
    e = O.philox_normal(1, 0, 1 << 18)
    assert e.dtype == np.float32 and np.isfinite(e).all()
    assert abs(e.mean()) < 0.01 and abs(e.std() - 1.0) < 0.01
    assert abs(((e ** 4).mean()) - 3.0) < 0.1


def test_kat_m1_d1_hand_computed():
    """M = 1, D = 1: everything is scalar arithmetic."""
    z, x, ell_raw, os_raw, m, s, w, b = 0.3, 1.1, 0.2, -0.4, 0.7, 1.6, -0.5, 0.25
    f64 = dict(dtype=torch.float64)
    p = {"inducing_points": torch.tensor([[z]], **f64), "raw_lengthscale": torch.tensor([[ell_raw]], **f64),
         "raw_outputscale": torch.tensor(os_raw, **f64), "variational_mean": torch.tensor([m], **f64),
         "variational_stddev": torch.tensor([s], **f64), "weights": torch.tensor([[w]], **f64),
         "bias": torch.tensor([b], **f64), "raw_noise": torch.zeros(1, **f64)}
    ell = math.log1p(math.exp(ell_raw))
    osv = math.log1p(math.exp(os_raw))
    k = osv * math.exp(-0.5 * ((x - z) / ell) ** 2)
    Lc = math.sqrt(osv + 1e-4)
    a = k / Lc
    mean_want = a * m + x * w + b
    var_want = osv + 1e-4 + (s * s - 1) * a * a
    for fn in (O.svgp_predict_closed_form, O.svgp_predict_reference_order):
        mean, var = fn(p, torch.tensor([[[x]]], dtype=torch.float64))
        assert abs(mean.item() - mean_want) < 1e-12
        assert abs(var.item() - var_want) < 1e-12
    kl_want = 0.5 * (s * s + m * m - 1 - math.log(s * s))
    assert abs(O.kl_meanfield(p).item() - kl_want) < 1e-12


def test_kat_prior_regimes():
    D, M = 8, 16
    p = O.clone_params(O.init_params_exercise(D, M, 3), torch.float64)
    x, _, _, _ = O.make_inputs(4, 5, D, 4, dtype=torch.float64)
    osv = O.softplus(p["raw_outputscale"]).item()
    # q(u) = prior (m = 0, s = 1): KL = 0, variance = outputscale + jitter, mean = linear mean
    p0 = dict(p, variational_mean=torch.zeros(M, dtype=torch.float64),
              variational_stddev=torch.ones(M, dtype=torch.float64))
    mean, var = O.svgp_predict_closed_form(p0, x)
    assert O.kl_meanfield(p0).item() == 0.0
    assert torch.allclose(var, torch.full_like(var, osv + 1e-4), atol=1e-14)
    assert torch.allclose(mean, (x @ p["weights"]).squeeze(-1) + p["bias"], atol=1e-14)
    # far-away inputs: K(x, Z) = 0 => same as above whatever q(u) is
    mean, var = O.svgp_predict_closed_form(p, x + 1e3)
    assert torch.allclose(var, torch.full_like(var, osv + 1e-4), atol=1e-14)
    assert torch.allclose(mean, ((x + 1e3) @ p["weights"]).squeeze(-1) + p["bias"], atol=1e-9)


def test_kat_at_inducing_points_matches_unwhitened_marginal():
    """x = Z_j: the predictive equals the un-whitened q(u) = N(L m, L diag(s^2) L^T) marginal j plus the
    linear mean, up to the 1e-4 jitter."""
    D, M = 6, 12
    p = O.clone_params(O.init_params_exercise(D, M, 5), torch.float64)
    Z = p["inducing_points"]
    ell = O.softplus(p["raw_lengthscale"]).reshape(D)
    osv = O.softplus(p["raw_outputscale"])
    Lc = torch.linalg.cholesky(O.rbf_scale_direct(Z, Z, ell, osv) + O.JITTER * torch.eye(M, dtype=torch.float64))
    mean, var = O.svgp_predict_closed_form(p, Z.unsqueeze(0))
    want_mean = Lc @ p["variational_mean"] + (Z @ p["weights"]).squeeze(-1) + p["bias"]
    want_var = (Lc.pow(2) * p["variational_stddev"].pow(2)).sum(-1)
    assert (mean[0] - want_mean).abs().max() < 5e-3
    assert (var[0] - want_var).abs().max() < 5e-3


@pytest.mark.parametrize("B,L,D,M", [(4, 6, 16, 32), (2, 24, 64, 128), (3, 5, 7, 9)])
def test_reference_order_equals_closed_form(B, L, D, M):
    p = O.init_params_exercise(D, M, 1)
    x, _, _, _ = O.make_inputs(B, L, D, 2)
    p64 = O.clone_params(p, torch.float64)
    m_ref, v_ref = O.svgp_predict_reference_order(p64, x.double())
    m_cf, v_cf = O.svgp_predict_closed_form(p64, x.double())
    assert (m_ref - m_cf).abs().max() < 1e-11 and (v_ref - v_cf).abs().max() < 1e-11
    # fp32 reference order (what gpytorch computes) stays within 1e-5 of the truth in this regime
    m32, v32 = O.svgp_predict_reference_order(p, x)
    assert ((m32.double() - m_cf).abs().max() / m_cf.abs().max()) < 1e-5
    assert ((v32.double() - v_cf).abs().max() / v_cf.abs().max()) < 1e-5


@pytest.mark.parametrize("B,L,D,M", [(3, 4, 5, 6), (2, 6, 16, 40)])
def test_kernel_order_backward_matches_autograd(B, L, D, M):
    p = O.init_params_exercise(D, M, 1, dtype=torch.float64)
    x, y, gm, gv = O.make_inputs(B, L, D, 2, dtype=torch.float64)
    pg = O.clone_params(p, requires_grad=True)
    xg = x.clone().requires_grad_(True)
    mm, vv = O.svgp_predict_closed_form(pg, xg)
    gkl = 0.37
    ((gm * mm).sum() + (gv * vv).sum() + gkl * O.kl_meanfield(pg)).backward()
    G = O.svgp_backward_kernel_order(p, x, gm, gv, gkl)

    def rel(a, b):
        return ((a - b).abs().max() / (b.abs().max() + 1e-300)).item()
    assert rel(G.dx, xg.grad) < 1e-10
    for k in ["inducing_points", "raw_lengthscale", "raw_outputscale", "variational_mean", "variational_stddev",
              "weights", "bias"]:
        assert rel(getattr(G, k), pg[k].grad) < 1e-10, k


def test_gradcheck_closed_form():
    D, M = 3, 4
    p = O.init_params_exercise(D, M, 1, dtype=torch.float64)
    x, y, _, _ = O.make_inputs(2, 3, D, 2, dtype=torch.float64)
    names = list(p.keys())

    def f(xx, *vals):
        q = dict(zip(names, vals))
        return O.mll_error(q, xx, y, float(D), reference_order=False)
    vals = [p[k].clone().requires_grad_(True) for k in names]
    assert torch.autograd.gradcheck(f, (x.clone().requires_grad_(True), *vals), eps=1e-6, atol=1e-6)


def test_elbo_formula():
    B, L = 3, 5
    g = torch.Generator().manual_seed(0)
    mean, y = torch.randn(B, L, generator=g, dtype=torch.float64), torch.randn(B, L, generator=g, dtype=torch.float64)
    var = torch.rand(B, L, generator=g, dtype=torch.float64) + 0.2
    noise, kl, nd = torch.tensor(0.7, dtype=torch.float64), torch.tensor(1.3, dtype=torch.float64), 32.0
    e = O.elbo_per_window(mean, var, y, noise, kl, nd)
    # E_q[log N(y | f, noise)] with f ~ N(mean, var), by definition
    want = (torch.distributions.Normal(mean, noise.sqrt()).log_prob(y) - 0.5 * var / noise).sum(-1) / L - kl / nd
    assert torch.allclose(e, want, atol=1e-12)


def test_init_params_reference_matches_rng_order():
    p = O.init_params_reference(8, seed=77, M=5)
    torch.manual_seed(77)
    Z = torch.randn(5, 8)
    w = torch.randn(8, 1)
    b = torch.randn(1)
    assert torch.equal(p["inducing_points"], Z) and torch.equal(p["weights"], w) and torch.equal(p["bias"], b)
    assert p["variational_mean"].abs().max() == 0 and (p["variational_stddev"] == 1).all()


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "svgp_*.npz"))))
def test_oracle_reproduces_golden(path):
    g = np.load(path)
    p = {k[2:]: torch.from_numpy(g[k]).double().requires_grad_(True) for k in g.files if k.startswith("p_")}
    x = torch.from_numpy(g["x"]).double().requires_grad_(True)
    mean, var = O.svgp_predict_closed_form(p, x)
    assert np.allclose(mean.detach().numpy(), g["mean"], rtol=0, atol=1e-12)
    assert np.allclose(var.detach().numpy(), g["var"], rtol=0, atol=1e-12)
    kl = O.kl_meanfield(p)
    D = x.shape[-1]
    elbo = O.elbo_per_window(mean, var, torch.from_numpy(g["y"]).double(), O.noise_variance(p), kl, float(D))
    assert np.allclose(elbo.detach().numpy(), g["elbo"], atol=1e-12)
    seed, off, stream = (int(v) for v in g["philox_seed_offset_stream"])
    eps = torch.from_numpy(O.philox_normal(seed, off, mean.numel(), stream)).double().reshape(mean.shape)
    sample = O.rsample(mean, var, eps)
    assert np.allclose(sample.detach().numpy(), g["sample"], atol=1e-12)
    # the reference-order restatement (gpytorch's op sequence) agrees with the fixtures too
    m_ref, v_ref = O.svgp_predict_reference_order({k: v.detach() for k, v in p.items()}, x.detach())
    assert np.allclose(m_ref.numpy(), g["mean"], atol=1e-10) and np.allclose(v_ref.numpy(), g["var"], atol=1e-10)


def test_deepgp2_and_exact_gp_oracle_shapes():
    D, H, M = 6, 3, 8
    p1 = O.init_params_hidden_layer(D, H, M, 1, dtype=torch.float64)
    p2 = O.init_params_exercise(H, M, 2, dtype=torch.float64)
    x, _, _, _ = O.make_inputs(2, 4, D, 3, dtype=torch.float64)
    eps = torch.randn(2, 4, H, dtype=torch.float64)
    mean, var, h = O.deepgp2_predict(p1, p2, x, eps)
    assert mean.shape == (2, 4) and var.shape == (2, 4) and h.shape == (2, 4, H) and (var > 0).all()
    # exact GP posterior interpolates the training targets when noise is small
    tx = torch.linspace(0, 1, 7, dtype=torch.float64).unsqueeze(-1)
    ty = torch.sin(6 * tx).squeeze(-1)
    m, c = O.exact_gp_posterior(tx, ty, tx, torch.zeros(1, dtype=torch.float64), torch.tensor(-1.0, dtype=torch.float64),
                                torch.tensor(0.5, dtype=torch.float64), torch.tensor(-12.0, dtype=torch.float64))
    assert (m - ty).abs().max() < 5e-3 and torch.diagonal(c).min() > -1e-9

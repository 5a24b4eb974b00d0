"""The rows of SURVEY 8 composed as one training step on the device, against the fp64 oracle composition:

    window sampler (base_train.batch_sampled_data)              Utils/base_train.py:100-153, train.py:160-161
      -> embeddings of the encoder / decoder windows            (stand-in for the forecaster: two nn.Linear)
      -> ONE fused GP blur of both activations + ELBO           denoise_model_2.py:50-51, forecast_denoising.py:87-89
      -> x + proj_up(mean) on both                              denoise_model_2.py:36-38
      -> final projection + MSE + clip(lam) * mll_error         forecast_denoising.py:84, 102-104
      -> backward to every parameter
"""
import contextlib
import io
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from sampler_cases import CASES, column_definition, make_frame  # noqa: E402

from oracle import gp_oracle as O  # noqa: E402
from oracle import sampler_oracle as SO  # noqa: E402

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-300)).item()


def test_whole_step_matches_oracle_composition(cuda):
    from fine_grained_gaussian_process_forcasting_b200 import base_train as BT, gpcompat, step_ops
    from fine_grained_gaussian_process_forcasting_b200.DeepGP import DeepGPp
    from test_gpu_modules import load_params
    c = CASES["traffic_like"]
    D, M = 32, 128
    args = (c["train_percent"], c["max_samples"], c["time_steps"], c["num_encoder_steps"], c["pred_len"])
    with contextlib.redirect_stdout(io.StringIO()):
        train, _, _ = BT.batch_sampled_data(make_frame(c), *args, column_definition(), c["batch_size"], device=cuda)
        want_batches = SO.batch_sampled_data(make_frame(c), *args, column_definition(), c["batch_size"])[0]
    enc, dec, y = next(iter(train))
    enc0, dec0, y0 = (torch.from_numpy(a) for a in want_batches[0])
    assert torch.equal(enc.cpu(), enc0) and torch.equal(dec.cpu(), dec0) and torch.equal(y.cpu(), y0)
    B, Le, F = enc.shape
    Ld, P = dec.shape[1], y.shape[1]
    assert Ld == P                                               # the decoder window is the horizon here (8 steps)

    torch.manual_seed(7)
    emb_e, emb_d = torch.nn.Linear(F, D), torch.nn.Linear(F, D)
    proj_up, final = torch.nn.Linear(1, D), torch.nn.Linear(D, 1)
    lam = torch.nn.Parameter(torch.tensor([0.004]))
    p = O.init_params_exercise(D, M, 21)
    mods = [emb_e, emb_d, proj_up, final]
    with gpcompat.num_likelihood_samples(1):
        gp = DeepGPp(D, 3, num_inducing=M).to(cuda)
        load_params(gp, p)
        dm = [torch.nn.Linear(m.in_features, m.out_features).to(cuda) for m in mods]
        for a, b_ in zip(dm, mods):
            a.load_state_dict(b_.state_dict())
        lam_d = torch.nn.Parameter(lam.detach().to(cuda))
        he, hd = dm[0](enc), dm[1](dec)                          # [B, Le, D], [B, Ld, D]
        flat = torch.cat([he.reshape(-1, D), hd.reshape(-1, D)], 0)
        out_e, out_d = gp.blur_segments(flat, [(B, Le), (B, Ld)], y_last=y.squeeze(-1))
        he_n = step_ops.blur_apply(he, out_e.mean, dm[2].weight, dm[2].bias)
        hd_n = step_ops.blur_apply(hd, out_d.mean, dm[2].weight, dm[2].bias)
        final_out, loss, mse = step_ops.forecast_loss(dm[3], hd_n, y, out_d.elbo, lam_d)
        total = loss + 1e-3 * he_n.square().mean()               # the encoder side feeds the denoiser in the reference
        total.backward()
    torch.cuda.synchronize()

    # ---- the same step through the oracle, float64 on the host
    p64 = O.clone_params(p, torch.float64, requires_grad=True)
    p64["raw_noise"] = torch.zeros(1, dtype=torch.float64, requires_grad=True)
    m64 = [torch.nn.Linear(m.in_features, m.out_features).double() for m in mods]
    for a, b_ in zip(m64, mods):
        a.load_state_dict(b_.state_dict())
    lam64 = lam.detach().double().requires_grad_(True)
    he64, hd64 = m64[0](enc0.double()), m64[1](dec0.double())
    me, ve = O.svgp_predict_closed_form(p64, he64)
    md, vd = O.svgp_predict_closed_form(p64, hd64)
    elbo64 = O.elbo_per_window(md, vd, y0.double().squeeze(-1), O.noise_variance(p64), O.kl_meanfield(p64), float(D))
    he_n64 = O.blur_apply_reference(he64, me.unsqueeze(0), m64[2].weight, m64[2].bias)
    hd_n64 = O.blur_apply_reference(hd64, md.unsqueeze(0), m64[2].weight, m64[2].bias)
    f64, loss64, mse64 = O.forecast_loss_reference(hd_n64, m64[3].weight, m64[3].bias, y0.double(), elbo64, lam64)
    total64 = loss64 + 1e-3 * he_n64.square().mean()
    total64.backward()

    assert rel(final_out, f64) < 1e-4 and rel(loss, loss64) < 1e-4 and rel(mse, mse64) < 1e-4
    assert rel(out_d.elbo.reshape(-1), elbo64.reshape(-1)) < 1e-4
    for a, b_ in zip(dm, m64):
        assert rel(a.weight.grad, b_.weight.grad) < 2e-4 and rel(a.bias.grad, b_.bias.grad) < 2e-4
    assert rel(lam_d.grad, lam64.grad) < 1e-4
    hl = gp.hidden_layer
    assert rel(hl.variational_strategy.inducing_points.grad, p64["inducing_points"].grad) < 2e-4
    assert rel(hl.variational_strategy._variational_distribution.variational_mean.grad, p64["variational_mean"].grad) < 2e-4
    assert rel(hl.covar_module.base_kernel.raw_lengthscale.grad.reshape(-1), p64["raw_lengthscale"].grad.reshape(-1)) < 2e-4

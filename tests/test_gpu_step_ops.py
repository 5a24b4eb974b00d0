"""SURVEY section 8 (f) ranks 1 and 2 on the GPU: the fused blur application (x + proj_up(mean)) and the fused loss
assembly (final projection + MSE + clipped-lambda ELBO term) against the oracle's plain-torch restatement of the
reference lines in float64 (values and every gradient).  Tolerance 2e-6 (fp32 arithmetic, fp64 final reductions)."""
import pytest
import torch

from oracle import gp_oracle as O

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = a.detach().double().cpu().reshape(-1)
    b = b.detach().double().cpu().reshape(-1)
    return ((a - b).abs().max() / (b.abs().max() + 1e-300)).item()


@pytest.mark.parametrize("B,L,D", [(256, 24, 64), (7, 13, 10), (64, 192, 32), (3, 5, 128), (33, 24, 16)])
def test_blur_apply_matches_reference(cuda, B, L, D):
    from fine_grained_gaussian_process_forcasting_b200.step_ops import blur_apply
    g = torch.Generator().manual_seed(B * 1000 + D)
    x = torch.randn(B, L, D, generator=g)
    eps = torch.randn(1, B, L, generator=g)
    lin = torch.nn.Linear(1, D)
    go = torch.randn(B, L, D, generator=g)
    x64, e64 = x.double().requires_grad_(True), eps.double().requires_grad_(True)
    w64, b64 = lin.weight.detach().double().requires_grad_(True), lin.bias.detach().double().requires_grad_(True)
    ref = O.blur_apply_reference(x64, e64, w64, b64)
    ref.backward(go.double())
    xc, ec = x.to(cuda).requires_grad_(True), eps.to(cuda).requires_grad_(True)
    wc, bc = lin.weight.detach().to(cuda).requires_grad_(True), lin.bias.detach().to(cuda).requires_grad_(True)
    out = blur_apply(xc, ec[0], wc, bc)
    out.backward(go.to(cuda))
    torch.cuda.synchronize()
    assert rel(out, ref) < 2e-6
    assert rel(xc.grad, x64.grad) == 0.0
    assert rel(ec.grad, e64.grad) < 2e-6 and rel(wc.grad, w64.grad) < 2e-6 and rel(bc.grad, b64.grad) < 2e-6


def test_add_gp_noise_mirrors_reference_call(cuda):
    from fine_grained_gaussian_process_forcasting_b200.DeepGP import DeepGPp
    from fine_grained_gaussian_process_forcasting_b200 import gpcompat
    from fine_grained_gaussian_process_forcasting_b200.step_ops import add_gp_noise
    B, L, D = 8, 24, 32
    with gpcompat.num_likelihood_samples(1):
        gp = DeepGPp(D, 5, num_inducing=32).to(cuda)
        proj_up = torch.nn.Linear(1, D).to(cuda)
        x = torch.randn(B, L, D, device=cuda, requires_grad=True)
        x_noisy, dist = add_gp_noise(gp, proj_up, x)
        mean = dist.mean                                        # [1, B, L]
        ref = x + proj_up(mean.permute(1, 2, 0))
        assert x_noisy.shape == (B, L, D) and rel(x_noisy, ref) < 2e-6
        x_noisy.square().sum().backward()
        assert x.grad is not None and proj_up.weight.grad is not None
        assert gp.hidden_layer.variational_strategy.inducing_points.grad is not None


@pytest.mark.parametrize("B,Ltot,P,D,lam", [(256, 48, 24, 64, 0.003), (5, 30, 7, 10, -0.2), (64, 24, 24, 32, 0.2),
                                             (9, 40, 24, 128, 0.005)])
def test_forecast_loss_matches_reference(cuda, B, Ltot, P, D, lam):
    from fine_grained_gaussian_process_forcasting_b200.step_ops import forecast_loss
    g = torch.Generator().manual_seed(B + 17 * D)
    dec = torch.randn(B, Ltot, D, generator=g)
    y = torch.randn(B, P, 1, generator=g)
    elbo = torch.randn(1, B, generator=g)
    lin = torch.nn.Linear(D, 1)
    lamt = torch.tensor([lam])
    gfin = torch.randn(B, P, 1, generator=g)
    d64 = dec.double().requires_grad_(True)
    w64, b64 = lin.weight.detach().double().requires_grad_(True), lin.bias.detach().double().requires_grad_(True)
    e64, l64 = elbo.double().requires_grad_(True), lamt.double().requires_grad_(True)
    fin_r, loss_r, mse_r = O.forecast_loss_reference(d64[:, -P:, :], w64, b64, y.double(), e64, l64)
    (1.7 * loss_r + 0.3 * mse_r + (gfin.double() * fin_r).sum()).backward()
    dc = dec.to(cuda).requires_grad_(True)
    linc = torch.nn.Linear(D, 1).to(cuda)
    with torch.no_grad():
        linc.weight.copy_(lin.weight); linc.bias.copy_(lin.bias)
    ec, lc = elbo.to(cuda).requires_grad_(True), lamt.to(cuda).requires_grad_(True)
    fin, loss, mse = forecast_loss(linc, dc[:, -P:, :], y.to(cuda), ec, lc)     # the slice is read in place
    (1.7 * loss + 0.3 * mse + (gfin.to(cuda) * fin).sum()).backward()
    torch.cuda.synchronize()
    assert fin.shape == (B, P, 1)
    assert rel(fin, fin_r) < 2e-6 and rel(loss, loss_r) < 2e-6 and rel(mse, mse_r) < 2e-6
    assert rel(dc.grad, d64.grad) < 2e-6
    assert rel(linc.weight.grad, w64.grad) < 5e-6 and rel(linc.bias.grad, b64.grad) < 5e-6
    assert rel(ec.grad, e64.grad) < 2e-6
    assert (lc.grad.double().cpu() - l64.grad).abs().max().item() <= 2e-6 * max(1.0, l64.grad.abs().max().item())

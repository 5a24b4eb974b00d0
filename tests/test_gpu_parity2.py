"""Round-2 parity cases (VERDICT r01 "weak" list): ill-conditioned Kzz, the two-layer stack on the tensor-core path,
independent samples along S > 1, sharded multi-output counters.

Error metric everywhere: max|cuda - oracle| / max|oracle| against the float64 closed-form oracle (a NORM-wise bound,
not element-wise relative).  The ill-conditioned case follows SURVEY section 7: the criterion is
err(cuda) <= err(reference-order fp32), i.e. the CUDA path must not be less accurate than the reference's own fp32
pipeline on the same inputs."""
import numpy as np
import pytest
import torch

from oracle import gp_oracle as O

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = a.detach().double().cpu().reshape(-1)
    b = b.detach().double().cpu().reshape(-1)
    return ((a - b).abs().max() / (b.abs().max() + 1e-300)).item()


def _ill_conditioned_params(D, M, seed, eps):
    """Inducing points in near-duplicate pairs (|z_2i - z_2i+1| ~ eps): cond(Kzz + 1e-4 I) ~ 1e5 .. 1e6."""
    p = O.init_params_exercise(D, M, seed)
    g = torch.Generator().manual_seed(seed + 100)
    Z = p["inducing_points"].clone()
    Z[1::2] = Z[0::2][: Z[1::2].shape[0]] + eps * torch.randn(Z[1::2].shape, generator=g)
    p["inducing_points"] = Z
    return p


@pytest.mark.parametrize("D,M,eps", [(64, 64, 2e-2), (64, 256, 2e-2), (32, 256, 5e-3)])
def test_ill_conditioned_kzz(cuda, D, M, eps):
    from fine_grained_gaussian_process_forcasting_b200 import ops
    B, L = 16, 24
    p32 = _ill_conditioned_params(D, M, 31, eps)
    x32, _, gm32, gv32 = O.make_inputs(B, L, D, seed=32)
    p64 = O.clone_params(p32, torch.float64)
    ell = O.softplus(p64["raw_lengthscale"]).reshape(D)
    Kzz = O.rbf_scale_direct(p64["inducing_points"], p64["inducing_points"], ell, O.softplus(p64["raw_outputscale"])) \
        + O.JITTER * torch.eye(M, dtype=torch.float64)
    cond = torch.linalg.cond(Kzz).item()
    assert cond > 3e4, cond                                     # the case is what it claims to be
    mean_o, var_o = O.svgp_predict_closed_form(p64, x32.double())
    # the reference's own op order in fp32 (fp64 Cholesky / solve, fp32 kernel build) on the same inputs
    mean_r, var_r = O.svgp_predict_reference_order(p32, x32)
    e_ref = max(rel(mean_r, mean_o), rel(var_r, var_o))
    pd = {k: v.to(cuda) for k, v in p32.items()}
    mean, var, _, _, info = ops.svgp_predict(x32.to(cuda), pd["inducing_points"], pd["raw_lengthscale"],
                                             pd["raw_outputscale"], pd["variational_mean"], pd["variational_stddev"],
                                             pd["weights"], pd["bias"])
    torch.cuda.synchronize()
    assert int(info.max().item()) == 0
    e_new = max(rel(mean, mean_o), rel(var, var_o))
    print(f"ill-conditioned D={D} M={M}: cond(Kzz + jI) = {cond:.2e}  err(cuda) = {e_new:.2e}  err(reference-order fp32) = {e_ref:.2e}")
    # never worse than 3x the reference's own fp32 noise floor, and within the 3xTF32 allowance of the north star
    assert e_new <= max(3.0 * e_ref, 2e-5), (e_new, e_ref)
    assert e_new < 1e-3


def _load_layer(layer, p):
    vs = layer.variational_strategy
    with torch.no_grad():
        vs.inducing_points.copy_(p["inducing_points"])
        layer.covar_module.base_kernel.raw_lengthscale.copy_(p["raw_lengthscale"])
        if "raw_outputscale" in p:
            layer.covar_module.raw_outputscale.copy_(p["raw_outputscale"].reshape(layer.covar_module.raw_outputscale.shape))
        vs._variational_distribution.variational_mean.copy_(p["variational_mean"])
        vs._variational_distribution._variational_stddev.copy_(p["variational_stddev"])
        layer.mean_module.weights.copy_(p["weights"])
        layer.mean_module.bias.copy_(p["bias"])
        vs.variational_params_initialized.fill_(1)


def test_two_layer_deepgp_tensor_core_path(cuda):
    """configs[3] shape of the two-layer stack on the tcgen05 path: D=64, H=10, M=256 (layer 2 input width 10)."""
    from fine_grained_gaussian_process_forcasting_b200.DeepGP import DeepGP2
    from fine_grained_gaussian_process_forcasting_b200 import gpcompat
    B, L, D, H, M = 12, 24, 64, 10, 256
    p1 = O.init_params_hidden_layer(D, H, M, 5)
    p2 = O.init_params_exercise(H, M, 6)
    x, y, _, _ = O.make_inputs(B, L, D, 7)
    with gpcompat.num_likelihood_samples(1):
        net = DeepGP2(D, 11, hidden_dims=H, num_inducing=M).to(cuda)
        _load_layer(net.hidden_layer, p1)
        _load_layer(net.last_layer, p2)
        net.hidden_layer.set_rng(77, 0, stream=1)
        xd = x.to(cuda).requires_grad_(True)
        mean, dist = net.predict(xd)
        mll = gpcompat.DeepApproximateMLL(gpcompat.VariationalELBO(net.likelihood, net, D))
        loss = -mll(dist, y.to(cuda).unsqueeze(0)).mean()
        loss.backward()
    N = B * L
    eps = torch.stack([torch.from_numpy(O.philox_normal(77, h * N, N, 1)).reshape(B, L) for h in range(H)], -1).double()
    p1_64 = O.clone_params(p1, torch.float64, requires_grad=True)
    p2_64 = O.clone_params(p2, torch.float64, requires_grad=True)
    x64 = x.double().requires_grad_(True)
    mo, vo, ho = O.deepgp2_predict(p1_64, p2_64, x64, eps)
    kl = O.kl_hidden_layer(p1_64) + O.kl_meanfield(p2_64)
    noise = O.softplus(torch.zeros((), dtype=torch.float64)) + O.NOISE_LOWER
    lo = -O.elbo_per_window(mo, vo, y.double(), noise, kl, float(D)).mean()
    lo.backward()
    l1, l2 = net.hidden_layer, net.last_layer
    errs = {"mean": rel(mean[0], mo), "var": rel(dist.variance[0], vo), "loss": abs(loss.item() - lo.item()) / abs(lo.item()),
            "dx": rel(xd.grad, x64.grad),
            "dZ1": rel(l1.variational_strategy.inducing_points.grad, p1_64["inducing_points"].grad),
            "dZ2": rel(l2.variational_strategy.inducing_points.grad, p2_64["inducing_points"].grad),
            "dw1": rel(l1.mean_module.weights.grad, p1_64["weights"].grad)}
    print("two-layer tensor path:", {k: f"{v:.1e}" for k, v in errs.items()})
    assert errs["mean"] < 1e-4 and errs["var"] < 1e-4 and errs["loss"] < 1e-4      # 3xTF32 class (north star: 1e-3)
    assert errs["dx"] < 5e-4 and errs["dZ1"] < 5e-4 and errs["dZ2"] < 5e-4 and errs["dw1"] < 5e-4


def test_samples_independent_along_likelihood_samples(cuda):
    """S > 1 (gpytorch's default num_likelihood_samples = 10): the sample handed to the next layer must differ
    along S (gpytorch draws Normal(mean, sqrt(var)).rsample() on the EXPANDED distribution)."""
    from fine_grained_gaussian_process_forcasting_b200.DeepGP import DeepGP2, DeepGPp
    from fine_grained_gaussian_process_forcasting_b200 import gpcompat
    B, L, D, H, M = 3, 8, 16, 4, 32
    x = torch.randn(B, L, D, generator=torch.Generator().manual_seed(0)).to(cuda)
    with gpcompat.num_likelihood_samples(4):
        net = DeepGP2(D, 3, hidden_dims=H, num_inducing=M).to(cuda)
        h = net.hidden_layer(x)                                   # MultitaskMultivariateNormal, batch [4, B]
        assert tuple(h.mean.shape) == (4, B, L, H)
        s = h.rsample()
        assert tuple(s.shape) == (4, B, L, H)
        assert not torch.equal(s[0], s[1]) and not torch.equal(s[1], s[2])
        assert torch.equal(h.mean[0], h.mean[1])                  # the mean itself is only expanded
        mean, dist = net.predict(x)
        assert tuple(mean.shape) == (4, B, L) and not torch.equal(mean[0], mean[1])
    with gpcompat.num_likelihood_samples(1):
        one = DeepGPp(D, 3, num_inducing=M).to(cuda)
        out = one.blur(x)
        assert tuple(out.sample.shape) == (1, B, L)               # S = 1 keeps the fused sample (a view)


def test_sharded_multi_output_counters_match_single_rank(cuda):
    """with batch sharding, GP h of a multi-output layer must draw the counters h * N_global + global n,
    so that two shards reproduce the single-rank samples bit for bit."""
    from fine_grained_gaussian_process_forcasting_b200.DeepGP import DeepGP2
    from fine_grained_gaussian_process_forcasting_b200.distributed import ShardedGPBlur
    from fine_grained_gaussian_process_forcasting_b200 import gpcompat
    B, L, D, H, M = 8, 6, 16, 3, 32
    x = torch.randn(B, L, D, generator=torch.Generator().manual_seed(1)).to(cuda)
    with gpcompat.num_likelihood_samples(1):
        net = DeepGP2(D, 5, hidden_dims=H, num_inducing=M).to(cuda)
        wrap = ShardedGPBlur(net, broadcast=False)
        with torch.no_grad():
            wrap.step_index = 0
            wrap(x, None, first_global_window=0, global_windows=B)
            full = net.hidden_layer(x).rsample() if False else None
        # layer-1 samples of the full batch vs the two half batches with global offsets
        def layer1_sample(xs, first):
            wrap.step_index = 0
            Lh = net.hidden_layer
            total = B * L
            Lh._rng_offset = first * L
            Lh._rng_h_stride = total
            with torch.no_grad():
                return Lh(xs).sample_value.clone()
        s_full = layer1_sample(x, 0)
        s_a = layer1_sample(x[:B // 2], 0)
        s_b = layer1_sample(x[B // 2:], B // 2)
        assert torch.equal(torch.cat([s_a, s_b], dim=1), s_full)


def test_not_psd_raises_after_jitter_retries(cuda):
    """gpytorch's psd_safe_cholesky convention: retry with extra jitter, then NotPSDError (default check is ON)."""
    from fine_grained_gaussian_process_forcasting_b200.DeepGP import DeepGPp
    from fine_grained_gaussian_process_forcasting_b200 import gpcompat
    with gpcompat.num_likelihood_samples(1):
        m = DeepGPp(16, 1, num_inducing=32).to(cuda)
        x = torch.randn(4, 8, 16, device=cuda)
        m.predict(x)                                              # healthy parameters: fine
        with torch.no_grad():
            m.hidden_layer.variational_strategy.inducing_points[3, 0] = float("nan")
        with pytest.raises(gpcompat.NotPSDError):
            m.predict(x)
        with gpcompat.check_cholesky(False):                      # opt-out: flag only
            m.predict(x)
            assert int(m.hidden_layer.last_info.max().item()) != 0


def test_grad_sink_matches_autograd_accumulation(cuda):
    """FlatGradBucket(module=...) lets the M x M backward kernel accumulate into the flat buffer: same gradients as the
    ordinary autograd accumulation, also over two backward passes (accumulate semantics of .grad)."""
    from fine_grained_gaussian_process_forcasting_b200.DeepGP import DeepGPp
    from fine_grained_gaussian_process_forcasting_b200.distributed import FlatGradBucket, gp_parameters
    from fine_grained_gaussian_process_forcasting_b200 import gpcompat
    B, L, D, M = 6, 24, 32, 128
    p = O.init_params_exercise(D, M, 41)
    x, y, _, _ = O.make_inputs(B, L, D, 42)

    def run(use_sink, passes):
        with gpcompat.num_likelihood_samples(1):
            m = DeepGPp(D, 1, num_inducing=M).to(cuda)
            _load_layer(m.hidden_layer, p)
            bucket = FlatGradBucket(gp_parameters(m), module=m) if use_sink else None
            assert (m.hidden_layer._grad_sink is not None) == use_sink
            for _ in range(passes):
                out = m.blur(x.to(cuda).requires_grad_(True), y.to(cuda))
                (out.mean.sum() + out.sample.sum() - out.elbo.mean()).backward()
                m.hidden_layer._rng_offset = 0
            torch.cuda.synchronize()
            return {n: q.grad.detach().clone().reshape(-1) for n, q in m.named_parameters()}

    for passes in (1, 2):
        ga, gb = run(False, passes), run(True, passes)
        for n in ga:
            assert rel(gb[n], ga[n]) < 1e-6, (n, passes, rel(gb[n], ga[n]))


def test_gpytorch_shim_path_runs_forward_and_backward_on_gpu(cuda):
    """The `import gpytorch` shim on a GPU: a user-side deep GP layer written ONLY against the
    gpytorch names the reference imports (DeepGP.py:6-11), evaluated with the reference's call protocol (predict ->
    likelihood(dist).mean, DeepApproximateMLL(VariationalELBO(...))(dist, y), forecast_denoising.py:87-89), checked
    against the oracle.  (The reference's own DeepGP.py cannot travel to the GPU box; tests/test_host.py builds it on
    the shim in the CPU container.)"""
    import fine_grained_gaussian_process_forcasting_b200 as pkg
    pkg.install_gpytorch_shim()
    import gpytorch
    from gpytorch.means import LinearMean
    from gpytorch.kernels import RBFKernel, ScaleKernel
    from gpytorch.variational import VariationalStrategy, MeanFieldVariationalDistribution
    from gpytorch.distributions import MultivariateNormal
    from gpytorch.models.deep_gps import DeepGPLayer, DeepGP
    from gpytorch.likelihoods import GaussianLikelihood
    from gpytorch.mlls import DeepApproximateMLL, VariationalELBO

    B, L, D, M = 5, 24, 32, 128

    class Layer(DeepGPLayer):
        def __init__(self):
            z = torch.randn(M, D)
            q = MeanFieldVariationalDistribution(num_inducing_points=M, batch_shape=torch.Size([]))
            super().__init__(VariationalStrategy(self, z, q, learn_inducing_locations=True), D, None)
            self.mean_module = LinearMean(D)
            self.covar_module = ScaleKernel(RBFKernel(batch_shape=torch.Size([]), ard_num_dims=D), batch_shape=torch.Size([]),
                                            ard_num_dims=None)

        def forward(self, x):
            return MultivariateNormal(self.mean_module(x), self.covar_module(x))

    class Net(DeepGP):
        def __init__(self):
            super().__init__()
            self.hidden_layer = Layer()
            self.likelihood = GaussianLikelihood()

        def forward(self, x):
            return self.hidden_layer(x)

    p = O.init_params_exercise(D, M, 51)
    x, y, _, _ = O.make_inputs(B, L, D, 52)
    with gpytorch.settings.num_likelihood_samples(1):
        net = Net().to(cuda)
        _load_layer(net.hidden_layer, p)
        xd = x.to(cuda).requires_grad_(True)
        dist = net(xd)
        mean = net.likelihood(dist).mean
        mll = DeepApproximateMLL(VariationalELBO(net.likelihood, net, D))
        loss = -mll(dist, y.to(cuda).unsqueeze(0)).mean()
        loss.backward()
    p64 = O.clone_params(p, torch.float64, requires_grad=True)
    p64["raw_noise"] = torch.zeros(1, dtype=torch.float64)
    x64 = x.double().requires_grad_(True)
    lo = O.mll_error(p64, x64, y.double(), float(D), reference_order=False)
    lo.backward()
    mo, vo = O.svgp_predict_closed_form(p64, x64)
    assert tuple(mean.shape) == (1, B, L)
    assert rel(mean[0], mo) < 1e-4 and rel(dist.variance[0], vo) < 1e-4
    assert abs(loss.item() - lo.item()) < 1e-4 * abs(lo.item())
    assert rel(xd.grad, x64.grad) < 2e-4
    assert rel(net.hidden_layer.variational_strategy.inducing_points.grad, p64["inducing_points"].grad) < 2e-4


@pytest.mark.parametrize("D,M,Ls", [(64, 256, (48, 24)), (32, 32, (24, 7, 13)), (64, 128, (40, 24))])
def test_blur_segments_equals_separate_calls(cuda, D, M, Ls):
    """DeepGPp.blur_segments (ONE fused evaluation over the concatenated activations of a step, per-segment upstream
    gradients read in place by the backward kernels) == one blur call per activation: outputs bit-identical (same
    kernels, same Philox counters), gradients to fp32 accumulation order."""
    from fine_grained_gaussian_process_forcasting_b200.DeepGP import DeepGPp
    from fine_grained_gaussian_process_forcasting_b200 import gpcompat, ops
    B = 16
    p = O.init_params_exercise(D, M, 51)
    g = torch.Generator().manual_seed(52)
    xs = [torch.randn(B, L, D, generator=g) for L in Ls]
    y = torch.randn(B, Ls[-1], generator=g)
    gms = [torch.randn(1, B, L, generator=g).to(cuda) for L in Ls]
    gss = [torch.randn(1, B, L, generator=g).to(cuda) for L in Ls]
    gss[0] = None                                     # a segment without a gradient on its sample

    def run(fused):
        with gpcompat.num_likelihood_samples(1):
            m = DeepGPp(D, 3, num_inducing=M).to(cuda)
            _load_layer(m.hidden_layer, p)
            m.hidden_layer._rng_offset = 0
            if fused:
                flat = torch.cat([x.reshape(-1, D) for x in xs]).to(cuda)
                views, o = [], 0
                for x in xs:
                    views.append(flat[o:o + x.shape[0] * x.shape[1]].view(x.shape))
                    o += x.shape[0] * x.shape[1]
                xf = ops.as_one_buffer(views)
                assert xf.data_ptr() == flat.data_ptr() and xf.shape == flat.shape      # a view, no copy
                xf = xf.detach().requires_grad_(True)
                outs = m.blur_segments(xf, [(B, L) for L in Ls], y.to(cuda))
                leaves = [xf]
            else:
                leaves = [x.to(cuda).requires_grad_(True) for x in xs]
                outs = [m.blur(lv, y.to(cuda) if i == len(xs) - 1 else None) for i, lv in enumerate(leaves)]
            heads, grads = [], []
            for o_, gm_, gs_ in zip(outs, gms, gss):
                heads.append(o_.mean); grads.append(gm_)
                if gs_ is not None:
                    heads.append(o_.sample); grads.append(gs_)
            heads.append(outs[-1].elbo); grads.append(torch.full_like(outs[-1].elbo, -1.0 / B))
            torch.autograd.backward(heads, grads)
            torch.cuda.synchronize()
            dx = torch.cat([lv.grad.reshape(-1, D) for lv in leaves])
            pg = {n: q.grad.detach().clone().reshape(-1) for n, q in m.named_parameters()}
            return outs, dx, pg

    oa, dxa, ga = run(False)
    ob, dxb, gb = run(True)
    for a_, b_ in zip(oa, ob):
        assert torch.equal(a_.mean, b_.mean) and torch.equal(a_.variance, b_.variance) and torch.equal(a_.sample, b_.sample)
    assert torch.equal(oa[-1].elbo, ob[-1].elbo)
    assert rel(dxb, dxa) < 1e-6
    for n in ga:
        assert rel(gb[n], ga[n]) < 2e-5, (n, rel(gb[n], ga[n]))

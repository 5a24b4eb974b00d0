"""ATA attention head: oracle vs vectors from the UNMODIFIED reference head
(tests/golden/ata_ref_*.npz, generator tests/golden/make_ata_golden.py) on the CPU; fused CUDA core and the module
mirror vs both on the GPU."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from ata_cases import CASES, make_inputs  # noqa: E402

from oracle import ata_oracle as AO  # noqa: E402
from fine_grained_gaussian_process_forcasting_b200 import ATA as ATAmod  # noqa: E402

TOL = 2e-5        # fp32 pipeline (conv, batch norm, exp): max-abs error / max-abs reference


def golden(name):
    return dict(np.load(os.path.join(HERE, "golden", f"ata_ref_{name}.npz")))


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-300)).item()


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_vectors(name):
    b, h, l, lk, dk, seed = CASES[name]
    Q, K, V, Gc = make_inputs(name)
    Q, K, V = (t.clone().requires_grad_(True) for t in (Q, K, V))
    ctx, attn = AO.ata_forward(Q, K, V, AO.init_weights(dk, h, seed), dk)
    (ctx * Gc).sum().backward()
    g = golden(name)
    assert rel(ctx.detach(), g["context"]) < 1e-6 and rel(attn.detach(), g["attn"]) < 1e-6
    assert rel(Q.grad, g["gQ"]) < 1e-5 and rel(K.grad, g["gK"]) < 1e-5 and rel(V.grad, g["gV"]) < 1e-6


def test_module_mirror_on_host():
    """Same sub-module names / shapes and the same initial weights as the reference constructor draws (RNG order)."""
    head = ATAmod.ATA(d_k=4, device="cpu", h=8, seed=1234)
    keys = set(head.state_dict())
    for side in ("q", "k"):
        for i in range(4):
            for leaf in ("0.weight", "0.bias", "1.weight", "1.bias", "1.running_mean", "1.running_var", "1.num_batches_tracked"):
                assert f"conv_list_{side}.{i}.{leaf}" in keys
    assert {"proj_back_q.weight", "proj_back_q.bias", "proj_back_k.weight", "proj_back_k.bias"} <= keys
    w = AO.init_weights(4, 8, 1234)
    for i, f in enumerate(AO.FILTERS):
        assert torch.equal(head.conv_list_k[i][0].weight, w[f"k{f}"][0]) and torch.equal(head.conv_list_q[i][0].bias, w[f"q{f}"][1])
    x = torch.randn(2, 8, 12, 4)
    with pytest.raises(RuntimeError):
        head(x, x, x)                       # no CPU fallback


def test_cached_head_replays_constructor_side_effects():
    """ATA.cached: same weights as a fresh construction and the same generator states afterwards (torch, numpy,
    random), on the first and on later calls; running statistics frozen, weights without gradient."""
    import random
    fresh = ATAmod.ATA(d_k=4, device="cpu", h=2, seed=99)
    want = (torch.get_rng_state().clone(), np.random.get_state()[1].copy(), random.getstate())
    ATAmod.ATA._cache.clear()
    for _ in range(3):
        torch.manual_seed(5); np.random.seed(5); random.seed(5)       # whatever the caller's generators held before
        torch.randn(7)
        head = ATAmod.ATA.cached(d_k=4, device="cpu", h=2, seed=99)
        assert torch.equal(torch.get_rng_state(), want[0]) and np.array_equal(np.random.get_state()[1], want[1])
        assert random.getstate() == want[2]
        for (n1, a), (n2, b) in zip(fresh.state_dict().items(), head.state_dict().items()):
            assert n1 == n2 and torch.equal(a, b)
    assert head is ATAmod.ATA.cached(d_k=4, device="cpu", h=2, seed=99) and head.training
    assert all(not q.requires_grad for q in head.parameters())
    assert all(m.momentum == 0.0 for m in head.modules() if isinstance(m, torch.nn.BatchNorm1d))
    assert ATAmod.ATA.cached(d_k=4, device="cpu", h=2, seed=100) is not head


# ------------------------------------------------------------------------------------------------------------------
@pytest.fixture
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda:0")


@pytest.fixture
def exact_convs():
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False          # the golden vectors are CPU fp32; cuDNN would pick TF32 convolutions
    yield
    torch.backends.cudnn.allow_tf32 = old


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_module_matches_reference_vectors(cuda, exact_convs, name):
    b, h, l, lk, dk, seed = CASES[name]
    Q, K, V, Gc = make_inputs(name)
    Q, K, V = (t.to(cuda).requires_grad_(True) for t in (Q, K, V))       # keeps the [b, l, h, d] memory of q_s / k_s / v_s
    head = ATAmod.ATA(d_k=dk, device="cpu", h=h, seed=seed).to(cuda)     # CPU generator, like the golden run
    context, attn = head(Q=Q, K=K, V=V, need_attn=True)
    assert context.shape == (b, h, l, dk) and context.transpose(1, 2).is_contiguous()
    (context * Gc.to(cuda)).sum().backward()
    g = golden(name)
    assert rel(context.detach(), g["context"]) < TOL and rel(attn, g["attn"]) < TOL
    assert rel(Q.grad, g["gQ"]) < 10 * TOL and rel(K.grad, g["gK"]) < 10 * TOL and rel(V.grad, g["gV"]) < TOL
    assert head(Q=Q, K=K, V=V)[1] is None                                # default: attn is not materialised


@pytest.mark.gpu
@pytest.mark.parametrize("b,h,l,lk,dk,dv", [(256, 8, 192, 192, 4, 4), (16, 8, 24, 192, 4, 4), (5, 4, 300, 37, 8, 8),
                                            (3, 2, 9, 17, 16, 16), (2, 1, 33, 65, 3, 7), (2, 2, 40, 40, 64, 64)])
def test_fused_core_vs_fp64(cuda, b, h, l, lk, dk, dv):
    """The fused core against the reference's op sequence in float64 on the same device, forward and gradients;
    first row = the encoder self-attention of configs[1] (b = 256, 8 heads, l = 192)."""
    g = torch.Generator(device=cuda).manual_seed(b * 1000 + l)
    G = 4 * dk
    qp = torch.relu(torch.randn(b, h, l, G, device=cuda, generator=g)).requires_grad_(True)     # post-ReLU, many exact zeros
    kp = torch.relu(torch.randn(b, h, lk, G, device=cuda, generator=g)).requires_grad_(True)
    v = torch.randn(b, lk, h, dv, device=cuda, generator=g).transpose(1, 2).requires_grad_(True)
    gc = torch.randn(b, h, l, dv, device=cuda, generator=g)
    ctx, qpool, kpool = ATAmod.ata_core(qp, kp, v, dk)
    (ctx * gc).sum().backward()
    got = (ctx.detach(), qp.grad.clone(), kp.grad.clone(), v.grad.clone())
    q64, k64, v64 = (t.detach().double().requires_grad_(True) for t in (qp, kp, v))
    want_ctx, _ = AO.core(q64, k64, v64, dk)
    (want_ctx * gc.double()).sum().backward()
    assert torch.equal(qpool, q64.max(-1).values.float()) and torch.equal(kpool, k64.max(-1).values.float())
    for a_, w_ in zip(got, (want_ctx.detach(), q64.grad, k64.grad, v64.grad)):
        assert rel(a_, w_) < TOL
    # deterministic: a second evaluation is bit-identical
    qp.grad = kp.grad = v.grad = None
    ctx2, _, _ = ATAmod.ata_core(qp, kp, v, dk)
    (ctx2 * gc).sum().backward()
    assert torch.equal(ctx2, got[0]) and torch.equal(qp.grad, got[1]) and torch.equal(kp.grad, got[2]) and torch.equal(v.grad, got[3])


@pytest.mark.gpu
def test_cached_head_equals_fresh_head_on_device(cuda, exact_convs):
    """The cached instance, called repeatedly, gives what a freshly constructed head gives (multi_head_attention.py:49-51
    builds one per forward): same context, same input gradients, same CUDA generator state afterwards."""
    b, h, l, dk, seed = 4, 8, 24, 4, 321
    g = torch.Generator().manual_seed(1)
    Q, K, V = (torch.randn(b, l, h, dk, generator=g).transpose(1, 2).to(cuda).requires_grad_(True) for _ in range(3))
    gc = torch.randn(b, h, l, dk, generator=g).to(cuda)
    ATAmod.ATA._cache.clear()
    for it in range(3):
        fresh = ATAmod.ATA(d_k=dk, device=cuda, h=h, seed=seed)
        state_fresh = torch.cuda.get_rng_state(cuda).clone()
        c1, _ = fresh(Q=Q, K=K, V=V)
        g1 = torch.autograd.grad(c1, (Q, K, V), gc)
        torch.cuda.manual_seed(1000 + it)
        head = ATAmod.ATA.cached(d_k=dk, device=cuda, h=h, seed=seed)
        assert torch.equal(torch.cuda.get_rng_state(cuda), state_fresh)
        c2, _ = head(Q=Q, K=K, V=V)
        g2 = torch.autograd.grad(c2, (Q, K, V), gc)
        # (the cached head evaluates the four stacks of a side as ONE 9-tap convolution: same arithmetic up to the
        # summation order inside cuDNN)
        assert rel(c2, c1) < TOL and all(rel(b_, a) < 10 * TOL for a, b_ in zip(g1, g2))
        assert getattr(head, "_fused", False)


@pytest.mark.gpu
@pytest.mark.parametrize("b,h,l,lk,dk", [(3, 8, 24, 40, 4), (2, 4, 7, 12, 8), (256, 8, 192, 192, 4)])
def test_fused_stacks_layout_equals_cat_layout(cuda, b, h, l, lk, dk):
    """The core on ONE [b, 4 C, l] convolution output (filter stacks as channel groups) is bit-identical to the core on
    the reference's torch.cat(dim=0) + reshape of the four stacks, forward and backward."""
    g = torch.Generator(device=cuda).manual_seed(l)
    C, nf = h * dk, 4
    yq = torch.relu(torch.randn(b, nf * C, l, device=cuda, generator=g)).requires_grad_(True)
    yk = torch.relu(torch.randn(b, nf * C, lk, device=cuda, generator=g)).requires_grad_(True)
    v = torch.randn(b, lk, h, dk, device=cuda, generator=g).transpose(1, 2).requires_grad_(True)
    gc = torch.randn(b, h, l, dk, device=cuda, generator=g)
    c1 = ATAmod.ata_core_fused_stacks(yq, yk, v, dk, nf)[0]
    g1 = torch.autograd.grad(c1, (yq, yk, v), gc)

    def cat_view(y, L):      # what ATA.py:50-59 builds: stacks [b, C, L] -> cat over the batch axis -> [b, h, L, 4 d_k]
        return torch.cat([y[:, i * C:(i + 1) * C, :] for i in range(nf)], dim=0).reshape(b, h, L, -1)

    c2 = ATAmod.ata_core(cat_view(yq, l), cat_view(yk, lk), v, dk)[0]
    g2 = torch.autograd.grad(c2, (yq, yk, v), gc)
    assert torch.equal(c1, c2) and all(torch.equal(a, b_) for a, b_ in zip(g1, g2))

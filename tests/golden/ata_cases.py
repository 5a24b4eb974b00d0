"""Seeded cases for the ATA head parity tests (shared by the golden generator and the tests)."""
import torch

CASES = {
    # name: (b, h, l, l_k, d_k, seed)      d_v = d_k as in multi_head_attention.py (d_k = d_v = d_model / n_heads)
    "self_small": (3, 8, 24, 24, 4, 1234),     # decoder self-attention shape of the reference default (d_model 32, 8 heads)
    "cross": (2, 8, 24, 40, 4, 77),            # decoder -> encoder cross attention (l != l_k)
    "odd": (2, 2, 12, 12, 6, 5),               # d_v = 6 (padded to 8 lanes in the kernel), G = 24
}


def make_inputs(name):
    b, h, l, lk, dk, seed = CASES[name]
    g = torch.Generator().manual_seed(seed + 1)
    # q_s / k_s / v_s of multi_head_attention.py:44-46 are [b, l, h, d] buffers viewed as [b, h, l, d]
    Q = torch.randn(b, l, h, dk, generator=g).transpose(1, 2)
    K = torch.randn(b, lk, h, dk, generator=g).transpose(1, 2)
    V = torch.randn(b, lk, h, dk, generator=g).transpose(1, 2)
    Gc = torch.randn(b, h, l, dk, generator=g)
    return Q, K, V, Gc

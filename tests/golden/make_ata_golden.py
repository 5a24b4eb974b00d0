"""Generates tests/golden/ata_ref_*.npz with the UNMODIFIED reference head (/root/reference/forecasting_models/ATA.py,
CPU, fp32; only possible in the build container):

    python tests/golden/make_ata_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ata_cases import CASES, make_inputs  # noqa: E402

sys.path.insert(0, "/root/reference")
from forecasting_models.ATA import ATA  # noqa: E402

for name, (b, h, l, lk, dk, seed) in CASES.items():
    Q, K, V, Gc = make_inputs(name)
    Q, K, V = (t.clone().requires_grad_(True) for t in (Q, K, V))
    head = ATA(d_k=dk, device="cpu", h=h, seed=seed)
    context, attn = head(Q=Q, K=K, V=V)
    (context * Gc).sum().backward()
    np.savez_compressed(os.path.join(HERE, f"ata_ref_{name}.npz"), context=context.detach().numpy(),
                        attn=attn.detach().numpy(), gQ=Q.grad.numpy(), gK=K.grad.numpy(), gV=V.grad.numpy())
    print(name, tuple(context.shape), float(context.abs().max()))

"""CPU oracle for the ATA attention head.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import
this module; the product (``fine_grained_gaussian_process_forcasting_b200/ATA.py`` + ``csrc/gpblur_ata.cu``) never
routes through it.

PARITY PINNED: the head is plain torch code of the reference itself.  ``tests/golden/make_ata_golden.py`` runs the
UNMODIFIED ``forecasting_models/ATA.py`` from /root/reference (CPU, fp32) in the build container and stores context,
attention and input gradients for seeded inputs (``tests/golden/ata_ref_*.npz``); ``tests/test_ata.py`` checks this
restatement against them.

Restated (functional form, explicit weights) from /root/reference/forecasting_models/ATA.py:
* constructor: seeding and parameter order ........ :8-38
* reshape (memory re-interpretation, no transpose)  :47-48
* Conv1d + BatchNorm1d (batch statistics) + ReLU .. :19-33, :50-51
* cat over the batch axis + reshapes + top-1 ...... :53-60
* rank-one scores / softmax / context ............. :62-65
"""
from __future__ import annotations

import math
import random

import numpy as np
import torch
import torch.nn.functional as F

FILTERS = (1, 3, 7, 9)


def init_weights(d_k: int, h: int, seed: int):
    """Parameters in the order the reference constructor draws them (conv_list_k, conv_list_q, proj_back_q/k)."""
    torch.manual_seed(seed)
    random.seed(seed)
    np.random.seed(seed)
    C = d_k * h
    w = {}
    for side in ("k", "q"):
        for f in FILTERS:
            conv = torch.nn.Conv1d(C, C, f, padding=int((f - 1) / 2))
            w[f"{side}{f}"] = (conv.weight.detach().clone(), conv.bias.detach().clone())
    return w


def multiscale(x: torch.Tensor, weights, side: str) -> torch.Tensor:
    """[b, C, l] -> cat over the batch axis of relu(batch_norm(conv_f(x))) for the four filter lengths: [4 b, C, l]."""
    outs = []
    for f in FILTERS:
        wt, bs = weights[f"{side}{f}"]
        y = F.conv1d(x, wt.to(x), bs.to(x), padding=int((f - 1) / 2))
        y = F.batch_norm(y, None, None, None, None, training=True, momentum=0.1, eps=1e-5)   # fresh BN: weight 1, bias 0
        outs.append(torch.relu(y))
    return torch.cat(outs, dim=0)


def core(q_proj: torch.Tensor, k_proj: torch.Tensor, V: torch.Tensor, d_k: int):
    """ATA.py:56-65 on the pooled projections: (context, attn)."""
    Q = q_proj.max(dim=-1, keepdim=True).values          # topk(k = 1)
    K = k_proj.max(dim=-1, keepdim=True).values
    scores = torch.einsum("bhqd,bhkd->bhqk", Q, K) / math.sqrt(d_k)
    attn = torch.softmax(scores, -1)
    return torch.einsum("bhqk,bhkd->bhqd", attn, V), attn


def ata_forward(Q: torch.Tensor, K: torch.Tensor, V: torch.Tensor, weights, d_k: int):
    b, h, l, _ = Q.shape
    l_k = K.shape[2]
    q_proj = multiscale(Q.reshape(b, -1, l), weights, "q").reshape(b, h, l, -1)
    k_proj = multiscale(K.reshape(b, -1, l_k), weights, "k").reshape(b, h, l_k, -1)
    return core(q_proj, k_proj, V, d_k)

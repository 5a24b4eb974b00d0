"""ATA attention head with a fused core.

Mirror of ``/root/reference/forecasting_models/ATA.py`` (class name, constructor ``(d_k, device, h, seed)``, the seeding
side effects and the order in which the sub-modules draw their initial weights, ``forward(Q, K, V)``), as
``modules/multi_head_attention.py:49-51`` instantiates and calls it.  The multi-scale ``Conv1d + BatchNorm1d + ReLU``
stacks stay library calls (cuDNN through torch - a different, GEMM-shaped workload); everything after them -

    Q_proj = Q_p.reshape(b, h, l, -1);  Q, _ = torch.topk(Q_proj, dim=-1, k=1)            ATA.py:56-60
    scores = einsum('bhqd,bhkd->bhqk', Q, K) / sqrt(d_k);  attn = softmax(scores, -1)      ATA.py:62-64
    context = einsum('bhqk,bhkd->bhqd', attn, V)                                           ATA.py:65

- is ONE CUDA kernel forward and ONE backward (``csrc/gpblur_ata.cu``): after the top-1 pooling the scores are rank
one, so ``scores`` / ``attn`` ``[b, h, l, l_k]`` never exist in HBM.  ``forward`` returns ``(context, attn)`` like the
reference; ``attn`` is ``None`` unless ``need_attn=True`` (keyword extension: the only caller discards it,
multi_head_attention.py:50, 95-97), in which case it is rebuilt from the pooled vectors, detached.

No CPU fallback: the core raises on non-CUDA tensors (CPU restatement: oracle/ata_oracle.py).
"""
from __future__ import annotations

import math
import random

import numpy as np
import torch
import torch.nn as nn

from . import _cabi
from .ops import _f32c, _need_cuda, _ptr, _stream


class _AtaCoreFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qp, kp, v, scale):
        _need_cuda(qp, kp, v)
        qp, kp = _f32c(qp), _f32c(kp)
        B, H, Lq, G = qp.shape
        Lk, DV = kp.shape[2], v.shape[-1]
        if kp.shape != (B, H, Lk, G) or v.shape != (B, H, Lk, DV):
            raise ValueError(f"ata_core: shapes {tuple(qp.shape)} {tuple(kp.shape)} {tuple(v.shape)}")
        if v.dtype != torch.float32 or v.stride(-1) != 1:
            v = _f32c(v)
        dev = qp.device
        out = torch.empty(B, Lq, H, DV, device=dev, dtype=torch.float32)
        q_pool = torch.empty(B, H, Lq, device=dev, dtype=torch.float32)
        k_pool = torch.empty(B, H, Lk, device=dev, dtype=torch.float32)
        q_arg = torch.empty(B, H, Lq, device=dev, dtype=torch.int32)
        k_arg = torch.empty(B, H, Lk, device=dev, dtype=torch.int32)
        lse = torch.empty(B, H, Lq, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            rc = _cabi.lib().gpblur_ata_forward(_ptr(qp), _ptr(kp), _ptr(v), v.stride(0), v.stride(1), v.stride(2), B, H,
                                                Lq, Lk, G, DV, float(scale), _ptr(out), _ptr(q_pool), _ptr(k_pool),
                                                _ptr(q_arg), _ptr(k_arg), _ptr(lse), _stream())
        _cabi.check(rc, "gpblur_ata_forward")
        ctx.save_for_backward(out, v, q_pool, k_pool, q_arg, k_arg, lse)
        ctx.dims = (B, H, Lq, Lk, G, DV, float(scale))
        ctx.mark_non_differentiable(q_pool, k_pool)
        # [b, h, l, dv] as the reference returns it - a view of the [b, l, h, dv] buffer, so that the caller's
        # context.transpose(1, 2).contiguous() (multi_head_attention.py:95) is free
        return out.transpose(1, 2), q_pool, k_pool

    @staticmethod
    def backward(ctx, g_context, _gq, _gk):
        out, v, q_pool, k_pool, q_arg, k_arg, lse = ctx.saved_tensors
        B, H, Lq, Lk, G, DV, scale = ctx.dims
        dev = out.device
        g = _f32c(g_context.transpose(1, 2))                      # [B, Lq, H, DV]
        g_qp = torch.empty(B, H, Lq, G, device=dev, dtype=torch.float32)
        g_kp = torch.empty(B, H, Lk, G, device=dev, dtype=torch.float32)
        g_v = torch.empty(B, Lk, H, DV, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            rc = _cabi.lib().gpblur_ata_backward(_ptr(g), _ptr(out), _ptr(v), v.stride(0), v.stride(1), v.stride(2),
                                                 _ptr(q_pool), _ptr(k_pool), _ptr(q_arg), _ptr(k_arg), _ptr(lse), B, H, Lq,
                                                 Lk, G, DV, scale, _ptr(g_qp), _ptr(g_kp), _ptr(g_v), _stream())
        _cabi.check(rc, "gpblur_ata_backward")
        return g_qp, g_kp, g_v.transpose(1, 2), None


def ata_core(q_proj: torch.Tensor, k_proj: torch.Tensor, v: torch.Tensor, d_k: int):
    """(context [b, h, l, d_v], q_pool [b, h, l], k_pool [b, h, l_k]) of ATA.py:56-65 for ``Q_proj [b, h, l, G]``,
    ``K_proj [b, h, l_k, G]`` and ``V [b, h, l_k, d_v]``."""
    return _AtaCoreFunction.apply(q_proj, k_proj, v, 1.0 / math.sqrt(d_k))


class ATA(nn.Module):
    def __init__(self, d_k, device, h, seed):
        super(ATA, self).__init__()
        torch.manual_seed(seed)          # ATA.py:12-14: the head is re-created (and re-seeded) on every forward
        random.seed(seed)
        np.random.seed(seed)
        self.d_k = d_k
        self.filter_length = [1, 3, 7, 9]

        def stack():                     # ATA.py:19-33, parameter draws in the reference's order (k first, then q)
            return nn.ModuleList([
                nn.Sequential(nn.Conv1d(in_channels=d_k * h, out_channels=d_k * h, kernel_size=f, padding=int((f - 1) / 2),
                                        device=device),
                              nn.BatchNorm1d(d_k * h, device=device),
                              nn.ReLU())
                for f in self.filter_length])

        self.conv_list_k = stack()
        self.conv_list_q = stack().to(device)
        self.proj_back_q = nn.Linear(d_k * len(self.filter_length), self.d_k, device=device)   # unused by forward, as in
        self.proj_back_k = nn.Linear(d_k * len(self.filter_length), self.d_k, device=device)   # the reference (ATA.py:35-36)
        self.factor = 1

    def forward(self, Q, K, V, *, need_attn: bool = False):
        b, h, l, d_k = Q.shape
        l_k = K.shape[2]
        Q = Q.reshape(b, -1, l)          # ATA.py:47-48: a re-interpretation of the [b, h, l, d_k] memory, not a transpose
        K = K.reshape(b, -1, l_k)
        n = len(self.filter_length)
        Q_l = [self.conv_list_q[i](Q) for i in range(n)]
        K_l = [self.conv_list_k[i](K) for i in range(n)]
        # ATA.py:53-59: cat over the batch axis, then two reshapes of contiguous memory = one reshape
        Q_proj = torch.cat(Q_l, dim=0).reshape(b, h, l, -1)
        K_proj = torch.cat(K_l, dim=0).reshape(b, h, l_k, -1)
        context, q_pool, k_pool = ata_core(Q_proj, K_proj, V, self.d_k)
        attn = None
        if need_attn:
            attn = torch.softmax(q_pool.unsqueeze(-1) * k_pool.unsqueeze(-2) / np.sqrt(self.d_k), -1)
        return context, attn

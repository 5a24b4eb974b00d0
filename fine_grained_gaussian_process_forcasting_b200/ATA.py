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
import torch.nn.functional as F

from . import _cabi
from .ops import _f32c, _need_cuda, _ptr, _stream


class _AtaCoreFunction(torch.autograd.Function):
    """nf = 0: qp [B, H, Lq, G], kp [B, H, Lk, G] (the reference's cat + reshape).  nf > 0: qp [B, nf * C, Lq],
    kp [B, nf * C, Lk] - the nf filter stacks as channel groups of one convolution output (C = H * d_k, G = nf * d_k);
    the kernel reads every group where the cat would have put it, so the result is the same."""

    @staticmethod
    def forward(ctx, qp, kp, v, scale, nf, H):
        _need_cuda(qp, kp, v)
        qp, kp = _f32c(qp), _f32c(kp)
        B = qp.shape[0]
        if nf:
            Lq, Lk = qp.shape[2], kp.shape[2]
            G = qp.shape[1] // H                         # nf * d_k
            ok = qp.dim() == 3 and kp.shape[:2] == qp.shape[:2] and qp.shape[1] % (nf * H) == 0
        else:
            _, H, Lq, G = qp.shape
            Lk = kp.shape[2]
            ok = kp.shape == (B, H, Lk, G)
        DV = v.shape[-1]
        if not ok or v.shape != (B, H, Lk, DV):
            raise ValueError(f"ata_core: shapes {tuple(qp.shape)} {tuple(kp.shape)} {tuple(v.shape)} (nf = {nf})")
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
            rc = _cabi.lib().gpblur_ata_forward_fused_stacks(
                _ptr(qp), _ptr(kp), _ptr(v), v.stride(0), v.stride(1), v.stride(2), B, H, Lq, Lk, G, DV, int(nf),
                float(scale), _ptr(out), _ptr(q_pool), _ptr(k_pool), _ptr(q_arg), _ptr(k_arg), _ptr(lse), _stream())
        _cabi.check(rc, "gpblur_ata_forward")
        ctx.save_for_backward(out, v, q_pool, k_pool, q_arg, k_arg, lse)
        ctx.dims = (B, H, Lq, Lk, G, DV, float(scale), int(nf), qp.shape, kp.shape)
        ctx.mark_non_differentiable(q_pool, k_pool)
        # [b, h, l, dv] as the reference returns it - a view of the [b, l, h, dv] buffer, so that the caller's
        # context.transpose(1, 2).contiguous() (multi_head_attention.py:95) is free
        return out.transpose(1, 2), q_pool, k_pool

    @staticmethod
    def backward(ctx, g_context, _gq, _gk):
        out, v, q_pool, k_pool, q_arg, k_arg, lse = ctx.saved_tensors
        B, H, Lq, Lk, G, DV, scale, nf, q_shape, k_shape = ctx.dims
        dev = out.device
        g = _f32c(g_context.transpose(1, 2))                      # [B, Lq, H, DV]
        g_qp = torch.empty(q_shape, device=dev, dtype=torch.float32)
        g_kp = torch.empty(k_shape, device=dev, dtype=torch.float32)
        g_v = torch.empty(B, Lk, H, DV, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            rc = _cabi.lib().gpblur_ata_backward_fused_stacks(
                _ptr(g), _ptr(out), _ptr(v), v.stride(0), v.stride(1), v.stride(2), _ptr(q_pool), _ptr(k_pool),
                _ptr(q_arg), _ptr(k_arg), _ptr(lse), B, H, Lq, Lk, G, DV, nf, scale, _ptr(g_qp), _ptr(g_kp), _ptr(g_v),
                _stream())
        _cabi.check(rc, "gpblur_ata_backward")
        return g_qp, g_kp, g_v.transpose(1, 2), None, None, None


def ata_core(q_proj: torch.Tensor, k_proj: torch.Tensor, v: torch.Tensor, d_k: int):
    """(context [b, h, l, d_v], q_pool [b, h, l], k_pool [b, h, l_k]) of ATA.py:56-65 for ``Q_proj [b, h, l, G]``,
    ``K_proj [b, h, l_k, G]`` and ``V [b, h, l_k, d_v]``."""
    return _AtaCoreFunction.apply(q_proj, k_proj, v, 1.0 / math.sqrt(d_k), 0, q_proj.shape[1])


def ata_core_fused_stacks(q_stacks: torch.Tensor, k_stacks: torch.Tensor, v: torch.Tensor, d_k: int, n_filters: int):
    """The same for the filter stacks as channel groups of ONE convolution output: ``q_stacks [b, n_filters * h * d_k,
    l]`` (stack i in channels [i C, (i + 1) C)) instead of ``torch.cat([stack_0, ...], dim=0)``."""
    return _AtaCoreFunction.apply(q_stacks, k_stacks, v, 1.0 / math.sqrt(d_k), int(n_filters), v.shape[1])


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

    # ---- the head is RE-CREATED on every forward by its only caller (multi_head_attention.py:49-51): sixteen modules
    # built and initialised on the device, each time with the same seed, i.e. the same weights ----
    _cache = {}

    @classmethod
    def cached(cls, d_k, device, h, seed):
        """The module ``ATA(d_k, device, h, seed)`` would construct, built once per (d_k, device, h, seed) and handed
        out again - with the constructor's side effects replayed: Python / numpy / torch generators are left exactly as
        a fresh construction leaves them (seeded, the torch generators advanced past the weight draws).  The cached
        instance behaves like a fresh one on every call: it stays in training mode (a module created inside ``forward``
        always is), its batch-norm running statistics never move (momentum 0) and its weights take no gradient (the
        reference computes their gradients and throws them away with the module)."""
        dev = torch.device(device)
        key = (int(d_k), str(dev), int(h), int(seed))
        ent = cls._cache.get(key)
        if ent is None:
            mod = cls(d_k, device, h, seed)
            for m in mod.modules():
                if isinstance(m, nn.BatchNorm1d):
                    m.momentum = 0.0
            for prm in mod.parameters():
                prm.requires_grad_(False)
            mod._fuse_stacks()
            cuda_state = torch.cuda.get_rng_state(dev) if dev.type == "cuda" else None
            cls._cache[key] = (mod, torch.get_rng_state(), cuda_state)
            return mod
        mod, cpu_state, cuda_state = ent
        torch.manual_seed(seed)          # every generator re-seeded, as ATA.py:12 does
        random.seed(seed)
        np.random.seed(seed)
        torch.set_rng_state(cpu_state)   # ... and advanced past the weight initialisation
        if cuda_state is not None:
            torch.cuda.set_rng_state(cuda_state, dev)
        return mod

    def _fuse_stacks(self):
        """Frozen weights (``cached``): the four `same`-padded convolutions of a side become ONE 9-tap convolution with
        4 C output channels (shorter filters centred, zero taps around them) and the four fresh batch norms one
        affine-free batch norm over 4 C channels - per-channel statistics, so nothing mixes."""
        kmax = max(self.filter_length)
        for side, stacks in (("q", self.conv_list_q), ("k", self.conv_list_k)):
            C = stacks[0][0].weight.shape[0]
            w = torch.zeros(len(stacks) * C, C, kmax, device=stacks[0][0].weight.device)
            for i, f in enumerate(self.filter_length):
                o = (kmax - f) // 2
                w[i * C:(i + 1) * C, :, o:o + f] = stacks[i][0].weight.detach()
            self.register_buffer(f"_w_{side}", w, persistent=False)
            self.register_buffer(f"_b_{side}", torch.cat([st[0].bias.detach() for st in stacks]), persistent=False)
        self._fused = True

    def _stacks(self, x, side):
        y = F.conv1d(x, getattr(self, f"_w_{side}"), getattr(self, f"_b_{side}"), padding=(max(self.filter_length) - 1) // 2)
        return torch.relu(F.batch_norm(y, None, None, None, None, True, 0.0, 1e-5))

    def forward(self, Q, K, V, *, need_attn: bool = False):
        b, h, l, d_k = Q.shape
        l_k = K.shape[2]
        n = len(self.filter_length)
        if getattr(self, "_fused", False) and (h * l) % n == 0 and (h * l_k) % n == 0:
            # one convolution + one batch norm + one ReLU per side, no cat: the core reads the stacks in place
            context, q_pool, k_pool = ata_core_fused_stacks(self._stacks(Q.reshape(b, -1, l), "q"),
                                                            self._stacks(K.reshape(b, -1, l_k), "k"), V, self.d_k, n)
            attn = None
            if need_attn:
                attn = torch.softmax(q_pool.unsqueeze(-1) * k_pool.unsqueeze(-2) / np.sqrt(self.d_k), -1)
            return context, attn
        Q = Q.reshape(b, -1, l)          # ATA.py:47-48: a re-interpretation of the [b, h, l, d_k] memory, not a transpose
        K = K.reshape(b, -1, l_k)
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

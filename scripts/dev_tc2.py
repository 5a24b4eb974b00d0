"""Bring-up / A-B timing of the TS-form point kernels against the fp64 oracle.
usage: python scripts/dev_tc2.py [fwd|bwd|all] (GPU box only)"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import gp_oracle as O  # noqa: E402
from fine_grained_gaussian_process_forcasting_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def rel(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-300)).item()


def params(D, M, seed=11):
    p32 = O.init_params_exercise(D, M, seed=seed)
    pd = {k: v.to(dev) for k, v in p32.items()}
    return p32, pd


def stage_of(pd, D):
    return ops.param_stage_raw(pd["inducing_points"], pd["raw_lengthscale"].reshape(-1), pd["raw_outputscale"].reshape(1),
                               pd["variational_mean"], pd["variational_stddev"], pd["weights"].reshape(-1), pd["bias"])


def check_fwd(B, L, D, M):
    p32, pd = params(D, M)
    x32, _, _, _ = O.make_inputs(B, L, D, seed=12)
    p64 = O.clone_params(p32, torch.float64)
    mean_o, var_o = O.svgp_predict_closed_form(p64, x32.double())
    ell = O.softplus(p64["raw_lengthscale"]).reshape(D)
    os_ = O.softplus(p64["raw_outputscale"])
    Z = p64["inducing_points"]
    Kzz = O.rbf_scale_direct(Z, Z, ell, os_) + O.JITTER * torch.eye(M, dtype=torch.float64)
    Kzx = O.rbf_scale_direct(Z, x32.double().reshape(-1, D), ell, os_)
    A_o = torch.linalg.solve_triangular(torch.linalg.cholesky(Kzz), Kzx, upper=False).t()
    stage, kl, info = stage_of(pd, D)
    xd = x32.to(dev).reshape(-1, D)
    os.environ["GPBLUR_TC_V"] = "2"
    mean, var, sample, ws = ops.point_forward_raw(stage, xd, M, 1234, 5, 0, True, True)
    torch.cuda.synchronize()
    A = ops.debug_fetch(3, B * L, D, M, ws)[:, :M]
    print(f"fwd v2 B={B} L={L} D={D} M={M}: mean {rel(mean.reshape(B, L), mean_o):.2e} var {rel(var.reshape(B, L), var_o):.2e} "
          f"A {rel(A, A_o):.2e}", flush=True)


def time_fwd(B, L, D, M, iters=20):
    p32, pd = params(D, M)
    stage, kl, info = stage_of(pd, D)
    x = torch.randn(B * L, D, device=dev)
    N = B * L
    ws = torch.empty(ops.workspace_bytes(N, D, M, True), device=dev, dtype=torch.uint8)
    outb = torch.empty(3 * N, device=dev)
    res = {}
    for v in ("2", "1"):
        os.environ["GPBLUR_TC_V"] = v
        for _ in range(3):
            ops.point_forward_raw(stage, x, M, 1, 0, 0, True, True, out=outb, ws=ws)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            ops.point_forward_raw(stage, x, M, 1, 0, 0, True, True, out=outb, ws=ws)
        e1.record()
        torch.cuda.synchronize()
        res[v] = e0.elapsed_time(e1) / iters
    # includes the per-call D2D copy of the parameter stage
    print(f"time fwd B={B} L={L} D={D} M={M}: new {res['2']*1e3:.1f} us  old {res['1']*1e3:.1f} us", flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "fwd"
    if what in ("fwd", "all"):
        for shp in [(2, 24, 64, 128), (8, 192, 32, 256), (16, 24, 64, 128), (3, 24, 32, 300), (2, 24, 64, 1024), (8, 24, 10, 256),
                    (1100, 1, 128, 256), (300, 24, 64, 256)]:
            check_fwd(*shp)
        for shp in [(8192, 24, 64, 256), (256, 192, 64, 256), (8192, 24, 64, 128), (8192, 24, 64, 512), (8192, 24, 64, 1024)]:
            time_fwd(*shp)

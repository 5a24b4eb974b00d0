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
    for v in ("2",):
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
    print(f"time fwd B={B} L={L} D={D} M={M}: {res['2']*1e3:.1f} us", flush=True)


def check_bwd(B, L, D, M):
    p32 = O.init_params_exercise(D, M, seed=21)
    x32, y32, gm32, gv32 = O.make_inputs(B, L, D, seed=22)
    gs32 = torch.randn(B, L, generator=torch.Generator().manual_seed(23))
    gkl = 0.37
    seed, offset, stream = 99, 1000, 2
    p64 = O.clone_params(p32, torch.float64, requires_grad=True)
    x64 = x32.double().requires_grad_(True)
    mo, vo = O.svgp_predict_closed_form(p64, x64)
    eps = torch.from_numpy(O.philox_normal(seed, offset, B * L, stream)).double().reshape(B, L)
    so = O.rsample(mo, vo, eps)
    loss = (gm32.double() * mo).sum() + (gv32.double() * vo).sum() + (gs32.double() * so).sum() + gkl * O.kl_meanfield(p64)
    loss.backward()
    pd = {k: v.to(dev).requires_grad_(True) for k, v in p32.items()}
    xd = x32.to(dev).requires_grad_(True)
    mean, var, sample, kl, info = ops.svgp_predict(xd, pd["inducing_points"], pd["raw_lengthscale"], pd["raw_outputscale"],
                                                   pd["variational_mean"], pd["variational_stddev"], pd["weights"], pd["bias"],
                                                   seed=seed, offset=offset, stream_id=stream, want_sample=True)
    lc = (gm32.to(dev) * mean).sum() + (gv32.to(dev) * var).sum() + (gs32.to(dev) * sample).sum() + gkl * kl
    lc.backward()
    torch.cuda.synchronize()
    errs = {"dx": rel(xd.grad, x64.grad)}
    for k in ["inducing_points", "raw_lengthscale", "raw_outputscale", "variational_mean", "variational_stddev", "weights", "bias"]:
        errs[k[:8]] = rel(pd[k].grad, p64[k].grad)
    print(f"bwd B={B} L={L} D={D} M={M}: " + " ".join(f"{k}={v:.1e}" for k, v in errs.items()), flush=True)


def time_bwd(B, L, D, M, iters=10):
    from fine_grained_gaussian_process_forcasting_b200 import _cabi
    p32, pd = params(D, M)
    pdg = {k: v.clone().requires_grad_(True) for k, v in pd.items()}
    x = torch.randn(B, L, D, device=dev, requires_grad=True)
    gm = torch.randn(B, L, device=dev)

    def step():
        mean, var, sample, kl, info = ops.svgp_predict(x, pdg["inducing_points"], pdg["raw_lengthscale"], pdg["raw_outputscale"],
                                                       pdg["variational_mean"], pdg["variational_stddev"], pdg["weights"],
                                                       pdg["bias"], seed=1, offset=0, stream_id=0, want_sample=True)
        torch.autograd.backward([mean, sample], [gm, gm])
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    _cabi.profile_enable(True)
    for _ in range(iters):
        step()
    torch.cuda.synchronize()
    prof = _cabi.profile_collect()
    _cabi.profile_enable(False)
    print(f"stages B={B} L={L} D={D} M={M}: " + " ".join(f"{k}={ms / max(cnt, 1) * 1e3:.1f}us" for k, (ms, cnt) in prof.items() if cnt), flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "fwd"
    if what in ("fwd", "all"):
        for shp in [(2, 24, 64, 128), (8, 192, 32, 256), (16, 24, 64, 128), (3, 24, 32, 300), (2, 24, 64, 1024), (8, 24, 10, 256),
                    (1100, 1, 128, 256), (300, 24, 64, 256)]:
            check_fwd(*shp)
        for shp in [(8192, 24, 64, 256), (256, 192, 64, 256), (8192, 24, 64, 128), (8192, 24, 64, 512), (8192, 24, 64, 1024)]:
            time_fwd(*shp)
    if what == "stages":                       # python scripts/dev_tc2.py stages B L D M [iters]
        shp = tuple(int(v) for v in sys.argv[2:6])
        time_bwd(*shp, iters=int(sys.argv[6]) if len(sys.argv) > 6 else 200)
        sys.exit(0)
    if what == "wdbg":
        B, L, D, M = (int(v) for v in sys.argv[2:6])
        p32 = O.init_params_exercise(D, M, seed=21)
        x32, y32, gm32, gv32 = O.make_inputs(B, L, D, seed=22)
        pd = {k: v.to(dev) for k, v in p32.items()}
        N = B * L
        xd = x32.to(dev).reshape(N, D)
        args = (pd["inducing_points"], pd["raw_lengthscale"].reshape(-1), pd["raw_outputscale"].reshape(1),
                pd["variational_mean"], pd["variational_stddev"], pd["weights"].reshape(-1), pd["bias"])
        mean, var, sample, kl, info, ws = ops.svgp_forward_raw(xd, *args, 1, 0, 0, True, True)
        gm = gm32.to(dev).reshape(N).contiguous(); gv = gv32.to(dev).reshape(N).contiguous()
        gkl = torch.zeros(1, device=dev)
        dx, bucket = ops.svgp_backward_raw(xd, *args, gm, gv, None, gkl, var, 1, 0, 0, ws)
        torch.cuda.synchronize()
        W = ops.debug_fetch(5, N, D, M, ws)[:, :M].double().cpu()
        p64 = O.clone_params(p32, torch.float64)
        ell = O.softplus(p64["raw_lengthscale"]).reshape(D); os_ = O.softplus(p64["raw_outputscale"])
        Z = p64["inducing_points"]
        Kzz = O.rbf_scale_direct(Z, Z, ell, os_) + O.JITTER * torch.eye(M, dtype=torch.float64)
        K = O.rbf_scale_direct(x32.double().reshape(N, D), Z, ell, os_)          # [N, M]
        Lc = torch.linalg.cholesky(Kzz)
        Linv = torch.linalg.solve_triangular(Lc, torch.eye(M, dtype=torch.float64), upper=False)
        Aa = K @ Linv.t()
        cvec = p64["variational_stddev"] ** 2 - 1
        beta = Linv.t() @ p64["variational_mean"]
        vo = O.svgp_predict_closed_form(p64, x32.double())[1].reshape(N)
        gvv = gv32.double().reshape(N).clone(); gvv[vo <= 1e-6] = 0
        kbar = gm32.double().reshape(N, 1) * beta.reshape(1, M) + 2 * gvv.reshape(N, 1) * ((Aa * cvec) @ Linv)
        Wo = kbar * K
        err = (W - Wo).abs()
        print("W rel err", (err.max() / Wo.abs().max()).item())
        for c in range(M // 32):
            blk = err[:, c * 32:(c + 1) * 32]
            print(f"  chunk {c}: max err {blk.max().item():.3e} (ref max {Wo[:, c*32:(c+1)*32].abs().max().item():.3e})  "
                  f"rows with err>1e-4: {(blk.max(dim=1).values > 1e-4).sum().item()} / {N}")
        sys.exit(0)
    if what in ("bwd", "all"):
        for shp in [(16, 24, 64, 128), (16, 24, 32, 128), (8, 24, 64, 256), (8, 192, 32, 256), (8, 24, 16, 256), (4, 24, 64, 512),
                    (2, 24, 64, 1024), (5, 13, 48, 100), (3, 24, 32, 300), (8, 24, 10, 256), (40, 24, 20, 128), (1100, 1, 128, 256),
                    (700, 24, 64, 256)]:
            check_bwd(*shp)
        for shp in [(8192, 24, 64, 256), (256, 192, 64, 256), (8192, 24, 64, 128), (8192, 24, 64, 1024)]:
            time_bwd(*shp)

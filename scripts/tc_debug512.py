import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fine_grained_gaussian_process_forcasting_b200 import ops
from oracle import gp_oracle as O
dev = torch.device("cuda")
D = 64
for M in (256, 512, 1024):
    for N in (96, 1000):
        p = {k: v.to(dev) for k, v in O.init_params_exercise(D, M, 1).items()}
        x = torch.randn(N, D, device=dev)
        res = {}
        for mode in ("tc_train", "tc_infer", "ffma"):
            os.environ["GPBLUR_TC"] = "-1" if mode == "ffma" else "0"
            out = ops.svgp_forward_raw(x, p["inducing_points"], p["raw_lengthscale"].reshape(-1), p["raw_outputscale"].reshape(1),
                                       p["variational_mean"], p["variational_stddev"], p["weights"].reshape(-1), p["bias"], 0, 0, 0, False,
                                       mode != "tc_infer")
            torch.cuda.synchronize()
            res[mode] = (out[0].clone(), out[1].clone())
        os.environ["GPBLUR_TC"] = "0"
        for mode in ("tc_train", "tc_infer"):
            em = (res[mode][0] - res["ffma"][0]).abs().max().item() / res["ffma"][0].abs().max().item()
            ev = (res[mode][1] - res["ffma"][1]).abs().max().item() / res["ffma"][1].abs().max().item()
            bad = ((res[mode][0] - res["ffma"][0]).abs() > 1e-3).nonzero().flatten()[:8].tolist()
            print(f"M={M} N={N} {mode}: mean {em:.2e} var {ev:.2e} bad rows {bad}")

"""M x M forward stage cold (256 MiB L2 flush before every launch, as bench.py does between steps) vs warm
(back-to-back), event-timed per launch, with the in-kernel phase stamps of the last cold launch.
usage: python scripts/mm_cold_probe.py [M]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from fine_grained_gaussian_process_forcasting_b200 import ops
from oracle import gp_oracle as O
dev = torch.device("cuda")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 256
D = 64
p = {k: v.to(dev) for k, v in O.init_params_exercise(D, M, 1).items()}
args = (p["inducing_points"], p["raw_lengthscale"].reshape(-1), p["raw_outputscale"].reshape(1), p["variational_mean"],
        p["variational_stddev"], p["weights"].reshape(-1), p["bias"])
flush = torch.empty(64 << 20, device=dev)
stage = torch.empty(ops.param_stage_bytes(D, M), device=dev, dtype=torch.uint8)
kl = torch.empty(1, device=dev); info = torch.empty(1, device=dev, dtype=torch.int32)
for _ in range(10):
    ops.param_stage_raw(*args, out=(stage, kl, info))
for mode in ("warm", "cold"):
    ts = []
    for _ in range(30):
        if mode == "cold":
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.param_stage_raw(*args, out=(stage, kl, info)); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    t = ops.debug_fetch(4, 64, D, M, stage).cpu().tolist() if hasattr(ops, "debug_fetch") else None
    line = f"M={M} {mode}: median {np.median(ts):.1f} us  min {min(ts):.1f}"
    if t:
        f = [(t[i + 1] - t[i]) / 1e3 for i in range(0, 7)]
        line += f"  in-kernel phases {[round(v, 1) for v in f]} total {sum(f):.1f}"
    print(line)

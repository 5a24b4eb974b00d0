"""Event-timed M x M parameter stage alone (back-to-back launches, steady clocks).
usage: python scripts/mm_only.py [M] [iters]    (GPBLUR_MM_STOP=k: the forward kernel returns after phase k)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fine_grained_gaussian_process_forcasting_b200 import ops
from oracle import gp_oracle as O
dev = torch.device("cuda")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 300
D = 64
p = {k: v.to(dev) for k, v in O.init_params_exercise(D, M, 1).items()}
args = (p["inducing_points"], p["raw_lengthscale"].reshape(-1), p["raw_outputscale"].reshape(1), p["variational_mean"],
        p["variational_stddev"], p["weights"].reshape(-1), p["bias"])
for _ in range(20):
    ops.param_stage_raw(*args)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    ops.param_stage_raw(*args)
e1.record()
torch.cuda.synchronize()
print(f"M={M} stop={os.environ.get('GPBLUR_MM_STOP', '-')}: param stage {e0.elapsed_time(e1) / iters * 1e3:.1f} us per launch", flush=True)

import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fine_grained_gaussian_process_forcasting_b200 import ops
from oracle import gp_oracle as O
dev = torch.device("cuda")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 256
D, N = 64, 256
p = {k: v.to(dev) for k, v in O.init_params_exercise(D, M, 1).items()}
x = torch.randn(N, D, device=dev)
for it in range(4):
    out = ops.svgp_forward_raw(x, p["inducing_points"], p["raw_lengthscale"].reshape(-1), p["raw_outputscale"].reshape(1),
                               p["variational_mean"], p["variational_stddev"], p["weights"].reshape(-1), p["bias"], 0, 0, 0, False, True)
torch.cuda.synchronize()
print("ok")

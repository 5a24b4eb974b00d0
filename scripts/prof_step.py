"""One GP-blur call (forward + backward) for ncu: python scripts/prof_step.py B L D M [iters]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fine_grained_gaussian_process_forcasting_b200 import ops
from oracle import gp_oracle as O
dev = torch.device("cuda:0")
B, L, D, M = (int(v) for v in sys.argv[1:5])
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 3
p32 = O.init_params_exercise(D, M, seed=11)
pdg = {k: v.to(dev).requires_grad_(True) for k, v in p32.items()}
x = torch.randn(B, L, D, device=dev, requires_grad=True)
gm = torch.randn(B, L, device=dev)
for _ in range(iters):
    mean, var, sample, kl, info = ops.svgp_predict(x, pdg["inducing_points"], pdg["raw_lengthscale"], pdg["raw_outputscale"],
                                                   pdg["variational_mean"], pdg["variational_stddev"], pdg["weights"],
                                                   pdg["bias"], seed=1, offset=0, stream_id=0, want_sample=True)
    torch.autograd.backward([mean, sample], [gm, gm])
torch.cuda.synchronize()
print("done")

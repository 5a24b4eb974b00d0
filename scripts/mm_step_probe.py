"""Per-block-column timestamps of the forward M x M kernel (GPBLUR_MM_STOP=9 probe): start of step, end of the diagonal
factorisation.  usage: GPBLUR_MM_STOP=9 python scripts/mm_step_probe.py [M]"""
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
for it in range(100):
    out = ops.svgp_forward_raw(x, p["inducing_points"], p["raw_lengthscale"].reshape(-1), p["raw_outputscale"].reshape(1),
                               p["variational_mean"], p["variational_stddev"], p["weights"].reshape(-1), p["bias"], 0, 0, 0, False, True)
torch.cuda.synchronize()
t = ops.debug_fetch(4, N, D, M, out[-1]).cpu().tolist()
if os.environ.get("GPBLUR_MM_STOP") == "8":
    names = ["T->F (factor | operand loads)", "publish stores", "product 1 + park + barrier", "product 2 + epilogue", "-> next T"]
    for i, n in enumerate(names):
        print(f"step 2: {n}: {(t[17 + i] - t[16 + i]) / 1e3:.2f} us")
    sys.exit(0)
nb = min(8, M // 32)
for kb in range(nb):
    s0, s1 = t[16 + 2 * kb], t[17 + 2 * kb]
    nxt = t[16 + 2 * kb + 2] if kb + 1 < nb else t[3]
    print(f"kb={kb}: load+factor+inverse {(s1 - s0) / 1e3:.2f} us, items + grid barrier {(nxt - s1) / 1e3:.2f} us")

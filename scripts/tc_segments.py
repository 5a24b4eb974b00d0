import sys, os
os.environ["GPBLUR_TC_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fine_grained_gaussian_process_forcasting_b200 import ops
from oracle import gp_oracle as O
dev = torch.device("cuda")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 256
D, N = 64, 8192 * 24
p = {k: v.to(dev) for k, v in O.init_params_exercise(D, M, 1).items()}
x = torch.randn(N, D, device=dev)
for it in range(3):
    out = ops.svgp_forward_raw(x, p["inducing_points"], p["raw_lengthscale"].reshape(-1), p["raw_outputscale"].reshape(1),
                               p["variational_mean"], p["variational_stddev"], p["weights"].reshape(-1), p["bias"], 0, 0, 0, True, True)
torch.cuda.synchronize()
t = ops.debug_fetch(4, N, D, M, out[-1]).cpu().tolist()
names = ["phaseA", "drain", "tmem+exp", "acquire", "split+store", "commit(sync/issue)", "block epilogue", "tile head"]
ntiles = (N + 127) // 128
per_cta = (ntiles + 147) // 148
for who, off in (("thread 0 (issuer)", 0), ("thread 32", 8)):
    tot = sum(t[off:off + 8])
    print(who, "total cycles", tot, "per tile", tot // per_cta)
    for i, nm in enumerate(names):
        print(f"   {nm:22s} {t[off + i]:10d}  {100 * t[off + i] / max(tot, 1):5.1f}%  per tile {t[off + i] // per_cta}")


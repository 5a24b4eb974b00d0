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
args = (p["inducing_points"], p["raw_lengthscale"].reshape(-1), p["raw_outputscale"].reshape(1), p["variational_mean"],
        p["variational_stddev"], p["weights"].reshape(-1), p["bias"])
x = torch.randn(N, D, device=dev)
g = torch.randn(N, device=dev)
for it in range(3):
    mean, var, sample, kl, info, ws = ops.svgp_forward_raw(x, *args, 0, 0, 0, True, True)
    dx, sgrad = ops.point_backward_raw(x, M, g, g, None, var, 0, 0, 0, ws)   # no M x M backward: it reuses the stamp slots
torch.cuda.synchronize()
t = ops.debug_fetch(4, N, D, M, ws).cpu().tolist()
names = ["tile head + first loads", "acquire (MMA s-2 retired)", "split + store", "next loads + commit", "drain", "epilogue"]
ntiles = (N + 127) // 128
per_cta = (ntiles + 147) // 148
tot = sum(t[24:30])
print("dx kernel, thread 0: total cycles", tot, "per tile", tot // per_cta)
for i, nm in enumerate(names):
    print(f"   {nm:28s} {t[24 + i]:10d}  {100 * t[24 + i] / max(tot, 1):5.1f}%  per tile {t[24 + i] // per_cta}")

bn = ["tile head", "phase A", "acquire (MMA s+2 retired)", "split + store (global A loads)", "next loads + commit", "interleaved epilogue", "drain", "final chunks + row sums"]
tot = sum(t[16:24])
print("bwd kernel, thread 0: total cycles", tot, "per tile", tot // per_cta)
for i, nm in enumerate(bn):
    print(f"   {nm:32s} {t[16 + i]:10d}  {100 * t[16 + i] / max(tot, 1):5.1f}%  per tile {t[16 + i] // per_cta}")

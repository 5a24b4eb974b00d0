"""Event trace of CTA 0 of the tensor-core dx kernel: issuer vs producer timestamps per pipeline slab (cycles)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fine_grained_gaussian_process_forcasting_b200 import ops
from oracle import gp_oracle as O
dev = torch.device("cuda")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 256
D, N = 64, 8192 * 24
trace = torch.zeros(2048, device=dev, dtype=torch.int64)
os.environ["GPBLUR_TRACE_PTR"] = hex(trace.data_ptr())
p = {k: v.to(dev) for k, v in O.init_params_exercise(D, M, 1).items()}
args = (p["inducing_points"], p["raw_lengthscale"].reshape(-1), p["raw_outputscale"].reshape(1), p["variational_mean"],
        p["variational_stddev"], p["weights"].reshape(-1), p["bias"])
x = torch.randn(N, D, device=dev)
g = torch.randn(N, device=dev)
for it in range(2):
    mean, var, sample, kl, info, ws = ops.svgp_forward_raw(x, *args, 0, 0, 0, True, True)
    dx, sgrad = ops.point_backward_raw(x, M, g, g, None, var, 0, 0, 0, ws)
torch.cuda.synchronize()
t = trace.cpu().tolist()
t0 = t[8 * 6]       # slab 8 = first slab of the second tile
print("slab | issuer: wait-start A-ready B-full issued committed | thread 0: acquire-start acquired stored arrived | thread 255: acquired arrived")
for sl in range(8, 32):
    i = [v - t0 for v in t[sl * 6: sl * 6 + 5]]
    p0 = [v - t0 for v in t[512 + sl * 4: 512 + sl * 4 + 4]]
    p1 = [v - t0 for v in t[768 + sl * 4: 768 + sl * 4 + 4]]
    print(f"{sl:3d} | {i[0]:7d} {i[1]:7d} {i[2]:7d} {i[3]:7d} {i[4]:7d} | {p0[0]:7d} {p0[1]:7d} {p0[2]:7d} {p0[3]:7d} | {p1[1]:7d} {p1[3]:7d}")

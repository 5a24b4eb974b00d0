"""Event trace of CTA 0 of the tensor-core forward point kernel: issuer vs producer timestamps per pipeline slab."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fine_grained_gaussian_process_forcasting_b200 import ops
from oracle import gp_oracle as O
dev = torch.device("cuda")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 256
D, N = 64, 8192 * 24
trace = torch.zeros(2048, device=dev, dtype=torch.int64)
os.environ["GPBLUR_FWD_TRACE_PTR"] = hex(trace.data_ptr())
p = {k: v.to(dev) for k, v in O.init_params_exercise(D, M, 1).items()}
args = (p["inducing_points"], p["raw_lengthscale"].reshape(-1), p["raw_outputscale"].reshape(1), p["variational_mean"],
        p["variational_stddev"], p["weights"].reshape(-1), p["bias"])
x = torch.randn(N, D, device=dev)
for it in range(2):
    out = ops.svgp_forward_raw(x, *args, 0, 0, 0, True, True)
torch.cuda.synchronize()
t = trace.cpu().tolist()
t0 = t[10 * 6]
print("slab | issuer: wait-start A-ready B-full issued loop-end late-request | owner lane 0: exp-done acquired stored committed")
for sl in range(10, 40):
    i = [v - t0 for v in t[sl * 6: sl * 6 + 5]] + [t[sl * 6 + 5]]
    p0 = [v - t0 if v else 0 for v in t[512 + sl * 4: 512 + sl * 4 + 4]]
    print(f"{sl:3d} | {i[0]:7d} {i[1]:7d} {i[2]:7d} {i[3]:7d} {i[4]:7d} {i[5]:2d} | {p0[0]:7d} {p0[1]:7d} {p0[2]:7d} {p0[3]:7d}")

import sys, os, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from fine_grained_gaussian_process_forcasting_b200.graphs import GraphedStep
from fine_grained_gaussian_process_forcasting_b200.distributed import FlatGradBucket, gp_parameters
name = sys.argv[1] if len(sys.argv) > 1 else "c1"
pre_eager = int(sys.argv[2]) if len(sys.argv) > 2 else 1
wl = bench.WORKLOADS[name]
dev = torch.device("cuda")
model = bench.make_model(wl, dev).train()
bucket = FlatGradBucket(gp_parameters(model))
B, D, calls = wl["B"], wl["D"], wl["calls"]
xs = [torch.randn(B, L, D, device=dev) for L in calls]
y = torch.randn(1, B, calls[-1], device=dev)
g_elbo = torch.full((1, B), -1.0 / B, device=dev)
gms = [torch.randn(1, B, L, device=dev) for L in calls]
def body(*ins):
    bucket.zero()
    outs, grads = [], []
    for c, L in enumerate(calls):
        x = ins[c].detach().requires_grad_(True)
        last = c == len(calls) - 1
        out = model.blur(x, ins[-1] if last else None, num_data=D)
        outs += [out.mean, out.sample]; grads += [gms[c], gms[c]]
        if last:
            outs.append(out.elbo); grads.append(g_elbo)
    torch.autograd.backward(outs, grads)
    return (out.elbo,)
for _ in range(pre_eager):
    body(*xs, y)
torch.cuda.synchronize()
try:
    g = GraphedStep(model, body, xs + [y])
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    print(name, "graph ok:", e0.elapsed_time(e1) / 50, "ms/replay")
except Exception:
    traceback.print_exc()

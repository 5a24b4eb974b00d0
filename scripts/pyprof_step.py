"""cProfile of the Python side of a training step through the module API (host overhead per step)."""
import cProfile, pstats, sys, os, io, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "c1"
wl = bench.WORKLOADS[name]
dev = torch.device("cuda")
model = bench.make_model(wl, dev).train()
B, D, calls = wl["B"], wl["D"], wl["calls"]
from fine_grained_gaussian_process_forcasting_b200.gpcompat import DeepGPLayer
layers = [m for m in model.modules() if isinstance(m, DeepGPLayer)]
xs = [torch.randn(B, L, D, device=dev) for L in calls]
y = torch.randn(1, B, calls[-1], device=dev)
g_elbo = torch.full((1, B), -1.0 / B, device=dev)
gms = [torch.randn(1, B, L, device=dev) for L in calls]

def step():
    for ly in layers:
        ly.invalidate_param_stage()
    outs, grads = [], []
    for c, L in enumerate(calls):
        x = xs[c].detach().requires_grad_(True)
        last = c == len(calls) - 1
        out = model.blur(x, y if last else None, num_data=D)
        outs += [out.mean, out.sample]; grads += [gms[c], gms[c]]
        if last:
            outs.append(out.elbo); grads.append(g_elbo)
    torch.autograd.backward(outs, grads)

for _ in range(20):
    step()
torch.cuda.synchronize()
n = 300
t0 = time.perf_counter()
for _ in range(n):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"{name}: host {1e3*(t1-t0)/n:.3f} ms/step issue time, {1e3*(t2-t0)/n:.3f} ms/step incl. drain")
pr = cProfile.Profile()
pr.enable()
for _ in range(n):
    step()
pr.disable()
torch.cuda.synchronize()
for key in ("tottime", "cumulative"):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats(key).print_stats(38)
    print(s.getvalue())

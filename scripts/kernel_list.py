"""Kernels of ONE eager step of a bench workload in launch order (torch.profiler / CUPTI): name, stream, duration, and the
CPU op that launched it.  usage: python scripts/kernel_list.py [workload]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from torch.profiler import profile, ProfilerActivity
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
wl = bench.WORKLOADS[name]
dev = torch.device("cuda")
model = bench.make_model(wl, dev).train()
B, D, calls = wl["B"], wl["D"], wl["calls"]
from fine_grained_gaussian_process_forcasting_b200.gpcompat import DeepGPLayer
from fine_grained_gaussian_process_forcasting_b200.distributed import FlatGradBucket
layers = [m for m in model.modules() if isinstance(m, DeepGPLayer)]
bucket = FlatGradBucket([p for p in model.parameters()], module=model)
xs = [torch.randn(B, L, D, device=dev) for L in calls]
y = torch.randn(1, B, calls[-1], device=dev)
g_elbo = torch.full((1, B), -1.0 / B, device=dev)
gms = [torch.randn(1, B, L, device=dev) for L in calls]

def step():
    bucket.zero()
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

for _ in range(5):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start if evs else 0
for e in evs:
    print(f"{e.time_range.start - t0:9.1f} us  +{e.time_range.end - e.time_range.start:7.1f} us  {e.name[:110]}")

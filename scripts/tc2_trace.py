"""Event trace of CTA 0 of the TS-form forward kernel (clock64 stamps).  usage: python scripts/tc2_trace.py [M] [B]"""
import os, sys
import ctypes as C
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fine_grained_gaussian_process_forcasting_b200 import ops, _cabi
from oracle import gp_oracle as O
dev = torch.device("cuda:0")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 256
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
D, L = 64, 24
p32 = O.init_params_exercise(D, M, seed=11)
pd = {k: v.to(dev) for k, v in p32.items()}
stage, kl, info = ops.param_stage_raw(pd["inducing_points"], pd["raw_lengthscale"].reshape(-1), pd["raw_outputscale"].reshape(1),
                                      pd["variational_mean"], pd["variational_stddev"], pd["weights"].reshape(-1), pd["bias"])
x = torch.randn(B * L, D, device=dev)
for _ in range(2):
    ops.point_forward_raw(stage, x, M, 1, 0, 0, True, True)
torch.cuda.synchronize()
buf = torch.zeros(4096, dtype=torch.int64, device=dev)
MODE = sys.argv[3] if len(sys.argv) > 3 else "fwd"
if MODE == "fwd":
    _cabi.lib().gpblur_debug_set_trace(C.c_void_p(buf.data_ptr()))
    ops.point_forward_raw(stage, x, M, 1, 0, 0, True, True)
    torch.cuda.synchronize()
    _cabi.lib().gpblur_debug_set_trace(None)
else:
    N = B * L
    mean, var, sample, ws = ops.point_forward_raw(stage, x, M, 1, 0, 0, True, True)
    gm = torch.randn(N, device=dev); gv = torch.randn(N, device=dev)
    for _ in range(2):
        ops.point_backward_raw(x, M, gm, gv, None, var, 1, 0, 0, ws, stage=stage)
    torch.cuda.synchronize()
    _cabi.lib().gpblur_debug_set_trace(C.c_void_p(buf.data_ptr()))
    ops.point_backward_raw(x, M, gm, gv, None, var, 1, 0, 0, ws, stage=stage)
    torch.cuda.synchronize()
    _cabi.lib().gpblur_debug_set_trace(None)
t = buf.cpu()
iss = t[:768].reshape(96, 8)
prod = t[1024:1024 + 768].reshape(96, 8)
t0 = int(iss[0, 0])
print("issuer: slab rows | wait_a_start a_ready b_full issued requested | d(issue) d(total)  [cycles rel. to slab 0 start]")
for g in range(64):
    r = iss[g]
    if r[0] == 0: break
    print(f"  {g:3d} N={int(r[5]):3d} | start {int(r[0])-t0:7d} ready {int(r[2])-t0:7d} issued {int(r[3])-t0:7d} | "
          f"wait {int(r[2]-r[0]):5d} issue {int(r[3]-r[2]):5d}")
print("producers (thread 0 of each group), whitening slabs: start ld_done exp_done acquired published")
for g in range(40):
    r = prod[g]
    if r[0] == 0: continue
    print(f"  {g:3d} | {int(r[0])-t0:7d} ld {int(r[1]-r[0]):5d} exp {int(r[2]-r[1]):5d} acq {int(r[3]-r[2]):5d} st+pub {int(r[4]-r[3]):5d}")
ep = t[3072:3072 + 384].reshape(2, 24, 8)
print("epilogue chunks: group, chunk | start, wait, tmem_ld, fma(+ldg), store")
for gg in range(2):
    for i in range(10):
        r = ep[gg, i]
        if r[0] == 0: break
        print(f"  g{gg} c{int(r[6])} | {int(r[0])-t0:7d} wait {int(r[1]-r[0]):5d} ld {int(r[2]-r[1]):5d} fma {int(r[3]-r[2]):5d} store {int(r[4]-r[3]):5d}")
sf = t[2048:2048 + 128].reshape(32, 4)
print("s_full waits per pass: group0 (start, dur) group1 (start, dur)")
for i in range(8):
    r = sf[i]
    print(f"  pass {i}: g0 {int(r[0])-t0:7d} +{int(r[1]-r[0]):5d}   g1 {int(r[2])-t0:7d} +{int(r[3]-r[2]):5d}")

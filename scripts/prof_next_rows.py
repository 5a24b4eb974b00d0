"""The SURVEY 8(f) kernels (window gather, blur application, loss assembly, ATA core) a few times each, for ncu:
python scripts/prof_next_rows.py [iters]"""
import math, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fine_grained_gaussian_process_forcasting_b200 import ATA as ata_mod, base_train as bt, step_ops
dev = torch.device("cuda:0")
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
g = torch.Generator(device=dev).manual_seed(1)
rs = np.random.RandomState(3)
rows, F, T, ne, pl, nwin = 1 << 20, 4, 240, 192, 24, 65536
ws = bt.WindowSet(rs.randn(rows, F).astype(np.float32), rs.randn(rows).astype(np.float32),
                  rs.randint(0, rows - T, size=nwin).astype(np.int64), T, ne, pl, device=dev)
N, D = 256 * (192 + 24), 64
x = torch.randn(N, D, device=dev, generator=g, requires_grad=True)
mean = torch.randn(N, device=dev, generator=g, requires_grad=True)
up = torch.nn.Linear(1, D).to(dev)
go = torch.randn(N, D, device=dev, generator=g)
hdec = torch.randn(8192, 48, D, device=dev, generator=g, requires_grad=True)
yt = torch.randn(8192, 24, 1, device=dev, generator=g)
elbo = torch.randn(8192, device=dev, generator=g, requires_grad=True)
lam = torch.full((1,), 0.003, device=dev, requires_grad=True)
fin = torch.nn.Linear(D, 1).to(dev)
b, h, l, dk = 256, 8, 192, 4
qp = torch.relu(torch.randn(b, h, l, 4 * dk, device=dev, generator=g)).requires_grad_(True)
kp = torch.relu(torch.randn(b, h, l, 4 * dk, device=dev, generator=g)).requires_grad_(True)
v = torch.randn(b, l, h, dk, device=dev, generator=g).transpose(1, 2).requires_grad_(True)
gc = torch.randn(b, h, l, dk, device=dev, generator=g)
for _ in range(iters):
    ws.gather(0, nwin)
    ws.gather(0, 256)
    o = step_ops.blur_apply(x, mean, up.weight, up.bias)
    torch.autograd.grad(o, (x, mean, up.weight, up.bias), go)
    f, loss, mse = step_ops.forecast_loss(fin, hdec[:, -24:, :], yt, elbo, lam)
    torch.autograd.grad(loss, (hdec, fin.weight, fin.bias, elbo, lam))
    c = ata_mod.ata_core(qp, kp, v, dk)[0]
    torch.autograd.grad(c, (qp, kp, v), gc)
torch.cuda.synchronize()
print("done")

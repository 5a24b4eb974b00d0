"""Debug helper: where do a single-shot run and a two-shard run of the forward differ (they must be bit-equal)?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fine_grained_gaussian_process_forcasting_b200 import ops
from oracle import gp_oracle as O
cuda = torch.device("cuda")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 128
B, L, D = 1024, 24, 64
p = {k: v.to(cuda) for k, v in O.init_params_exercise(D, M, 9).items()}
x = torch.randn(B, L, D, device=cuda, generator=torch.Generator(device=cuda).manual_seed(1))
def run(xs, off):
    return ops.svgp_predict(xs, p["inducing_points"], p["raw_lengthscale"], p["raw_outputscale"], p["variational_mean"],
                            p["variational_stddev"], p["weights"], p["bias"], seed=5, offset=off, stream_id=0, want_sample=True)
m_all, v_all, s_all, _, _ = run(x, 0)
m_all2, v_all2, _, _, _ = run(x, 0)
h = 384
m_a, v_a, s_a, _, _ = run(x[:h], 0)
m_b, v_b, s_b, _, _ = run(x[h:], h * L)
for name, full, parts in (("mean", m_all, torch.cat([m_a, m_b])), ("var", v_all, torch.cat([v_a, v_b])), ("rerun mean", m_all, m_all2)):
    d = (full - parts).abs().reshape(-1)
    bad = torch.nonzero(d > 0).reshape(-1)
    print(name, "max diff", d.max().item(), "n bad", bad.numel(), "of", d.numel())
    if bad.numel():
        tiles = torch.unique(bad // 128)
        print("   bad tiles:", tiles[:40].tolist(), "... rows in tile of first bad:", (bad[:16] % 128).tolist())

"""Generates tests/golden/svgp_*.npz: seeded inputs + float64 oracle outputs and gradients for small
instances of the BASELINE configs.  gpytorch (where the reference's arithmetic lives) is not installable
in this environment, so these vectors come from the oracle's closed-form restatement
(oracle/gp_oracle.py); they pin the oracle against drift and give the CUDA path a fixed target.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import gp_oracle as O  # noqa: E402

CASES = {
    # name: (B, L, D, M)  small-batch instances of C1, C2 (dec side), C3 and a ragged case
    "c1_small": (8, 24, 64, 32),
    "c2_dec_small": (4, 24, 32, 256),
    "c3_small": (4, 24, 64, 128),
    "ragged": (3, 7, 5, 3),
}


def make(name, B, L, D, M):
    p32 = O.init_params_exercise(D, M, seed=101)
    x32, y32, gm32, gv32 = O.make_inputs(B, L, D, seed=102)
    p = O.clone_params(p32, torch.float64, requires_grad=True)
    x = x32.double().requires_grad_(True)
    mean, var = O.svgp_predict_closed_form(p, x)
    kl = O.kl_meanfield(p)
    elbo = O.elbo_per_window(mean, var, y32.double(), O.noise_variance(p), kl, float(D))
    eps = torch.from_numpy(O.philox_normal(4321, 17, B * L, 0)).double().reshape(B, L)
    sample = O.rsample(mean, var, eps)
    loss = -elbo.mean() + (gm32.double() * mean).sum() + (gv32.double() * sample).sum()
    loss.backward()
    out = {"x": x32.numpy(), "y": y32.numpy(), "g_mean": gm32.numpy(), "g_sample": gv32.numpy(),
           "mean": mean.detach().numpy(), "var": var.detach().numpy(), "kl": kl.detach().numpy(),
           "elbo": elbo.detach().numpy(), "sample": sample.detach().numpy(), "loss": loss.detach().numpy(),
           "dx": x.grad.numpy(), "philox_seed_offset_stream": np.array([4321, 17, 0])}
    for k, v in p32.items():
        out["p_" + k] = v.numpy()
        out["d_" + k] = p[k].grad.numpy()
    np.savez_compressed(os.path.join(HERE, f"svgp_{name}.npz"), **out)


if __name__ == "__main__":
    for name, shp in CASES.items():
        make(name, *shp)
        print("wrote", name)
    r = O.philox_bits(1234, (1 << 33) + 5, 64, 3)
    np.save(os.path.join(HERE, "philox_bits_seed1234.npy"), r)

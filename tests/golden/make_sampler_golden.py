"""Generates tests/golden/sampler_ref_*.npz with the UNMODIFIED reference sampler
(/root/reference/Utils/base_train.py, imported from /root/reference - only possible in the build container; the vectors
travel, the reference does not):

    python tests/golden/make_sampler_golden.py

Each file holds the seeded synthetic frame (columns as arrays) and every batch of the three loaders
``batch_sampled_data`` returns for it."""
import contextlib
import io
import os
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from sampler_cases import CASES, column_definition, make_frame  # noqa: E402

sys.path.insert(0, "/root/reference")
from Utils import base_train  # noqa: E402
from Utils.base import DataTypes, InputTypes  # noqa: E402


def main():
    for name, c in CASES.items():
        df = make_frame(c)
        coldef = [(n, DataTypes(dt), InputTypes(it)) for n, dt, it in column_definition()]
        with contextlib.redirect_stdout(io.StringIO()):
            loaders = base_train.batch_sampled_data(df, c["train_percent"], c["max_samples"], c["time_steps"],
                                                    c["num_encoder_steps"], c["pred_len"], coldef, c["batch_size"])
        out = {}
        for split, loader in zip(("train", "valid", "test"), loaders):
            bs = list(loader)
            out[f"{split}_n"] = np.int64(len(bs))
            for i, (enc, dec, y) in enumerate(bs):
                out[f"{split}_{i}_enc"], out[f"{split}_{i}_dec"], out[f"{split}_{i}_y"] = enc.numpy(), dec.numpy(), y.numpy()
        np.savez_compressed(os.path.join(HERE, f"sampler_ref_{name}.npz"), **out)
        print(name, {k: int(out[k]) for k in out if k.endswith("_n")})


if __name__ == "__main__":
    main()

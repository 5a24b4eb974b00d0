"""Seeded synthetic frames for the window-sampler parity tests (shared by the golden generator and the tests)."""
import numpy as np
import pandas as pd

CASES = {
    # traffic-like: F = 4 input columns (16-byte rows: vector path), several entities, max_samples below the pool
    "traffic_like": dict(seed=11, entities=[90, 75, 120, 64], time_steps=48, num_encoder_steps=32, pred_len=8,
                         train_percent=0.8, max_samples=(40, 12), batch_size=8),
    # ragged: one entity too short for a window, max_samples ABOVE the number of valid locations of the small splits
    # (zero-filled tail), batch size that does not divide max_samples (drop_last)
    "ragged": dict(seed=12, entities=[30, 5, 41, 17, 26], time_steps=16, num_encoder_steps=8, pred_len=4,
                   train_percent=0.8, max_samples=(50, 30), batch_size=7),
    # no decoder steps between the encoder window and the horizon; shuffled input rows (the sort is part of the path)
    "no_dec": dict(seed=13, entities=[33, 29], time_steps=12, num_encoder_steps=9, pred_len=3,
                   train_percent=0.6, max_samples=(10, 6), batch_size=2, shuffle=True),
}


def column_definition():
    """(name, DataTypes value, InputTypes value) - the layout of /root/reference/data/traffic.py:26-33."""
    return [("id", 0, 4), ("hours_from_start", 0, 5), ("values", 0, 0), ("time_on_day", 0, 2), ("day_of_week", 0, 2),
            ("categorical_id", 1, 3)]


def make_frame(c):
    rng = np.random.RandomState(c["seed"])
    parts = []
    for e, n in enumerate(c["entities"]):
        t = np.arange(n, dtype=np.float64)
        parts.append(pd.DataFrame({
            "id": np.full(n, float(e * 3 + 1)), "hours_from_start": t, "values": rng.randn(n),
            "time_on_day": (t % 24) / 24.0 - 0.5 + 1e-9 * rng.randn(n), "day_of_week": np.floor(t / 24) % 7,
            "categorical_id": np.full(n, e, dtype=np.int64)}))
    df = pd.concat(parts, ignore_index=True)
    if c.get("shuffle"):
        df = df.sample(frac=1.0, random_state=c["seed"]).reset_index(drop=True)
    return df

"""CPU oracle for the window sampler.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import
this module; the product (``fine_grained_gaussian_process_forcasting_b200/base_train.py`` + ``gpblur_window_gather``)
never routes through it.

PARITY PINNED: unlike the GP arithmetic (gpytorch, not installable), this path is plain pandas / numpy code of the
reference itself.  ``tests/golden/make_sampler_golden.py`` imports the UNMODIFIED ``Utils/base_train.py`` from
/root/reference in the build container and stores its batches for seeded synthetic frames
(``tests/golden/sampler_ref_*.npz``); ``tests/test_sampler.py`` checks this restatement against those vectors
bit for bit.

Restated, statement by statement:
* ``sample_train_val_test`` ... /root/reference/Utils/base_train.py:29-97
* ``batch_sampled_data`` ...... /root/reference/Utils/base_train.py:100-153
* column lookup ............... /root/reference/Utils/utils.py:2-14, enums /root/reference/Utils/base.py
"""
from __future__ import annotations

import random

import numpy as np

ID, TIME, TARGET = 4, 5, 0        # InputTypes values, /root/reference/Utils/base.py


def single_col(input_type, column_definition):
    cols = [t[0] for t in column_definition if int(t[2]) == input_type]
    if len(cols) != 1:
        raise ValueError("Invalid number of columns for {}".format(input_type))
    return cols[0]


def sample_windows(ddf, max_samples, time_steps, num_encoder_steps, pred_len, column_definition):
    """base_train.py:29-97: every window is sliced out of its entity's frame and copied into float64 arrays."""
    id_col, time_col, target_col = (single_col(t, column_definition) for t in (ID, TIME, TARGET))
    enc_input_cols = [t[0] for t in column_definition if int(t[2]) not in (ID, TIME)]
    locations, frames = [], {}
    for identifier, df in ddf.groupby(id_col):                                   # :43-51
        if len(df) >= time_steps:
            locations += [(identifier, time_steps + i) for i in range(len(df) - time_steps + 1)]
            frames[identifier] = df
    if 0 < max_samples < len(locations):                                          # :53-63
        ranges = [locations[i] for i in np.random.choice(len(locations), max_samples, replace=False)]
    else:
        ranges = [locations[i] for i in np.random.choice(len(locations), len(locations), replace=False)]
    F = len(enc_input_cols)
    enc = np.zeros((max_samples, num_encoder_steps, F))                           # :65-70 (float64, zero-initialised)
    dec = np.zeros((max_samples, time_steps - num_encoder_steps - pred_len, F))
    inputs = np.zeros((max_samples, time_steps, F))
    outputs = np.zeros((max_samples, time_steps, 1))
    for i, (identifier, start_idx) in enumerate(ranges):                          # :72-80
        sliced = frames[identifier].iloc[start_idx - time_steps:start_idx]
        enc[i] = sliced[enc_input_cols].iloc[:num_encoder_steps]
        dec[i] = sliced[enc_input_cols].iloc[num_encoder_steps:-pred_len]
        inputs[i] = sliced[enc_input_cols]
        outputs[i] = sliced[[target_col]]
    return {"enc_inputs": enc.astype(np.float32), "dec_inputs": dec.astype(np.float32), "inputs": inputs.astype(np.float32),
            "outputs": outputs[:, -pred_len:, :].astype(np.float32),
            "input_arima": outputs[:, :-pred_len, :].astype(np.float32)}          # FloatTensor(...) rounds to fp32 (:136-146)


def batches(sample, batch_size):
    """DataLoader(TensorDataset(enc, dec, outputs), batch_size, drop_last=True) (base_train.py:136-151)."""
    n = sample["enc_inputs"].shape[0] // batch_size
    return [tuple(sample[k][i * batch_size:(i + 1) * batch_size] for k in ("enc_inputs", "dec_inputs", "outputs"))
            for i in range(n)]


def batch_sampled_data(data, train_percent, max_samples, time_steps, num_encoder_steps, pred_len, column_definition,
                       batch_size):
    """base_train.py:100-153."""
    np.random.seed(2436)
    random.seed(2436)
    data.sort_values(by=[single_col(ID, column_definition), single_col(TIME, column_definition)], inplace=True)
    train_len = int(len(data) * train_percent)
    valid_len = int((len(data) - train_len) / 2)
    splits = (data[:train_len], data[train_len:-valid_len], data)
    maxes = (max_samples[0], max_samples[1], max_samples[1])
    return [batches(sample_windows(s, m, time_steps, num_encoder_steps, pred_len, column_definition), batch_size)
            for s, m in zip(splits, maxes)]

"""Window sampler with a device-resident series table.

Mirror of the reference's ``Utils/base_train.py`` for the data path that feeds the GP blur: the same function names,
argument meaning, RNG consumption (``np.random.seed(2436)`` + three ``np.random.choice`` draws) and batch contents
(``enc_inputs``, ``dec_inputs``, ``outputs[:, -pred_len:, :]``, fp32, ``drop_last=True`` batches in sample order) -

* ``sample_train_val_test``  /root/reference/Utils/base_train.py:29-97
* ``batch_sampled_data``     /root/reference/Utils/base_train.py:100-153

but nothing is materialised per window on the host.  The reference slices every sampled window out of a pandas frame
(``.iloc`` per window, float64 arrays of [max_samples, time_steps, F]), builds ``TensorDataset``s on the host and ships
every batch with ``.to(device)`` (/root/reference/train.py:160-161).  Here each split keeps ONE fp32 table
``[rows, F]`` in HBM (entities back to back, uploaded once through pinned memory) plus the int64 start row of every
sampled window; a batch is ONE launch of ``gpblur_window_gather`` (csrc/gpblur_sampler.cu) on the consumer's stream.
The loaders yield CUDA tensors, so the caller's ``.to(self.device)`` is a no-op.

No CPU fallback: gathering raises without a CUDA device (the CPU restatement lives in oracle/sampler_oracle.py).
"""
from __future__ import annotations

import enum
import random
from typing import Iterator, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _cabi
from .ops import _ptr, _stream


class DataTypes(enum.IntEnum):
    """/root/reference/Utils/base.py: numerical type of a column (same values)."""
    REAL_VALUED = 0
    CATEGORICAL = 1
    DATE = 2


class InputTypes(enum.IntEnum):
    """/root/reference/Utils/base.py: input type of a column (same values; compares equal to the reference's enum)."""
    TARGET = 0
    OBSERVED_INPUT = 1
    KNOWN_INPUT = 2
    STATIC_INPUT = 3
    ID = 4
    TIME = 5


def get_single_col_by_input_type(input_type, column_definition):
    """/root/reference/Utils/utils.py:2-14 (same error)."""
    cols = [tup[0] for tup in column_definition if tup[2] == input_type]
    if len(cols) != 1:
        raise ValueError('Invalid number of columns for {}'.format(input_type))
    return cols[0]


class WindowSet:
    """The sampled windows of one split: host-side index (numpy) + device-side table, gathered on demand.

    ``starts[k]`` is the first row (in ``table``) of sample ``k``; ``-1`` marks the zero-filled tail the reference
    leaves when ``max_samples`` exceeds the number of valid sampling locations."""

    def __init__(self, table: np.ndarray, target: np.ndarray, starts: np.ndarray, time_steps: int,
                 num_encoder_steps: int, pred_len: int, time_values=None, id_values=None, device=None):
        self.table_host = np.ascontiguousarray(table, dtype=np.float32)
        self.target_host = np.ascontiguousarray(target, dtype=np.float32)
        self.starts_host = np.ascontiguousarray(starts, dtype=np.int64)
        self.time_steps, self.num_encoder_steps, self.pred_len = int(time_steps), int(num_encoder_steps), int(pred_len)
        self.num_decoder_steps = self.time_steps - self.num_encoder_steps - self.pred_len
        if self.num_decoder_steps < 0:
            raise ValueError("time_steps < num_encoder_steps + pred_len")
        self._time_values, self._id_values = time_values, id_values
        self.device = None
        self.table = self.target = self.starts = None
        if device is not None:
            self.to(device)

    # ---- host side -------------------------------------------------------------------------------------------
    def __len__(self) -> int:
        return int(self.starts_host.shape[0])

    @property
    def input_size(self) -> int:
        return int(self.table_host.shape[1])

    def object_column(self, which: str) -> np.ndarray:
        """'time' / 'identifier' [max_samples, time_steps, 1] object arrays (base_train.py:65-66, 79-80); host only."""
        vals = self._time_values if which == "time" else self._id_values
        out = np.empty((len(self), self.time_steps, 1), dtype=object)
        live = self.starts_host >= 0
        rows = self.starts_host[live][:, None] + np.arange(self.time_steps)[None, :]
        out[live, :, 0] = np.asarray(vals, dtype=object)[rows]
        return out

    # ---- device side -----------------------------------------------------------------------------------------
    def to(self, device) -> "WindowSet":
        """Upload the table / target / start rows once (pinned staging, asynchronous copies)."""
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("WindowSet gathers only on CUDA devices (B200 / sm_100a); there is no CPU fallback")

        def up(a):
            t = torch.from_numpy(a)
            if t.numel():
                t = t.pin_memory()
            return t.to(device, non_blocking=True)

        self.table, self.target, self.starts = up(self.table_host), up(self.target_host), up(self.starts_host)
        self.device = device
        return self

    def _gather(self, table, target, lo, hi, n_enc, pred_len):
        """One launch: (table rows [0, n_enc), table rows [n_enc, T - pred_len), target rows [T - pred_len, T)) of every
        window lo .. hi - 1."""
        if self.table is None:
            raise RuntimeError("WindowSet.gather: call .to(cuda_device) first (no CPU fallback)")
        lo, hi = max(0, int(lo)), min(len(self), int(hi))
        b, F, dev, T = max(0, hi - lo), int(table.shape[1]), self.device, self.time_steps
        enc = torch.empty(b, n_enc, F, device=dev, dtype=torch.float32)
        dec = torch.empty(b, T - n_enc - pred_len, F, device=dev, dtype=torch.float32)
        y = torch.empty(b, pred_len, 1, device=dev, dtype=torch.float32)
        if b:
            with torch.cuda.device(dev):
                rc = _cabi.lib().gpblur_window_gather(
                    _ptr(table), _ptr(target), table.shape[0], F, _ptr(self.starts[lo:hi]), b, T, n_enc, pred_len,
                    _ptr(enc) if enc.numel() else None, _ptr(dec) if dec.numel() else None,
                    _ptr(y) if y.numel() else None, _stream())
            _cabi.check(rc, "gpblur_window_gather")
        return enc, dec, y

    def gather(self, lo: int, hi: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """(enc_inputs [b, n_enc, F], dec_inputs [b, n_dec, F], outputs [b, pred_len, 1]) of samples lo .. hi - 1."""
        return self._gather(self.table, self.target, lo, hi, self.num_encoder_steps, self.pred_len)

    def __getitem__(self, key: str):
        """The reference's ``sampled_data`` dictionary (base_train.py:82-93), every entry gathered on the device."""
        n, T = len(self), self.time_steps
        if key in ("enc_inputs", "dec_inputs", "outputs"):
            return self.gather(0, n)[("enc_inputs", "dec_inputs", "outputs").index(key)]
        if key == "inputs":                      # the whole window of the input columns
            return self._gather(self.table, self.target, 0, n, T, 0)[0]
        if key == "input_arima":                 # outputs[:, :-pred_len, :]: the TARGET column before the horizon
            return self._gather(self.target.reshape(-1, 1), self.target, 0, n, T - self.pred_len, self.pred_len)[0]
        if key == "active_entries":              # np.ones_like(outputs[:, num_encoder_steps:, :])
            return torch.ones(n, T - self.num_encoder_steps, 1, device=self.device, dtype=torch.float32)
        if key in ("time", "identifier"):
            return self.object_column(key)
        raise KeyError(key)


def sample_train_val_test(ddf, max_samples, time_steps, num_encoder_steps, pred_len, column_definition, tgt_all=False,
                          device=None) -> WindowSet:
    """/root/reference/Utils/base_train.py:29-97 with the same sampling: entities in ``groupby(id)`` order, sampling
    locations ``(identifier, time_steps + i)``, ``np.random.choice(len, max_samples, replace=False)`` on numpy's GLOBAL
    generator (or a full permutation + the reference's message when ``max_samples`` is not below the number of
    locations).  Returns the window index instead of materialised arrays."""
    id_col = get_single_col_by_input_type(InputTypes.ID, column_definition)
    time_col = get_single_col_by_input_type(InputTypes.TIME, column_definition)
    target_col = get_single_col_by_input_type(InputTypes.TARGET, column_definition)
    enc_input_cols = [tup[0] for tup in column_definition if tup[2] not in {InputTypes.ID, InputTypes.TIME}]
    if max_samples <= 0:
        raise ValueError("max_samples must be positive")   # the reference indexes zero-row arrays (IndexError)

    tables, targets, times, ids, counts, offsets = [], [], [], [], [], []
    rows = 0
    for identifier, df in ddf.groupby(id_col):
        n = len(df)
        if n >= time_steps:
            # float64 -> float32 exactly as np.zeros(float64)[...] = frame; torch.FloatTensor(...) rounds
            tables.append(df[enc_input_cols].to_numpy(dtype=np.float64).astype(np.float32))
            targets.append(df[target_col].to_numpy(dtype=np.float64).astype(np.float32))
            times.append(df[time_col].to_numpy(dtype=object))
            ids.append(df[id_col].to_numpy(dtype=object))
            counts.append(n - time_steps + 1)
            offsets.append(rows)
            rows += n
    n_loc = int(sum(counts))
    if 0 < max_samples < n_loc:
        picks = np.random.choice(n_loc, max_samples, replace=False)
    else:
        print("maximum samples exceeds {}".format(n_loc))
        picks = np.random.choice(n_loc, n_loc, replace=False)      # n_loc <= max_samples here
    starts = np.full(max_samples, -1, dtype=np.int64)
    if n_loc:
        cum = np.cumsum(np.asarray(counts, dtype=np.int64))
        ent = np.searchsorted(cum, picks, side="right")
        within = picks - (cum[ent] - np.asarray(counts, dtype=np.int64)[ent])
        starts[:len(picks)] = np.asarray(offsets, dtype=np.int64)[ent] + within
    F = len(enc_input_cols)
    table = np.concatenate(tables, axis=0) if tables else np.zeros((0, F), np.float32)
    target = np.concatenate(targets, axis=0) if targets else np.zeros((0,), np.float32)
    tv = np.concatenate(times) if times else np.zeros((0,), object)
    iv = np.concatenate(ids) if ids else np.zeros((0,), object)
    return WindowSet(table, target, starts, time_steps, num_encoder_steps, pred_len, tv, iv, device=device)


class DeviceWindowLoader:
    """What ``torch.utils.data.DataLoader(TensorDataset(enc, dec, y), batch_size, drop_last=True)`` yields
    (base_train.py:149-151) - consecutive batches in sample order, the ragged tail dropped - gathered on the device.

    ``shard=(rank, world)`` (keyword, data-parallel training): every rank walks the SAME global
    batches and gathers only its contiguous slice of each (``distributed.shard_range``), so the union over the ranks is
    exactly the single-device batch and the Philox / window indices stay global."""

    def __init__(self, windows: WindowSet, batch_size: int, drop_last: bool = True, *, shard: Tuple[int, int] = (0, 1)):
        self.windows, self.batch_size, self.drop_last = windows, int(batch_size), bool(drop_last)
        rank, world = int(shard[0]), int(shard[1])
        if not 0 <= rank < world:
            raise ValueError(f"shard {shard}: need 0 <= rank < world")
        self.shard = (rank, world)

    def __len__(self) -> int:
        n = len(self.windows)
        return n // self.batch_size if self.drop_last else -(-n // self.batch_size)

    def batch_range(self, i: int) -> Tuple[int, int]:
        """[lo, hi) of this rank's windows in global batch i."""
        lo = i * self.batch_size
        hi = min(len(self.windows), lo + self.batch_size)
        rank, world = self.shard
        n = hi - lo
        base, extra = divmod(n, world)                    # the first `extra` ranks take one more window
        start = lo + rank * base + min(rank, extra)
        return start, start + base + (1 if rank < extra else 0)

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]:
        for i in range(len(self)):
            yield self.windows.gather(*self.batch_range(i))


def sampled_windows(data, train_percent, max_samples, time_steps, num_encoder_steps, pred_len, column_definition,
                    tgt_all=False, device=None):
    """The host half of ``batch_sampled_data`` (base_train.py:116-134): seeds (2436), the in-place (id, time) sort of
    the caller's frame, the row split (test = everything) and the three window indices.  ``device=None`` keeps them on
    the host (index only; gathering needs ``.to(cuda)``)."""
    np.random.seed(2436)
    random.seed(2436)
    time_col = get_single_col_by_input_type(InputTypes.TIME, column_definition)
    id_col = get_single_col_by_input_type(InputTypes.ID, column_definition)
    data.sort_values(by=[id_col, time_col], inplace=True)
    train_len = int(len(data) * train_percent)
    valid_len = int((len(data) - train_len) / 2)
    train = data[:train_len]
    valid = data[train_len:-valid_len]
    test = data
    train_max, valid_max = max_samples
    args = (time_steps, num_encoder_steps, pred_len, column_definition)
    sample_train = sample_train_val_test(train, train_max, *args, device=device)
    sample_valid = sample_train_val_test(valid, valid_max, *args, device=device)
    sample_test = sample_train_val_test(test, valid_max, *args, tgt_all, device=device)
    return sample_train, sample_valid, sample_test


def batch_sampled_data(data, train_percent, max_samples, time_steps, num_encoder_steps, pred_len, column_definition,
                       batch_size, tgt_all=False, device: Optional[torch.device] = None,
                       shard: Tuple[int, int] = (0, 1)):
    """/root/reference/Utils/base_train.py:100-153: three loaders (train, valid, test) of ``batch_size`` windows,
    ``drop_last=True``.  ``device`` (keyword, default: the current CUDA device) is where the tables live;
    ``shard=(rank, world)``: this rank's slice of every global batch (see DeviceWindowLoader)."""
    if device is None:
        if not torch.cuda.is_available():
            raise RuntimeError("batch_sampled_data: no CUDA device (B200 / sm_100a); there is no CPU fallback")
        device = torch.device("cuda", torch.cuda.current_device())
    sets = sampled_windows(data, train_percent, max_samples, time_steps, num_encoder_steps, pred_len, column_definition,
                           tgt_all, device=device)
    return tuple(DeviceWindowLoader(s, batch_size, shard=shard) for s in sets)

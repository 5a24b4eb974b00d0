"""Window sampler: oracle vs the vectors the UNMODIFIED reference produced
(tests/golden/sampler_ref_*.npz, generator tests/golden/make_sampler_golden.py), the product's host index vs the same
vectors (CPU), and the device gather vs both (GPU, bit-exact)."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from sampler_cases import CASES, column_definition, make_frame  # noqa: E402

from oracle import sampler_oracle as SO  # noqa: E402
from fine_grained_gaussian_process_forcasting_b200 import base_train as BT  # noqa: E402


def golden(name):
    z = np.load(os.path.join(HERE, "golden", f"sampler_ref_{name}.npz"))
    return {split: [(z[f"{split}_{i}_enc"], z[f"{split}_{i}_dec"], z[f"{split}_{i}_y"]) for i in range(int(z[f"{split}_n"]))]
            for split in ("train", "valid", "test")}


def args_of(c):
    return (c["train_percent"], c["max_samples"], c["time_steps"], c["num_encoder_steps"], c["pred_len"])


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def same(a, b):
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_vectors(name):
    c = CASES[name]
    got = quiet(SO.batch_sampled_data, make_frame(c), *args_of(c), column_definition(), c["batch_size"])
    want = golden(name)
    for split, bs in zip(("train", "valid", "test"), got):
        assert len(bs) == len(want[split]) and len(bs) > 0
        for g, w in zip(bs, want[split]):
            assert all(same(np.asarray(x), y) for x, y in zip(g, w))


def host_gather(ws, lo, hi):
    """numpy emulation of gpblur_window_gather on the host index (test-side only)."""
    T, ne, pl = ws.time_steps, ws.num_encoder_steps, ws.pred_len
    F = ws.input_size
    enc, dec = np.zeros((hi - lo, ne, F), np.float32), np.zeros((hi - lo, T - ne - pl, F), np.float32)
    y = np.zeros((hi - lo, pl, 1), np.float32)
    for k, s in enumerate(ws.starts_host[lo:hi]):
        if s >= 0:
            enc[k], dec[k] = ws.table_host[s:s + ne], ws.table_host[s + ne:s + T - pl]
            y[k, :, 0] = ws.target_host[s + T - pl:s + T]
    return enc, dec, y


@pytest.mark.parametrize("name", sorted(CASES))
def test_host_index_matches_reference_vectors(name):
    """The product's sampling (entity order, RNG draws, start rows, fp32 rounding, zero-filled tail, sort side effect)."""
    c = CASES[name]
    df = make_frame(c)
    sets = quiet(BT.sampled_windows, df, *args_of(c), column_definition())
    assert df["id"].is_monotonic_increasing                       # the caller's frame is sorted in place, as the reference does
    want = golden(name)
    bsz = c["batch_size"]
    for split, ws in zip(("train", "valid", "test"), sets):
        assert len(ws) // bsz == len(want[split])
        for i, w in enumerate(want[split]):
            assert all(same(g, y) for g, y in zip(host_gather(ws, i * bsz, (i + 1) * bsz), w))


def test_host_contract():
    coldef = column_definition()
    with pytest.raises(ValueError):
        BT.get_single_col_by_input_type(BT.InputTypes.KNOWN_INPUT, coldef)        # two KNOWN_INPUT columns
    assert BT.get_single_col_by_input_type(BT.InputTypes.TARGET, coldef) == "values"
    c = CASES["ragged"]
    ws = quiet(BT.sampled_windows, make_frame(c), *args_of(c), coldef)[1]
    assert (ws.starts_host < 0).any()                             # valid split: fewer locations than max_samples
    with pytest.raises(RuntimeError):
        ws.gather(0, 4)                                           # no CPU fallback
    with pytest.raises(RuntimeError):
        ws.to("cpu")
    ident = ws["identifier"]
    assert ident.shape == (len(ws), c["time_steps"], 1) and ident.dtype == object


def test_loader_shards_partition_every_batch():
    """Data-parallel loaders: the ranks' slices of a global batch are disjoint, contiguous, in rank order and agree with
    distributed.shard_range."""
    from fine_grained_gaussian_process_forcasting_b200.distributed import shard_range
    c = CASES["ragged"]
    ws = quiet(BT.sampled_windows, make_frame(c), *args_of(c), column_definition())[0]
    for world in (1, 2, 3, 8):
        loaders = [BT.DeviceWindowLoader(ws, c["batch_size"], shard=(r, world)) for r in range(world)]
        for i in range(len(loaders[0])):
            pos = i * c["batch_size"]
            for r, ld in enumerate(loaders):
                lo, hi = ld.batch_range(i)
                st, cnt = shard_range(c["batch_size"], r, world)
                assert (lo, hi) == (i * c["batch_size"] + st, i * c["batch_size"] + st + cnt) and lo == pos
                pos = hi
            assert pos == (i + 1) * c["batch_size"]
    with pytest.raises(ValueError):
        BT.DeviceWindowLoader(ws, 4, shard=(2, 2))


# ------------------------------------------------------------------------------------------------------------------
@pytest.fixture
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda:0")


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_device_loaders_match_reference_vectors(cuda, name):
    c = CASES[name]
    loaders = quiet(BT.batch_sampled_data, make_frame(c), *args_of(c), column_definition(), c["batch_size"], device=cuda)
    want = golden(name)
    for split, loader in zip(("train", "valid", "test"), loaders):
        assert len(loader) == len(want[split])
        for (enc, dec, y), w in zip(loader, want[split]):
            assert enc.is_cuda and enc.to(cuda) is enc            # train.py:160-161 `.to(self.device)` is a no-op
            assert all(same(g.cpu().numpy(), v) for g, v in zip((enc, dec, y), w))
        again = list(loader)                                      # loaders are re-iterable, one pass per epoch
        assert len(again) == len(want[split])


@pytest.mark.gpu
def test_device_dictionary_entries_match_oracle(cuda):
    c = CASES["ragged"]
    coldef = column_definition()
    np.random.seed(5)
    ws = quiet(BT.sample_train_val_test, make_frame(c), 25, c["time_steps"], c["num_encoder_steps"], c["pred_len"], coldef,
               device=cuda)
    np.random.seed(5)
    want = SO.sample_windows(make_frame(c), 25, c["time_steps"], c["num_encoder_steps"], c["pred_len"], coldef)
    for k in ("enc_inputs", "dec_inputs", "outputs", "inputs", "input_arima"):
        assert same(ws[k].cpu().numpy(), want[k]), k
    assert ws["active_entries"].shape == (25, c["time_steps"] - c["num_encoder_steps"], 1)


@pytest.mark.gpu
@pytest.mark.parametrize("F,B,T,ne,pl", [(3, 17, 20, 11, 4), (5, 33, 9, 9, 0), (4, 64, 40, 24, 8), (8, 1, 3, 0, 3),
                                         (4, 2048, 216, 192, 24)])
def test_gather_kernel_bit_exact(cuda, F, B, T, ne, pl):
    """Scalar (F % 4 != 0) and 16-byte paths, dead windows (start < 0 or past the table), empty segments, and the
    traffic-shape batch (time_steps 216 = 192 encoder + 24 horizon, F = 4) at 8 x the reference batch size."""
    rng = np.random.RandomState(F * 1000 + B)
    rows = 5000
    table = rng.randn(rows, F).astype(np.float32)
    target = rng.randn(rows).astype(np.float32)
    starts = rng.randint(0, rows - T + 1, size=B).astype(np.int64)
    starts[::7] = -1
    if B > 3:
        starts[3] = rows - T + 1                                  # would run past the table: treated as dead
    ws = BT.WindowSet(table, target, starts, T, ne, pl, device=cuda)
    enc, dec, y = ws.gather(0, B)
    torch.cuda.synchronize()
    live = (starts >= 0) & (starts + T <= rows)
    idx = np.where(live, starts, 0)[:, None] + np.arange(T)[None, :]
    full = table[idx] * live[:, None, None]
    tfull = target[idx] * live[:, None]
    assert same(enc.cpu().numpy(), full[:, :ne].astype(np.float32))
    assert same(dec.cpu().numpy(), full[:, ne:T - pl].astype(np.float32))
    assert same(y.cpu().numpy()[:, :, 0], tfull[:, T - pl:].astype(np.float32))
    e0, d0, y0 = ws.gather(5, 5)                                  # empty batch
    assert e0.shape[0] == 0 and d0.shape[0] == 0 and y0.shape[0] == 0

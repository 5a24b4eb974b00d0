"""CUDA-graph capture of a whole GP-blur training step (forward + hand-written backward).

The small configurations of the path are launch-bound: one step of the reference's default shape is ~17 kernel
launches of 10-200 us each, and the host (Python autograd + ctypes) needs ~0.8 ms to issue them.  ``GraphedStep``
records the launches once and replays them with ONE ``cudaGraphLaunch`` per step:

* inputs live in static device buffers (``copy_`` new data into ``.inputs`` before ``replay()``);
* parameter gradients accumulate into the parameters' existing ``.grad`` tensors (use ``FlatGradBucket`` so that
  they are one flat buffer), outputs are static tensors returned by ``replay()``;
* the fused sampler stays fresh across replays: every GP layer gets a device-resident Philox offset word
  (``rng_offset_dev``) that the captured kernels add to their counters, and the graph itself bumps that word at
  the end of each replay by the number of counters the step consumed.

Nothing here is specific to benchmarking: any callable built from this package's modules can be captured.
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch

from .gpcompat import DeepGPLayer


class GraphedStep:
    def __init__(self, module: torch.nn.Module, step_fn: Callable[..., Sequence[torch.Tensor]],
                 example_inputs: Sequence[torch.Tensor], warmup: int = 3, world: int = 1, rank: int = 0):
        """``step_fn(*inputs)`` runs forward AND backward through ``module`` and returns the tensors the caller
        wants to read (e.g. the per-window ELBO); ``example_inputs`` fix the shapes.  The function must not
        synchronise with the host (no ``.item()``).  ``world`` / ``rank``: with batch-sharded data parallelism the
        Philox offset word starts at rank * (counters per step) and advances by world * (counters per step), so
        that the ranks draw disjoint counters."""
        dev = example_inputs[0].device
        self.module = module
        self.layers = [m for m in module.modules() if isinstance(m, DeepGPLayer)]
        self.inputs = [t.clone() for t in example_inputs]
        for ly in self.layers:                       # device-resident Philox offsets (see module docstring)
            ly.rng_offset_dev = torch.zeros(1, device=dev, dtype=torch.int64)
            ly._rng_offset = 0
        self._consumed = {}
        self._offsets = [ly.rng_offset_dev for ly in self.layers]   # the captured kernels read these words

        def run():
            for ly in self.layers:
                ly.invalidate_param_stage()          # parameters change between replays: never reuse a stage
                ly._rng_offset = 0
            outs = step_fn(*self.inputs)
            for ly in self.layers:                   # counters consumed by this step (host-side bookkeeping)
                ly.rng_offset_dev.add_(world * ly._rng_offset)
                self._consumed[id(ly)] = ly._rng_offset
            return outs

        for ly in self.layers:                       # drop autograd graphs of earlier eager steps: their AccumulateGrad
            ly.invalidate_param_stage()              # nodes are bound to the eager stream and would break the capture
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                run()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        # anomaly mode (the reference switches it on globally, forecast_denoising.py:11) checks every gradient for
        # NaN on the HOST, which cannot be captured: it is off during the capture and restored afterwards
        anomaly = torch.is_anomaly_enabled()
        check_nan = torch.is_anomaly_check_nan_enabled() if hasattr(torch, "is_anomaly_check_nan_enabled") else True
        if anomaly:
            torch.autograd.set_detect_anomaly(False)
        try:
            with torch.cuda.graph(self.graph):
                self.outputs = run()
        finally:
            if anomaly:
                torch.autograd.set_detect_anomaly(True, check_nan=check_nan)
        for ly in self.layers:
            ly.invalidate_param_stage()              # tensors of the capture pool must not leak into eager calls
            ly.rng_offset_dev.fill_(rank * self._consumed.get(id(ly), 0))

    def replay(self):
        self.graph.replay()
        return self.outputs

    def check_info(self):
        """Cholesky status of the last replay (the capture cannot synchronise on it): raises NotPSDError if any layer's
        Kzz + jitter was not positive definite.  Synchronises with the device."""
        from .ops import NotPSDError
        for ly in self.layers:
            if ly.last_info is not None and int(ly.last_info.max().item()) != 0:
                raise NotPSDError(f"Kzz + jitter is not positive definite (pivot {int(ly.last_info.max().item())}) in a "
                                  f"graph replay; re-run the step eagerly to get the jitter retries")

    def set_rng_offset(self, value: int):
        for ly in self.layers:
            ly.rng_offset_dev.fill_(int(value))


class CallStreams:
    """Issue the independent GP calls of one step (the reference makes two: encoder- and decoder-side,
    denoise_model_2.py:50-51) on separate CUDA streams, so that a call that cannot fill the GPU runs in the tail
    of the other one's persistent kernels.  Call 0 stays on the current stream; autograd replays every call's
    backward on the stream of its forward.  Works eagerly and inside a ``GraphedStep`` capture (fork / join events
    become graph dependencies).

        cs = CallStreams(device, 2)
        enc_out = cs.run(0, lambda: model.blur(x_enc))
        dec_out = cs.run(1, lambda: model.blur(x_dec, y))
        cs.join()                      # before the outputs are consumed on the current stream
        torch.autograd.backward(...); cs.join()
    """

    def __init__(self, device, n_calls: int):
        self.device = torch.device(device)
        self.side = [None] + [torch.cuda.Stream(device=self.device) for _ in range(max(0, n_calls - 1))]

    def run(self, i: int, fn):
        side = self.side[i]
        if side is None:
            return fn()
        main = torch.cuda.current_stream(self.device)
        side.wait_stream(main)
        torch.cuda.set_stream(side)
        try:
            return fn()
        finally:
            torch.cuda.set_stream(main)

    def join(self):
        main = torch.cuda.current_stream(self.device)
        for side in self.side[1:]:
            main.wait_stream(side)

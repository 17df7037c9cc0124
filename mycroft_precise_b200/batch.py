"""Many independent audio streams per call -- the throughput path (no reference equivalent: the
reference runs one Listener per process).  Semantically ``S`` x [Listener.update ->
TriggerDetector.update] per tick, all state device-resident.
"""
import numpy as np

from .core import PreciseB200
from .model_io import GruModel
from .params import ListenerParams


class StreamBatch:
    def __init__(self, model: GruModel, n_streams: int, params: ListenerParams = None, chunk_samples=1024,
                 device=0, sensitivity=0.5, trigger_level=3):
        self.pr = params or ListenerParams()
        self.n_streams = n_streams
        self.core = PreciseB200(self.pr, hidden=model.hidden, max_streams=n_streams, chunk_samples=chunk_samples,
                                device=device, sensitivity=sensitivity, trigger_level=trigger_level,
                                activation=model.activation, recurrent_activation=model.recurrent_activation)
        self.core.load_weights(model.kernel, model.recurrent, model.bias, model.dense_w, model.dense_b)
        torch = self.core.torch
        self.count = torch.zeros(1, dtype=torch.int64, device=self.core.device)
        self._out = None

    def update(self, pcm, ids=None):
        """pcm: int16 CUDA tensor [n, chunk_samples] -> dict(raw f32[n], conf f64[n], fired u8[n]).
        ``self.count`` (int64[1], device) accumulates fired streams until reset_count()."""
        n = pcm.shape[0]
        if self._out is None or self._out['conf'].shape[0] != n:
            torch = self.core.torch
            dev = self.core.device
            self._out = dict(raw=torch.empty(n, dtype=torch.float32, device=dev),
                             conf=torch.empty(n, dtype=torch.float64, device=dev),
                             fired=torch.empty(n, dtype=torch.uint8, device=dev))
        return self.core.update(pcm, ids, self._out, self.count)

    def update_host(self, pcm_np: np.ndarray, conf_np: np.ndarray, raw_np=None, fired_np=None) -> int:
        return self.core.update_host(pcm_np, conf_np, raw_np, fired_np)

    def reset_count(self):
        self.count.zero_()

    def clear(self, ids=None):
        self.core.clear(ids=ids) if ids is not None else self.core.clear()

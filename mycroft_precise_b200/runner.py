"""Drop-in mirrors of the reference's inference classes, backed by the CUDA library.

  B200Runner    ~ precise.network_runner.Runner subclasses (KerasRunner / TensorFlowRunner,
                  network_runner.py:31-95): predict(inputs[N,T,F]) -> [N,1] float32, run(inp) -> float.
                  Usable as ``Listener(model, chunk, runner_cls=B200Runner)`` with the reference's
                  own Listener.
  B200Listener  ~ precise.network_runner.Listener (network_runner.py:98-153): same constructor
                  shape, update / update_vectors / clear, same input polymorphism
                  (BinaryIO | ndarray | bytes) and EOFError on an empty read.  The carry buffer and
                  the 29-row window are plain host arrays exactly as in the reference; vectorize_raw,
                  runner.run and ThresholdDecoder.decode run on the GPU.
  B200Engine    ~ precise_runner.runner.Engine (runner.py:22-33) for PreciseRunner: get_prediction
                  (chunk: bytes) -> float through the stateful device path (pb_update_host).
  Engine, TriggerDetector: interface-identical copies of the tiny client-side classes
                  (runner.py:22-33, :115-142) so this package works without the reference installed.
"""
from os.path import splitext

import numpy as np

from .core import PreciseB200
from .model_io import GruModel, load_weights, MODEL_SUFFIXES
from .params import ListenerParams, load_params


class Engine(object):
    """runner/precise_runner/runner.py:22-33"""

    def __init__(self, chunk_size=2048):
        self.chunk_size = chunk_size

    def start(self):
        pass

    def stop(self):
        pass

    def get_prediction(self, chunk):
        raise NotImplementedError


class TriggerDetector:
    """Host copy of runner/precise_runner/runner.py:115-142 for single-stream clients; the batched
    path runs the same state machine on the device (gru_kernels.cuh: epilogue)."""

    def __init__(self, chunk_size, sensitivity=0.5, trigger_level=3):
        self.chunk_size = chunk_size
        self.sensitivity = sensitivity
        self.trigger_level = trigger_level
        self.activation = 0

    def update(self, prob):
        chunk_activated = prob > 1.0 - self.sensitivity
        if chunk_activated or self.activation < 0:
            self.activation += 1
            has_activated = self.activation > self.trigger_level
            if has_activated or chunk_activated and self.activation < 0:
                self.activation = -(8 * 2048) // self.chunk_size
            if has_activated:
                return True
        elif self.activation > 0:
            self.activation -= 1
        return False


def _resolve_model(model):
    """model: GruModel | path to a weights file.  Returns (GruModel, ListenerParams).

    Same suffix dispatch as Listener.find_runner (network_runner.py:111-119: '.net' Keras HDF5, '.pb' frozen
    GraphDef) plus this package's own '.npz'; the readers live in model_io.load_weights."""
    if isinstance(model, GruModel):
        return model, None
    ext = splitext(model)[-1]
    if ext not in MODEL_SUFFIXES:
        raise ValueError('File extension of ' + model + ' must be: ' + str(list(MODEL_SUFFIXES)))
    return load_weights(model), load_params(model)


class B200Runner:
    """``Runner`` plug-in (network_runner.py:31-42).  ``model_name`` is a weights file or GruModel."""

    def __init__(self, model_name, params: ListenerParams = None, device=0):
        model, pr = _resolve_model(model_name)
        self.pr = params or pr or ListenerParams()
        self.core = PreciseB200(self.pr, hidden=model.hidden, device=device, activation=model.activation,
                                recurrent_activation=model.recurrent_activation)
        if model.feature_size != self.core.feature_size:
            raise ValueError('model expects %d features, params give %d' % (model.feature_size, self.core.feature_size))
        self.core.load_weights(model.kernel, model.recurrent, model.bias, model.dense_w, model.dense_b)

    def predict(self, inputs: np.ndarray) -> np.ndarray:
        torch = self.core.torch
        x = torch.as_tensor(np.ascontiguousarray(inputs, dtype=np.float32)).to(self.core.device)
        return self.core.predict(x).cpu().numpy()[:, None]

    def run(self, inp: np.ndarray) -> float:
        return self.predict(inp[np.newaxis])[0][0]


class B200Listener:
    """``Listener`` mirror (network_runner.py:98-153)."""

    def __init__(self, model_name, chunk_size: int = -1, runner_cls: type = None, params: ListenerParams = None,
                 device=0):
        self.window_audio = np.array([], dtype=np.float32)
        model, pr = _resolve_model(model_name) if model_name is not None and model_name != '' else (None, None)
        self.pr = params or pr or ListenerParams()
        self.mfccs = np.zeros((self.pr.n_features, self.pr.n_mfcc))
        self.chunk_size = chunk_size
        if runner_cls is not None:
            self.runner = runner_cls(model_name)
            self.core = getattr(self.runner, 'core', None) or PreciseB200(self.pr, device=device)
        else:
            self.runner = B200Runner(model, self.pr, device)
            self.core = self.runner.core

    def clear(self):
        self.window_audio = np.array([], dtype=np.float32)
        self.mfccs = np.zeros((self.pr.n_features, self.pr.n_mfcc))

    def _vectorize_raw(self, audio: np.ndarray) -> np.ndarray:
        if len(audio) == 0:
            raise ValueError('Cannot vectorize empty audio!')
        torch = self.core.torch
        a = torch.as_tensor(np.ascontiguousarray(audio, dtype=np.float32)).to(self.core.device)
        return self.core.mfcc(a[None])[0].cpu().numpy().astype(np.float64)

    def update_vectors(self, stream) -> np.ndarray:
        if isinstance(stream, np.ndarray):
            buffer_audio = stream
        else:
            if isinstance(stream, (bytes, bytearray)):
                chunk = stream
            else:
                chunk = stream.read(self.chunk_size)
            if len(chunk) == 0:
                raise EOFError
            buffer_audio = np.frombuffer(chunk, dtype='<i2').astype(np.float32, order='C') / 32768.0
        self.window_audio = np.concatenate((self.window_audio, buffer_audio.astype(np.float32)))
        if len(self.window_audio) >= self.pr.window_samples:
            new_features = self._vectorize_raw(self.window_audio)
            self.window_audio = self.window_audio[len(new_features) * self.pr.hop_samples:]
            if len(new_features) > len(self.mfccs):
                new_features = new_features[-len(self.mfccs):]
            self.mfccs = np.concatenate((self.mfccs[len(new_features):], new_features))
        return self.mfccs

    def update_raw(self, stream) -> float:
        mfccs = self.update_vectors(stream)
        if self.pr.use_delta:
            deltas = np.zeros_like(mfccs)
            deltas[1:] = mfccs[1:] - mfccs[:-1]
            mfccs = np.concatenate([mfccs, deltas], -1)
        return self.runner.run(mfccs)

    def update(self, stream) -> float:
        raw = self.update_raw(stream)
        torch = self.core.torch
        r = torch.tensor([raw], dtype=torch.float32, device=self.core.device)
        return float(self.core.decode(r).cpu()[0])


class B200Engine(Engine):
    """``Engine`` plug-in for the reference's PreciseRunner (runner.py:145-243).

    Every ``get_prediction(chunk)`` is one stateful device tick for one stream through the
    host-buffer ABI call (pb_update_host): PCM up, confidence down.
    """

    def __init__(self, model_file, chunk_size=2048, params: ListenerParams = None, device=0):
        Engine.__init__(self, chunk_size)
        if chunk_size % 2:
            raise ValueError('chunk_size is in bytes of int16 audio and must be even')
        self.model, pr = _resolve_model(model_file)
        self.pr = params or pr or ListenerParams()
        self.device = device
        self.core = None

    def start(self):
        m = self.model
        self.core = PreciseB200(self.pr, hidden=m.hidden, max_streams=1, chunk_samples=self.chunk_size // 2,
                                device=self.device, activation=m.activation,
                                recurrent_activation=m.recurrent_activation)
        self.core.load_weights(m.kernel, m.recurrent, m.bias, m.dense_w, m.dense_b)
        # pinned staging: pb_update_host then works in place on these buffers (no staged copies)
        from .core import pinned_empty
        self._pcm, self._pcm_p = pinned_empty((1, self.chunk_size // 2), np.int16)
        self._conf, self._conf_p = pinned_empty((1,), np.float64)

    def stop(self):
        if self.core is not None:
            from .core import pinned_free
            self.core.close()
            self.core = None
            pinned_free(self._pcm_p)
            pinned_free(self._conf_p)
            self._pcm = self._conf = None

    def get_prediction(self, chunk):
        if len(chunk) != self.chunk_size:
            raise ValueError('Invalid chunk size: ' + str(len(chunk)))
        self._pcm[0, :] = np.frombuffer(chunk, dtype='<i2')
        self.core.update_host(self._pcm, self._conf)
        return float(self._conf[0])

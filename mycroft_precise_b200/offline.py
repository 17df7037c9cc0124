"""Batch / offline callers of the hot path, composed from the same kernels.

  vectorize_raw   ~ precise/vectorization.py:46-50   (whole buffer -> MFCC frames)
  vectorize       ~ precise/vectorization.py:62-84   (crop to the last max_samples, left-zero-pad / crop to n_features rows)
  vectorize_delta ~ precise/vectorization.py:87-89
  evaluate        ~ precise/scripts/simulate.py:92-104 and annoyance_estimator.py:115-130
                    (whole-file MFCC, 29-row windows every chunk_size // hop_samples frames, Runner.predict)

The gather of overlapping windows is a strided view on the device tensor (torch, plumbing); MFCC and
the network run in the CUDA library.
"""
import numpy as np

from .core import PreciseB200


def _as_device_audio(core: PreciseB200, audio):
    torch = core.torch
    if isinstance(audio, np.ndarray):
        if audio.dtype == np.int16:
            return torch.from_numpy(np.ascontiguousarray(audio)).to(core.device)
        return torch.from_numpy(np.ascontiguousarray(audio, dtype=np.float32)).to(core.device)
    return audio


def vectorize_raw(core: PreciseB200, audio):
    """audio: 1-D (or [S, L]) int16 / float array or CUDA tensor -> [n_frames, F] (or [S, n_frames, F]) CUDA tensor."""
    a = _as_device_audio(core, audio)
    if a.numel() == 0:
        raise ValueError('Cannot vectorize empty audio!')
    one = a.dim() == 1
    out = core.mfcc(a[None] if one else a)
    return out[0] if one else out


def add_deltas(features):
    """[..., T, F] -> [..., T, 2F]; delta[0] = 0 (precise/vectorization.py:53-59)."""
    import torch
    deltas = torch.zeros_like(features)
    deltas[..., 1:, :] = features[..., 1:, :] - features[..., :-1, :]
    return torch.cat([features, deltas], -1)


def vectorize(core: PreciseB200, audio):
    """Fixed-size network input [n_features, F] for one clip (precise/vectorization.py:62-84)."""
    torch = core.torch
    pr = core.params
    a = _as_device_audio(core, audio)
    if a.shape[-1] > pr.max_samples:
        a = a[..., -pr.max_samples:].contiguous()
    feats = vectorize_raw(core, a)
    n = feats.shape[-2]
    if n < pr.n_features:
        pad = torch.zeros(feats.shape[:-2] + (pr.n_features - n, feats.shape[-1]), dtype=feats.dtype, device=feats.device)
        feats = torch.cat([pad, feats], -2)
    if n > pr.n_features:
        feats = feats[..., -pr.n_features:, :]
    return feats


def vectorize_delta(core: PreciseB200, audio):
    return add_deltas(vectorize(core, audio))


def sliding_windows(core: PreciseB200, mfccs, chunk_size_bytes: int):
    """mfccs [n_frames, F] -> [N, n_features, F]: rows i-n_features..i for i in range(n_features, n_frames, hops)
    (simulate.py:96-99; chunk_size is in bytes of int16 audio as everywhere in the reference)."""
    pr = core.params
    hops = chunk_size_bytes // pr.hop_samples
    if hops < 1:
        raise ValueError('chunk_size smaller than one hop')
    T = pr.n_features
    n = mfccs.shape[0]
    ends = range(T, n, hops)
    if len(ends) == 0:
        return mfccs.new_zeros((0, T, mfccs.shape[1]))
    m = mfccs.contiguous()
    F = m.shape[1]
    view = m.as_strided((len(ends), T, F), (hops * F, F, 1))
    return view.contiguous()


def evaluate(core: PreciseB200, audio, chunk_size_bytes: int = 2048):
    """``SimulateScript.evaluate``: network outputs [N] (float32 CUDA tensor) for every window of one recording."""
    mf = vectorize_raw(core, audio)
    win = sliding_windows(core, mf, chunk_size_bytes)
    if core.params.use_delta:
        win = add_deltas(win)
    if win.shape[0] == 0:
        return win.new_zeros((0,))
    return core.predict(win)

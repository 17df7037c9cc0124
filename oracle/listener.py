"""Oracle restatement of the streaming state machine.  TEST INFRASTRUCTURE ONLY.

Follows reference ``precise/network_runner.py:98-153`` (class ``Listener``) and
``precise/util.py:35-37`` (``buffer_to_audio``).  The state machine (carry buffer, frame
release schedule, 29-row MFCC window) is pinned: ``tests/golden/listener_golden.npz`` was
produced by the reference's real ``Listener`` class driven through its ``runner_cls`` seam
(``tests/golden/make_golden.py``).  The MFCC / GRU arithmetic it calls is the unpinned
restatement in ``oracle/mfcc.py`` / ``oracle/gru.py``.
"""
import numpy as np

from .decoder import OracleDecoder
from .gru import GruWeights, run as gru_run
from .mfcc import vectorize_raw, add_deltas
from .params import OracleParams
from .trigger import OracleTrigger


def buffer_to_audio(buffer: bytes) -> np.ndarray:
    """util.py:35-37 (np.fromstring replaced by its documented equivalent np.frombuffer)."""
    return np.frombuffer(buffer, dtype='<i2').astype(np.float32, order='C') / 32768.0


class OracleListener:
    def __init__(self, weights: GruWeights, pr: OracleParams = None, chunk_size: int = -1):
        self.pr = pr or OracleParams()
        self.weights = weights
        self.chunk_size = chunk_size
        self.decoder = OracleDecoder(self.pr.threshold_config, self.pr.threshold_center)
        self.clear()

    def clear(self):                                   # network_runner.py:121-123
        self.window_audio = np.array([])
        self.mfccs = np.zeros((self.pr.n_features, self.pr.n_mfcc))

    def update_vectors(self, stream) -> np.ndarray:    # network_runner.py:125-146
        if isinstance(stream, np.ndarray):
            buffer_audio = stream
        else:
            chunk = stream if isinstance(stream, (bytes, bytearray)) else stream.read(self.chunk_size)
            if len(chunk) == 0:
                raise EOFError
            buffer_audio = buffer_to_audio(chunk)
        self.window_audio = np.concatenate((self.window_audio, buffer_audio))
        if len(self.window_audio) >= self.pr.window_samples:
            new = vectorize_raw(self.window_audio, self.pr)
            self.window_audio = self.window_audio[len(new) * self.pr.hop_samples:]
            if len(new) > len(self.mfccs):
                new = new[-len(self.mfccs):]
            self.mfccs = np.concatenate((self.mfccs[len(new):], new))
        return self.mfccs

    def update_raw(self, stream) -> float:
        """Network output before decoding (what ``runner.run`` returns, network_runner.py:152)."""
        mfccs = self.update_vectors(stream)
        if self.pr.use_delta:
            mfccs = add_deltas(mfccs)
        return gru_run(self.weights, mfccs)

    def update(self, stream) -> float:                 # network_runner.py:148-153
        return self.decoder.decode(self.update_raw(stream))


def run_streams(weights, pcm_i16: np.ndarray, chunk_samples: int, pr: OracleParams = None,
                sensitivity=0.5, trigger_level=3):
    """Drive S independent oracle listeners + trigger detectors over pcm_i16[S, n_samples].

    Returns (raw[S, K] float32, conf[S, K] float64, fired[S, K] bool) for the
    K = n_samples // chunk_samples complete chunks.  Chunks are fed as float32 ndarrays
    int16/32768, which is what ``buffer_to_audio`` produces for the same bytes.
    """
    pr = pr or OracleParams()
    S, n = pcm_i16.shape
    K = n // chunk_samples
    raw = np.zeros((S, K), dtype=np.float32)
    conf = np.zeros((S, K), dtype=np.float64)
    fired = np.zeros((S, K), dtype=bool)
    for s in range(S):
        lis = OracleListener(weights, pr)
        det = OracleTrigger(chunk_samples * 2, sensitivity, trigger_level)
        for k in range(K):
            chunk = pcm_i16[s, k * chunk_samples:(k + 1) * chunk_samples].astype(np.float32) / 32768.0
            r = lis.update_raw(chunk)
            raw[s, k] = r
            conf[s, k] = lis.decoder.decode(r)
            fired[s, k] = det.update(conf[s, k])
    return raw, conf, fired

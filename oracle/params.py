"""Oracle restatement of the audio-pipeline parameter object.  TEST INFRASTRUCTURE ONLY.

Follows reference ``precise/params.py:29-144`` (class ``ListenerParams`` and the global
default instance ``pr``): same field names, same derived sizes, same rounding rules.
Pinned: ``tests/golden/params_golden.json`` was produced by importing the reference class.
"""
from dataclasses import dataclass, field
from math import floor
from typing import Tuple

# reference precise/params.py:121-133 (class Vectorizer)
VEC_MELS = 1
VEC_MFCCS = 2
VEC_SPEECHPY_MFCCS = 3


@dataclass
class OracleParams:
    # defaults are the reference's global ``pr`` (precise/params.py:140-144)
    buffer_t: float = 1.5
    window_t: float = 0.1
    hop_t: float = 0.05
    sample_rate: int = 16000
    sample_depth: int = 2
    n_fft: int = 512
    n_filt: int = 20
    n_mfcc: int = 13
    use_delta: bool = False
    vectorizer: int = VEC_MFCCS
    threshold_config: Tuple[Tuple[float, float], ...] = field(default=((6, 4),))
    threshold_center: float = 0.2

    # precise/params.py:85-87
    @property
    def window_samples(self) -> int:
        return int(self.sample_rate * self.window_t + 0.5)

    # precise/params.py:90-92
    @property
    def hop_samples(self) -> int:
        return int(self.sample_rate * self.hop_t + 0.5)

    # precise/params.py:74-77
    @property
    def buffer_samples(self) -> int:
        samples = int(self.sample_rate * self.buffer_t + 0.5)
        return self.hop_samples * (samples // self.hop_samples)

    # precise/params.py:80-82
    @property
    def n_features(self) -> int:
        return 1 + int(floor((self.buffer_samples - self.window_samples) / self.hop_samples))

    # precise/params.py:95-97
    @property
    def max_samples(self) -> int:
        return int(self.buffer_t * self.sample_rate)

    # precise/params.py:100-109
    @property
    def feature_size(self) -> int:
        n = {VEC_MFCCS: self.n_mfcc, VEC_MELS: self.n_filt,
             VEC_SPEECHPY_MFCCS: self.n_mfcc}[self.vectorizer]
        return 2 * n if self.use_delta else n

"""Host mirror of the reference's parameter object.

Same field names, defaults, derived properties and `.params` JSON convention as
``precise/params.py:29-165`` (class ListenerParams, global ``pr``, inject_params/save_params), but
an ordinary value object instead of a mutated global.
"""
import json
from math import floor
from os.path import isfile


class Vectorizer:
    """precise/params.py:121-133"""
    mels = 1
    mfccs = 2
    speechpy_mfccs = 3


_FIELDS = ('buffer_t', 'window_t', 'hop_t', 'sample_rate', 'sample_depth', 'n_fft', 'n_filt',
           'n_mfcc', 'use_delta', 'vectorizer', 'threshold_config', 'threshold_center')


class ListenerParams:
    def __init__(self, buffer_t=1.5, window_t=0.1, hop_t=0.05, sample_rate=16000, sample_depth=2,
                 n_fft=512, n_filt=20, n_mfcc=13, use_delta=False, vectorizer=Vectorizer.mfccs,
                 threshold_config=((6, 4),), threshold_center=0.2):
        self.buffer_t = buffer_t
        self.window_t = window_t
        self.hop_t = hop_t
        self.sample_rate = sample_rate
        self.sample_depth = sample_depth
        self.n_fft = n_fft
        self.n_filt = n_filt
        self.n_mfcc = n_mfcc
        self.use_delta = bool(use_delta)
        self.vectorizer = vectorizer
        self.threshold_config = tuple(tuple(p) for p in threshold_config)
        self.threshold_center = threshold_center

    # derived sizes: precise/params.py:74-109
    @property
    def buffer_samples(self):
        samples = int(self.sample_rate * self.buffer_t + 0.5)
        return self.hop_samples * (samples // self.hop_samples)

    @property
    def n_features(self):
        return 1 + int(floor((self.buffer_samples - self.window_samples) / self.hop_samples))

    @property
    def window_samples(self):
        return int(self.sample_rate * self.window_t + 0.5)

    @property
    def hop_samples(self):
        return int(self.sample_rate * self.hop_t + 0.5)

    @property
    def max_samples(self):
        return int(self.buffer_t * self.sample_rate)

    @property
    def feature_size(self):
        n = {Vectorizer.mfccs: self.n_mfcc, Vectorizer.mels: self.n_filt,
             Vectorizer.speechpy_mfccs: self.n_mfcc}[self.vectorizer]
        return 2 * n if self.use_delta else n

    def to_dict(self):
        return {k: getattr(self, k) for k in _FIELDS}

    def __repr__(self):
        return 'ListenerParams(%s)' % ', '.join('%s=%r' % kv for kv in self.to_dict().items())


def load_params(model_name: str) -> ListenerParams:
    """``inject_params`` (precise/params.py:150-159): read ``<model>.params`` if present.

    A file without a 'vectorizer' key selects the legacy speechpy vectorizer
    (``compatibility_params``, params.py:147), which this implementation refuses at create time.
    """
    params_file = model_name + '.params'
    try:
        with open(params_file) as f:
            d = dict(vectorizer=Vectorizer.speechpy_mfccs)
            d.update(json.load(f))
        return ListenerParams(**{k: v for k, v in d.items() if k in _FIELDS})
    except (OSError, ValueError, TypeError):
        if isfile(model_name):
            print('Warning: Failed to load parameters from ' + params_file)
    return ListenerParams()


def save_params(model_name: str, params: ListenerParams):
    """``save_params`` (precise/params.py:162-165)."""
    with open(model_name + '.params', 'w') as f:
        json.dump(params.to_dict(), f)

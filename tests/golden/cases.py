"""Case tables shared by make_golden.py (build container) and the replay tests (anywhere)."""
import numpy as np

DECODER_CASES = [
    (((6, 4),), 0.2),                 # reference default (params.py:143)
    (((6, 4),), 0.5),
    (((0.5, 2.0), (5.0, 1.5)), 0.3),
    (((-3.0, 1.0),), 0.7),
    (((2.2, 0.0),), 0.5),             # std 0 -> out_range 0 branch (threshold_decoder.py:48-49)
]


TRIGGER_CASES = [(2048, 0.5, 3), (2048, 0.2, 0), (1024, 0.5, 3), (4096, 0.8, 1), (3000, 0.5, 2), (2048, 0.5, 10)]


def make_pcm(seed, n, kind='noise'):
    rs = np.random.RandomState(seed)
    if kind == 'noise':
        return np.clip(rs.randn(n) * 3000, -32768, 32767).astype('<i2')
    if kind == 'zero':
        return np.zeros(n, dtype='<i2')
    if kind == 'dc':
        return np.full(n, 32767, dtype='<i2')
    if kind == 'tone':
        t = np.arange(n)
        return (12000 * np.sin(2 * np.pi * 440.0 * t / 16000) + rs.randn(n) * 50).astype('<i2')
    raise ValueError(kind)


LISTENER_CASES = [
    # (pcm kind, chunk_samples, n_chunks)
    ('noise', 1024, 60), ('noise', 512, 70), ('noise', 800, 50), ('noise', 3000, 20),
    ('noise', 333, 90), ('tone', 1024, 45), ('zero', 1024, 40), ('dc', 1024, 40),
]



"""Oracle restatement of the threshold decoder.  TEST INFRASTRUCTURE ONLY.  Pinned by import.

Follows reference ``precise/threshold_decoder.py:38-70`` and the scalar helpers
``precise/functions.py:94-108`` (sigmoid, asigmoid, pdf).  ``tests/golden/decoder_golden.npz``
holds outputs of the reference class itself (imported unmodified in the build container);
``tests/test_oracle_golden.py`` requires bit-equality with them.
"""
from math import exp, log, sqrt, pi

import numpy as np


def sigmoid(x: float) -> float:            # functions.py:94-96
    return 1 / (1 + exp(-x))


def asigmoid(x) -> float:                  # functions.py:99-101: -log(1 / x - 1)
    """Listener.update passes the np.float32 scalar that Runner.run returns (network_runner.py:73-74, :153), so under
    NumPy >= 2 promotion (the reference as it runs in this image) ``1 / x - 1`` is float32 arithmetic and only the
    logarithm is double; a python float (or NumPy 1.16's legacy promotion) evaluates it in float64.  Spelled out here
    so the oracle does not depend on the installed NumPy's promotion rules; tests/golden/decoder_golden.npz holds the
    reference class's own outputs for both kinds of argument (dec32_*, dec_*)."""
    if isinstance(x, np.float32):
        with np.errstate(all='ignore'):
            return -log(float(np.float32(1) / x - np.float32(1)))
    return -log(1 / x - 1)


def pdf(x, mu, std):                       # functions.py:104-108
    if std == 0:
        return 0
    return (1.0 / (std * sqrt(2 * pi))) * np.exp(-(x - mu) ** 2 / (2 * std ** 2))


class OracleDecoder:
    def __init__(self, mu_stds, center=0.5, resolution=200, min_z=-4, max_z=4):
        # threshold_decoder.py:38-43
        self.min_out = int(min(mu + min_z * std for mu, std in mu_stds))
        self.max_out = int(max(mu + max_z * std for mu, std in mu_stds))
        self.out_range = self.max_out - self.min_out
        pts = np.linspace(self.min_out, self.max_out, resolution * self.out_range)
        pd = np.sum([pdf(pts, mu, std) for mu, std in mu_stds], axis=0) / (resolution * len(mu_stds))
        self.cd = np.cumsum(pd)
        self.center = center

    def index(self, raw_output: float) -> int:
        """LUT index selected by decode() (exposed so tests can report bin flips)."""
        ratio = (asigmoid(raw_output) - self.min_out) / self.out_range
        ratio = min(max(ratio, 0.0), 1.0)
        return int(ratio * (len(self.cd) - 1) + 0.5)

    def decode(self, raw_output: float) -> float:
        # threshold_decoder.py:45-57
        if raw_output == 1.0 or raw_output == 0.0:
            return raw_output
        if self.out_range == 0:
            cp = int(raw_output > self.min_out)
        else:
            cp = self.cd[self.index(raw_output)]
        if cp < self.center:
            return 0.5 * cp / self.center
        return 0.5 + 0.5 * (cp - self.center) / (1 - self.center)

    def encode(self, threshold: float) -> float:
        # threshold_decoder.py:59-66
        threshold = 0.5 * threshold / self.center
        if threshold < 0.5:
            cp = threshold * self.center * 2
        else:
            cp = (threshold - 0.5) * 2 * (1 - self.center) + self.center
        ratio = np.searchsorted(self.cd, cp) / len(self.cd)
        return sigmoid(self.min_out + self.out_range * ratio)

"""ctypes wrapper of oracle/c/precise_oracle.c -- TEST INFRASTRUCTURE ONLY.

A second, independent restatement of the same reference algorithm in plain C.  Used (a) by tests/ to cross-check the
numpy oracle and (b) by bench.py as a compiled multi-threaded CPU baseline beside the numpy port.  The product never
imports it.
"""
import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, 'c', '_build', 'libprecise_oracle.so')


class po_params(C.Structure):
    _fields_ = [('sample_rate', C.c_int), ('window', C.c_int), ('hop', C.c_int), ('n_fft', C.c_int), ('n_filt', C.c_int),
                ('n_mfcc', C.c_int), ('n_features', C.c_int), ('hidden', C.c_int), ('n_thresholds', C.c_int),
                ('mu', C.c_double * 8), ('sd', C.c_double * 8), ('center', C.c_double), ('sensitivity', C.c_double),
                ('trigger_level', C.c_int), ('chunk_samples', C.c_int)]


def build():
    subprocess.check_call(['make', '-C', os.path.join(_HERE, 'c')], stdout=subprocess.DEVNULL)


def _lib():
    if not os.path.isfile(_SO):
        build()
    lib = C.CDLL(_SO)
    lib.po_model_create.restype = C.c_void_p
    lib.po_model_create.argtypes = [C.POINTER(po_params), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float]
    lib.po_model_destroy.argtypes = [C.c_void_p]
    lib.po_mfcc.restype = C.c_int
    lib.po_mfcc.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    lib.po_gru.restype = C.c_float
    lib.po_gru.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.po_decode.restype = C.c_double
    lib.po_decode.argtypes = [C.c_void_p, C.c_float]
    lib.po_run_streams.restype = C.c_long
    lib.po_run_streams.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


class COracle:
    def __init__(self, weights, pr, chunk_samples=1024, sensitivity=0.5, trigger_level=3):
        self.lib = _lib()
        p = po_params()
        p.sample_rate, p.window, p.hop, p.n_fft = pr.sample_rate, pr.window_samples, pr.hop_samples, pr.n_fft
        p.n_filt, p.n_mfcc, p.n_features, p.hidden = pr.n_filt, pr.n_mfcc, pr.n_features, weights.H
        p.n_thresholds = len(pr.threshold_config)
        for i, (mu, sd) in enumerate(pr.threshold_config):
            p.mu[i], p.sd[i] = mu, sd
        p.center, p.sensitivity, p.trigger_level, p.chunk_samples = pr.threshold_center, sensitivity, trigger_level, chunk_samples
        self._w = [np.ascontiguousarray(a, dtype=np.float32) for a in (weights.kernel, weights.recurrent, weights.bias, weights.dense_w)]
        self.m = self.lib.po_model_create(C.byref(p), *[a.ctypes.data for a in self._w], C.c_float(weights.dense_b))
        self.n_out = min(pr.n_filt, pr.n_mfcc)
        self.chunk = chunk_samples

    def __del__(self):
        try:
            self.lib.po_model_destroy(self.m)
        except Exception:
            pass

    def mfcc(self, audio):
        a = np.ascontiguousarray(audio, dtype=np.float64)
        out = np.zeros((max(1, len(a)), self.n_out))
        n = self.lib.po_mfcc(self.m, a.ctypes.data, len(a), out.ctypes.data, len(out))
        return out[:n].copy()

    def gru(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        lg = C.c_float()
        p = self.lib.po_gru(self.m, x.ctypes.data, C.cast(C.byref(lg), C.c_void_p))
        return float(p), float(lg.value)

    def decode(self, raw):
        return float(self.lib.po_decode(self.m, C.c_float(raw)))

    def run_streams(self, pcm_i16, threads=1):
        """pcm_i16 [S, K*chunk] -> (raw f32[S,K], conf f64[S,K], fired bool[S,K], detections); streams sharded over threads."""
        S, n = pcm_i16.shape
        K = n // self.chunk
        pcm = np.ascontiguousarray(pcm_i16[:, :K * self.chunk], dtype=np.int16)
        raw = np.zeros((S, K), np.float32); conf = np.zeros((S, K)); fired = np.zeros((S, K), np.uint8)
        bounds = np.linspace(0, S, max(1, min(threads, S)) + 1).astype(int)

        def work(i):
            lo, hi = int(bounds[i]), int(bounds[i + 1])
            if hi <= lo:
                return 0
            return self.lib.po_run_streams(self.m, pcm[lo:hi].ctypes.data, hi - lo, K, raw[lo:hi].ctypes.data,
                                           conf[lo:hi].ctypes.data, fired[lo:hi].ctypes.data)
        with ThreadPoolExecutor(len(bounds) - 1) as ex:
            det = sum(ex.map(work, range(len(bounds) - 1)))
        return raw, conf, fired.astype(bool), int(det)

"""ctypes binding of libprecise_b200.so (C ABI: include/precise_b200.h) + a tensor-level wrapper.

PyTorch is used for device memory, streams and (in dist.py) torch.distributed only; every
computation below is a call into the CUDA library.  No fallback: a missing library or device raises.
"""
import ctypes as C
import os

import numpy as np

from .params import ListenerParams

_HERE = os.path.dirname(os.path.abspath(__file__))
PB_MAX_THRESHOLDS = 8
PB_ABI_VERSION = 1


class PBError(RuntimeError):
    pass


class pb_config(C.Structure):
    _fields_ = [
        ('abi_version', C.c_int32), ('device', C.c_int32), ('max_streams', C.c_int32),
        ('chunk_samples', C.c_int32),
        ('sample_rate', C.c_int32), ('window_samples', C.c_int32), ('hop_samples', C.c_int32),
        ('n_fft', C.c_int32), ('n_filt', C.c_int32), ('n_mfcc', C.c_int32), ('n_features', C.c_int32),
        ('use_delta', C.c_int32), ('vectorizer', C.c_int32),
        ('hidden', C.c_int32), ('activation', C.c_int32), ('recurrent_activation', C.c_int32),
        ('n_thresholds', C.c_int32),
        ('threshold_mu', C.c_double * PB_MAX_THRESHOLDS), ('threshold_std', C.c_double * PB_MAX_THRESHOLDS),
        ('threshold_center', C.c_double),
        ('sensitivity', C.c_double), ('trigger_level', C.c_int32), ('decode_legacy_f64', C.c_int32),
    ]


# name -> (restype, argtypes); must list every symbol include/precise_b200.h declares
_VP, _I64, _I32 = C.c_void_p, C.c_int64, C.c_int32
SYMBOLS = {
    'pb_config_default': (C.c_int, [C.POINTER(pb_config)]),
    'pb_create': (C.c_int, [C.POINTER(pb_config), C.POINTER(_VP)]),
    'pb_destroy': (None, [_VP]),
    'pb_load_weights': (C.c_int, [_VP, _VP, _VP, _VP, _VP, C.c_float]),
    'pb_mfcc_frames': (_I64, [_VP, _I64]),
    'pb_feature_size': (_I32, [_VP]),
    'pb_mfcc_width': (_I32, [_VP]),
    'pb_mfcc': (C.c_int, [_VP, _VP, _I64, _I64, _VP, _VP]),
    'pb_mfcc_f32': (C.c_int, [_VP, _VP, _I64, _I64, _VP, _VP]),
    'pb_predict': (C.c_int, [_VP, _VP, _I64, _VP, _VP, _VP]),
    'pb_decode': (C.c_int, [_VP, _VP, _I64, _VP, _VP]),
    'pb_update': (C.c_int, [_VP, _VP, _VP, _I64, _VP, _VP, _VP, _VP, _VP]),
    'pb_update_vectors': (C.c_int, [_VP, _VP, _VP, _I64, _VP]),
    'pb_update_host': (C.c_int, [_VP, _VP, _VP, _I64, _VP, _VP, _VP, _VP]),
    'pb_read_window': (C.c_int, [_VP, _VP, _I64, _VP, _VP]),
    'pb_clear': (C.c_int, [_VP, _VP, _I64, _VP]),
    'pb_host_alloc': (C.c_int, [C.POINTER(_VP), C.c_uint64]),
    'pb_host_free': (C.c_int, [_VP]),
    'pb_profile_enable': (C.c_int, [_VP, C.c_int]),
    'pb_profile_reset': (C.c_int, [_VP]),
    'pb_profile_read': (C.c_int, [_VP, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    'pb_get_filterbank': (C.c_int, [_VP, _VP]),
    'pb_get_cdf': (_I64, [_VP, _VP, _I64, C.POINTER(_I32), C.POINTER(_I32)]),
    'pb_set_cdf': (C.c_int, [_VP, _VP, _I64]),
    'pb_debug_force_generic': (C.c_int, [_VP, C.c_int]),
    'pb_debug_gru_mode': (C.c_int, [_VP, C.c_int]),
    'pb_debug_k1_mode': (C.c_int, [_VP, C.c_int]),
    'pb_debug_tc_dft_power': (C.c_int, [_VP, _VP]),
    'pb_debug_tc_mfcc_frame': (C.c_int, [_VP, _VP, _VP]),
    'pb_debug_tc3_mfcc_frame': (C.c_int, [_VP, _VP, _VP, _VP]),
    'pb_debug_counters': (C.c_int, [_VP, C.POINTER(C.c_longlong)]),
    'pb_last_error': (C.c_char_p, []),
    'pb_abi_version': (C.c_int, []),
    'pb_build_info': (C.c_char_p, []),
}

_lib = None


def lib_path() -> str:
    return os.environ.get('PRECISE_B200_LIB', os.path.join(_HERE, 'csrc', 'libprecise_b200.so'))


def get_lib():
    """Load the CUDA library (once).  Fails loudly: there is no other implementation."""
    global _lib
    if _lib is None:
        path = lib_path()
        if not os.path.isfile(path):
            raise PBError('%s is missing: build it with `make -C %s` (or __graft_entry__.build()); '
                          'there is no CPU fallback' % (path, os.path.join(_HERE, 'csrc')))
        lib = C.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)          # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if lib.pb_abi_version() != PB_ABI_VERSION:
            raise PBError('ABI mismatch: library %d, binding %d' % (lib.pb_abi_version(), PB_ABI_VERSION))
        _lib = lib
    return _lib


_EXC = {-1: ValueError, -2: NotImplementedError, -3: PBError, -4: PBError, -5: EOFError}


def check(rc: int):
    if rc != 0:
        msg = get_lib().pb_last_error().decode('utf-8', 'replace')
        raise _EXC.get(rc, PBError)(msg)


def make_config(pr: ListenerParams, hidden=20, max_streams=1, chunk_samples=1024, device=0,
                sensitivity=0.5, trigger_level=3, activation='linear',
                recurrent_activation='hard_sigmoid', decode_legacy_f64=False) -> pb_config:
    cfg = pb_config()
    check(get_lib().pb_config_default(C.byref(cfg)))
    cfg.device = device
    cfg.max_streams = max_streams
    cfg.chunk_samples = chunk_samples
    cfg.sample_rate = pr.sample_rate
    cfg.window_samples = pr.window_samples
    cfg.hop_samples = pr.hop_samples
    cfg.n_fft = pr.n_fft
    cfg.n_filt = pr.n_filt
    cfg.n_mfcc = pr.n_mfcc
    cfg.n_features = pr.n_features
    cfg.use_delta = int(pr.use_delta)
    cfg.vectorizer = pr.vectorizer
    cfg.hidden = hidden
    cfg.activation = {'linear': 0, 'tanh': 1}[activation]
    cfg.recurrent_activation = {'hard_sigmoid': 0, 'sigmoid': 1}[recurrent_activation]
    tc = pr.threshold_config
    if not 1 <= len(tc) <= PB_MAX_THRESHOLDS:
        raise ValueError('threshold_config must hold 1..%d (mu, std) pairs' % PB_MAX_THRESHOLDS)
    cfg.n_thresholds = len(tc)
    for i, (mu, std) in enumerate(tc):
        cfg.threshold_mu[i] = mu
        cfg.threshold_std[i] = std
    cfg.threshold_center = pr.threshold_center
    cfg.sensitivity = sensitivity
    cfg.trigger_level = trigger_level
    cfg.decode_legacy_f64 = int(bool(decode_legacy_f64))
    return cfg


def numpy_cdf(threshold_config, resolution=200, min_z=-4, max_z=4):
    """The CDF table exactly as ThresholdDecoder.__init__ builds it (threshold_decoder.py:38-43,
    :68-70, functions.pdf :104-108) -- same numpy calls, so the table is bit-identical to the
    reference's.  This is table construction (6400 doubles, once per handle), not hot-path compute."""
    from math import sqrt, pi
    mu_stds = threshold_config
    min_out = int(min(mu + min_z * std for mu, std in mu_stds))
    max_out = int(max(mu + max_z * std for mu, std in mu_stds))
    out_range = max_out - min_out
    points = np.linspace(min_out, max_out, resolution * out_range)

    def pdf(x, mu, std):
        if std == 0:
            return 0
        return (1.0 / (std * sqrt(2 * pi))) * np.exp(-(x - mu) ** 2 / (2 * std ** 2))

    pd = np.sum([pdf(points, mu, std) for mu, std in mu_stds], axis=0) / (resolution * len(mu_stds))
    return np.ascontiguousarray(np.cumsum(pd), dtype=np.float64), min_out, max_out


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _np_ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _check_np(name, a, dtype, shape, optional=True):
    """The C ABI takes raw host pointers: refuse anything whose dtype / shape / layout it would misread."""
    if a is None:
        if optional:
            return
        raise ValueError('%s is required' % name)
    if not isinstance(a, np.ndarray) or a.dtype != np.dtype(dtype) or not a.flags.c_contiguous or tuple(a.shape) != tuple(shape):
        raise ValueError('%s must be a C-contiguous %s array of shape %s, got %s %s' % (
            name, np.dtype(dtype).name, tuple(shape), getattr(a, 'dtype', type(a)), getattr(a, 'shape', None)))


class PreciseB200:
    """One library handle: tables + per-stream state for ``max_streams`` streams on one GPU.

    Tensor-level API (torch CUDA tensors in, torch CUDA tensors out, asynchronous on the current
    torch stream).  Higher-level mirrors of the reference classes live in runner.py / batch.py.
    """

    def __init__(self, params: ListenerParams = None, hidden=20, max_streams=1, chunk_samples=1024,
                 device=0, sensitivity=0.5, trigger_level=3, activation='linear',
                 recurrent_activation='hard_sigmoid', decode_legacy_f64=False, check_ids=False):
        import torch
        self.torch = torch
        self.lib = get_lib()
        self.params = params or ListenerParams()
        if not torch.cuda.is_available():
            raise PBError('no CUDA device: mycroft_precise_b200 has no CPU path')
        self.device = torch.device('cuda', device)
        self.cfg = make_config(self.params, hidden, max_streams, chunk_samples, device, sensitivity,
                               trigger_level, activation, recurrent_activation, decode_legacy_f64)
        self.check_ids = bool(check_ids)      # debug: also verify 0 <= ids < max_streams and uniqueness (a device sync)
        h = C.c_void_p()
        check(self.lib.pb_create(C.byref(self.cfg), C.byref(h)))
        self._h = h
        self.max_streams = max_streams
        self.chunk_samples = chunk_samples
        self.hidden = hidden
        self.n_features = self.params.n_features
        self.mfcc_width = int(self.lib.pb_mfcc_width(h))
        self.feature_size = int(self.lib.pb_feature_size(h))
        cd, lo, hi = numpy_cdf(self.params.threshold_config)
        if hi > lo:                      # out_range 0 (threshold_decoder.py:48-49): the table is never indexed
            check(self.lib.pb_set_cdf(h, cd.ctypes.data_as(C.c_void_p), len(cd)))
        self._count = torch.zeros(1, dtype=torch.int64, device=self.device)

    def close(self):
        if getattr(self, '_h', None):
            self.lib.pb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    # ---- model
    def load_weights(self, kernel, recurrent, bias, dense_w, dense_b):
        F, H = self.feature_size, self.hidden
        k = np.ascontiguousarray(kernel, dtype=np.float32)
        u = np.ascontiguousarray(recurrent, dtype=np.float32)
        b = np.ascontiguousarray(bias, dtype=np.float32).reshape(-1)
        w = np.ascontiguousarray(dense_w, dtype=np.float32).reshape(-1)
        if k.shape != (F, 3 * H) or u.shape != (H, 3 * H) or b.shape != (3 * H,) or w.shape != (H,):
            raise ValueError('weight shapes %s %s %s %s do not match F=%d, H=%d' % (k.shape, u.shape, b.shape, w.shape, F, H))
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        check(self.lib.pb_load_weights(self._h, vp(k), vp(u), vp(b), vp(w), float(np.asarray(dense_b).reshape(-1)[0])))

    # ---- stateless pieces
    def mfcc_frames(self, n_samples: int) -> int:
        return int(self.lib.pb_mfcc_frames(self._h, n_samples))

    def mfcc(self, pcm):
        """pcm [S, L] int16 or float32 CUDA tensor -> [S, n_frames, mfcc_width] float32."""
        torch = self.torch
        if pcm.dim() == 1:
            pcm = pcm[None]
        pcm = pcm.contiguous()
        S, L = pcm.shape
        nf = self.mfcc_frames(L)
        out = torch.empty((S, nf, self.mfcc_width), dtype=torch.float32, device=self.device)
        if pcm.dtype == torch.int16:
            check(self.lib.pb_mfcc(self._h, _ptr(pcm), S, L, _ptr(out), self._stream()))
        elif pcm.dtype == torch.float32:
            check(self.lib.pb_mfcc_f32(self._h, _ptr(pcm), S, L, _ptr(out), self._stream()))
        else:
            raise ValueError('pcm must be int16 or float32')
        return out

    def predict(self, inputs, want_logit=False):
        """inputs [N, n_features, feature_size] float32 CUDA -> prob [N] float32 (Runner.predict)."""
        torch = self.torch
        inputs = inputs.contiguous()
        if inputs.dim() != 3 or inputs.shape[1] != self.n_features or inputs.shape[2] != self.feature_size:
            raise ValueError('inputs must be [N, %d, %d], got %s' % (self.n_features, self.feature_size, tuple(inputs.shape)))
        if inputs.dtype != torch.float32:
            raise ValueError('inputs must be float32')
        N = inputs.shape[0]
        out = torch.empty(N, dtype=torch.float32, device=self.device)
        logit = torch.empty(N, dtype=torch.float32, device=self.device) if want_logit else None
        check(self.lib.pb_predict(self._h, _ptr(inputs), N, _ptr(out), _ptr(logit), self._stream()))
        return (out, logit) if want_logit else out

    def decode(self, raw):
        torch = self.torch
        raw = raw.contiguous().to(torch.float32)
        out = torch.empty(raw.numel(), dtype=torch.float64, device=self.device)
        check(self.lib.pb_decode(self._h, _ptr(raw), raw.numel(), _ptr(out), self._stream()))
        return out.view(raw.shape)

    # ---- argument checks: the C ABI reads raw device pointers, so dtype / device / size / layout are verified here
    def _check_t(self, name, t, dtype, numel, optional=True):
        if t is None:
            if optional:
                return
            raise ValueError('%s is required' % name)
        if (not isinstance(t, self.torch.Tensor) or t.dtype != dtype or t.device != self.device
                or t.numel() != numel or not t.is_contiguous()):
            raise ValueError('%s must be a contiguous %s tensor with %d elements on %s, got %s %s on %s' % (
                name, dtype, numel, self.device, getattr(t, 'dtype', type(t)), tuple(getattr(t, 'shape', ())),
                getattr(t, 'device', None)))

    def _check_ids(self, ids, n):
        self._check_t('ids', ids, self.torch.int32, n)
        if ids is not None and self.check_ids and n:
            lo, hi = int(ids.min()), int(ids.max())
            if lo < 0 or hi >= self.max_streams:
                raise ValueError('stream ids must lie in [0, %d), got [%d, %d]' % (self.max_streams, lo, hi))
            if int(self.torch.unique(ids).numel()) != n:
                raise ValueError('stream ids must be unique within a tick')

    def _check_pcm(self, pcm):
        torch = self.torch
        if (not isinstance(pcm, torch.Tensor) or pcm.dtype != torch.int16 or pcm.dim() != 2 or pcm.shape[1] != self.chunk_samples
                or not pcm.is_contiguous() or pcm.device != self.device):
            raise ValueError('pcm must be a contiguous int16 [n, %d] tensor on %s' % (self.chunk_samples, self.device))
        if pcm.shape[0] > self.max_streams:
            raise ValueError('n = %d exceeds max_streams = %d' % (pcm.shape[0], self.max_streams))
        return pcm.shape[0]

    # ---- stateful tick
    def update(self, pcm, ids=None, out=None, count=None):
        """pcm [n, chunk_samples] int16 CUDA.  Returns dict(raw, conf, fired) (+ count accumulates)."""
        torch = self.torch
        n = self._check_pcm(pcm)
        self._check_ids(ids, n)
        if out is not None:
            self._check_t("out['raw']", out.get('raw'), torch.float32, n)
            self._check_t("out['conf']", out.get('conf'), torch.float64, n, optional=False)
            self._check_t("out['fired']", out.get('fired'), torch.uint8, n)
        self._check_t('count', count, torch.int64, 1)
        if out is None:
            out = dict(raw=torch.empty(n, dtype=torch.float32, device=self.device),
                       conf=torch.empty(n, dtype=torch.float64, device=self.device),
                       fired=torch.empty(n, dtype=torch.uint8, device=self.device))
        check(self.lib.pb_update(self._h, _ptr(pcm), _ptr(ids), n, _ptr(out.get('raw')), _ptr(out['conf']),
                                 _ptr(out.get('fired')), _ptr(count), self._stream()))
        return out

    def update_vectors(self, pcm, ids=None):
        n = self._check_pcm(pcm)
        self._check_ids(ids, n)
        check(self.lib.pb_update_vectors(self._h, _ptr(pcm), _ptr(ids), n, self._stream()))

    def read_window(self, n=None, ids=None):
        torch = self.torch
        n = (ids.numel() if ids is not None else (self.max_streams if n is None else n))
        self._check_ids(ids, n)
        out = torch.empty((n, self.n_features, self.mfcc_width), dtype=torch.float32, device=self.device)
        check(self.lib.pb_read_window(self._h, _ptr(ids), n, _ptr(out), self._stream()))
        return out

    def clear(self, n=None, ids=None):
        n = (ids.numel() if ids is not None else (self.max_streams if n is None else n))
        self._check_ids(ids, n)
        check(self.lib.pb_clear(self._h, _ptr(ids), n, self._stream()))

    def update_host(self, pcm_np, conf_np, raw_np=None, fired_np=None, ids_np=None) -> int:
        """Host-buffer tick (numpy arrays, ideally backed by pinned memory).  Returns this tick's count."""
        if not isinstance(pcm_np, np.ndarray) or pcm_np.ndim != 2:
            raise ValueError('pcm_np must be an int16 array of shape (n, %d)' % self.chunk_samples)
        n = pcm_np.shape[0]
        _check_np('pcm_np', pcm_np, np.int16, (n, self.chunk_samples), optional=False)
        _check_np('conf_np', conf_np, np.float64, (n,), optional=False)
        _check_np('raw_np', raw_np, np.float32, (n,))
        _check_np('fired_np', fired_np, np.uint8, (n,))
        _check_np('ids_np', ids_np, np.int32, (n,))
        if ids_np is not None and self.check_ids and n:
            if ids_np.min() < 0 or ids_np.max() >= self.max_streams or len(np.unique(ids_np)) != n:
                raise ValueError('stream ids must be unique and lie in [0, %d)' % self.max_streams)
        cnt = C.c_uint64(0)
        vp = _np_ptr
        check(self.lib.pb_update_host(self._h, vp(pcm_np), vp(ids_np), n, vp(raw_np), vp(conf_np), vp(fired_np),
                                      C.cast(C.byref(cnt), C.c_void_p)))
        return int(cnt.value)

    def debug_counters(self):
        out = (C.c_longlong * 4)()
        check(self.lib.pb_debug_counters(self._h, out))
        return list(out)

    def gru_mode(self, mode):
        check(self.lib.pb_debug_gru_mode(self._h, int(mode)))

    def force_generic(self, on=True):
        check(self.lib.pb_debug_force_generic(self._h, int(on)))

    def k1_mode(self, mode):
        """1 = experimental tensor-core DFT tick (csrc/mfcc_tc.cuh, not yet validated on hardware), 0 = default kernels."""
        check(self.lib.pb_debug_k1_mode(self._h, int(mode)))

    # ---- profiling / introspection
    def profile(self, on=True):
        check(self.lib.pb_profile_enable(self._h, int(on)))
        check(self.lib.pb_profile_reset(self._h))

    def profile_read(self):
        ms = (C.c_double * 4)()
        ln = (C.c_uint64 * 4)()
        check(self.lib.pb_profile_read(self._h, ms, ln))
        return list(ms), list(ln)

    def filterbank(self) -> np.ndarray:
        fb = np.zeros((self.params.n_filt, self.params.n_fft // 2 + 1), dtype=np.float64)
        check(self.lib.pb_get_filterbank(self._h, fb.ctypes.data_as(C.c_void_p)))
        return fb

    def cdf(self):
        lo, hi = C.c_int32(), C.c_int32()
        n = self.lib.pb_get_cdf(self._h, None, 0, C.byref(lo), C.byref(hi))
        cd = np.zeros(n, dtype=np.float64)
        self.lib.pb_get_cdf(self._h, cd.ctypes.data_as(C.c_void_p), n, C.byref(lo), C.byref(hi))
        return cd, lo.value, hi.value


def pinned_empty(shape, dtype):
    """numpy array backed by cudaHostAlloc memory (for update_host at PCIe rate)."""
    lib = get_lib()
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    p = C.c_void_p()
    check(lib.pb_host_alloc(C.byref(p), nbytes))
    buf = (C.c_char * max(nbytes, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    return arr, p


def pinned_free(p):
    check(get_lib().pb_host_free(p))

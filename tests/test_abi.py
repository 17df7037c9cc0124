"""CPU-only: the C-ABI library loads, exports every symbol include/precise_b200.h declares, the
ctypes struct matches the C struct, and the product fails loudly without a GPU (no fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as g
    from mycroft_precise_b200.core import lib_path
    if not os.path.isfile(lib_path()):
        g.build()
    from mycroft_precise_b200.core import get_lib
    return get_lib()


def header_symbols():
    src = open(os.path.join(ROOT, 'include', 'precise_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(pb_[a-z0-9_]+)\s*\(', src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from mycroft_precise_b200.core import SYMBOLS
    names = header_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(SYMBOLS) == names


def test_config_struct_layout_and_defaults(lib):
    from mycroft_precise_b200.core import pb_config, PB_ABI_VERSION
    cfg = pb_config()
    assert lib.pb_config_default(C.byref(cfg)) == 0
    assert cfg.abi_version == PB_ABI_VERSION == lib.pb_abi_version()
    got = (cfg.chunk_samples, cfg.sample_rate, cfg.window_samples, cfg.hop_samples, cfg.n_fft, cfg.n_filt,
           cfg.n_mfcc, cfg.n_features, cfg.use_delta, cfg.vectorizer, cfg.hidden, cfg.n_thresholds,
           cfg.threshold_mu[0], cfg.threshold_std[0], cfg.threshold_center, cfg.sensitivity, cfg.trigger_level)
    # reference defaults: precise/params.py:140-144, model.py:40, runner.py:23,121
    assert got == (1024, 16000, 1600, 800, 512, 20, 13, 29, 0, 2, 20, 1, 6.0, 4.0, 0.2, 0.5, 3)
    assert C.sizeof(pb_config) == 17 * 4 + 4 + 8 * 8 * 2 + 8 + 8 + 4 + 4       # incl. padding before the doubles


def test_make_config_follows_listener_params(lib):
    from mycroft_precise_b200 import ListenerParams
    from mycroft_precise_b200.core import make_config
    pr = ListenerParams(n_filt=40, n_mfcc=40, threshold_config=((1, 2), (3, 4)), threshold_center=0.4)
    cfg = make_config(pr, hidden=128, max_streams=7, chunk_samples=512)
    assert (cfg.n_filt, cfg.n_mfcc, cfg.hidden, cfg.max_streams, cfg.chunk_samples, cfg.n_thresholds) == (40, 40, 128, 7, 512, 2)
    assert list(cfg.threshold_mu)[:2] == [1, 3] and list(cfg.threshold_std)[:2] == [2, 4]
    assert (cfg.window_samples, cfg.hop_samples, cfg.n_features) == (1600, 800, 29)


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    import mycroft_precise_b200 as m
    with pytest.raises(m.PBError):
        m.PreciseB200()
    from mycroft_precise_b200.core import pb_config
    cfg = pb_config()
    lib.pb_config_default(C.byref(cfg))
    h = C.c_void_p()
    rc = lib.pb_create(C.byref(cfg), C.byref(h))
    assert rc == -3 and not h.value and b'cuda' in lib.pb_last_error().lower()


def test_argument_validation_without_gpu(lib):
    from mycroft_precise_b200.core import pb_config
    cfg = pb_config()
    lib.pb_config_default(C.byref(cfg))
    h = C.c_void_p()
    cfg.n_fft = 2048
    assert lib.pb_create(C.byref(cfg), C.byref(h)) == -2          # n_fft > 1024: unsupported, reported before any CUDA call
    cfg.n_fft = 512
    cfg.vectorizer = 7
    assert lib.pb_create(C.byref(cfg), C.byref(h)) == -1          # unknown vectorizer
    cfg.vectorizer = 2
    cfg.abi_version = 99
    assert lib.pb_create(C.byref(cfg), C.byref(h)) == -1
    cfg.abi_version = 1
    cfg.max_streams = 0
    assert lib.pb_create(C.byref(cfg), C.byref(h)) == -1
    assert lib.pb_mfcc(None, None, 1, 1, None, None) == -1


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'mycroft_precise_b200')
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                assert 'oracle' not in open(os.path.join(dp, f)).read().replace('# oracle', ''), f


def test_params_mirror_matches_reference_golden(golden_dir):
    import json
    from mycroft_precise_b200 import ListenerParams
    for c in json.load(open(os.path.join(golden_dir, 'params_golden.json'))):
        p = ListenerParams(**c['fields'])
        for k, v in c['derived'].items():
            assert getattr(p, k) == v


def test_host_trigger_mirror_matches_reference_golden(golden_dir):
    import numpy as np
    from mycroft_precise_b200 import TriggerDetector
    from golden.cases import TRIGGER_CASES
    g = np.load(os.path.join(golden_dir, 'trigger_golden.npz'))
    for i, (chunk, sens, lvl) in enumerate(TRIGGER_CASES):
        det = TriggerDetector(chunk, sens, lvl)
        assert [det.update(float(p)) for p in g['probs_%d' % i]] == list(g['fired_%d' % i])


def test_numpy_cdf_equals_reference_table(golden_dir):
    import numpy as np
    from mycroft_precise_b200.core import numpy_cdf
    from golden.cases import DECODER_CASES
    g = np.load(os.path.join(golden_dir, 'decoder_golden.npz'))
    for i, (cfg, _) in enumerate(DECODER_CASES):
        cd, lo, hi = numpy_cdf(cfg)
        assert np.array_equal(cd, g['cd_%d' % i]) and [lo, hi] == list(g['meta_%d' % i][:2])

"""Oracle vs fixtures produced by the reference's own code (tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest

from oracle.decoder import OracleDecoder
from oracle.gru import GruWeights
from oracle.listener import OracleListener
from oracle.params import OracleParams
from oracle.trigger import OracleTrigger

from golden.cases import DECODER_CASES, TRIGGER_CASES, LISTENER_CASES


def test_params_match_reference_class(golden_dir):
    cases = json.load(open(os.path.join(golden_dir, 'params_golden.json')))
    assert len(cases) >= 6
    for c in cases:
        f = dict(c['fields'])
        f['threshold_config'] = tuple(tuple(x) for x in f['threshold_config'])
        p = OracleParams(**f)
        for k, v in c['derived'].items():
            assert getattr(p, k) == v, (k, f)


def test_default_sizes():
    p = OracleParams()
    assert (p.window_samples, p.hop_samples, p.buffer_samples, p.max_samples,
            p.n_features, p.feature_size) == (1600, 800, 24000, 24000, 29, 13)


@pytest.mark.parametrize('i', range(len(DECODER_CASES)))
def test_decoder_bit_equal(golden_dir, i):
    g = np.load(os.path.join(golden_dir, 'decoder_golden.npz'))
    cfg, center = DECODER_CASES[i]
    d = OracleDecoder(cfg, center)
    assert [d.min_out, d.max_out, d.out_range] == list(g['meta_%d' % i])
    assert np.array_equal(d.cd, g['cd_%d' % i])
    got = np.array([d.decode(float(r)) for r in g['raws']])
    assert np.array_equal(got, g['dec_%d' % i])
    with np.errstate(all='ignore'):
        got32 = np.array([d.decode(r) for r in g['raws']])          # np.float32 scalars in, like Listener.update
    assert np.array_equal(got32, g['dec32_%d' % i])
    if d.out_range:
        enc = np.array([d.encode(t) for t in np.linspace(0.02, 0.98, 49)])
        assert np.array_equal(enc, g['enc_%d' % i])


def test_decoder_kat(golden_dir):
    # values probed from the imported reference class
    g = np.load(os.path.join(golden_dir, 'decoder_golden.npz'))
    want = [0.0, 1.0, 0.16724202203314845, 0.4274002278048867, 0.6019128072385334, 0.7436703577978986]
    assert np.allclose(g['kat'], want, rtol=0, atol=1e-15)
    d = OracleDecoder(((6, 4),), 0.2)
    assert len(d.cd) == 6400 and abs(d.cd[-1] - 0.99978) < 1e-5
    assert [d.decode(v) for v in (0.0, 1.0, 0.5, 0.9, 0.99, 0.999)] == list(g['kat'])


@pytest.mark.parametrize('i', range(len(TRIGGER_CASES)))
def test_trigger_traces(golden_dir, i):
    g = np.load(os.path.join(golden_dir, 'trigger_golden.npz'))
    chunk, sens, lvl = TRIGGER_CASES[i]
    det = OracleTrigger(chunk, sens, lvl)
    fired, act = [], []
    for p in g['probs_%d' % i]:
        fired.append(det.update(float(p)))
        act.append(det.activation)
    assert np.array_equal(fired, g['fired_%d' % i])
    assert np.array_equal(act, g['act_%d' % i])
    assert g['fired_%d' % i].sum() > 0


def test_trigger_kat(golden_dir):
    g = np.load(os.path.join(golden_dir, 'trigger_golden.npz'))
    det = OracleTrigger(2048, 0.5, 3)
    got = [det.update(p) for p in [0.9] * 6 + [0.1] * 3 + [0.9] * 10]
    assert list(g['kat']) == got
    assert [i for i, f in enumerate(got) if f] == [3]


@pytest.mark.parametrize('i', range(len(LISTENER_CASES)))
def test_listener_state_machine(golden_dir, i):
    """oracle.listener == the reference's real Listener class on the same chunks."""
    g = np.load(os.path.join(golden_dir, 'listener_golden.npz'))
    w = GruWeights(g['kernel'], g['recurrent'], g['bias'], g['dense_w'], g['dense_b'])
    kind, chunk, n_chunks = LISTENER_CASES[i]
    assert int(g['chunk_%d' % i]) == chunk
    pcm = g['pcm_%d' % i]
    lis = OracleListener(w)
    confs, rings = [], []
    for k in range(n_chunks):
        a = pcm[k * chunk:(k + 1) * chunk].astype(np.float32) / 32768.0
        confs.append(lis.update(a))
        rings.append(lis.mfccs.copy())
    assert np.array_equal(np.array(confs), g['conf_%d' % i])
    assert np.array_equal(np.array(rings)[::5], g['ring_%d' % i])
    assert len(lis.window_audio) == int(g['carry_%d' % i])


def test_listener_bytes_equals_ndarray(golden_dir):
    g = np.load(os.path.join(golden_dir, 'listener_golden.npz'))
    w = GruWeights(g['kernel'], g['recurrent'], g['bias'], g['dense_w'], g['dense_b'])
    pcm = g['pcm_0'][:1024 * 12]
    a, b = OracleListener(w), OracleListener(w)
    for k in range(12):
        c = pcm[k * 1024:(k + 1) * 1024]
        assert a.update(c.tobytes()) == b.update(c.astype(np.float32) / 32768.0)
    with pytest.raises(EOFError):
        a.update(b'')

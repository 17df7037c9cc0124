"""The plain-C restatement (oracle/c) against the numpy oracle: two independent implementations of the same reference
algorithm must agree -- MFCC in float64 to 1e-9, the fp32 network to 1e-6, decode / trigger exactly."""
import numpy as np
import pytest

from oracle import mfcc as om
from oracle import gru as og
from oracle.cport import COracle
from oracle.decoder import OracleDecoder
from oracle.listener import run_streams
from oracle.params import OracleParams
from golden.cases import make_pcm


@pytest.fixture(scope='module')
def co():
    w = og.GruWeights.random(13, 20, seed=0, scale=0.1)
    return COracle(w, OracleParams()), w


def test_mfcc_matches_numpy_oracle(co):
    c, _ = co
    pr = OracleParams()
    for kind, n in (('noise', 24000), ('tone', 9999), ('zero', 1600), ('dc', 3000)):
        a = make_pcm(3, n, kind).astype(np.float32) / 32768.0
        got, want = c.mfcc(a), om.vectorize_raw(a, pr)
        assert got.shape == want.shape
        assert np.max(np.abs(got - want)) < 1e-9
    assert c.mfcc(np.zeros(1599)).shape == (0, 13)


def test_network_and_decoder(co):
    c, w = co
    x = (np.random.RandomState(1).randn(50, 29, 13) * 3).astype(np.float32)
    p64, l64 = og.gru_forward(w, x, np.float64)
    d = OracleDecoder(((6, 4),), 0.2)
    for i in range(50):
        p, lg = c.gru(x[i])
        assert abs(p - p64[i]) < 1e-6 and abs(lg - l64[i]) < 1e-5
        assert c.decode(np.float32(p)) == d.decode(float(np.float32(p)))


def test_streaming_state_machine_and_trigger():
    w = og.GruWeights.random(13, 20, seed=0, scale=0.1)
    w.dense_b = 3.0
    pr = OracleParams()
    for chunk in (1024, 800, 333):
        S, K = 6, 40
        pcm = np.stack([make_pcm(40 + s, K * chunk, 'noise' if s < 4 else ('zero' if s == 4 else 'dc')) for s in range(S)])
        c = COracle(w, pr, chunk_samples=chunk)
        raw, conf, fired, det = c.run_streams(pcm, threads=3)
        oraw, oconf, ofired = run_streams(w, pcm, chunk)
        assert np.max(np.abs(raw - oraw)) < 1e-6
        assert np.max(np.abs(conf - oconf)) < 1e-3                 # a neighbouring LUT bin at most
        assert (conf == oconf).mean() > 0.99
        assert np.array_equal(fired, ofired) and det == ofired.sum() > 0

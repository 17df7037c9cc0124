"""-m gpu: the batch/offline callers and the precise-engine wire protocol."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')

from oracle import gru as og          # noqa: E402
from oracle import mfcc as om         # noqa: E402
from oracle.listener import run_streams   # noqa: E402
from oracle.params import OracleParams   # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_vectorize_pad_crop_delta():
    import mycroft_precise_b200 as m
    from mycroft_precise_b200 import offline
    core = m.PreciseB200()
    opr = OracleParams()
    rs = np.random.RandomState(1)
    for n in (1600, 4000, 24000, 40001):
        a = (rs.randn(n) * 0.1).astype(np.float32)
        got = offline.vectorize(core, a).cpu().numpy()
        want = om.vectorize(a, opr)
        assert got.shape == (29, 13) and np.max(np.abs(got - want)) < 2e-4
        d = offline.vectorize_delta(core, a).cpu().numpy()
        assert np.max(np.abs(d - om.add_deltas(want))) < 4e-4
    i16 = (rs.randn(30000) * 3000).astype(np.int16)
    got = offline.vectorize(core, i16).cpu().numpy()
    assert np.max(np.abs(got - om.vectorize(i16[-24000:].astype(np.float32) / 32768.0, opr))) < 2e-4
    with pytest.raises(ValueError):
        offline.vectorize_raw(core, np.zeros(0, np.float32))
    core.close()


@pytest.mark.parametrize('chunk_bytes', [2048, 4096, 800])
def test_simulate_evaluate(chunk_bytes):
    import mycroft_precise_b200 as m
    from mycroft_precise_b200 import offline
    model = m.GruModel.random(13, 20, seed=2, scale=0.1)
    core = m.PreciseB200(hidden=20)
    core.load_weights(model.kernel, model.recurrent, model.bias, model.dense_w, model.dense_b)
    rs = np.random.RandomState(3)
    audio = (rs.randn(16000 * 20) * 0.1).astype(np.float32)          # 20 s recording
    got = offline.evaluate(core, audio, chunk_bytes).cpu().numpy()
    opr = OracleParams()
    mf = om.vectorize_raw(audio, opr)
    hops = chunk_bytes // opr.hop_samples
    inputs = np.array([mf[i - 29:i] for i in range(29, len(mf), hops)])     # simulate.py:96-99
    w = og.GruWeights(model.kernel, model.recurrent, model.bias, model.dense_w, model.dense_b)
    want = og.predict(w, inputs)[:, 0]
    assert got.shape == want.shape and len(got) > 50
    assert np.max(np.abs(got - want)) < 1e-4
    core.close()


def test_engine_wire_protocol(tmp_path):
    import mycroft_precise_b200 as m
    model = m.GruModel.random(13, 20, seed=4, scale=0.1)
    path = str(tmp_path / 'model.npz')
    m.save_weights(path, model)
    rs = np.random.RandomState(5)
    pcm = np.clip(rs.randn(1024 * 30) * 3000, -32768, 32767).astype('<i2')
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, '-m', 'mycroft_precise_b200.engine', path, '2048'], input=pcm.tobytes(),
                       capture_output=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    lines = r.stdout.split(b'\n')[:-1]
    assert len(lines) == 30
    for ln in lines:
        assert re.match(rb'[01]\.[0-9]+', ln) or re.match(rb'[0-9.]+e-[0-9]+', ln)    # reference test: test_engine.py:50
    w = og.GruWeights(model.kernel, model.recurrent, model.bias, model.dense_w, model.dense_b)
    raw, conf, fired = run_streams(w, pcm[None], 1024)
    got = np.array([float(x) for x in lines])
    from oracle.decoder import OracleDecoder
    from oracle.params import OracleParams
    d = OracleDecoder(OracleParams().threshold_config, OracleParams().threshold_center)
    step = np.max(np.abs(np.diff(d.cd))) * 2.5                     # at most a neighbouring LUT bin
    assert np.max(np.abs(got - conf[0])) <= step
    assert np.mean(got != conf[0]) < 0.1
    # a short last read still gets its answer (reference: Listener.update processes whatever stream.read returned)
    tail = np.clip(rs.randn(700) * 3000, -32768, 32767).astype('<i2')
    r = subprocess.run([sys.executable, '-m', 'mycroft_precise_b200.engine', path, '2048'], input=pcm.tobytes() + tail.tobytes(),
                       capture_output=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    lines2 = r.stdout.split(b'\n')[:-1]
    assert len(lines2) == 31 and lines2[:30] == lines
    from oracle.listener import OracleListener
    lis = OracleListener(w)
    for k in range(30):
        lis.update(pcm[k * 1024:(k + 1) * 1024].astype(np.float32) / 32768.0)
    assert abs(float(lines2[30]) - lis.update(tail.astype(np.float32) / 32768.0)) <= step
    # chunks that complete more than 8 frames per tick go through the stateless mirror: same numbers, chunking-independent state
    r = subprocess.run([sys.executable, '-m', 'mycroft_precise_b200.engine', path, str(2 * 10240)], input=pcm.tobytes(),
                       capture_output=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    big = [float(x) for x in r.stdout.split(b'\n')[:-1]]
    assert len(big) == 3 and all(abs(big[i] - conf[0, 10 * i + 9]) <= step for i in range(3))
    # chunk_size = -1: read everything, one prediction (precise/scripts/engine.py: default)
    r = subprocess.run([sys.executable, '-m', 'mycroft_precise_b200.engine', path], input=pcm.tobytes(),
                       capture_output=True, env=env, timeout=300)
    assert r.returncode == 0 and len(r.stdout.split(b'\n')) == 2
    assert abs(float(r.stdout) - conf[0, -1]) <= step          # state is chunking-independent
    r = subprocess.run([sys.executable, '-m', 'mycroft_precise_b200.engine', '-v'], capture_output=True, env=env, timeout=120)
    assert r.stdout.strip() == m.__version__.encode()


def test_model_formats_through_drop_in_classes(tmp_path):
    """The three model file formats the reference's Listener.find_runner dispatches on (network_runner.py:111-119: .net Keras HDF5,
    .pb frozen graph) plus this package's .npz, each through B200Runner, B200Engine and the precise-engine CLI: same numbers as
    the in-memory model.  (.net written by tests/h5_writer.py, .pb by the GraphDef writer of tests/test_model_io.py -- files from
    Keras / TensorFlow themselves cannot be produced in this image.)"""
    import mycroft_precise_b200 as m
    sys.path.insert(0, os.path.dirname(__file__))
    from h5_writer import keras_model_file
    from test_model_io import _const_node, _ld, _model_config
    model = m.GruModel.random(13, 20, seed=12, scale=0.1)
    paths = {}
    paths['npz'] = str(tmp_path / 'model.npz')
    m.save_weights(paths['npz'], model)
    paths['pb'] = str(tmp_path / 'model.pb')
    open(paths['pb'], 'wb').write(_ld(1, _ld(1, b'import/net_input') + _ld(2, b'Placeholder')) + _const_node('net/kernel', model.kernel)
                                   + _const_node('net/recurrent_kernel', model.recurrent) + _const_node('net/bias', model.bias)
                                   + _const_node('dense_1/kernel', model.dense_w.reshape(20, 1)) + _const_node('dense_1/bias', np.float32([model.dense_b])))
    paths['net'] = str(tmp_path / 'model.net')
    open(paths['net'], 'wb').write(keras_model_file(model.kernel, model.recurrent, model.bias, model.dense_w, model.dense_b, _model_config()))
    rs = np.random.RandomState(3)
    x = rs.randn(7, 29, 13).astype(np.float32)
    pcm = np.clip(rs.randn(1024 * 28) * 3000, -32768, 32767).astype('<i2')
    want_p = m.B200Runner(model).predict(x)
    eng = m.B200Engine(model, 2048); eng.start()
    want_e = [eng.get_prediction(pcm[k * 1024:(k + 1) * 1024].tobytes()) for k in range(28)]
    eng.stop()
    env = dict(os.environ, PYTHONPATH=ROOT)
    for kind, path in paths.items():
        assert np.array_equal(m.B200Runner(path).predict(x), want_p), kind
        eng = m.B200Engine(path, 2048); eng.start()
        got_e = [eng.get_prediction(pcm[k * 1024:(k + 1) * 1024].tobytes()) for k in range(28)]
        eng.stop()
        assert got_e == want_e, kind
        r = subprocess.run([sys.executable, '-m', 'mycroft_precise_b200.engine', path, '2048'], input=pcm.tobytes(), capture_output=True, env=env, timeout=300)
        assert r.returncode == 0, (kind, r.stderr.decode()[-1500:])
        assert [float(v) for v in r.stdout.split(b'\n')[:-1]] == want_e, kind

"""CPU: weight containers and the TensorFlow-free GraphDef (.pb) importer."""
import json
import os
import struct

import numpy as np
import pytest

from mycroft_precise_b200 import GruModel, ListenerParams, Vectorizer, load_params, load_weights, save_weights
from mycroft_precise_b200.params import save_params


def _vi(n):
    out = b''
    while True:
        b = n & 0x7f
        n >>= 7
        out += bytes([b | (0x80 if n else 0)])
        if not n:
            return out


def _ld(fno, payload):
    return _vi(fno << 3 | 2) + _vi(len(payload)) + payload


def _const_node(name, arr, as_content=True):
    shape = b''.join(_ld(2, _vi(1 << 3) + _vi(d)) for d in arr.shape)
    tensor = _vi(1 << 3) + _vi(1) + _ld(2, shape)
    if as_content:
        tensor += _ld(4, arr.astype('<f4').tobytes())
    else:
        tensor += _ld(5, arr.astype('<f4').tobytes())          # packed float_val
    attr_val = _ld(8, tensor)
    attr = _ld(1, b'value') + _ld(2, attr_val)
    dtype_attr = _ld(1, b'dtype') + _ld(2, _vi(6 << 3) + _vi(1))
    return _ld(1, _ld(1, name.encode()) + _ld(2, b'Const') + _ld(5, dtype_attr) + _ld(5, attr))


def test_pb_import_roundtrip(tmp_path):
    m = GruModel.random(13, 20, seed=5, scale=0.2)
    other = _ld(1, _ld(1, b'import/net_input') + _ld(2, b'Placeholder'))
    blob = (other + _const_node('net/kernel', m.kernel) + _const_node('net/recurrent_kernel', m.recurrent, as_content=False)
            + _const_node('net/bias', m.bias) + _const_node('dense_1/kernel', m.dense_w.reshape(20, 1))
            + _const_node('dense_1/bias', np.float32([m.dense_b]))
            + _ld(1, _ld(1, b'net/while/add/y') + _ld(2, b'Const')))
    p = tmp_path / 'model.pb'
    p.write_bytes(blob)
    g = load_weights(str(p))
    assert g.hidden == 20 and g.feature_size == 13
    for a, b in ((g.kernel, m.kernel), (g.recurrent, m.recurrent), (g.bias, m.bias), (g.dense_w, m.dense_w)):
        assert np.array_equal(a, b)
    assert g.dense_b == pytest.approx(m.dense_b)


def test_npz_roundtrip_and_params_file(tmp_path):
    m = GruModel.random(26, 32, seed=1)
    path = str(tmp_path / 'w.npz')
    save_weights(path, m)
    g = load_weights(path)
    assert np.array_equal(g.recurrent, m.recurrent) and g.activation == 'linear'
    pr = ListenerParams(n_mfcc=13, use_delta=True, threshold_config=((5, 3),), threshold_center=0.3)
    save_params(path, pr)
    q = load_params(path)
    assert q.to_dict() == pr.to_dict() and q.feature_size == 26
    # a .params file written before the 'vectorizer' field existed selects speechpy (params.py:147)
    d = pr.to_dict(); d.pop('vectorizer')
    json.dump(d, open(path + '.params', 'w'))
    assert load_params(path).vectorizer == Vectorizer.speechpy_mfccs
    assert load_params(str(tmp_path / 'missing.npz')).to_dict() == ListenerParams().to_dict()
    with pytest.raises(ValueError):
        load_weights(str(tmp_path / 'model.h5x'))


class _FakeH5Group(dict):
    """Mapping with the part of the h5py Group / File interface that model_from_keras_h5 uses."""
    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.attrs = {}


def _keras_like_file(m, with_model_weights=True, ract='hard_sigmoid'):
    """Layout written by Keras 2.x model.save for precise/model.py:72-82 (layer 'net' = GRU, 'dense_1' = Dense)."""
    import json
    net = _FakeH5Group({'net': _FakeH5Group({'kernel:0': m.kernel, 'recurrent_kernel:0': m.recurrent, 'bias:0': m.bias})})
    dense = _FakeH5Group({'dense_1': _FakeH5Group({'kernel:0': m.dense_w.reshape(-1, 1), 'bias:0': np.array([m.dense_b], np.float32)})})
    weights = _FakeH5Group({'net': net, 'dense_1': dense})
    f = _FakeH5Group({'model_weights': weights, 'optimizer_weights': _FakeH5Group()}) if with_model_weights else weights
    cfg = {'class_name': 'Sequential', 'config': [
        {'class_name': 'GRU', 'config': {'name': 'net', 'units': m.hidden, 'activation': 'linear',
                                         'recurrent_activation': ract, 'dropout': 0.2}},
        {'class_name': 'Dense', 'config': {'name': 'dense_1', 'units': 1, 'activation': 'sigmoid'}}]}
    if with_model_weights:
        f.attrs['model_config'] = json.dumps(cfg).encode('utf-8')
    return f


def test_keras_h5_layout_extraction():
    from mycroft_precise_b200.model_io import GruModel, model_from_keras_h5
    m = GruModel.random(13, 20, seed=4)
    for full in (True, False):
        got = model_from_keras_h5(_keras_like_file(m, with_model_weights=full))
        for k in ('kernel', 'recurrent', 'bias', 'dense_w'):
            assert np.array_equal(getattr(got, k), getattr(m, k))
        assert got.dense_b == m.dense_b and got.activation == 'linear' and got.recurrent_activation == 'hard_sigmoid'
    got = model_from_keras_h5(_keras_like_file(m, ract='sigmoid'))
    assert got.recurrent_activation == 'sigmoid'
    # Keras >= 2.2 nests the layer list one level deeper
    f = _keras_like_file(m)
    import json
    cfg = json.loads(f.attrs['model_config'])
    cfg['config'] = {'name': 'sequential_1', 'layers': cfg['config']}
    f.attrs['model_config'] = json.dumps(cfg)
    assert model_from_keras_h5(f).hidden == 20


def test_net_file_that_is_not_hdf5_fails_loudly(tmp_path):
    from mycroft_precise_b200.model_io import load_weights
    p = tmp_path / 'm.net'
    p.write_bytes(b'\x89HDF\r\n\x1a\n')                      # signature only: truncated
    with pytest.raises(Exception):
        load_weights(str(p))


def _model_config(ract='hard_sigmoid'):
    return json.dumps({'class_name': 'Sequential', 'config': [
        {'class_name': 'GRU', 'config': {'name': 'net', 'units': 20, 'activation': 'linear', 'recurrent_activation': ract}},
        {'class_name': 'Dense', 'config': {'name': 'dense_1', 'units': 1, 'activation': 'sigmoid'}}]})


@pytest.mark.parametrize('split', [False, True])
def test_builtin_hdf5_reader_on_keras_layout(tmp_path, split):
    """mycroft_precise_b200/h5_import.py against a file written from the same format specification by tests/h5_writer.py
    (groups, nested groups, float datasets, string / string-array attributes, continuation blocks)."""
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    from h5_writer import keras_model_file
    from mycroft_precise_b200.h5_import import H5File, H5FormatError
    from mycroft_precise_b200.model_io import model_from_keras_h5
    m = GruModel.random(13, 20, seed=6)
    data = keras_model_file(m.kernel, m.recurrent, m.bias, m.dense_w, m.dense_b, _model_config('sigmoid'), split_headers=split)
    p = tmp_path / 'model.net'
    p.write_bytes(data)
    with H5File(str(p)) as f:
        assert sorted(f.keys()) == ['model_weights', 'optimizer_weights']
        assert bytes(f.attrs['keras_version']) == b'2.1.5'
        assert list(f['model_weights'].attrs['layer_names']) == [b'net', b'dense_1']
        assert f['model_weights/net/net/kernel:0'].shape == (13, 60)
        assert f['optimizer_weights'].keys() == []
        got = model_from_keras_h5(f)
    for k in ('kernel', 'recurrent', 'bias', 'dense_w'):
        assert np.array_equal(getattr(got, k), getattr(m, k))
    assert got.dense_b == m.dense_b and got.recurrent_activation == 'sigmoid' and got.activation == 'linear'
    with pytest.raises(H5FormatError):
        H5File(data=b'not an hdf5 file' * 64)
    with pytest.raises(H5FormatError):
        H5File(data=data[:8] + b'\x02' + data[9:])           # superblock version 2 (libver='latest'): refused, not misparsed


def test_load_weights_net_without_h5py_uses_builtin_reader(tmp_path):
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    from h5_writer import keras_model_file
    try:
        import h5py  # noqa: F401
        pytest.skip('h5py present: load_net prefers it')
    except ImportError:
        pass
    m = GruModel.random(13, 20, seed=7)
    p = tmp_path / 'm.net'
    p.write_bytes(keras_model_file(m.kernel, m.recurrent, m.bias, m.dense_w, m.dense_b, _model_config()))
    got = load_weights(str(p))
    assert np.array_equal(got.kernel, m.kernel) and got.recurrent_activation == 'hard_sigmoid'

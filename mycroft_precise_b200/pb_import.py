"""Weights from a frozen TensorFlow GraphDef (``.pb``), the production model format of the reference
(``precise-convert``, precise/scripts/convert.py:41-85; loaded by TensorFlowRunner, network_runner.py:45-74).

TensorFlow is not needed: a GraphDef is plain protobuf wire format and only five Const tensors matter.
Field numbers (tensorflow/core/framework/*.proto, TF 1.13):
  GraphDef.node = 1;  NodeDef.name = 1, .op = 2, .attr = 5 (map<string, AttrValue>: key = 1, value = 2)
  AttrValue.tensor = 8;  TensorProto.dtype = 1, .tensor_shape = 2, .tensor_content = 4, .float_val = 5
  TensorShapeProto.dim = 2;  Dim.size = 1
Keras names the variables <layer>/kernel, <layer>/recurrent_kernel, <layer>/bias; the reference's GRU layer is named
'net' (precise/model.py:81) and is followed by one Dense layer.  convert_variables_to_constants keeps those names.
"""
import struct

import numpy as np

from .model_io import GruModel


def _varint(buf, pos):
    r = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        r |= (b & 0x7f) << shift
        if not b & 0x80:
            return r, pos
        shift += 7


def _fields(buf):
    """Yield (field_number, wire_type, value) for one message; length-delimited values are memoryviews."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = bytes(buf[pos:pos + 8]); pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]; pos += ln
        elif wt == 5:
            v = bytes(buf[pos:pos + 4]); pos += 4
        else:
            raise ValueError('unsupported protobuf wire type %d' % wt)
        yield fno, wt, v


def _tensor(buf) -> np.ndarray:
    dtype, shape, content, floats = 1, [], None, []
    for fno, wt, v in _fields(buf):
        if fno == 1:
            dtype = v
        elif fno == 2:
            for f2, _, dim in _fields(v):
                if f2 == 2:
                    size = 0
                    for f3, _, s in _fields(dim):
                        if f3 == 1:
                            size = s
                    shape.append(size)
        elif fno == 4:
            content = bytes(v)
        elif fno == 5:
            if wt == 2:                                   # packed
                floats += list(struct.unpack('<%df' % (len(v) // 4), bytes(v)))
            else:
                floats.append(struct.unpack('<f', v)[0])
    if dtype != 1:                                        # DT_FLOAT
        return None
    n = int(np.prod(shape)) if shape else 1
    if content is not None:
        a = np.frombuffer(content, dtype='<f4')
    elif len(floats) == 1 and n > 1:
        a = np.full(n, floats[0], dtype=np.float32)       # TF splat encoding
    else:
        a = np.asarray(floats, dtype=np.float32)
    return a.reshape(shape) if a.size == n else None


def read_const_tensors(path: str) -> dict:
    """name -> float32 ndarray for every float Const node of the GraphDef."""
    data = memoryview(open(path, 'rb').read())
    out = {}
    for fno, wt, node in _fields(data):
        if fno != 1 or wt != 2:
            continue
        name, op, tensor = None, None, None
        for f, w, v in _fields(node):
            if f == 1:
                name = bytes(v).decode()
            elif f == 2:
                op = bytes(v).decode()
            elif f == 5:
                key, val = None, None
                for f2, _, v2 in _fields(v):
                    if f2 == 1:
                        key = bytes(v2).decode()
                    elif f2 == 2:
                        val = v2
                if key == 'value' and val is not None:
                    for f3, _, v3 in _fields(val):
                        if f3 == 8:
                            tensor = _tensor(v3)
        if op == 'Const' and tensor is not None and name:
            out[name] = tensor
    return out


def load_pb(path: str) -> GruModel:
    t = read_const_tensors(path)

    def find(suffix, ndim):
        c = [k for k, v in t.items() if k.endswith(suffix) and v.ndim == ndim]
        return c

    rec = find('/recurrent_kernel', 2)
    if len(rec) != 1:
        raise ValueError('expected exactly one GRU layer in %s, found %s' % (path, rec))
    layer = rec[0][:-len('/recurrent_kernel')]
    kernel, recurrent, bias = t[layer + '/kernel'], t[rec[0]], t[layer + '/bias']
    H = recurrent.shape[0]
    if bias.ndim == 2:                                    # reset_after=True layout [2, 3H]: not the reference's network
        raise ValueError('GRU with reset_after=True is not supported')
    dense = [k[:-len('/kernel')] for k, v in t.items() if k.endswith('/kernel') and v.ndim == 2 and v.shape == (H, 1)]
    if len(dense) != 1:
        raise ValueError('expected one Dense(1) layer after the GRU, found %s' % dense)
    return GruModel(kernel, recurrent, bias, t[dense[0] + '/kernel'].reshape(-1), t[dense[0] + '/bias'].reshape(-1)[0])

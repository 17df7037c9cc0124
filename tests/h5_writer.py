"""Test helper: writes the subset of HDF5 that h5py's default format uses for Keras model files (superblock 0, symbol-table
groups, version-1 object headers, contiguous datasets, version-1 attribute messages), straight from the published HDF5
file-format specification.  Only used to exercise mycroft_precise_b200/h5_import.py -- it is not a general HDF5 writer."""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


def _pad8(b):
    return b + b'\x00' * (-len(b) % 8)


def _dtype_msg(dt):
    dt = np.dtype(dt)
    if dt.kind == 'f':
        exp_bits, man_bits = {4: (8, 23), 8: (11, 52)}[dt.itemsize]
        props = struct.pack('<HHBBBBI', 0, 8 * dt.itemsize, man_bits, exp_bits, 0, man_bits, (1 << (exp_bits - 1)) - 1)
        return struct.pack('<BBBBI', 0x11, 0x20, 8 * dt.itemsize - 1, 0, dt.itemsize) + props
    if dt.kind in 'iu':
        return struct.pack('<BBBBI', 0x10, 0x08 if dt.kind == 'i' else 0, 0, 0, dt.itemsize) + struct.pack('<HH', 0, 8 * dt.itemsize)
    if dt.kind == 'S':
        return struct.pack('<BBBBI', 0x13, 0, 0, 0, dt.itemsize)
    raise ValueError(dt)


def _space_msg(shape):
    return struct.pack('<BBBBI', 1, len(shape), 0, 0, 0) + b''.join(struct.pack('<Q', s) for s in shape)


def _msg(mtype, payload):
    payload = _pad8(payload)
    return struct.pack('<HHBBBB', mtype, len(payload), 0, 0, 0, 0) + payload


def _attr_msg(name, value):
    arr = np.asarray(value)
    if arr.dtype.kind == 'U':
        arr = np.char.encode(arr, 'utf-8')
    nm = name.encode('utf-8') + b'\x00'
    dt, sp = _dtype_msg(arr.dtype), _space_msg(arr.shape)
    body = struct.pack('<BBHHH', 1, 0, len(nm), len(dt), len(sp)) + _pad8(nm) + _pad8(dt) + _pad8(sp) + arr.tobytes()
    return _msg(0x000C, body)


class Writer:
    def __init__(self):
        self.buf = bytearray(b'\x00' * 96)                   # superblock placeholder (56 + 40 bytes)

    def _alloc(self, blob):
        while len(self.buf) % 8:
            self.buf.append(0)
        addr = len(self.buf)
        self.buf += blob
        return addr

    def _object_header(self, msgs, split=False):
        """version-1 object header; split=True pushes the last message into a continuation block (as libhdf5 does when an
        attribute is added after creation)."""
        if split and len(msgs) > 1:
            tail = msgs[-1]
            cont_addr = self._alloc(tail)
            msgs = msgs[:-1] + [_msg(0x0010, struct.pack('<QQ', cont_addr, len(tail)))]
            n = len(msgs) + 1
        else:
            n = len(msgs)
        body = b''.join(msgs)
        return self._alloc(struct.pack('<BBHII', 1, 0, n, 1, len(body)) + b'\x00' * 4 + body)

    def dataset(self, array, attrs=None, split=False):
        arr = np.ascontiguousarray(array)
        data_addr = self._alloc(arr.tobytes())
        msgs = [_msg(0x0001, _space_msg(arr.shape)), _msg(0x0003, _dtype_msg(arr.dtype)),
                _msg(0x0008, struct.pack('<BBQQ', 3, 1, data_addr, arr.nbytes))]
        msgs += [_attr_msg(k, v) for k, v in (attrs or {}).items()]
        return self._object_header(msgs, split)

    def group(self, children, attrs=None, split=False):
        names = sorted(children)
        heap_data = bytearray(b'\x00' * 8)                   # offset 0: the empty name
        offs = {}
        for nme in names:
            offs[nme] = len(heap_data)
            heap_data += nme.encode('utf-8') + b'\x00'
            while len(heap_data) % 8:
                heap_data.append(0)
        seg = self._alloc(bytes(heap_data))
        heap = self._alloc(b'HEAP' + struct.pack('<BBBBQQQ', 0, 0, 0, 0, len(heap_data), UNDEF, seg))
        snod = b'SNOD' + struct.pack('<BBH', 1, 0, len(names))
        for nme in names:
            snod += struct.pack('<QQII', offs[nme], children[nme], 0, 0) + b'\x00' * 16
        snod_addr = self._alloc(snod)
        tree = b'TREE' + struct.pack('<BBHQQ', 0, 0, 1, UNDEF, UNDEF) + struct.pack('<QQQ', 0, snod_addr, offs[names[-1]] if names else 0)
        tree_addr = self._alloc(tree)
        msgs = [_msg(0x0011, struct.pack('<QQ', tree_addr, heap))] + [_attr_msg(k, v) for k, v in (attrs or {}).items()]
        return self._object_header(msgs, split), tree_addr, heap

    def finish(self, root):
        root_addr, tree, heap = root
        sb = b'\x89HDF\r\n\x1a\n' + struct.pack('<BBBBBBBBHHI', 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
        sb += struct.pack('<QQQQ', 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack('<QQII', 0, root_addr, 1, 0) + struct.pack('<QQ', tree, heap)
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


def keras_model_file(kernel, recurrent, bias, dense_w, dense_b, model_config_json, split_headers=False):
    """Bytes of a file laid out like keras.models.Model.save() of precise/model.py:72-82 (layers 'net' and 'dense_1')."""
    w = Writer()

    def layer(name, weights):
        inner, _, _ = w.group({k: w.dataset(v) for k, v in weights.items()})
        names = np.array([('%s/%s' % (name, k)).encode() for k in weights])
        return w.group({name: inner}, {'weight_names': names}, split=split_headers)[0]

    net = layer('net', {'kernel:0': kernel, 'recurrent_kernel:0': recurrent, 'bias:0': bias})
    dense = layer('dense_1', {'kernel:0': np.asarray(dense_w, np.float32).reshape(-1, 1), 'bias:0': np.asarray([dense_b], np.float32)})
    mw, _, _ = w.group({'net': net, 'dense_1': dense},
                       {'layer_names': np.array([b'net', b'dense_1']), 'backend': np.bytes_(b'tensorflow'), 'keras_version': np.bytes_(b'2.1.5')})
    opt, _, _ = w.group({})
    root = w.group({'model_weights': mw, 'optimizer_weights': opt},
                   {'keras_version': np.bytes_(b'2.1.5'), 'backend': np.bytes_(b'tensorflow'),
                    'model_config': np.bytes_(model_config_json.encode('utf-8'))}, split=split_headers)
    return w.finish(root)

"""Minimal read-only HDF5 parser for Keras 2.x model files (``.net``), for machines without h5py.

Covers what ``keras.models.Model.save`` / ``save_weights`` write through h5py with its default (``libver='earliest'``) file
format -- the only thing ``model_io.model_from_keras_h5`` needs:

  * superblock version 0 / 1, 8-byte offsets and lengths;
  * old-style groups (symbol-table message -> v1 B-tree of SNOD nodes + local heap);
  * version-1 object headers with continuation blocks;
  * datasets: contiguous or compact layout (chunked without filters too), little-endian IEEE floats / integers, fixed strings;
  * attributes (message versions 1-3): scalars and 1-D arrays of floats, integers, fixed-length strings, and variable-length
    strings through the global heap.

Anything else (superblock >= 2, link-message groups, compressed chunks, compound types) raises ``H5FormatError`` -- use h5py or
the frozen ``.pb`` in that case.  Written from the published HDF5 file-format specification (version 1.1 / 2.0 documents); no
Keras-written file exists in this build image to test against, so the parser is exercised by files produced by the writer
in ``tests/h5_writer.py``, which follows the same specification (parity with real files: unpinned).
"""
import struct

import numpy as np

SIGNATURE = b'\x89HDF\r\n\x1a\n'
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5FormatError(ValueError):
    pass


class _Reader:
    def __init__(self, buf):
        self.buf = buf

    def u(self, off, n):
        return int.from_bytes(self.buf[off:off + n], 'little')

    def bytes(self, off, n):
        if off < 0 or off + n > len(self.buf):
            raise H5FormatError('read of %d bytes at %d beyond the end of the file' % (n, off))
        return self.buf[off:off + n]


def _pad8(n):
    return (n + 7) & ~7


class _Datatype:
    def __init__(self, raw):
        if len(raw) < 8:
            raise H5FormatError('truncated datatype message')
        self.cls = raw[0] & 0x0F
        self.version = raw[0] >> 4
        self.bits = raw[1] | (raw[2] << 8) | (raw[3] << 16)
        self.size = struct.unpack_from('<I', raw, 4)[0]
        self.props = raw[8:]
        self.vlen_string = False
        if self.cls == 9:                                   # variable length: bits 0-3 type (1 = string)
            self.vlen_string = (self.bits & 0x0F) == 1
            if not self.vlen_string:
                raise H5FormatError('variable-length sequences are not supported')

    def numpy(self):
        order = '>' if (self.bits & 1) else '<'
        if self.cls == 1:
            if self.size not in (2, 4, 8):
                raise H5FormatError('float of %d bytes' % self.size)
            return np.dtype(order + 'f%d' % self.size)
        if self.cls == 0:
            signed = bool(self.bits & 0x08)
            return np.dtype(order + ('i' if signed else 'u') + str(self.size))
        if self.cls == 3:
            return np.dtype('S%d' % self.size)
        raise H5FormatError('datatype class %d is not supported' % self.cls)


def _parse_dataspace(raw):
    ver, rank, flags = raw[0], raw[1], raw[2]
    if ver == 1:
        off = 8
    elif ver == 2:
        off = 4
        if raw[3] == 2:                                      # null dataspace
            return None
    else:
        raise H5FormatError('dataspace message version %d' % ver)
    return tuple(struct.unpack_from('<Q', raw, off + 8 * i)[0] for i in range(rank))


class H5File:
    """``H5File(path)`` or ``H5File(data=bytes)``; behaves like the root group (mapping of groups / numpy arrays, ``.attrs``)."""

    def __init__(self, path=None, data=None):
        if data is None:
            with open(path, 'rb') as fh:
                data = fh.read()
        self._r = _Reader(data)
        base = 0
        while True:                                          # the superblock may sit at 0, 512, 1024, ...
            if data[base:base + 8] == SIGNATURE:
                break
            base = 512 if base == 0 else base * 2
            if base + 8 > len(data):
                raise H5FormatError('not an HDF5 file (signature not found)')
        r = self._r
        ver = r.u(base + 8, 1)
        if ver not in (0, 1):
            raise H5FormatError('superblock version %d: written with libver="latest"? use h5py for this file' % ver)
        if r.u(base + 13, 1) != 8 or r.u(base + 14, 1) != 8:
            raise H5FormatError('only 8-byte offsets and lengths are supported')
        p = base + 24 + (4 if ver == 1 else 0)
        self._base = r.u(p, 8)
        root_entry = p + 32
        self._root = self._group_from_entry(root_entry)

    # ---- mapping interface of the root group
    def __getitem__(self, k):
        return self._root[k]

    def __contains__(self, k):
        return k in self._root

    def keys(self):
        return self._root.keys()

    @property
    def attrs(self):
        return self._root.attrs

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    # ---- low level
    def _addr(self, a):
        if a == UNDEF:
            raise H5FormatError('undefined address')
        return a + self._base

    def _group_from_entry(self, off):
        r = self._r
        ohdr = r.u(off + 8, 8)
        return self._object(self._addr(ohdr))

    def _messages(self, addr):
        """(type, flags, payload bytes) of a version-1 object header, following continuation blocks."""
        r = self._r
        if r.bytes(addr, 4) == b'OHDR':
            raise H5FormatError('version-2 object header: file written with libver="latest"; use h5py')
        if r.u(addr, 1) != 1:
            raise H5FormatError('object header version %d' % r.u(addr, 1))
        n_msgs = r.u(addr + 2, 2)
        size = r.u(addr + 8, 4)
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < n_msgs:
            pos, left = blocks.pop(0)
            end = pos + left
            while pos + 8 <= end and len(out) < n_msgs:
                mtype, msize, mflags = r.u(pos, 2), r.u(pos + 2, 2), r.u(pos + 4, 1)
                payload = r.bytes(pos + 8, msize)
                pos += 8 + msize
                if mtype == 0x0010:                          # continuation
                    blocks.append((self._addr(struct.unpack_from('<Q', payload, 0)[0]), struct.unpack_from('<Q', payload, 8)[0]))
                out.append((mtype, mflags, payload))
        return out

    def _object(self, addr):
        msgs = self._messages(addr)
        attrs = {}
        stab = dtype = space = layout = None
        filtered = False
        for mtype, mflags, p in msgs:
            if mtype == 0x0011:
                stab = (struct.unpack_from('<Q', p, 0)[0], struct.unpack_from('<Q', p, 8)[0])
            elif mtype == 0x0003:
                dtype = _Datatype(p)
            elif mtype == 0x0001:
                space = _parse_dataspace(p)
            elif mtype == 0x0008:
                layout = p
            elif mtype == 0x000B:
                filtered = True
            elif mtype == 0x000C:
                k, v = self._attribute(p)
                attrs[k] = v
            elif mtype in (0x0002, 0x0006):
                raise H5FormatError('link-message (new style) group: use h5py for this file')
        if stab is not None:
            return H5Group(self, stab[0], stab[1], attrs)
        if dtype is not None and layout is not None:
            if filtered:
                raise H5FormatError('filtered (compressed) dataset: use h5py for this file')
            return self._dataset(dtype, space, layout)
        return H5Group(self, None, None, attrs)              # an empty group

    def _attribute(self, p):
        ver = p[0]
        name_sz, dt_sz, sp_sz = struct.unpack_from('<HHH', p, 2)
        if ver == 1:
            off = 8
            name = p[off:off + name_sz]; off += _pad8(name_sz)
            dt = p[off:off + dt_sz]; off += _pad8(dt_sz)
            sp = p[off:off + sp_sz]; off += _pad8(sp_sz)
        elif ver in (2, 3):
            off = 8 + (1 if ver == 3 else 0)
            name = p[off:off + name_sz]; off += name_sz
            dt = p[off:off + dt_sz]; off += dt_sz
            sp = p[off:off + sp_sz]; off += sp_sz
        else:
            raise H5FormatError('attribute message version %d' % ver)
        name = name.split(b'\x00')[0].decode('utf-8')
        return name, self._decode(_Datatype(dt), _parse_dataspace(sp), p[off:])

    def _decode(self, dtype, shape, raw):
        if shape is None:
            return None
        n = int(np.prod(shape)) if shape else 1
        if dtype.vlen_string:
            out = []
            for i in range(n):
                ln, gaddr, idx = struct.unpack_from('<IQI', raw, 16 * i)
                out.append(self._global_heap(self._addr(gaddr), idx)[:ln])
            return out[0] if not shape else np.array(out, dtype=object)
        dt = dtype.numpy()
        arr = np.frombuffer(raw[:n * dt.itemsize], dtype=dt).reshape(shape if shape else ())
        if dt.kind == 'S':
            arr = np.char.rstrip(arr, b'\x00') if shape else np.bytes_(bytes(arr).rstrip(b'\x00'))
        elif not shape:
            arr = arr[()]
        else:
            arr = arr.copy()
        return arr

    def _global_heap(self, addr, index):
        r = self._r
        if r.bytes(addr, 4) != b'GCOL':
            raise H5FormatError('global heap collection expected at %d' % addr)
        size = r.u(addr + 8, 8)
        pos, end = addr + 16, addr + size
        while pos + 16 <= end:
            idx, osize = r.u(pos, 2), r.u(pos + 8, 8)
            if idx == 0:
                break
            if idx == index:
                return r.bytes(pos + 16, osize)
            pos += 16 + _pad8(osize)
        raise H5FormatError('global heap object %d not found' % index)

    def _dataset(self, dtype, shape, layout):
        r = self._r
        ver = layout[0]
        dt = dtype.numpy()
        n = int(np.prod(shape)) if shape else 1
        if ver == 3:
            cls = layout[1]
            if cls == 1:
                addr = struct.unpack_from('<Q', layout, 2)[0]
                raw = b'' if addr == UNDEF else r.bytes(self._addr(addr), n * dt.itemsize)
            elif cls == 0:
                sz = struct.unpack_from('<H', layout, 2)[0]
                raw = layout[4:4 + sz]
            elif cls == 2:
                nd = layout[2]
                btree = struct.unpack_from('<Q', layout, 3)[0]
                cdims = struct.unpack_from('<%dI' % nd, layout, 11)
                return self._chunked(dt, shape, self._addr(btree), cdims[:-1])
            else:
                raise H5FormatError('data layout class %d' % cls)
        elif ver in (1, 2):
            nd, cls = layout[1], layout[2]
            off = 8
            if cls == 1:
                addr = struct.unpack_from('<Q', layout, off)[0]
                raw = r.bytes(self._addr(addr), n * dt.itemsize)
            elif cls == 0:
                off += 4 * nd
                sz = struct.unpack_from('<I', layout, off)[0]
                raw = layout[off + 4:off + 4 + sz]
            else:
                raise H5FormatError('chunked layout of message version %d' % ver)
        else:
            raise H5FormatError('data layout message version %d' % ver)
        if len(raw) < n * dt.itemsize:
            return np.zeros(shape if shape else (), dt)      # never written: fill value
        arr = np.frombuffer(raw[:n * dt.itemsize], dtype=dt).reshape(shape if shape else ()).copy()
        return arr.astype(dt.newbyteorder('=')) if dt.byteorder == '>' else arr

    def _chunked(self, dt, shape, btree, cdims):
        out = np.zeros(shape, dt)
        nd = len(shape)
        r = self._r

        def walk(addr):
            if r.bytes(addr, 4) != b'TREE' or r.u(addr + 4, 1) != 1:
                raise H5FormatError('chunk B-tree expected at %d' % addr)
            level, used = r.u(addr + 5, 1), r.u(addr + 6, 2)
            pos = addr + 24
            ksz = 8 + 8 * (nd + 1)
            for _ in range(used):
                csize, fmask = r.u(pos, 4), r.u(pos + 4, 4)
                offs = [r.u(pos + 8 + 8 * i, 8) for i in range(nd)]
                child = self._addr(r.u(pos + ksz, 8))
                pos += ksz + 8
                if level > 0:
                    walk(child)
                    continue
                if fmask:
                    raise H5FormatError('filtered chunk')
                chunk = np.frombuffer(r.bytes(child, csize), dtype=dt)[:int(np.prod(cdims))].reshape(cdims)
                sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, shape))
                out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]
        walk(btree)
        return out


class H5Group:
    def __init__(self, f, btree, heap, attrs):
        self._f, self._btree, self._heap = f, btree, heap
        self.attrs = attrs
        self._entries = None
        self._cache = {}

    def _load(self):
        if self._entries is not None:
            return
        self._entries = {}
        if self._btree is None:
            return
        f, r = self._f, self._f._r
        heap = f._addr(self._heap)
        if r.bytes(heap, 4) != b'HEAP':
            raise H5FormatError('local heap expected at %d' % heap)
        seg = f._addr(r.u(heap + 24, 8))

        def name_at(off):
            end = r.buf.index(b'\x00', seg + off)
            return r.buf[seg + off:end].decode('utf-8')

        def walk(addr):
            if r.bytes(addr, 4) != b'TREE' or r.u(addr + 4, 1) != 0:
                raise H5FormatError('group B-tree expected at %d' % addr)
            level, used = r.u(addr + 5, 1), r.u(addr + 6, 2)
            pos = addr + 24 + 8                              # skip key 0
            for _ in range(used):
                child = f._addr(r.u(pos, 8))
                pos += 16                                    # child + next key
                if level > 0:
                    walk(child)
                    continue
                if r.bytes(child, 4) != b'SNOD':
                    raise H5FormatError('symbol node expected at %d' % child)
                for i in range(r.u(child + 6, 2)):
                    e = child + 8 + 40 * i
                    self._entries[name_at(r.u(e, 8))] = f._addr(r.u(e + 8, 8))
        walk(f._addr(self._btree))

    def keys(self):
        self._load()
        return list(self._entries.keys())

    def __contains__(self, k):
        self._load()
        return k.split('/')[0] in self._entries if k else False

    def __getitem__(self, k):
        self._load()
        head, _, rest = k.strip('/').partition('/')
        if head not in self._entries:
            raise KeyError(k)
        if head not in self._cache:
            self._cache[head] = self._f._object(self._entries[head])
        obj = self._cache[head]
        return obj[rest] if rest else obj

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return len(self.keys())

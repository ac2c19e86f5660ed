"""Minimal pure-Python HDF5 reader for The Payne's network files (no h5py, no libhdf5).

The reference stores its emulators as HDF5 written by h5py with default settings
(Payne/train/trainflux.py:221-235, :560-570; read back by Payne/train/NNmodels.py:44-89 and
Payne/predict/predictspec.py:43-59): superblock version 0, old-style groups (symbol-table
message -> v1 B-tree of symbol nodes + local heap), version-1 object headers, datasets that are
contiguous or chunked + gzip (``compression='gzip'`` on every ``model/*`` tensor), fixed-length
byte strings for ``label_i``.  That subset of the HDF5 1.x file-format specification is what
this module parses:

    read(path) -> {"xmin": ndarray, "model/lin1.weight": ndarray, ...}

Not supported (a clear IOError says so): superblock >= 2 / new-style groups (libver='latest'),
compound / variable-length / reference datatypes, external storage, filters other than
deflate, shuffle and fletcher32.

``write(path, datasets)`` emits the same subset (contiguous, or chunked + gzip).  It exists so the
reader can be round-trip tested on machines without the reference tree; files it writes follow
the specification but have not been opened with libhdf5 in this image.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

SIG = b'\x89HDF\r\n\x1a\n'
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(IOError):
    pass


# ------------------------------------------------------------------------------------ reader
class _File:
    def __init__(self, buf):
        self.b = buf
        if buf[:8] != SIG:
            raise H5Error('not an HDF5 file (signature at offset 0 missing; user blocks are not supported)')
        ver = buf[8]
        if ver > 1:
            raise H5Error('HDF5 superblock version %d (libver="latest") is not supported; rewrite the file with '
                          'h5py defaults or convert it to .npz' % ver)
        self.O, self.L = buf[13], buf[14]
        if self.O != 8 or self.L != 8:
            raise H5Error('only 8-byte offsets/lengths are supported (got %d/%d)' % (self.O, self.L))
        p = 24 + (4 if ver == 1 else 0)
        self.base = self.u64(p)
        p += 4 * 8                     # base, free-space, end-of-file, driver-info addresses
        # root group symbol table entry
        self.root_header = self.u64(p + 8)
        self.root_cache = self.u32(p + 16)
        self.root_scratch = (self.u64(p + 24), self.u64(p + 32))

    def u16(self, p): return struct.unpack_from('<H', self.b, p)[0]
    def u32(self, p): return struct.unpack_from('<I', self.b, p)[0]
    def u64(self, p): return struct.unpack_from('<Q', self.b, p)[0]

    # -- object headers -------------------------------------------------------------------
    def messages(self, addr):
        """[(type, payload bytes)] of a version-1 object header, continuation blocks followed."""
        a = addr + self.base
        if self.b[a:a + 4] == b'OHDR':
            raise H5Error('version-2 object headers are not supported')
        if self.b[a] != 1:
            raise H5Error('unexpected object header version %d at %#x' % (self.b[a], a))
        nmsg, size = self.u16(a + 2), self.u32(a + 8)
        blocks = [(a + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            p, n = blocks.pop(0)
            end = p + n
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize = self.u16(p), self.u16(p + 2)
                body = self.b[p + 8:p + 8 + msize]
                if mtype == 0x0010:
                    blocks.append((self.u64(p + 8) + self.base, self.u64(p + 16)))
                out.append((mtype, body))
                p += 8 + msize
        return out

    # -- groups ---------------------------------------------------------------------------
    def heap_name(self, heap_addr, off):
        h = heap_addr + self.base
        if self.b[h:h + 4] != b'HEAP':
            raise H5Error('local heap signature missing at %#x' % h)
        data = self.u64(h + 24) + self.base
        end = self.b.index(b'\x00', data + off)
        return self.b[data + off:end].decode('utf-8')

    def group_entries(self, btree, heap):
        """[(name, object header address)] under a symbol-table group."""
        out = []
        stack = [btree]
        while stack:
            n = stack.pop() + self.base
            sig = self.b[n:n + 4]
            if sig == b'TREE':
                if self.b[n + 4] != 0:
                    raise H5Error('group B-tree node of the wrong type at %#x' % n)
                used = self.u16(n + 6)
                p = n + 8 + 16          # past the sibling addresses
                for i in range(used):
                    stack.append(self.u64(p + 8 + i * 16))      # key (8), child (8) interleaved
            elif sig == b'SNOD':
                nsym = self.u16(n + 6)
                for i in range(nsym):
                    e = n + 8 + 40 * i
                    out.append((self.heap_name(heap, self.u64(e)), self.u64(e + 8)))
            else:
                raise H5Error('unknown group node signature %r at %#x' % (sig, n))
        return sorted(out)

    # -- datasets -------------------------------------------------------------------------
    def dataspace(self, m):
        ver, rank, flags = m[0], m[1], m[2]
        if ver == 1:
            p = 8
        elif ver == 2:
            p = 4
            if m[3] == 2:
                return None            # null dataspace
        else:
            raise H5Error('dataspace message version %d' % ver)
        return tuple(struct.unpack_from('<Q', m, p + 8 * i)[0] for i in range(rank))

    def datatype(self, m):
        cls, bits0, size = m[0] & 15, m[1], struct.unpack_from('<I', m, 4)[0]
        order = '>' if (bits0 & 1) else '<'
        if cls == 0:
            return np.dtype('%s%s%d' % (order, 'i' if bits0 & 8 else 'u', size))
        if cls == 1:
            return np.dtype('%sf%d' % (order, size))
        if cls == 3:
            return np.dtype('S%d' % size)
        raise H5Error('datatype class %d (compound / variable-length / ...) is not supported' % cls)

    def filters(self, m):
        ver, n = m[0], m[1]
        p = 8 if ver == 1 else 2
        out = []
        for _ in range(n):
            fid = struct.unpack_from('<H', m, p)[0]
            if ver == 1 or fid >= 256:
                nlen = struct.unpack_from('<H', m, p + 2)[0]
                p += 4
            else:
                nlen = 0
                p += 2
            ncd = struct.unpack_from('<H', m, p + 2)[0]
            p += 4
            p += (nlen + 7) // 8 * 8 if ver == 1 else nlen
            cd = struct.unpack_from('<%dI' % ncd, m, p)
            p += 4 * ncd
            if ver == 1 and ncd % 2:
                p += 4
            out.append((fid, cd))
        return out

    def chunks(self, btree, ndim):
        """[(offsets, filter mask, address, stored size)] of a chunked dataset (v1 B-tree, type 1)."""
        out = []
        stack = [btree]
        ksz = 8 + 8 * (ndim + 1)
        while stack:
            n = stack.pop() + self.base
            if self.b[n:n + 4] != b'TREE' or self.b[n + 4] != 1:
                raise H5Error('chunk B-tree node expected at %#x' % n)
            level, used = self.b[n + 5], self.u16(n + 6)
            p = n + 8 + 16
            for i in range(used):
                k = p + i * (ksz + 8)
                child = self.u64(k + ksz)
                if level > 0:
                    stack.append(child)
                else:
                    offs = struct.unpack_from('<%dQ' % ndim, self.b, k + 8)
                    out.append((offs, self.u32(k + 4), child, self.u32(k)))
        return out

    def dataset(self, msgs):
        shape = dtype = layout = None
        filt = []
        for t, m in msgs:
            if t == 0x0001:
                shape = self.dataspace(m)
            elif t == 0x0003:
                dtype = self.datatype(m)
            elif t == 0x0008:
                layout = m
            elif t == 0x000B:
                filt = self.filters(m)
        if dtype is None or layout is None:
            return None
        if shape is None:
            return np.zeros(0, dtype=dtype.newbyteorder('='))
        count = int(np.prod(shape)) if shape else 1
        ver = layout[0]
        if ver != 3:
            raise H5Error('data layout message version %d is not supported (HDF5 >= 1.6.3 writes version 3)' % ver)
        cls = layout[1]
        if cls == 0:                                   # compact
            n = struct.unpack_from('<H', layout, 2)[0]
            raw = bytes(layout[4:4 + n])
        elif cls == 1:                                 # contiguous
            addr, n = struct.unpack_from('<QQ', layout, 2)
            raw = b'\x00' * (count * dtype.itemsize) if addr == UNDEF else bytes(
                self.b[addr + self.base:addr + self.base + n])
        elif cls == 2:                                 # chunked
            nd = layout[2] - 1
            bt = struct.unpack_from('<Q', layout, 3)[0]
            cdims = struct.unpack_from('<%dI' % (nd + 1), layout, 11)[:nd]
            arr = np.zeros(shape, dtype=dtype)
            if bt != UNDEF:
                for offs, mask, addr, size in self.chunks(bt, nd):
                    data = bytes(self.b[addr + self.base:addr + self.base + size])
                    for i in range(len(filt) - 1, -1, -1):     # undo the pipeline back to front
                        if mask & (1 << i):
                            continue
                        fid, cd = filt[i]
                        if fid == 1:
                            data = zlib.decompress(data)
                        elif fid == 2:
                            es = cd[0] if cd else dtype.itemsize
                            n = len(data) // es
                            data = np.frombuffer(data[:n * es], np.uint8).reshape(es, n).T.tobytes() + data[n * es:]
                        elif fid == 3:
                            data = data[:-4]
                        else:
                            raise H5Error('HDF5 filter id %d is not supported' % fid)
                    chunk = np.frombuffer(data, dtype=dtype, count=int(np.prod(cdims))).reshape(cdims)
                    sel = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, shape))
                    arr[sel] = chunk[tuple(slice(0, s.stop - s.start) for s in sel)]
            return arr.astype(dtype.newbyteorder('='))
        else:
            raise H5Error('data layout class %d' % cls)
        arr = np.frombuffer(raw, dtype=dtype, count=count).reshape(shape)
        return arr.astype(dtype.newbyteorder('='))

    # -- walk -----------------------------------------------------------------------------
    def walk(self):
        out = {}
        seen = set()

        def visit(prefix, header):
            if header in seen:
                return
            seen.add(header)
            msgs = self.messages(header)
            st = [m for t, m in msgs if t == 0x0011]
            if st:
                bt, heap = struct.unpack_from('<QQ', st[0], 0)
                for name, child in self.group_entries(bt, heap):
                    visit(prefix + name + '/', child)
                return
            if any(t in (0x0002, 0x0006) for t, _ in msgs) and not any(t == 0x0008 for t, _ in msgs):
                raise H5Error('new-style (link message) groups are not supported')
            d = self.dataset(msgs)
            if d is not None:
                out[prefix.rstrip('/')] = d

        visit('', self.root_header)
        return out


def read(path):
    """All datasets of an HDF5 file as ``{"group/name": ndarray}`` (scalars come back 0-d)."""
    with open(path, 'rb') as f:
        buf = f.read()
    return _File(buf).walk()


# ------------------------------------------------------------------------------------ writer
def _dtype_msg(dt):
    dt = np.dtype(dt)
    if dt.kind == 'f':
        # IEEE little-endian: class 1 version 1; bit fields: byte order 0, mantissa normalisation 2 (implied msb),
        # sign location in byte 1 of the bit field
        size = dt.itemsize
        exp_bits, mant_bits, bias = {4: (8, 23, 127), 8: (11, 52, 1023)}[size]
        head = struct.pack('<BBBBI', 0x11, 0x20, size * 8 - 1, 0, size)
        props = struct.pack('<HHBBBBI', 0, size * 8, mant_bits, exp_bits, 0, mant_bits, bias)
        return head + props
    if dt.kind in 'iu':
        head = struct.pack('<BBBBI', 0x10, 0x08 if dt.kind == 'i' else 0, 0, 0, dt.itemsize)
        return head + struct.pack('<HH', 0, dt.itemsize * 8)
    if dt.kind == 'S':
        return struct.pack('<BBBBI', 0x13, 0x00, 0, 0, dt.itemsize)      # null-terminated ASCII
    raise H5Error('cannot write dtype %s' % dt)


def _msg(mtype, body):
    body = body + b'\x00' * (-len(body) % 8)
    return struct.pack('<HHBBBB', mtype, len(body), 0, 0, 0, 0) + body


def _header(msgs):
    body = b''.join(msgs)
    return struct.pack('<BBHII', 1, 0, len(msgs), 1, len(body)) + b'\x00' * 4 + body


class _Writer:
    LEAF_K = 512          # up to 1024 links per group in a single symbol node

    def __init__(self, gzip=None):
        self.buf = bytearray()
        self.gzip = gzip or (lambda name, arr: False)

    def alloc(self, data):
        self.buf += b'\x00' * (-len(self.buf) % 8)
        a = len(self.buf)
        self.buf += data
        return a

    def dataset(self, arr, gzip=False):
        arr = np.asarray(arr)
        if arr.ndim:
            arr = np.ascontiguousarray(arr)          # (ascontiguousarray would turn a 0-d scalar into shape (1,))
        if arr.dtype.kind == 'U':
            arr = np.char.encode(arr, 'ascii')
        if arr.dtype.byteorder == '>':
            arr = arr.astype(arr.dtype.newbyteorder('<'))
        if gzip and arr.ndim >= 1 and arr.size:
            return self.dataset_gzip(arr)
        raw = arr.tobytes()
        addr = self.alloc(raw) if raw else UNDEF
        space = struct.pack('<BBBB4x', 1, arr.ndim, 0, 0) + b''.join(struct.pack('<Q', s) for s in arr.shape)
        layout = struct.pack('<BBQQ', 3, 1, addr, len(raw))
        fill = struct.pack('<BBBB', 2, 2, 2, 0)          # fill value v2: late allocation, never written, undefined
        return self.alloc(_header([_msg(0x0001, space), _msg(0x0003, _dtype_msg(arr.dtype)), _msg(0x0005, fill),
                                   _msg(0x0008, layout)]))

    def dataset_gzip(self, arr, level=4):
        """Chunked + deflate, the layout h5py's ``compression='gzip'`` produces: chunks split the first
        axis (at most 64 of them, one leaf of the version-1 chunk B-tree)."""
        nd = arr.ndim
        rows = max(1, -(-arr.shape[0] // 64))
        rows = max(rows, min(arr.shape[0], 3))          # a ragged last chunk exercises edge clipping
        cdims = (rows,) + arr.shape[1:]
        keys = []
        for r0 in range(0, arr.shape[0], rows):
            chunk = np.zeros(cdims, dtype=arr.dtype)
            part = arr[r0:r0 + rows]
            chunk[:part.shape[0]] = part
            data = zlib.compress(chunk.tobytes(), level)
            keys.append(((r0,) + (0,) * (nd - 1), len(data), self.alloc(data)))
        node = b'TREE' + struct.pack('<BBHQQ', 1, 0, len(keys), UNDEF, UNDEF)
        for offs, size, addr in keys:
            node += struct.pack('<II', size, 0) + b''.join(struct.pack('<Q', o) for o in offs + (0,))
            node += struct.pack('<Q', addr)
        node += struct.pack('<II', 0, 0) + b''.join(struct.pack('<Q', o) for o in (arr.shape[0],) + (0,) * nd)
        node += b'\x00' * ((2 * 32 - len(keys)) * (8 + 8 * (nd + 1) + 8))
        bt = self.alloc(node)
        space = struct.pack('<BBBB4x', 1, nd, 0, 0) + b''.join(struct.pack('<Q', s) for s in arr.shape)
        layout = struct.pack('<BBBQ', 3, 2, nd + 1, bt) + b''.join(struct.pack('<I', c) for c in cdims)
        layout += struct.pack('<I', arr.dtype.itemsize)
        pipeline = struct.pack('<BB6x', 1, 1) + struct.pack('<HHHH', 1, 0, 1, 1) + struct.pack('<II', level, 0)
        fill = struct.pack('<BBBB', 2, 3, 2, 0)
        return self.alloc(_header([_msg(0x0001, space), _msg(0x0003, _dtype_msg(arr.dtype)), _msg(0x0005, fill),
                                   _msg(0x000B, pipeline), _msg(0x0008, layout)]))

    def group(self, tree):
        """tree: {name: ndarray | dict}; returns (object header address, btree, heap)."""
        names = sorted(tree)
        if len(names) > 2 * self.LEAF_K:
            raise H5Error('too many links in one group for this writer')
        heap_data = bytearray(b'\x00' * 8)
        offs = {}
        for n in names:
            offs[n] = len(heap_data)
            heap_data += n.encode('utf-8') + b'\x00'
            heap_data += b'\x00' * (-len(heap_data) % 8)
        entries = b''
        for n in names:
            v = tree[n]
            if isinstance(v, dict):
                hdr, bt, hp = self.group(v)
                entries += struct.pack('<QQII', offs[n], hdr, 1, 0) + struct.pack('<QQ', bt, hp)
            else:
                entries += struct.pack('<QQII', offs[n], self.dataset(v, self.gzip(n, v)), 0, 0) + b'\x00' * 16
        entries += b'\x00' * (40 * (2 * self.LEAF_K - len(names)))
        snod = self.alloc(b'SNOD' + struct.pack('<BBH', 1, 0, len(names)) + entries)
        data_addr = self.alloc(bytes(heap_data))
        heap = self.alloc(b'HEAP' + struct.pack('<B3xQQQ', 0, len(heap_data), UNDEF, data_addr))
        last = offs[names[-1]] if names else 0
        # one leaf: key0 = 0 (empty string), child, key1 = heap offset of the largest name; room for 2K entries
        node = b'TREE' + struct.pack('<BBHQQ', 0, 0, 1, UNDEF, UNDEF) + struct.pack('<QQQ', 0, snod, last)
        node += b'\x00' * (16 * (2 * 16) + 8 - 24)
        bt = self.alloc(node)
        hdr = self.alloc(_header([_msg(0x0011, struct.pack('<QQ', bt, heap))]))
        return hdr, bt, heap


def write(path, datasets, gzip=()):
    """Write ``{"group/name": array}`` as an HDF5 file (superblock 0).  Datasets are contiguous; those
    whose last path component starts with one of the ``gzip`` prefixes (or all, ``gzip=True``) are stored
    chunked + deflate like h5py's ``compression='gzip'``."""
    tree = {}
    for key, arr in datasets.items():
        node = tree
        parts = key.strip('/').split('/')
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = np.asarray(arr)
    if gzip is True:
        rule = lambda name, arr: True
    else:
        rule = lambda name, arr: any(name.startswith(g) for g in gzip)
    w = _Writer(rule)
    w.buf += b'\x00' * 96                      # superblock (56 bytes + 40-byte root entry), filled in last
    hdr, bt, heap = w.group(tree)
    eof = len(w.buf)
    sb = SIG + struct.pack('<BBBBBBBBHHI', 0, 0, 0, 0, 0, 8, 8, 0, _Writer.LEAF_K, 16, 0)
    sb += struct.pack('<QQQQ', 0, UNDEF, eof, UNDEF)
    sb += struct.pack('<QQII', 0, hdr, 1, 0) + struct.pack('<QQ', bt, heap)
    assert len(sb) == 96
    w.buf[:96] = sb
    with open(path, 'wb') as f:
        f.write(bytes(w.buf))
    return path

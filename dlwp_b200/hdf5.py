"""
A small HDF5 reader and writer in pure Python / numpy -- enough of the file format for the reference's model files.

The reference stores networks with `keras.Model.save('<name>.keras')` (DLWP/util.py:139-141), i.e. h5py / libhdf5; neither
is installable here, so this module restates the parts of the public "HDF5 File Format Specification" (version 3.0) that
such files use, and writes files a stock libhdf5 can open:

reader  superblock v0/v1 (+ user block) and v2/v3; object headers v1 and v2 (continuation blocks); old-style groups
        (symbol table message -> v1 B-tree + local heap + SNOD nodes) and compact new-style groups (link messages);
        dataspace v1/v2; datatypes fixed-point, IEEE float, fixed-length string, variable-length string (global heap);
        data layout v1-v3 compact / contiguous / chunked (v1 chunk B-tree, deflate / shuffle / fletcher32 filters);
        attribute messages v1-v3.
writer  superblock v0, object headers v1, old-style groups, contiguous datasets, attributes with fixed-length strings /
        numeric scalars and arrays -- what h5py writes with its default `libver='earliest'`.

Pinned against a file produced by libhdf5 itself: scipy ships a MATLAB v7.3 file
(scipy/io/matlab/tests/data/testhdf5_7.4_GLNX86.mat: 512-byte user block, superblock v0, B-tree group, float64 dataset,
string attribute); tests/test_hdf5_cpu.py reads it and checks the known contents.  Host-side interchange only -- nothing
here is on the GPU hot path.
"""

import struct
import zlib

import numpy as np

SIGNATURE = b'\x89HDF\r\n\x1a\n'
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(ValueError):
    pass


# ====================================================================================================================
# Reader
# ====================================================================================================================

class _Obj(object):
    def __init__(self, f, addr, name):
        self._f, self._addr, self.name = f, addr, name
        self._msgs = f._object_header(addr)
        self._attrs = None

    @property
    def attrs(self):
        if self._attrs is None:
            self._attrs = {}
            for t, _, data in self._msgs:
                if t == 0x000C:
                    k, v = self._f._attribute(data)
                    self._attrs[k] = v
        return self._attrs


class Dataset(_Obj):
    def __init__(self, f, addr, name):
        super(Dataset, self).__init__(f, addr, name)
        m = {t: d for t, _, d in self._msgs}
        self.shape = f._dataspace(m[0x0001])
        self.dtype, self._vlen = f._datatype(m[0x0003])
        self._layout = m[0x0008]
        self._filters = f._filters(m[0x000B]) if 0x000B in m else []

    def __getitem__(self, key):
        return self.read()[key]

    def read(self):
        f, buf = self._f, self._f.buf
        if self.shape is None:
            return np.zeros((0,), self.dtype)
        n = int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1
        item = self.dtype.itemsize
        d = self._layout
        ver = d[0]
        if ver == 3:
            cls = d[1]
            if cls == 0:
                size, = struct.unpack_from('<H', d, 2)
                raw = bytes(d[4:4 + size])
            elif cls == 1:
                addr, = f._unpack_o(d, 2)
                size, = f._unpack_l(d, 2 + f.O)
                raw = b'\x00' * (n * item) if addr == f.undef else bytes(buf[f.base + addr:f.base + addr + n * item])
            elif cls == 2:
                nd = d[2]
                bt, = f._unpack_o(d, 3)
                cdims = struct.unpack_from('<%dI' % nd, d, 3 + f.O)
                return self._read_chunked(bt, cdims)
            else:
                raise H5Error('data layout class %d is not supported' % cls)
        elif ver in (1, 2):
            nd, cls = d[1], d[2]
            p = 8
            addr = None
            if cls != 0:
                addr, = f._unpack_o(d, p)
                p += f.O
            dims = struct.unpack_from('<%dI' % nd, d, p)
            p += 4 * nd
            if cls == 0:
                size, = struct.unpack_from('<I', d, p)
                raw = bytes(d[p + 4:p + 4 + size])
            elif cls == 1:
                raw = bytes(buf[f.base + addr:f.base + addr + n * item])
            else:
                return self._read_chunked(addr, dims)
        else:
            raise H5Error('data layout message version %d is not supported (libver=latest files)' % ver)
        if self._vlen:
            return f._vlen_strings(raw, n).reshape(self.shape)
        return np.frombuffer(raw, self.dtype, n).reshape(self.shape).copy()

    def _read_chunked(self, btree, cdims):
        f = self._f
        if self._vlen:
            raise H5Error('chunked variable-length datasets are not supported')
        nd = len(cdims) - 1
        cshape = tuple(int(c) for c in cdims[:nd])
        out = np.zeros(self.shape, self.dtype)
        if btree == f.undef:
            return out
        for offs, fmask, addr, size in f._chunk_leaves(btree, nd):
            raw = bytes(f.buf[f.base + addr:f.base + addr + size])
            for i, (fid, cd) in reversed(list(enumerate(self._filters))):
                if fmask & (1 << i):
                    continue
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:
                    es = cd[0] if cd else self.dtype.itemsize
                    a = np.frombuffer(raw, np.uint8)
                    m = len(raw) // es
                    raw = a[:m * es].reshape(es, m).T.tobytes() + a[m * es:].tobytes()
                elif fid == 3:
                    raw = raw[:-4]
                else:
                    raise H5Error('filter %d is not supported' % fid)
            chunk = np.frombuffer(raw, self.dtype, int(np.prod(cshape))).reshape(cshape)
            sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cshape, self.shape))
            out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]
        return out


class Group(_Obj):
    def __init__(self, f, addr, name):
        super(Group, self).__init__(f, addr, name)
        self._links = None

    def _load(self):
        if self._links is not None:
            return
        f = self._f
        links = {}
        for t, _, d in self._msgs:
            if t == 0x0011:
                bt, = f._unpack_o(d, 0)
                hp, = f._unpack_o(d, f.O)
                links.update(f._symbol_table(bt, hp))
            elif t == 0x0006:
                k, a = f._link(d)
                if k is not None:
                    links[k] = a
            elif t == 0x0002:
                # link info: a defined fractal-heap address means dense link storage
                flags = d[1]
                p = 2 + (8 if flags & 1 else 0)
                fh, = f._unpack_o(d, p)
                if fh != f.undef:
                    raise H5Error('dense link storage (libver=latest groups with many members) is not supported')
        self._links = links

    def keys(self):
        self._load()
        return list(self._links)

    def __contains__(self, k):
        try:
            self[k]
            return True
        except KeyError:
            return False

    def __iter__(self):
        return iter(self.keys())

    def __getitem__(self, path):
        obj = self
        for part in [p for p in path.split('/') if p]:
            if not isinstance(obj, Group):
                raise KeyError(path)
            obj._load()
            if part not in obj._links:
                raise KeyError('%s (no member %r in %s)' % (path, part, obj.name))
            obj = obj._f._open(obj._links[part], (obj.name.rstrip('/') + '/' + part))
        return obj


class File(Group):
    """Read-only view of an HDF5 file: `f['a/b']`, `.keys()`, `.attrs`, `Dataset.read()`."""

    def __init__(self, path_or_bytes):
        if isinstance(path_or_bytes, (bytes, bytearray, memoryview)):
            self.buf = bytes(path_or_bytes)
        else:
            with open(path_or_bytes, 'rb') as fh:
                self.buf = fh.read()
        self._cache = {}
        self._gheaps = {}
        start = 0
        while True:
            if self.buf[start:start + 8] == SIGNATURE:
                break
            start = 512 if start == 0 else start * 2
            if start + 8 > len(self.buf):
                raise H5Error('not an HDF5 file (no signature at 0, 512, 1024, ...)')
        b = self.buf
        ver = b[start + 8]
        if ver in (0, 1):
            self.O, self.L = b[start + 13], b[start + 14]
            p = start + 24 + (4 if ver == 1 else 0)
            self._fmt()
            self.base, = self._unpack_o(b, p)
            p += 4 * self.O
            root_addr, = self._unpack_o(b, p + self.O)
            self.base = start if self.base == 0 and start else self.base   # libhdf5 stores the user block size here
        elif ver in (2, 3):
            self.O, self.L = b[start + 9], b[start + 10]
            self._fmt()
            self.base, = self._unpack_o(b, start + 12)
            root_addr, = self._unpack_o(b, start + 12 + 3 * self.O)
            self.base = start if self.base == 0 and start else self.base
        else:
            raise H5Error('superblock version %d is not supported' % ver)
        super(File, self).__init__(self, root_addr, '/')

    def _fmt(self):
        codes = {2: 'H', 4: 'I', 8: 'Q'}
        self._ofmt, self._lfmt = '<' + codes[self.O], '<' + codes[self.L]
        self.undef = (1 << (8 * self.O)) - 1

    def _unpack_o(self, buf, off):
        return struct.unpack_from(self._ofmt, buf, off)

    def _unpack_l(self, buf, off):
        return struct.unpack_from(self._lfmt, buf, off)

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    # -- object headers ------------------------------------------------------------------------------------------
    def _open(self, addr, name):
        if addr not in self._cache:
            msgs = self._object_header(addr)
            types = {t for t, _, _ in msgs}
            cls = Dataset if 0x0008 in types else Group
            self._cache[addr] = cls(self, addr, name)
        return self._cache[addr]

    def _object_header(self, addr):
        b, p = self.buf, self.base + addr
        if b[p:p + 4] == b'OHDR':
            return self._object_header_v2(p)
        ver, _, nmsg, _, size = struct.unpack_from('<BBHII', b, p)
        if ver != 1:
            raise H5Error('object header version %d at %#x' % (ver, addr))
        msgs, blocks, count = [], [(p + 16, size)], 0
        while blocks and count < nmsg:
            q, n = blocks.pop(0)
            end = q + n
            while q + 8 <= end and count < nmsg:
                t, sz, fl = struct.unpack_from('<HHB', b, q)
                data = b[q + 8:q + 8 + sz]
                q += 8 + sz
                count += 1
                if t == 0x0010:
                    off, = self._unpack_o(data, 0)
                    ln, = self._unpack_l(data, self.O)
                    blocks.append((self.base + off, ln))
                elif t != 0:
                    if fl & 0x02:
                        raise H5Error('shared object header messages are not supported')
                    msgs.append((t, fl, data))
        return msgs

    def _object_header_v2(self, p):
        b = self.buf
        flags = b[p + 5]
        q = p + 6
        if flags & 0x20:
            q += 16
        if flags & 0x10:
            q += 4
        w = 1 << (flags & 3)
        size = int.from_bytes(b[q:q + w], 'little')
        q += w
        msgs, blocks = [], [(q, size)]
        while blocks:
            q, n = blocks.pop(0)
            end = q + n
            while q + 4 <= end:
                t = b[q]
                sz, = struct.unpack_from('<H', b, q + 1)
                fl = b[q + 3]
                q += 4 + (2 if flags & 0x04 else 0)
                data = b[q:q + sz]
                q += sz
                if t == 0x10:
                    off, = self._unpack_o(data, 0)
                    ln, = self._unpack_l(data, self.O)
                    blocks.append((self.base + off + 4, ln - 8))   # skip 'OCHK', drop the checksum
                elif t != 0:
                    if fl & 0x02:
                        raise H5Error('shared object header messages are not supported')
                    msgs.append((t, fl, data))
        return msgs

    # -- groups --------------------------------------------------------------------------------------------------
    def _heap_string(self, heap_addr, off):
        b, p = self.buf, self.base + heap_addr
        if b[p:p + 4] != b'HEAP':
            raise H5Error('bad local heap signature')
        seg, = self._unpack_o(b, p + 8 + 2 * self.L)
        s = self.base + seg + off
        e = b.index(b'\x00', s)
        return b[s:e].decode('utf-8')

    def _symbol_table(self, btree, heap):
        b = self.buf
        out = {}
        p = self.base + btree
        if b[p:p + 4] != b'TREE':
            raise H5Error('bad B-tree signature')
        ntype, level, used = struct.unpack_from('<BBH', b, p + 4)
        q = p + 8 + 2 * self.O
        for i in range(used):
            child, = self._unpack_o(b, q + self.L + i * (self.L + self.O))
            if level > 0:
                out.update(self._symbol_table(child, heap))
                continue
            s = self.base + child
            if b[s:s + 4] != b'SNOD':
                raise H5Error('bad symbol table node signature')
            nsym, = struct.unpack_from('<H', b, s + 6)
            e = s + 8
            for _ in range(nsym):
                noff, = self._unpack_o(b, e)
                oaddr, = self._unpack_o(b, e + self.O)
                out[self._heap_string(heap, noff)] = oaddr
                e += 2 * self.O + 24
        return out

    def _link(self, d):
        flags = d[1]
        p = 2
        ltype = 0
        if flags & 0x08:
            ltype = d[p]
            p += 1
        if flags & 0x04:
            p += 8
        if flags & 0x10:
            p += 1
        w = 1 << (flags & 3)
        n = int.from_bytes(d[p:p + w], 'little')
        p += w
        name = bytes(d[p:p + n]).decode('utf-8')
        p += n
        if ltype != 0:
            return None, None          # soft / external links are not followed
        addr, = self._unpack_o(d, p)
        return name, addr

    # -- messages ------------------------------------------------------------------------------------------------
    def _dataspace(self, d):
        ver, rank, flags = d[0], d[1], d[2]
        if ver == 1:
            p = 8
        elif ver == 2:
            if d[3] == 2:
                return None            # null dataspace
            p = 4
        else:
            raise H5Error('dataspace version %d' % ver)
        return tuple(int(self._unpack_l(d, p + i * self.L)[0]) for i in range(rank))

    def _datatype(self, d):
        cv, b0, b1, b2, size = struct.unpack_from('<BBBBI', d, 0)
        cls = cv & 0x0F
        order = '>' if b0 & 1 else '<'
        if cls == 0:
            return np.dtype('%s%s%d' % (order, 'i' if b0 & 0x08 else 'u', size)), False
        if cls == 1:
            return np.dtype('%sf%d' % (order, size)), False
        if cls == 3:
            return np.dtype('S%d' % size), False
        if cls == 9 and (b0 & 0x0F) == 1:
            return np.dtype(object), True       # variable-length string
        raise H5Error('datatype class %d is not supported' % cls)

    def _filters(self, d):
        ver, n = d[0], d[1]
        p = 8 if ver == 1 else 2
        out = []
        for _ in range(n):
            fid, = struct.unpack_from('<H', d, p)
            p += 2
            nlen = 0
            if ver == 1 or fid >= 256:
                nlen, = struct.unpack_from('<H', d, p)
                p += 2
            _, ncd = struct.unpack_from('<HH', d, p)
            p += 4
            if ver == 1:
                nlen = (nlen + 7) // 8 * 8
            p += nlen
            cd = struct.unpack_from('<%dI' % ncd, d, p)
            p += 4 * ncd
            if ver == 1 and ncd % 2:
                p += 4
            out.append((fid, cd))
        return out

    def _chunk_leaves(self, btree, nd):
        b, p = self.buf, self.base + btree
        if b[p:p + 4] != b'TREE':
            raise H5Error('bad chunk B-tree signature')
        _, level, used = struct.unpack_from('<BBH', b, p + 4)
        q = p + 8 + 2 * self.O
        ksize = 8 + 8 * (nd + 1)
        for i in range(used):
            k = q + i * (ksize + self.O)
            size, fmask = struct.unpack_from('<II', b, k)
            offs = struct.unpack_from('<%dQ' % nd, b, k + 8)
            child, = self._unpack_o(b, k + ksize)
            if level > 0:
                for leaf in self._chunk_leaves(child, nd):
                    yield leaf
            else:
                yield offs, fmask, child, size

    def _global_heap_object(self, addr, index):
        if addr not in self._gheaps:
            b, p = self.buf, self.base + addr
            if b[p:p + 4] != b'GCOL':
                raise H5Error('bad global heap signature')
            total, = self._unpack_l(b, p + 8)
            objs, q, end = {}, p + 8 + self.L, p + total
            while q + 8 + self.L <= end:
                idx, = struct.unpack_from('<H', b, q)
                size, = self._unpack_l(b, q + 8)
                if idx == 0:
                    break
                objs[idx] = b[q + 8 + self.L:q + 8 + self.L + size]
                q += 8 + self.L + (size + 7) // 8 * 8
            self._gheaps[addr] = objs
        return self._gheaps[addr][index]

    def _vlen_strings(self, raw, n):
        out = np.empty((n,), object)
        step = 4 + self.O + 4
        for i in range(n):
            ln, = struct.unpack_from('<I', raw, i * step)
            addr, = self._unpack_o(raw, i * step + 4)
            idx, = struct.unpack_from('<I', raw, i * step + 4 + self.O)
            out[i] = b'' if (addr == 0 or addr == self.undef) else bytes(self._global_heap_object(addr, idx)[:ln])
        return out

    def _attribute(self, d):
        ver = d[0]
        nsz, tsz, ssz = struct.unpack_from('<HHH', d, 2)
        p = 8 + (1 if ver == 3 else 0)
        pad = (lambda v: (v + 7) // 8 * 8) if ver == 1 else (lambda v: v)
        name = bytes(d[p:p + nsz]).split(b'\x00')[0].decode('utf-8')
        p += pad(nsz)
        dtype, vlen = self._datatype(d[p:p + tsz])
        p += pad(tsz)
        shape = self._dataspace(d[p:p + ssz])
        p += pad(ssz)
        if shape is None:
            return name, None
        n = int(np.prod(shape, dtype=np.int64)) if shape else 1
        if vlen:
            v = self._vlen_strings(bytes(d[p:]), n).reshape(shape)
        else:
            v = np.frombuffer(bytes(d[p:p + n * dtype.itemsize]), dtype, n).reshape(shape).copy()
        return name, (v[()] if shape == () else v)


# ====================================================================================================================
# Writer
# ====================================================================================================================

class WGroup(object):
    """Group under construction: `g.attrs[...] = value`, `g.create_group(name)`, `g.create_dataset(name, data)`."""

    def __init__(self):
        self.attrs = {}
        self.members = {}

    def create_group(self, name):
        node = self
        for part in [p for p in name.split('/') if p]:
            nxt = node.members.get(part)
            if nxt is None:
                nxt = node.members[part] = WGroup()
            elif not isinstance(nxt, WGroup):
                raise H5Error('%r is a dataset' % part)
            node = nxt
        return node

    def create_dataset(self, name, data):
        parts = [p for p in name.split('/') if p]
        node = self.create_group('/'.join(parts[:-1])) if len(parts) > 1 else self
        ds = node.members[parts[-1]] = WDataset(np.asarray(data))
        return ds


class WDataset(object):
    def __init__(self, data):
        self.data = data
        self.attrs = {}


_LEAF_K, _INTERNAL_K = 16, 16


def _dtype_message(dt):
    dt = np.dtype(dt)
    if dt.kind == 'f' and dt.itemsize in (4, 8):
        prec = 8 * dt.itemsize
        props = (0, prec, 23, 8, 0, 23, 127) if dt.itemsize == 4 else (0, prec, 52, 11, 0, 52, 1023)
        return struct.pack('<BBBBI', 0x11, 0x20, prec - 1, 0, dt.itemsize) + struct.pack('<HHBBBBI', *props)
    if dt.kind in 'iu':
        return struct.pack('<BBBBI', 0x10, 0x08 if dt.kind == 'i' else 0, 0, 0, dt.itemsize) + \
            struct.pack('<HH', 0, 8 * dt.itemsize)
    if dt.kind == 'S':
        return struct.pack('<BBBBI', 0x13, 0x01, 0, 0, max(1, dt.itemsize))      # null-padded ASCII
    raise H5Error('cannot store dtype %s' % dt)


def _dataspace_message(shape):
    rank = len(shape)
    return struct.pack('<BBBBI', 1, rank, 0, 0, 0) + b''.join(struct.pack('<Q', int(s)) for s in shape)


def _pad8(b):
    return b + b'\x00' * (-len(b) % 8)


def _as_array(v):
    if isinstance(v, str):
        v = v.encode('utf-8')
    if isinstance(v, (list, tuple)) and v and isinstance(v[0], (str, bytes)):
        v = [s.encode('utf-8') if isinstance(s, str) else s for s in v]
    a = np.asarray(v)
    if a.dtype.kind == 'U':
        a = np.char.encode(a, 'utf-8')
    if a.dtype.kind == 'S' and a.dtype.itemsize == 0:
        a = a.astype('S1')
    if a.dtype.kind == 'f' and a.dtype.itemsize == 2:
        a = a.astype(np.float32)
    if a.dtype.kind == 'b':
        a = a.astype(np.int8)
    if a.dtype.byteorder == '>':
        a = a.astype(a.dtype.newbyteorder('<'))
    return a


def _attr_message(name, value):
    a = _as_array(value)
    nm = name.encode('utf-8') + b'\x00'
    dt, ds = _dtype_message(a.dtype), _dataspace_message(a.shape)
    body = struct.pack('<BBHHH', 1, 0, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) + a.tobytes()
    return 0x000C, body


class _Image(object):
    def __init__(self):
        self.buf = bytearray(96)           # superblock v0 with 8-byte offsets / lengths

    def alloc(self, data):
        self.buf += b'\x00' * (-len(self.buf) % 8)
        off = len(self.buf)
        self.buf += data
        return off

    def object_header(self, msgs):
        body = b''
        for t, data in msgs:
            data = _pad8(data)
            if len(data) > 0xFFF8:
                raise H5Error('object header message of %d bytes exceeds the 64 KiB limit of the v1 format' % len(data))
            body += struct.pack('<HHBBBB', t, len(data), 0, 0, 0, 0) + data
        return self.alloc(struct.pack('<BBHII', 1, 0, len(msgs), 1, len(body)) + b'\x00' * 4 + body)

    def dataset(self, ds):
        a = _as_array(ds.data)
        addr = self.alloc(a.tobytes()) if a.size else UNDEF
        msgs = [(0x0001, _dataspace_message(a.shape)), (0x0003, _dtype_message(a.dtype)),
                (0x0005, struct.pack('<BBBB', 2, 2, 2, 0)),                      # fill value v2: late alloc, none defined
                (0x0008, struct.pack('<BBQQ', 3, 1, addr, a.nbytes))]            # layout v3, contiguous
        msgs += [_attr_message(k, v) for k, v in ds.attrs.items()]
        return self.object_header(msgs)

    def group(self, g):
        """Returns (object header address, B-tree address, local heap address)."""
        entries = []
        for name in sorted(g.members, key=lambda s: s.encode('utf-8')):
            m = g.members[name]
            if isinstance(m, WGroup):
                oh, bt, hp = self.group(m)
                entries.append((name, oh, 1, struct.pack('<QQ', bt, hp)))
            else:
                entries.append((name, self.dataset(m), 0, b'\x00' * 16))
        # local heap: the empty string at offset 0, then the member names
        seg = bytearray(b'\x00' * 8)
        offs = []
        for name, _, _, _ in entries:
            offs.append(len(seg))
            seg += _pad8(name.encode('utf-8') + b'\x00')
        free_off = len(seg)
        seg += struct.pack('<QQ', 1, 16)   # one free block at the end: next = 1 (none), size 16
        seg_addr = self.alloc(bytes(seg))
        heap = self.alloc(b'HEAP' + struct.pack('<BBBB', 0, 0, 0, 0) + struct.pack('<QQQ', len(seg), free_off, seg_addr))
        per = 2 * _LEAF_K
        chunks = [list(range(i, min(i + per, len(entries)))) for i in range(0, len(entries), per)]
        if len(chunks) > 2 * _INTERNAL_K:
            raise H5Error('more than %d members in one group' % (per * 2 * _INTERNAL_K))
        keys, children = [0], []
        for idx in chunks:
            body = b'SNOD' + struct.pack('<BBH', 1, 0, len(idx))
            for i in idx:
                name, oh, ctype, scratch = entries[i]
                body += struct.pack('<QQII', offs[i], oh, ctype, 0) + scratch
            body += b'\x00' * (40 * (per - len(idx)))
            children.append(self.alloc(body))
            keys.append(offs[idx[-1]])
        node = b'TREE' + struct.pack('<BBH', 0, 0, len(children)) + struct.pack('<QQ', UNDEF, UNDEF)
        for i in range(2 * _INTERNAL_K):
            node += struct.pack('<Q', keys[i] if i < len(keys) else 0)
            node += struct.pack('<Q', children[i] if i < len(children) else 0)
        node += struct.pack('<Q', keys[-1] if len(keys) > 2 * _INTERNAL_K else 0)
        btree = self.alloc(node)
        msgs = [(0x0011, struct.pack('<QQ', btree, heap))] + [_attr_message(k, v) for k, v in g.attrs.items()]
        return self.object_header(msgs), btree, heap


class FileWriter(WGroup):
    """`w = FileWriter(); w.attrs[...] = ...; w.create_dataset('a/b', arr); w.save(path)`."""

    def tobytes(self):
        img = _Image()
        oh, bt, hp = img.group(self)
        eof = len(img.buf)
        sb = SIGNATURE + struct.pack('<BBBBBBBB', 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack('<HHI', _LEAF_K, _INTERNAL_K, 0)
        sb += struct.pack('<QQQQ', 0, UNDEF, eof, UNDEF)
        sb += struct.pack('<QQII', 0, oh, 1, 0) + struct.pack('<QQ', bt, hp)
        assert len(sb) == 96
        img.buf[:96] = sb
        return bytes(img.buf)

    def save(self, path):
        with open(path, 'wb') as f:
            f.write(self.tobytes())

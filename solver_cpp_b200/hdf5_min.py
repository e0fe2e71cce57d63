"""Minimal pure-Python reader of classic-format HDF5 files (superblock 0, symbol-table groups, version-1 object headers,
contiguous / chunked layout version 3 with deflate, shuffle and fletcher32 filters, fixed-point and IEEE float datasets) --
what DAMASK's DADF5 result files are made of (/root/reference/src/HDF5_utilities.f90, src/result.f90; layout as read by
/root/reference/python/damask/_result.py).  There is no h5py in this environment.  Host-side plumbing:
used to turn a real DAMASK run shipped with the reference into golden vectors (tests/golden/make_golden_g4.py).
Not a general HDF5 implementation: anything outside the list above raises NotImplementedError."""
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


def fletcher32(data):
    """checksum of HDF5's fletcher32 filter (H5_checksum_fletcher32 of libhdf5: big-endian 16-bit words, the two running
    sums folded to 16 bits after every 360 words, an odd trailing byte taken as the high byte of a last word)"""
    n = len(data) // 2
    w16 = np.frombuffer(data, '>u2', n)
    nfull = n // 360
    # per-block sums of all full blocks at once, slab by slab (no int64 copy of the whole data set); the folding recurrence
    # over the blocks stays sequential, on two numbers per block
    S1, S2 = np.empty(nfull, np.int64), np.empty(nfull, np.int64)
    wts = np.arange(360, 0, -1, dtype=np.int64)
    SLAB = 1 << 14
    for b0 in range(0, nfull, SLAB):
        blk = w16[b0 * 360:min(b0 + SLAB, nfull) * 360].reshape(-1, 360).astype(np.int64)
        S1[b0:b0 + len(blk)] = blk.sum(1)
        S2[b0:b0 + len(blk)] = blk @ wts
    s1 = s2 = 0
    for a1, a2 in zip(S1.tolist(), S2.tolist()):
        s2 += 360 * s1 + a2
        s1 += a1
        s1 = (s1 & 0xFFFF) + (s1 >> 16)
        s2 = (s2 & 0xFFFF) + (s2 >> 16)
    if n % 360:
        blk = w16[nfull * 360:].astype(np.int64)
        m = blk.size
        s2 += m * s1 + int((blk * np.arange(m, 0, -1)).sum())
        s1 += int(blk.sum())
        s1 = (s1 & 0xFFFF) + (s1 >> 16)
        s2 = (s2 & 0xFFFF) + (s2 >> 16)
    if len(data) % 2:
        s1 += data[-1] << 8
        s2 += s1
        s1 = (s1 & 0xFFFF) + (s1 >> 16)
        s2 = (s2 & 0xFFFF) + (s2 >> 16)
    s1 = (s1 & 0xFFFF) + (s1 >> 16)
    s2 = (s2 & 0xFFFF) + (s2 >> 16)
    return ((s2 << 16) | s1) & 0xFFFFFFFF


class H5File:
    verify_checksums = True             # fletcher32-filtered chunks are checked against their stored checksum

    def __init__(self, path):
        self.b = open(path, 'rb').read()
        b = self.b
        if b[:8] != b'\x89HDF\r\n\x1a\n':
            raise ValueError('not an HDF5 file')
        if b[8] != 0:
            raise NotImplementedError('superblock version %d' % b[8])
        if b[13] != 8 or b[14] != 8:
            raise NotImplementedError('offset/length size other than 8')
        base = struct.unpack_from('<Q', b, 24)[0]
        if base != 0:
            raise NotImplementedError('non-zero base address')
        # root group symbol table entry after base, free-space, EOF and driver addresses
        self.root = self._symbol_entry(24 + 32)

    # ---- low level
    def _symbol_entry(self, off):
        name_off, ohdr, cache = struct.unpack_from('<QQI', self.b, off)
        e = {'name_off': name_off, 'ohdr': ohdr, 'cache': cache}
        if cache == 1:
            e['btree'], e['heap'] = struct.unpack_from('<QQ', self.b, off + 24)
        return e

    def _messages(self, addr):
        """(type, flags, payload bytes) of a version-1 object header, following continuation blocks"""
        b = self.b
        if b[addr] != 1:
            raise NotImplementedError('object header version %d' % b[addr])
        n_msg = struct.unpack_from('<H', b, addr + 2)[0]
        size = struct.unpack_from('<I', b, addr + 8)[0]
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < n_msg:
            pos, length = blocks.pop(0)
            end = pos + length
            while pos + 8 <= end and len(out) < n_msg:
                mtype, msize, flags = struct.unpack_from('<HHB', b, pos)
                data = b[pos + 8:pos + 8 + msize]
                pos += 8 + msize
                if mtype == 0x0010:
                    o, l = struct.unpack_from('<QQ', data, 0)
                    blocks.append((o, l))
                out.append((mtype, flags, data))
        return out

    def _heap_data(self, heap_addr):
        if self.b[heap_addr:heap_addr + 4] != b'HEAP':
            raise ValueError('local heap signature')
        return struct.unpack_from('<Q', self.b, heap_addr + 24)[0]

    def _group_entries(self, btree, heap):
        """name -> object header address of a symbol-table group"""
        b = self.b
        data = self._heap_data(heap)
        out = {}

        def walk(node):
            if b[node:node + 4] != b'TREE':
                raise ValueError('B-tree signature')
            ntype, level, used = struct.unpack_from('<BBH', b, node + 4)
            if ntype != 0:
                raise ValueError('group B-tree expected')
            pos = node + 24
            for k in range(used):
                child = struct.unpack_from('<Q', b, pos + 8)[0]
                pos += 16
                if level > 0:
                    walk(child)
                else:
                    if b[child:child + 4] != b'SNOD':
                        raise ValueError('symbol node signature')
                    n = struct.unpack_from('<H', b, child + 6)[0]
                    for s in range(n):
                        e = self._symbol_entry(child + 8 + 40 * s)
                        p = data + e['name_off']
                        name = b[p:b.index(b'\0', p)].decode()
                        out[name] = e['ohdr']
        walk(btree)
        return out

    def _children(self, ohdr):
        for mtype, _, data in self._messages(ohdr):
            if mtype == 0x0011:
                btree, heap = struct.unpack_from('<QQ', data, 0)
                return self._group_entries(btree, heap)
        return None

    # ---- public
    def keys(self, path='/'):
        c = self._children(self._resolve(path))
        if c is None:
            raise KeyError('%s is not a group' % path)
        return sorted(c)

    def is_group(self, path):
        return self._children(self._resolve(path)) is not None

    def _resolve(self, path):
        addr = self.root['ohdr']
        for part in [p for p in path.split('/') if p]:
            c = self._children(addr)
            if c is None or part not in c:
                raise KeyError(path)
            addr = c[part]
        return addr

    def read(self, path):
        """dataset -> numpy array (fixed-point and floating-point element types)"""
        shape, dtype, layout, filters = None, None, None, []
        for mtype, _, d in self._messages(self._resolve(path)):
            if mtype == 0x0001:
                ver, rank, flags = d[0], d[1], d[2]
                off = 8 if ver == 1 else 4
                shape = tuple(struct.unpack_from('<%dQ' % rank, d, off)) if rank else ()
            elif mtype == 0x0003:
                dtype = self._dtype(d)
            elif mtype == 0x0008:
                layout = d
            elif mtype == 0x000B:
                filters = self._filters(d)
        if shape is None or dtype is None or layout is None:
            raise ValueError('%s is not a dataset' % path)
        n = int(np.prod(shape)) if shape else 1
        if layout[0] != 3:
            raise NotImplementedError('data layout version %d' % layout[0])
        cls = layout[1]
        if cls == 0:
            size = struct.unpack_from('<H', layout, 2)[0]
            raw = layout[4:4 + size]
        elif cls == 1:
            addr, size = struct.unpack_from('<QQ', layout, 2)
            raw = b'' if addr == UNDEF else self.b[addr:addr + size]
        elif cls == 2:
            nd = layout[2]
            btree = struct.unpack_from('<Q', layout, 3)[0]
            cdims = struct.unpack_from('<%dI' % nd, layout, 11)
            return self._read_chunked(btree, shape, cdims[:-1], dtype, filters)
        else:
            raise NotImplementedError('layout class %d' % cls)
        if len(raw) < n * dtype.itemsize:
            return np.zeros(shape, dtype)
        return np.frombuffer(raw, dtype, n).reshape(shape).copy()

    @staticmethod
    def _dtype(d):
        return H5File._dtype_len(d)[0]

    @staticmethod
    def _dtype_len(d):
        """(numpy dtype, encoded length) of a datatype message; variable-length strings come back as the 16-byte
        global-heap reference records (resolved by attrs())"""
        cls, ver = d[0] & 0x0F, d[0] >> 4
        size = struct.unpack_from('<I', d, 4)[0]
        if cls in (0, 1) and d[1] & 1:
            raise NotImplementedError('big-endian data')
        if cls == 0:
            return np.dtype('<%s%d' % ('i' if d[1] & 0x08 else 'u', size)), 12
        if cls == 1:
            return np.dtype('<f%d' % size), 20
        if cls == 3:
            return np.dtype('S%d' % size), 8
        if cls == 9:
            _, n = H5File._dtype_len(d[8:])
            return np.dtype([('len', '<u4'), ('heap', '<u8'), ('index', '<u4')]), 8 + n
        if cls == 6 and ver in (1, 2, 3):
            n_mem = d[1] | (d[2] << 8)
            pos = 8
            names, formats, offsets = [], [], []
            for _ in range(n_mem):
                end = d.index(b'\0', pos)
                names.append(d[pos:end].decode())
                pos = (pos + (end - pos) // 8 * 8 + 8) if ver < 3 else end + 1
                if ver == 3:
                    nb = max(1, (int(size).bit_length() + 7) // 8)
                    offsets.append(int.from_bytes(d[pos:pos + nb], 'little'))
                    pos += nb
                else:
                    offsets.append(struct.unpack_from('<I', d, pos)[0])
                    pos += 4
                    if ver == 1:
                        if d[pos]:
                            raise NotImplementedError('array members of a version-1 compound type')
                        pos += 28
                dt, n = H5File._dtype_len(d[pos:])
                formats.append(dt)
                pos += n
            return np.dtype({'names': names, 'formats': formats, 'offsets': offsets, 'itemsize': size}), pos
        return np.dtype('V%d' % size), len(d)        # anything else: raw records

    def _global_heap_object(self, addr, index):
        b = self.b
        if b[addr:addr + 4] != b'GCOL':
            raise ValueError('global heap signature')
        size = struct.unpack_from('<Q', b, addr + 8)[0]
        pos, end = addr + 16, addr + size
        while pos + 16 <= end:
            idx, _, _, n = struct.unpack_from('<HHIQ', b, pos)
            if idx == index:
                return b[pos + 16:pos + 16 + n]
            if idx == 0:
                break
            pos += 16 + (n + 7) // 8 * 8
        raise KeyError('global heap object %d' % index)

    def attrs(self, path='/'):
        """attributes of a group or dataset: name -> numpy array / scalar / str (version 1-3 attribute messages; fixed-
        and variable-length strings decoded)"""
        out = {}
        for mtype, _, d in self._messages(self._resolve(path)):
            if mtype != 0x000C:
                continue
            ver = d[0]
            if ver not in (1, 2, 3):
                raise NotImplementedError('attribute message version %d' % ver)
            n_name, n_type, n_space = struct.unpack_from('<HHH', d, 2)
            pos = 8 if ver < 3 else 9
            rnd = (lambda n: (n + 7) // 8 * 8) if ver == 1 else (lambda n: n)
            name = d[pos:pos + n_name].split(b'\0')[0].decode()
            pos += rnd(n_name)
            t = d[pos:pos + n_type]
            dtype = self._dtype(t)
            vlen = (t[0] & 0x0F) == 9
            pos += rnd(n_type)
            sp = d[pos:pos + n_space]
            rank = sp[1]
            shape = tuple(struct.unpack_from('<%dQ' % rank, sp, 8 if sp[0] == 1 else 4)) if rank else ()
            pos += rnd(n_space)
            n = int(np.prod(shape)) if shape else 1
            a = np.frombuffer(d[pos:pos + n * dtype.itemsize], dtype, n)
            if vlen:
                a = np.array([self._global_heap_object(int(r['heap']), int(r['index']))[:int(r['len'])].decode()
                              for r in a], dtype=object)
            elif dtype.kind == 'S':
                a = np.array([x.split(b'\0')[0].decode() for x in a], dtype=object)
            a = a.reshape(shape)
            out[name] = a[()] if shape == () else a.copy()
        return out

    @staticmethod
    def _filters(d):
        ver, nf = d[0], d[1]
        pos = 8 if ver == 1 else 2
        out = []
        for _ in range(nf):
            fid = struct.unpack_from('<H', d, pos)[0]
            if ver == 1 or fid >= 256:
                nlen = struct.unpack_from('<H', d, pos + 2)[0]
                flags, ncl = struct.unpack_from('<HH', d, pos + 4)
                pos += 8
                if ver == 1:
                    nlen = (nlen + 7) // 8 * 8
                pos += nlen
            else:
                flags, ncl = struct.unpack_from('<HH', d, pos + 2)
                pos += 6
            cd = struct.unpack_from('<%dI' % ncl, d, pos)
            pos += 4 * ncl
            if ver == 1 and ncl % 2:
                pos += 4
            out.append((fid, cd))
        return out

    def _read_chunked(self, btree, shape, cdims, dtype, filters):
        b = self.b
        out = np.zeros(shape, dtype)
        nd = len(shape)
        if btree == UNDEF:
            return out

        def walk(node):
            if b[node:node + 4] != b'TREE':
                raise ValueError('B-tree signature')
            ntype, level, used = struct.unpack_from('<BBH', b, node + 4)
            if ntype != 1:
                raise ValueError('chunk B-tree expected')
            ksize = 8 + 8 * (nd + 1)
            pos = node + 24
            for _ in range(used):
                csize, mask = struct.unpack_from('<II', b, pos)
                offs = struct.unpack_from('<%dQ' % nd, b, pos + 8)
                child = struct.unpack_from('<Q', b, pos + ksize)[0]
                pos += ksize + 8
                if level > 0:
                    walk(child)
                    continue
                raw = b[child:child + csize]
                for k in range(len(filters) - 1, -1, -1):
                    if mask & (1 << k):
                        continue
                    fid, cd = filters[k]
                    if fid == 3:
                        raw, stored = raw[:-4], struct.unpack('<I', raw[-4:])[0]
                        if self.verify_checksums and fletcher32(raw) != stored:
                            raise ValueError('fletcher32 checksum of a chunk at %d does not match' % child)
                    elif fid == 1:
                        raw = zlib.decompress(raw)
                    elif fid == 2:
                        es = cd[0] if cd else dtype.itemsize
                        a = np.frombuffer(raw, np.uint8)
                        m = a.size // es
                        raw = a[:m * es].reshape(es, m).T.tobytes() + a[m * es:].tobytes()
                    else:
                        raise NotImplementedError('HDF5 filter %d' % fid)
                chunk = np.frombuffer(raw, dtype, int(np.prod(cdims))).reshape(cdims)
                sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, shape))
                out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]
        walk(btree)
        return out


# ------------------------------------------------------------------------------------------------------------ writer
def _pad8(b):
    return b + b'\0' * (-len(b) % 8)


def _encode_dtype(dt):
    """datatype message of a numpy dtype (version 1 encodings, little endian)"""
    dt = np.dtype(dt)
    if dt.kind in 'iu':
        return struct.pack('<BBBBIHH', 0x10, 0x08 if dt.kind == 'i' else 0, 0, 0, dt.itemsize, 0, 8 * dt.itemsize)
    if dt.kind == 'f' and dt.itemsize in (4, 8):
        e, m, bias = ((8, 23, 127), (11, 52, 1023))[dt.itemsize == 8]
        return struct.pack('<BBBBIHHBBBBI', 0x11, 0x20, 8 * dt.itemsize - 1, 0, dt.itemsize, 0, 8 * dt.itemsize, m, e, 0, m, bias)
    if dt.kind == 'S':
        return struct.pack('<BBBBI', 0x13, 0x01, 0, 0, dt.itemsize)                   # null-padded, ASCII
    if dt.names:
        out = struct.pack('<BBBBI', 0x16, len(dt.names) & 0xFF, len(dt.names) >> 8, 0, dt.itemsize)
        for name in dt.names:
            sub, off = dt.fields[name][:2]
            if sub.shape:
                raise NotImplementedError('array members of a compound type')
            out += _pad8(name.encode() + b'\0') + struct.pack('<IB3xI4x16x', off, 0, 0) + _encode_dtype(sub)
        return out
    raise NotImplementedError('HDF5 encoding of dtype %s' % dt)


def _encode_space(shape):
    return struct.pack('<BBB5x', 1, len(shape), 0) + b''.join(struct.pack('<Q', n) for n in shape)


def _as_h5_array(v):
    """numpy array of a type the writer can encode; str and lists of str become fixed-length byte strings"""
    if isinstance(v, str):
        v = v.encode()
    if isinstance(v, bytes):
        return np.array(v, dtype='S%d' % max(1, len(v)))
    a = np.asarray(v)
    if a.dtype.kind in 'UO':
        enc = [str(x).encode() for x in a.reshape(-1)]
        a = np.array(enc, dtype='S%d' % max(1, max(map(len, enc), default=1))).reshape(a.shape)
    if a.dtype.kind == 'b':
        a = a.astype(np.int8)
    if a.dtype.byteorder == '>':
        a = a.astype(a.dtype.newbyteorder('<'))
    return a


class _Node:
    def __init__(self, data=None):
        self.data = data                # None: group; dataset: (shape, dtype, n_bytes, file address of the raw data)
        self.attrs = {}
        self.children = {}


class H5Writer:
    """Writer of the same classic HDF5 subset H5File reads: superblock 0, symbol-table groups (group B-tree of any depth, leaf
    K 4 / internal K 16 like libhdf5's defaults), version-1 object headers, little-endian datasets of integer / float /
    fixed-length string / compound types stored contiguously or (checksum=True) as one fletcher32-filtered chunk the way
    DAMASK's result datasets are, version-1 attributes.  The raw data of a dataset goes to the file (path
    given to the constructor) the moment dataset() is called -- a result file of a 256^3 run is larger than one wants to
    hold in host memory -- the small metadata (groups, headers, attributes) is kept and written by save() / close();
    without a path everything goes to memory and tobytes() returns the file image.
    Written for solver_cpp_b200.dadf5 (DADF5 result files, /root/reference/src/result.f90, src/HDF5_utilities.f90).
    No libhdf5 / h5py exists in this environment: files are checked by reading them back with H5File, which itself is
    checked against result files written by real DAMASK (tests/test_hdf5_cpu.py)."""
    LEAF_K, NODE_K = 4, 16

    def __init__(self, path=None):
        import io
        self.root = _Node()
        self.path = path
        self.fh = open(path, 'wb') if path is not None else io.BytesIO()
        self.fh.write(b'\0' * 96)                      # superblock, filled in at the end
        self.pos = 96
        self.closed = False

    def _walk(self, path, create):
        node = self.root
        for part in [p for p in path.split('/') if p]:
            if node.data is not None:
                raise KeyError('%s: a dataset has no children' % path)
            if part not in node.children:
                if not create:
                    raise KeyError(path)
                node.children[part] = _Node()
            node = node.children[part]
        return node

    def group(self, path, **attrs):
        g = self._walk(path, True)
        if g.data is not None:
            raise KeyError('%s is a dataset' % path)
        g.attrs.update(attrs)
        return path

    CHUNK_K = 32                        # libhdf5's default for chunk B-trees (not stored in a version-0 superblock)

    def dataset(self, path, array, checksum=False, **attrs):
        """checksum=True: stored the way DAMASK's result datasets are -- chunked layout, one chunk spanning the dataset,
        fletcher32 filter (the checksum behind the data) -- instead of contiguous"""
        parent, _, name = path.rstrip('/').rpartition('/')
        g = self._walk(parent, True)
        if g.data is not None or name in g.children:
            raise KeyError('%s exists or lies below a dataset' % path)
        a = np.ascontiguousarray(_as_h5_array(array))
        _encode_dtype(a.dtype)                                              # refuse unsupported element types here
        if self.closed:
            raise ValueError('H5Writer: file already written')
        raw = a.tobytes()
        # HDF5 caps a chunk at 4 GiB (32-bit chunk size in the B-tree key): larger datasets (F at 512^3 is 9.7 GB) are stored
        # contiguous, without the checksum, instead of failing at close() after all raw data has been streamed
        chunked = bool(checksum) and 0 < len(raw) < (1 << 32) - 8 and a.ndim > 0
        if chunked:
            raw += struct.pack('<I', fletcher32(raw))
        n = _Node((a.shape, a.dtype, len(raw), self.pos if raw else UNDEF, chunked))
        self.fh.write(_pad8(raw))
        self.pos += len(raw) + (-len(raw) % 8)
        n.attrs.update(attrs)
        g.children[name] = n

    def set_attrs(self, path, **attrs):
        self._walk(path, False).attrs.update(attrs)

    # ---- serialisation
    def _alloc(self, blob):
        addr = self.pos + len(self.buf)
        self.buf += _pad8(blob)
        return addr

    def _header(self, messages):
        body = b''
        for mtype, flags, data in messages:
            data = _pad8(data)
            if len(data) > 0xFFFF:
                raise ValueError('object header message of %d bytes (attribute too large)' % len(data))
            body += struct.pack('<HHB3x', mtype, len(data), flags) + data
        return self._alloc(struct.pack('<BBHII4x', 1, 0, len(messages), 1, len(body)) + body)

    @staticmethod
    def _attr_messages(attrs):
        out = []
        for name, v in attrs.items():
            a = _as_h5_array(v)
            nm, t, sp = name.encode() + b'\0', _encode_dtype(a.dtype), _encode_space(a.shape)
            out.append((0x000C, 0, struct.pack('<BBHHH', 1, 0, len(nm), len(t), len(sp)) + _pad8(nm) + _pad8(t) + _pad8(sp) +
                        np.ascontiguousarray(a).tobytes()))
        return out

    def _write_node(self, node):
        """-> (object header address, B-tree address, heap address); the latter two 0 for a dataset"""
        if node.data is not None:
            shape, dtype, n_bytes, addr, chunked = node.data
            msgs = [(0x0001, 0, _encode_space(shape)), (0x0003, 1, _encode_dtype(dtype))]
            if chunked:
                # chunk B-tree (node type 1) with one entry: key (chunk bytes, filter mask, offsets + 0), child, end key
                nd = len(shape)
                key = 8 + 8 * (nd + 1)
                blob = b'TREE' + struct.pack('<BBHQQ', 1, 0, 1, UNDEF, UNDEF)
                blob += struct.pack('<II', n_bytes, 0) + struct.pack('<%dQ' % (nd + 1), *([0] * (nd + 1))) + struct.pack('<Q', addr)
                blob += struct.pack('<II', 0, 0) + struct.pack('<%dQ' % (nd + 1), *(list(shape) + [dtype.itemsize]))
                blob += b'\0' * (24 + 2 * self.CHUNK_K * 8 + (2 * self.CHUNK_K + 1) * key - len(blob))
                btree = self._alloc(blob)
                msgs += [(0x0005, 1, struct.pack('<BBBBI', 2, 3, 2, 1, 0)),
                         (0x000B, 1, struct.pack('<BB6xHHHH', 1, 1, 3, 16, 0, 0) + b'fletcher32'.ljust(16, b'\0')),
                         (0x0008, 0, struct.pack('<BBBQ', 3, 2, nd + 1, btree) +
                          struct.pack('<%dI' % (nd + 1), *(list(shape) + [dtype.itemsize])))]
            else:
                msgs += [(0x0005, 1, struct.pack('<BBBBI', 2, 1, 2, 1, 0)),
                         (0x0008, 0, struct.pack('<BBQQ', 3, 1, addr, n_bytes))]
            return self._header(msgs + self._attr_messages(node.attrs)), 0, 0
        # children first, then the local heap of their names, the symbol nodes and the B-tree over them
        names = sorted(node.children, key=lambda s: s.encode())
        done = {n: self._write_node(node.children[n]) for n in names}
        heap = bytearray(8)
        off = {}
        for n in names:
            off[n] = len(heap)
            heap += _pad8(n.encode() + b'\0')
        free = len(heap)
        heap += struct.pack('<QQ', 1, 16)                                   # one free block, end of list
        data_addr = self._alloc(bytes(heap))
        heap_addr = self._alloc(b'HEAP' + struct.pack('<B3xQQQ', 0, len(heap), free, data_addr))
        level = []                                                          # (address, heap offset of the largest name below)
        for s in range(0, len(names), 2 * self.LEAF_K):
            part = names[s:s + 2 * self.LEAF_K]
            blob = b'SNOD' + struct.pack('<BBH', 1, 0, len(part))
            for n in part:
                ohdr, bt, hp = done[n]
                is_group = node.children[n].data is None
                blob += struct.pack('<QQI4xQQ', off[n], ohdr, 1 if is_group else 0, bt, hp)
            blob += b'\0' * (8 + 2 * self.LEAF_K * 40 - len(blob))
            level.append((self._alloc(blob), off[part[-1]]))
        depth = 0
        node_size = 24 + (4 * self.NODE_K + 1) * 8
        while True:
            groups = [level[s:s + 2 * self.NODE_K] for s in range(0, len(level), 2 * self.NODE_K)] or [[]]
            base = self.pos + len(self.buf)
            addrs = [base + k * node_size for k in range(len(groups))]
            nxt = []
            for k, grp in enumerate(groups):
                first_key = 0 if k == 0 else groups[k - 1][-1][1]
                blob = b'TREE' + struct.pack('<BBHQQ', 0, depth, len(grp), addrs[k - 1] if k else UNDEF,
                                             addrs[k + 1] if k + 1 < len(groups) else UNDEF) + struct.pack('<Q', first_key)
                for child, key in grp:
                    blob += struct.pack('<QQ', child, key)
                blob += b'\0' * (node_size - len(blob))
                assert self._alloc(blob) == addrs[k]
                nxt.append((addrs[k], grp[-1][1] if grp else 0))
            if len(nxt) == 1:
                btree = nxt[0][0]
                break
            level, depth = nxt, depth + 1
        ohdr = self._header([(0x0011, 0, struct.pack('<QQ', btree, heap_addr))] + self._attr_messages(node.attrs))
        return ohdr, btree, heap_addr

    def close(self):
        """writes the metadata behind the raw data and the superblock in front; the writer is finished afterwards"""
        if self.closed:
            return
        self.buf = bytearray()
        ohdr, btree, heap = self._write_node(self.root)
        self.fh.write(bytes(self.buf))
        eof = self.pos + len(self.buf)
        del self.buf
        sb = b'\x89HDF\r\n\x1a\n' + struct.pack('<BBBBBBBBHHI', 0, 0, 0, 0, 0, 8, 8, 0, self.LEAF_K, self.NODE_K, 0)
        sb += struct.pack('<QQQQ', 0, UNDEF, eof, UNDEF) + struct.pack('<QQI4xQQ', 0, ohdr, 1, btree, heap)
        self.fh.seek(0)
        self.fh.write(sb)
        self.closed = True
        if self.path is not None:
            self.fh.close()

    def tobytes(self):
        if self.path is not None:
            raise ValueError('H5Writer: streaming to %s; tobytes() is for the in-memory mode' % self.path)
        self.close()
        return self.fh.getvalue()

    def save(self, path=None):
        """in-memory mode: write the image to `path`; streaming mode: finish the file given to the constructor"""
        if self.path is not None:
            if path is not None and path != self.path:
                raise ValueError('H5Writer: already streaming to %s' % self.path)
            self.close()
            return
        with open(path, 'wb') as f:
            f.write(self.tobytes())


def check_group_structure(f, path='/'):
    """Strict walk of a symbol-table group the way libhdf5 searches it (by the B-tree KEYS, which H5File ignores): names in
    a symbol node sorted and within (left key, right key], sibling pointers consistent, heap free list inside the data
    segment.  Recurses into sub-groups; returns the number of links visited.  Raises ValueError on any violation."""
    b = f.b
    addr = f._resolve(path)
    sym = [d for t, _, d in f._messages(addr) if t == 0x0011]
    if not sym:
        return 0
    btree, heap = struct.unpack_from('<QQ', sym[0], 0)
    if b[heap:heap + 4] != b'HEAP':
        raise ValueError('%s: heap signature' % path)
    hsize, hfree, hdata = struct.unpack_from('<QQQ', b, heap + 8)
    pos = hfree
    while pos != 1:
        if pos + 16 > hsize:
            raise ValueError('%s: heap free list leaves the data segment' % path)
        pos, blk = struct.unpack_from('<QQ', b, hdata + pos)
        if blk < 16:
            raise ValueError('%s: heap free block smaller than its header' % path)
    name_at = lambda off: b[hdata + off:b.index(b'\0', hdata + off)]
    leaf_k, node_k = struct.unpack_from('<HH', b, 16)
    count = [0]

    def walk(node, lo, hi, first):
        if b[node:node + 4] != b'TREE' or b[node + 4] != 0:
            raise ValueError('%s: group B-tree node expected' % path)
        level, used = b[node + 5], struct.unpack_from('<H', b, node + 6)[0]
        if used > 2 * node_k:
            raise ValueError('%s: B-tree node over-full' % path)
        keys = [struct.unpack_from('<Q', b, node + 24 + 16 * k)[0] for k in range(used + 1)]
        kids = [struct.unpack_from('<Q', b, node + 32 + 16 * k)[0] for k in range(used)]
        if used and (lo is not None and name_at(keys[0]) != lo or hi is not None and name_at(keys[-1]) != hi):
            raise ValueError('%s: B-tree node keys disagree with the parent' % path)
        for k, child in enumerate(kids):
            left, right = name_at(keys[k]), name_at(keys[k + 1])
            if level > 0:
                walk(child, left, right, first and k == 0)
                continue
            if b[child:child + 4] != b'SNOD':
                raise ValueError('%s: symbol node signature' % path)
            n = struct.unpack_from('<H', b, child + 6)[0]
            if not 0 < n <= 2 * leaf_k:
                raise ValueError('%s: symbol node with %d entries' % (path, n))
            names = [name_at(struct.unpack_from('<Q', b, child + 8 + 40 * s)[0]) for s in range(n)]
            if names != sorted(names) or names[-1] != right or not (names[0] > left or (first and k == 0 and left == b'')):
                raise ValueError('%s: symbol node names %r outside (%r, %r]' % (path, names, left, right))
            count[0] += n
    walk(btree, None, None, True)
    for k in f.keys(path):
        q = path.rstrip('/') + '/' + k
        if f.is_group(q):
            count[0] += check_group_structure(f, q)
    return count[0]

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


class H5File:
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
        cls, ver = d[0] & 0x0F, d[0] >> 4
        size = struct.unpack_from('<I', d, 4)[0]
        if d[1] & 1:
            raise NotImplementedError('big-endian data')
        if cls == 0:
            return np.dtype('<%s%d' % ('i' if d[1] & 0x08 else 'u', size))
        if cls == 1:
            return np.dtype('<f%d' % size)
        if cls == 3:
            return np.dtype('S%d' % size)
        return np.dtype('V%d' % size)               # compound etc.: raw records

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
                        raw = raw[:-4]
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

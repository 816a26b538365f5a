"""Minimal pure-Python HDF5 reader - just enough for the reference's ``testdata/*.h5``.

h5py and the HDF5 C library are not installed in the build container, so the golden
fixtures are extracted with this reader.  Supported: superblock v0/v1, v1 object headers
(with continuation blocks), old-style groups (symbol-table B-tree + local heap), fixed
point / IEEE float little-endian datatypes, contiguous and chunked (v1 B-tree) layouts,
deflate and shuffle filters.  Only ``tests/golden/make_golden.py`` uses it.
"""
import struct
import zlib

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"


class H5File:
    def __init__(self, path):
        with open(path, "rb") as fh:
            self.buf = fh.read()
        if self.buf[:8] != _SIG:
            raise ValueError("not an HDF5 file")
        ver = self.buf[8]
        if ver not in (0, 1):
            raise NotImplementedError(f"superblock version {ver}")
        self.O = self.buf[13]
        self.L = self.buf[14]
        pos = 24 + (4 if ver == 1 else 0)
        pos += 4 * self.O                      # base, free-space, EOF, driver-info addresses
        # root group symbol table entry
        _name_off, pos = self._off(pos)
        root_hdr, pos = self._off(pos)
        self.root = self._read_group(root_hdr)

    # -- primitive readers -------------------------------------------------
    def _off(self, pos):
        return int.from_bytes(self.buf[pos:pos + self.O], "little"), pos + self.O

    def _len(self, pos):
        return int.from_bytes(self.buf[pos:pos + self.L], "little"), pos + self.L

    # -- object headers ----------------------------------------------------
    def _messages(self, addr):
        ver, _, nmsg, _refc, hsize = struct.unpack_from("<BBHII", self.buf, addr)
        if ver != 1:
            raise NotImplementedError(f"object header version {ver}")
        blocks = [(addr + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and len(out) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", self.buf, pos)
                data = self.buf[pos + 8:pos + 8 + msize]
                pos += 8 + msize
                if mtype == 0x10:          # continuation
                    coff, p2 = self._off_b(data, 0)
                    clen, _ = self._len_b(data, p2)
                    blocks.append((coff, clen))
                out.append((mtype, data))
        return out

    def _off_b(self, b, pos):
        return int.from_bytes(b[pos:pos + self.O], "little"), pos + self.O

    def _len_b(self, b, pos):
        return int.from_bytes(b[pos:pos + self.L], "little"), pos + self.L

    # -- groups --------------------------------------------------------------
    def _read_group(self, hdr_addr):
        for mtype, data in self._messages(hdr_addr):
            if mtype == 0x11:              # symbol table message
                btree, p = self._off_b(data, 0)
                heap, _ = self._off_b(data, p)
                return self._walk_group_btree(btree, self._heap_data(heap))
        raise ValueError("object is not an old-style group")

    def _heap_data(self, addr):
        assert self.buf[addr:addr + 4] == b"HEAP"
        pos = addr + 8
        size, pos = self._len(pos)
        _free, pos = self._len(pos)
        daddr, pos = self._off(pos)
        return self.buf[daddr:daddr + size]

    def _walk_group_btree(self, addr, heap):
        assert self.buf[addr:addr + 4] == b"TREE"
        _ntype, level, used = struct.unpack_from("<BBH", self.buf, addr + 4)
        pos = addr + 8 + 2 * self.O
        entries = {}
        pos += self.L                           # key 0
        for _ in range(used):
            child, pos = self._off(pos)
            pos += self.L                       # next key
            if level > 0:
                entries.update(self._walk_group_btree(child, heap))
            else:
                entries.update(self._read_snod(child, heap))
        return entries

    def _read_snod(self, addr, heap):
        assert self.buf[addr:addr + 4] == b"SNOD"
        nsym = struct.unpack_from("<H", self.buf, addr + 6)[0]
        pos = addr + 8
        out = {}
        for _ in range(nsym):
            name_off, pos = self._off(pos)
            hdr, pos = self._off(pos)
            pos += 4 + 4 + 16
            name = heap[name_off:heap.index(b"\0", name_off)].decode()
            out[name] = hdr
        return out

    def keys(self):
        return list(self.root)

    # -- datasets ------------------------------------------------------------
    def __getitem__(self, name):
        msgs = self._messages(self.root[name])
        shape = dtype = layout = None
        filters = []
        for mtype, d in msgs:
            if mtype == 0x01:
                ver, rank, flags = d[0], d[1], d[2]
                pos = 8 if ver == 1 else 4
                shape = tuple(int.from_bytes(d[pos + i * self.L:pos + (i + 1) * self.L], "little")
                              for i in range(rank))
            elif mtype == 0x03:
                cls = d[0] & 0x0F
                size = struct.unpack_from("<I", d, 4)[0]
                if d[1] & 1:
                    raise NotImplementedError("big-endian data")
                if cls == 0:
                    signed = bool(d[1] & 0x08)
                    dtype = np.dtype(f"<{'i' if signed else 'u'}{size}")
                elif cls == 1:
                    dtype = np.dtype(f"<f{size}")
                else:
                    raise NotImplementedError(f"datatype class {cls}")
            elif mtype == 0x08:
                if d[0] != 3:
                    raise NotImplementedError(f"layout version {d[0]}")
                if d[1] == 1:
                    a, p = self._off_b(d, 2)
                    n, _ = self._len_b(d, p)
                    layout = ("contiguous", a, n)
                elif d[1] == 2:
                    ndim = d[2]
                    a, p = self._off_b(d, 3)
                    cdims = struct.unpack_from(f"<{ndim}I", d, p)
                    layout = ("chunked", a, cdims[:-1])
                else:
                    raise NotImplementedError("compact layout")
            elif mtype == 0x0B:
                ver, nf = d[0], d[1]
                pos = 8 if ver == 1 else 2
                for _ in range(nf):
                    fid = struct.unpack_from("<H", d, pos)[0]
                    pos += 2
                    nlen = 0
                    if ver == 1 or fid >= 256:
                        nlen = struct.unpack_from("<H", d, pos)[0]
                        pos += 2
                    _fl, ncd = struct.unpack_from("<HH", d, pos)
                    pos += 4
                    pos += (nlen + 7) // 8 * 8 if ver == 1 else nlen
                    pos += 4 * ncd
                    if ver == 1 and ncd % 2:
                        pos += 4
                    filters.append(fid)
        if layout[0] == "contiguous":
            arr = np.frombuffer(self.buf, dtype, count=int(np.prod(shape)), offset=layout[1])
            return arr.reshape(shape).copy()
        out = np.zeros(shape, dtype)
        self._read_chunks(layout[1], layout[2], filters, dtype, out)
        return out

    def _read_chunks(self, addr, cdims, filters, dtype, out):
        assert self.buf[addr:addr + 4] == b"TREE"
        _ntype, level, used = struct.unpack_from("<BBH", self.buf, addr + 4)
        rank = len(cdims)
        pos = addr + 8 + 2 * self.O
        for _ in range(used):
            csize, _mask = struct.unpack_from("<II", self.buf, pos)
            offs = struct.unpack_from(f"<{rank + 1}Q", self.buf, pos + 8)
            pos += 8 + 8 * (rank + 1)
            child, pos = self._off(pos)
            if level > 0:
                self._read_chunks(child, cdims, filters, dtype, out)
                continue
            raw = self.buf[child:child + csize]
            for fid in reversed(filters):
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:
                    n = len(raw) // dtype.itemsize
                    raw = np.frombuffer(raw, np.uint8).reshape(dtype.itemsize, n).T.tobytes()
                elif fid == 3:
                    raw = raw[:-4]
                else:
                    raise NotImplementedError(f"filter {fid}")
            chunk = np.frombuffer(raw, dtype).reshape(cdims)
            sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, out.shape))
            out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]

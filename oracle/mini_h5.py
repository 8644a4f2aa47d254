"""Minimal pure-Python reader for the HDF5 files meshio wrote for the reference's meshes
(``meshes/**/mesh.h5``: superblock v0, v1 object headers, chunked + gzip datasets ``/data0`` points,
``/data1`` cells, ``/data2`` material).  TEST INFRASTRUCTURE ONLY: used by
tests/golden/make_golden.py in the build container (h5py is not installed) to build dolfin-free
surrogate foreground matrices on the real meshes.  Not a general HDF5 implementation."""
from __future__ import annotations

import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5File:
    def __init__(self, path):
        self.d = open(path, "rb").read()
        d = self.d
        assert d[:8] == b"\x89HDF\r\n\x1a\n" and d[8] == 0, "only superblock version 0 is supported"
        assert d[13] == 8 and d[14] == 8, "8-byte offsets/lengths expected"
        # superblock v0: 24 bytes of header fields, then base, free-space, eof, driver addresses, root entry
        root_entry = 24 + 4 * 8
        self.root_header = struct.unpack_from("<Q", d, root_entry + 8)[0]
        cache_type = struct.unpack_from("<I", d, root_entry + 16)[0]
        assert cache_type == 1, "root symbol table entry without cached B-tree/heap addresses"
        self.root_btree, self.root_heap = struct.unpack_from("<QQ", d, root_entry + 24)
        self.links = self._read_group(self.root_btree, self.root_heap)

    # ---- groups (symbol tables)
    def _heap_data(self, addr):
        d = self.d
        assert d[addr:addr + 4] == b"HEAP"
        return struct.unpack_from("<Q", d, addr + 24)[0]

    def _read_group(self, btree, heap):
        d = self.d
        heap_data = self._heap_data(heap)
        out = {}

        def walk(addr):
            assert d[addr:addr + 4] == b"TREE" and d[addr + 4] == 0
            level = d[addr + 5]
            n = struct.unpack_from("<H", d, addr + 6)[0]
            pos = addr + 8 + 16
            children = []
            for k in range(n):
                pos += 8  # key k
                children.append(struct.unpack_from("<Q", d, pos)[0])
                pos += 8
            for c in children:
                if level > 0:
                    walk(c)
                else:
                    assert d[c:c + 4] == b"SNOD"
                    ns = struct.unpack_from("<H", d, c + 6)[0]
                    for e in range(ns):
                        ent = c + 8 + e * 40
                        name_off, hdr = struct.unpack_from("<QQ", d, ent)
                        s = heap_data + name_off
                        name = d[s:d.index(b"\0", s)].decode()
                        out[name] = hdr

        walk(btree)
        return out

    # ---- datasets
    def _messages(self, addr):
        d = self.d
        assert d[addr] == 1, "only version 1 object headers are supported"
        nmsg = struct.unpack_from("<H", d, addr + 2)[0]
        size = struct.unpack_from("<I", d, addr + 8)[0]
        blocks = [(addr + 16, size)]
        msgs = []
        while blocks and len(msgs) < nmsg:
            pos, left = blocks.pop(0)
            end = pos + left
            while pos + 8 <= end and len(msgs) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", d, pos)
                body = pos + 8
                if mtype == 0x10:  # continuation
                    caddr, clen = struct.unpack_from("<QQ", d, body)
                    blocks.append((caddr, clen))
                msgs.append((mtype, body, msize))
                pos = body + msize
        return msgs

    def read(self, name) -> np.ndarray:
        d = self.d
        msgs = self._messages(self.links[name])
        shape = dtype = layout = None
        filters = []
        for mtype, body, msize in msgs:
            if mtype == 1:  # dataspace
                ver, rank, flags = d[body], d[body + 1], d[body + 2]
                off = body + (8 if ver == 1 else 4)
                shape = struct.unpack_from("<" + "Q" * rank, d, off)
            elif mtype == 3:  # datatype
                cls = d[body] & 0x0F
                size = struct.unpack_from("<I", d, body + 4)[0]
                if cls == 0:
                    signed = (d[body + 1] >> 3) & 1
                    dtype = np.dtype(("<i" if signed else "<u") + str(size))
                elif cls == 1:
                    dtype = np.dtype("<f" + str(size))
                else:
                    raise NotImplementedError(f"datatype class {cls}")
            elif mtype == 8:  # layout
                ver = d[body]
                assert ver == 3, "layout message version 3 expected"
                lclass = d[body + 1]
                if lclass == 1:
                    a, s = struct.unpack_from("<QQ", d, body + 2)
                    layout = ("contiguous", a, s)
                elif lclass == 2:
                    rank = d[body + 2]
                    a = struct.unpack_from("<Q", d, body + 3)[0]
                    dims = struct.unpack_from("<" + "I" * rank, d, body + 11)
                    layout = ("chunked", a, dims)
                else:
                    raise NotImplementedError("compact layout")
            elif mtype == 11:  # filter pipeline
                ver, nf = d[body], d[body + 1]
                pos = body + (8 if ver == 1 else 2)
                for _ in range(nf):
                    fid, namelen, _fl, ncd = struct.unpack_from("<HHHH", d, pos)
                    pos += 8
                    if ver == 1 or fid >= 256:
                        pos += (namelen + 7) // 8 * 8 if ver == 1 else namelen
                    cd = struct.unpack_from("<" + "I" * ncd, d, pos)
                    pos += 4 * ncd
                    if ver == 1 and ncd % 2:
                        pos += 4
                    filters.append((fid, cd))
        assert shape is not None and dtype is not None and layout is not None
        n = int(np.prod(shape)) if shape else 1
        if layout[0] == "contiguous":
            return np.frombuffer(d, dtype=dtype, count=n, offset=layout[1]).reshape(shape).copy()
        out = np.zeros(shape, dtype=dtype)
        cdims = layout[2][:-1]  # last entry is the element size
        rank = len(cdims)

        def unfilter(raw, mask):
            for k in range(len(filters) - 1, -1, -1):
                if mask & (1 << k):
                    continue
                fid, cd = filters[k]
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:
                    es = cd[0] if cd else dtype.itemsize
                    a = np.frombuffer(raw, dtype=np.uint8)
                    m = a.size // es
                    raw = a[:m * es].reshape(es, m).T.tobytes() + a[m * es:].tobytes()
                else:
                    raise NotImplementedError(f"filter {fid}")
            return raw

        def walk(addr):
            assert d[addr:addr + 4] == b"TREE" and d[addr + 4] == 1
            level = d[addr + 5]
            nent = struct.unpack_from("<H", d, addr + 6)[0]
            pos = addr + 8 + 16
            keysize = 8 + 8 * (rank + 1)
            for _ in range(nent):
                csize, fmask = struct.unpack_from("<II", d, pos)
                offs = struct.unpack_from("<" + "Q" * (rank + 1), d, pos + 8)
                child = struct.unpack_from("<Q", d, pos + keysize)[0]
                pos += keysize + 8
                if level > 0:
                    walk(child)
                else:
                    raw = unfilter(d[child:child + csize], fmask)
                    chunk = np.frombuffer(raw, dtype=dtype, count=int(np.prod(cdims))).reshape(cdims)
                    sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs[:rank], cdims, shape))
                    out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]

        if layout[1] != UNDEF:
            walk(layout[1])
        return out


def read_mesh(path):
    """(points [nv, dim] f64, cells [nc, dim+1] int, material [nc] f64) of a reference ``mesh.h5``."""
    f = H5File(path)
    return f.read("data0"), f.read("data1").astype(np.int64), f.read("data2")

"""Minimal read-only HDF5 parser for the mesh files DOLFINx/XDMF writes.

No h5py/libhdf5 in this image, and the reference's fixtures
(`cpp/fenicsx-sf/tests/test_operators3d/mesh.h5`) are plain: superblock version 0, old-style groups
(symbol table + v1 B-tree + local heap), version-1 object headers, contiguous uncompressed datasets
of fixed-point / IEEE float types.  Exactly that subset is supported; anything else raises.
"""
import struct

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class Hdf5Error(ValueError):
    pass


class File:
    def __init__(self, path):
        with open(path, "rb") as f:
            self.b = f.read()
        b = self.b
        if b[:8] != _SIG:
            raise Hdf5Error("not an HDF5 file")
        if b[8] != 0:
            raise Hdf5Error(f"superblock version {b[8]} not supported (only 0)")
        self.O, self.L = b[13], b[14]
        if self.O != 8 or self.L != 8:
            raise Hdf5Error("only 8-byte offsets/lengths supported")
        # 8 sig + 8 version bytes + 2+2 K + 4 flags = 24, then base, free, eof, driver (4*O)
        self.base = self._u(24)
        root_entry = 24 + 4 * self.O
        self.root = self._symtab_entry(root_entry)

    # -- primitives ----------------------------------------------------------------------------
    def _u(self, off, n=8):
        return int.from_bytes(self.b[off:off + n], "little")

    def _symtab_entry(self, off):
        name_off = self._u(off)
        hdr = self._u(off + 8)
        cache = self._u(off + 16, 4)
        btree = heap = None
        if cache == 1:
            btree, heap = self._u(off + 24), self._u(off + 32)
        return dict(name_off=name_off, header=hdr, btree=btree, heap=heap)

    def _messages(self, addr):
        """Yield (type, bytes) for every message of a version-1 object header (+ continuations)."""
        b = self.b
        if b[addr] != 1:
            raise Hdf5Error(f"object header version {b[addr]} not supported (only 1)")
        nmsg = self._u(addr + 2, 2)
        size = self._u(addr + 8, 4)
        blocks = [(addr + 16, size)]
        seen = 0
        while blocks and seen < nmsg:
            pos, left = blocks.pop(0)
            end = pos + left
            while pos + 8 <= end and seen < nmsg:
                mtype, msize = self._u(pos, 2), self._u(pos + 2, 2)
                data = b[pos + 8:pos + 8 + msize]
                pos += 8 + msize
                seen += 1
                if mtype == 0x10:                      # continuation
                    blocks.append((self._u_from(data, 0), self._u_from(data, 8)))
                else:
                    yield mtype, data

    @staticmethod
    def _u_from(data, off, n=8):
        return int.from_bytes(data[off:off + n], "little")

    # -- groups --------------------------------------------------------------------------------
    def _group_tables(self, header_addr):
        for mtype, data in self._messages(header_addr):
            if mtype == 0x11:                          # symbol table message
                return self._u_from(data, 0), self._u_from(data, 8)
        raise Hdf5Error("object is not an old-style group")

    def _heap_data(self, heap_addr):
        if self.b[heap_addr:heap_addr + 4] != b"HEAP":
            raise Hdf5Error("bad local heap")
        return self._u(heap_addr + 24)

    def _btree_entries(self, addr, heap_data):
        b = self.b
        if b[addr:addr + 4] != b"TREE":
            raise Hdf5Error("bad B-tree node")
        level, used = b[addr + 5], self._u(addr + 6, 2)
        pos = addr + 8 + 2 * self.O                    # past the sibling pointers
        out = {}
        for i in range(used):
            child = self._u(pos + self.L + i * (self.L + self.O))
            if level > 0:
                out.update(self._btree_entries(child, heap_data))
                continue
            if b[child:child + 4] != b"SNOD":
                raise Hdf5Error("bad symbol table node")
            nsym = self._u(child + 6, 2)
            for k in range(nsym):
                e = self._symtab_entry(child + 8 + 40 * k)
                s = heap_data + e["name_off"]
                name = b[s:b.index(b"\0", s)].decode()
                out[name] = e
        return out

    def listdir(self, path="/"):
        return sorted(self._children(self._resolve(path)))

    def _children(self, entry):
        btree, heap = entry.get("btree"), entry.get("heap")
        if btree is None:
            btree, heap = self._group_tables(entry["header"])
        return self._btree_entries(btree, self._heap_data(heap))

    def _resolve(self, path):
        e = self.root
        for part in [p for p in path.split("/") if p]:
            ch = self._children(e)
            if part not in ch:
                raise KeyError(path)
            e = ch[part]
        return e

    # -- datasets ------------------------------------------------------------------------------
    def read(self, path):
        e = self._resolve(path)
        shape = dtype = addr = nbytes = None
        for mtype, data in self._messages(e["header"]):
            if mtype == 0x01:                          # dataspace
                ver, rank = data[0], data[1]
                off = 8 if ver == 1 else 4
                shape = tuple(self._u_from(data, off + 8 * i) for i in range(rank))
            elif mtype == 0x03:                        # datatype
                cls, size = data[0] & 0x0F, self._u_from(data, 4, 4)
                if cls == 0:
                    signed = bool(data[1] & 0x08)
                    dtype = np.dtype(f"<{'i' if signed else 'u'}{size}")
                elif cls == 1:
                    dtype = np.dtype(f"<f{size}")
                else:
                    raise Hdf5Error(f"datatype class {cls} not supported")
                if data[1] & 0x01:
                    dtype = dtype.newbyteorder(">")
            elif mtype == 0x08:                        # data layout
                ver = data[0]
                if ver == 3:
                    if data[1] != 1:
                        raise Hdf5Error("only contiguous datasets supported")
                    addr, nbytes = self._u_from(data, 2), self._u_from(data, 10)
                elif ver in (1, 2):
                    rank, cls = data[1], data[2]
                    if cls != 1:
                        raise Hdf5Error("only contiguous datasets supported")
                    addr = self._u_from(data, 8)
                else:
                    raise Hdf5Error(f"layout version {ver} not supported")
            elif mtype == 0x0B:
                raise Hdf5Error("filtered (compressed) datasets not supported")
        if shape is None or dtype is None or addr is None or addr == _UNDEF:
            raise Hdf5Error(f"{path}: not a contiguous dataset with allocated storage")
        n = int(np.prod(shape)) if shape else 1
        start = self.base + addr
        arr = np.frombuffer(self.b, dtype=dtype, count=n, offset=start)
        return arr.reshape(shape).copy()

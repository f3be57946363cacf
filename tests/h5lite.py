"""A small, independent HDF5 reader for the tests (pure Python + numpy, no libhdf5).

It follows the HDF5 File Format Specification for the subset libhdf5 writes with default
property lists: superblock 0/1 (optionally behind a user block), version-1 object headers
with continuation blocks, symbol-table groups (version-1 B-tree + local heap + symbol
nodes), dataspace v1/v2, datatypes fixed / float / compound v1-v3 / array v2-v3, layout
v1-v3 contiguous or compact.

It is written separately from host/heroam_h5.c so that the two can check each other, and
it is itself pinned against a file written by the real library
(tests/test_h5.py::test_reader_parses_a_file_written_by_libhdf5).
"""
from __future__ import annotations

import struct

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(Exception):
    pass


class Dataset:
    def __init__(self, shape, dtype, data, messages):
        self.shape = shape
        self.dtype = dtype
        self.data = data
        self.messages = messages  # [(type, bytes)]


class File:
    def __init__(self, path: str):
        with open(path, "rb") as f:
            self.buf = f.read()
        at = 0
        while True:
            if self.buf[at:at + 8] == SIGNATURE:
                break
            at = 512 if at == 0 else at * 2
            if at >= len(self.buf):
                raise H5Error("no HDF5 signature")
        sb = self.buf[at:]
        self.super_version = sb[8]
        if self.super_version > 1:
            raise H5Error(f"superblock version {self.super_version} not supported")
        self.free_space_version, self.root_version, _, self.shared_version = sb[9], sb[10], sb[11], sb[12]
        self.size_of_offsets, self.size_of_lengths = sb[13], sb[14]
        if (self.size_of_offsets, self.size_of_lengths) != (8, 8):
            raise H5Error("only 8-byte offsets / lengths")
        self.leaf_k, self.internal_k = struct.unpack_from("<HH", sb, 16)
        self.consistency_flags = struct.unpack_from("<I", sb, 20)[0]
        pos = 24 if self.super_version == 0 else 28
        self.base_address, self.free_space_address, self.eof_address, self.driver_address = struct.unpack_from("<4Q", sb, pos)
        self.base = at  # where addresses are measured from
        root = sb[pos + 32:pos + 72]
        self.root_name_offset, self.root_header, self.root_cache_type = struct.unpack_from("<QQI", root, 0)
        self.root_scratch = struct.unpack_from("<QQ", root, 24)

    # -- primitives ---------------------------------------------------------------------
    def at(self, addr: int, n: int) -> bytes:
        lo = self.base + addr
        if addr == UNDEF or lo + n > len(self.buf):
            raise H5Error(f"read of {n} bytes at {addr} outside the file")
        return self.buf[lo:lo + n]

    def messages(self, addr: int):
        """[(type, flags, data)] of a version-1 object header, continuation blocks followed."""
        pre = self.at(addr, 16)
        version, _, count, refs, size = struct.unpack_from("<BBHII", pre, 0)
        if version != 1:
            raise H5Error(f"object header version {version}")
        out = []
        blocks = [(addr + 16, size)]
        while blocks and count:
            baddr, bsize = blocks.pop(0)
            blk = self.at(baddr, bsize)
            pos = 0
            while count and pos + 8 <= bsize:
                mtype, msize, flags = struct.unpack_from("<HHB", blk, pos)
                if msize % 8:
                    raise H5Error("message size not a multiple of 8")
                data = blk[pos + 8:pos + 8 + msize]
                if len(data) != msize:
                    raise H5Error("message overruns its block")
                count -= 1
                if mtype == 0x0010:
                    blocks.append(struct.unpack_from("<QQ", data, 0))
                else:
                    out.append((mtype, flags, data))
                pos += 8 + msize
        if count:
            raise H5Error("object header lists more messages than it holds")
        return out

    # -- groups -------------------------------------------------------------------------
    def group_entries(self, header: int):
        """{name: (header address, cache type, scratch)} of a symbol-table group, in B-tree order."""
        stab = [d for t, _, d in self.messages(header) if t == 0x0011]
        if not stab:
            raise H5Error("not a symbol-table group")
        btree, heap = struct.unpack_from("<QQ", stab[0], 0)
        hh = self.at(heap, 32)
        if hh[:4] != b"HEAP" or hh[4] != 0:
            raise H5Error("bad local heap")
        hsize, hfree, hdata = struct.unpack_from("<QQQ", hh, 8)
        if hfree != 1 and hfree + 16 > hsize:
            raise H5Error("local heap free list outside the data segment")
        names = self.at(hdata, hsize)
        out = {}
        self._order = []

        def name_at(off):
            end = names.index(b"\0", off)
            return names[off:end].decode()

        def walk(node, depth, lower_key):
            hd = self.at(node, 24)
            if hd[:4] != b"TREE" or hd[4] != 0:
                raise H5Error("bad group B-tree node")
            level, used = hd[5], struct.unpack_from("<H", hd, 6)[0]
            if used > 2 * self.internal_k:
                raise H5Error("B-tree node over-full")
            self.at(node, 24 + 16 * 2 * self.internal_k + 8)  # a full node must lie inside the file
            body = self.at(node + 24, 16 * used + 8)
            keys = [struct.unpack_from("<Q", body, 16 * i)[0] for i in range(used + 1)]
            kids = [struct.unpack_from("<Q", body, 16 * i + 8)[0] for i in range(used)]
            for i, kid in enumerate(kids):
                lo, hi = name_at(keys[i]), name_at(keys[i + 1])
                if level > 0:
                    walk(kid, depth + 1, keys[i])
                    continue
                sn = self.at(kid, 8)
                if sn[:4] != b"SNOD" or sn[4] != 1:
                    raise H5Error("bad symbol node")
                n = struct.unpack_from("<H", sn, 6)[0]
                if n > 2 * self.leaf_k:
                    raise H5Error("symbol node over-full")
                ents = self.at(kid + 8, 40 * 2 * self.leaf_k)
                for e in range(n):
                    noff, ohdr, cache = struct.unpack_from("<QQI", ents, 40 * e)
                    name = name_at(noff)
                    # B-tree invariant: key[i] < every name in child i <= key[i+1] (strcmp order)
                    if not (lo.encode() < name.encode() <= hi.encode()):
                        raise H5Error(f"name {name!r} outside its key interval ({lo!r}, {hi!r}]")
                    if self._order and not (self._order[-1].encode() < name.encode()):
                        raise H5Error("names not in strcmp order")
                    self._order.append(name)
                    out[name] = (ohdr, cache, struct.unpack_from("<QQ", ents, 40 * e + 24))

        walk(btree, 0, 0)
        return out

    # -- datasets -----------------------------------------------------------------------
    def _dtype(self, d: bytes, pos: int = 0):
        """(numpy dtype, encoded length) of the datatype message at d[pos:]."""
        cls, ver = d[pos] & 0x0F, d[pos] >> 4
        bits = d[pos + 1] | (d[pos + 2] << 8) | (d[pos + 3] << 16)
        size = struct.unpack_from("<I", d, pos + 4)[0]
        order = ">" if bits & 1 else "<"
        if cls == 0:
            offset, precision = struct.unpack_from("<HH", d, pos + 8)
            if offset != 0 or precision != 8 * size:
                raise H5Error("padded integer")
            return np.dtype(f"{order}{'i' if bits & 8 else 'u'}{size}"), 12
        if cls == 1:
            offset, precision, eloc, esize, mloc, msize, bias = struct.unpack_from("<HHBBBBI", d, pos + 8)
            want = {4: (31, 23, 8, 0, 23, 127), 8: (63, 52, 11, 0, 52, 1023)}[size]
            if ((bits >> 8) & 0xFF, eloc, esize, mloc, msize, bias) != want or (bits >> 4) & 3 != 2:
                raise H5Error("not an IEEE float")
            return np.dtype(f"{order}f{size}"), 20
        if cls == 10:
            nd = d[pos + 8]
            if ver == 2:
                dims = struct.unpack_from(f"<{nd}I", d, pos + 12)
                head = 12 + 8 * nd
            elif ver == 3:
                dims = struct.unpack_from(f"<{nd}I", d, pos + 9)
                head = 9 + 4 * nd
            else:
                raise H5Error(f"array datatype version {ver}")
            base, blen = self._dtype(d, pos + head)
            if base.itemsize * int(np.prod(dims)) != size:
                raise H5Error("array size mismatch")
            return np.dtype((base, tuple(dims))), head + blen
        if cls == 6:
            n = bits & 0xFFFF
            p = pos + 8
            names, formats, offsets = [], [], []
            for _ in range(n):
                end = d.index(b"\0", p)
                name = d[p:end].decode()
                ln = end - p + 1
                p += ln if ver >= 3 else (ln + 7) // 8 * 8
                if ver >= 3:
                    width = 1 if size < 256 else 2 if size < 65536 else 4
                    off = int.from_bytes(d[p:p + width], "little")
                    p += width
                else:
                    off = struct.unpack_from("<I", d, p)[0]
                    p += 4
                if ver == 1:
                    p += 28
                mt, mlen = self._dtype(d, p)
                p += mlen
                names.append(name); formats.append(mt); offsets.append(off)
            return np.dtype({"names": names, "formats": formats, "offsets": offsets, "itemsize": size}), p - pos
        raise H5Error(f"datatype class {cls} not supported")

    def dataset(self, header: int) -> Dataset:
        msgs = self.messages(header)
        shape = dtype = None
        data = None
        for t, _, d in msgs:
            if t == 0x0001:
                ver, rank, flags = d[0], d[1], d[2]
                head = 8 if ver == 1 else 4
                shape = struct.unpack_from(f"<{rank}Q", d, head)
            elif t == 0x0003:
                dtype, _ = self._dtype(d)
        for t, _, d in msgs:
            if t != 0x0008:
                continue
            nbytes = int(np.prod(shape)) * dtype.itemsize
            if d[0] == 3:
                if d[1] == 1:
                    addr, size = struct.unpack_from("<QQ", d, 2)
                    if size != nbytes:
                        raise H5Error("layout size differs from dataspace x datatype")
                    data = self.at(addr, nbytes)
                elif d[1] == 0:
                    size = struct.unpack_from("<H", d, 2)[0]
                    data = d[4:4 + size]
                else:
                    raise H5Error("chunked layout")
            elif d[0] in (1, 2):
                rank, lclass = d[1], d[2]
                if lclass != 1:
                    raise H5Error("only contiguous v1/v2 layouts")
                addr = struct.unpack_from("<Q", d, 8)[0]
                data = self.at(addr, nbytes)
        if shape is None or dtype is None or data is None:
            raise H5Error("dataset header incomplete")
        return Dataset(shape, dtype, np.frombuffer(data, dtype=dtype).reshape(shape), msgs)

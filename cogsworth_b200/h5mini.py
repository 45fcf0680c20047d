"""A minimal HDF5 reader / writer for the ``orbits`` group of a cogsworth population file.

``Population.save`` writes four plain arrays with h5py (reference ``src/cogsworth/pop.py:1853-1876``)::

    orbits/offsets  int64 (n + 1,)    orbits/pos  f64 (3, total)    orbits/vel  f64 (3, total)    orbits/t  f64 (total,)

h5py's defaults for such arrays are the oldest, simplest structures of the format (HDF5 File Format Specification
version 1.1): superblock version 0, groups as symbol tables (a version-1 B-tree over symbol-table nodes plus a local
heap of names), version-1 object headers, contiguous dataset storage, little-endian fixed- and floating-point types.
That subset is small enough to read and write with ``struct`` alone, which is what this module does -- so the orbit
wire format works without h5py (absent from the build image) and stays the reference's own: a file written here opens
with h5py / ``Population.orbits``' load branch (pop.py:696-719), and ``read_datasets`` reads the files h5py writes for
these arrays.  Anything outside the subset (chunked / compressed datasets, version-2 object headers of
``libver='latest'``, other datatypes) raises ``NotImplementedError`` naming what was met.
"""
from __future__ import annotations

import struct

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K, INTERNAL_K = 4, 16          # group B-tree fan-outs of a default file
FREE_NULL = 1                       # "no free block" in a local heap header (libhdf5's H5HL_FREE_NULL)

__all__ = ["write_datasets", "read_datasets", "list_datasets"]


def _pad8(n: int) -> int:
    return (n + 7) & ~7


# ------------------------------------------------------------------------------------------------ writer
def _msg(mtype: int, data: bytes, flags: int = 0) -> bytes:
    data = data + b"\0" * (_pad8(len(data)) - len(data))
    return struct.pack("<HHB3x", mtype, len(data), flags) + data


def _object_header(messages) -> bytes:
    body = b"".join(messages)
    # version, reserved, number of messages, reference count, header size; the messages start 8-byte aligned
    return struct.pack("<BxHII4x", 1, len(messages), 1, len(body)) + body


def _dataspace(shape) -> bytes:
    return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", int(d)) for d in shape)


def _datatype(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt == np.dtype("<i8"):
        return struct.pack("<BBBBI", 0x10, 0x08, 0, 0, 8) + struct.pack("<HH", 0, 64)
    if dt == np.dtype("<f8"):
        return struct.pack("<BBBBI", 0x11, 0x20, 0x3F, 0, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
    raise NotImplementedError(f"h5mini writes int64 and float64 datasets only, not {dt}")


def _group_blocks(names, header_addrs, at):
    """(local heap + its data, B-tree node, symbol-table node) of a group holding `names`, laid out from address `at`;
    returns (bytes, btree address, heap address)."""
    order = sorted(range(len(names)), key=lambda i: names[i].encode())
    if len(order) > 2 * LEAF_K:
        raise NotImplementedError("h5mini writes groups of at most 8 members")
    heap_data = bytearray(b"\0" * 8)                      # offset 0: the empty name
    offs = {}
    for i in order:
        offs[i] = len(heap_data)
        nm = names[i].encode() + b"\0"
        heap_data += nm + b"\0" * (_pad8(len(nm)) - len(nm))
    heap_addr = at
    heap_hdr_size = 32
    heap_data_addr = heap_addr + heap_hdr_size
    heap = b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), FREE_NULL, heap_data_addr) + bytes(heap_data)
    btree_addr = heap_addr + len(heap)
    btree_size = 24 + 2 * INTERNAL_K * 8 + (2 * INTERNAL_K + 1) * 8
    snod_addr = btree_addr + btree_size
    last_off = offs[order[-1]] if order else 0
    btree = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if order else 0, UNDEF, UNDEF)
    btree += struct.pack("<QQQ", 0, snod_addr, last_off)
    btree += b"\0" * (btree_size - len(btree))
    snod = b"SNOD" + struct.pack("<BxH", 1, len(order))
    for i in order:
        snod += struct.pack("<QQII16x", offs[i], header_addrs[i], 0, 0)
    snod += b"\0" * (8 + 2 * LEAF_K * 40 - len(snod))
    return heap + btree + snod, btree_addr, heap_addr


def write_datasets(file_name: str, group: str, arrays: dict) -> None:
    """Write a NEW HDF5 file whose root holds one group `group` with the given int64 / float64 arrays as contiguous
    datasets (row-major, little-endian)."""
    names = list(arrays)
    data = [np.ascontiguousarray(arrays[k]) for k in names]
    for k, a in zip(names, data):
        if a.dtype not in (np.dtype("<i8"), np.dtype("<f8")):
            raise NotImplementedError(f"dataset {k!r}: h5mini writes int64 and float64 only, not {a.dtype}")
    # ---- layout: superblock | root header | root group blocks | group header | group blocks | dataset headers | raw data
    pos = 96
    root_hdr_addr = pos
    root_hdr_size = 16 + 8 + 16
    pos += root_hdr_size
    root_blocks_addr = pos
    blocks_size = 32 + _pad8(8 + sum(_pad8(len(n.encode()) + 1) for n in [group])) + 0
    # sizes depend on the names only: build twice (first with dummy addresses) to learn them
    dummy, _, _ = _group_blocks([group], [0], root_blocks_addr)
    pos += len(dummy)
    grp_hdr_addr = pos
    pos += root_hdr_size
    grp_blocks_addr = pos
    dummy2, _, _ = _group_blocks(names, [0] * len(names), grp_blocks_addr)
    pos += len(dummy2)
    ds_hdr_addrs, ds_hdr_sizes = [], []
    for a in data:
        size = 16 + (8 + _pad8(8 + 8 * a.ndim)) + (8 + _pad8(len(_datatype(a.dtype)))) + (8 + _pad8(18))
        ds_hdr_addrs.append(pos)
        ds_hdr_sizes.append(size)
        pos += size
    raw_addrs = []
    for a in data:
        pos = _pad8(pos)
        raw_addrs.append(pos)
        pos += a.nbytes
    eof = pos
    # ---- bytes
    root_blocks, root_btree, root_heap = _group_blocks([group], [grp_hdr_addr], root_blocks_addr)
    grp_blocks, grp_btree, grp_heap = _group_blocks(names, ds_hdr_addrs, grp_blocks_addr)
    sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII", 0, root_hdr_addr, 1, 0) + struct.pack("<QQ", root_btree, root_heap)
    assert len(sb) == 96
    root_hdr = _object_header([_msg(0x0011, struct.pack("<QQ", root_btree, root_heap))])
    grp_hdr = _object_header([_msg(0x0011, struct.pack("<QQ", grp_btree, grp_heap))])
    assert len(root_hdr) == root_hdr_size == len(grp_hdr) and len(root_blocks) == len(dummy) and len(grp_blocks) == len(dummy2)
    with open(file_name, "wb") as f:
        f.write(sb + root_hdr + root_blocks + grp_hdr + grp_blocks)
        for a, addr, size in zip(data, raw_addrs, ds_hdr_sizes):
            hdr = _object_header([_msg(0x0001, _dataspace(a.shape), 0), _msg(0x0003, _datatype(a.dtype), 1),
                                  _msg(0x0008, struct.pack("<BBQQ", 3, 1, addr, a.nbytes))])
            assert len(hdr) == size
            f.write(hdr)
        for a, addr in zip(data, raw_addrs):
            f.write(b"\0" * (addr - f.tell()))
            f.write(a.tobytes())
        assert f.tell() == eof


# ------------------------------------------------------------------------------------------------ reader
class _File:
    def __init__(self, file_name):
        self.f = open(file_name, "rb")
        self.base = None
        for off in (0, 512, 1024, 2048, 4096):        # a user block may precede the superblock
            self.f.seek(off)
            if self.f.read(8) == SIGNATURE:
                self.base = off
                break
        if self.base is None:
            raise ValueError(f"{file_name} is not an HDF5 file")
        ver = self.f.read(1)[0]
        if ver > 1:
            raise NotImplementedError(f"HDF5 superblock version {ver} (written with libver='latest'?) is outside h5mini's subset")
        self.f.seek(self.base + 13)
        so, sl = self.f.read(2)
        if (so, sl) != (8, 8):
            raise NotImplementedError("only 8-byte offsets / lengths")
        self.f.seek(self.base + 8 + 8 + 8 + (4 if ver == 1 else 0))
        base_addr, _, _, _ = struct.unpack("<QQQQ", self.f.read(32))
        _, self.root_header, cache, _ = struct.unpack("<QQII", self.f.read(24))
        self.base = base_addr                         # every address in the file is relative to it (= the user block size)

    def read(self, addr, n):
        self.f.seek(self.base + addr)
        return self.f.read(n)

    def close(self):
        self.f.close()

    # -- object headers ---------------------------------------------------------------------------
    def messages(self, addr):
        head = self.read(addr, 16)
        if head[:4] == b"OHDR":
            raise NotImplementedError("version-2 object header (libver='latest') is outside h5mini's subset")
        ver, nmsg, _, size = struct.unpack("<BxHII", head[:12])
        if ver != 1:
            raise NotImplementedError(f"object header version {ver}")
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            a, sz = blocks.pop(0)
            buf = self.read(a, sz)
            p = 0
            while p + 8 <= len(buf) and len(out) < nmsg:
                mtype, msize, flags = struct.unpack("<HHB", buf[p:p + 5])
                body = buf[p + 8:p + 8 + msize]
                p += 8 + msize
                if mtype == 0x0010:                   # continuation block
                    blocks.append(struct.unpack("<QQ", body[:16]))
                out.append((mtype, body))
        return out

    # -- groups -----------------------------------------------------------------------------------
    def members(self, header_addr):
        st = [b for t, b in self.messages(header_addr) if t == 0x0011]
        if not st:
            if any(t in (0x0002, 0x0006) for t, _ in self.messages(header_addr)):
                raise NotImplementedError("new-style group (link messages) is outside h5mini's subset")
            return None                               # not a group
        btree, heap = struct.unpack("<QQ", st[0][:16])
        hh = self.read(heap, 32)
        assert hh[:4] == b"HEAP"
        dsize, _, daddr = struct.unpack("<QQQ", hh[8:32])
        names = self.read(daddr, dsize)
        out = {}

        def walk(node):
            h = self.read(node, 24)
            assert h[:4] == b"TREE"
            level, used = h[5], struct.unpack("<H", h[6:8])[0]
            body = self.read(node + 24, (2 * used + 1) * 8)
            for k in range(used):
                child = struct.unpack("<Q", body[8 + 16 * k:16 + 16 * k])[0]
                if level > 0:
                    walk(child)
                    continue
                s = self.read(child, 8)
                assert s[:4] == b"SNOD"
                nsym = struct.unpack("<H", s[6:8])[0]
                ents = self.read(child + 8, 40 * nsym)
                for e in range(nsym):
                    noff, haddr = struct.unpack("<QQ", ents[40 * e:40 * e + 16])
                    end = names.index(b"\0", noff)
                    out[names[noff:end].decode()] = haddr
        if btree != UNDEF:
            walk(btree)
        return out

    # -- datasets ---------------------------------------------------------------------------------
    def dataset(self, header_addr):
        shape = dtype = layout = None
        for t, b in self.messages(header_addr):
            if t == 0x0001:
                ver, rank = b[0], b[1]
                start = 8 if ver == 1 else 4
                shape = struct.unpack(f"<{rank}Q", b[start:start + 8 * rank])
            elif t == 0x0003:
                cls, size = b[0] & 0x0F, struct.unpack("<I", b[4:8])[0]
                order = ">" if (b[1] & 1) else "<"
                if cls == 0:
                    dtype = np.dtype(f"{order}{'i' if b[1] & 0x08 else 'u'}{size}")
                elif cls == 1:
                    dtype = np.dtype(f"{order}f{size}")
                else:
                    raise NotImplementedError(f"datatype class {cls}")
            elif t == 0x0008:
                ver = b[0]
                if ver in (1, 2):                     # HDF5 <= 1.6 writers: dimensionality, class, address, 32-bit dims
                    if b[2] != 1:
                        raise NotImplementedError("only contiguous datasets (layout message version 1 / 2)")
                    layout = ("contiguous", struct.unpack("<Q", b[8:16])[0], 0)
                    continue
                if ver != 3:
                    raise NotImplementedError(f"data layout message version {ver}")
                if b[1] == 1:
                    layout = ("contiguous",) + struct.unpack("<QQ", b[2:18])
                elif b[1] == 0:
                    n = struct.unpack("<H", b[2:4])[0]
                    layout = ("compact", b[4:4 + n])
                else:
                    raise NotImplementedError("chunked dataset (compression / resizable) is outside h5mini's subset")
        if shape is None or dtype is None or layout is None:
            return None
        n = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
        if layout[0] == "compact":
            raw = layout[1]
        elif layout[1] == UNDEF:
            raw = b"\0" * (n * dtype.itemsize)
        else:
            raw = self.read(layout[1], n * dtype.itemsize)
        return np.frombuffer(raw, dtype=dtype, count=n).reshape(shape).astype(dtype.newbyteorder("="), copy=True)


def _open_group(F: _File, group: str):
    addr = F.root_header
    for part in [p for p in group.split("/") if p]:
        m = F.members(addr)
        if m is None or part not in m:
            raise KeyError(group)
        addr = m[part]
    return addr


def list_datasets(file_name: str, group: str = "/") -> dict:
    """{name: object header address} of the members of `group`."""
    F = _File(file_name)
    try:
        return F.members(_open_group(F, group)) or {}
    finally:
        F.close()


def read_datasets(file_name: str, group: str, names=None) -> dict:
    """The named (default: all) datasets of `group` as numpy arrays."""
    F = _File(file_name)
    try:
        m = F.members(_open_group(F, group))
        if m is None:
            raise KeyError(f"{group} is not a group")
        out = {}
        for k in (names if names is not None else m):
            if k not in m:
                raise KeyError(f"{group}/{k}")
            out[k] = F.dataset(m[k])
        return out
    finally:
        F.close()

"""Minimal HDF5 writer/reader for the flat files ``SPGG.run`` produces.

The reference writes one flat HDF5 file through ``h5py`` (``src/model/spgg.py:339,
397-402, 595-633``): root group only, contiguous little-endian ``float64`` / ``int64``
datasets, no attributes, no chunking, no compression.  ``h5py`` / ``libhdf5`` are not
installed in the build image, so this module writes that subset of the HDF5 1.8 file
format directly (superblock version 0, version-1 object headers, one symbol-table
group with a version-1 B-tree + local heap, contiguous layout version 3) and reads it
back.  If ``h5py`` is importable, ``open_file`` prefers it.

Validation status: round trip (writer -> reader) and structure checks in
``tests/test_h5lite.py``; not cross-checked against libhdf5 in this image.
"""
from __future__ import annotations

import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * ((-len(b)) % 8)


def _dtype_message(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt == np.float64:
        # class 1 (floating point), version 1; LE, IEEE implied-msb mantissa, sign bit 63
        head = bytes([0x11, 0x20, 0x3F, 0x00]) + struct.pack("<I", 8)
        prop = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
        return head + prop
    if dt == np.float32:
        head = bytes([0x11, 0x20, 0x1F, 0x00]) + struct.pack("<I", 4)
        prop = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
        return head + prop
    if dt.kind in "iu" and dt.itemsize in (1, 2, 4, 8):
        signed = 0x08 if dt.kind == "i" else 0x00
        head = bytes([0x10, signed, 0x00, 0x00]) + struct.pack("<I", dt.itemsize)
        prop = struct.pack("<HH", 0, 8 * dt.itemsize)
        return head + prop
    raise TypeError(f"h5lite: unsupported dtype {dt}")


def _message(mtype: int, data: bytes, flags: int = 0) -> bytes:
    data = _pad8(data)
    return struct.pack("<HHB3x", mtype, len(data), flags) + data


def _object_header(messages) -> bytes:
    body = b"".join(messages)
    # version 1, reserved, #messages, reference count 1, header size; prefix padded to 16 bytes
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body


class _OnDisk:
    """A dataset already streamed to the file (large arrays are not kept in memory)."""

    def __init__(self, path, addr, shape, dtype):
        self.path, self.addr, self.shape, self.dtype = path, addr, tuple(shape), np.dtype(dtype)

    def load(self):
        n = int(np.prod(self.shape)) if self.shape else 1
        return np.fromfile(self.path, dtype=self.dtype, count=n, offset=self.addr).reshape(self.shape)


class _Writer:
    """Streams every dataset's bytes to the file when it is created (like h5py does - a 16.7 M-site
    lattice snapshot is written once, straight from the array's buffer) and appends the metadata
    (root header, B-tree, heap, symbol node, dataset headers) behind the data on ``close``; the
    superblock at offset 0 is written last.  Addresses in HDF5 are arbitrary, so readers do not care
    where the metadata lives."""

    KEEP_BYTES = 1 << 20   # smaller arrays stay readable through ``File[name]`` without touching the disk

    def __init__(self, path):
        self.path = path
        self.datasets = {}
        self._meta = {}       # name -> (data address or UNDEF, shape, dtype, nbytes)
        self._f = open(path, "wb")
        self._f.write(b"\0" * 96)   # superblock placeholder
        self._pos = 96

    def create_dataset(self, name, data=None, **_ignored):
        arr = np.asarray(data)
        if arr.dtype == np.bool_:
            arr = arr.astype(np.int8)
        if arr.dtype.kind == "f" and arr.dtype.itemsize not in (4, 8):
            arr = arr.astype(np.float64)
        if arr.dtype.kind not in "fiu":
            arr = arr.astype(np.float64)
        if name in self.datasets:
            raise ValueError(f"Unable to create dataset (name already exists): {name}")
        _dtype_message(arr.dtype)    # unsupported dtypes fail here, before anything is written
        arr = np.ascontiguousarray(arr, dtype=arr.dtype.newbyteorder("<"))
        addr = UNDEF
        if arr.nbytes:
            pad = (-self._pos) % 8
            if pad:
                self._f.write(b"\0" * pad)
                self._pos += pad
            addr = self._pos
            self._f.write(memoryview(arr.reshape(-1)).cast("B"))
            self._pos += arr.nbytes
        self._meta[name] = (addr, arr.shape, arr.dtype, arr.nbytes)
        self.datasets[name] = arr.copy() if arr.nbytes <= self.KEEP_BYTES else _OnDisk(self.path, addr, arr.shape, arr.dtype)
        if isinstance(self.datasets[name], _OnDisk):
            self._f.flush()
        return self.datasets[name]

    def close(self):
        if self._f is None:
            return
        names = sorted(self._meta, key=lambda s: s.encode())  # strcmp order
        n = len(names)
        leaf_k = max(4, (n + 1) // 2)          # one SNOD holds 2*leaf_k symbols
        internal_k = 16
        # ---- local heap data: "" at offset 0, then the names
        heap = bytearray(_pad8(b"\0"))
        name_off = {}
        for nm in names:
            name_off[nm] = len(heap)
            heap += _pad8(nm.encode() + b"\0")
        heap_data = bytes(heap)
        # ---- layout of the metadata block, behind the data
        meta0 = self._pos + ((-self._pos) % 8)
        off_root_hdr = meta0
        root_hdr = _object_header([_message(0x0011, struct.pack("<QQ", 0, 0))])  # size only; emitted below
        off_btree = off_root_hdr + len(root_hdr)
        btree_size = 24 + (2 * internal_k + 1) * 8 + 2 * internal_k * 8
        off_heap = off_btree + btree_size
        heap_hdr_size = 32
        off_heap_data = off_heap + heap_hdr_size
        off_snod = off_heap_data + len(heap_data)
        snod_size = 8 + 2 * leaf_k * 40
        pos = off_snod + snod_size
        hdr_addr, headers = {}, {}
        for nm in names:
            addr, dims, dt, nbytes = self._meta[nm]
            space = struct.pack("<BBB5x", 1, len(dims), 0) + b"".join(struct.pack("<Q", d) for d in dims)
            fill = struct.pack("<BBBB", 2, 2, 2, 0)   # v2: late allocation, write if set, undefined
            layout = struct.pack("<BBQQ", 3, 1, addr if nbytes else UNDEF, nbytes)
            headers[nm] = _object_header([
                _message(0x0001, space), _message(0x0003, _dtype_message(dt), 1),
                _message(0x0005, fill, 1), _message(0x0008, layout)])
            hdr_addr[nm] = pos
            pos += len(headers[nm])
            pos += (-pos) % 8
        eof = pos
        # ---- emit the metadata block
        out = bytearray(eof - meta0)
        def put(addr, b):
            out[addr - meta0:addr - meta0 + len(b)] = b
        put(off_root_hdr, _object_header([_message(0x0011, struct.pack("<QQ", off_btree, off_heap))]))
        # B-tree node: group node, level 0, one child (the SNOD) when there are symbols
        used = 1 if n else 0
        bt = b"TREE" + struct.pack("<BBHQQ", 0, 0, used, UNDEF, UNDEF)
        keys_children = struct.pack("<Q", 0)
        if n:
            keys_children += struct.pack("<QQ", off_snod, name_off[names[-1]])
        bt += keys_children
        bt += b"\0" * (btree_size - len(bt))
        put(off_btree, bt)
        # local heap header: no free blocks (free-list head = 1 == H5HL_FREE_NULL)
        put(off_heap, b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), 1, off_heap_data))
        put(off_heap_data, heap_data)
        sn = b"SNOD" + struct.pack("<BBH", 1, 0, n)
        for nm in names:
            sn += struct.pack("<QQII16x", name_off[nm], hdr_addr[nm], 0, 0)
        sn += b"\0" * (snod_size - len(sn))
        put(off_snod, sn)
        for nm in names:
            put(hdr_addr[nm], headers[nm])
        self._f.write(b"\0" * (meta0 - self._pos))
        self._f.write(bytes(out))
        # ---- superblock (version 0) and the root symbol-table entry: name offset, header address,
        # cache type 1, reserved, scratch = B-tree and heap addresses
        sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, leaf_k, internal_k, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        sb += struct.pack("<QQII", 0, off_root_hdr, 1, 0) + struct.pack("<QQ", off_btree, off_heap)
        assert len(sb) == 96
        self._f.seek(0)
        self._f.write(sb)
        self._f.close()
        self._f = None


class _Dataset:
    def __init__(self, arr):
        self._arr = arr
        self.shape = arr.shape
        self.dtype = arr.dtype

    def __getitem__(self, key):
        return self._arr[key]

    def __array__(self, dtype=None, copy=None):
        return self._arr if dtype is None else self._arr.astype(dtype)

    def __len__(self):
        return len(self._arr)


def _read(path):
    with open(path, "rb") as f:
        buf = f.read()
    if buf[:8] != SIGNATURE:
        raise OSError(f"{path}: not an HDF5 file")
    if buf[8] != 0 or buf[13] != 8 or buf[14] != 8:
        raise OSError("h5lite reads only superblock version 0 files with 8-byte offsets")
    root_hdr, = struct.unpack_from("<Q", buf, 56 + 8)
    btree, heap = struct.unpack_from("<QQ", buf, 56 + 24)
    assert buf[heap:heap + 4] == b"HEAP"
    heap_data, = struct.unpack_from("<Q", buf, heap + 24)

    def name_at(off):
        end = buf.index(b"\0", heap_data + off)
        return buf[heap_data + off:end].decode()

    def walk(node):
        assert buf[node:node + 4] == b"TREE"
        _ntype, level, used = struct.unpack_from("<BBH", buf, node + 4)
        for k in range(used):
            child, = struct.unpack_from("<Q", buf, node + 24 + 8 + 16 * k)
            if level:
                yield from walk(child)
            else:
                assert buf[child:child + 4] == b"SNOD"
                nsym, = struct.unpack_from("<H", buf, child + 6)
                for s in range(nsym):
                    noff, haddr = struct.unpack_from("<QQ", buf, child + 8 + 40 * s)
                    yield name_at(noff), haddr

    out = {}
    for name, haddr in walk(btree):
        ver, _r, nmsg, _rc, hsize = struct.unpack_from("<BBHII", buf, haddr)
        assert ver == 1
        p, shape, dt, daddr, dsize = haddr + 16, None, None, None, None
        for _ in range(nmsg):
            mtype, msize = struct.unpack_from("<HH", buf, p)
            d = p + 8
            if mtype == 0x0001:
                rank = buf[d + 1]
                shape = tuple(struct.unpack_from("<Q", buf, d + 8 + 8 * i)[0] for i in range(rank))
            elif mtype == 0x0003:
                cls = buf[d] & 0x0F
                size, = struct.unpack_from("<I", buf, d + 4)
                if cls == 1:
                    dt = np.dtype(f"<f{size}")
                else:
                    dt = np.dtype(("<i" if buf[d + 1] & 0x08 else "<u") + str(size))
            elif mtype == 0x0008:
                assert buf[d] == 3 and buf[d + 1] == 1, "contiguous layout v3 only"
                daddr, dsize = struct.unpack_from("<QQ", buf, d + 2)
            p = d + msize
        n = int(np.prod(shape)) if shape else 1
        if n == 0 or daddr == UNDEF:
            arr = np.zeros(shape, dt)
        else:
            arr = np.frombuffer(buf, dt, count=n, offset=daddr).reshape(shape).copy()
        out[name] = arr
    return out


class File:
    """``h5py.File``-like object for the flat-file subset (modes ``"w"`` and ``"r"``)."""

    def __init__(self, name, mode="r"):
        self.filename = name
        self.mode = mode
        if mode.startswith("w"):
            self._w = _Writer(name)
            self._data = None
        elif mode.startswith("r"):
            self._w = None
            self._data = _read(name)
        else:
            raise ValueError(f"h5lite: unsupported mode {mode!r}")

    def create_dataset(self, name, data=None, **kw):
        if self._w is None:
            raise OSError("file is not open for writing")
        return self._w.create_dataset(name, data=data, **kw)

    def keys(self):
        return (self._w.datasets if self._w is not None else self._data).keys()

    def __contains__(self, name):
        return name in self.keys()

    def __getitem__(self, name):
        src = self._w.datasets if self._w is not None else self._data
        d = src[name]
        return _Dataset(d.load() if isinstance(d, _OnDisk) else d)

    def close(self):
        if self._w is not None:
            self._w.close()
            self._w = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


def open_file(name, mode="r"):
    """``h5py.File`` when h5py is installed, else the built-in subset."""
    try:
        import h5py  # type: ignore
        return h5py.File(name, mode)
    except ImportError:
        return File(name, mode)

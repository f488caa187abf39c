"""``h5lite``: the flat-file HDF5 subset ``SPGG.run`` writes (reference spgg.py:339, 397-402,
595-633 write through h5py, which is not installed here).  Round trip and structure
checks against the HDF5 1.8 file-format specification (superblock v0, symbol-table group,
v1 object headers, contiguous layout v3)."""
import struct

import numpy as np
import pytest

from spgg_b200 import h5lite


def _sample():
    rs = np.random.RandomState(0)
    return {
        "coop_rate_history": rs.rand(17),
        "it_records_final": rs.rand(17, 6),
        "switch_C_to_D": rs.randint(0, 100, 17).astype(np.int64),
        "Sn_final": rs.randint(0, 2, (12, 12)).astype(np.int64),
        "R_final": rs.randint(-10, 11, (12, 12)).astype(np.float64),
        "rep_hist_final": rs.randint(0, 50, 20).astype(np.int64),
        "q_c_pos_6_6_final": np.zeros(0),
        "reputation_reward_ratio": np.array([1.0, np.nan, 3.0]),
    }


def test_round_trip(tmp_path):
    path = str(tmp_path / "a.h5")
    data = _sample()
    with h5lite.File(path, "w") as f:
        for k, v in data.items():
            f.create_dataset(k, data=v)
        assert "Sn_final" in f
    with h5lite.File(path, "r") as f:
        assert sorted(f.keys()) == sorted(data)
        for k, v in data.items():
            got = f[k][:]
            assert got.dtype == v.dtype and got.shape == v.shape, k
            assert np.array_equal(got, v, equal_nan=True), k
            assert np.array_equal(np.array(f[k]), v, equal_nan=True)


def test_many_datasets_and_duplicate_name(tmp_path):
    path = str(tmp_path / "b.h5")
    with h5lite.File(path, "w") as f:
        for i in range(90):                       # the reference writes ~50-86 datasets per run
            f.create_dataset(f"series_{i:03d}", data=np.full(5, float(i)))
        with pytest.raises(ValueError):
            f.create_dataset("series_000", data=np.zeros(1))
    with h5lite.File(path, "r") as f:
        assert len(list(f.keys())) == 90
        assert f["series_042"][0] == 42.0


def test_file_structure_follows_the_format_spec(tmp_path):
    path = str(tmp_path / "c.h5")
    with h5lite.File(path, "w") as f:
        f.create_dataset("b", data=np.arange(4, dtype=np.int64))
        f.create_dataset("a", data=np.arange(3.0))
    buf = open(path, "rb").read()
    assert buf[:8] == b"\x89HDF\r\n\x1a\n"
    # superblock v0: versions, sizes of offsets/lengths, group K values
    assert buf[8] == 0 and buf[13] == 8 and buf[14] == 8
    base, _free, eof, _drv = struct.unpack_from("<QQQQ", buf, 24)
    assert base == 0 and eof == len(buf)
    # root symbol-table entry caches the B-tree and heap addresses
    _noff, root_hdr, cache, _res = struct.unpack_from("<QQII", buf, 56)
    btree, heap = struct.unpack_from("<QQ", buf, 80)
    assert cache == 1 and buf[btree:btree + 4] == b"TREE" and buf[heap:heap + 4] == b"HEAP"
    assert buf[root_hdr] == 1                     # object header version 1
    # one leaf: symbols sorted by name, as the library's binary search expects
    child, = struct.unpack_from("<Q", buf, btree + 24 + 8)
    assert buf[child:child + 4] == b"SNOD"
    nsym, = struct.unpack_from("<H", buf, child + 6)
    assert nsym == 2
    heap_data, = struct.unpack_from("<Q", buf, heap + 24)
    offs = [struct.unpack_from("<Q", buf, child + 8 + 40 * s)[0] for s in range(nsym)]
    names = [buf[heap_data + o:buf.index(b"\0", heap_data + o)].decode() for o in offs]
    assert names == ["a", "b"]
    # every dataset header and data block is 8-byte aligned
    for s in range(nsym):
        haddr, = struct.unpack_from("<Q", buf, child + 8 + 40 * s + 8)
        assert haddr % 8 == 0


def test_unsupported_inputs(tmp_path):
    with pytest.raises(ValueError):
        h5lite.File(str(tmp_path / "x.h5"), "a")
    bad = tmp_path / "bad.h5"
    bad.write_bytes(b"not an hdf5 file")
    with pytest.raises(OSError):
        h5lite.File(str(bad), "r")
    with h5lite.File(str(tmp_path / "y.h5"), "w") as f:
        f.create_dataset("flags", data=np.array([True, False]))
        f.create_dataset("f32", data=np.arange(3, dtype=np.float32))
    with h5lite.File(str(tmp_path / "y.h5"), "r") as f:
        assert f["flags"][:].tolist() == [1, 0]
        assert f["f32"].dtype == np.float32
        with pytest.raises(OSError):
            f.create_dataset("z", data=np.zeros(1))


def test_large_datasets_stream_to_disk_and_read_back(tmp_path):
    """Datasets above the in-memory threshold are written when created (no staging copy of the
    whole file) and stay readable through the open writer and after reopening."""
    path = str(tmp_path / "big.h5")
    rs = np.random.RandomState(0)
    big = rs.rand(600, 600)                       # 2.9 MB > KEEP_BYTES
    ints = rs.randint(0, 2, (700, 700))           # int64, 3.9 MB
    with h5lite.File(path, "w") as f:
        f.create_dataset("small_before", data=np.arange(5.0))
        f.create_dataset("R_snapshot_1", data=big)
        f.create_dataset("Sn_snapshot_1", data=ints)
        f.create_dataset("small_after", data=np.arange(7, dtype=np.int64))
        assert np.array_equal(f["R_snapshot_1"][:], big)          # read back from the open writer
        assert f["Sn_snapshot_1"].shape == (700, 700)
        big[0, 0] = -1.0                                          # the file already holds the data
    with h5lite.File(path, "r") as f:
        assert sorted(f.keys()) == ["R_snapshot_1", "Sn_snapshot_1", "small_after", "small_before"]
        got = f["R_snapshot_1"][:]
        assert got[0, 0] != -1.0 and np.array_equal(got[1:], big[1:])
        assert np.array_equal(f["Sn_snapshot_1"][:], ints)
        assert f["small_after"][:].tolist() == list(range(7))
    # every data block is 8-byte aligned and inside the file
    buf = open(path, "rb").read()
    eof, = struct.unpack_from("<Q", buf, 40)
    assert eof == len(buf)

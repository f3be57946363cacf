"""Per-step trajectory files in the reference's HDF5 layout (reference h5file.c:21-112),
written by host/heroam_h5.c without libhdf5.

Three readers look at every file: the C reader of host/heroam_h5.c (the `open_file` /
`read_timestep_from_hdf5` replacement), the independent Python reader tests/h5lite.py, and --
for the structural invariants the real library relies on -- explicit checks of the B-tree
ordering, node sizes and addresses.  h5lite itself is pinned against a file written by the
real libhdf5 that happens to ship with scipy's test data.
"""
import os
import struct

import numpy as np
import pytest

import h5lite
from sphere_roaming_b200 import host
from sphere_roaming_b200.types import PARTICLE_DTYPE

LIBHDF5_FILE = os.path.join(os.path.dirname(__import__("scipy").__file__), "io", "matlab", "tests", "data",
                            "testhdf5_7.4_GLNX86.mat")

#: the compound of h5file.c:31-47: (name, offset, numpy type)
MEMBERS = [("index", 0, "<u4"), ("success", 4, "<u4"), ("time", 8, "<f4"), ("force_mag", 12, "<f4"),
           ("velocity_sq", 16, "<f4"), ("orientation", 20, ("<f4", (4,))), ("position", 36, ("<f4", (3,))),
           ("velocity", 48, ("<f4", (3,))), ("ang_velocity", 60, ("<f4", (3,))), ("force", 72, ("<f4", (3,)))]


def random_records(rng, n):
    raw = rng.integers(0, 256, size=n * PARTICLE_DTYPE.itemsize, dtype=np.uint8)
    rec = raw.view(PARTICLE_DTYPE).copy()
    rec["index"] = np.arange(n)
    return rec


@pytest.mark.skipif(not os.path.exists(LIBHDF5_FILE), reason="scipy test data not installed")
def test_reader_parses_a_file_written_by_libhdf5():
    """Pins tests/h5lite.py: superblock behind a 512-byte user block, symbol-table root group,
    local heap, B-tree, symbol node, version-1 object header, float64 dataset."""
    f = h5lite.File(LIBHDF5_FILE)
    assert (f.super_version, f.base, f.leaf_k, f.internal_k) == (0, 512, 4, 16)
    entries = f.group_entries(f.root_header)
    assert list(entries) == ["testdouble"]
    ds = f.dataset(entries["testdouble"][0])
    assert ds.dtype == np.dtype("<f8") and ds.shape == (9, 1)
    np.testing.assert_allclose(ds.data.ravel(), np.arange(9) * np.pi / 4, rtol=0, atol=1e-15)


@pytest.mark.parametrize("n_steps,no", [(1, 1), (3, 4), (9, 2), (257, 5), (3000, 3), (40000, 1)])  # 1..3 B-tree levels
def test_write_then_read_back(tmp_path, n_steps, no):
    rng = np.random.default_rng(n_steps * 131 + no)
    path = str(tmp_path / "trajectory_000007.h5")
    steps = [random_records(rng, no) for _ in range(n_steps)]
    with host.TrajectoryFileWriter(path) as w:
        for k, rec in enumerate(steps):
            w.save_timestep(k, rec)
    # (1) the C reader: H5Gget_num_objs + read_timestep_from_hdf5
    with host.TrajectoryFileReader(path) as r:
        assert r.num_steps() == n_steps
        for k in (range(n_steps) if n_steps < 300 else rng.choice(n_steps, 100, replace=False)):
            got = r.read_timestep(int(k))
            assert got.tobytes() == steps[int(k)].tobytes()
        with pytest.raises(KeyError):
            r.read_timestep(n_steps)
    # (2) the independent reader
    f = h5lite.File(path)
    assert (f.super_version, f.base, f.leaf_k, f.internal_k) == (0, 0, 4, 16)
    assert f.eof_address == os.path.getsize(path)
    assert f.root_cache_type == 1
    root = f.group_entries(f.root_header)
    assert sorted(root) == sorted(f"Step_{k}" for k in range(n_steps))
    assert list(root) == sorted(root)  # B-tree order is strcmp order, not numeric order
    # the root symbol-table entry caches what the root header's message says
    stab = [d for t, _, d in f.messages(f.root_header) if t == 0x0011][0]
    assert struct.unpack("<QQ", stab[:16]) == f.root_scratch
    for k in (range(n_steps) if n_steps < 300 else rng.choice(n_steps, 60, replace=False)):
        header, cache, scratch = root[f"Step_{int(k)}"]
        assert cache == 1
        gstab = [d for t, _, d in f.messages(header) if t == 0x0011][0]
        assert struct.unpack("<QQ", gstab[:16]) == scratch
        members = f.group_entries(header)
        assert list(members) == ["particles"]
        ds = f.dataset(members["particles"][0])
        assert ds.shape == (no,)
        assert ds.dtype.itemsize == 84
        assert [(n, ds.dtype.fields[n][1], ds.dtype.fields[n][0]) for n in ds.dtype.names] == \
               [(n, off, np.dtype(t)) for n, off, t in MEMBERS]
        assert ds.data.tobytes() == steps[int(k)].tobytes()
        # the messages libhdf5 writes for H5Dcreate with default property lists
        assert [t for t, _, _ in ds.messages] == [0x0001, 0x0003, 0x0005, 0x0008, 0x0012]


def test_datatype_message_bytes(tmp_path):
    """The encoded compound, byte for byte, against the format specification: class 6 version 2
    (the version libhdf5 uses once a member is an array), array members class 10 version 2."""
    path = str(tmp_path / "t.h5")
    with host.TrajectoryFileWriter(path) as w:
        w.save_timestep(0, np.zeros(1, dtype=PARTICLE_DTYPE))
    f = h5lite.File(path)
    ds = f.dataset(f.group_entries(f.group_entries(f.root_header)["Step_0"][0])["particles"][0])
    dt = [d for t, _, d in ds.messages if t == 0x0003][0]
    u32 = bytes([0x10, 0, 0, 0]) + struct.pack("<IHH", 4, 0, 32)
    f32 = bytes([0x11, 0x20, 0x1F, 0]) + struct.pack("<IHHBBBBI", 4, 0, 32, 23, 8, 0, 23, 127)

    def arr(n):
        return bytes([0x2A, 0, 0, 0]) + struct.pack("<IB3xII", 4 * n, 1, n, 0) + f32

    def member(name, off, body):
        raw = name.encode() + b"\0"
        raw += b"\0" * (-len(raw) % 8)
        return raw + struct.pack("<I", off) + body

    want = bytes([0x26, 10, 0, 0]) + struct.pack("<I", 84)
    for name, off, t in MEMBERS:
        want += member(name, off, u32 if t == "<u4" else f32 if t == "<f4" else arr(t[1][0]))
    want += b"\0" * (-len(want) % 8)
    assert dt == want


def test_steps_out_of_order_and_varying_counts(tmp_path):
    rng = np.random.default_rng(5)
    path = str(tmp_path / "t.h5")
    order = [5, 0, 12, 3, 100, 7]
    recs = {k: random_records(rng, 1 + k % 4) for k in order}
    with host.TrajectoryFileWriter(path) as w:
        for k in order:
            w.save_timestep(k, recs[k])
        with pytest.raises(OSError):
            w.save_timestep(3, recs[3])  # a group can be created once
    with host.TrajectoryFileReader(path) as r:
        assert r.num_steps() == len(order)
        for k in order:
            assert r.read_timestep(k).tobytes() == recs[k].tobytes()
    f = h5lite.File(path)
    assert list(f.group_entries(f.root_header)) == sorted(f"Step_{k}" for k in order)


def test_empty_file_and_empty_trajectory(tmp_path):
    path = str(tmp_path / "empty.h5")
    host.TrajectoryFileWriter(path).close()
    with host.TrajectoryFileReader(path) as r:
        assert r.num_steps() == 0
    f = h5lite.File(path)
    assert f.group_entries(f.root_header) == {}
    path = str(tmp_path / "zero.h5")
    with host.TrajectoryFileWriter(path) as w:
        w.save_timestep(0, np.zeros(0, dtype=PARTICLE_DTYPE))
    with host.TrajectoryFileReader(path) as r:
        assert len(r.read_timestep(0)) == 0


def test_c_reader_reads_libhdf5_groups():
    """The C reader walks a symbol-table root group written by the real library (it finds no
    Step_ groups there, but counts the objects as H5Gget_num_objs would)."""
    if not os.path.exists(LIBHDF5_FILE):
        pytest.skip("scipy test data not installed")
    with host.TrajectoryFileReader(LIBHDF5_FILE) as r:
        assert r.num_steps() == 1


def test_reader_rejects_garbage(tmp_path):
    p = tmp_path / "junk.h5"
    p.write_bytes(b"not an hdf5 file" * 100)
    with pytest.raises(OSError):
        host.TrajectoryFileReader(str(p))

"""The minimal HDF5 writer/reader of the host layer (alis_b200/host/h5lite.c; reference h5write /
h5read, world.c:469-537, which use libhdf5 -- absent from this image).

* round trips of rank 1..3 datasets through the C functions;
* the written bytes parsed by an independent Python walk of the on-disk structures (HDF5 File
  Format Specification: version-0 superblock -> root symbol-table entry -> group B-tree -> symbol
  table node -> version-1 object header -> dataspace / datatype / layout messages);
* files the writer never produces (float64, big-endian, compact layout, a user block in front) are
  built here byte by byte and must be read correctly; unsupported ones must be refused with a message.
The end-to-end use (sliceFreqDom -> h5 -> guidedWave*) is pinned against the unmodified reference
by the guided_y / guided_xz golden configurations."""
import ctypes as C
import os
import struct

import numpy as np
import pytest

from tests import helpers

LIB = os.path.join(helpers.TEST_BIN, "libalis_oracle.so")
pytestmark = pytest.mark.skipif(not os.path.exists(LIB), reason="tests/bin not built (run __graft_entry__.build())")
UNDEF = 0xFFFFFFFFFFFFFFFF


@pytest.fixture(scope="module")
def lib():
    os.environ.setdefault("LD_LIBRARY_PATH", "")
    C.CDLL(os.path.join(helpers.ROOT, "oracle", "liboracle_alis.so"), mode=C.RTLD_GLOBAL)
    L = C.CDLL(LIB)
    L.h5lite_write.restype = C.c_int
    L.h5lite_write.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_ulonglong), C.POINTER(C.c_float)]
    L.h5lite_read.restype = C.c_void_p
    L.h5lite_read.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_ulonglong), C.c_char_p, C.c_size_t]
    return L


def write(lib, path, a):
    a = np.ascontiguousarray(a, np.float32)
    dims = (C.c_ulonglong * 3)(*a.shape)
    assert lib.h5lite_write(path.encode(), a.ndim, dims, a.ctypes.data_as(C.POINTER(C.c_float))) == 0


def read(lib, path):
    rank, dims, err = C.c_int(), (C.c_ulonglong * 3)(), C.create_string_buffer(200)
    p = lib.h5lite_read(path.encode(), C.byref(rank), dims, err, 200)
    if not p:
        raise RuntimeError(err.value.decode())
    shape = tuple(dims[q] for q in range(rank.value))
    out = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(int(np.prod(shape)),)).copy().reshape(shape)
    C.CDLL(None).free(C.c_void_p(p))
    return out


def parse(raw):
    """independent walk of the classic HDF5 structures -> (dims, dtype info, data address, data size)"""
    assert raw[:8] == b"\x89HDF\r\n\x1a\n"
    sb_ver, fs_ver, root_ver, _, shm_ver, so, sl, _ = raw[8:16]
    assert (sb_ver, fs_ver, root_ver, shm_ver, so, sl) == (0, 0, 0, 0, 8, 8)
    leaf_k, internal_k, flags = struct.unpack_from("<HHI", raw, 16)
    base, free, eof, driver = struct.unpack_from("<4Q", raw, 24)
    assert (base, free, driver) == (0, UNDEF, UNDEF) and eof == len(raw)
    name_off, root_oh, cache, _, btree, heap = struct.unpack_from("<QQIIQQ", raw, 56)
    assert cache == 1 and name_off == 0
    # root object header: a symbol table message with the same two addresses
    ver, _, nmsg, refs, size = struct.unpack_from("<BBHII", raw, root_oh)
    assert (ver, nmsg, refs) == (1, 1, 1)
    mtype, msize = struct.unpack_from("<HH", raw, root_oh + 16)
    assert mtype == 0x11 and struct.unpack_from("<QQ", raw, root_oh + 24) == (btree, heap)
    # B-tree leaf -> symbol table node
    assert raw[btree:btree + 4] == b"TREE"
    ntype, level, used, left, right = struct.unpack_from("<BBHQQ", raw, btree + 4)
    assert (ntype, level, used, left, right) == (0, 0, 1, UNDEF, UNDEF)
    key0, snod, key1 = struct.unpack_from("<QQQ", raw, btree + 24)
    assert heap >= btree + 24 + (2 * internal_k + 1) * 8 + 2 * internal_k * 8       # the node has its full size
    assert raw[heap:heap + 4] == b"HEAP"
    seg_size, free_head, seg = struct.unpack_from("<QQQ", raw, heap + 8)
    names = raw[seg:seg + seg_size]
    assert names[key0:key0 + 1] == b"\0" and names[key1:key1 + 5] == b"data\0"
    assert struct.unpack_from("<QQ", names, free_head) == (1, seg_size - free_head)
    assert raw[snod:snod + 4] == b"SNOD" and struct.unpack_from("<BBH", raw, snod + 4) == (1, 0, 1)
    link, oh = struct.unpack_from("<QQ", raw, snod + 8)
    assert link == key1 and oh >= snod + 8 + 2 * leaf_k * 40
    # dataset object header
    ver, _, nmsg, refs, size = struct.unpack_from("<BBHII", raw, oh)
    assert ver == 1 and refs == 1
    pos, info = oh + 16, {}
    for _ in range(nmsg):
        mtype, msize, mflags = struct.unpack_from("<HHB", raw, pos)
        d = raw[pos + 8:pos + 8 + msize]
        assert msize % 8 == 0
        if mtype == 1:
            assert d[0] == 1 and d[2] == 0
            info["dims"] = struct.unpack_from(f"<{d[1]}Q", d, 8)
        elif mtype == 3:
            info["class"], info["order"], info["size"] = d[0] & 15, d[1] & 1, struct.unpack_from("<I", d, 4)[0]
            assert d[0] >> 4 == 1 and struct.unpack_from("<HHBBBBI", d, 8) == (0, 32, 23, 8, 0, 23, 127)
        elif mtype == 8:
            assert d[0] == 3 and d[1] == 1
            info["addr"], info["bytes"] = struct.unpack_from("<QQ", d, 2)
        elif mtype == 5:
            assert d[0] == 2
        pos += 8 + msize
    assert pos == oh + 16 + size
    return info


@pytest.mark.parametrize("shape", [(7,), (5, 9), (3, 4, 6), (1, 1), (64, 33)])
def test_round_trip_and_on_disk_structure(lib, tmp_path, shape):
    rng = np.random.RandomState(len(shape))
    a = rng.randn(*shape).astype(np.float32)
    path = str(tmp_path / "a.h5")
    write(lib, path, a)
    assert np.array_equal(read(lib, path), a)
    raw = open(path, "rb").read()
    info = parse(raw)
    assert info["dims"] == shape and (info["class"], info["order"], info["size"]) == (1, 0, 4)
    assert info["addr"] % 8 == 0 and info["bytes"] == a.nbytes and info["addr"] + a.nbytes == len(raw)
    assert np.array_equal(np.frombuffer(raw, "<f4", a.size, info["addr"]).reshape(shape), a)


def test_reader_accepts_other_producers_layouts(lib, tmp_path):
    a = np.arange(24, dtype=np.float32).reshape(4, 6) / 7
    path = str(tmp_path / "a.h5")
    write(lib, path, a)
    raw = bytearray(open(path, "rb").read())
    info = parse(bytes(raw))
    dt = raw.find(bytes([0x11, 0x20, 31, 0]) + struct.pack("<I", 4))          # the datatype message body
    assert dt > 0
    # (1) big-endian float64 data
    be = bytearray(raw[:info["addr"]])
    be[dt + 1] |= 1
    be[dt + 4:dt + 8] = struct.pack("<I", 8)
    lay = be.find(bytes([3, 1]) + struct.pack("<Q", info["addr"]))
    be[lay + 10:lay + 18] = struct.pack("<Q", a.size * 8)
    be += a.astype(">f8").tobytes()
    be[40:48] = struct.pack("<Q", len(be))
    p2 = str(tmp_path / "be.h5")
    open(p2, "wb").write(be)
    assert np.array_equal(read(lib, p2), a)
    # (2) a 512-byte user block in front of the superblock
    p3 = str(tmp_path / "ub.h5")
    open(p3, "wb").write(b"\0" * 512 + bytes(raw))
    assert np.array_equal(read(lib, p3), a)
    # (3) an unwritten dataset (address undefined) reads as the fill value 0
    un = bytearray(raw)
    lay = un.find(bytes([3, 1]) + struct.pack("<Q", info["addr"]))
    un[lay + 2:lay + 10] = struct.pack("<Q", UNDEF)
    p4 = str(tmp_path / "un.h5")
    open(p4, "wb").write(un)
    assert not read(lib, p4).any()


def test_reader_refuses_what_it_cannot_read(lib, tmp_path):
    a = np.ones((3, 3), np.float32)
    path = str(tmp_path / "a.h5")
    write(lib, path, a)
    raw = bytearray(open(path, "rb").read())
    for name, edit, msg in (("v2", lambda b: b.__setitem__(8, 2), "superblock version"),
                            ("chunked", lambda b: b.__setitem__(b.find(bytes([3, 1]) + struct.pack("<Q", parse(bytes(raw))["addr"])) + 1, 2), "chunked"),
                            ("notfloat", lambda b: b.__setitem__(b.find(bytes([0x11, 0x20, 31, 0])), 0x10), "float")):
        b = bytearray(raw)
        edit(b)
        p = str(tmp_path / (name + ".h5"))
        open(p, "wb").write(b)
        with pytest.raises(RuntimeError, match=msg):
            read(lib, p)
    open(str(tmp_path / "junk.h5"), "wb").write(b"not an hdf5 file" * 10)
    with pytest.raises(RuntimeError, match="not an HDF5 file"):
        read(lib, str(tmp_path / "junk.h5"))
    with pytest.raises(RuntimeError, match="cannot open"):
        read(lib, str(tmp_path / "missing.h5"))


def test_reader_bounds_hostile_sizes(lib, tmp_path):
    """sizes read from the file are checked against the message that carries them: a compact layout whose
    declared data size runs past its message, and a dataspace whose dimension product overflows, are
    refused with a message instead of being copied / allocated"""
    a = np.ones((3, 3), np.float32)
    path = str(tmp_path / "a.h5")
    write(lib, path, a)
    raw = bytearray(open(path, "rb").read())
    info = parse(bytes(raw))
    lay = raw.find(bytes([3, 1]) + struct.pack("<Q", info["addr"]))
    # (1) turn the contiguous layout message into a compact one that claims 60000 bytes of data inside an 18-byte message
    b = bytearray(raw)
    b[lay + 1] = 0
    b[lay + 2:lay + 4] = struct.pack("<H", 60000)
    open(str(tmp_path / "compact.h5"), "wb").write(b)
    with pytest.raises(RuntimeError, match="truncated header message"):
        read(lib, str(tmp_path / "compact.h5"))
    # (2) absurd dimensions
    b = bytearray(raw)
    ds = b.find(bytes([1, 2, 0, 0, 0, 0, 0, 0]) + struct.pack("<Q", 3))
    assert ds > 0
    b[ds + 8:ds + 16] = struct.pack("<Q", 1 << 62)
    b[ds + 16:ds + 24] = struct.pack("<Q", 1 << 62)
    open(str(tmp_path / "huge.h5"), "wb").write(b)
    with pytest.raises(RuntimeError, match="too large"):
        read(lib, str(tmp_path / "huge.h5"))

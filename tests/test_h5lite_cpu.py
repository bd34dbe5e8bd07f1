"""h5lite (avlmaps_b200/utils/h5lite.py): the dependency-free reader / writer of `vlmaps.h5df`
(reference avlmaps/utils/mapping_utils.py:469-541 goes through h5py, which is absent here).

Pins: (1) a file written by libhdf5 itself -- scipy ships MATLAB v7.3 test data, an HDF5 file with a 512-byte user
block, superblock v0, a symbol-table root group and a contiguous float64 dataset; (2) the structures our writer emits
are walked field by field next to the same structures of that libhdf5 file; (3) round trips of every map field;
(4) chunked + deflate + shuffle storage through a file assembled by hand in the test; (5) damaged files fail loudly.
"""
from __future__ import annotations

import struct
import zlib
from pathlib import Path

import numpy as np
import pytest

from avlmaps_b200.utils import h5lite, mapping_utils


def _scipy_hdf5_file():
    import scipy.io

    p = Path(scipy.io.__file__).parent / "matlab" / "tests" / "data" / "testhdf5_7.4_GLNX86.mat"
    if not p.exists():
        pytest.skip("scipy's MATLAB v7.3 test file is not installed")
    return p


def _map_fields(rng, v=300, d=64):
    return {
        "mapped_iter_list": np.arange(11, dtype=np.int32),
        "grid_feat": rng.standard_normal((v, d)).astype(np.float32),
        "grid_pos": rng.integers(-2, 100, (v, 3)).astype(np.int32),
        "weight": rng.random(v).astype(np.float32),
        "occupied_ids": rng.integers(-1, v, (12, 12, 5)).astype(np.int32),
        "grid_rgb": rng.integers(0, 255, (v, 3)).astype(np.uint8),
    }


def test_reads_a_file_written_by_libhdf5():
    p = _scipy_hdf5_file()
    with h5lite.File(p) as f:
        assert f.superblock_version == 0 and f._buf.base == 512       # MATLAB's user block shifts every address
        assert f.keys() == ["testdouble"] and "testdouble" in f and "nope" not in f
        ds = f["testdouble"]
        assert ds.shape == (9, 1) and ds.dtype == np.dtype("<f8")
        want = np.arange(9, dtype=np.float64).reshape(9, 1) * (np.pi / 4)
        assert np.allclose(ds[:], want, rtol=0, atol=1e-15)           # scipy's own expectation: 0 .. 2 pi
        assert np.array_equal(ds.memmap(), ds.read())
        assert ds.offset == 512 + 0xE00
    with pytest.raises(KeyError):
        with h5lite.File(p) as f:
            f["nope"]


def _walk(path):
    """Field-by-field parse of the structures of a v0 file with a symbol-table root group (independent of the
    reader's code paths: plain struct.unpack at the offsets the format specification gives)."""
    b = Path(path).read_bytes()
    s = b.index(h5lite.SIGNATURE)
    v = {}
    (v["sb_version"], v["fs_version"], v["root_version"], _, v["shm_version"], v["so"], v["sl"], _, v["leaf_k"],
     v["internal_k"], v["flags"]) = struct.unpack_from("<8BHHI", b, s + 8)
    v["base"], v["freespace"], v["eof"], v["driver"] = struct.unpack_from("<4Q", b, s + 24)
    name_off, root, cache, _, btree, heap = struct.unpack_from("<QQIIQQ", b, s + 56)
    v["root_entry"] = (name_off, cache)
    base = v["base"]
    at = lambda a: base + a                                           # noqa: E731
    # root object header
    ver, _, nmsg, ref, hsize = struct.unpack_from("<BBHII", b, at(root))
    v["root_ohdr"] = (ver, ref)
    mtype, msize, mflags = struct.unpack_from("<HHB", b, at(root) + 16)
    assert mtype == 0x11 and struct.unpack_from("<QQ", b, at(root) + 24) == (btree, heap)
    # local heap
    assert b[at(heap):at(heap) + 4] == b"HEAP"
    hver, seg_size, free_head, seg_addr = struct.unpack_from("<B3xQQQ", b, at(heap) + 4)
    v["heap_version"] = hver
    seg = b[at(seg_addr):at(seg_addr) + seg_size]
    assert seg[:8] == b"\0" * 8                                       # offset 0 = the empty string
    nxt, fsize = struct.unpack_from("<QQ", seg, free_head)
    assert nxt == 1 and free_head + fsize == seg_size                 # one free block reaching the end of the segment
    # B-tree root: a leaf with one child
    assert b[at(btree):at(btree) + 4] == b"TREE"
    ntype, level, used, left, right = struct.unpack_from("<BBHQQ", b, at(btree) + 4)
    assert (ntype, level, used, left, right) == (0, 0, 1, h5lite.UNDEF, h5lite.UNDEF)
    key0, snod, key1 = struct.unpack_from("<QQQ", b, at(btree) + 24)
    assert key0 == 0
    assert b[at(snod):at(snod) + 4] == b"SNOD"
    sver, _, nsym = struct.unpack_from("<BBH", b, at(snod) + 4)
    v["snod_version"] = sver
    names, headers = [], []
    for i in range(nsym):
        off, ohdr, ctype, _ = struct.unpack_from("<QQII", b, at(snod) + 8 + 40 * i)
        assert ctype == 0
        names.append(seg[off:seg.index(b"\0", off)].decode())
        headers.append(ohdr)
        last_off = off
    assert key1 == last_off                                           # right key = heap offset of the largest name
    assert names == sorted(names, key=str.encode)
    v["names"] = names
    # first dataset header: messages by type
    msgs = {}
    ver, _, nmsg, ref, hsize = struct.unpack_from("<BBHII", b, at(headers[0]))
    q = at(headers[0]) + 16
    for _ in range(nmsg):
        mtype, msize, mflags = struct.unpack_from("<HHB", b, q)
        msgs[mtype] = (mflags, b[q + 8:q + 8 + msize])
        q += 8 + msize
        if q >= at(headers[0]) + 16 + hsize:
            break
    v["dataset_msgs"] = msgs
    v["file_size"] = len(b)
    return v


def test_writer_emits_the_structures_libhdf5_does(tmp_path):
    ref = _walk(_scipy_hdf5_file())
    p = tmp_path / "one.h5"
    h5lite.write_file(p, {"testdouble": np.arange(9, dtype=np.float64).reshape(9, 1) * (np.pi / 4)})
    got = _walk(p)
    for k in ("sb_version", "fs_version", "root_version", "shm_version", "so", "sl", "internal_k", "freespace", "driver",
              "root_entry", "root_ohdr", "heap_version", "snod_version", "names"):
        assert got[k] == ref[k], k
    assert got["base"] == 0 and got["eof"] == got["file_size"] and ref["eof"] == ref["file_size"]
    # datatype and dataspace messages: byte for byte what libhdf5 wrote for the same array
    assert got["dataset_msgs"][3] == ref["dataset_msgs"][3]
    assert got["dataset_msgs"][1] == ref["dataset_msgs"][1]
    # fill value: same allocation / write time / "defined, size 0" fields (libhdf5 1.6 wrote message v1, we write v2)
    assert got["dataset_msgs"][5][1][1:8] == ref["dataset_msgs"][5][1][1:8]
    # layout: ours is message v3, contiguous, sized; the address lands inside the file
    ver, cls, addr, size = struct.unpack_from("<BBQQ", got["dataset_msgs"][8][1])
    assert (ver, cls, size) == (3, 1, 72) and addr + size <= got["file_size"]


def test_roundtrip_of_every_map_field(tmp_path):
    rng = np.random.default_rng(0)
    d = _map_fields(rng)
    d.update(pcd_min=rng.random(3), pcd_max=rng.random(3), cs=np.asarray(0.05), init_height_id=np.array(3, np.int32),
             empty=np.zeros((0, 64), np.float32), half=rng.random(5).astype(np.float16), be=np.arange(5, dtype=">i8"),
             u16=np.arange(4, dtype=np.uint16), f64=rng.random((3, 2, 2)), noncontig=np.arange(20, dtype=np.int64)[::2])
    p = tmp_path / "vlmaps.h5df"
    h5lite.write_file(p, d)
    with h5lite.File(p) as f:
        assert sorted(f.keys()) == sorted(d) and len(f) == len(d)
        for k, v in d.items():
            a = f[k][()]
            assert np.array_equal(a, v) and a.dtype == v.dtype and np.shape(a) == v.shape, k
        assert f["grid_feat"].offset % h5lite.DATA_ALIGN == 0          # big arrays start on a page
        assert np.array_equal(f["grid_feat"].memmap(), d["grid_feat"])
        assert f["empty"].offset is None and f["empty"].read().shape == (0, 64)
    assert not (tmp_path / "vlmaps.h5df.tmp").exists()                 # written to a temporary, then renamed
    back = h5lite.read_file(p, ["weight", "absent"])
    assert list(back) == ["weight"]


def test_mapping_utils_write_real_hdf5_without_h5py(tmp_path, monkeypatch):
    monkeypatch.setattr(mapping_utils, "_have_h5py", lambda: False)
    rng = np.random.default_rng(1)
    d = _map_fields(rng)
    p = tmp_path / "vlmaps.h5df"
    mapping_utils.save_3d_map(p, d["grid_feat"], d["grid_pos"], d["weight"], d["occupied_ids"], set(range(11)), d["grid_rgb"])
    assert p.read_bytes()[:8] == h5lite.SIGNATURE and not Path(str(p) + ".npz").exists()
    it, gf, gp, w, occ, rgb = mapping_utils.load_3d_map(p)
    assert it == list(range(11)) and np.array_equal(gf, d["grid_feat"]) and np.array_equal(occ, d["occupied_ids"])
    assert np.array_equal(rgb, d["grid_rgb"]) and np.array_equal(gp, d["grid_pos"]) and np.array_equal(w, d["weight"])
    # init_height_id makes it the 7-tuple (mapping_utils.py:536-539); grid_rgb may be absent (None)
    mapping_utils.save_3d_map(p, d["grid_feat"], d["grid_pos"], d["weight"], d["occupied_ids"], [4], None, init_height_id=7)
    out = mapping_utils.load_3d_map(p)
    assert len(out) == 7 and out[5] is None and int(out[6]) == 7 and out[0] == [4]
    # multi-floor twin: nine fields, cs comes back as a scalar
    q = tmp_path / "mf.h5df"
    mapping_utils.save_3d_map_multi_floor(q, d["grid_feat"], d["grid_pos"], d["weight"], d["grid_rgb"], d["occupied_ids"],
                                          {0, 2}, np.array([-1.0, 0.0, 2.5]), np.array([3.0, 2.0, 9.5]), 0.05)
    out = mapping_utils.load_3d_map_multi_floor(q)
    assert out[0] == [0, 2] and np.array_equal(out[6], [-1.0, 0.0, 2.5]) and out[8] == 0.05 and np.ndim(out[8]) == 0
    with pytest.raises(KeyError, match="pcd_min"):
        mapping_utils.load_3d_map_multi_floor(p)                       # a single-floor file is not a multi-floor map
    # area map file
    r = tmp_path / "area.h5df"
    mapping_utils.save_clip_sparse_map(r, d["grid_feat"], [np.eye(4), 2 * np.eye(4)])
    cm, poses = mapping_utils.load_clip_sparse_map(r)
    assert np.array_equal(cm, d["grid_feat"]) and poses.shape == (2, 4, 4)
    with pytest.raises(FileNotFoundError):
        mapping_utils.load_3d_map(tmp_path / "missing.h5df")


def test_load_3d_map_memory_maps_grid_feat_on_request(tmp_path, monkeypatch):
    monkeypatch.setattr(mapping_utils, "_have_h5py", lambda: False)
    rng = np.random.default_rng(4)
    d = _map_fields(rng)
    p = tmp_path / "vlmaps.h5df"
    mapping_utils.save_3d_map(p, d["grid_feat"], d["grid_pos"], d["weight"], d["occupied_ids"], [1, 2], d["grid_rgb"])
    it, gf, gp, w, occ, rgb = mapping_utils.load_3d_map(p, mmap_feat=True)
    assert isinstance(gf, np.memmap) and not gf.flags.writeable and gf.flags.c_contiguous
    assert np.array_equal(gf, d["grid_feat"]) and it == [1, 2] and np.array_equal(gp, d["grid_pos"])
    assert not isinstance(gp, np.memmap)
    # a legacy .npz twin has nothing to map: the copy comes back instead
    q = tmp_path / "old.h5df"
    np.savez(str(q) + ".npz", **d)
    gf2 = mapping_utils.load_3d_map(q, mmap_feat=True)[1]
    assert not isinstance(gf2, np.memmap) and np.array_equal(gf2, d["grid_feat"])


def test_legacy_npz_twin_still_loads(tmp_path):
    rng = np.random.default_rng(2)
    d = _map_fields(rng)
    p = tmp_path / "vlmaps.h5df"
    np.savez(str(p) + ".npz", **d)
    assert mapping_utils.map_file_exists(p)
    it, gf, *_ = mapping_utils.load_3d_map(p)
    assert it == list(range(11)) and np.array_equal(gf, d["grid_feat"])


def _patch_layout_to_chunked(path, name, arr, chunk, level=6, shuffle=True):
    """Re-point dataset `name` of an h5lite-written file at chunked + filtered storage appended to the file:
    rewrites its object header in place (continuation block at the end of the file), v1 chunk B-tree."""
    b = bytearray(Path(path).read_bytes())
    with h5lite.File(path) as f:
        hdr = None
        for n, a in f._links.items():
            if n == name:
                hdr = a
    rank = arr.ndim
    chunks = []
    grid = [range(0, s, c) for s, c in zip(arr.shape, chunk)]
    for offs in np.stack(np.meshgrid(*grid, indexing="ij"), -1).reshape(-1, rank):
        block = np.zeros(chunk, arr.dtype)
        sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunk, arr.shape))
        block[tuple(slice(0, x.stop - x.start) for x in sl)] = arr[sl]
        raw = block.tobytes()
        if shuffle:
            raw = np.frombuffer(raw, np.uint8).reshape(-1, arr.dtype.itemsize).T.tobytes()
        raw = zlib.compress(raw, level)
        chunks.append((tuple(int(o) for o in offs), raw))
    pos = (len(b) + 7) // 8 * 8
    b.extend(b"\0" * (pos - len(b)))
    entries = []
    for offs, raw in chunks:
        entries.append((offs, len(b), len(raw)))
        b.extend(raw + b"\0" * (-len(raw) % 8))
    btree_addr = len(b)
    node = b"TREE" + struct.pack("<BBHQQ", 1, 0, len(entries), h5lite.UNDEF, h5lite.UNDEF)
    for offs, addr, n in entries:
        node += struct.pack("<II", n, 0) + struct.pack(f"<{rank + 1}Q", *offs, 0) + struct.pack("<Q", addr)
    node += struct.pack("<II", 0, 0) + struct.pack(f"<{rank + 1}Q", *arr.shape, 0)
    b.extend(node)
    # new messages in a continuation block: layout v3 chunked + filter pipeline v1 (shuffle, deflate)
    layout = struct.pack("<BBB", 3, 2, rank + 1) + struct.pack("<Q", btree_addr) + struct.pack(f"<{rank + 1}I", *chunk, arr.dtype.itemsize)
    filt = struct.pack("<BB6x", 1, 2 if shuffle else 1)
    if shuffle:
        filt += struct.pack("<HHHH", 2, 0, 0, 1) + struct.pack("<I", arr.dtype.itemsize) + b"\0" * 4
    filt += struct.pack("<HHHH", 1, 0, 0, 1) + struct.pack("<I", level) + b"\0" * 4
    cont = h5lite._message_v1(h5lite.MSG_LAYOUT, layout) + h5lite._message_v1(h5lite.MSG_FILTERS, filt)
    cont_addr = len(b)
    b.extend(cont)
    # in the header: turn the old layout message into NIL and the spare NIL into the continuation message
    ver, _, nmsg, ref, hsize = struct.unpack_from("<BBHII", b, hdr)
    q = hdr + 16
    for _ in range(nmsg):
        mtype, msize = struct.unpack_from("<HH", b, q)
        if mtype == h5lite.MSG_LAYOUT:
            struct.pack_into("<H", b, q, h5lite.MSG_NIL)
        elif mtype == h5lite.MSG_NIL:
            struct.pack_into("<HH", b, q, h5lite.MSG_CONTINUATION, 16)
            struct.pack_into("<QQ", b, q + 8, cont_addr, len(cont))
        q += 8 + msize
    struct.pack_into("<H", b, hdr + 2, nmsg + 2)
    struct.pack_into("<Q", b, 40, len(b))                              # end-of-file address in the superblock
    Path(path).write_bytes(bytes(b))


@pytest.mark.parametrize("shuffle", [True, False])
def test_reads_chunked_deflate_shuffle(tmp_path, shuffle):
    rng = np.random.default_rng(3)
    arr = rng.integers(0, 50, (37, 21)).astype(np.int32)               # ragged against the 16 x 8 chunks
    p = tmp_path / "chunked.h5"
    h5lite.write_file(p, {"a": arr, "b": np.arange(3, dtype=np.float32)})
    _patch_layout_to_chunked(p, "a", arr, (16, 8), shuffle=shuffle)
    with h5lite.File(p) as f:
        ds = f["a"]
        assert ds._layout["class"] == 2 and ds._layout["chunk"] == (16, 8) and len(ds._filters) == (2 if shuffle else 1)
        assert ds.offset is None and np.array_equal(ds.read(), arr)
        assert np.array_equal(f["b"][:], np.arange(3, dtype=np.float32))
        with pytest.raises(h5lite.H5Error, match="not contiguous"):
            ds.memmap()


def test_damaged_and_unsupported_files_fail_loudly(tmp_path):
    p = tmp_path / "x.h5"
    p.write_bytes(b"not an hdf5 file at all" * 10)
    with pytest.raises(h5lite.H5Error, match="signature"):
        h5lite.File(p)
    good = tmp_path / "good.h5"
    h5lite.write_file(good, {"grid_feat": np.ones((100, 64), np.float32)})
    raw = good.read_bytes()
    p.write_bytes(raw[:5000])                                          # data region cut off
    with h5lite.File(p) as f:
        with pytest.raises(h5lite.H5Error, match="file ends|outside"):
            f["grid_feat"].read()
    p.write_bytes(raw[:200])                                           # metadata cut off
    with pytest.raises(h5lite.H5Error):
        h5lite.File(p)
    bad = bytearray(raw)
    bad[8] = 9                                                         # unknown superblock version
    p.write_bytes(bytes(bad))
    with pytest.raises(h5lite.H5Error, match="superblock version 9"):
        h5lite.File(p)
    with pytest.raises(h5lite.H5Error, match="read-only"):
        h5lite.File(good, "w")
    with pytest.raises(h5lite.H5Error, match="bool"):
        h5lite.write_file(p, {"mask": np.zeros(3, bool)})
    with pytest.raises(h5lite.H5Error, match="dtype"):
        h5lite.write_file(p, {"c": np.zeros(3, np.complex64)})
    with pytest.raises(h5lite.H5Error, match="name"):
        h5lite.write_file(p, {"a/b": np.zeros(3)})
    with pytest.raises(h5lite.H5Error, match="at most"):
        h5lite.write_file(p, {f"d{i}": np.zeros(1) for i in range(40)})


def test_superblock_v2_and_v2_object_headers_with_link_messages(tmp_path):
    """A new-style file (libver='latest' shape): superblock v2, OHDR v2 root group with compact link messages,
    OHDR v2 dataset with a compact layout.  Assembled by hand from the format specification."""
    arr = np.arange(6, dtype="<i2").reshape(2, 3)

    def ohdr2(msgs):
        body = b"".join(struct.pack("<BHB", t, len(m), 0) + m for t, m in msgs)
        head = b"OHDR" + struct.pack("<BB", 2, 0) + struct.pack("<B", len(body))
        return head + body + b"\0\0\0\0"                               # checksum (not verified by the reader)

    space = struct.pack("<BBBB", 2, 2, 0, 1) + struct.pack("<QQ", 2, 3)
    dtype = h5lite._datatype_message(arr.dtype)
    layout = struct.pack("<BBH", 3, 0, arr.nbytes) + arr.tobytes()
    ds = ohdr2([(1, space), (3, dtype), (8, layout)])
    big = np.arange(40, dtype="<f4")
    sb_size = 48
    ds_addr = sb_size
    # a second dataset with a layout message v4, contiguous (what libver >= v110 writes), raw data after the headers
    ds2_addr = ds_addr + len(ds)
    space2 = struct.pack("<BBBB", 2, 1, 0, 1) + struct.pack("<Q", 40)
    probe = ohdr2([(1, space2), (3, h5lite._datatype_message(big.dtype)), (8, struct.pack("<BBQQ", 4, 1, 0, big.nbytes))])
    link = struct.pack("<BB", 1, 0) + struct.pack("<B", 4) + b"tiny" + struct.pack("<Q", ds_addr)
    link2 = struct.pack("<BB", 1, 0) + struct.pack("<B", 3) + b"big" + struct.pack("<Q", ds2_addr)
    link_info = struct.pack("<BB", 0, 0) + struct.pack("<QQ", h5lite.UNDEF, h5lite.UNDEF)
    root = ohdr2([(2, link_info), (6, link), (6, link2)])
    root_addr = ds2_addr + len(probe)
    data_addr = root_addr + len(root)
    ds2 = ohdr2([(1, space2), (3, h5lite._datatype_message(big.dtype)), (8, struct.pack("<BBQQ", 4, 1, data_addr, big.nbytes))])
    assert len(ds2) == len(probe)
    eof = data_addr + big.nbytes
    sb = h5lite.SIGNATURE + struct.pack("<BBBB", 2, 8, 8, 0) + struct.pack("<QQQQ", 0, h5lite.UNDEF, eof, root_addr) + b"\0" * 4
    assert len(sb) == sb_size
    p = tmp_path / "v2.h5"
    p.write_bytes(sb + ds + ds2 + root + big.tobytes())
    with h5lite.File(p) as f:
        assert f.superblock_version == 2 and f.keys() == ["tiny", "big"]
        assert np.array_equal(f["tiny"][:], arr) and f["tiny"].dtype == np.dtype("<i2")
        assert np.array_equal(f["big"][:], big) and f["big"].offset == data_addr and np.array_equal(f["big"].memmap(), big)
    # a chunked layout v4 is named as unsupported instead of being misread
    bad = ohdr2([(1, space2), (3, h5lite._datatype_message(big.dtype)), (8, struct.pack("<BBQQ", 4, 2, data_addr, big.nbytes))])
    p.write_bytes(sb + ds + bad + root + big.tobytes())
    with h5lite.File(p) as f, pytest.raises(h5lite.H5Error, match="layout message v4"):
        f["big"]

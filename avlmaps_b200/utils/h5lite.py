"""Dependency-free reader / writer for the HDF5 files of the map path (`vlmaps.h5df`).

The reference stores a map with h5py (`save_3d_map` / `load_3d_map`, reference
avlmaps/utils/mapping_utils.py:469-541; multi-floor twin avlmaps/map/vlmap_builder_multi_floor.py:370-393;
`save_clip_sparse_map` mapping_utils.py:637-647): a handful of plain numeric datasets in the root group, created
with `f.create_dataset(name, data=array)` -- no chunking, no filters, no attributes.  This module reads that
family of files without h5py (and a good deal more, see below) and writes the same structure, so that a map made
by the reference loads here -- and a dataset's bytes can be handed to the device straight from the file
(`Dataset.offset` / `Dataset.memmap()`), which is what `load_map` needs for a 2 GB `grid_feat`.

Reader: superblock versions 0-3 (user block / base address honoured), object headers v1 and v2 (with
continuation blocks), groups as symbol tables (v1 B-tree + local heap + SNOD) or as compact link messages, nested
groups, dataspaces v1/v2 (scalar, simple), fixed-point and IEEE floating-point datatypes of either byte order,
layouts compact / contiguous / chunked (v1 B-tree index; deflate and shuffle filters).  Anything else (dense link
storage in fractal heaps, compound / variable-length types, virtual or external storage) raises `H5Error` naming the
feature instead of returning wrong bytes.

Writer: superblock v0, root group as a symbol table with ONE leaf node (its capacity is declared in the superblock's
"group leaf node K"), v1 object headers with dataspace v1, datatype v1, fill value v2 and a contiguous layout v3
message -- the structures libhdf5 itself emits under `libver="earliest"` (h5py's default).  Raw data is aligned to
4096 bytes.  Layout of every structure follows the published "HDF5 File Format Specification Version 3.0".
"""
from __future__ import annotations

import struct
import zlib
from pathlib import Path
from typing import Dict, Iterator, List, Optional, Tuple

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF

MSG_NIL, MSG_DATASPACE, MSG_LINK_INFO, MSG_DATATYPE, MSG_FILL_OLD, MSG_FILL, MSG_LINK = 0, 1, 2, 3, 4, 5, 6
MSG_LAYOUT, MSG_GROUP_INFO, MSG_FILTERS, MSG_ATTRIBUTE, MSG_CONTINUATION, MSG_SYMBOL_TABLE = 8, 10, 11, 12, 16, 17


class H5Error(RuntimeError):
    """The file is not HDF5, is damaged, or uses a feature this reader does not implement."""


# ------------------------------------------------------------------------------------------------ reader
class _Buf:
    """Random-access little-endian reads on a file object; addresses are relative to the base address."""

    def __init__(self, fh, base: int, size_of_offsets: int, size_of_lengths: int, file_size: int):
        self.fh, self.base, self.so, self.sl, self.file_size = fh, base, size_of_offsets, size_of_lengths, file_size

    def read(self, addr: int, n: int) -> bytes:
        pos = self.base + addr
        if addr == self.undef or pos < 0 or pos + n > self.file_size:
            raise H5Error(f"read of {n} bytes at address {addr:#x} is outside the file ({self.file_size} bytes)")
        self.fh.seek(pos)
        data = self.fh.read(n)
        if len(data) != n:
            raise H5Error(f"short read at address {addr:#x}")
        return data

    @property
    def undef(self) -> int:
        return (1 << (8 * self.so)) - 1

    def uint(self, data: bytes, pos: int, n: int) -> int:
        return int.from_bytes(data[pos:pos + n], "little")


class Dataset:
    """One dataset: shape, dtype and where its bytes are.  `read()` returns a fresh C-contiguous array."""

    def __init__(self, file: "File", name: str, shape: Tuple[int, ...], dtype: np.dtype, layout: dict, filters: list):
        self.file, self.name, self.shape, self.dtype, self._layout, self._filters = file, name, shape, dtype, layout, filters

    # --- h5py-like conveniences (what the reference's load functions use: f[k][:], f[k][()], "k" in f)
    def __getitem__(self, key):
        arr = self.read()
        if key == () or key is Ellipsis:
            return arr[()] if arr.ndim == 0 else arr
        return arr[key]

    def __array__(self, dtype=None, copy=None):
        arr = self.read()
        return arr.astype(dtype) if dtype is not None else arr

    @property
    def size(self) -> int:
        return int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1

    @property
    def nbytes(self) -> int:
        return self.size * self.dtype.itemsize

    @property
    def offset(self) -> Optional[int]:
        """Absolute file offset of the raw bytes of a contiguous dataset (None otherwise, or when never written)."""
        if self._layout["class"] != 1 or self._layout["address"] is None:
            return None
        return self.file._buf.base + self._layout["address"]

    def memmap(self) -> np.ndarray:
        """Read-only memory map of a contiguous dataset (no copy; pages stream in as the device upload reads them)."""
        off = self.offset
        if off is None or self.size == 0:
            raise H5Error(f"dataset {self.name!r} is not contiguous in the file; use read()")
        return np.memmap(self.file.path, dtype=self.dtype, mode="r", offset=off, shape=self.shape)

    def read(self) -> np.ndarray:
        buf, lay = self.file._buf, self._layout
        n = self.nbytes
        if lay["class"] == 0:                                   # compact: bytes live in the object header
            raw = lay["data"][:n]
            if len(raw) < n:
                raise H5Error(f"compact dataset {self.name!r} holds {len(raw)} bytes, needs {n}")
            return np.frombuffer(raw, dtype=self.dtype).reshape(self.shape).copy()
        if lay["class"] == 1:                                   # contiguous
            if lay["address"] is None or n == 0:                # storage never allocated -> fill value (zeros)
                return np.zeros(self.shape, dtype=self.dtype)
            if lay["size"] < n:
                raise H5Error(f"dataset {self.name!r}: layout holds {lay['size']} bytes, dataspace needs {n}")
            out = np.empty(self.shape, dtype=self.dtype)
            self.file._fh.seek(buf.base + lay["address"])
            got = self.file._fh.readinto(memoryview(out.reshape(-1).view(np.uint8)))
            if got != n:
                raise H5Error(f"dataset {self.name!r}: file ends after {got} of {n} bytes")
            return out
        if lay["class"] == 2:
            return self._read_chunked()
        raise H5Error(f"dataset {self.name!r}: layout class {lay['class']} (virtual / unknown) is not supported")

    # --- chunked storage, v1 B-tree index
    def _read_chunked(self) -> np.ndarray:
        lay, rank = self._layout, len(self.shape)
        chunk = lay["chunk"]
        if len(chunk) != rank:
            raise H5Error(f"dataset {self.name!r}: chunk rank {len(chunk)} does not match dataspace rank {rank}")
        out = np.zeros(self.shape, dtype=self.dtype)
        if lay["address"] is None:
            return out
        for offs, addr, nbytes, mask in self._chunks(lay["address"], rank):
            raw = self.file._buf.read(addr, nbytes)
            for i, (fid, cd) in reversed(list(enumerate(self._filters))):
                if mask >> i & 1:
                    continue
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:
                    width = cd[0] if cd else self.dtype.itemsize
                    a = np.frombuffer(raw, dtype=np.uint8)
                    m = a.size // width
                    raw = a[:m * width].reshape(width, m).T.tobytes() + a[m * width:].tobytes()
                elif fid == 3:                                  # fletcher32: 4 checksum bytes trail the chunk
                    raw = raw[:-4]
                else:
                    raise H5Error(f"dataset {self.name!r}: filter id {fid} is not supported")
            block = np.frombuffer(raw, dtype=self.dtype, count=int(np.prod(chunk))).reshape(chunk)
            dst = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunk, self.shape))
            src = tuple(slice(0, d.stop - d.start) for d in dst)
            out[dst] = block[src]
        return out

    def _chunks(self, addr: int, rank: int) -> Iterator[Tuple[Tuple[int, ...], int, int, int]]:
        buf = self.file._buf
        head = buf.read(addr, 8 + 2 * buf.so)
        if head[:4] != b"TREE" or head[4] != 1:
            raise H5Error(f"dataset {self.name!r}: chunk index at {addr:#x} is not a v1 raw-data B-tree")
        level, used = head[5], buf.uint(head, 6, 2)
        key_size = 8 + 8 * (rank + 1)
        body = buf.read(addr + len(head), used * (key_size + buf.so) + key_size)
        for i in range(used):
            p = i * (key_size + buf.so)
            nbytes, mask = struct.unpack_from("<II", body, p)
            offs = struct.unpack_from(f"<{rank}Q", body, p + 8)
            child = buf.uint(body, p + key_size, buf.so)
            if level:
                yield from self._chunks(child, rank)
            else:
                yield offs, child, nbytes, mask


class Group:
    def __init__(self, file: "File", name: str, links: Dict[str, int]):
        self.file, self.name, self._links = file, name, links

    def keys(self):
        return list(self._links)

    def __iter__(self):
        return iter(self._links)

    def __len__(self):
        return len(self._links)

    def __contains__(self, name: str) -> bool:
        try:
            self._resolve(name)
            return True
        except KeyError:
            return False

    def _resolve(self, name: str):
        node = self
        parts = [p for p in name.split("/") if p]
        if name.startswith("/"):
            node = self.file
        for i, part in enumerate(parts):
            if not isinstance(node, Group) or part not in node._links:
                raise KeyError(name)
            node = self.file._open(node._links[part], (node.name.rstrip("/") + "/" + part))
        return node

    def __getitem__(self, name: str):
        return self._resolve(name)


class File(Group):
    """`with h5lite.File(path) as f: f["grid_feat"][:]` -- the read-only subset of h5py.File the map path uses."""

    def __init__(self, path, mode: str = "r"):
        if mode != "r":
            raise H5Error("h5lite.File is read-only; write with h5lite.write_file")
        self.path = str(path)
        self._fh = open(self.path, "rb")
        try:
            self._cache: Dict[int, object] = {}
            root_addr = self._superblock()
            root = self._open(root_addr, "/")
            if not isinstance(root, Group):
                raise H5Error("root object is not a group")
            Group.__init__(self, self, "/", root._links)
        except Exception:
            self._fh.close()
            raise

    def close(self):
        self._fh.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # --- superblock
    def _superblock(self) -> int:
        fh = self._fh
        fh.seek(0, 2)
        file_size = fh.tell()
        pos = 0
        while True:                                             # user block: signature at 0, 512, 1024, 2048, ...
            if pos + 8 > file_size:
                raise H5Error(f"{self.path}: no HDF5 signature found")
            fh.seek(pos)
            if fh.read(8) == SIGNATURE:
                break
            pos = 512 if pos == 0 else pos * 2
        fh.seek(pos)
        sb = fh.read(min(128, file_size - pos))
        version = sb[8]
        if version in (0, 1):
            so, sl = sb[13], sb[14]
            p = 24 + (4 if version == 1 else 0)
            base = int.from_bytes(sb[p:p + so], "little")
            p += 4 * so                                         # base, free-space, end-of-file, driver-info addresses
            # root group symbol table entry: link name offset, object header address, cache type, reserved, scratch
            root_addr = int.from_bytes(sb[p + so:p + 2 * so], "little")
        elif version in (2, 3):
            so, sl = sb[9], sb[10]
            base = int.from_bytes(sb[12:12 + so], "little")
            root_addr = int.from_bytes(sb[12 + 3 * so:12 + 4 * so], "little")
        else:
            raise H5Error(f"{self.path}: superblock version {version} is not supported")
        if so not in (2, 4, 8) or sl not in (2, 4, 8):
            raise H5Error(f"{self.path}: size of offsets / lengths {so} / {sl} is not supported")
        # the base address is normally where the superblock sits (a user block shifts both)
        self._buf = _Buf(fh, base if base != (1 << 8 * so) - 1 else pos, so, sl, file_size)
        self.superblock_version = version
        return root_addr

    # --- object headers
    def _messages(self, addr: int) -> List[Tuple[int, int, bytes]]:
        """[(type, flags, body)] of the object header at `addr`, continuation blocks followed."""
        buf = self._buf
        first = buf.read(addr, 16)
        msgs: List[Tuple[int, int, bytes]] = []
        if first[:4] == b"OHDR":
            if first[4] != 2:
                raise H5Error(f"object header v{first[4]} at {addr:#x} is not supported")
            flags = first[5]
            p = 6 + (16 if flags & 0x20 else 0) + (4 if flags & 0x10 else 0)
            nsz = 1 << (flags & 3)
            head = buf.read(addr, p + nsz)
            size0 = buf.uint(head, p, nsz)
            order = 2 if flags & 0x04 else 0
            blocks = [(addr + p + nsz, size0)]
            while blocks:
                baddr, bsize = blocks.pop(0)
                data = buf.read(baddr, bsize)
                q = 0
                while q + 4 + order <= len(data):
                    mtype, msize, mflags = data[q], buf.uint(data, q + 1, 2), data[q + 3]
                    q += 4 + order
                    body = data[q:q + msize]
                    q += msize
                    if mtype == MSG_CONTINUATION:
                        caddr, clen = buf.uint(body, 0, buf.so), buf.uint(body, buf.so, buf.sl)
                        if buf.read(caddr, 4) != b"OCHK":
                            raise H5Error(f"continuation block at {caddr:#x} lacks the OCHK signature")
                        blocks.append((caddr + 4, clen - 8))    # minus signature and trailing checksum
                    elif mtype != MSG_NIL:
                        msgs.append((mtype, mflags, body))
            return msgs
        if first[0] != 1:
            raise H5Error(f"no object header at address {addr:#x}")
        nmsgs, hsize = buf.uint(first, 2, 2), buf.uint(first, 8, 4)
        blocks = [(addr + 16, hsize)]
        seen = 0
        while blocks and seen < nmsgs:
            baddr, bsize = blocks.pop(0)
            data = buf.read(baddr, bsize)
            q = 0
            while q + 8 <= len(data) and seen < nmsgs:
                mtype, msize, mflags = buf.uint(data, q, 2), buf.uint(data, q + 2, 2), data[q + 4]
                body = data[q + 8:q + 8 + msize]
                q += 8 + msize
                seen += 1
                if mtype == MSG_CONTINUATION:
                    blocks.append((buf.uint(body, 0, buf.so), buf.uint(body, buf.so, buf.sl)))
                elif mtype != MSG_NIL:
                    msgs.append((mtype, mflags, body))
        return msgs

    def _open(self, addr: int, name: str):
        if addr in self._cache:
            return self._cache[addr]
        msgs = self._messages(addr)
        types = {t for t, _, _ in msgs}
        if MSG_LAYOUT in types and MSG_DATATYPE in types and MSG_DATASPACE in types:
            obj = self._dataset(name, msgs)
        elif MSG_SYMBOL_TABLE in types or MSG_LINK_INFO in types or MSG_LINK in types or MSG_GROUP_INFO in types:
            obj = Group(self, name, self._links_of(msgs))
        elif not msgs:
            obj = Group(self, name, {})
        else:
            raise H5Error(f"object {name!r} is neither a dataset nor a group (a committed datatype?)")
        self._cache[addr] = obj
        return obj

    # --- groups
    def _links_of(self, msgs) -> Dict[str, int]:
        buf = self._buf
        links: Dict[str, int] = {}
        for mtype, mflags, body in msgs:
            if mflags & 0x02:
                raise H5Error("shared object header messages are not supported")
            if mtype == MSG_SYMBOL_TABLE:
                btree, heap = buf.uint(body, 0, buf.so), buf.uint(body, buf.so, buf.so)
                hh = buf.read(heap, 8 + 2 * buf.sl + buf.so)
                if hh[:4] != b"HEAP":
                    raise H5Error(f"no local heap at {heap:#x}")
                seg_size, seg_addr = buf.uint(hh, 8, buf.sl), buf.uint(hh, 8 + 2 * buf.sl, buf.so)
                names = buf.read(seg_addr, seg_size)
                for name_off, ohdr in self._group_entries(btree):
                    end = names.index(b"\0", name_off)
                    links[names[name_off:end].decode("utf-8")] = ohdr
            elif mtype == MSG_LINK:
                version, flags = body[0], body[1]
                if version != 1:
                    raise H5Error(f"link message v{version} is not supported")
                p = 2
                ltype = 0
                if flags & 0x08:
                    ltype = body[p]
                    p += 1
                if flags & 0x04:
                    p += 8
                if flags & 0x10:
                    p += 1
                nlen_size = 1 << (flags & 3)
                nlen = buf.uint(body, p, nlen_size)
                p += nlen_size
                lname = body[p:p + nlen].decode("utf-8")
                p += nlen
                if ltype == 0:                                  # hard link; soft / external links are skipped
                    links[lname] = buf.uint(body, p, buf.so)
            elif mtype == MSG_LINK_INFO:
                flags = body[1]
                p = 2 + (8 if flags & 1 else 0)
                fractal = buf.uint(body, p, buf.so)
                if fractal != buf.undef:
                    raise H5Error("group with dense link storage (fractal heap) is not supported")
        return links

    def _group_entries(self, addr: int) -> Iterator[Tuple[int, int]]:
        buf = self._buf
        head = buf.read(addr, 8)
        if head[:4] == b"SNOD":
            n = buf.uint(head, 6, 2)
            esize = 2 * buf.so + 24
            data = buf.read(addr + 8, n * esize)
            for i in range(n):
                yield buf.uint(data, i * esize, buf.so), buf.uint(data, i * esize + buf.so, buf.so)
            return
        if head[:4] != b"TREE" or head[4] != 0:
            raise H5Error(f"no group B-tree node at {addr:#x}")
        used = buf.uint(head, 6, 2)
        body = buf.read(addr + 8 + 2 * buf.so, used * (buf.sl + buf.so) + buf.sl)
        for i in range(used):                                   # key0 child0 key1 child1 ... keyN
            child = buf.uint(body, i * (buf.sl + buf.so) + buf.sl, buf.so)
            yield from self._group_entries(child)

    # --- datasets
    def _dataset(self, name: str, msgs) -> Dataset:
        buf = self._buf
        shape = dtype = layout = None
        filters: list = []
        for mtype, mflags, body in msgs:
            if mtype in (MSG_DATASPACE, MSG_DATATYPE, MSG_LAYOUT, MSG_FILTERS) and mflags & 0x02:
                raise H5Error(f"dataset {name!r}: shared object header messages are not supported")
            if mtype == MSG_DATASPACE:
                version, rank = body[0], body[1]
                if version == 1:
                    p = 8
                elif version == 2:
                    p = 4
                    if body[3] == 2:
                        raise H5Error(f"dataset {name!r} has a null dataspace")
                else:
                    raise H5Error(f"dataset {name!r}: dataspace v{version} is not supported")
                shape = tuple(buf.uint(body, p + i * buf.sl, buf.sl) for i in range(rank))
            elif mtype == MSG_DATATYPE:
                dtype = _numpy_dtype(body, name)
            elif mtype == MSG_LAYOUT:
                layout = self._layout(body, name)
            elif mtype == MSG_FILTERS:
                filters = _filters(body, name)
        if layout["class"] != 2 and filters:
            raise H5Error(f"dataset {name!r}: filters on a non-chunked layout")
        if layout["class"] == 2:
            layout["chunk"] = layout["chunk"][:len(shape)]      # the trailing entry is the element size
        return Dataset(self, name, shape, dtype, layout, filters)

    def _layout(self, body: bytes, name: str) -> dict:
        buf = self._buf
        version = body[0]
        if version in (1, 2):
            rank, cls = body[1], body[2]
            p = 8
            addr = None
            if cls != 0:
                addr = buf.uint(body, p, buf.so)
                p += buf.so
            dims = struct.unpack_from(f"<{rank}I", body, p)
            p += 4 * rank
            if cls == 2:
                p += 4                                          # element size
                return {"class": 2, "address": None if addr == buf.undef else addr, "chunk": tuple(dims)}
            if cls == 1:
                return {"class": 1, "address": None if addr == buf.undef else addr, "size": UNDEF}
            size = buf.uint(body, p, 4)
            return {"class": 0, "data": body[p + 4:p + 4 + size]}
        if version in (3, 4):                                   # v4 (libver >= v110) encodes compact / contiguous like v3
            cls = body[1]
            if cls == 0:
                size = buf.uint(body, 2, 2)
                return {"class": 0, "data": body[4:4 + size]}
            if cls == 1:
                addr, size = buf.uint(body, 2, buf.so), buf.uint(body, 2 + buf.so, buf.sl)
                return {"class": 1, "address": None if addr == buf.undef else addr, "size": size}
            if cls == 2 and version == 4:
                raise H5Error(f"dataset {name!r}: chunked layout message v4 (single-chunk / array / B-tree v2 indexes, "
                              "libver >= v110) is not supported; contiguous and v1-B-tree chunked datasets are")
            if cls == 2:
                rank = body[2]
                addr = buf.uint(body, 3, buf.so)
                dims = struct.unpack_from(f"<{rank}I", body, 3 + buf.so)
                return {"class": 2, "address": None if addr == buf.undef else addr, "chunk": tuple(dims[:-1]) + (dims[-1],)}
            return {"class": cls}
        raise H5Error(f"dataset {name!r}: data layout message v{version} is not supported "
                      "(written with libver='latest'? this reader handles the v1 B-tree chunk index only)")


def _numpy_dtype(body: bytes, name: str) -> np.dtype:
    cls, version = body[0] & 0x0F, body[0] >> 4
    bits0, bits1 = body[1], body[2]
    size = int.from_bytes(body[4:8], "little")
    if version not in (1, 2, 3):
        raise H5Error(f"dataset {name!r}: datatype message v{version} is not supported")
    if cls == 0:                                                # fixed point: bit 0 byte order, bit 3 signed
        order = ">" if bits0 & 1 else "<"
        offset, precision = struct.unpack_from("<HH", body, 8)
        if size not in (1, 2, 4, 8) or offset != 0 or precision != 8 * size:
            raise H5Error(f"dataset {name!r}: {precision}-bit integer in {size} bytes is not supported")
        return np.dtype(f"{order}{'i' if bits0 & 0x08 else 'u'}{size}")
    if cls == 1:                                                # floating point: IEEE binary16/32/64 only
        if bits0 & 0x40:
            raise H5Error(f"dataset {name!r}: VAX-endian floats are not supported")
        order = ">" if bits0 & 1 else "<"
        offset, precision, eloc, esize, mloc, msize = struct.unpack_from("<HHBBBB", body, 8)
        bias = int.from_bytes(body[16:20], "little")
        ieee = {2: (16, 10, 5, 0, 10, 15, 15), 4: (32, 23, 8, 0, 23, 127, 31), 8: (64, 52, 11, 0, 52, 1023, 63)}.get(size)
        if ieee is None or (precision, eloc, esize, mloc, msize, bias, bits1) != ieee or offset != 0:
            raise H5Error(f"dataset {name!r}: non-IEEE floating-point type of {size} bytes is not supported")
        return np.dtype(f"{order}f{size}")
    kinds = {2: "time", 3: "string", 4: "bitfield", 5: "opaque", 6: "compound", 7: "reference", 8: "enum",
             9: "variable-length", 10: "array"}
    raise H5Error(f"dataset {name!r}: datatype class {kinds.get(cls, cls)} is not supported")


def _filters(body: bytes, name: str) -> list:
    version, n = body[0], body[1]
    out = []
    p = 8 if version == 1 else 2
    for _ in range(n):
        fid = int.from_bytes(body[p:p + 2], "little")
        p += 2
        nlen = 0
        if version == 1 or fid >= 256:
            nlen = int.from_bytes(body[p:p + 2], "little")
            p += 2
        p += 2                                                  # flags
        ncd = int.from_bytes(body[p:p + 2], "little")
        p += 2 + nlen
        cd = struct.unpack_from(f"<{ncd}I", body, p)
        p += 4 * ncd
        if version == 1 and ncd % 2:
            p += 4
        out.append((fid, cd))
    return out


# ------------------------------------------------------------------------------------------------ writer
def _align(n: int, a: int) -> int:
    return (n + a - 1) // a * a


def _datatype_message(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    big = 1 if dt.byteorder == ">" else 0
    if dt.kind in "iu":
        bits0 = big | (0x08 if dt.kind == "i" else 0)
        return struct.pack("<BBBBIHH", 0x10, bits0, 0, 0, dt.itemsize, 0, 8 * dt.itemsize)
    if dt.kind == "f" and dt.itemsize in (2, 4, 8):
        precision, eloc, esize, mloc, msize, bias, sign = {
            2: (16, 10, 5, 0, 10, 15, 15), 4: (32, 23, 8, 0, 23, 127, 31), 8: (64, 52, 11, 0, 52, 1023, 63)}[dt.itemsize]
        # bits 4-5 = 2: the most significant mantissa bit is implied (IEEE normalisation); byte 1 = sign position
        return struct.pack("<BBBBIHHBBBBI", 0x11, big | 0x20, sign, 0, dt.itemsize, 0, precision, eloc, esize, mloc, msize, bias)
    raise H5Error(f"cannot store dtype {dt} (only integers and IEEE floats; cast bool masks to uint8 first)")


def _message_v1(mtype: int, body: bytes, flags: int = 0) -> bytes:
    body = body + b"\0" * (_align(len(body), 8) - len(body))
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _object_header_v1(messages: List[bytes]) -> bytes:
    blob = b"".join(messages)
    # version, reserved, number of messages, reference count, header size, 4 bytes to reach 8-byte alignment
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(blob)) + blob


LEAF_K = 16          # one symbol table node holds up to 2 * LEAF_K entries; declared in the superblock
INTERNAL_K = 16
DATA_ALIGN = 4096


def write_file(path, datasets: Dict[str, np.ndarray]) -> None:
    """Write `datasets` (name -> array) as contiguous datasets of the root group of a new HDF5 file."""
    if len(datasets) > 2 * LEAF_K:
        raise H5Error(f"at most {2 * LEAF_K} datasets per file")
    arrays = {}
    for name, arr in datasets.items():
        if not name or "/" in name or "\0" in name:
            raise H5Error(f"bad dataset name {name!r}")
        a = np.asarray(arr)
        if a.dtype == np.bool_:
            raise H5Error(f"dataset {name!r}: bool needs h5py's enum type; cast to uint8")
        arrays[name] = np.ascontiguousarray(a) if a.ndim else a
    names = sorted(arrays, key=lambda s: s.encode("utf-8"))     # symbol table entries are ordered by name bytes

    # local heap data segment: offset 0 holds the empty string (8 zero bytes), names are NUL-terminated and 8-aligned
    seg = bytearray(8)
    name_off = {}
    for n in names:
        name_off[n] = len(seg)
        raw = n.encode("utf-8") + b"\0"
        seg += raw + b"\0" * (_align(len(raw), 8) - len(raw))
    free_off = len(seg)
    seg += struct.pack("<QQ", 1, 32) + b"\0" * 16               # one free block (next = 1: end of list), 32 bytes
    seg_size = len(seg)

    sb_size = 8 + 8 + 4 + 4 + 4 * 8 + 40                        # 96 bytes
    # every header carries a 24-byte NIL message: room for the continuation message libhdf5 needs if someone
    # later opens the file read-write and adds an attribute (libhdf5's own headers keep such slack too)
    spare = _message_v1(MSG_NIL, b"\0" * 16)
    root_ohdr = _object_header_v1([_message_v1(MSG_SYMBOL_TABLE, b"\0" * 16), spare])
    addr_root = sb_size
    addr_btree = _align(addr_root + len(root_ohdr), 8)
    btree_size = 8 + 2 * 8 + 2 * INTERNAL_K * 8 + (2 * INTERNAL_K + 1) * 8
    addr_heap = addr_btree + btree_size
    addr_seg = addr_heap + 32
    addr_snod = _align(addr_seg + seg_size, 8)
    snod_size = 8 + 2 * LEAF_K * 40
    pos = addr_snod + snod_size

    # dataset object headers (their layout message needs the data address: headers first, then the data region)
    hdr_addr, hdr_len = {}, {}
    for n in names:
        hdr_addr[n] = pos
        a = arrays[n]
        body_len = 16 + sum(8 + _align(l, 8) for l in (8 + 8 * a.ndim, len(_datatype_message(a.dtype)), 8, 18, 16))
        hdr_len[n] = body_len
        pos = _align(pos + body_len, 8)
    data_addr = {}
    for n in names:
        a = arrays[n]
        if a.nbytes == 0:
            data_addr[n] = UNDEF
            continue
        pos = _align(pos, DATA_ALIGN if a.nbytes >= DATA_ALIGN else 8)
        data_addr[n] = pos
        pos += a.nbytes
    eof = pos

    def dataset_header(n: str) -> bytes:
        a = arrays[n]
        space = struct.pack("<BBBB4x", 1, a.ndim, 0, 0) + b"".join(struct.pack("<Q", d) for d in a.shape)
        fill = struct.pack("<BBBBI", 2, 2, 2, 1, 0)             # v2, allocate late, write if set, default value
        layout = struct.pack("<BBQQ", 3, 1, data_addr[n], a.nbytes)
        h = _object_header_v1([_message_v1(MSG_DATASPACE, space), _message_v1(MSG_DATATYPE, _datatype_message(a.dtype), 1),
                               _message_v1(MSG_FILL, fill, 1), _message_v1(MSG_LAYOUT, layout), spare])
        assert len(h) == hdr_len[n], (len(h), hdr_len[n])
        return h

    superblock = (SIGNATURE + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", LEAF_K, INTERNAL_K, 0)
                  + struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
                  + struct.pack("<QQII", 0, addr_root, 1, 0) + struct.pack("<QQ", addr_btree, addr_heap))
    assert len(superblock) == sb_size
    root_ohdr = _object_header_v1([_message_v1(MSG_SYMBOL_TABLE, struct.pack("<QQ", addr_btree, addr_heap), 1), spare])
    last = name_off[names[-1]] if names else 0
    btree = (b"TREE" + struct.pack("<BBH", 0, 0, 1 if names else 0) + struct.pack("<QQ", UNDEF, UNDEF)
             + (struct.pack("<QQQ", 0, addr_snod, last) if names else b""))
    btree += b"\0" * (btree_size - len(btree))
    heap = b"HEAP" + struct.pack("<B3xQQQ", 0, seg_size, free_off, addr_seg)
    snod = b"SNOD" + struct.pack("<BBH", 1, 0, len(names))
    for n in names:
        snod += struct.pack("<QQII16x", name_off[n], hdr_addr[n], 0, 0)
    snod += b"\0" * (snod_size - len(snod))

    tmp = Path(str(path) + ".tmp")
    with open(tmp, "wb") as f:
        for addr, blob in ((0, superblock), (addr_root, root_ohdr), (addr_btree, btree), (addr_heap, heap),
                           (addr_seg, bytes(seg)), (addr_snod, snod)):
            f.seek(addr)
            f.write(blob)
        for n in names:
            f.seek(hdr_addr[n])
            f.write(dataset_header(n))
        for n in names:
            if data_addr[n] != UNDEF:
                f.seek(data_addr[n])
                arrays[n].reshape(-1).view(np.uint8).tofile(f) if arrays[n].ndim else f.write(arrays[n].tobytes())
        f.truncate(eof)
    tmp.replace(path)


def read_file(path, names: Optional[List[str]] = None) -> Dict[str, np.ndarray]:
    """All (or the named) root-level datasets of `path` as arrays."""
    with File(path) as f:
        return {k: f[k].read() for k in (names if names is not None else f.keys()) if k in f and isinstance(f[k], Dataset)}

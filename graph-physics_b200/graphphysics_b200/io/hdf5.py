"""Minimal HDF5 reader / writer in pure Python + numpy (no h5py), for the files the reference's data path uses
(SURVEY §8f N4): the trajectory files of `H5Dataset` (graphphysics/dataset/h5_dataset.py, graphphysics/utils/
hierarchical.py:10-86: `file[trajectory][feature][()]`), the heavy-data files behind `XDMFDataset`
(graphphysics/dataset/xdmf_dataset.py; meshio's `xdmf` writer: `/data0`, `/data1`, ...) and the prediction archives
(graphphysics/utils/meshio_mesh.py:119-233).

Reader: superblock version 0 / 1 (what h5py and meshio write by default), version-1 object headers with
continuation blocks, symbol-table groups (B-tree v1 + local heap), datasets with contiguous, compact or chunked layout
(chunk B-tree v1), the deflate and shuffle filters, little- or big-endian fixed-point and IEEE floating-point types of
1 / 2 / 4 / 8 bytes.  Anything else (new-style groups, variable-length types, external storage) raises
`NotImplementedError` naming the feature.

Writer: the same on-disk structures (superblock version 0, one full-size symbol-table node and B-tree node per group,
version-1 object headers, contiguous little-endian datasets).  Validated by round trips through the reader above, which
itself is validated on h5py / meshio-written files; libhdf5 is not in this image, so reading these files back with h5py
is untested.  Used for the prediction archives and by the tests.

    f = H5File("train.h5");  f.keys();  f["0"]["velocity"][()]  ->  numpy array        (read-only mapping)
    write_h5("out.h5", {"data0": points, "traj": {"velocity": v}})
"""
from __future__ import annotations

import struct
import zlib
from typing import Dict, Iterator, List, Optional, Tuple, Union

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Dataset:
    """A dataset: `shape`, `dtype`, and `[()]` / `read()` for the whole array (the access pattern of the reference,
    hierarchical.py:80-84)."""

    def __init__(self, f: "H5File", name: str, shape: Tuple[int, ...], dtype: np.dtype, layout: dict, filters: List[Tuple[int, tuple]]):
        self._f, self.name, self.shape, self.dtype, self._layout, self._filters = f, name, shape, dtype, layout, filters

    @property
    def ndim(self) -> int:
        return len(self.shape)

    def __len__(self) -> int:
        return self.shape[0] if self.shape else 0

    def read(self) -> np.ndarray:
        buf, lay = self._f._buf, self._layout
        n = int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1
        if lay["class"] == 0:                                   # compact: the data sits in the object header
            raw = lay["data"]
            return np.frombuffer(raw, dtype=self.dtype, count=n).reshape(self.shape).copy()
        if lay["class"] == 1:                                   # contiguous
            if lay["addr"] == _UNDEF:                           # never written: fill value 0
                return np.zeros(self.shape, self.dtype)
            return np.frombuffer(buf, dtype=self.dtype, count=n, offset=lay["addr"]).reshape(self.shape).copy()
        # chunked: walk the chunk B-tree, inflate every chunk, paste it at its offset
        cdims = lay["chunk"]
        out = np.zeros(self.shape, self.dtype)
        if lay["addr"] == _UNDEF:
            return out
        for offs, addr, size, mask in self._f._chunks(lay["addr"], len(cdims)):
            raw = bytes(buf[addr:addr + size])
            for k, (fid, cd) in reversed(list(enumerate(self._filters))):
                if mask & (1 << k):
                    continue                                    # this filter was skipped for this chunk
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:                                  # shuffle: bytes of every element are stored plane by plane
                    es = self.dtype.itemsize
                    raw = np.frombuffer(raw, np.uint8).reshape(es, -1).T.tobytes()
                elif fid == 3:                                  # fletcher32: checksum appended, data unchanged
                    raw = raw[:-4]
                else:
                    raise NotImplementedError(f"HDF5 filter id {fid} (dataset {self.name})")
            chunk = np.frombuffer(raw, dtype=self.dtype, count=int(np.prod(cdims))).reshape(cdims)
            sel_out = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, self.shape))
            sel_in = tuple(slice(0, s.stop - s.start) for s in sel_out)
            out[sel_out] = chunk[sel_in]
        return out

    def __getitem__(self, key):
        a = self.read()
        return a if key == () or key is Ellipsis else a[key]

    def __array__(self, dtype=None):
        a = self.read()
        return a.astype(dtype) if dtype is not None else a


class H5Group:
    def __init__(self, f: "H5File", name: str, entries: Dict[str, int]):
        self._f, self.name, self._entries = f, name, entries

    def keys(self) -> List[str]:
        return list(self._entries)

    def __iter__(self) -> Iterator[str]:
        return iter(self._entries)

    def __len__(self) -> int:
        return len(self._entries)

    def __contains__(self, key: str) -> bool:
        return key.strip("/").split("/")[0] in self._entries

    def __getitem__(self, key: str) -> Union["H5Group", H5Dataset]:
        parts = [p for p in key.split("/") if p]
        node: Union[H5Group, H5Dataset] = self
        for p in parts:
            if not isinstance(node, H5Group):
                raise KeyError(key)
            if p not in node._entries:
                raise KeyError(f"{key!r} (no member {p!r} in {node.name!r})")
            node = node._f._object(node._entries[p], (node.name.rstrip("/") + "/" + p))
        return node

    def items(self):
        return [(k, self[k]) for k in self._entries]


class H5File(H5Group):
    """Read-only view of an HDF5 file (context manager; the whole file is memory-mapped)."""

    def __init__(self, path: str, mode: str = "r"):
        if mode != "r":
            raise ValueError("H5File is read-only; use write_h5() to create files")
        self.path = path
        self._buf = np.memmap(path, dtype=np.uint8, mode="r")
        b = self._buf
        if bytes(b[:8]) != _SIG:
            raise ValueError(f"{path}: not an HDF5 file")
        ver = int(b[8])
        if ver not in (0, 1):
            raise NotImplementedError(f"{path}: HDF5 superblock version {ver} (files written with libver='latest'); versions 0 and 1 are supported")
        if int(b[13]) != 8 or int(b[14]) != 8:
            raise NotImplementedError(f"{path}: size of offsets / lengths {int(b[13])} / {int(b[14])} (8 / 8 supported)")
        off = 24 if ver == 0 else 28                             # v1 adds indexed-storage K + reserved
        self._base = self._u64(off)
        root_entry = off + 32                                    # base, free-space, end-of-file, driver-info addresses
        root_ohdr = self._u64(root_entry + 8)
        cache_type = self._u32(root_entry + 16)
        if cache_type == 1:
            entries = self._group_entries(self._u64(root_entry + 24), self._u64(root_entry + 32))
        else:
            entries = self._object(root_ohdr, "/")._entries
        super().__init__(self, "/", entries)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def close(self):
        self._buf = None

    # ---------------------------------------------------------------- primitives
    def _u16(self, o): return int(struct.unpack_from("<H", self._buf, o)[0])
    def _u32(self, o): return int(struct.unpack_from("<I", self._buf, o)[0])
    def _u64(self, o): return int(struct.unpack_from("<Q", self._buf, o)[0])

    def _heap_string(self, heap_data: int, off: int) -> str:
        b = self._buf
        end = heap_data + off
        while b[end] != 0:
            end += 1
        return bytes(b[heap_data + off:end]).decode("utf-8")

    def _group_entries(self, btree: int, heap: int) -> Dict[str, int]:
        """Symbol-table group: local heap (names) + B-tree v1 of symbol-table nodes -> {name: object header address}."""
        b = self._buf
        if bytes(b[heap:heap + 4]) != b"HEAP":
            raise ValueError("HDF5: bad local heap signature")
        heap_data = self._u64(heap + 24)
        out: Dict[str, int] = {}

        def walk(addr: int):
            sig = bytes(b[addr:addr + 4])
            if sig == b"TREE":
                level, used = int(b[addr + 5]), self._u16(addr + 6)
                p = addr + 24                                   # keys (8) and children (8) interleaved, starting with key 0
                for i in range(used):
                    child = self._u64(p + 8 + 16 * i)
                    walk(child)
            elif sig == b"SNOD":
                n = self._u16(addr + 6)
                for i in range(n):
                    e = addr + 8 + 40 * i
                    out[self._heap_string(heap_data, self._u64(e))] = self._u64(e + 8)
            else:
                raise ValueError(f"HDF5: unexpected node signature {sig!r} in a group B-tree")

        walk(btree)
        return out

    def _messages(self, addr: int) -> List[Tuple[int, int, int]]:
        """(type, offset, size) of every message of a version-1 object header, continuation blocks included."""
        b = self._buf
        if bytes(b[addr:addr + 4]) == b"OHDR":
            raise NotImplementedError("HDF5 version-2 object headers (files written with libver='latest')")
        if int(b[addr]) != 1:
            raise ValueError(f"HDF5: object header version {int(b[addr])} at {addr}")
        n_msgs, hdr_size = self._u16(addr + 2), self._u32(addr + 8)
        blocks = [(addr + 16, hdr_size)]
        msgs = []
        while blocks and len(msgs) < n_msgs:
            p, size = blocks.pop(0)
            end = p + size
            while p + 8 <= end and len(msgs) < n_msgs:
                mtype, msize = self._u16(p), self._u16(p + 2)
                body = p + 8
                if mtype == 0x10:                               # continuation
                    blocks.append((self._u64(body), self._u64(body + 8)))
                msgs.append((mtype, body, msize))
                p = body + msize
        return msgs

    def _object(self, addr: int, name: str):
        msgs = self._messages(addr)
        by_type = {}
        for t, o, s in msgs:
            by_type.setdefault(t, (o, s))
        if 0x11 in by_type:                                      # symbol table message: a group
            o, _ = by_type[0x11]
            return H5Group(self, name, self._group_entries(self._u64(o), self._u64(o + 8)))
        if 0x08 not in by_type:
            if 0x02 in by_type or 0x06 in by_type:
                raise NotImplementedError(f"{name}: new-style (link-message) groups; write the file with the default libver")
            raise ValueError(f"{name}: object is neither a group nor a dataset")
        shape = self._dataspace(*by_type[0x01])
        dtype = self._datatype(by_type[0x03][0], name)
        layout = self._layout(by_type[0x08][0], name)
        filters = self._filters(*by_type[0x0B]) if 0x0B in by_type else []
        if layout["class"] == 2:
            layout["chunk"] = layout["chunk"][:len(shape)]
        return H5Dataset(self, name, shape, dtype, layout, filters)

    def _dataspace(self, o: int, size: int) -> Tuple[int, ...]:
        b = self._buf
        ver, rank = int(b[o]), int(b[o + 1])
        p = o + (8 if ver == 1 else 4)
        return tuple(self._u64(p + 8 * i) for i in range(rank))

    def _datatype(self, o: int, name: str) -> np.dtype:
        b = self._buf
        cls, bits0 = int(b[o]) & 0x0F, int(b[o + 1])
        size = self._u32(o + 4)
        order = ">" if (bits0 & 1) else "<"
        if cls == 0:                                             # fixed point
            signed = bool(bits0 & 0x08)
            if size not in (1, 2, 4, 8):
                raise NotImplementedError(f"{name}: {size}-byte integers")
            return np.dtype(f"{order}{'i' if signed else 'u'}{size}")
        if cls == 1:                                             # floating point
            if size not in (2, 4, 8):
                raise NotImplementedError(f"{name}: {size}-byte floats")
            return np.dtype(f"{order}f{size}")
        raise NotImplementedError(f"{name}: HDF5 datatype class {cls} (only integers and floats are supported)")

    def _layout(self, o: int, name: str) -> dict:
        b = self._buf
        ver, cls = int(b[o]), int(b[o + 1])
        if ver != 3:
            raise NotImplementedError(f"{name}: data layout message version {ver} (3 supported)")
        if cls == 0:
            size = self._u16(o + 2)
            return {"class": 0, "data": bytes(b[o + 4:o + 4 + size])}
        if cls == 1:
            return {"class": 1, "addr": self._u64(o + 2), "size": self._u64(o + 10)}
        if cls == 2:
            nd = int(b[o + 2])
            addr = self._u64(o + 3)
            dims = tuple(self._u32(o + 11 + 4 * i) for i in range(nd))      # last entry = element size
            return {"class": 2, "addr": addr, "chunk": dims[:-1], "ndims": nd}
        raise NotImplementedError(f"{name}: data layout class {cls}")

    def _filters(self, o: int, size: int) -> List[Tuple[int, tuple]]:
        b = self._buf
        ver, n = int(b[o]), int(b[o + 1])
        p = o + (8 if ver == 1 else 2)
        out = []
        for _ in range(n):
            fid = self._u16(p)
            if ver == 1 or fid >= 256:
                nlen = self._u16(p + 2)
                flags, ncd = self._u16(p + 4), self._u16(p + 6)
                p += 8 + (((nlen + 7) // 8) * 8 if ver == 1 else nlen)
            else:
                flags, ncd = self._u16(p + 2), self._u16(p + 4)
                p += 6
            cd = tuple(self._u32(p + 4 * i) for i in range(ncd))
            p += 4 * ncd
            if ver == 1 and ncd % 2:
                p += 4
            out.append((fid, cd))
        return out

    def _chunks(self, addr: int, rank: int):
        """Leaves of a chunk B-tree v1: (offsets, address, stored size, filter mask)."""
        b = self._buf
        if bytes(b[addr:addr + 4]) != b"TREE" or int(b[addr + 4]) != 1:
            raise ValueError("HDF5: bad chunk B-tree node")
        level, used = int(b[addr + 5]), self._u16(addr + 6)
        key_size = 8 + 8 * (rank + 1)
        p = addr + 24
        for i in range(used):
            k = p + i * (key_size + 8)
            size, mask = self._u32(k), self._u32(k + 4)
            offs = tuple(self._u64(k + 8 + 8 * d) for d in range(rank))
            child = self._u64(k + key_size)
            if level == 0:
                yield offs, child, size, mask
            else:
                yield from self._chunks(child, rank)


# ------------------------------------------------------------------------------------------------ writer
def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


def _dtype_msg(dt: np.dtype) -> bytes:
    dt = np.dtype(dt).newbyteorder("<")
    size = dt.itemsize
    if dt.kind in "iu":
        bits = (0x08 if dt.kind == "i" else 0)
        return struct.pack("<BBBBI", 0x10 | 0, bits, 0, 0, size) + struct.pack("<HH", 0, size * 8)
    if dt.kind == "f":
        # class 1, version 1; bit field: little-endian, IEEE implied mantissa normalisation (bits 4-5 = 2), sign position
        sign_pos = size * 8 - 1
        exp_bits, man_bits, bias = {2: (5, 10, 15), 4: (8, 23, 127), 8: (11, 52, 1023)}[size]
        return (struct.pack("<BBBBI", 0x10 | 1, 0x20, sign_pos, 0, size) +
                struct.pack("<HHBBBBI", 0, size * 8, man_bits, exp_bits, 0, man_bits, bias))
    raise NotImplementedError(f"write_h5: dtype {dt}")


def _msg(mtype: int, body: bytes, flags: int = 0) -> bytes:
    body = body + b"\0" * (_pad8(len(body)) - len(body))
    return struct.pack("<HHBBBB", mtype, len(body), flags, 0, 0, 0) + body


def _ohdr(msgs: List[bytes]) -> bytes:
    data = b"".join(msgs)
    return struct.pack("<BBHII", 1, 0, len(msgs), 1, len(data)) + b"\0" * 4 + data


class _Writer:
    def __init__(self):
        self.buf = bytearray()

    def alloc(self, data: bytes, align: int = 8) -> int:
        pad = (-len(self.buf)) % align
        self.buf += b"\0" * pad
        addr = len(self.buf)
        self.buf += data
        return addr

    def dataset(self, arr: np.ndarray) -> int:
        arr = np.asarray(arr)
        if arr.ndim and not arr.flags["C_CONTIGUOUS"]:
            arr = np.ascontiguousarray(arr)
        if arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("<"))
        if arr.dtype == np.bool_:
            arr = arr.astype(np.uint8)
        raw = arr.tobytes()
        data_addr = self.alloc(raw) if raw else _UNDEF
        space = struct.pack("<BBBB4x", 1, arr.ndim, 0, 0) + b"".join(struct.pack("<Q", int(d)) for d in arr.shape)
        layout = struct.pack("<BBQQ", 3, 1, data_addr, len(raw))
        return self.alloc(_ohdr([_msg(0x01, space), _msg(0x03, _dtype_msg(arr.dtype), flags=1), _msg(0x08, layout)]))

    leaf_k = 4          # set by write_h5 before any group is written (a symbol-table node holds up to 2K entries)

    def group(self, members: Dict[str, int]) -> Tuple[int, int, int]:
        """One symbol-table node holding every member (the superblock's leaf K is set large enough) -> (object header,
        B-tree, heap) addresses."""
        names = sorted(members)                                  # the B-tree orders names bytewise
        heap_data = bytearray(b"\0" * 8)                         # offset 0: the empty string (key 0 of the B-tree)
        offs = {}
        for n in names:
            offs[n] = len(heap_data)
            enc = n.encode("utf-8") + b"\0"
            heap_data += enc + b"\0" * (_pad8(len(enc)) - len(enc))
        free_off = len(heap_data)
        heap_data += struct.pack("<QQ", 1, 16)                   # one free block at the end: (next = 1 -> none, size)
        heap_data_addr = self.alloc(bytes(heap_data))
        heap = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), free_off, heap_data_addr))
        snod = b"SNOD" + struct.pack("<BBH", 1, 0, len(names))
        for n in names:
            snod += struct.pack("<QQII16x", offs[n], members[n], 0, 0)
        snod += b"\0" * (8 + 2 * self.leaf_k * 40 - len(snod))          # nodes have their full size on disk
        snod_addr = self.alloc(snod)
        last_key = offs[names[-1]] if names else 0
        tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if names else 0, _UNDEF, _UNDEF)
        tree += struct.pack("<QQQ", 0, snod_addr, last_key)
        tree += b"\0" * (24 + (2 * 16 + 1) * 8 + 2 * 16 * 8 - len(tree))   # internal K = 16: 2K+1 keys, 2K children
        tree_addr = self.alloc(tree)
        ohdr = self.alloc(_ohdr([_msg(0x11, struct.pack("<QQ", tree_addr, heap))]))
        return ohdr, tree_addr, heap

    def node(self, obj) -> Tuple[int, Optional[Tuple[int, int]]]:
        if isinstance(obj, dict):
            members = {str(k): self.node(v)[0] for k, v in obj.items()}
            o, t, h = self.group(members)
            return o, (t, h)
        return self.dataset(np.asarray(obj)), None


def write_h5(path: str, tree: Dict[str, object]) -> None:
    """Write a nested dict (dict = group, array-like = dataset) as an HDF5 file h5py / meshio can read."""
    def count(d):
        return max([len(d)] + [count(v) for v in d.values() if isinstance(v, dict)]) if d else 0
    leaf_k = max(4, (count(tree) + 1) // 2 + 1)                  # a symbol-table node holds up to 2K entries
    w = _Writer()
    w.leaf_k = leaf_k
    w.buf += b"\0" * 96                                          # superblock v0 (56) + root symbol-table entry (40)
    root_ohdr, (tree_addr, heap_addr) = w.node(tree)
    eof = len(w.buf)
    sb = _SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, leaf_k, 16, 0)
    sb += struct.pack("<QQQQ", 0, _UNDEF, eof, _UNDEF)
    sb += struct.pack("<QQII", 0, root_ohdr, 1, 0) + struct.pack("<QQ", tree_addr, heap_addr)
    assert len(sb) == 96
    w.buf[:96] = sb
    with open(path, "wb") as fh:
        fh.write(bytes(w.buf))

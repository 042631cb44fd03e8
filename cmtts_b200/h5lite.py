"""Minimal read-only HDF5 reader (numpy only) for Keras weight files.

The zero-shot path's speaker encoder (reference: deepspeaker/embedding.py:8-11, `model.m.load_weights(ckpt, by_name=True)`)
ships its weights as an HDF5 file written by h5py / Keras 2.2.4 (`ResCNN_triplet_training_checkpoint_265.h5`); this image
has neither h5py nor TensorFlow.  Such files use the oldest on-disk structures only — superblock version 0, "old style"
groups (symbol-table message -> version-1 B-tree -> SNOD symbol nodes + local heap), version-1 object headers, datasets
with contiguous (or compact / chunked) layout of little-endian fixed-width numbers — which is all this module reads
(HDF5 File Format Specification version 1.1 / 2.0, sections III.A-III.D and IV.A).  Anything else (new-style groups with
link messages / fractal heaps, variable-length types, filters other than deflate / shuffle) raises `H5Error`.

    f = H5File(path)
    f.keys("model_weights")                 # group members, in B-tree (name) order
    f["model_weights/conv64-s/conv64-s/kernel:0"]   # -> numpy array
    f.datasets("model_weights")             # {relative path: array} of every dataset below a group
"""
from __future__ import annotations

import struct
import zlib
from typing import Dict, List, Tuple

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(RuntimeError):
    pass


class _Obj:
    """Parsed object header: the messages this reader needs."""

    def __init__(self):
        self.btree = None        # (B-tree address, local-heap address) of a group
        self.shape = None
        self.dtype = None
        self.layout = None       # ("contiguous", addr, size) | ("compact", bytes) | ("chunked", btree addr, chunk dims)
        self.filters: List[int] = []


class H5File:
    def __init__(self, path: str):
        with open(path, "rb") as fh:
            self.buf = fh.read()
        b = self.buf
        if b[:8] != b"\x89HDF\r\n\x1a\n":
            raise H5Error(f"{path}: not an HDF5 file (no signature at offset 0)")
        if b[8] not in (0, 1):
            raise H5Error(f"{path}: superblock version {b[8]} (only 0 / 1: files written with the default h5py settings)")
        self.so, self.sl = b[13], b[14]
        if (self.so, self.sl) != (8, 8):
            raise H5Error("only 8-byte offsets / lengths are supported")
        p = 24 if b[8] == 0 else 28                       # version 1 adds indexed-storage K + reserved
        self.base, _free, self.eof, _drv = struct.unpack_from("<4Q", b, p)
        p += 32
        # root group symbol-table entry: link name offset, object header address, cache type, reserved, scratch
        _name, self.root_addr, cache, _r = struct.unpack_from("<QQII", b, p)
        self._objs: Dict[int, _Obj] = {}

    # ---------------------------------------------------------------- low-level structures
    def _obj(self, addr: int) -> _Obj:
        if addr in self._objs:
            return self._objs[addr]
        b = self.buf
        a = addr + self.base
        ver, _r, nmsg, _ref, hsize = struct.unpack_from("<BBHII", b, a)
        if ver != 1:
            raise H5Error(f"object header version {ver} at {addr:#x} (only version 1)")
        o = _Obj()
        blocks = [(a + 16, hsize)]                        # messages start 8-byte aligned after the 12-byte prefix
        seen = 0
        while blocks and seen < nmsg:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and seen < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", b, p)
                body = p + 8
                seen += 1
                if mtype == 0x0010:                       # continuation
                    off, ln = struct.unpack_from("<QQ", b, body)
                    blocks.append((off + self.base, ln))
                elif mtype == 0x0011:                     # symbol table (old-style group)
                    o.btree = struct.unpack_from("<QQ", b, body)
                elif mtype == 0x0001:
                    o.shape = self._dataspace(body)
                elif mtype == 0x0003:
                    o.dtype = self._datatype(body)
                elif mtype == 0x0008:
                    o.layout = self._layout(body)
                elif mtype == 0x000B:
                    o.filters = self._filters(body)
                elif mtype in (0x0002, 0x0006):
                    raise H5Error("new-style groups (link info / link messages) are not supported")
                p = body + msize
        self._objs[addr] = o
        return o

    def _dataspace(self, p: int) -> Tuple[int, ...]:
        b = self.buf
        ver, rank, flags = b[p], b[p + 1], b[p + 2]
        if ver == 1:
            q = p + 8
        elif ver == 2:
            q = p + 4
        else:
            raise H5Error(f"dataspace message version {ver}")
        return tuple(struct.unpack_from(f"<{rank}Q", b, q)) if rank else ()

    def _datatype(self, p: int) -> np.dtype:
        b = self.buf
        cls, bits0 = b[p] & 0x0F, b[p + 1]
        size = struct.unpack_from("<I", b, p + 4)[0]
        order = ">" if (bits0 & 1) else "<"
        if cls == 1 and size in (2, 4, 8):
            return np.dtype(f"{order}f{size}")
        if cls == 0 and size in (1, 2, 4, 8):
            return np.dtype(f"{order}{'i' if (bits0 & 8) else 'u'}{size}")
        if cls == 3:
            return np.dtype(f"S{size}")
        raise H5Error(f"datatype class {cls} size {size} is not supported")

    def _layout(self, p: int):
        b = self.buf
        ver = b[p]
        if ver == 3:
            cls = b[p + 1]
            if cls == 0:
                n = struct.unpack_from("<H", b, p + 2)[0]
                return ("compact", bytes(b[p + 4:p + 4 + n]))
            if cls == 1:
                addr, size = struct.unpack_from("<QQ", b, p + 2)
                return ("contiguous", addr, size)
            if cls == 2:
                rank = b[p + 2]
                addr = struct.unpack_from("<Q", b, p + 3)[0]
                dims = struct.unpack_from(f"<{rank}I", b, p + 11)
                return ("chunked", addr, dims)
            raise H5Error(f"data layout class {cls}")
        if ver in (1, 2):
            rank, cls = b[p + 1], b[p + 2]
            q = p + 8
            if cls == 0:
                dims = struct.unpack_from(f"<{rank}I", b, q)
                n = struct.unpack_from("<I", b, q + 4 * rank)[0]
                return ("compact", bytes(b[q + 4 * rank + 4:q + 4 * rank + 4 + n]))
            addr = struct.unpack_from("<Q", b, q)[0]
            dims = struct.unpack_from(f"<{rank}I", b, q + 8)
            if cls == 1:
                return ("contiguous", addr, None)
            return ("chunked", addr, dims)
        raise H5Error(f"data layout message version {ver}")

    def _filters(self, p: int) -> List[int]:
        b = self.buf
        ver, n = b[p], b[p + 1]
        q = p + (8 if ver == 1 else 2)
        ids = []
        for _ in range(n):
            fid = struct.unpack_from("<H", b, q)[0]
            if ver == 1 or fid >= 256:
                nlen, _fl, ncd = struct.unpack_from("<HHH", b, q + 2)
                q += 8
            else:                                         # version 2 omits the name length of predefined filters
                _fl, ncd = struct.unpack_from("<HH", b, q + 2)
                nlen, q = 0, q + 6
            q += (nlen + 7) // 8 * 8 if ver == 1 else nlen
            q += 4 * ncd + (4 if (ver == 1 and ncd % 2) else 0)
            ids.append(fid)
        return ids

    def _heap_name(self, heap_addr: int, off: int) -> str:
        b = self.buf
        a = heap_addr + self.base
        if b[a:a + 4] != b"HEAP":
            raise H5Error("local heap signature missing")
        data = struct.unpack_from("<Q", b, a + 24)[0] + self.base
        end = b.index(b"\x00", data + off)
        return b[data + off:end].decode("utf-8")

    def _group_entries(self, btree: int, heap: int) -> List[Tuple[str, int]]:
        """(name, object header address) of every member, walking the version-1 B-tree (node type 0) in key order."""
        b = self.buf
        out: List[Tuple[str, int]] = []
        a = btree + self.base
        if b[a:a + 4] != b"TREE" or b[a + 4] != 0:
            raise H5Error("group B-tree node signature / type mismatch")
        level, used = b[a + 5], struct.unpack_from("<H", b, a + 6)[0]
        p = a + 24                                        # after left / right sibling addresses
        for i in range(used):
            child = struct.unpack_from("<Q", b, p + 8)[0]  # key_i (8) then child_i (8)
            p += 16
            if level > 0:
                out += self._group_entries(child, heap)
                continue
            s = child + self.base
            if b[s:s + 4] != b"SNOD":
                raise H5Error("symbol table node signature missing")
            nsym = struct.unpack_from("<H", b, s + 6)[0]
            for k in range(nsym):
                e = s + 8 + 40 * k
                name_off, hdr = struct.unpack_from("<QQ", b, e)
                out.append((self._heap_name(heap, name_off), hdr))
        return out

    # ---------------------------------------------------------------- public API
    def _resolve(self, path: str) -> int:
        addr = self.root_addr
        for part in [s for s in path.split("/") if s]:
            o = self._obj(addr)
            if o.btree is None:
                raise KeyError(f"{path}: not a group on the way to {part!r}")
            members = dict(self._group_entries(*o.btree))
            if part not in members:
                raise KeyError(f"{path}: no member {part!r}")
            addr = members[part]
        return addr

    def keys(self, path: str = "/") -> List[str]:
        o = self._obj(self._resolve(path))
        if o.btree is None:
            raise KeyError(f"{path} is not a group")
        return [n for n, _ in self._group_entries(*o.btree)]

    def is_group(self, path: str) -> bool:
        return self._obj(self._resolve(path)).btree is not None

    def __getitem__(self, path: str) -> np.ndarray:
        return self._read(self._obj(self._resolve(path)), path)

    def _read(self, o: _Obj, what: str) -> np.ndarray:
        if o.layout is None or o.dtype is None or o.shape is None:
            raise KeyError(f"{what} is not a dataset")
        n = int(np.prod(o.shape, dtype=np.int64)) if o.shape else 1
        nbytes = n * o.dtype.itemsize
        kind = o.layout[0]
        if kind == "compact":
            raw = o.layout[1][:nbytes]
        elif kind == "contiguous":
            addr = o.layout[1]
            if addr == UNDEF:
                return np.zeros(o.shape, o.dtype.newbyteorder("="))
            raw = self.buf[addr + self.base:addr + self.base + nbytes]
        else:
            return self._read_chunked(o).astype(o.dtype.newbyteorder("="))
        if len(raw) != nbytes:
            raise H5Error(f"{what}: short read")
        return np.frombuffer(raw, dtype=o.dtype).reshape(o.shape).astype(o.dtype.newbyteorder("="))

    def _read_chunked(self, o: _Obj) -> np.ndarray:
        _k, btree, cdims = o.layout
        rank = len(o.shape)
        chunk = tuple(cdims[:rank])
        for f in o.filters:
            if f not in (1, 2):
                raise H5Error(f"filter id {f} is not supported (deflate = 1, shuffle = 2 only)")
        out = np.zeros(o.shape, o.dtype)
        esz = o.dtype.itemsize

        def walk(addr):
            b = self.buf
            a = addr + self.base
            if b[a:a + 4] != b"TREE" or b[a + 4] != 1:
                raise H5Error("chunk B-tree node signature / type mismatch")
            level, used = b[a + 5], struct.unpack_from("<H", b, a + 6)[0]
            p = a + 24
            ksz = 8 + 8 * (rank + 1)
            for _ in range(used):
                size, _mask = struct.unpack_from("<II", b, p)
                offs = struct.unpack_from(f"<{rank + 1}Q", b, p + 8)[:rank]
                child = struct.unpack_from("<Q", b, p + ksz)[0]
                p += ksz + 8
                if level > 0:
                    walk(child)
                    continue
                raw = bytes(b[child + self.base:child + self.base + size])
                for f in reversed(o.filters):
                    if f == 1:
                        raw = zlib.decompress(raw)
                    elif f == 2:
                        raw = np.frombuffer(raw, np.uint8).reshape(esz, -1).T.tobytes()
                blk = np.frombuffer(raw, o.dtype, count=int(np.prod(chunk))).reshape(chunk)
                sl = tuple(slice(s, min(s + c, d)) for s, c, d in zip(offs, chunk, o.shape))
                out[sl] = blk[tuple(slice(0, x.stop - x.start) for x in sl)]

        walk(btree)
        return out

    def datasets(self, path: str = "/") -> Dict[str, np.ndarray]:
        """Every dataset below `path`, keyed by its path relative to it."""
        res: Dict[str, np.ndarray] = {}

        def rec(addr: int, prefix: str):
            o = self._obj(addr)
            if o.btree is not None:
                for name, child in self._group_entries(*o.btree):
                    rec(child, f"{prefix}/{name}" if prefix else name)
            elif o.layout is not None:
                res[prefix] = self._read(o, prefix)

        rec(self._resolve(path), "")
        return res

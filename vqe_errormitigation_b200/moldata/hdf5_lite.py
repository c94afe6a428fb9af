"""Minimal read-only HDF5 reader (superblock v0, symbol-table groups, v1 object headers).

Just enough of the HDF5 file format to read the openfermion ``MolecularData`` saves the
reference keeps under ``src/mol_data/H2_*.hdf5`` without ``h5py``:
numeric scalars and n-d arrays stored compact, contiguous or chunked(+deflate).
This replaces ``MolecularData.load()`` as used at ``src/data_containers/model_hydrogen.py:56-57``
and ``src/processors/processor_quantum.py:75-77`` of the reference.

Written from the public HDF5 file-format specification (v1.1 structures); it is not a general
HDF5 implementation: no v2 object headers, fractal heaps, variable-length or compound types.
"""
from __future__ import annotations

import struct
import zlib
from typing import Dict, Iterator, Optional, Tuple

import numpy as np

_SIGNATURE = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class Hdf5FormatError(ValueError):
    pass


class _Dataset:
    __slots__ = ("shape", "dtype", "layout", "filters", "raw")

    def __init__(self):
        self.shape: Tuple[int, ...] = ()
        self.dtype: Optional[np.dtype] = None
        self.layout = None
        self.filters = []
        self.raw = None


class Hdf5File:
    """``Hdf5File(path)[name]`` returns a numpy array (0-d arrays come back as Python scalars)."""

    def __init__(self, path: str):
        with open(path, "rb") as fh:
            self._buf = fh.read()
        b = self._buf
        if b[:8] != _SIGNATURE:
            raise Hdf5FormatError(f"{path}: not an HDF5 file")
        if b[8] != 0:
            raise Hdf5FormatError(f"{path}: superblock version {b[8]} unsupported (only v0)")
        self._so, self._sl = b[13], b[14]
        if (self._so, self._sl) != (8, 8):
            raise Hdf5FormatError("only 8-byte offsets/lengths supported")
        base = 24
        self._base_addr = self._u64(base)
        root_entry = base + 4 * 8
        self._objects: Dict[str, int] = {}
        _name_off, ohdr, cache_type = struct.unpack_from("<QQI", b, root_entry)
        if cache_type == 1:
            btree, heap = struct.unpack_from("<QQ", b, root_entry + 24)
        else:
            btree, heap = self._group_from_header(ohdr)
        for name, addr in self._walk_group(btree, heap):
            self._objects[name] = addr

    # -- primitive readers -------------------------------------------------------------------
    def _u16(self, off):
        return struct.unpack_from("<H", self._buf, off)[0]

    def _u32(self, off):
        return struct.unpack_from("<I", self._buf, off)[0]

    def _u64(self, off):
        return struct.unpack_from("<Q", self._buf, off)[0]

    # -- groups ------------------------------------------------------------------------------
    def _heap_data(self, heap_addr: int) -> int:
        if self._buf[heap_addr:heap_addr + 4] != b"HEAP":
            raise Hdf5FormatError("bad local heap signature")
        return self._u64(heap_addr + 24)

    def _walk_group(self, btree: int, heap: int) -> Iterator[Tuple[str, int]]:
        data_seg = self._heap_data(heap)
        stack = [btree]
        while stack:
            node = stack.pop()
            b = self._buf
            if b[node:node + 4] != b"TREE":
                raise Hdf5FormatError("bad group B-tree signature")
            node_type, level = b[node + 4], b[node + 5]
            if node_type != 0:
                raise Hdf5FormatError("expected a group B-tree node")
            used = self._u16(node + 6)
            p = node + 24
            children = []
            for _ in range(used):
                p += 8  # key
                children.append(self._u64(p))
                p += 8
            if level > 0:
                stack.extend(reversed(children))
                continue
            for snod in children:
                if b[snod:snod + 4] != b"SNOD":
                    raise Hdf5FormatError("bad symbol node signature")
                nsym = self._u16(snod + 6)
                for i in range(nsym):
                    e = snod + 8 + 40 * i
                    name_off, ohdr = struct.unpack_from("<QQ", b, e)
                    s = data_seg + name_off
                    name = b[s:b.index(b"\0", s)].decode("ascii")
                    yield name, ohdr

    def _group_from_header(self, ohdr: int) -> Tuple[int, int]:
        for mtype, body, _size in self._messages(ohdr):
            if mtype == 0x0011:
                return self._u64(body), self._u64(body + 8)
        raise Hdf5FormatError("root group has no symbol-table message")

    # -- object headers ------------------------------------------------------------------------
    def _messages(self, ohdr: int) -> Iterator[Tuple[int, int, int]]:
        b = self._buf
        if b[ohdr] != 1:
            raise Hdf5FormatError(f"object header version {b[ohdr]} unsupported")
        nmsg = self._u16(ohdr + 2)
        hdr_size = self._u32(ohdr + 8)
        blocks = [(ohdr + 16, hdr_size)]
        seen = 0
        while blocks and seen < nmsg:
            start, length = blocks.pop(0)
            p, end = start, start + length
            while p + 8 <= end and seen < nmsg:
                mtype, msize = struct.unpack_from("<HH", b, p)
                body = p + 8
                seen += 1
                if mtype == 0x0010:
                    blocks.append((self._u64(body), self._u64(body + 8)))
                else:
                    yield mtype, body, msize
                p = body + msize

    def _parse_dataset(self, ohdr: int) -> _Dataset:
        ds = _Dataset()
        b = self._buf
        for mtype, body, msize in self._messages(ohdr):
            if mtype == 0x0001:  # dataspace
                version, rank, flags = b[body], b[body + 1], b[body + 2]
                p = body + (8 if version == 1 else 4)
                ds.shape = tuple(self._u64(p + 8 * i) for i in range(rank))
            elif mtype == 0x0003:  # datatype
                cls = b[body] & 0x0F
                bits0 = b[body + 1]
                size = self._u32(body + 4)
                order = ">" if (bits0 & 1) else "<"
                if cls == 1:
                    ds.dtype = np.dtype(f"{order}f{size}")
                elif cls == 0:
                    signed = bool(bits0 & 0x08)
                    ds.dtype = np.dtype(f"{order}{'i' if signed else 'u'}{size}")
                elif cls == 3:
                    ds.dtype = np.dtype(f"S{size}")
                else:
                    ds.dtype = None
            elif mtype == 0x0008:  # layout
                version = b[body]
                if version != 3:
                    raise Hdf5FormatError(f"layout message version {version} unsupported")
                lclass = b[body + 1]
                if lclass == 0:
                    size = self._u16(body + 2)
                    ds.layout = ("compact", body + 4, size)
                elif lclass == 1:
                    ds.layout = ("contiguous", self._u64(body + 2), self._u64(body + 10))
                elif lclass == 2:
                    ndim = b[body + 2]
                    btree = self._u64(body + 3)
                    dims = tuple(self._u32(body + 11 + 4 * i) for i in range(ndim))
                    ds.layout = ("chunked", btree, dims)
                else:
                    raise Hdf5FormatError("unknown layout class")
            elif mtype == 0x000B:  # filter pipeline
                version, nfilt = b[body], b[body + 1]
                p = body + (8 if version == 1 else 2)
                for _ in range(nfilt):
                    fid = self._u16(p)
                    if version == 1 or fid >= 256:
                        name_len = self._u16(p + 2)
                        p2 = p + 8
                    else:
                        name_len = 0
                        p2 = p + 6
                    ncd = self._u16(p + 6 if (version == 1 or fid >= 256) else p + 4)
                    p2 += name_len
                    p2 += 4 * ncd
                    if version == 1 and ncd % 2:
                        p2 += 4
                    ds.filters.append(fid)
                    p = p2
        return ds

    def _chunks(self, btree: int, ndim: int) -> Iterator[Tuple[Tuple[int, ...], int, int, int]]:
        b = self._buf
        stack = [btree]
        while stack:
            node = stack.pop()
            if b[node:node + 4] != b"TREE" or b[node + 4] != 1:
                raise Hdf5FormatError("bad chunk B-tree node")
            level = b[node + 5]
            used = self._u16(node + 6)
            key_size = 8 + 8 * ndim
            p = node + 24
            for _ in range(used):
                csize, fmask = struct.unpack_from("<II", b, p)
                offs = tuple(self._u64(p + 8 + 8 * i) for i in range(ndim))
                child = self._u64(p + key_size)
                if level == 0:
                    yield offs, csize, fmask, child
                else:
                    stack.append(child)
                p += key_size + 8

    # -- public ---------------------------------------------------------------------------------
    def keys(self):
        return sorted(self._objects)

    def __contains__(self, name: str) -> bool:
        return name in self._objects

    def __getitem__(self, name: str):
        ds = self._parse_dataset(self._objects[name])
        if ds.dtype is None or ds.layout is None:
            raise Hdf5FormatError(f"dataset {name!r}: unsupported datatype or layout")
        count = int(np.prod(ds.shape)) if ds.shape else 1
        nbytes = count * ds.dtype.itemsize
        kind = ds.layout[0]
        if kind == "compact":
            raw = self._buf[ds.layout[1]:ds.layout[1] + ds.layout[2]]
            arr = np.frombuffer(raw[:nbytes], dtype=ds.dtype)
        elif kind == "contiguous":
            addr = ds.layout[1]
            if addr == _UNDEF:
                arr = np.zeros(count, dtype=ds.dtype)
            else:
                arr = np.frombuffer(self._buf[addr:addr + nbytes], dtype=ds.dtype)
        else:
            _, btree, cdims = ds.layout
            rank = len(ds.shape)
            out = np.zeros(ds.shape, dtype=ds.dtype)
            for offs, csize, fmask, addr in self._chunks(btree, rank + 1):
                raw = self._buf[addr:addr + csize]
                for fid in reversed(ds.filters):
                    if fid == 1 and not (fmask & 1):
                        raw = zlib.decompress(raw)
                    elif fid == 2:  # shuffle
                        isz = ds.dtype.itemsize
                        raw = np.frombuffer(raw, np.uint8).reshape(isz, -1).T.tobytes()
                    elif fid not in (1, 3):
                        raise Hdf5FormatError(f"filter {fid} unsupported")
                chunk = np.frombuffer(raw, dtype=ds.dtype).reshape(cdims[:rank])
                sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs[:rank], cdims[:rank], ds.shape))
                sub = tuple(slice(0, s.stop - s.start) for s in sl)
                out[sl] = chunk[sub]
            arr = out.reshape(-1)
        arr = arr.astype(ds.dtype.newbyteorder("="), copy=True).reshape(ds.shape)
        if arr.shape == ():
            return arr.item()
        return arr

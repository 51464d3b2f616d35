"""Minimal pure-Python HDF5 reader for Keras weight files.

The reference loads per-filter surrogates with ``keras.saving.load_model``
(``nmma/em/model.py:635-648``); h5py / keras are not part of this stack, so the
dense-layer tensors are pulled straight out of the container.  Supported subset
(what legacy Keras-2 ``.h5`` files and the ``model.weights.h5`` member of a
Keras-3 ``.keras`` zip use): superblock v0/v1 (and v2/v3 with v2 object
headers), old-style groups (symbol table B-tree v1 + local heap), new-style
compact link messages, contiguous / compact / unfiltered-chunked little-endian
float / integer datasets.
"""
from __future__ import annotations

import struct
from typing import Dict, Iterator, Tuple

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class H5FormatError(ValueError):
    pass


class H5File:
    """Read-only view of an HDF5 file held in memory."""

    def __init__(self, source):
        if isinstance(source, (bytes, bytearray, memoryview)):
            self.buf = bytes(source)
        else:
            with open(source, "rb") as fh:
                self.buf = fh.read()
        if self.buf[:8] != _SIG:
            raise H5FormatError("not an HDF5 file (bad signature)")
        ver = self.buf[8]
        if ver in (0, 1):
            so, sl = self.buf[13], self.buf[14]
            if so != 8 or sl != 8:
                raise H5FormatError("only 8-byte offsets/lengths are supported")
            base = 24 if ver == 0 else 28
            self.base_addr = self._u64(base)
            # root symbol-table entry follows the four superblock addresses
            self.root_addr = self._u64(base + 32 + 8)
        elif ver in (2, 3):
            if self.buf[9] != 8 or self.buf[10] != 8:
                raise H5FormatError("only 8-byte offsets/lengths are supported")
            self.base_addr = self._u64(12)
            self.root_addr = self._u64(12 + 24)
        else:
            raise H5FormatError(f"unsupported superblock version {ver}")

    # ---- primitive readers -------------------------------------------------
    def _u8(self, o):
        return self.buf[o]

    def _u16(self, o):
        return struct.unpack_from("<H", self.buf, o)[0]

    def _u32(self, o):
        return struct.unpack_from("<I", self.buf, o)[0]

    def _u64(self, o):
        return struct.unpack_from("<Q", self.buf, o)[0]

    def _uvar(self, o, n):
        return int.from_bytes(self.buf[o:o + n], "little")

    # ---- object headers ----------------------------------------------------
    def _messages(self, addr) -> Iterator[Tuple[int, int, int]]:
        """Yield (type, offset, size) of every header message of an object."""
        addr += self.base_addr
        if self.buf[addr:addr + 4] == b"OHDR":
            yield from self._messages_v2(addr)
            return
        if self.buf[addr] != 1:
            raise H5FormatError("unsupported object header version")
        nmsg = self._u16(addr + 2)
        size = self._u32(addr + 8)
        blocks = [(addr + 16, size)]
        seen = 0
        while blocks and seen < nmsg:
            off, length = blocks.pop(0)
            end = off + length
            while off + 8 <= end and seen < nmsg:
                mtype = self._u16(off)
                msize = self._u16(off + 2)
                body = off + 8
                seen += 1
                if mtype == 0x0010:  # continuation
                    blocks.append((self._u64(body) + self.base_addr, self._u64(body + 8)))
                else:
                    yield mtype, body, msize
                off = body + msize

    def _messages_v2(self, addr):
        flags = self.buf[addr + 5]
        off = addr + 6
        if flags & 0x20:
            off += 16
        if flags & 0x10:
            off += 4
        nsz = 1 << (flags & 3)
        chunk0 = self._uvar(off, nsz)
        off += nsz
        track_order = bool(flags & 0x04)
        blocks = [(off, chunk0)]
        while blocks:
            off, length = blocks.pop(0)
            end = off + length
            while off + 4 <= end:
                mtype = self.buf[off]
                msize = self._u16(off + 1)
                body = off + 4 + (2 if track_order else 0)
                if body + msize > end:
                    break
                if mtype == 0x10:
                    caddr = self._u64(body) + self.base_addr
                    clen = self._u64(body + 8)
                    if self.buf[caddr:caddr + 4] != b"OCHK":
                        raise H5FormatError("bad continuation chunk")
                    blocks.append((caddr + 4, clen - 8))
                elif mtype != 0:
                    yield mtype, body, msize
                off = body + msize

    # ---- groups ------------------------------------------------------------
    def _heap_string(self, heap_addr, offset) -> str:
        heap_addr += self.base_addr
        if self.buf[heap_addr:heap_addr + 4] != b"HEAP":
            raise H5FormatError("bad local heap")
        data = self._u64(heap_addr + 24) + self.base_addr
        start = data + offset
        end = self.buf.index(b"\x00", start)
        return self.buf[start:end].decode("utf-8")

    def _btree_group(self, btree_addr, heap_addr, out: Dict[str, int]):
        a = btree_addr + self.base_addr
        if self.buf[a:a + 4] != b"TREE":
            raise H5FormatError("bad group B-tree")
        level = self.buf[a + 5]
        nent = self._u16(a + 6)
        p = a + 24
        for i in range(nent):
            child = self._u64(p + 8 + i * 16)
            if level > 0:
                self._btree_group(child, heap_addr, out)
            else:
                s = child + self.base_addr
                if self.buf[s:s + 4] != b"SNOD":
                    raise H5FormatError("bad symbol node")
                nsym = self._u16(s + 6)
                for k in range(nsym):
                    e = s + 8 + k * 40
                    name = self._heap_string(heap_addr, self._u64(e))
                    out[name] = self._u64(e + 8)

    def links(self, addr) -> Dict[str, int]:
        """Child name -> object header address for the group at ``addr``."""
        out: Dict[str, int] = {}
        for mtype, body, msize in self._messages(addr):
            if mtype == 0x0011:  # symbol table
                self._btree_group(self._u64(body), self._u64(body + 8), out)
            elif mtype == 0x0006:  # link message (new-style compact group)
                flags = self.buf[body + 1]
                o = body + 2
                ltype = 0
                if flags & 0x08:
                    ltype = self.buf[o]
                    o += 1
                if flags & 0x04:
                    o += 8
                if flags & 0x10:
                    o += 1
                nlen_sz = 1 << (flags & 3)
                nlen = self._uvar(o, nlen_sz)
                o += nlen_sz
                name = self.buf[o:o + nlen].decode("utf-8")
                o += nlen
                if ltype == 0:
                    out[name] = self._u64(o)
            elif mtype == 0x0002:
                # link-info with dense storage is not needed for Keras files
                fheap = self._u64(body + 2 + (8 if self.buf[body + 1] & 1 else 0))
                if fheap != _UNDEF:
                    raise H5FormatError("dense link storage not supported")
        return out

    # ---- datasets ----------------------------------------------------------
    def is_dataset(self, addr) -> bool:
        return any(m[0] == 0x0008 for m in self._messages(addr))

    def read_dataset(self, addr) -> np.ndarray:
        shape = None
        dtype = None
        layout = None
        for mtype, body, msize in self._messages(addr):
            if mtype == 0x0001:
                ver = self.buf[body]
                rank = self.buf[body + 1]
                o = body + (8 if ver == 1 else 4)
                shape = tuple(self._u64(o + 8 * i) for i in range(rank))
            elif mtype == 0x0003:
                cls = self.buf[body] & 0x0F
                bits0 = self.buf[body + 1]
                size = self._u32(body + 4)
                if bits0 & 1:
                    raise H5FormatError("big-endian data not supported")
                if cls == 1:
                    dtype = np.dtype(f"<f{size}")
                elif cls == 0:
                    signed = bool(bits0 & 0x08)
                    dtype = np.dtype(f"<{'i' if signed else 'u'}{size}")
                else:
                    raise H5FormatError(f"unsupported datatype class {cls}")
            elif mtype == 0x0008:
                ver = self.buf[body]
                if ver != 3:
                    raise H5FormatError(f"unsupported layout message version {ver}")
                lclass = self.buf[body + 1]
                if lclass == 1:
                    layout = ("contiguous", self._u64(body + 2), self._u64(body + 10))
                elif lclass == 0:
                    n = self._u16(body + 2)
                    layout = ("compact", body + 4, n)
                elif lclass == 2:
                    rank = self.buf[body + 2]
                    bt = self._u64(body + 3)
                    dims = tuple(self._u32(body + 11 + 4 * i) for i in range(rank))
                    layout = ("chunked", bt, dims)
                else:
                    raise H5FormatError("unknown layout class")
            elif mtype == 0x000B:
                raise H5FormatError("filtered (compressed) datasets not supported")
        if shape is None or dtype is None or layout is None:
            raise H5FormatError("object is not a simple dataset")
        count = int(np.prod(shape)) if shape else 1
        if layout[0] == "contiguous":
            if layout[1] == _UNDEF:
                return np.zeros(shape, dtype)
            o = layout[1] + self.base_addr
            arr = np.frombuffer(self.buf, dtype, count, o)
        elif layout[0] == "compact":
            arr = np.frombuffer(self.buf, dtype, count, layout[1])
        else:
            return self._read_chunked(layout[1], layout[2], shape, dtype)
        return arr.reshape(shape).copy()

    def _read_chunked(self, btree, cdims, shape, dtype):
        rank = len(shape)
        out = np.zeros(shape, dtype)
        cshape = cdims[:rank]

        def walk(addr):
            a = addr + self.base_addr
            if self.buf[a:a + 4] != b"TREE" or self.buf[a + 4] != 1:
                raise H5FormatError("bad chunk B-tree")
            level = self.buf[a + 5]
            nent = self._u16(a + 6)
            keysz = 8 + 8 * (rank + 1)
            p = a + 24
            for i in range(nent):
                k = p + i * (keysz + 8)
                csize = self._u32(k)
                if self._u32(k + 4):
                    raise H5FormatError("filtered chunks not supported")
                offs = tuple(self._u64(k + 8 + 8 * j) for j in range(rank))
                child = self._u64(k + keysz)
                if level > 0:
                    walk(child)
                else:
                    n = int(np.prod(cshape))
                    chunk = np.frombuffer(self.buf, dtype, n, child + self.base_addr).reshape(cshape)
                    sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cshape, shape))
                    sub = tuple(slice(0, s.stop - s.start) for s in sl)
                    out[sl] = chunk[sub]

        walk(btree)
        return out

    # ---- traversal ---------------------------------------------------------
    def walk_datasets(self, addr=None, prefix="") -> Iterator[Tuple[str, int]]:
        if addr is None:
            addr = self.root_addr
        for name, child in sorted(self.links(addr).items()):
            path = f"{prefix}/{name}"
            if self.is_dataset(child):
                yield path, child
            else:
                yield from self.walk_datasets(child, path)

    def datasets(self, under="") -> Dict[str, np.ndarray]:
        out = {}
        for path, addr in self.walk_datasets():
            if path.startswith(under):
                out[path] = self.read_dataset(addr)
        return out

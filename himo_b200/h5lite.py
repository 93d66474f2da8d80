"""h5lite -- a self-contained reader / writer for the HDF5 subset the HiMo / OpenSceneFlow pipeline uses.

The reference keeps every scene in `<scene_id>.h5`: one group per timestamp holding plain `create_dataset(name, data=...)`
arrays (OSF/dataprocess/extract_av2.py:225-238, dataprocess/extract_sca.py:76-93), read back by `HDF5Dataset`
(OSF/src/dataset.py:313-364) and extended in place by the writers of results and labels (`r+`: `del f[key][name]`,
`f[key].create_dataset(name, data=...)`, OSF/src/trainer.py:337-343, OSF/src/runner.py:187-190, OSF/process.py:97-101).
This image has neither h5py nor libhdf5, so the `.h5` contract is served by this module -- pure Python + numpy, the
format as published in the HDF5 File Format Specification (versions 1.0 / 1.1 of the superblock):

  * superblock version 0 / 1 (optionally behind a user block), 8-byte offsets and lengths;
  * "old-style" groups: version-1 object header with a Symbol Table message -> version-1 B-tree (node type 0) ->
    symbol-table nodes (SNOD) -> names in a local heap -- what libhdf5 (and therefore h5py with its default
    `libver='earliest'`) writes;
  * datasets: version-1 object headers (with continuation blocks) holding Dataspace (v1 / v2), Datatype (fixed point,
    IEEE float, enum over an integer = h5py's encoding of numpy bool), Data Layout v1 / v2 / v3, contiguous or compact,
    and chunked with an optional shuffle / deflate pipeline on the READ side;
  * writer: new files, nested groups, contiguous datasets, and in-place insertion / replacement of a dataset in an existing
    group of an existing file (also one written by the real library): the object header and data are appended, the name goes
    into the group's local heap (grown by relocation when full), the entry into its symbol-table node (split when full).

Not supported (raises `H5Unsupported`, never guesses): superblock v2 / v3 and "new-style" groups (link messages, fractal
heaps -- `libver='latest'`), compound / variable-length / string datatypes, external storage, big-endian data.

Validation (tests/test_h5lite.py): the READER is pinned on a file written by the real HDF5 library that ships in this
image (scipy's `testhdf5_7.4_GLNX86.mat`, a MATLAB v7.3 file = HDF5 with a 512-byte user block; copy under tests/golden/):
structure and the known values of its dataset; the WRITER through the pinned reader, through byte-level comparison of
the messages it emits with that file's, and by inserting datasets into a copy of that library-written file.
"""
from __future__ import annotations

import os
import struct
import zlib
from typing import Dict, Iterator, List, Optional, Tuple

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(Exception):
    pass


class H5Unsupported(H5Error):
    pass


# ------------------------------------------------------------------------------------------------ datatypes
def _dtype_from_message(d: bytes) -> Tuple[np.dtype, int]:
    """Datatype message -> (numpy dtype, bytes consumed).  numpy bool for h5py's FALSE/TRUE enum."""
    cls, ver = d[0] & 0x0F, d[0] >> 4
    bits0, bits1 = d[1], d[2]
    size = struct.unpack_from("<I", d, 4)[0]
    if cls in (0, 1) and (bits0 & 1):
        raise H5Unsupported("big-endian data")
    if cls == 0:                                      # fixed point: bit 3 of class bits 0 = signed
        signed = bool(bits0 & 0x08)
        return np.dtype(("<i" if signed else "<u") + str(size)), 8 + 4
    if cls == 1:                                      # IEEE float
        if size not in (2, 4, 8):
            raise H5Unsupported(f"float of {size} bytes")
        return np.dtype("<f" + str(size)), 8 + 12
    if cls == 8:                                      # enum: base type, names, values
        n = bits0 | (bits1 << 8)
        base, used = _dtype_from_message(d[8:])
        p = 8 + used
        names = []
        for _ in range(n):
            e = d.index(b"\0", p)
            names.append(d[p:e].decode())
            p = e + 1 if ver >= 3 else p + ((e - p + 8) // 8) * 8
        vals = np.frombuffer(d, dtype=base, count=n, offset=p)
        p += n * base.itemsize
        if sorted(names) == ["FALSE", "TRUE"] and base.itemsize == 1 and sorted(int(v) for v in vals) == [0, 1]:
            return np.dtype(bool), p
        return base, p                                # any other enum: its integer values
    raise H5Unsupported(f"datatype class {cls}")


def _dtype_message(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt == np.dtype(bool):                          # h5py: ENUM {FALSE = 0, TRUE = 1} over int8 (version-1 layout: padded names)
        base = _dtype_message(np.dtype("i1"))
        names = b"FALSE\0\0\0" + b"TRUE\0\0\0\0"
        return struct.pack("<BBBBI", 0x18, 2, 0, 0, 1) + base + names + bytes([0, 1])
    if dt.kind in "iu":
        bits = 0x08 if dt.kind == "i" else 0x00
        return struct.pack("<BBBBI", 0x10, bits, 0, 0, dt.itemsize) + struct.pack("<HH", 0, dt.itemsize * 8)
    if dt.kind == "f":
        sz = dt.itemsize
        exp_bits, man_bits, bias = {2: (5, 10, 15), 4: (8, 23, 127), 8: (11, 52, 1023)}[sz]
        # class bits: byte order 0, padding 0, mantissa normalisation 2 (implied msb) at bits 4-5, sign location in byte 1
        return (struct.pack("<BBBBI", 0x11, 0x20, sz * 8 - 1, 0, sz) +
                struct.pack("<HHBBBBI", 0, sz * 8, man_bits, exp_bits, 0, man_bits, bias))
    raise H5Unsupported(f"cannot store dtype {dt}")


# ------------------------------------------------------------------------------------------------ reading
class _Obj:
    def __init__(self, f: "File", addr: int):
        self.f, self.addr = f, addr
        self.msgs = f._read_object_header(addr)

    def msg(self, t: int) -> Optional[bytes]:
        for mt, d in self.msgs:
            if mt == t:
                return d
        return None


class Dataset:
    def __init__(self, obj: _Obj, name: str):
        self._o, self.name = obj, name
        ds = obj.msg(0x0001)
        if ds is None or obj.msg(0x0003) is None or obj.msg(0x0008) is None:
            raise H5Error(f"{name}: not a dataset")
        ver, rank, flags = ds[0], ds[1], ds[2]
        off = 8 if ver == 1 else 4
        self.shape = tuple(struct.unpack_from("<Q", ds, off + 8 * k)[0] for k in range(rank))
        self.dtype, _ = _dtype_from_message(obj.msg(0x0003))

    def __getitem__(self, key):
        return self.read()[key]

    def __array__(self, dtype=None):
        a = self.read()
        return a if dtype is None else a.astype(dtype)

    def read(self) -> np.ndarray:
        f, lay = self._o.f, self._o.msg(0x0008)
        n = int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1
        nbytes = n * self.dtype.itemsize
        ver = lay[0]
        if ver in (1, 2):
            ndim, cls = lay[1], lay[2]
            p = 8
            addr = None
            if cls != 0:
                addr = struct.unpack_from("<Q", lay, p)[0]
                p += 8
            dims = struct.unpack_from("<" + "I" * ndim, lay, p)
            p += 4 * ndim
            if cls == 0:
                size = struct.unpack_from("<I", lay, p)[0]
                raw = lay[p + 4:p + 4 + size]
            elif cls == 1:
                raw = f._read(addr, nbytes) if addr != UNDEF else bytes(nbytes)
            else:
                raw = self._read_chunked(addr, dims[:-1])
        elif ver == 3:
            cls = lay[1]
            if cls == 0:
                size = struct.unpack_from("<H", lay, 2)[0]
                raw = lay[4:4 + size]
            elif cls == 1:
                addr, size = struct.unpack_from("<QQ", lay, 2)
                raw = f._read(addr, nbytes) if addr != UNDEF else bytes(nbytes)
            elif cls == 2:
                ndim = lay[2]
                addr = struct.unpack_from("<Q", lay, 3)[0]
                dims = struct.unpack_from("<" + "I" * ndim, lay, 11)
                raw = self._read_chunked(addr, dims[:-1])
            else:
                raise H5Unsupported(f"layout class {cls}")
        else:
            raise H5Unsupported(f"data layout version {ver}")
        return np.frombuffer(raw, dtype=self.dtype, count=n).reshape(self.shape).copy()

    def _read_chunked(self, btree_addr: int, chunk: Tuple[int, ...]) -> bytes:
        """Version-1 B-tree of raw-data chunks (node type 1), optional shuffle + deflate pipeline."""
        f = self._o.f
        filters = []
        fp = self._o.msg(0x000B)
        if fp is not None:
            ver, nf = fp[0], fp[1]
            p = 8 if ver == 1 else 2
            for _ in range(nf):
                fid = struct.unpack_from("<H", fp, p)[0]
                if ver == 1 or fid >= 256:
                    nlen, _fl, ncd = struct.unpack_from("<HHH", fp, p + 2)
                    p += 8
                    p += (nlen + 7) // 8 * 8 if ver == 1 else nlen
                else:
                    _fl, ncd = struct.unpack_from("<HH", fp, p + 2)
                    p += 6
                p += 4 * ncd
                if ver == 1 and ncd % 2:
                    p += 4
                filters.append(fid)
        out = np.zeros(self.shape, dtype=self.dtype)
        rank = len(self.shape)
        if btree_addr == UNDEF:
            return out.tobytes()
        isz = self.dtype.itemsize

        def walk(addr):
            hdr = f._read(addr, 24)
            if hdr[:4] != b"TREE" or hdr[4] != 1:
                raise H5Error("bad chunk B-tree node")
            level, n = hdr[5], struct.unpack_from("<H", hdr, 6)[0]
            ksz = 8 + 8 * (rank + 1)
            body = f._read(addr + 24, n * (ksz + 8) + ksz)
            for i in range(n):
                k0 = i * (ksz + 8)
                csize, _mask = struct.unpack_from("<II", body, k0)
                offs = struct.unpack_from("<" + "Q" * (rank + 1), body, k0 + 8)[:rank]
                child = struct.unpack_from("<Q", body, k0 + ksz)[0]
                if level > 0:
                    walk(child)
                    continue
                raw = f._read(child, csize)
                for fid in reversed(filters):
                    if fid == 1:
                        raw = zlib.decompress(raw)
                    elif fid == 2:
                        a = np.frombuffer(raw, np.uint8).reshape(isz, -1)
                        raw = a.T.tobytes()
                    else:
                        raise H5Unsupported(f"filter {fid}")
                blk = np.frombuffer(raw, dtype=self.dtype, count=int(np.prod(chunk))).reshape(chunk)
                sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunk, self.shape))
                out[sl] = blk[tuple(slice(0, s.stop - s.start) for s in sl)]
        walk(btree_addr)
        return out.tobytes()


class Group:
    def __init__(self, f: "File", addr: int, name: str = "/"):
        self.f, self.addr, self.name = f, addr, name
        st = _Obj(f, addr).msg(0x0011)
        if st is None:
            if _Obj(f, addr).msg(0x0002) is not None or _Obj(f, addr).msg(0x0006) is not None:
                raise H5Unsupported("new-style group (link messages): write the file with libver='earliest'")
            raise H5Error(f"{name}: not a group")
        self.btree, self.heap = struct.unpack_from("<QQ", st, 0)

    # --- symbol table walk
    def _heap(self):
        h = self.f._read(self.heap, 32)
        if h[:4] != b"HEAP":
            raise H5Error("bad local heap")
        size, free, data = struct.unpack_from("<QQQ", h, 8)
        return size, free, data

    def _entries(self) -> Iterator[Tuple[str, int, int, int]]:
        """(name, object header address, SNOD address, entry index) in name order."""
        _, _, hdata = self._heap()
        f = self.f

        def name_at(off):
            raw = f._read(hdata + off, 256)
            while b"\0" not in raw:
                raw += f._read(hdata + off + len(raw), 256)
            return raw[:raw.index(b"\0")].decode()

        def walk(addr):
            hdr = f._read(addr, 24)
            if hdr[:4] != b"TREE" or hdr[4] != 0:
                raise H5Error("bad group B-tree node")
            level, n = hdr[5], struct.unpack_from("<H", hdr, 6)[0]
            body = f._read(addr + 24, n * 16 + 8)
            for i in range(n):
                child = struct.unpack_from("<Q", body, i * 16 + 8)[0]
                if level > 0:
                    yield from walk(child)
                else:
                    sn = f._read(child, 8)
                    if sn[:4] != b"SNOD":
                        raise H5Error("bad symbol table node")
                    cnt = struct.unpack_from("<H", sn, 6)[0]
                    ent = f._read(child + 8, cnt * 40)
                    for k in range(cnt):
                        lno, oh = struct.unpack_from("<QQ", ent, k * 40)
                        yield name_at(lno), oh, child, k
        if self.btree != UNDEF:
            yield from walk(self.btree)

    def keys(self) -> List[str]:
        return [e[0] for e in self._entries()]

    def __contains__(self, name: str) -> bool:
        return any(e[0] == name for e in self._entries()) if "/" not in name.strip("/") else self._resolve(name, True) is not None

    def _resolve(self, path: str, quiet=False):
        node = self
        parts = [p for p in path.split("/") if p]
        for i, part in enumerate(parts):
            if not isinstance(node, Group):
                if quiet:
                    return None
                raise KeyError(path)
            hit = next((e for e in node._entries() if e[0] == part), None)
            if hit is None:
                if quiet:
                    return None
                raise KeyError(f"{path!r}: no member {part!r} in {node.name!r}")
            node = node.f._open(hit[1], (node.name.rstrip("/") + "/" + part))
        return node

    def __getitem__(self, path: str):
        return self._resolve(path)

    # --- writing (File opened with mode 'w' / 'r+' / 'a')
    def create_group(self, name: str) -> "Group":
        self.f._need_write()
        parts = [p for p in name.split("/") if p]
        node = self
        for part in parts:
            ex = node._resolve(part, quiet=True)
            if ex is None:
                addr = node.f._new_group()
                node._insert(part, addr)
                ex = Group(node.f, addr, node.name.rstrip("/") + "/" + part)
            node = ex
        return node

    def require_group(self, name: str) -> "Group":
        return self.create_group(name)

    def create_dataset(self, name: str, data) -> Dataset:
        """`group.create_dataset(name, data=...)`: contiguous, no filters, like every writer of the reference.  An existing
        member of that name is REPLACED (the reference's `del f[key][name]` + create, OSF/src/trainer.py:340-343)."""
        self.f._need_write()
        if "/" in name.strip("/"):
            head, _, tail = name.strip("/").rpartition("/")
            return self.create_group(head).create_dataset(tail, data)
        arr = np.asarray(data)
        if arr.ndim and not arr.flags.c_contiguous:
            arr = np.ascontiguousarray(arr)
        if arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("<"))
        addr = self.f._new_dataset(arr)
        self._insert(name.strip("/"), addr)
        return Dataset(_Obj(self.f, addr), name)

    def __delitem__(self, name: str):
        """Remove a member's entry (its storage stays in the file as dead space, as with libhdf5 before a repack)."""
        self.f._need_write()
        hit = next((e for e in self._entries() if e[0] == name), None)
        if hit is None:
            raise KeyError(name)
        _, _, snod, k = hit
        f = self.f
        cnt = struct.unpack("<H", f._read(snod + 6, 2))[0]
        ent = bytearray(f._read(snod + 8, cnt * 40))
        del ent[k * 40:(k + 1) * 40]
        f._write(snod + 6, struct.pack("<H", cnt - 1))
        f._write(snod + 8, bytes(ent) + bytes(40))
        self._fix_keys()

    def _heap_add(self, name: str) -> int:
        """Append a NUL-terminated name to the local heap's data segment; returns its offset."""
        f = self.f
        raw = name.encode() + b"\0"
        need = (len(raw) + 7) // 8 * 8
        size, free, data = self._heap()
        # walk the free list for a block that fits
        null = lambda v: v in (UNDEF, 1) or v + 16 > size          # H5HL_FREE_NULL is 1 in libhdf5; be liberal in what we read
        prev, cur = None, free
        while not null(cur):
            nxt, bsz = struct.unpack("<QQ", f._read(data + cur, 16))
            if bsz >= need:
                rest = bsz - need
                if rest >= 16:                                   # keep the tail as a free block
                    new_free = cur + need
                    f._write(data + new_free, struct.pack("<QQ", nxt, rest))
                    link = new_free
                else:
                    need = bsz
                    link = nxt
                if prev is None:
                    f._write(self.heap + 16, struct.pack("<Q", link))
                else:
                    f._write(data + prev, struct.pack("<Q", link))
                f._write(data + cur, raw + bytes(need - len(raw)))
                return cur
            prev, cur = cur, nxt
        # no room: relocate the data segment to the end of the file at twice the size, old contents first
        new_size = max(2 * size, size + need + 64)
        old = f._read(data, size)
        new_data = f._alloc(new_size)
        f._write(new_data, old + bytes(new_size - size))
        off = size
        tail = new_size - size - need
        link = 1
        if tail >= 16:
            link = off + need
            f._write(new_data + link, struct.pack("<QQ", 1 if null(free) else free, tail))
        elif not null(free):
            link = free
        f._write(new_data + off, raw + bytes(need - len(raw)))
        f._write(self.heap + 8, struct.pack("<QQQ", new_size, link, new_data))
        return off

    def _insert(self, name: str, obj_addr: int):
        f = self.f
        hit = next((e for e in self._entries() if e[0] == name), None)
        if hit is not None:                                      # replace: repoint the existing entry
            _, _, snod, k = hit
            f._write(snod + 8 + k * 40 + 8, struct.pack("<QII", obj_addr, 0, 0) + bytes(16))
            return
        K = f.leaf_k
        hdr = f._read(self.btree, 24)
        level, n = hdr[5], struct.unpack_from("<H", hdr, 6)[0]
        if level != 0:
            raise H5Unsupported("insertion into a group whose B-tree has more than one level")
        body = bytearray(f._read(self.btree + 24, 2 * f.internal_k * 16 + 8))
        name_off = self._heap_add(name)
        _, _, hdata = self._heap()

        def nm(off):
            raw = f._read(hdata + off, 512)
            return raw[:raw.index(b"\0")]
        key = name.encode()
        if n == 0:                                               # empty group: first symbol-table node
            sn = f._alloc(8 + 2 * K * 40)
            f._write(sn, b"SNOD" + bytes([1, 0]) + struct.pack("<H", 1) + struct.pack("<QQII", name_off, obj_addr, 0, 0) +
                     bytes(16) + bytes((2 * K - 1) * 40))
            struct.pack_into("<QQQ", body, 0, 0, sn, name_off)
            f._write(self.btree + 6, struct.pack("<H", 1))
            f._write(self.btree + 24, bytes(body[:24]))
            return
        # child i holds the names in (key[i], key[i+1]]: the first child whose upper key is >= the new name, else the last
        ci = n - 1
        for i in range(n):
            upper = struct.unpack_from("<Q", body, (i + 1) * 16)[0]
            if key <= nm(upper):
                ci = i
                break
        sn = struct.unpack_from("<Q", body, ci * 16 + 8)[0]
        cnt = struct.unpack("<H", f._read(sn + 6, 2))[0]
        ents = [f._read(sn + 8 + k * 40, 40) for k in range(cnt)]
        names = [nm(struct.unpack_from("<Q", e, 0)[0]) for e in ents]
        pos = sum(1 for x in names if x < key)
        ents.insert(pos, struct.pack("<QQII", name_off, obj_addr, 0, 0) + bytes(16))
        if len(ents) <= 2 * K:
            f._write(sn + 6, struct.pack("<H", len(ents)))
            f._write(sn + 8, b"".join(ents))
        else:                                                    # split the node: lower K + 1 stay, upper K move
            if n >= 2 * f.internal_k:
                raise H5Unsupported("group B-tree root is full (h5lite writes single-level group B-trees: up to 2K * 2K_internal = "
                                    "256 members appended in name order)")
            if pos == len(ents) - 1 and ci == n - 1:            # appending in key order (timestamps, z0 z1 ...): leave the node
                lo, hi = ents[:2 * K], ents[2 * K:]             # full instead of half full -- 2K entries per node, as
            else:                                               # libhdf5's right-edge insertion does
                lo, hi = ents[:K + 1], ents[K + 1:]
            f._write(sn + 6, struct.pack("<H", len(lo)))
            f._write(sn + 8, b"".join(lo) + bytes((2 * K - len(lo)) * 40))
            sn2 = f._alloc(8 + 2 * K * 40)
            f._write(sn2, b"SNOD" + bytes([1, 0]) + struct.pack("<H", len(hi)) + b"".join(hi) + bytes((2 * K - len(hi)) * 40))
            # shift the children / keys after ci and insert (key = largest name of the lower node, child = sn2)
            tail = bytes(body[(ci + 1) * 16:(n * 16 + 8)])
            struct.pack_into("<QQ", body, (ci + 1) * 16, 0, sn2)
            body[(ci + 2) * 16:(ci + 2) * 16 + len(tail)] = tail
            n += 1
            f._write(self.btree + 6, struct.pack("<H", n))
            f._write(self.btree + 24, bytes(body[:n * 16 + 8]))
        self._fix_keys()

    def _fix_keys(self):
        """Recompute the keys of the (single-level) B-tree root: key[0] = "", key[i+1] = largest name of child i."""
        f = self.f
        hdr = f._read(self.btree, 24)
        n = struct.unpack_from("<H", hdr, 6)[0]
        body = bytearray(f._read(self.btree + 24, n * 16 + 8))
        struct.pack_into("<Q", body, 0, 0)
        last = 0
        for i in range(n):
            sn = struct.unpack_from("<Q", body, i * 16 + 8)[0]
            cnt = struct.unpack("<H", f._read(sn + 6, 2))[0]
            if cnt:
                last = struct.unpack("<Q", f._read(sn + 8 + (cnt - 1) * 40, 8))[0]
            struct.pack_into("<Q", body, (i + 1) * 16, last)
        f._write(self.btree + 24, bytes(body))


class File(Group):
    """`h5lite.File(path, mode)` with mode 'r', 'r+', 'w' or 'a' -- the subset of `h5py.File` the pipeline uses:
    `f[path]`, `name in f`, `f.keys()`, `f.create_group`, `group.create_dataset(name, data=...)`, `del group[name]`,
    context manager."""

    def __init__(self, path: str, mode: str = "r"):
        if mode not in ("r", "r+", "w", "a"):
            raise ValueError(mode)
        exists = os.path.exists(path)
        if mode == "a":
            mode = "r+" if exists else "w"
        self.path, self.mode = path, mode
        self.writable = mode != "r"
        if mode == "w":
            self.fh = open(path, "w+b")
            self.base = 0
            self.leaf_k, self.internal_k = 4, 16
            self.eof = 0
            self._format_new()
        else:
            self.fh = open(path, "r+b" if self.writable else "rb")
            self._parse_superblock()
        Group.__init__(self, self, self.root_addr, "/")

    # ---- context manager
    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
        return False

    def close(self):
        if self.fh is not None:
            if self.writable:
                self._store_eof()
                self.fh.flush()
            self.fh.close()
            self.fh = None

    # ---- raw access (addresses are relative to the base address, as the format defines them)
    def _read(self, addr: int, n: int) -> bytes:
        self.fh.seek(self.base + addr)
        b = self.fh.read(n)
        return b + bytes(n - len(b))

    def _write(self, addr: int, data: bytes):
        self.fh.seek(self.base + addr)
        self.fh.write(data)
        self.eof = max(self.eof, addr + len(data))

    def _alloc(self, n: int, align: int = 8) -> int:
        addr = (self.eof + align - 1) // align * align
        self.eof = addr + n
        self.fh.seek(self.base + addr)
        self.fh.write(bytes(n))
        return addr

    def _need_write(self):
        if not self.writable:
            raise H5Error("file is open read-only")

    # ---- superblock
    def _parse_superblock(self):
        self.fh.seek(0, 2)
        flen = self.fh.tell()
        off = 0
        while True:
            self.fh.seek(off)
            if self.fh.read(8) == SIGNATURE:
                break
            off = 512 if off == 0 else off * 2
            if off >= flen:
                raise H5Error(f"{self.path}: not an HDF5 file")
        self.sb_off = off
        self.fh.seek(off)
        sb = self.fh.read(128)
        ver = sb[8]
        if ver not in (0, 1):
            raise H5Unsupported(f"superblock version {ver} (file written with libver='latest'?)")
        if sb[13] != 8 or sb[14] != 8:
            raise H5Unsupported("offsets / lengths that are not 8 bytes")
        self.leaf_k, self.internal_k = struct.unpack_from("<HH", sb, 16)
        p = 24 + (4 if ver == 1 else 0)
        self.base, _free, self.eof, _drv = struct.unpack_from("<QQQQ", sb, p)
        self._eof_pos = off + p + 16
        p += 32
        _lno, self.root_addr, cache = struct.unpack_from("<QQI", sb, p)

    def _store_eof(self):
        self.fh.seek(self._eof_pos)
        self.fh.write(struct.pack("<Q", self.eof))

    def _format_new(self):
        """Superblock v0 + root group, laid out like libhdf5 lays out a fresh file."""
        self.sb_off = 0
        self.eof = 96                                            # superblock (56 bytes + root symbol-table entry 40)
        self.root_addr = self._new_group()
        st = _Obj(self, self.root_addr).msg(0x0011)
        btree, heap = struct.unpack_from("<QQ", st, 0)
        sb = SIGNATURE + bytes([0, 0, 0, 0, 0, 8, 8, 0]) + struct.pack("<HHI", self.leaf_k, self.internal_k, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, self.eof, UNDEF)
        sb += struct.pack("<QQII", 0, self.root_addr, 1, 0) + struct.pack("<QQ", btree, heap)
        self._eof_pos = 24 + 16
        self.fh.seek(0)
        self.fh.write(sb)

    # ---- object headers
    def _read_object_header(self, addr: int) -> List[Tuple[int, bytes]]:
        hdr = self._read(addr, 16)
        if hdr[:4] == b"OHDR":
            raise H5Unsupported("version-2 object header (libver='latest')")
        if hdr[0] != 1:
            raise H5Error(f"bad object header at {addr}")
        nmsg = struct.unpack_from("<H", hdr, 2)[0]
        hsize = struct.unpack_from("<I", hdr, 8)[0]
        blocks = [(addr + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            a, ln = blocks.pop(0)
            blk = self._read(a, ln)
            p = 0
            while p + 8 <= ln and len(out) < nmsg:
                t, sz, _fl = struct.unpack_from("<HHB", blk, p)
                d = blk[p + 8:p + 8 + sz]
                out.append((t, d))
                if t == 0x0010:
                    o, l2 = struct.unpack_from("<QQ", d, 0)
                    blocks.append((o, l2))
                p += 8 + sz
        return out

    def _open(self, addr: int, name: str):
        o = _Obj(self, addr)
        if o.msg(0x0011) is not None:
            return Group(self, addr, name)
        return Dataset(o, name)

    @staticmethod
    def _msg(t: int, data: bytes, flags: int = 0) -> bytes:
        pad = (-len(data)) % 8
        return struct.pack("<HHBBBB", t, len(data) + pad, flags, 0, 0, 0) + data + bytes(pad)

    def _write_header(self, msgs: List[bytes], min_size: int = 0) -> int:
        body = b"".join(msgs)
        if len(body) < min_size:                                  # pad with one NIL message, as the library does
            body += self._msg(0x0000, bytes(min_size - len(body) - 8))
            msgs = msgs + [b""]
        addr = self._alloc(16 + len(body))
        self._write(addr, struct.pack("<BBHII", 1, 0, len(msgs), 1, len(body)) + bytes(4) + body)
        return addr

    def _new_group(self) -> int:
        K, IK = self.leaf_k, self.internal_k
        heap_data_size = 88
        heap = self._alloc(32)
        data = self._alloc(heap_data_size)
        # offset 0 holds the empty string (the B-tree's first key); the rest is one free block
        self._write(data, bytes(8) + struct.pack("<QQ", 1, heap_data_size - 8))    # next = 1 = "no next" as the library writes it
        self._write(heap, b"HEAP" + bytes(4) + struct.pack("<QQQ", heap_data_size, 8, data))
        bt = self._alloc(24 + 2 * IK * 16 + 8)
        self._write(bt, b"TREE" + bytes([0, 0]) + struct.pack("<H", 0) + struct.pack("<QQ", UNDEF, UNDEF))
        return self._write_header([self._msg(0x0011, struct.pack("<QQ", bt, heap), 0)], min_size=24)

    def _new_dataset(self, arr: np.ndarray) -> int:
        shape = arr.shape
        rank = len(shape)
        nbytes = arr.nbytes
        data_addr = UNDEF
        if nbytes:
            data_addr = self._alloc(nbytes)
            self._write(data_addr, arr.tobytes())
        dataspace = struct.pack("<BBBBI", 1, rank, 0, 0, 0) + b"".join(struct.pack("<Q", s) for s in shape)
        msgs = [
            self._msg(0x0001, dataspace),
            self._msg(0x0003, _dtype_message(arr.dtype), 1),
            self._msg(0x0005, bytes([1, 2, 2, 1, 0, 0, 0, 0]), 1),  # fill value exactly as the library-written fixture has it: v1, late
                                                                   # allocation, write "if set", defined with size 0
            self._msg(0x0008, bytes([3, 1]) + struct.pack("<QQ", data_addr, nbytes)),   # layout v3, contiguous
        ]
        return self._write_header(msgs, min_size=sum(len(m) for m in msgs) + 24)

"""Minimal HDF5 reader / writer for the reference's data files (no h5py / hdf5storage in this image).

Upstream stores its synthetic observations with ``hdf5storage.write`` (src/data_generation_2sam_more_loss.py:
256-268) and reads them back with ``hdf5storage.read`` (main_custom_training.py:76): a MATLAB-7.3 style HDF5
file -- 512-byte user block, version-0 superblock, old-style root group (B-tree v1 + local heap + symbol
nodes), one dataset per dictionary key, float64, stored TRANSPOSED (MATLAB order) and, in the shipped
``data_fem_test_big_noise.h5``, chunked with shuffle + deflate + fletcher32.

``read`` understands exactly that subset (contiguous / compact / chunked v3 layouts, the three filters,
fixed-point and IEEE float types); ``write`` produces the same kind of file with contiguous, unfiltered
datasets and the ``MATLAB_class`` attribute hdf5storage looks for.  Arrays are transposed on the way in and out
like hdf5storage's ``matlab_compatible`` mode does, so ``read(write(d)) == d``.
"""
from __future__ import annotations

import struct
import time
import zlib

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


# ------------------------------------------------------------------------------------------------ reader
class _File:
    def __init__(self, buf):
        self.b = buf
        pos = 0
        while buf[pos:pos + 8] != _SIG:   # the superblock sits at 0, 512, 1024, ...
            pos = 512 if pos == 0 else pos * 2
            if pos >= len(buf):
                raise ValueError("not an HDF5 file")
        if buf[pos + 8] != 0:
            raise ValueError(f"superblock version {buf[pos + 8]} not supported (expected 0)")
        if buf[pos + 13] != 8 or buf[pos + 14] != 8:
            raise ValueError("only 8-byte offsets and lengths are supported")
        self.base = struct.unpack_from("<Q", buf, pos + 24)[0]
        self.root = self.symbol_entry(pos + 56)

    def u(self, fmt, off):
        return struct.unpack_from("<" + fmt, self.b, off)

    def symbol_entry(self, off):
        name_off, header, cache = self.u("QQI", off)
        scratch = self.u("QQ", off + 24)
        return {"name_off": name_off, "header": header, "cache": cache, "btree": scratch[0], "heap": scratch[1]}

    # -- old-style groups
    def group_entries(self, btree, heap):
        h = self.base + heap
        if self.b[h:h + 4] != b"HEAP":
            raise ValueError("bad local heap")
        data = self.base + self.u("Q", h + 24)[0]
        out = []

        def walk(addr):
            a = self.base + addr
            if self.b[a:a + 4] == b"SNOD":
                n = self.u("H", a + 6)[0]
                for i in range(n):
                    e = self.symbol_entry(a + 8 + 40 * i)
                    s = data + e["name_off"]
                    e["name"] = bytes(self.b[s:self.b.index(b"\0", s)]).decode()
                    out.append(e)
                return
            if self.b[a:a + 4] != b"TREE" or self.b[a + 4] != 0:
                raise ValueError("bad group B-tree node")
            n = self.u("H", a + 6)[0]
            for i in range(n):
                walk(self.u("Q", a + 24 + 8 + 16 * i)[0])

        walk(btree)
        return out

    # -- object headers (version 1)
    def messages(self, addr):
        a = self.base + addr
        if self.b[a] != 1:
            raise ValueError(f"object header version {self.b[a]} not supported (expected 1)")
        nmsg = self.u("H", a + 2)[0]
        size = self.u("I", a + 8)[0]
        blocks, msgs = [(a + 16, size)], []
        while blocks and len(msgs) < nmsg:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(msgs) < nmsg:
                mtype, msize, _flags = self.u("HHB", p)
                body = p + 8
                if mtype == 0x10:
                    off, ln = self.u("QQ", body)
                    blocks.append((self.base + off, ln))
                msgs.append((mtype, body, msize))
                p = body + msize
        return msgs

    def dataset(self, addr):
        shape = dtype = layout = None
        filters, attrs = [], {}
        for mtype, p, size in self.messages(addr):
            if mtype == 0x01:
                ver, rank = self.b[p], self.b[p + 1]
                start = p + (8 if ver == 1 else 4)
                shape = self.u(f"{rank}Q", start) if rank else ()
            elif mtype == 0x03:
                dtype = self.datatype(p)
            elif mtype == 0x08:
                if self.b[p] != 3:
                    raise ValueError("only version-3 data layouts are supported")
                cls = self.b[p + 1]
                if cls == 0:
                    n = self.u("H", p + 2)[0]
                    layout = ("compact", p + 4, n)
                elif cls == 1:
                    layout = ("contiguous",) + self.u("QQ", p + 2)
                else:
                    nd = self.b[p + 2]
                    layout = ("chunked", self.u("Q", p + 3)[0], self.u(f"{nd}I", p + 11))
            elif mtype == 0x0B:
                filters = self.filter_pipeline(p)
            elif mtype == 0x0C:
                k, v = self.attribute(p)
                attrs[k] = v
        return shape, dtype, layout, filters, attrs

    def datatype(self, p):
        cls, b0 = self.b[p] & 15, self.b[p + 1]
        size = self.u("I", p + 4)[0]
        order = ">" if b0 & 1 else "<"
        if cls == 1:
            return np.dtype(f"{order}f{size}")
        if cls == 0:
            return np.dtype(f"{order}{'i' if b0 & 8 else 'u'}{size}")
        if cls == 3:
            return np.dtype(f"S{size}")
        raise ValueError(f"datatype class {cls} not supported")

    def filter_pipeline(self, p):
        ver, n = self.b[p], self.b[p + 1]
        q = p + (8 if ver == 1 else 2)
        ids = []
        for _ in range(n):
            fid = self.u("H", q)[0]
            if ver == 1 or fid >= 256:
                nlen, _fl, nvals = self.u("HHH", q + 2)
                q += 8 + ((nlen + 7) & ~7 if ver == 1 else nlen)
            else:
                _fl, nvals = self.u("HH", q + 2)
                q += 6
            q += 4 * nvals + (4 if ver == 1 and nvals % 2 else 0)
            ids.append(fid)
        return ids

    def attribute(self, p):
        if self.b[p] != 1:
            return f"@{p}", None
        nlen, tlen, slen = self.u("HHH", p + 2)
        pad = lambda v: (v + 7) & ~7
        q = p + 8
        name = self.b[q:q + nlen].split(b"\0")[0].decode()
        q += pad(nlen)
        dt = self.datatype(q)
        q += pad(tlen)
        rank = self.b[q + 1]
        shape = self.u(f"{rank}Q", q + 8) if rank else ()
        q += pad(slen)
        cnt = int(np.prod(shape)) if shape else 1
        val = np.frombuffer(self.b, dtype=dt, count=cnt, offset=q)
        return name, (val[0] if not shape else val.reshape(shape))

    # -- raw data
    def read_data(self, shape, dtype, layout, filters):
        n = int(np.prod(shape)) if shape else 1
        if layout[0] == "compact":
            return np.frombuffer(self.b, dtype=dtype, count=n, offset=layout[1]).reshape(shape).copy()
        if layout[0] == "contiguous":
            if layout[1] == _UNDEF:
                return np.zeros(shape, dtype=dtype)
            return np.frombuffer(self.b, dtype=dtype, count=n, offset=self.base + layout[1]).reshape(shape).copy()
        _, btree, cdims = layout
        cshape, esize = cdims[:-1], cdims[-1]
        out = np.zeros(shape, dtype=dtype)
        rank = len(shape)

        def walk(addr):
            a = self.base + addr
            if self.b[a:a + 4] != b"TREE" or self.b[a + 4] != 1:
                raise ValueError("bad chunk B-tree node")
            level, used = self.b[a + 5], self.u("H", a + 6)[0]
            ksz = 8 + 8 * (rank + 1)
            for i in range(used):
                k = a + 24 + i * (ksz + 8)
                nbytes, mask = self.u("II", k)
                offs = self.u(f"{rank}Q", k + 8)
                child = self.u("Q", k + ksz)[0]
                if level > 0:
                    walk(child)
                    continue
                raw = bytes(self.b[self.base + child:self.base + child + nbytes])
                for pos in range(len(filters) - 1, -1, -1):
                    if mask & (1 << pos):
                        continue
                    fid = filters[pos]
                    if fid == 3:      # fletcher32: checksum in the last four bytes
                        raw = raw[:-4]
                    elif fid == 1:    # deflate
                        raw = zlib.decompress(raw)
                    elif fid == 2:    # shuffle
                        raw = np.frombuffer(raw, dtype=np.uint8).reshape(esize, -1).T.tobytes()
                    else:
                        raise ValueError(f"filter {fid} not supported")
                chunk = np.frombuffer(raw, dtype=dtype).reshape(cshape)
                sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cshape, shape))
                out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]

        walk(btree)
        return out


def read(path=None, filename=None, matlab_compatible=True):
    """``hdf5storage.read(path='/', filename=...)`` for the files described above: a dict name -> ndarray of
    the root group's datasets (arrays transposed back from MATLAB order)."""
    fname = filename if filename is not None else path
    with open(fname, "rb") as fh:
        f = _File(fh.read())
    out = {}
    for e in f.group_entries(f.root["btree"], f.root["heap"]):
        if e["cache"] == 1:   # a sub-group (hdf5storage's #refs# etc.): not part of the flat data files
            continue
        shape, dtype, layout, filters, _attrs = f.dataset(e["header"])
        if dtype is None or layout is None:
            continue
        a = f.read_data(shape, dtype, layout, filters)
        out[e["name"]] = np.ascontiguousarray(a.T) if matlab_compatible else a
    return out


# ------------------------------------------------------------------------------------------------ writer
def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


def _msg(mtype, body, flags=0):
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _dataspace(shape):
    return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", s) for s in shape)


_F64 = struct.pack("<BBBBI", 0x11, 0x20, 0x3F, 0x00, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)


def _string_type(n):
    return struct.pack("<BBBBI", 0x13, 0x00, 0x00, 0x00, n)


def _attribute(name, text):
    nm = name.encode() + b"\0"
    dt, ds = _string_type(len(text)), _dataspace(())
    return (struct.pack("<BxHHH", 1, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) + text.encode())


def _fletcher32(raw):
    """HDF5's Fletcher-32 (H5_checksum_fletcher32): 16-bit big-endian words, sums modulo 65535."""
    b = np.frombuffer(raw + (b"\0" if len(raw) % 2 else b""), dtype=">u2").astype(np.uint64)
    s1 = int(b.sum() % 65535)
    s2 = int((b * np.arange(len(b), 0, -1, dtype=np.uint64)).sum() % 65535)
    return struct.pack("<I", (s2 << 16) | s1)


def _filter_pipeline_msg(esize, level=4):
    def one(fid, vals):
        body = struct.pack("<HHHH", fid, 0, 1 if fid != 3 else 0, len(vals)) + b"".join(struct.pack("<I", v) for v in vals)
        return body + (b"\0" * 4 if len(vals) % 2 else b"")
    return struct.pack("<BB6x", 1, 3) + one(2, [esize]) + one(1, [level]) + one(3, [])


def write(data, path=None, filename=None, matlab_compatible=True, compress=False):
    """``hdf5storage.write(data=dict, path='/', filename=...)``: every value becomes a float64 dataset of the
    root group, stored transposed (MATLAB order) with the ``MATLAB_class`` attribute; contiguous, or with
    ``compress=True`` chunked with shuffle + deflate + fletcher32 like the file upstream ships."""
    fname = filename if filename is not None else path
    names = sorted(data)   # symbol-table entries are ordered by name
    base = 512
    blob = bytearray()     # everything behind the user block; addresses are relative to `base`

    def put(b, align=8):
        blob.extend(b"\0" * (-len(blob) % align))
        at = len(blob)
        blob.extend(b)
        return at

    leaf_k = max(4, (len(names) + 1) // 2)
    blob.extend(b"\0" * 96)                       # superblock, filled in last
    heap_data = bytearray(b"\0" * 8)              # offset 0: the empty name
    name_off = {}
    for nm in names:
        name_off[nm] = len(heap_data)
        heap_data.extend(_pad8(nm.encode() + b"\0"))
    heap_data_at = put(bytes(heap_data))
    heap_at = put(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), _UNDEF, heap_data_at))
    headers = {}
    for nm in names:
        a = np.asarray(data[nm], dtype=np.float64)
        a = np.atleast_2d(a)
        stored = np.ascontiguousarray(a.T) if matlab_compatible else np.ascontiguousarray(a)
        msgs = [_msg(0x01, _dataspace(stored.shape)), _msg(0x03, _F64, flags=1),
                _msg(0x05, struct.pack("<BBBB", 2, 2, 2, 0))]                       # fill value: undefined, late alloc
        if compress:
            # chunks along the last (long) axis; per chunk: shuffle -> deflate -> fletcher32
            rank, step = stored.ndim, max(1, min(stored.shape[-1], 4096 // max(1, stored.shape[0])))
            cshape = stored.shape[:-1] + (step,)
            keys = []
            for o in range(0, stored.shape[-1], step):
                chunk = np.zeros(cshape)
                part = stored[..., o:o + step]
                chunk[..., :part.shape[-1]] = part
                raw = np.frombuffer(chunk.tobytes(), dtype=np.uint8).reshape(-1, 8).T.tobytes()
                raw = zlib.compress(raw, 4)
                raw += _fletcher32(raw)
                keys.append((len(raw), (0,) * (rank - 1) + (o,), put(raw)))
            if len(keys) > 64:
                raise ValueError("dataset too large for the single-node chunk index of this writer")
            node = b"TREE" + struct.pack("<BBHQQ", 1, 0, len(keys), _UNDEF, _UNDEF)
            for nbytes, offs, at in keys:
                node += struct.pack("<II", nbytes, 0) + struct.pack(f"<{rank + 1}Q", *offs, 0) + struct.pack("<Q", at)
            node += struct.pack("<II", 0, 0) + struct.pack(f"<{rank + 1}Q", *((0,) * (rank - 1) + (len(keys) * step,)), 0)
            node += b"\0" * ((64 - len(keys)) * (8 + 8 * (rank + 1) + 8))
            bt_at = put(node)
            msgs += [_msg(0x0B, _filter_pipeline_msg(8)),
                     _msg(0x08, struct.pack("<BBBQ", 3, 2, rank + 1, bt_at) + struct.pack(f"<{rank + 1}I", *cshape, 8))]
        else:
            raw_at = put(stored.tobytes())
            msgs.append(_msg(0x08, struct.pack("<BBQQ", 3, 1, raw_at, stored.nbytes)))   # contiguous layout
        msgs.append(_msg(0x0C, _attribute("MATLAB_class", "double")))
        body = b"".join(msgs)
        headers[nm] = put(struct.pack("<BxHII4x", 1, len(msgs), 1, len(body)) + body)
    snod = b"SNOD" + struct.pack("<BxH", 1, len(names))
    for nm in names:
        snod += struct.pack("<QQII16x", name_off[nm], headers[nm], 0, 0)
    snod += b"\0" * (40 * (2 * leaf_k - len(names)))
    snod_at = put(snod)
    last = name_off[names[-1]] if names else 0
    tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, _UNDEF, _UNDEF) + struct.pack("<QQQ", 0, snod_at, last)
    tree += b"\0" * (16 * (2 * 16 - 1))           # room for 2K entries of an internal node (K = 16)
    tree_at = put(tree)
    root_hdr = put(struct.pack("<BxHII4x", 1, 1, 1, 24) + _msg(0x11, struct.pack("<QQ", tree_at, heap_at)))
    eof = len(blob)
    sb = _SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, leaf_k, 16, 0)
    sb += struct.pack("<QQQQ", base, _UNDEF, eof, _UNDEF)
    sb += struct.pack("<QQII", 0, root_hdr, 1, 0) + struct.pack("<QQ", tree_at, heap_at)
    blob[:len(sb)] = sb
    head = ("MATLAB 7.3 MAT-file, Platform: vbfem-b200 h5io, Created on: "
            + time.strftime("%a %b %d %H:%M:%S %Y") + " HDF5 schema 1.00 .").encode()
    user = head.ljust(116, b" ") + b"\0" * 8 + b"\x00\x02IM"
    with open(fname, "wb") as fh:
        fh.write(user.ljust(base, b"\0"))
        fh.write(bytes(blob))
    return fname

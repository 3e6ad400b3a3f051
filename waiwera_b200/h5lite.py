"""A small reader / writer for the subset of HDF5 that Waiwera's files use (output, and the
restart files `initial.filename` points at).  There is no HDF5 library in this image (no libhdf5, h5py, netCDF-4), so this
module parses and writes the file format itself, from the published HDF5 File Format Specification (version 1 structures,
which is what PETSc's HDF5 viewer writes with the library's default "earliest" format bounds):

    superblock version 0; "old-style" groups (symbol-table message -> B-tree version 1 of group nodes + SNOD symbol
    nodes + local heap); object headers version 1 incl. continuation blocks; dataspace messages version 1 / 2;
    fixed-point and IEEE floating-point datatypes, fixed-length strings; data layout message version 3 (compact,
    contiguous, chunked with a version-1 B-tree of chunks) and versions 1 / 2;
    and, read only, what netCDF-4 adds to that (the reference's benchmark meshes are ExodusII files in that container):
    object headers version 2 ("OHDR" / "OCHK"), "new-style" groups (link messages in the header, or in a fractal heap
    with direct blocks under one indirect block, indexed by a version-2 B-tree of depth <= 1), attribute messages
    versions 1-3 (in the header or in a fractal heap), filter pipelines with deflate / shuffle / fletcher32.

Reading: H5File(path) -> .datasets() (paths), [path] -> numpy array, .groups(), .attrs(path).  Writing: write(path, {"a/b": array})
lays down the same structures with contiguous datasets.  The reader is checked against the reference's own files
(tools/make_golden.py reads them with it; tests/test_h5lite.py); the writer against the reader and against structural
checks -- it cannot be checked against libhdf5 here, which DESIGN.md states.

Reference: what the files hold is set up in src/flow_simulation.F90 (output_* routines: "time", "cell_fields/...",
"source_fields/...", "cell_index", "source_index") through PETSc's VecView on an HDF5 viewer (timestepping datasets
[time, ...], chunked); src/initial.F90:setup_initial_file reads them back for a restart."""
import struct

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(Exception):
    pass


class _Dataset:
    def __init__(self, shape, dtype, layout, maxshape=None):
        self.shape, self.dtype, self.layout, self.maxshape = tuple(shape), dtype, layout, maxshape


class H5File:
    """read-only view of a file in the subset above"""

    def __init__(self, path):
        self.path = path
        self.buf = open(path, "rb").read()
        b = self.buf
        if b[:8] != SIG:
            raise H5Error("%s: not an HDF5 file" % path)
        ver = b[8]
        if ver not in (0, 1):
            raise H5Error("%s: superblock version %d is not supported (only the version-1 structures are)" % (path, ver))
        self.O, self.L = b[13], b[14]
        if self.O != 8 or self.L != 8:
            raise H5Error("%s: %d-byte offsets / %d-byte lengths are not supported" % (path, self.O, self.L))
        pos = 16 + 2 + 2 + 4
        if ver == 1:
            pos += 4
        self.base = self._u(pos, 8)
        pos += 4 * 8                              # base, free-space info, end of file, driver info
        self.root_entry = self._symbol_entry(pos)
        self._objects = {}
        self._headers = {}
        self._walk("", self.root_entry["header"], self.root_entry)

    # ---- primitives
    def _u(self, pos, n):
        return int.from_bytes(self.buf[pos:pos + n], "little")

    def _symbol_entry(self, pos):
        e = {"name_off": self._u(pos, 8), "header": self._u(pos + 8, 8), "cache": self._u(pos + 16, 4)}
        if e["cache"] == 1:
            e["btree"], e["heap"] = self._u(pos + 24, 8), self._u(pos + 32, 8)
        return e

    def _messages(self, addr):
        """(type, payload bytes) of an object header version 1, following continuation blocks"""
        b = self.buf
        addr += self.base
        if b[addr:addr + 4] == b"OHDR":
            return self._messages_v2(addr)
        if b[addr] != 1:
            raise H5Error("object header version %d is not supported" % b[addr])
        nmsg = self._u(addr + 2, 2)
        size = self._u(addr + 8, 4)
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            pos, left = blocks.pop(0)
            end = pos + left
            while pos + 8 <= end and len(out) < nmsg:
                mtype, msize = self._u(pos, 2), self._u(pos + 2, 2)
                data = b[pos + 8:pos + 8 + msize]
                pos += 8 + msize
                if mtype == 0x10:
                    blocks.append((self.base + int.from_bytes(data[:8], "little"), int.from_bytes(data[8:16], "little")))
                out.append((mtype, data))
        return out

    def _messages_v2(self, addr):
        """(type, payload bytes) of an object header version 2 ("OHDR", with "OCHK" continuation chunks)"""
        b = self.buf
        if b[addr + 4] != 2:
            raise H5Error("object header version %d is not supported" % b[addr + 4])
        flags = b[addr + 5]
        pos = addr + 6 + (16 if flags & 0x20 else 0) + (4 if flags & 0x10 else 0)
        nsz = 1 << (flags & 3)
        size = self._u(pos, nsz)
        blocks = [(pos + nsz, size)]
        mhdr = 6 if flags & 0x04 else 4
        out = []
        while blocks:
            pos, left = blocks.pop(0)
            end = pos + left
            while pos + mhdr <= end:
                mtype, msize = b[pos], self._u(pos + 1, 2)
                data = b[pos + mhdr:pos + mhdr + msize]
                pos += mhdr + msize
                if mtype == 0x10:
                    caddr, clen = self.base + int.from_bytes(data[:8], "little"), int.from_bytes(data[8:16], "little")
                    if b[caddr:caddr + 4] != b"OCHK":
                        raise H5Error("bad object header continuation chunk")
                    blocks.append((caddr + 4, clen - 8))          # signature in front, checksum behind
                elif mtype != 0:
                    out.append((mtype, data))
        return out

    # ---- version-2 groups: link messages in the header, or in a fractal heap ("dense" storage)
    def _link(self, d):
        """(name, object header address or None for soft / external links, bytes used) of a link message"""
        flags = d[1]
        pos = 2
        ltype = 0
        if flags & 0x08:
            ltype = d[pos]
            pos += 1
        if flags & 0x04:
            pos += 8
        if flags & 0x10:
            pos += 1
        nsz = 1 << (flags & 3)
        nlen = int.from_bytes(d[pos:pos + nsz], "little")
        pos += nsz
        name = bytes(d[pos:pos + nlen]).decode()
        pos += nlen
        if ltype == 0:
            return name, int.from_bytes(d[pos:pos + 8], "little"), pos + 8
        vlen = int.from_bytes(d[pos:pos + 2], "little")
        return name, None, pos + 2 + vlen

    def _fractal_heap(self, addr):
        """header fields and the direct blocks [(heap offset, file address, size)] of a fractal heap"""
        b = self.buf
        h = self.base + addr
        if b[h:h + 4] != b"FRHP" or b[h + 4] != 0:
            raise H5Error("bad fractal heap header")
        if self._u(h + 7, 2):
            raise H5Error("filtered fractal heaps are not supported")
        hflags = b[h + 9]
        max_man = self._u(h + 10, 4)
        pos = h + 14 + 12 * 8                          # next huge id ... number of tiny objects
        width = self._u(pos, 2)
        start, max_direct = self._u(pos + 2, 8), self._u(pos + 10, 8)
        max_bits = self._u(pos + 18, 2)
        root, nrows = self._u(pos + 22, 8), self._u(pos + 30, 2)
        offb = (max_bits + 7) // 8
        info = {"off_bytes": offb, "len_bytes": (min(max_direct, max_man).bit_length() - 1) // 8 + 1,
                "block_header": 5 + 8 + offb + (4 if hflags & 2 else 0)}
        blocks = []
        if root != UNDEF:
            if nrows == 0:
                blocks.append((0, self.base + root, start))
            else:
                r = self.base + root
                if b[r:r + 4] != b"FHIB":
                    raise H5Error("bad fractal heap indirect block")
                pos = r + 5 + 8 + offb
                off = 0
                for row in range(nrows):
                    size = start if row < 2 else start << (row - 1)
                    if size > max_direct:
                        raise H5Error("fractal heaps with nested indirect blocks are not supported")
                    for _ in range(width):
                        child = self._u(pos, 8)
                        pos += 8
                        if child != UNDEF:
                            blocks.append((off, self.base + child, size))
                        off += size
        info["blocks"] = blocks
        return info

    def _heap_objects(self, heap_addr, btree_addr, id_offset):
        """the managed objects of a fractal heap that a version-2 B-tree's records point at (heap id at id_offset of each record)"""
        b = self.buf
        heap = self._fractal_heap(heap_addr)
        t = self.base + btree_addr
        if b[t:t + 4] != b"BTHD":
            raise H5Error("bad version-2 B-tree header")
        node_size, rec_size, depth = self._u(t + 6, 4), self._u(t + 10, 2), self._u(t + 12, 2)
        root, nroot = self._u(t + 16, 8), self._u(t + 24, 2)
        max_leaf = (node_size - 10) // rec_size
        nsz = (max_leaf.bit_length() + 7) // 8
        records = []

        def node(addr, nrec, level):
            n = self.base + addr
            if b[n:n + 4] != (b"BTIN" if level else b"BTLF"):
                raise H5Error("bad version-2 B-tree node")
            recs = [b[n + 6 + k * rec_size:n + 6 + (k + 1) * rec_size] for k in range(nrec)]
            if level == 0:
                records.extend(recs)
                return
            if level > 1:
                raise H5Error("version-2 B-trees deeper than 1 are not supported")
            pos = n + 6 + nrec * rec_size
            for k in range(nrec + 1):
                node(self._u(pos, 8), self._u(pos + 8, nsz), level - 1)
                pos += 8 + nsz
                if k < nrec:
                    records.append(recs[k])

        if root != UNDEF and nroot:
            node(root, nroot, depth)
        out = []
        for rec in records:
            hid = rec[id_offset:]
            if (hid[0] >> 4) & 3 != 0:
                raise H5Error("huge / tiny fractal heap objects are not supported")
            off = int.from_bytes(hid[1:1 + heap["off_bytes"]], "little")
            ln = int.from_bytes(hid[1 + heap["off_bytes"]:1 + heap["off_bytes"] + heap["len_bytes"]], "little")
            for boff, baddr, bsize in heap["blocks"]:
                if boff <= off < boff + bsize:
                    out.append(b[baddr + off - boff:baddr + off - boff + ln])
                    break
            else:
                raise H5Error("fractal heap object outside the direct blocks")
        return out

    def _links(self, msgs):
        """[(name, header address)] of a version-2 group"""
        out = []
        for t, d in msgs:
            if t == 0x06:
                out.append(self._link(d)[:2])
            elif t == 0x02:
                flags = d[1]
                pos = 2 + (8 if flags & 1 else 0)
                heap, btree = int.from_bytes(d[pos:pos + 8], "little"), int.from_bytes(d[pos + 8:pos + 16], "little")
                if heap != UNDEF:
                    out += [self._link(o)[:2] for o in self._heap_objects(heap, btree, 4)]
        return sorted((n, a) for n, a in out if a is not None)

    def _attributes(self, msgs):
        """{name: value} of the attributes with a plain datatype (numbers, fixed-length strings); others -> None"""
        raw = [d for t, d in msgs if t == 0x0C]
        for t, d in msgs:
            if t == 0x15:
                flags = d[1]
                pos = 2 + (2 if flags & 1 else 0)
                heap, btree = int.from_bytes(d[pos:pos + 8], "little"), int.from_bytes(d[pos + 8:pos + 16], "little")
                if heap != UNDEF:
                    raw += self._heap_objects(heap, btree, 0)
        out = {}
        for d in raw:
            ver = d[0]
            nlen, tlen, slen = (int.from_bytes(d[2 + 2 * i:4 + 2 * i], "little") for i in range(3))
            pos = 9 if ver == 3 else 8
            pad = (lambda n: (n + 7) & ~7) if ver == 1 else (lambda n: n)
            name = bytes(d[pos:pos + nlen]).split(b"\0")[0].decode()
            pos += pad(nlen)
            td = d[pos:pos + tlen]
            pos += pad(tlen)
            sd = d[pos:pos + slen]
            pos += pad(slen)
            try:
                dt = self._dtype(td)
            except H5Error:
                out[name] = None
                continue
            rank = sd[1]
            off = 8 if sd[0] == 1 else 4
            shape = [int.from_bytes(sd[off + 8 * i:off + 8 * i + 8], "little") for i in range(rank)]
            if sd[0] == 2 and sd[3] == 2:                 # null dataspace
                out[name] = np.zeros(0, dt)
                continue
            n = int(np.prod(shape)) if shape else 1
            v = np.frombuffer(bytes(d[pos:pos + n * dt.itemsize]), dt, n).reshape(shape)
            if dt.kind == "S":
                v = v.reshape(-1)[0].split(b"\0")[0].decode() if n == 1 else [x.split(b"\0")[0].decode() for x in v.reshape(-1)]
            out[name] = v
        return out

    def _heap_name(self, heap_addr, off):
        h = self.base + heap_addr
        if self.buf[h:h + 4] != b"HEAP":
            raise H5Error("bad local heap")
        data = self.base + self._u(h + 24, 8)
        end = self.buf.index(b"\0", data + off)
        return self.buf[data + off:end].decode()

    def _group_entries(self, btree, heap):
        """symbol table entries of a group: B-tree version 1 (node type 0) down to the SNOD leaves"""
        b = self.buf
        n = self.base + btree
        if b[n:n + 4] != b"TREE" or b[n + 4] != 0:
            raise H5Error("bad group B-tree node")
        level, used = b[n + 5], self._u(n + 6, 2)
        pos = n + 8 + 16
        out = []
        for k in range(used):
            child = self._u(pos + 8 + k * 16, 8)       # key (8), child (8), key, child, ..., key
            if level > 0:
                out += self._group_entries(child, heap)
            else:
                s = self.base + child
                if b[s:s + 4] != b"SNOD":
                    raise H5Error("bad symbol node")
                for i in range(self._u(s + 6, 2)):
                    e = self._symbol_entry(s + 8 + 40 * i)
                    e["name"] = self._heap_name(heap, e["name_off"])
                    out.append(e)
        return out

    def _walk(self, prefix, header, entry=None):
        if header in self._headers.values() and prefix:
            return                                          # a second hard link to an object already seen (or a cycle)
        self._headers[prefix or "/"] = header
        msgs = self._messages(header)
        types = {t for t, _ in msgs}
        if 0x11 in types or (entry and entry.get("cache") == 1):
            if 0x11 in types:
                d = [m for t, m in msgs if t == 0x11][0]
                btree, heap = int.from_bytes(d[:8], "little"), int.from_bytes(d[8:16], "little")
            else:
                btree, heap = entry["btree"], entry["heap"]
            self._objects[prefix or "/"] = ("group", None)
            for e in self._group_entries(btree, heap):
                self._walk((prefix + "/" + e["name"]).lstrip("/") if prefix else e["name"], e["header"], e)
        elif 0x02 in types or (0x06 in types and 0x08 not in types):
            self._objects[prefix or "/"] = ("group", None)
            for name, child in self._links(msgs):
                self._walk(prefix + "/" + name if prefix else name, child)
        elif 0x08 in types:
            self._objects[prefix] = ("dataset", self._dataset(msgs))

    # ---- datasets
    def _dataset(self, msgs):
        shape = maxshape = dtype = layout = None
        filters = []
        for t, d in msgs:
            if t == 0x01:
                ver, rank, flags = d[0], d[1], d[2]
                off = 8 if ver == 1 else 4
                shape = [int.from_bytes(d[off + 8 * i:off + 8 * i + 8], "little") for i in range(rank)]
                if flags & 1:
                    off += 8 * rank
                    maxshape = [int.from_bytes(d[off + 8 * i:off + 8 * i + 8], "little") for i in range(rank)]
            elif t == 0x03:
                dtype = self._dtype(d)
            elif t == 0x0B:
                filters = self._filters(d)
            elif t == 0x08:
                ver = d[0]
                if ver == 3:
                    cls = d[1]
                    if cls == 0:
                        n = int.from_bytes(d[2:4], "little")
                        layout = ("compact", bytes(d[4:4 + n]))
                    elif cls == 1:
                        layout = ("contiguous", int.from_bytes(d[2:10], "little"), int.from_bytes(d[10:18], "little"))
                    elif cls == 2:
                        nd = d[2]
                        bt = int.from_bytes(d[3:11], "little")
                        dims = [int.from_bytes(d[11 + 4 * i:15 + 4 * i], "little") for i in range(nd)]
                        layout = ("chunked", bt, dims)
                    else:
                        raise H5Error("layout class %d is not supported" % cls)
                elif ver in (1, 2):
                    nd, cls = d[1], d[2]
                    pos = 8
                    addr = None
                    if cls != 0:
                        addr = int.from_bytes(d[pos:pos + 8], "little")
                        pos += 8
                    dims = [int.from_bytes(d[pos + 4 * i:pos + 4 * i + 4], "little") for i in range(nd)]
                    pos += 4 * nd
                    if cls == 1:
                        layout = ("contiguous", addr, None)
                    elif cls == 2:
                        layout = ("chunked", addr, dims + [int.from_bytes(d[pos:pos + 4], "little")])
                    else:
                        n = int.from_bytes(d[pos:pos + 4], "little")
                        layout = ("compact", bytes(d[pos + 4:pos + 4 + n]))
                else:
                    raise H5Error("data layout message version %d is not supported" % ver)
        if shape is None or dtype is None or layout is None:
            raise H5Error("dataset without dataspace, datatype or layout")
        ds = _Dataset(shape, dtype, layout, maxshape)
        ds.filters = filters
        return ds

    @staticmethod
    def _dtype(d):
        cls, bits0, size = d[0] & 15, d[1], int.from_bytes(d[4:8], "little")
        order = ">" if bits0 & 1 else "<"
        if cls == 0:
            return np.dtype("%s%s%d" % (order, "i" if bits0 & 8 else "u", size))
        if cls == 1:
            return np.dtype("%sf%d" % (order, size))
        if cls == 3:
            return np.dtype("S%d" % size)
        raise H5Error("datatype class %d is not supported" % cls)

    @staticmethod
    def _filters(d):
        """[(filter id, client data)] of a filter pipeline message (versions 1 and 2)"""
        ver, n = d[0], d[1]
        pos = 8 if ver == 1 else 2
        out = []
        for _ in range(n):
            fid = int.from_bytes(d[pos:pos + 2], "little")
            pos += 2
            nlen = 0
            if ver == 1 or fid >= 256:
                nlen = int.from_bytes(d[pos:pos + 2], "little")
                pos += 2
            ncd = int.from_bytes(d[pos + 2:pos + 4], "little")
            pos += 4 + ((nlen + 7) & ~7 if ver == 1 else nlen)
            cd = [int.from_bytes(d[pos + 4 * i:pos + 4 * i + 4], "little") for i in range(ncd)]
            pos += 4 * ncd + (4 if ver == 1 and ncd % 2 else 0)
            if fid not in (1, 2, 3):
                raise H5Error("filter %d is not supported (deflate, shuffle and fletcher32 are)" % fid)
            out.append((fid, cd))
        return out

    @staticmethod
    def _unfilter(raw, filters, mask, itemsize):
        """undo the filter pipeline on one chunk (filters are applied in order on write: undone in reverse)"""
        import zlib
        for k in range(len(filters) - 1, -1, -1):
            fid, cd = filters[k]
            if mask >> k & 1:
                continue
            if fid == 3:
                raw = raw[:-4]
            elif fid == 1:
                raw = zlib.decompress(raw)
            elif fid == 2:
                size = cd[0] if cd else itemsize
                n = len(raw) // size
                body = np.frombuffer(raw, np.uint8, n * size).reshape(size, n).T.tobytes()
                raw = body + raw[n * size:]
        return raw

    def _chunks(self, addr, rank):
        """(offsets, address, bytes) of every chunk under a version-1 chunk B-tree node"""
        b = self.buf
        n = self.base + addr
        if b[n:n + 4] != b"TREE" or b[n + 4] != 1:
            raise H5Error("bad chunk B-tree node")
        level, used = b[n + 5], self._u(n + 6, 2)
        keysize = 8 + 8 * (rank + 1)
        pos = n + 8 + 16
        out = []
        for k in range(used):
            kpos = pos + k * (keysize + 8)
            nbytes, mask = self._u(kpos, 4), self._u(kpos + 4, 4)
            offs = [self._u(kpos + 8 + 8 * i, 8) for i in range(rank)]
            child = self._u(kpos + keysize, 8)
            if level > 0:
                out += self._chunks(child, rank)
            else:
                out.append((offs, child, nbytes, mask))
        return out

    def __getitem__(self, path):
        kind, ds = self._objects[path.strip("/")]
        if kind != "dataset":
            raise KeyError("%s is a group" % path)
        n = int(np.prod(ds.shape)) if ds.shape else 1
        if ds.layout[0] == "compact":
            return np.frombuffer(ds.layout[1], ds.dtype, n).reshape(ds.shape).copy()
        if ds.layout[0] == "contiguous":
            if ds.layout[1] == UNDEF or n == 0:
                return np.zeros(ds.shape, ds.dtype)
            return np.frombuffer(self.buf, ds.dtype, n, self.base + ds.layout[1]).reshape(ds.shape).copy()
        _, bt, cdims = ds.layout
        rank = len(ds.shape)
        out = np.zeros(ds.shape, ds.dtype)
        if bt == UNDEF or n == 0:
            return out
        cshape = cdims[:rank]
        for offs, addr, nbytes, mask in self._chunks(bt, rank):
            if ds.filters:
                raw = self._unfilter(self.buf[self.base + addr:self.base + addr + nbytes], ds.filters, mask, ds.dtype.itemsize)
                chunk = np.frombuffer(raw, ds.dtype, int(np.prod(cshape))).reshape(cshape)
            else:
                chunk = np.frombuffer(self.buf, ds.dtype, int(np.prod(cshape)), self.base + addr).reshape(cshape)
            sl_out = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cshape, ds.shape))
            sl_in = tuple(slice(0, s.stop - s.start) for s in sl_out)
            out[sl_out] = chunk[sl_in]
        return out

    def datasets(self):
        return sorted(k for k, (kind, _) in self._objects.items() if kind == "dataset")

    def groups(self):
        return sorted(k for k, (kind, _) in self._objects.items() if kind == "group" and k != "/")

    def attrs(self, path):
        """attributes of a dataset or group ("/" for the root)"""
        return self._attributes(self._messages(self._headers[path.strip("/") or "/"]))

    def shape(self, path):
        return self._objects[path.strip("/")][1].shape

    def __contains__(self, path):
        return path.strip("/") in self._objects


# ================================================================ writer

def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


def _dtype_message(dt):
    dt = np.dtype(dt)
    if dt.kind == "f":
        # IEEE little-endian: class 1 version 1; bit fields: byte order 0, padding 0, mantissa normalisation 2 (implied
        # msb) in bits 4-5, sign location in the second byte
        size = dt.itemsize
        sign = 8 * size - 1
        exp_bits, man_bits = (11, 52) if size == 8 else (8, 23)
        head = struct.pack("<BBBBI", 0x11, 0x20, sign, 0, size)
        props = struct.pack("<HHBBBBI", 0, 8 * size, man_bits, exp_bits, 0, man_bits, (1 << (exp_bits - 1)) - 1)
        return head + props
    if dt.kind in "iu":
        head = struct.pack("<BBBBI", 0x10, 8 if dt.kind == "i" else 0, 0, 0, dt.itemsize)
        return head + struct.pack("<HH", 0, 8 * dt.itemsize)
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0, 0, 0, dt.itemsize)     # null-terminated ASCII
    raise H5Error("cannot write dtype %s" % dt)


def _message(mtype, payload, flags=0):
    payload = _pad8(payload)
    return struct.pack("<HHBBBB", mtype, len(payload), flags, 0, 0, 0) + payload


def _object_header(messages):
    body = b"".join(messages)
    return struct.pack("<BBHII", 1, 0, len(messages), 1, len(body)) + b"\0" * 4 + body


class _Writer:
    LEAF_K, NODE_K = 4, 16          # group leaf node K (2K entries per symbol node), group internal node K

    def __init__(self):
        self.buf = bytearray()

    def alloc(self, data, align=8):
        self.buf += b"\0" * (-len(self.buf) % align)
        addr = len(self.buf)
        self.buf += data
        return addr

    def dataset(self, arr):
        arr = np.ascontiguousarray(arr)
        if arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("<"))
        data = arr.tobytes()
        addr = self.alloc(data) if data else UNDEF
        space = struct.pack("<BBB5x", 1, arr.ndim, 0) + b"".join(struct.pack("<Q", s) for s in arr.shape)
        layout = struct.pack("<BBQQ", 3, 1, addr, len(data))
        # fill value message, byte for byte what the library writes into these files (version 2, allocate early, write
        # if set, defined with size 0)
        fill = struct.pack("<BBBBI", 2, 1, 2, 1, 0)
        msgs = [_message(0x01, space), _message(0x03, _dtype_message(arr.dtype), 1), _message(0x05, fill), _message(0x08, layout)]
        return self.alloc(_object_header(msgs))

    def group(self, children):
        """children: {name: object header address}; returns (header address, btree address, heap address)"""
        names = sorted(children)                         # symbol nodes hold their entries in name order
        heap = bytearray(b"\0" * 8)                       # offset 0: the empty string (key of the left-most B-tree edge)
        offs = {}
        for nme in names:
            offs[nme] = len(heap)
            heap += _pad8(nme.encode() + b"\0")
        free = len(heap)
        heap += struct.pack("<QQ", 1, 16)                 # one free block at the end: next = 1 (none), size 16
        heap_data = self.alloc(bytes(heap))
        heap_addr = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), free, heap_data))
        per = 2 * self.LEAF_K
        leaves = [names[i:i + per] for i in range(0, len(names), per)] or [[]]
        if len(leaves) > 2 * self.NODE_K:
            raise H5Error("too many entries in one group for this writer (%d)" % len(names))
        snods, keys = [], [0]
        for leaf in leaves:
            ent = b""
            for nme in leaf:
                ent += struct.pack("<QQII16x", offs[nme], children[nme], 0, 0)
            ent += b"\0" * (40 * (per - len(leaf)))
            snods.append(self.alloc(b"SNOD" + struct.pack("<BBH", 1, 0, len(leaf)) + ent))
            keys.append(offs[leaf[-1]] if leaf else 0)
        body = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), UNDEF, UNDEF)
        for k in range(len(snods)):
            body += struct.pack("<QQ", keys[k], snods[k])
        body += struct.pack("<Q", keys[-1])
        body += b"\0" * ((2 * self.NODE_K + 1) * 8 + 2 * self.NODE_K * 8 - (len(body) - 24))
        btree = self.alloc(body)
        header = self.alloc(_object_header([_message(0x11, struct.pack("<QQ", btree, heap_addr))]))
        return header, btree, heap_addr


def write(path, datasets):
    """datasets: {"name" or "group/name": array-like}.  One level or several levels of groups."""
    w = _Writer()
    w.buf += b"\0" * 96                                   # superblock, filled in at the end
    tree = {}
    for name, arr in datasets.items():
        parts = [p for p in name.split("/") if p]
        node = tree
        for p in parts[:-1]:
            node = node.setdefault(p, {})
            if not isinstance(node, dict):
                raise H5Error("%s is both a dataset and a group" % p)
        node[parts[-1]] = np.asarray(arr)

    def emit(node):
        children = {}
        for nme, v in node.items():
            children[nme] = emit(v)[0] if isinstance(v, dict) else w.dataset(v)
        return w.group(children)

    header, btree, heap = emit(tree)
    eof = len(w.buf)
    sb = SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, _Writer.LEAF_K, _Writer.NODE_K, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII", 0, header, 1, 0) + struct.pack("<QQ", btree, heap)
    assert len(sb) == 96
    w.buf[:96] = sb
    with open(path, "wb") as f:
        f.write(bytes(w.buf))

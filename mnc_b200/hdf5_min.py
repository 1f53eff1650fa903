"""A small read-only HDF5 parser -- enough to load Caffe's `.caffemodel.h5` weight files without
h5py (`data/scripts/fetch_mnc_model.sh` downloads the trained MNC net
as `mnc_model.caffemodel.h5`; Caffe writes it with `Net::ToHDF5`, caffe-mnc/src/caffe/net.cpp:
920-975: group `data`, one group per layer, one float dataset per parameter blob, named by index).

Implemented from the HDF5 file-format specification (version 1.x structures, what libhdf5 1.8 /
h5py write by default):
  * superblock versions 0 and 1;
  * version-1 object headers (+ continuation blocks);
  * "old style" groups: symbol-table message -> version-1 B-tree of symbol-table nodes + local heap;
  * dataspace messages v1 / v2, datatype classes 0 (integers) and 1 (IEEE floats), little endian;
  * data layout message v3: compact, contiguous, and chunked (version-1 chunk B-tree) with the
    deflate and shuffle filters.
Anything else (superblock v2+, new-style link messages, other datatypes or filters) raises
`NotImplementedError` with the structure's name.  Checked in tests against the reference's own HDF5
test files (caffe-mnc/src/caffe/test/test_data/*.h5, when the tree is present) and against files
assembled byte by byte in the test.
"""
import struct
import zlib

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class _File(object):
    def __init__(self, buf):
        self.b = buf
        pos = 0
        while True:                       # the superblock may sit at 0, 512, 1024, ...
            if buf[pos:pos + 8] == _SIG:
                break
            pos = 512 if pos == 0 else pos * 2
            if pos + 8 > len(buf):
                raise ValueError("not an HDF5 file")
        ver = buf[pos + 8]
        if ver not in (0, 1):
            raise NotImplementedError("HDF5 superblock version %d" % ver)
        self.so = buf[pos + 13]           # size of offsets
        self.sl = buf[pos + 14]           # size of lengths
        p = pos + 24 + (4 if ver == 1 else 0)
        self.base = self.off(p)
        p += 4 * self.so                  # base, free-space, end-of-file, driver-info addresses
        # root group symbol-table entry: link name offset, object header address, cache, scratch
        self.root_header = self.off(p + self.so)

    def off(self, p):
        return int.from_bytes(self.b[p:p + self.so], "little")

    def length(self, p):
        return int.from_bytes(self.b[p:p + self.sl], "little")

    # ------------------------------------------------------------------ object headers
    def messages(self, addr):
        """[(type, flags, bytes)] of a version-1 object header, continuation blocks included."""
        b = self.b
        addr += self.base
        if b[addr] != 1:
            raise NotImplementedError("object header version %d (new-style groups / links)" % b[addr])
        nmsg = struct.unpack_from("<H", b, addr + 2)[0]
        size = struct.unpack_from("<I", b, addr + 8)[0]
        blocks = [(addr + 16, size)]      # messages start 8-byte aligned after the 12-byte prefix
        out = []
        while blocks and len(out) < nmsg:
            p, n = blocks.pop(0)
            end = p + n
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, mflags = struct.unpack_from("<HHB", b, p)
                body = b[p + 8:p + 8 + msize]
                if mtype == 0x10:         # continuation: more messages elsewhere
                    blocks.append((self.off(p + 8) + self.base, self.length(p + 8 + self.so)))
                out.append((mtype, mflags, body))
                p += 8 + msize
        return out

    # ------------------------------------------------------------------ groups
    def _heap_name(self, heap_addr, name_off):
        b = self.b
        heap_addr += self.base
        if b[heap_addr:heap_addr + 4] != b"HEAP":
            raise ValueError("local heap signature missing")
        data = self.off(heap_addr + 8 + 2 * self.sl) + self.base
        start = data + name_off
        end = start
        while b[end] != 0:
            end += 1
        return bytes(b[start:end]).decode("utf-8")

    def _group_btree(self, node_addr, heap_addr, out):
        b = self.b
        a = node_addr + self.base
        sig = bytes(b[a:a + 4])
        if sig == b"SNOD":
            n = struct.unpack_from("<H", b, a + 6)[0]
            p = a + 8
            for _ in range(n):
                out[self._heap_name(heap_addr, self.off(p))] = self.off(p + self.so)
                p += 2 * self.so + 24
            return
        if sig != b"TREE" or b[a + 4] != 0:
            raise ValueError("group B-tree node expected")
        used = struct.unpack_from("<H", b, a + 6)[0]
        p = a + 8 + 2 * self.so           # after the sibling addresses
        for _ in range(used):
            p += self.sl                  # key i
            self._group_btree(self.off(p), heap_addr, out)
            p += self.so

    def members(self, header_addr):
        """{name: object header address} of a group, {} for a non-group."""
        for mtype, _, body in self.messages(header_addr):
            if mtype == 0x11:             # symbol table message: B-tree address, local heap address
                out = {}
                self._group_btree(int.from_bytes(body[:self.so], "little"),
                                  int.from_bytes(body[self.so:2 * self.so], "little"), out)
                return out
            if mtype in (0x02, 0x06):
                raise NotImplementedError("new-style group (link info / link messages)")
        return {}

    # ------------------------------------------------------------------ datasets
    def dataset(self, header_addr):
        shape = dtype = layout = None
        filters = []
        for mtype, _, body in self.messages(header_addr):
            if mtype == 0x01:
                shape = self._dataspace(body)
            elif mtype == 0x03:
                dtype = self._datatype(body)
            elif mtype == 0x08:
                layout = bytes(body)
            elif mtype == 0x0B:
                filters = self._filters(body)
        if shape is None or dtype is None or layout is None:
            return None
        count = int(np.prod(shape)) if shape else 1
        ver, cls = layout[0], layout[1]
        if ver != 3:
            raise NotImplementedError("data layout message version %d" % ver)
        if cls == 0:                      # compact
            n = struct.unpack_from("<H", layout, 2)[0]
            raw = layout[4:4 + n]
        elif cls == 1:                    # contiguous
            addr = int.from_bytes(layout[2:2 + self.so], "little")
            if addr == _UNDEF & ((1 << (8 * self.so)) - 1):
                return np.zeros(shape, dtype=dtype)
            a = addr + self.base
            raw = self.b[a:a + count * dtype.itemsize]
        elif cls == 2:                    # chunked
            return self._chunked(layout, shape, dtype, filters)
        else:
            raise NotImplementedError("data layout class %d" % cls)
        return np.frombuffer(bytes(raw), dtype=dtype, count=count).reshape(shape)

    def _dataspace(self, body):
        ver, rank, flags = body[0], body[1], body[2]
        if ver == 1:
            p = 8
        elif ver == 2:
            if body[3] == 2:              # null dataspace
                return (0,)
            p = 4
        else:
            raise NotImplementedError("dataspace message version %d" % ver)
        return tuple(int.from_bytes(body[p + i * self.sl:p + (i + 1) * self.sl], "little")
                     for i in range(rank))

    @staticmethod
    def _datatype(body):
        cls, bits0 = body[0] & 0x0F, body[1]
        size = struct.unpack_from("<I", body, 4)[0]
        if bits0 & 1:
            raise NotImplementedError("big-endian datatype")
        if cls == 1 and size in (2, 4, 8):
            return np.dtype("<f%d" % size)
        if cls == 0 and size in (1, 2, 4, 8):
            return np.dtype("<%s%d" % ("i" if bits0 & 0x08 else "u", size))
        raise NotImplementedError("datatype class %d size %d" % (cls, size))

    @staticmethod
    def _filters(body):
        ver, n = body[0], body[1]
        p = 8 if ver == 1 else 2
        out = []
        for _ in range(n):
            fid = struct.unpack_from("<H", body, p)[0]
            if ver == 1 or fid >= 256:
                name_len = struct.unpack_from("<H", body, p + 2)[0]
                ncv = struct.unpack_from("<H", body, p + 6)[0]
                p += 8 + ((name_len + 7) // 8 * 8 if ver == 1 else name_len)
            else:
                ncv = struct.unpack_from("<H", body, p + 4)[0]
                p += 6
            cvals = struct.unpack_from("<%dI" % ncv, body, p)
            p += 4 * ncv + (4 if (ver == 1 and ncv % 2) else 0)
            out.append((fid, cvals))
        return out

    def _chunked(self, layout, shape, dtype, filters):
        ndim = layout[2]                  # dataset rank + 1 (the element size is the last "dimension")
        btree = int.from_bytes(layout[3:3 + self.so], "little")
        cdims = struct.unpack_from("<%dI" % ndim, layout, 3 + self.so)[:-1]
        out = np.zeros(shape, dtype=dtype)
        for f in filters:
            if f[0] not in (1, 2):
                raise NotImplementedError("HDF5 filter id %d" % f[0])
        if btree == (1 << (8 * self.so)) - 1:
            return out                    # undefined address: the dataset was never written (fill value 0)
        self._chunk_btree(btree, ndim, cdims, dtype, filters, out)
        return out

    def _chunk_btree(self, node_addr, ndim, cdims, dtype, filters, out):
        b = self.b
        a = node_addr + self.base
        if bytes(b[a:a + 4]) != b"TREE" or b[a + 4] != 1:
            raise ValueError("chunk B-tree node expected")
        level = b[a + 5]
        used = struct.unpack_from("<H", b, a + 6)[0]
        p = a + 8 + 2 * self.so
        key_size = 8 + 8 * ndim
        for _ in range(used):
            nbytes, mask = struct.unpack_from("<II", b, p)
            offs = struct.unpack_from("<%dQ" % ndim, b, p + 8)[:-1]
            child = self.off(p + key_size)
            p += key_size + self.so
            if level > 0:
                self._chunk_btree(child, ndim, cdims, dtype, filters, out)
                continue
            raw = bytes(b[child + self.base:child + self.base + nbytes])
            for i, (fid, cvals) in reversed(list(enumerate(filters))):
                if mask & (1 << i):
                    continue
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:            # shuffle: bytes of equal significance were stored together
                    es = cvals[0] if cvals else dtype.itemsize
                    arr = np.frombuffer(raw, dtype=np.uint8)
                    raw = arr.reshape(es, -1).T.tobytes()
            chunk = np.frombuffer(raw, dtype=dtype, count=int(np.prod(cdims))).reshape(cdims)
            sel = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, out.shape))
            out[sel] = chunk[tuple(slice(0, s.stop - s.start) for s in sel)]


def read_hdf5(path):
    """-> {"/path/to/dataset": ndarray} for every dataset in the file (groups are walked)."""
    with open(path, "rb") as f:
        buf = memoryview(f.read())
    hf = _File(buf)
    out = {}

    def walk(addr, prefix, depth=0):
        if depth > 32:
            raise ValueError("HDF5 group nesting too deep (cycle?)")
        for name, child in hf.members(addr).items():
            arr = hf.dataset(child)
            if arr is not None:
                out[prefix + "/" + name] = arr
            else:
                walk(child, prefix + "/" + name, depth + 1)

    walk(hf.root_header, "")
    return out


def load_caffemodel_h5(path):
    """Caffe `Net::ToHDF5` layout -> {layer name: [blob0, blob1, ...]} (the `data` group only; a
    `diff` group, if present, is ignored).  Layer names containing '/' come back joined."""
    layers = {}
    for key, arr in read_hdf5(path).items():
        parts = key.strip("/").split("/")
        if len(parts) < 3 or parts[0] != "data" or not parts[-1].isdigit():
            continue
        layers.setdefault("/".join(parts[1:-1]), {})[int(parts[-1])] = np.asarray(arr, dtype=np.float32)
    return {name: [blobs[i] for i in sorted(blobs)] for name, blobs in layers.items()}

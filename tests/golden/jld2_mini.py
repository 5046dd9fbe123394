"""Minimal pure-Python reader for the JLD2 ("HDF5-based Julia Data Format") fixtures
under /root/reference/test/inputs.  Only what those files need: superblock v2, v2 object
headers, link / dataspace / datatype / layout(v4 compact+contiguous) / continuation
messages, Float64 and Int64 scalars/arrays, and arrays of 8-byte object references.
Used ONLY by make_golden.py (runs in the build container where /root/reference exists).
"""
import struct
import numpy as np

BASE = 512


class JLD2File:
    def __init__(self, path):
        self.b = open(path, "rb").read()
        assert self.b[BASE:BASE + 8] == b"\x89HDF\r\n\x1a\n"
        root_rel = struct.unpack_from("<Q", self.b, BASE + 36)[0]
        self.root = self._links(BASE + root_rel)

    # -- object header walking -------------------------------------------------
    def _messages(self, addr):
        b = self.b
        assert b[addr:addr + 4] == b"OHDR", (addr, b[addr:addr + 4])
        flags = b[addr + 5]
        p = addr + 6
        if flags & 0x20:
            p += 16
        if flags & 0x10:
            p += 4
        w = 1 << (flags & 3)
        size = int.from_bytes(b[p:p + w], "little")
        p += w
        chunks = [(p, p + size)]
        out = []
        track = bool(flags & 0x04)
        while chunks:
            p, end = chunks.pop(0)
            while p + 4 <= end:
                mtype = b[p]
                msize = struct.unpack_from("<H", b, p + 1)[0]
                p += 4 + (2 if track else 0)
                data = b[p:p + msize]
                if mtype == 0x10:
                    off, ln = struct.unpack_from("<QQ", data, 0)
                    a = BASE + off
                    assert b[a:a + 4] == b"OCHK"
                    chunks.append((a + 4, a + ln - 4))
                elif mtype != 0:
                    out.append((mtype, data))
                p += msize
        return out

    def _links(self, addr):
        links = {}
        for mtype, d in self._messages(addr):
            if mtype != 0x06:
                continue
            flags = d[1]
            p = 2
            if flags & 0x08:
                p += 1
            if flags & 0x04:
                p += 8
            if flags & 0x10:
                p += 1
            w = 1 << (flags & 3)
            n = int.from_bytes(d[p:p + w], "little")
            p += w
            name = d[p:p + n].decode()
            p += n
            links[name] = BASE + struct.unpack_from("<Q", d, p)[0]
        return links

    def keys(self):
        return [k for k in self.root if not k.startswith("_")]

    # -- datasets ---------------------------------------------------------------
    def _dataset(self, addr):
        dims, cls, esize, raw = (), None, None, None
        for mtype, d in self._messages(addr):
            if mtype == 0x01:
                rank = d[1]
                dims = struct.unpack_from("<%dQ" % rank, d, 4) if rank else ()
            elif mtype == 0x03:
                cls = d[0] & 0x0F
                esize = struct.unpack_from("<I", d, 4)[0]
            elif mtype == 0x08:
                lclass = d[1]
                if lclass == 0:
                    n = struct.unpack_from("<H", d, 2)[0]
                    raw = d[4:4 + n]
                elif lclass == 1:
                    off, ln = struct.unpack_from("<QQ", d, 2)
                    raw = self.b[BASE + off:BASE + off + ln]
                else:
                    raise NotImplementedError("layout class %d" % lclass)
        return dims, cls, esize, raw

    def read(self, addr):
        dims, cls, esize, raw = self._dataset(addr)
        n = int(np.prod(dims)) if dims else 1
        if cls == 1 and esize == 8:
            a = np.frombuffer(raw[:8 * n], dtype="<f8").copy()
        elif cls == 0 and esize == 8:
            a = np.frombuffer(raw[:8 * n], dtype="<i8").copy()
        elif cls == 7:  # object references -> nested arrays
            offs = np.frombuffer(raw[:8 * n], dtype="<u8")
            items = [self.read(BASE + int(o)) for o in offs]
            if len(dims) <= 1:
                return items
            return items
        else:
            raise NotImplementedError((cls, esize, dims))
        if not dims:
            return a[0]
        # HDF5 dims are reversed w.r.t. Julia's column-major dims
        return a.reshape(tuple(dims)).T if len(dims) > 1 else a

    def __getitem__(self, name):
        return self.read(self.root[name])

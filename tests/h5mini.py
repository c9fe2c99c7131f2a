"""Minimal HDF5 reader (superblock v0, v1 object headers, v1 group B-trees, local heaps, contiguous
little-endian float32 datasets) -- enough for Keras 2.1 weight files.  TEST HELPER: an independent
Python restatement used to cross-check the engine's C++ reader (csrc/matnet.cu)."""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5File:
    def __init__(self, path):
        self.b = open(path, "rb").read()
        b = self.b
        assert b[:8] == b"\x89HDF\r\n\x1a\n" and b[8] == 0, "superblock v0 expected"
        assert b[13] == 8 and b[14] == 8, "8-byte offsets/lengths expected"
        # 8 sig, 8 version bytes, 2+2 group K, 4 flags, then base, free-space, eof, driver addresses
        self.base = struct.unpack_from("<Q", b, 24)[0]
        root_entry = 24 + 32
        self.root_header = struct.unpack_from("<Q", b, root_entry + 8)[0]

    # -- object headers (version 1) ------------------------------------------------------
    def messages(self, addr):
        b = self.b
        ver, _, nmsg, _refc, hsize = struct.unpack_from("<BBHII", b, addr)
        assert ver == 1
        out = []
        blocks = [(addr + 16, hsize)]
        while blocks and len(out) < nmsg:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and len(out) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", b, pos)
                body = pos + 8
                if mtype == 0x10:                      # continuation
                    off, length = struct.unpack_from("<QQ", b, body)
                    blocks.append((off, length))
                out.append((mtype, body, msize))
                pos = body + msize
        return out

    def symbol_table(self, addr):
        for mtype, body, _ in self.messages(addr):
            if mtype == 0x11:
                return struct.unpack_from("<QQ", self.b, body)
        return None

    def heap_data(self, heap_addr):
        assert self.b[heap_addr:heap_addr + 4] == b"HEAP"
        return struct.unpack_from("<Q", self.b, heap_addr + 24)[0]

    def group_links(self, addr):
        """name -> object header address of every link of the group at `addr`."""
        st = self.symbol_table(addr)
        if st is None:
            return {}
        btree, heap = st
        names = self.heap_data(heap)
        out = {}

        def walk(node):
            b = self.b
            if b[node:node + 4] == b"TREE":
                ntype, level, used = struct.unpack_from("<BBH", b, node + 4)
                assert ntype == 0
                pos = node + 24                         # sig 4, type 1, level 1, entries 2, left 8, right 8
                for i in range(used):
                    child = struct.unpack_from("<Q", b, pos + 8)[0]   # key (8) then child (8)
                    walk(child)
                    pos += 16
            else:
                assert b[node:node + 4] == b"SNOD", b[node:node + 4]
                n = struct.unpack_from("<H", b, node + 6)[0]
                pos = node + 8
                for i in range(n):
                    name_off, hdr = struct.unpack_from("<QQ", b, pos)
                    s = names + name_off
                    e = b.index(b"\0", s)
                    out[b[s:e].decode()] = hdr
                    pos += 40
        walk(btree)
        return out

    def dataset(self, addr):
        b = self.b
        dims, data_addr, size, is_f32 = None, None, None, False
        for mtype, body, msize in self.messages(addr):
            if mtype == 0x01:
                ver, rank, flags = struct.unpack_from("<BBB", b, body)
                off = body + (8 if ver == 1 else 4)
                dims = struct.unpack_from("<%dQ" % rank, b, off)
            elif mtype == 0x03:
                cls_ver, bits0 = b[body], b[body + 1]
                sz = struct.unpack_from("<I", b, body + 4)[0]
                is_f32 = (cls_ver & 0x0F) == 1 and sz == 4 and (bits0 & 1) == 0
            elif mtype == 0x08:
                ver = b[body]
                if ver == 3:
                    cls = b[body + 1]
                    assert cls == 1, "contiguous layout expected"
                    data_addr, size = struct.unpack_from("<QQ", b, body + 2)
                else:
                    raise ValueError("layout version %d" % ver)
        assert dims is not None and data_addr is not None and is_f32
        n = int(np.prod(dims)) if dims else 1
        assert size == 4 * n
        return np.frombuffer(b, dtype="<f4", count=n, offset=self.base + data_addr).reshape(dims).copy()

    def tensors(self):
        """{"layer/kernel:0": array} for every dataset, walking the whole tree."""
        out = {}

        def rec(addr, prefix):
            links = self.group_links(addr)
            if not links:
                return
            for name, hdr in links.items():
                if self.symbol_table(hdr) is not None:
                    rec(hdr, prefix + [name])
                else:
                    out["/".join(prefix + [name])] = self.dataset(hdr)
        rec(self.root_header, [])
        return out

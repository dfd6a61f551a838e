"""Minimal reader for the HDF5 subset Keras 2.3 / h5py 2.10 writes (superblock v0, v1 object headers, old-style groups
with B-tree + local heap, contiguous little-endian float datasets).  h5py is not installed in the build image; the only
consumer is `realistic_arm_limits_model.h5` (reference `env.py:67`, `load_model`), whose Dense weights feed the arm-limit
classifier of `enforce_realistic_human_joint_limits` (`env.py:353-387`)."""
from __future__ import annotations

import struct
from typing import Dict

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Lite:
    def __init__(self, path: str):
        with open(path, "rb") as f:
            self.b = f.read()
        assert self.b[:8] == b"\x89HDF\r\n\x1a\n", "not an HDF5 file"
        ver = self.b[8]
        assert ver == 0, f"superblock version {ver} not supported"
        assert self.b[13] == 8 and self.b[14] == 8, "only 8-byte offsets/lengths"
        # root group symbol table entry follows: base(8) free(8) eof(8) driver(8) at offset 24
        ste = 24 + 32
        self.root_header = struct.unpack_from("<Q", self.b, ste + 8)[0]

    # -- low level -------------------------------------------------------------------------------------------
    def _messages(self, addr: int):
        b = self.b
        ver, _, nmsg, _, hsize = struct.unpack_from("<BBHII", b, addr)
        assert ver == 1, f"object header version {ver} not supported"
        out = []
        blocks = [(addr + 16, hsize)]
        while blocks and len(out) < nmsg:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = struct.unpack_from("<HHB", b, pos)
                data = pos + 8
                if mtype == 0x10:                                  # continuation
                    off, ln = struct.unpack_from("<QQ", b, data)
                    blocks.append((off, ln))
                out.append((mtype, data, msize))
                pos = data + msize
        return out

    def _heap_string(self, heap_addr: int, off: int) -> str:
        assert self.b[heap_addr:heap_addr + 4] == b"HEAP"
        data_addr = struct.unpack_from("<Q", self.b, heap_addr + 24)[0]
        s = data_addr + off
        e = self.b.index(b"\x00", s)
        return self.b[s:e].decode()

    def _group_entries(self, btree: int, heap: int) -> Dict[str, int]:
        b = self.b
        out: Dict[str, int] = {}

        def walk(addr):
            assert b[addr:addr + 4] == b"TREE", "bad B-tree node"
            level, used = struct.unpack_from("<BH", b, addr + 5)
            pos = addr + 24
            for i in range(used):
                child = struct.unpack_from("<Q", b, pos + 8)[0]    # key(8) child(8) key(8) ...
                pos += 16
                if level > 0:
                    walk(child)
                else:
                    assert b[child:child + 4] == b"SNOD"
                    nsym = struct.unpack_from("<H", b, child + 6)[0]
                    for k in range(nsym):
                        e = child + 8 + 40 * k
                        name_off, hdr = struct.unpack_from("<QQ", b, e)
                        out[self._heap_string(heap, name_off)] = hdr
        walk(btree)
        return out

    def children(self, header_addr: int) -> Dict[str, int]:
        for mtype, data, size in self._messages(header_addr):
            if mtype == 0x11:                                      # symbol table message
                btree, heap = struct.unpack_from("<QQ", self.b, data)
                return self._group_entries(btree, heap)
        return {}

    def dataset(self, header_addr: int) -> np.ndarray:
        b = self.b
        dims = None; dtype = None; addr = None; nbytes = None
        for mtype, data, size in self._messages(header_addr):
            if mtype == 0x01:                                      # dataspace
                ver, rank, flags = struct.unpack_from("<BBB", b, data)
                off = data + (8 if ver == 1 else 4)
                dims = struct.unpack_from("<" + "Q" * rank, b, off)
            elif mtype == 0x03:                                    # datatype
                cls = b[data] & 0x0F
                tsize = struct.unpack_from("<I", b, data + 4)[0]
                assert cls == 1, "only floating-point datasets"
                dtype = {4: "<f4", 8: "<f8"}[tsize]
            elif mtype == 0x08:                                    # layout
                ver = b[data]
                assert ver == 3, f"layout version {ver} not supported"
                lclass = b[data + 1]
                if lclass == 1:                                    # contiguous
                    addr, nbytes = struct.unpack_from("<QQ", b, data + 2)
                elif lclass == 0:                                  # compact
                    nbytes = struct.unpack_from("<H", b, data + 2)[0]
                    addr = data + 4
                else:
                    raise ValueError("chunked datasets are not supported")
        assert dims is not None and dtype is not None and addr is not None
        n = int(np.prod(dims)) if dims else 1
        return np.frombuffer(b, dtype=dtype, count=n, offset=addr).reshape(dims).copy()

    def get(self, path: str) -> int:
        addr = self.root_header
        for part in [p for p in path.split("/") if p]:
            ch = self.children(addr)
            if part not in ch:
                raise KeyError(f"{part} not in {sorted(ch)}")
            addr = ch[part]
        return addr


def load_keras_dense_stack(path: str):
    """-> [(kernel[in,out], bias[out]), ...] of a Sequential Dense model saved by Keras (layers dense_1..dense_n)."""
    h = H5Lite(path)
    mw = h.get("model_weights")
    layers = sorted((n for n in h.children(mw) if n.startswith("dense")), key=lambda s: int(s.split("_")[1]))
    out = []
    for name in layers:
        g = h.children(h.children(mw)[name])
        inner = h.children(g[name]) if name in g else g
        k = h.dataset(inner["kernel:0"]); bb = h.dataset(inner["bias:0"])
        out.append((k.astype(np.float32), bb.astype(np.float32)))
    return out

"""Helpers for the reference-built seqset fixtures (tests/golden/ref_seqsets.npz, made by
tests/golden/make_ref_seqsets.py): member access, varbit decode, and reconstruction of the entry
sequences from the tables alone (the seqset is an FM-index-like structure: entry i with first base
b pops to the (i - fixed[b])-th set bit of prev_b, modules/bio_base/seqset.h)."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_Z = None


def fixture():
    global _Z
    if _Z is None:
        _Z = np.load(os.path.join(ROOT, "tests", "golden", "ref_seqsets.npz"))
    return _Z


def names():
    return [str(x) for x in fixture()["names"]]


def member(name, fn):
    return fixture()[f"{name}|{fn}"].tobytes()


class SpiralZip:
    """Member access by offset, the way the reference reads a spiral file (unzGetCurrentFileZStreamPos64,
    modules/io/spiral_file_mmap.cpp:82-125).  Array members carry a CRC field of 0 ("don't bother to fill
    crc", :421-423), so zipfile.read() -- which verifies the CRC -- only works for the JSON members."""

    def __init__(self, path):
        import struct
        import zipfile
        self.path = str(path)
        self.info = {i.filename: i for i in zipfile.ZipFile(self.path).infolist()}
        self.order = [i.filename for i in zipfile.ZipFile(self.path).infolist()]
        self._struct = struct

    def namelist(self):
        return list(self.order)

    def read(self, name):
        i = self.info[name]
        with open(self.path, "rb") as f:
            f.seek(i.header_offset)
            h = f.read(30)
            n, e = self._struct.unpack("<HH", h[26:30])
            f.seek(i.header_offset + 30 + n + e)
            return f.read(i.file_size)

    def crc_ok(self, name):
        import zlib
        return (zlib.crc32(self.read(name)) & 0xffffffff) == self.info[name].CRC


def varbit_decode(elements, bits, n):
    w = np.frombuffer(elements, dtype="<u8")
    if bits == 8:
        return np.frombuffer(elements, dtype=np.uint8)[:n].astype(np.uint16)
    idx = np.arange(n, dtype=np.uint64) * np.uint64(bits)
    lo = w[(idx >> np.uint64(6)).astype(np.int64)] >> (idx & np.uint64(63))
    nxt = np.minimum((idx >> np.uint64(6)).astype(np.int64) + 1, len(w) - 1)
    sh = (np.uint64(64) - (idx & np.uint64(63))) & np.uint64(63)
    hi = np.where((idx & np.uint64(63)) + np.uint64(bits) > 64, w[nxt] << sh, np.uint64(0))
    return ((lo | hi) & np.uint64((1 << bits) - 1)).astype(np.uint16)


def tables(name):
    n = json.loads(member(name, "seqset.json"))["num_entries"]
    meta_s = json.loads(member(name, "entry_sizes/packed_varbit_vector.json"))
    meta_h = json.loads(member(name, "shared/packed_varbit_vector.json"))
    sizes = varbit_decode(member(name, "entry_sizes/elements"), meta_s["bits_per_value"], n)
    shared = varbit_decode(member(name, "shared/elements"), meta_h["bits_per_value"], n)
    fixed = np.frombuffer(member(name, "fixed"), dtype="<u8")
    prev = np.stack([np.frombuffer(member(name, f"prev_{b}/bits"), dtype="<u8") for b in "ACGT"])
    return {"n": n, "sizes": sizes, "shared": shared, "fixed": fixed, "prev": prev, "meta_sizes": meta_s,
            "meta_shared": meta_h}


def entries_ascii(t):
    """[n, max_len] uint8 ASCII matrix (0 padded) of the entry sequences, from the tables alone."""
    n = t["n"]
    fixed = t["fixed"].astype(np.int64)
    first = np.zeros(n, dtype=np.int64)
    pop = np.zeros(n, dtype=np.int64)
    for b in range(4):
        lo, hi = fixed[b], fixed[b + 1]
        first[lo:hi] = b
        bits = np.unpackbits(t["prev"][b].view(np.uint8), bitorder="little")[:n]
        ones = np.flatnonzero(bits)
        assert len(ones) == hi - lo
        pop[lo:hi] = ones
    sizes = t["sizes"].astype(np.int64)
    maxlen = int(sizes.max())
    out = np.zeros((n, maxlen), dtype=np.uint8)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    idx = np.arange(n)
    for c in range(maxlen):
        live = sizes > c
        out[live, c] = acgt[first[idx[live]]]
        idx = pop[idx]
    return out, sizes


def as_reads(t):
    """(uint8 buffer, int64 offsets) of all entries, for add_reads"""
    mat, sizes = entries_ascii(t)
    mask = np.arange(mat.shape[1])[None, :] < sizes[:, None]
    buf = mat[mask]
    offs = np.zeros(len(sizes) + 1, dtype=np.int64)
    np.cumsum(sizes, out=offs[1:])
    return np.ascontiguousarray(buf), offs

"""ctypes binding of oracle/_ref/libref.so: the REFERENCE'S OWN classes for the hot path (kmer_counter, kmer_set,
fast_read_correct, build_seqset::correct_reads, part_repo, expander, builder, seqset), compiled from the sources where
they lie under /root/reference by `make -C oracle _ref` (see oracle/ref_shim.cpp for what is and is not the
reference's).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's CPU arms may import this; the product
(biograph_b200/, include/) never does.  The library is built in the development container (where /root/reference
exists) and travels to the GPU box as a built file; `available()` says whether it is there.
"""
import ctypes as C
import os
import shutil
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libref.so")
_lib = None


def available():
    """the library was built (development container) and loads here"""
    if not os.path.exists(LIB):
        return False
    try:
        lib()
        return True
    except OSError:
        return False


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB)
        L.ref_last_error.restype = C.c_char_p
        L.ref_open.restype = C.c_void_p
        L.ref_open.argtypes = [C.c_char_p, C.c_int, C.c_uint64]
        L.ref_close.argtypes = [C.c_void_p]
        L.ref_fast_read_correct.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int,
                                            C.c_char_p, C.POINTER(C.c_int)]
        L.ref_count_kmers.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_uint64,
                                      C.c_int, C.c_uint64]
        L.ref_counts.restype = C.c_int64
        L.ref_counts.argtypes = [C.c_void_p] + [C.POINTER(C.c_void_p)] * 4
        L.ref_solid_size.restype = C.c_int64
        L.ref_solid_size.argtypes = [C.c_void_p]
        L.ref_solid.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_correct.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_double, C.c_int]
        L.ref_corrected.restype = C.c_int64
        L.ref_corrected.argtypes = [C.c_void_p] + [C.POINTER(C.c_void_p)] * 3
        L.ref_seed.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int]
        L.ref_make_seqset.argtypes = [C.c_void_p]
        L.ref_make_seqset_file.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_seqset_size.restype = C.c_int64
        L.ref_seqset_size.argtypes = [C.c_void_p]
        L.ref_seqset_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.ref_make_readmap.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int]
        L.ref_open_seqset_file.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_seqset_uuid.restype = C.c_char_p
        L.ref_seqset_uuid.argtypes = [C.c_void_p]
        L.ref_readmap_rows.restype = C.c_int64
        L.ref_readmap_rows.argtypes = [C.c_char_p]
        L.ref_read_readmap_file.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p]
        L.ref_read_fastq.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_char_p,
                                     C.c_size_t]
        L.ref_biograph_dir.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p]
        L.ref_open_biograph.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_size_t]
        L.ref_merge.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ref_fast_migrate.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_char_p]
        L.ref_mergemap.restype = C.c_int64
        L.ref_mergemap.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        L.ref_flat.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
        L.ref_free.argtypes = [C.c_void_p]
        L.ref_members.restype = C.c_int64
        L.ref_members.argtypes = [C.c_void_p]
        L.ref_member_name.restype = C.c_char_p
        L.ref_member_name.argtypes = [C.c_void_p, C.c_int64]
        L.ref_member_data.restype = C.c_int64
        L.ref_member_data.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_void_p)]
        _lib = L
    return _lib


def _view(ptr, n, dtype):
    n = int(n)
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.frombuffer((C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr.value), dtype=dtype).copy()


def _pack(reads):
    if isinstance(reads, tuple):
        buf, offs = reads
        return (buf if isinstance(buf, bytes) else bytes(buf)), np.ascontiguousarray(offs, dtype=np.int64)
    bs = [r.encode() if isinstance(r, str) else bytes(r) for r in reads]
    offs = np.zeros(len(bs) + 1, dtype=np.int64)
    if bs:
        np.cumsum([len(b) for b in bs], out=offs[1:])
    return b"".join(bs), offs


def _zip_members(path):
    """{member path: bytes} of a spiral file (stored zip members, read by offset: the reference leaves the CRC fields unset)"""
    import struct
    import zipfile
    raw = open(path, "rb").read()
    out = {}
    for info in zipfile.ZipFile(path).infolist():
        o = info.header_offset
        sig, _, _, comp, _, _, _, _, _, nl, el = struct.unpack("<IHHHHHIIIHH", raw[o:o + 30])
        assert sig == 0x04034B50 and comp == 0
        out[info.filename] = raw[o + 30 + nl + el:o + 30 + nl + el + info.file_size]
    return out


def fast_read_correct(read, solid_kmers, k, max_corrections=2, min_good_run=2):
    """modules/bio_base/fast_read_correct.cpp:94-182 itself; same signature as oracle.fast_read_correct."""
    solid = np.ascontiguousarray(np.sort(np.asarray(solid_kmers, dtype=np.uint64)))
    rb = read.encode() if isinstance(read, str) else read
    out = C.create_string_buffer(max(1, len(rb)))
    corr = C.c_int(0)
    n = lib().ref_fast_read_correct(rb, len(rb), solid.ctypes.data, len(solid), k, max_corrections, min_good_run, out,
                                    C.byref(corr))
    if n < 0:
        raise RuntimeError(lib().ref_last_error().decode())
    return out.raw[:n].decode(), corr.value


class Run:
    """One pass of the reference's create flow; the stages can be called one by one."""

    def __init__(self, threads=0, max_mem_bytes=0, tmp_dir=None):
        self.tmp = tempfile.mkdtemp(prefix="bgx_ref_", dir=tmp_dir)
        self.h = lib().ref_open(self.tmp.encode(), threads or (os.cpu_count() or 1), max_mem_bytes)

    def close(self):
        if self.h:
            lib().ref_close(self.h)
            self.h = None
        shutil.rmtree(self.tmp, ignore_errors=True)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc:
            raise RuntimeError("reference: " + lib().ref_last_error().decode())

    def count_kmers(self, reads, k=30, min_count=5, counter_max_memory_bytes=0, force_exact_passes=0, genome_bases=0):
        """kmer_counter's two-stage count, then the solid kmer_set.  Returns (counts, solid): counts = every element
        extract_exact_counts yields, sorted by k-mer (k-mers the probabilistic pass filtered may be absent); solid =
        the kmer_set (ascending) with flag bits (bit0 fwd_starts_read, bit1 rev_starts_read).  genome_bases: the size
        of the --ref genome, which bounds the probabilistic table (0: the bases read)."""
        buf, offs = _pack(reads)
        self._ck(lib().ref_count_kmers(self.h, buf, offs.ctypes.data, len(offs) - 1, k, min_count,
                                       counter_max_memory_bytes, force_exact_passes, genome_bases))
        p = [C.c_void_p() for _ in range(4)]
        n = lib().ref_counts(self.h, *[C.byref(x) for x in p])
        kmers, fwd, rev, flags = (_view(p[0], n, np.uint64), _view(p[1], n, np.uint32), _view(p[2], n, np.uint32),
                                  _view(p[3], n, np.uint8))
        o = np.argsort(kmers, kind="stable")
        counts = {"kmers": kmers[o], "fwd": fwd[o], "rev": rev[o], "flags": flags[o]}
        ns = lib().ref_solid_size(self.h)
        sk = np.zeros(ns, dtype=np.uint64)
        sf = np.zeros(ns, dtype=np.uint8)
        self._ck(lib().ref_solid(self.h, sk.ctypes.data, sf.ctypes.data))
        return counts, {"kmers": sk, "flags": sf}

    def correct(self, reads, max_corrections=8, min_good_run=2, trim_after_portion=0.7, partition_depth=0):
        """build_seqset::correct_reads::correct over every read (seeds go to the part_repo).  Returns
        dict(seq, offs, kept) like oracle.correct_reads."""
        buf, offs = _pack(reads)
        self._ck(lib().ref_correct(self.h, buf, offs.ctypes.data, len(offs) - 1, max_corrections, min_good_run,
                                   trim_after_portion, partition_depth))
        p = [C.c_void_p() for _ in range(3)]
        n = lib().ref_corrected(self.h, *[C.byref(x) for x in p])
        o = _view(p[1], n + 1, np.int64)
        return {"seq": _view(p[0], o[n], np.uint8).tobytes(), "offs": o, "kept": _view(p[2], n, np.uint8)}

    def seed(self, reads, next_fwd=None, next_rev=None, partition_depth=2):
        buf, offs = _pack(reads)
        nf = None if next_fwd is None else np.ascontiguousarray(next_fwd, dtype=np.int32)
        nr = None if next_rev is None else np.ascontiguousarray(next_rev, dtype=np.int32)
        self._ck(lib().ref_seed(self.h, buf, offs.ctypes.data, len(offs) - 1, None if nf is None else nf.ctypes.data,
                                None if nr is None else nr.ctypes.data, partition_depth))

    def make_seqset(self, path=None):
        """expander x 4 + builder (SEQSETMain::make_seqset).  Same dict as oracle.seqset_staged.  path: the seqset
        is written there by the reference's own spiral_file_create_mmap (as <out>.bg/seqset) and reopened."""
        if path is None:
            self._ck(lib().ref_make_seqset(self.h))
        else:
            self._ck(lib().ref_make_seqset_file(self.h, os.fsencode(path)))
        return self.seqset_tables()

    def open_seqset_file(self, path):
        """a seqset spiral file through the reference's own reader (spiral_file_open_mmap + seqset); returns its
        tables.  flat() then walks every entry's sequence through the reference's seqset_flat."""
        self._ck(lib().ref_open_seqset_file(self.h, os.fsencode(path)))
        return self.seqset_tables()

    def seqset_uuid(self):
        return lib().ref_seqset_uuid(self.h).decode()

    def merge_from(self, runs):
        """`biograph merge`'s seqset path over the seqsets of `runs` (each has done make_seqset) into this run:
        seqset_flat_builder, make_mergemap, seqset_mergemap, seqset_merger.  Returns (merged tables, [mergemap bit
        words per input])."""
        arr = (C.c_void_p * len(runs))(*[r.h for r in runs])
        self._ck(lib().ref_merge(arr, len(runs), self.h))
        maps = []
        for i in range(len(runs)):
            p = C.c_void_p()
            nw = lib().ref_mergemap(self.h, i, C.byref(p))
            maps.append(_view(p, nw, np.uint64))
        return self.seqset_tables(), maps

    def flat(self):
        """seqset_flat over this run's seqset: the sequence of every entry (list of bytes)"""
        ps, po, n = C.c_void_p(), C.c_void_p(), C.c_int64()
        self._ck(lib().ref_flat(self.h, C.byref(ps), C.byref(po), C.byref(n)))
        offs = _view(po, n.value + 1, np.int64)
        seq = _view(ps, offs[-1], np.uint8).tobytes()
        lib().ref_free(ps)
        lib().ref_free(po)
        return [seq[offs[i]:offs[i + 1]] for i in range(n.value)]

    def seqset_tables(self):
        n = lib().ref_seqset_size(self.h)
        words = (n + 63) // 64
        sizes = np.zeros(n, dtype=np.uint16)
        shared = np.zeros(n, dtype=np.uint16)
        prev = np.zeros((4, words), dtype=np.uint64)
        fixed = np.zeros(5, dtype=np.uint64)
        stats = np.zeros(6, dtype=np.int64)
        self._ck(lib().ref_seqset_tables(self.h, sizes.ctypes.data, shared.ctypes.data, prev.ctypes.data,
                                         fixed.ctypes.data, stats.ctypes.data))
        return {"n": int(n), "sizes": sizes, "shared": shared, "prev": prev, "fixed": fixed, "stats": stats}

    def fast_migrate(self, input_index, old_readmap_path, new_readmap_path):
        """make_readmap::fast_migrate: the readmap file of merge input `input_index` onto this (merged) run's seqset;
        returns {member path: bytes} of the new file"""
        self._ck(lib().ref_fast_migrate(self.h, input_index, os.fsencode(old_readmap_path), os.fsencode(new_readmap_path)))
        return _zip_members(new_readmap_path)

    def make_readmap(self, reads, rec_offs, is_paired, keep_path=None):
        """make_readmap::do_make over the seqset just built (call before members()).  reads: the corrected reads;
        rec_offs[n_rec + 1]: record r holds reads [rec_offs[r], rec_offs[r + 1]) -- one read or two mates, as the
        corrected_reads stream holds them.  Returns {member path: bytes} of the readmap spiral file the reference
        wrote (stored zip members, read by offset: the reference leaves the CRC fields unset)."""
        buf, offs = _pack(reads)
        ro = np.ascontiguousarray(rec_offs, dtype=np.int64)
        path = keep_path or os.path.join(self.tmp, "ref.readmap")
        if os.path.exists(path):
            os.unlink(path)
        self._ck(lib().ref_make_readmap(self.h, os.fsencode(path), buf, offs.ctypes.data, ro.ctypes.data, len(ro) - 1,
                                        1 if is_paired else 0))
        return _zip_members(path)

    def members(self):
        """{member path: bytes} of the in-memory seqset spiral file, as the reference's encoders wrote them."""
        n = lib().ref_members(self.h)
        if n < 0:
            raise RuntimeError("reference: " + lib().ref_last_error().decode())
        out = {}
        for i in range(n):
            p = C.c_void_p()
            sz = lib().ref_member_data(self.h, i, C.byref(p))
            out[lib().ref_member_name(self.h, i).decode()] = _view(p, sz, np.uint8).tobytes()
        return out


def write_biograph(path, accession_id, reads, rec_offs=None, is_paired=False, threads=2):
    """A BioGraph directory as the reference writes one, from reads taken as already corrected: biograph_dir creates
    the layout, the builder's seqset goes to <path>/seqset through spiral_file_create_mmap, make_readmap::do_make to
    <path>/coverage/<sha1>.readmap, biograph_dir::save_metadata writes metadata/bg_info.json.  Returns the seqset tables."""
    import hashlib
    if lib().ref_biograph_dir(os.fsencode(path), accession_id.encode(), b""):
        raise RuntimeError("reference: " + lib().ref_last_error().decode())
    with Run(threads) as r:
        r.seed(reads)
        tables = r.make_seqset(os.path.join(path, "seqset"))
        tmp = os.path.join(path, "coverage", "tmp.readmap")
        r.make_readmap(reads, rec_offs if rec_offs is not None else list(range(len(reads) + 1)), is_paired, keep_path=tmp)
    sha = hashlib.sha1(open(tmp, "rb").read()).hexdigest()
    os.rename(tmp, os.path.join(path, "coverage", sha + ".readmap"))
    if lib().ref_biograph_dir(os.fsencode(path), accession_id.encode(), sha.encode()):
        raise RuntimeError("reference: " + lib().ref_last_error().decode())
    return tables


def read_fastq(path):
    """fastq_reader::read over a file through the reference's own file_reader: (reads accepted before the end or the
    first error, the io_exception's text or "")"""
    pb, po, n = C.c_void_p(), C.c_void_p(), C.c_int64()
    err = C.create_string_buffer(512)
    if lib().ref_read_fastq(os.fsencode(path), C.byref(pb), C.byref(po), C.byref(n), err, len(err)):
        raise RuntimeError("reference: " + lib().ref_last_error().decode())
    offs = _view(po, n.value + 1, np.int64)
    seq = _view(pb, offs[-1], np.uint8).tobytes().decode()
    lib().ref_free(pb)
    lib().ref_free(po)
    return [seq[offs[i]:offs[i + 1]] for i in range(n.value)], err.value.decode()


def open_biograph(path, sample=""):
    """a whole .bg directory through the reference's own biograph_dir / seqset / readmap (as its consumers open one):
    dict of what it sees (ids, entry and row counts, pair statistics).  sample: accession id or readmap id when the
    BioGraph holds more than one."""
    buf = C.create_string_buffer(1 << 14)
    if lib().ref_open_biograph(os.fsencode(path), sample.encode(), buf, len(buf)):
        raise RuntimeError("reference: " + lib().ref_last_error().decode())
    out = {}
    for line in buf.value.decode().splitlines():
        k, _, v = line.partition("=")
        out[k] = int(v) if v.isdigit() else v
    return out


def read_readmap_file(path):
    """a readmap spiral file through the reference's own reader (readmap::open_anonymous_readmap) and accessors:
    dict(entry_id, read_lengths, is_forward, mate_loop_ptr, seqset_uuid), one element per row"""
    p = os.fsencode(path)
    n = lib().ref_readmap_rows(p)
    if n < 0:
        raise RuntimeError("reference: " + lib().ref_last_error().decode())
    entry = np.zeros(n, dtype=np.uint64)
    ln = np.zeros(n, dtype=np.int32)
    fwd = np.zeros(n, dtype=np.uint8)
    ptr = np.zeros(n, dtype=np.uint64)
    uuid = C.create_string_buffer(64)
    if lib().ref_read_readmap_file(p, entry.ctypes.data, ln.ctypes.data, fwd.ctypes.data, ptr.ctypes.data, uuid):
        raise RuntimeError("reference: " + lib().ref_last_error().decode())
    return {"entry_id": entry, "read_lengths": ln, "is_forward": fwd, "mate_loop_ptr": ptr, "seqset_uuid": uuid.value.decode()}


def seqset_for_reads(reads, next_fwd=None, next_rev=None, threads=0, partition_depth=2):
    with Run(threads) as r:
        r.seed(reads, next_fwd, next_rev, partition_depth)
        return r.make_seqset()


def create(reads, k=30, min_count=5, max_corrections=8, min_good_run=2, trim_after_portion=0.7, threads=0,
           with_members=False):
    """The whole path on the reference's classes: reads -> (counts, solid, corrected, seqset[, members])."""
    with Run(threads) as r:
        counts, solid = r.count_kmers(reads, k, min_count)
        cr = r.correct(reads, max_corrections, min_good_run, trim_after_portion)
        ss = r.make_seqset()
        out = {"counts": counts, "solid": solid, "corrected": cr, "seqset": ss}
        if with_members:
            out["members"] = r.members()
        return out

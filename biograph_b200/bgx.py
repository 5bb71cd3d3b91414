"""ctypes binding of include/bgx.h (the C ABI of libbgx.so).

Mirrors the reference's stage interfaces for this path: add reads ->
count k-mers -> correct -> build seqset -> export tables.  Everything computes on the GPU; the
library refuses to create a context without a CUDA device."""
import ctypes as C
import json
import os
import subprocess
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class BgxError(RuntimeError):
    pass


class Options(C.Structure):
    _fields_ = [("kmer_size", C.c_int32), ("min_kmer_count", C.c_int32), ("max_corrections", C.c_int32),
                ("min_good_run", C.c_int32), ("trim_after_portion", C.c_float), ("device", C.c_int32),
                ("sort_key_bits", C.c_int32), ("count_batch_reads", C.c_int32)]


def lib_path():
    return os.path.join(_HERE, "libbgx.so")


def build_library(force=False):
    """Compile libbgx.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    args = ["make", "-C", os.path.join(_HERE, "csrc"), "-s", "-j8"]
    if force:
        subprocess.check_call(args + ["clean"])
    subprocess.check_call(args)
    return lib_path()


EXPORTS = ["bgx_default_options", "bgx_last_error", "bgx_version", "bgx_device_count", "bgx_create", "bgx_destroy",
           "bgx_free", "bgx_add_reads_ascii", "bgx_add_reads_fastq", "bgx_add_reads_packed", "bgx_add_reads_packed_async", "bgx_count_kmers", "bgx_export_kmers",
           "bgx_correct", "bgx_export_corrected", "bgx_export_reads", "bgx_build_seqset", "bgx_export_seqset",
           "bgx_export_entries_ascii", "bgx_lookup_reads", "bgx_build_readmap", "bgx_run", "bgx_reset_results", "bgx_clear_reads", "bgx_stats_json", "bgx_timer_start", "bgx_timer_stop",
           "bgx_launch_count", "bgx_debug_khash", "bgx_debug_count_plan", "bgx_debug_sort_pairs", "bgx_dist_unique_id", "bgx_dist_init", "bgx_seqset_layout", "bgx_seed_uncorrected", "bgx_export_varbit",
           "bgx_merge_seqsets", "bgx_export_mergemap", "bgx_migrate_bits", "bgx_export_flat_ascii"]


class SeqsetPart(C.Structure):
    """bgx_seqset_part (include/bgx.h): one input of a merge"""
    _fields_ = [("n_entries", C.c_uint64), ("sizes", C.c_void_p), ("prev_bits", C.c_void_p * 4)]


def load_library():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise BgxError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)")
    L = C.CDLL(path)
    vp, u64p = C.c_void_p, C.POINTER(C.c_uint64)
    L.bgx_last_error.restype = C.c_char_p
    L.bgx_version.restype = C.c_char_p
    L.bgx_default_options.argtypes = [C.POINTER(Options)]
    L.bgx_create.argtypes = [C.POINTER(Options), C.POINTER(vp)]
    L.bgx_destroy.argtypes = [vp]
    L.bgx_free.argtypes = [vp]
    L.bgx_add_reads_ascii.argtypes = [vp, vp, vp, C.c_uint64]
    L.bgx_add_reads_fastq.argtypes = [vp, C.c_char_p, C.c_uint64, u64p]
    L.bgx_add_reads_packed.argtypes = [vp, vp, vp, vp, vp, C.c_uint64]
    L.bgx_add_reads_packed_async.argtypes = [vp, vp, vp, vp, vp, C.c_uint64]
    L.bgx_count_kmers.argtypes = [vp]
    L.bgx_export_kmers.argtypes = [vp, C.c_uint32, u64p, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.bgx_correct.argtypes = [vp]
    L.bgx_export_corrected.argtypes = [vp, u64p, C.POINTER(vp), C.POINTER(vp), u64p, C.POINTER(vp), C.POINTER(vp),
                                       C.POINTER(vp)]
    L.bgx_export_reads.argtypes = [vp, u64p, C.POINTER(vp), C.POINTER(vp), u64p]
    L.bgx_build_seqset.argtypes = [vp]
    L.bgx_export_seqset.argtypes = [vp, u64p, C.POINTER(C.c_uint32), C.POINTER(vp), C.POINTER(vp), vp * 4, vp * 4,
                                    vp * 4, C.c_uint64 * 5]
    L.bgx_export_entries_ascii.argtypes = [vp, C.c_uint64, C.c_uint64, C.POINTER(vp), C.POINTER(vp)]
    L.bgx_lookup_reads.argtypes = [vp, u64p, C.POINTER(vp), C.POINTER(vp)]
    L.bgx_build_readmap.argtypes = [vp, C.c_int32, u64p, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp * 3),
                                             C.POINTER(vp * 3)]
    L.bgx_run.argtypes = [vp]
    L.bgx_reset_results.argtypes = [vp]
    L.bgx_clear_reads.argtypes = [vp]
    L.bgx_stats_json.argtypes = [vp, C.c_char_p, C.c_size_t]
    L.bgx_timer_start.argtypes = [vp]
    L.bgx_timer_stop.argtypes = [vp, C.POINTER(C.c_double)]
    L.bgx_launch_count.restype = C.c_uint64
    L.bgx_debug_sort_pairs.argtypes = [vp, vp, vp, C.c_uint64, C.c_int, C.c_int]
    L.bgx_debug_khash.restype = C.c_uint64
    L.bgx_debug_khash.argtypes = [C.c_uint64, C.c_int32, C.c_int32]
    L.bgx_debug_count_plan.restype = None
    L.bgx_debug_count_plan.argtypes = [C.c_uint64, C.c_uint64, C.c_int32, C.c_uint64, C.c_uint64, C.c_uint64, u64p,
                                       C.POINTER(C.c_int32)]
    L.bgx_dist_unique_id.argtypes = [vp]
    L.bgx_dist_init.argtypes = [vp, C.c_int32, C.c_int32, vp]
    L.bgx_seqset_layout.argtypes = [vp, C.c_uint64 * 6]
    L.bgx_seed_uncorrected.argtypes = [vp]
    L.bgx_export_varbit.argtypes = [vp, C.c_int32, C.POINTER(vp), u64p, C.POINTER(C.c_uint32), u64p]
    L.bgx_merge_seqsets.argtypes = [vp, C.POINTER(SeqsetPart), C.c_uint32, C.c_uint64]
    L.bgx_export_mergemap.argtypes = [vp, C.c_uint32, C.POINTER(vp * 3), u64p, u64p]
    L.bgx_migrate_bits.argtypes = [vp, C.c_uint32, vp, C.c_uint64, C.POINTER(vp * 3), u64p]
    L.bgx_export_flat_ascii.argtypes = [vp, C.c_uint32, C.c_uint64, C.c_uint64, C.POINTER(vp), C.POINTER(vp)]
    _LIB = L
    return L


def pack_reads_2bit(reads):
    """Host-side packing into the bgx_add_reads_packed layout (dna_sequence byte order, every read
    on an 8-byte boundary) + N mask.  reads: list of str/bytes or (buffer, offsets) pair.
    Returns (packed uint8[8*W], nmask uint32[W] or None, word_offs uint64[n+1], lens uint16[n])."""
    if isinstance(reads, np.ndarray) and reads.ndim == 2:
        return _pack_fixed_len(reads)
    if isinstance(reads, tuple):
        buf, offs = reads
        offs = np.asarray(offs, dtype=np.int64)
        arr = np.frombuffer(buf, dtype=np.uint8)
    else:
        bs = [r.encode() if isinstance(r, str) else bytes(r) for r in reads]
        offs = np.zeros(len(bs) + 1, dtype=np.int64)
        if bs:
            np.cumsum([len(b) for b in bs], out=offs[1:])
        arr = np.frombuffer(b"".join(bs), dtype=np.uint8)
    n = len(offs) - 1
    lens = np.diff(offs).astype(np.int64)
    nwords = (lens + 31) // 32
    word_offs = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(nwords, out=word_offs[1:])
    W = int(word_offs[n])
    code = np.zeros(256, dtype=np.uint8)
    isn = np.ones(256, dtype=bool)
    for i, ch in enumerate(b"ACGT"):
        code[ch] = i
        code[ch + 32] = i
        isn[ch] = isn[ch + 32] = False
    # position of every base in the padded (32 bases per word) layout
    read_of = np.repeat(np.arange(n), lens)
    within = np.arange(len(arr)) - np.repeat(offs[:-1], lens) if len(arr) else np.zeros(0, dtype=np.int64)
    dst = (word_offs[:-1].astype(np.int64)[read_of] * 32 + within) if len(arr) else np.zeros(0, dtype=np.int64)
    codes = np.zeros(W * 32, dtype=np.uint8)
    codes[dst] = code[arr]
    nflag = np.zeros(W * 32, dtype=bool)
    nflag[dst] = isn[arr]
    c4 = codes.reshape(-1, 4)
    packed = ((c4[:, 0] << 6) | (c4[:, 1] << 4) | (c4[:, 2] << 2) | c4[:, 3]).astype(np.uint8)
    nmask = None
    if nflag.any():
        bits = np.packbits(nflag.reshape(-1, 32), axis=1, bitorder="big")  # bit 31 = first base
        nmask = (bits[:, 0].astype(np.uint32) << 24) | (bits[:, 1].astype(np.uint32) << 16) | \
                (bits[:, 2].astype(np.uint32) << 8) | bits[:, 3].astype(np.uint32)
        nmask = np.ascontiguousarray(nmask, dtype=np.uint32)
    return np.ascontiguousarray(packed), nmask, word_offs, lens.astype(np.uint16)


def _pack_fixed_len(reads2d, chunk=1 << 18):
    """fast path of pack_reads_2bit for an [n, L] uint8 ASCII array"""
    n, L = reads2d.shape
    wpr = (L + 31) // 32
    code = np.zeros(256, dtype=np.uint8)
    isn = np.ones(256, dtype=bool)
    for i, ch in enumerate(b"ACGT"):
        code[ch] = i
        code[ch + 32] = i
        isn[ch] = isn[ch + 32] = False
    packed = np.zeros((n, wpr * 8), dtype=np.uint8)
    nmask = np.zeros((n, wpr), dtype=np.uint32)
    any_n = False
    for s0 in range(0, n, chunk):
        blk = reads2d[s0:s0 + chunk]
        c = np.zeros((blk.shape[0], wpr * 32), dtype=np.uint8)
        c[:, :L] = code[blk]
        c4 = c.reshape(blk.shape[0], wpr * 8, 4)
        packed[s0:s0 + chunk] = (c4[:, :, 0] << 6) | (c4[:, :, 1] << 4) | (c4[:, :, 2] << 2) | c4[:, :, 3]
        nf = isn[blk]
        if nf.any():
            any_n = True
            f = np.zeros((blk.shape[0], wpr * 32), dtype=bool)
            f[:, :L] = nf
            bits = np.packbits(f.reshape(blk.shape[0], wpr, 32), axis=2, bitorder="big").astype(np.uint32)
            nmask[s0:s0 + chunk] = (bits[:, :, 0] << 24) | (bits[:, :, 1] << 16) | (bits[:, :, 2] << 8) | bits[:, :, 3]
    word_offs = (np.arange(n + 1, dtype=np.uint64) * np.uint64(wpr))
    lens = np.full(n, L, dtype=np.uint16)
    return packed.reshape(-1), (nmask.reshape(-1) if any_n else None), word_offs, lens


class Bgx:
    """One seqset build on one GPU.  Method names follow the reference stages they replace."""

    def __init__(self, kmer_size=30, min_kmer_count=5, max_corrections=8, min_good_run=2, trim_after_portion=0.7,
                 device=0, sort_key_bits=0, count_batch_reads=0):
        self.L = load_library()
        o = Options()
        self.L.bgx_default_options(C.byref(o))
        o.kmer_size, o.min_kmer_count, o.max_corrections = kmer_size, min_kmer_count, max_corrections
        o.min_good_run, o.trim_after_portion, o.device, o.sort_key_bits = min_good_run, trim_after_portion, device, \
            sort_key_bits
        o.count_batch_reads = count_batch_reads
        self.h = C.c_void_p()
        if self.L.bgx_create(C.byref(o), C.byref(self.h)):
            raise BgxError(self.L.bgx_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.bgx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc:
            raise BgxError(self.L.bgx_last_error().decode())

    def _take(self, ptr, n, dtype):
        """numpy view of a library-owned (pinned) host buffer; bgx_free runs when the last view dies"""
        n = int(n)
        if not n:
            self.L.bgx_free(ptr)
            return np.zeros(0, dtype=dtype)
        buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr.value)
        weakref.finalize(buf, self.L.bgx_free, C.c_void_p(ptr.value))
        return np.frombuffer(buf, dtype=dtype)

    # -- prob_pass_processor::add ------------------------------------------------------------
    def add_reads(self, reads):
        """reads: list of str/bytes, or (bytes buffer, offsets int64[n+1])"""
        if isinstance(reads, tuple):
            buf, offs = reads
        else:
            bs = [r.encode() if isinstance(r, str) else bytes(r) for r in reads]
            offs = np.zeros(len(bs) + 1, dtype=np.uint64)
            if bs:
                np.cumsum([len(b) for b in bs], out=offs[1:])
            buf = b"".join(bs)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        if isinstance(buf, np.ndarray):
            p = buf.ctypes.data
        else:
            self._keep = C.create_string_buffer(buf, len(buf)) if len(buf) else C.create_string_buffer(1)
            p = C.addressof(self._keep)
        self._ck(self.L.bgx_add_reads_ascii(self.h, p, offs.ctypes.data, len(offs) - 1))

    def add_reads_fastq(self, text):
        """text: bytes of uncompressed FASTQ (whole records); returns the number of reads added"""
        n = C.c_uint64()
        self._ck(self.L.bgx_add_reads_fastq(self.h, text, len(text), C.byref(n)))
        return int(n.value)

    def add_reads_packed(self, packed, nmask, word_offs, lens):
        self._ck(self.L.bgx_add_reads_packed(self.h, packed.ctypes.data, None if nmask is None else nmask.ctypes.data,
                                             None if word_offs is None else word_offs.ctypes.data, lens.ctypes.data,
                                             len(lens)))

    def add_reads_packed_ptr(self, packed_ptr, nmask_ptr, word_offs_ptr, lens_ptr, n, overlap=False):
        """overlap=True: bgx_add_reads_packed_async (the buffers must outlive the next count_kmers / run)"""
        f = self.L.bgx_add_reads_packed_async if overlap else self.L.bgx_add_reads_packed
        self._ck(f(self.h, packed_ptr, nmask_ptr, word_offs_ptr, lens_ptr, n))

    # -- kmer_counter / run_kmerize_subtask ----------------------------------------------------
    def count_kmers(self):
        self._ck(self.L.bgx_count_kmers(self.h))

    def export_kmers(self, min_count=1):
        n = C.c_uint64()
        pk, pf, pr, pfl = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._ck(self.L.bgx_export_kmers(self.h, min_count, C.byref(n), C.byref(pk), C.byref(pf), C.byref(pr),
                                         C.byref(pfl)))
        return {"kmers": self._take(pk, n.value, np.uint64), "fwd": self._take(pf, n.value, np.uint32),
                "rev": self._take(pr, n.value, np.uint32), "flags": self._take(pfl, n.value, np.uint8)}

    # -- correct_reads::correct ------------------------------------------------------------------
    def correct(self):
        self._ck(self.L.bgx_correct(self.h))

    def seed_uncorrected(self):
        """seqset_for_reads seeding (no k-mer stage, no correction)"""
        self._ck(self.L.bgx_seed_uncorrected(self.h))

    def export_varbit(self, which):
        """packed_varbit_vector of entry_sizes (which=0) / shared (which=1)"""
        p, nw, b, mv = C.c_void_p(), C.c_uint64(), C.c_uint32(), C.c_uint64()
        self._ck(self.L.bgx_export_varbit(self.h, which, C.byref(p), C.byref(nw), C.byref(b), C.byref(mv)))
        return {"elements": self._take(p, nw.value, np.uint64), "bits_per_value": b.value, "max_value": mv.value}

    def export_corrected(self):
        n, nb = C.c_uint64(), C.c_uint64()
        pl, pb, pc, pf, pr = (C.c_void_p() for _ in range(5))
        self._ck(self.L.bgx_export_corrected(self.h, C.byref(n), C.byref(pl), C.byref(pb), C.byref(nb), C.byref(pc),
                                             C.byref(pf), C.byref(pr)))
        lens = self._take(pl, n.value, np.uint16)
        seq = self._take(pb, nb.value, np.uint8).tobytes()
        offs = np.zeros(n.value + 1, dtype=np.int64)
        np.cumsum(lens, out=offs[1:])
        return {"seq": seq, "offs": offs, "kept": (lens > 0).astype(np.uint8), "lens": lens,
                "corrections": self._take(pc, n.value, np.uint8).astype(np.int32),
                "next_fwd": self._take(pf, n.value, np.uint16).astype(np.int32),
                "next_rev": self._take(pr, n.value, np.uint16).astype(np.int32), "n_kept": int((lens > 0).sum())}

    def export_reads(self):
        """the resident reads back as ASCII: (bytes, lens uint16[n])"""
        n, nb = C.c_uint64(), C.c_uint64()
        pl, pb = C.c_void_p(), C.c_void_p()
        self._ck(self.L.bgx_export_reads(self.h, C.byref(n), C.byref(pl), C.byref(pb), C.byref(nb)))
        lens = self._take(pl, n.value, np.uint16)
        return self._take(pb, nb.value, np.uint8).tobytes(), lens

    # -- expander + builder -----------------------------------------------------------------------
    def build_seqset(self):
        self._ck(self.L.bgx_build_seqset(self.h))

    def run(self):
        self._ck(self.L.bgx_run(self.h))

    # -- multi-GPU ------------------------------------------------------------------------------------
    @staticmethod
    def unique_id():
        """rank 0: 128-byte NCCL id to broadcast to every rank"""
        L = load_library()
        buf = (C.c_uint8 * 128)()
        if L.bgx_dist_unique_id(buf):
            raise BgxError(L.bgx_last_error().decode())
        return bytes(buf)

    def dist_init(self, world_size, rank, unique_id):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._ck(self.L.bgx_dist_init(self.h, world_size, rank, buf))

    def seqset_layout(self):
        lay = (C.c_uint64 * 6)()
        self._ck(self.L.bgx_seqset_layout(self.h, lay))
        return dict(zip(("n", "n_global", "first", "prev_words", "sub_words", "acc_words"), (int(v) for v in lay)))

    def export_seqset(self, per_entry=True):
        """per_entry=False: without the uint16 sizes / shared arrays (a file writer takes them as
        packed_varbit_vector elements from export_varbit instead)"""
        n, ml = C.c_uint64(), C.c_uint32()
        ps, psh = C.c_void_p(), C.c_void_p()
        pb, psub, pacc = (C.c_void_p * 4)(), (C.c_void_p * 4)(), (C.c_void_p * 4)()
        fixed = (C.c_uint64 * 5)()
        self._ck(self.L.bgx_export_seqset(self.h, C.byref(n), C.byref(ml), C.byref(ps) if per_entry else None,
                                          C.byref(psh) if per_entry else None, pb, psub, pacc, fixed))
        N = n.value
        lay = self.seqset_layout()
        words, subw, accw = lay["prev_words"], lay["sub_words"], lay["acc_words"]
        out = {"n": N, "n_global": lay["n_global"], "first": lay["first"], "max_entry_len": ml.value,
               "fixed": np.array(list(fixed), dtype=np.uint64)}
        if per_entry:
            out["sizes"] = self._take(ps, N, np.uint16)
            out["shared"] = self._take(psh, N, np.uint16)
        out["prev"] = np.stack([self._take(C.c_void_p(pb[b]), words, np.uint64) for b in range(4)]) if True else None
        out["subaccum"] = [self._take(C.c_void_p(psub[b]), subw, np.uint64) for b in range(4)]
        out["accum"] = [self._take(C.c_void_p(pacc[b]), accw, np.uint64) for b in range(4)]
        return out

    def export_entries(self, first=0, count=None):
        if count is None:
            count = self.stats().get("entries", 0) - first
        count = int(count)
        pb, po = C.c_void_p(), C.c_void_p()
        self._ck(self.L.bgx_export_entries_ascii(self.h, first, count, C.byref(pb), C.byref(po)))
        offs = self._take(po, count + 1, np.uint64)
        seq = self._take(pb, int(offs[-1]) if count else 0, np.uint8).tobytes().decode()
        return [seq[int(offs[i]):int(offs[i + 1])] for i in range(count)]

    def lookup_reads(self):
        """make_readmap's entry lookups: (fwd_entry, rc_entry) uint64[n_reads], 2^64-1 for dropped reads"""
        n = C.c_uint64()
        pf, pr = C.c_void_p(), C.c_void_p()
        self._ck(self.L.bgx_lookup_reads(self.h, C.byref(n), C.byref(pf), C.byref(pr)))
        return self._take(pf, n.value, np.uint64), self._take(pr, n.value, np.uint64)

    def build_readmap(self, paired=False):
        """make_readmap::create_from_reads: dict of the readmap's payload arrays (paired: reads 2i, 2i+1 are mates)"""
        m = C.c_uint64()
        pl, pp, pf = C.c_void_p(), C.c_void_p(), C.c_void_p()
        src, dst = (C.c_void_p * 3)(), (C.c_void_p * 3)()
        self._ck(self.L.bgx_build_readmap(self.h, 1 if paired else 0, C.byref(m), C.byref(pl), C.byref(pp), C.byref(pf), C.byref(src),
                                                   C.byref(dst)))
        rows = int(m.value)
        n_ent = int(self.stats().get("entries", 0))

        def bc(arr, nbits):
            return {"bits": self._take(C.c_void_p(arr[0]), (nbits + 63) // 64, np.uint64),
                    "subaccum": self._take(C.c_void_p(arr[1]), (nbits + 511) // 512, np.uint64),
                    "accum": self._take(C.c_void_p(arr[2]), (nbits + 1 + 511) // 512, np.uint64)}
        return {"n_rows": rows, "read_lengths": self._take(pl, rows, np.uint16), "mate_loop_ptr": self._take(pp, rows, np.uint64),
                "is_forward": self._take(pf, (rows + 63) // 64, np.uint64), "source_to_mid": bc(src, n_ent),
                "dest_to_mid": bc(dst, rows)}

    # -- biograph merge (seqset_flat_builder + make_mergemap + seqset_merger) -----------------------
    def merge_seqsets(self, parts, parallel_splits=0):
        """parts: list of dicts {"sizes": uint16[n], "prev": uint64[4, ceil(n/64)]} (the members of a seqset
        file).  Afterwards export_seqset / export_varbit / export_entries return the merged seqset."""
        keep, arr = [], (SeqsetPart * len(parts))()
        for i, p in enumerate(parts):
            sizes = np.ascontiguousarray(p["sizes"], dtype=np.uint16)
            n = len(sizes)
            prev = [np.ascontiguousarray(p["prev"][b], dtype=np.uint64) for b in range(4)]
            for w in prev:
                if len(w) != (n + 63) // 64:
                    raise BgxError("merge_seqsets: prev bit vectors must hold ceil(n / 64) words")
            keep += [sizes] + prev
            arr[i].n_entries = n
            arr[i].sizes = sizes.ctypes.data
            for b in range(4):
                arr[i].prev_bits[b] = prev[b].ctypes.data
        self._ck(self.L.bgx_merge_seqsets(self.h, arr, len(parts), int(parallel_splits)))
        del keep

    def _bitcount3(self, arr, nbits):
        return {"bits": self._take(C.c_void_p(arr[0]), (nbits + 63) // 64, np.uint64),
                "subaccum": self._take(C.c_void_p(arr[1]), (nbits + 511) // 512, np.uint64),
                "accum": self._take(C.c_void_p(arr[2]), (nbits + 1 + 511) // 512, np.uint64), "nbits": nbits}

    def export_mergemap(self, part):
        """the `merged_entries` bitcount of input `part` (seqset_mergemap): bits / subaccum / accum, n_set"""
        out = (C.c_void_p * 3)()
        nb, ns = C.c_uint64(), C.c_uint64()
        self._ck(self.L.bgx_export_mergemap(self.h, part, C.byref(out), C.byref(nb), C.byref(ns)))
        r = self._bitcount3(out, int(nb.value))
        r["n_set"] = int(ns.value)
        return r

    def migrate_bits(self, part, bits, n_old):
        """make_readmap::fast_migrate for a bit vector over the entries of input `part` (read_ids/source_to_mid)"""
        bits = np.ascontiguousarray(bits, dtype=np.uint64)
        if len(bits) != (int(n_old) + 63) // 64:
            raise BgxError("migrate_bits: the bit vector must hold ceil(n_old / 64) words")
        out = (C.c_void_p * 3)()
        nb = C.c_uint64()
        self._ck(self.L.bgx_migrate_bits(self.h, part, bits.ctypes.data, int(n_old), C.byref(out), C.byref(nb)))
        return self._bitcount3(out, int(nb.value))

    def export_flat(self, part, first=0, count=None):
        """seqset_flat::get(i) of input `part` for i in [first, first + count): list of str"""
        pb, po = C.c_void_p(), C.c_void_p()
        count = int(count)
        self._ck(self.L.bgx_export_flat_ascii(self.h, part, first, count, C.byref(pb), C.byref(po)))
        offs = self._take(po, count + 1, np.uint64)
        seq = self._take(pb, int(offs[-1]) if count else 0, np.uint8).tobytes().decode()
        return [seq[int(offs[i]):int(offs[i + 1])] for i in range(count)]

    def reset_results(self):
        self._ck(self.L.bgx_reset_results(self.h))

    def clear_reads(self):
        self._ck(self.L.bgx_clear_reads(self.h))

    def timer_start(self):
        self._ck(self.L.bgx_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_double()
        self._ck(self.L.bgx_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def debug_sort_pairs(self, keys, vals, begin_bit=0, end_bit=64):
        """in-place stable radix sort of host uint64 arrays on the GPU (test hook)"""
        assert keys.dtype == np.uint64 and vals.dtype == np.uint64 and len(keys) == len(vals)
        self._ck(self.L.bgx_debug_sort_pairs(self.h, keys.ctypes.data, vals.ctypes.data, len(keys), begin_bit, end_bit))

    def launch_count(self):
        return int(self.L.bgx_launch_count())

    def stats(self):
        buf = C.create_string_buffer(1 << 16)
        self._ck(self.L.bgx_stats_json(self.h, buf, len(buf)))
        return json.loads(buf.value.decode())


def assemble_seqset(parts):
    """Concatenate the per-rank tables of a multi-GPU build (in rank order) into the whole seqset.
    Every rank but the last holds a multiple of 512 entries, so the bit vectors and their bitcount
    index simply concatenate; `fixed` and `max_entry_len` are global on every rank."""
    parts = sorted(parts, key=lambda p: p["first"])
    n = sum(p["n"] for p in parts)
    assert all(p["n_global"] == n for p in parts)
    out = {"n": n, "max_entry_len": parts[0]["max_entry_len"], "fixed": parts[0]["fixed"],
           "sizes": np.concatenate([p["sizes"] for p in parts]), "shared": np.concatenate([p["shared"] for p in parts]),
           "prev": np.stack([np.concatenate([p["prev"][b] for p in parts]) for b in range(4)]),
           "subaccum": [np.concatenate([p["subaccum"][b] for p in parts]) for b in range(4)],
           "accum": [np.concatenate([p["accum"][b] for p in parts]) for b in range(4)]}
    return out

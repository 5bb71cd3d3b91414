"""CPU restatement of `biograph merge`'s seqset path -- test infrastructure: imported
only by tests/ (same rule as oracle.py).  Sequences are ASCII `bytes` over ACGT (A<C<G<T is also the
byte order, so Python's bytes order is the reference's dna_slice order: lexicographic, a proper
prefix first, modules/bio_base/dna_sequence.cpp:528-566).

Follows
  seqset_flat        modules/bio_base/seqset_flat.{h,cpp}: get(i) = the sequence of entry i.  The flat
                     file's own layout (which trace claimed which entry) depends on thread timing and
                     is "not versioned ... only for the same run" (seqset_flat.h:8-10); the contract is
                     get(i), which is seqset_range::sequence (seqset.cpp:676-689): first base from
                     `fixed` (entry_get_base, :249-254), then the (i - fixed[b])-th set bit of prev_b
                     (inner_pop_front, :709-720).
  make_mergemap      modules/bio_base/make_mergemap.cpp:188-233 (count_range): k-way merge of the
                     parts' sorted entries; an entry that is equal to / a prefix of the next one in
                     the merged order is folded into it; bit x of part p = a member of merged entry
                     x's run came from part p.
  seqset_merger      modules/bio_base/seqset_merger.cpp:109-197 (merge_range), :80-107
                     (get_base_iterator), driven over generate_chunks(0, n, g_parallel_splits = 100000)
                     (modules/io/parallel.cpp:13,60-83).  The chunking is RESULT-VISIBLE: the prev bit
                     of a candidate b+x lands on the first entry OF THE CHUNK that holds the last entry
                     prefixed by x (get_base_iterator of the chunk's limit backs up over that
                     candidate, so every earlier chunk leaves it alone).  `biograph create` puts the
                     same bit on the first such entry of the whole seqset (bs/builder.cpp:85-107); both
                     are valid (seqset_range::pop_front widens over `shared`, seqset.cpp:611-628).
  fast_migrate       modules/bio_mapred/make_readmap.cpp:459-520: a readmap moves to the merged seqset by
                     sending every source position through find_count of the part's mergemap.

Pinned against the reference's own merge output: datasets/lambdaToyData/benchmark/family_lambda.bg =
`biograph merge --in proband --in father --in mother` of the three *_lambda.bg next to it (fixtures in
tests/golden/ref_seqsets.npz and ref_merge_readmaps.npz; tests/test_oracle_merge.py): every payload
member of the merged seqset -- prev bits included -- and of the three migrated readmaps."""
import bisect
import heapq

import numpy as np

K_PARALLEL_SPLITS = 100000  # g_parallel_splits, modules/io/parallel.cpp:13


# ---- seqset_flat ------------------------------------------------------------------------------------
def pop_front_table(fixed, prev01):
    """next[i] = inner_pop_front(entry i) for every entry; prev01: 4 arrays of n 0/1 values"""
    n = int(fixed[4])
    nxt = np.zeros(n, dtype=np.int64)
    first = np.zeros(n, dtype=np.int64)
    for b in range(4):
        lo, hi = int(fixed[b]), int(fixed[b + 1])
        ones = np.flatnonzero(np.asarray(prev01[b])[:n])
        assert len(ones) == hi - lo, "prev bit totals do not match `fixed`"
        nxt[lo:hi] = ones
        first[lo:hi] = b
    return first, nxt


def flat_sequences(fixed, prev01, sizes):
    """seqset_flat::get(i) for every i: list of bytes"""
    first, nxt = pop_front_table(fixed, prev01)
    n = len(sizes)
    sizes = np.asarray(sizes, dtype=np.int64)
    mx = int(sizes.max()) if n else 0
    mat = np.zeros((n, mx), dtype=np.uint8)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    cur = np.arange(n)
    for c in range(mx):
        live = sizes > c
        mat[live, c] = acgt[first[cur[live]]]
        cur = nxt[cur]
    return [mat[i, :sizes[i]].tobytes() for i in range(n)]


# ---- make_mergemap ----------------------------------------------------------------------------------
def is_equal_or_prefix(prefix, longer):  # make_mergemap.cpp:12-18
    return len(prefix) <= len(longer) and longer[:len(prefix)] == prefix


def make_mergemap(parts):
    """parts: list of sorted prefix-free lists of bytes.  Literal count_range over the whole range
    (:188-233): returns (merged sequences, [bit array per part])."""
    heap = []
    for p, ent in enumerate(parts):
        if ent:
            heapq.heappush(heap, (ent[0], p, 0))
    merged, marks = [], [[] for _ in parts]

    def advance(p, i):
        if i + 1 < len(parts[p]):
            heapq.heappush(heap, (parts[p][i + 1], p, i + 1))

    while heap:
        seq, p, i = heapq.heappop(heap)
        x = len(merged)
        marks[p].append(x)
        advance(p, i)
        while heap and is_equal_or_prefix(seq, heap[0][0]):
            seq, p, i = heapq.heappop(heap)
            marks[p].append(x)
            advance(p, i)
        merged.append(seq)  # seqset_merger::iterator::dereference: the longest member of the run
    bits = []
    for p in range(len(parts)):
        b = np.zeros(len(merged), dtype=np.uint8)
        b[np.asarray(marks[p], dtype=np.int64)] = 1
        assert int(b.sum()) == len(parts[p])  # seqset_merger.cpp:33: total_bits == flat size
        bits.append(b)
    return merged, bits


def make_mergemap_sorted(parts):
    """the same result from one sort of the union (what the GPU computes): a record is dropped when it
    is equal to / a prefix of its successor; its merged index is that of the next kept record."""
    allrec = sorted((s, p) for p, ent in enumerate(parts) for s in ent)
    merged, bits_idx = [], [[] for _ in parts]
    run = []
    for j, (s, p) in enumerate(allrec):
        run.append(p)
        last = j + 1 == len(allrec) or not is_equal_or_prefix(s, allrec[j + 1][0])
        if last:
            for q in run:
                bits_idx[q].append(len(merged))
            merged.append(s)
            run = []
    bits = []
    for p in range(len(parts)):
        b = np.zeros(len(merged), dtype=np.uint8)
        b[np.asarray(bits_idx[p], dtype=np.int64)] = 1
        bits.append(b)
    return merged, bits


# ---- seqset_merger ----------------------------------------------------------------------------------
def generate_chunks(n, nsplits=K_PARALLEL_SPLITS):  # modules/io/parallel.cpp:60-83
    out = []
    for i in range(nsplits):
        a, b = n * i // nsplits, n * (i + 1) // nsplits
        if a != b:
            out.append((a, b))
    return out


def _get_base_iterator(ent, base, it):  # seqset_merger.cpp:80-107; `it` an index, len(ent) = end()
    n = len(ent)
    if it == n:
        if base == 3:
            return n
        return bisect.bisect_left(ent, b"ACGT"[base + 1:base + 2])
    search = b"ACGT"[base:base + 1] + ent[it]
    res = bisect.bisect_left(ent, search)
    while res != 0:
        prev = ent[res - 1]
        m = min(len(prev), len(search))
        if prev[:m] == search[:m]:
            res -= 1
        else:
            break
    return res


def merge_range(ent, start, limit, sizes, shared, prev):  # seqset_merger.cpp:109-197
    n = len(ent)
    prev_seq = b"" if start == 0 else ent[start - 1]
    it = [_get_base_iterator(ent, b, start) for b in range(4)]
    lim = [_get_base_iterator(ent, b, limit) for b in range(4)]
    for cur in range(start, limit):
        cs = ent[cur]
        sizes[cur] = len(cs)
        for b in range(4):
            if it[b] == lim[b]:
                continue
            cand = ent[it[b]]
            ov = min(len(cand) - 1, len(cs))
            if cand[1:1 + ov] == cs[:ov]:
                prev[b][cur] = 1
                it[b] += 1
            else:
                assert cs[:ov] < cand[1:1 + ov], "Out-of-order prevs"
        m = min(len(cs), len(prev_seq))
        sh = m
        for i in range(m):
            if cs[i] != prev_seq[i]:
                sh = i
                break
        shared[cur] = sh
        prev_seq = cs
    assert it == lim, "merge_range: candidates left over"
    return n


def merge_tables(ent, nsplits=K_PARALLEL_SPLITS):
    """seqset_merger::build: literal, chunk by chunk"""
    n = len(ent)
    sizes = np.zeros(n, dtype=np.uint16)
    shared = np.zeros(n, dtype=np.uint16)
    prev = [np.zeros(n, dtype=np.uint8) for _ in range(4)]
    for a, b in generate_chunks(n, nsplits):
        merge_range(ent, a, b, sizes, shared, prev)
    fixed = np.zeros(5, dtype=np.uint64)
    for b in range(4):
        fixed[b + 1] = fixed[b] + np.uint64(int(prev[b].sum()))
    assert int(fixed[4]) == n, "Invalid seqset in finalize"  # seqset.cpp:123-126
    return {"n": n, "sizes": sizes, "shared": shared, "prev": prev, "fixed": fixed}


def chunk_start_of(e, n, nsplits=K_PARALLEL_SPLITS):
    """start of the generate_chunks chunk that holds entry e"""
    c = ((e + 1) * nsplits - 1) // n
    return n * c // nsplits


def merge_prev_closed_form(ent, nsplits=K_PARALLEL_SPLITS):
    """what the GPU computes: candidate b+x sets bit b at max(lo, chunk start of hi-1), [lo, hi) = the
    entries prefixed by x"""
    n = len(ent)
    prev = [np.zeros(n, dtype=np.uint8) for _ in range(4)]
    code = {65: 0, 67: 1, 71: 2, 84: 3}
    for e in ent:
        x = e[1:]
        lo = bisect.bisect_left(ent, x)
        hi = lo
        if len(x) == 0:
            hi = n
        else:
            # first entry after the ones that start with x: x with its last base bumped
            hi = bisect.bisect_left(ent, x + b"\xff")
        assert hi > lo, "Missing expansion?"
        pos = max(lo, chunk_start_of(hi - 1, n, nsplits))
        assert prev[code[e[0]]][pos] == 0
        prev[code[e[0]]][pos] = 1
    return prev


# ---- readmap migration ------------------------------------------------------------------------------
def migrate_source_bits(old_source01, mergemap01):
    """fast_migrate (make_readmap.cpp:472-486): bit e of the old sparse_multi source moves to
    find_count(e) of the mergemap = the position of its e-th set bit"""
    sel = np.flatnonzero(np.asarray(mergemap01))
    old = np.flatnonzero(np.asarray(old_source01))
    out = np.zeros(len(mergemap01), dtype=np.uint8)
    out[sel[old]] = 1
    return out

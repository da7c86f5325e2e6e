"""Larger shapes of the BASELINE.json configs: oracle comparison where the oracle finishes in seconds, and
size-independent properties (cross-kernel consistency, split invariance, determinism, backend agreement) beyond that."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from gtars_b200 import ffi
    c = ffi.Context(0)
    yield c
    c.close()


def _np(t):
    return t.cpu().numpy().view(np.uint32)


def _universe(ctx, kind, n=1_000_000, nested=0.0):
    from gtars_b200 import ffi, synth
    u = synth.make_universe(n, nested_frac=nested)
    offs = u["chrom_offsets"].numpy().astype(np.uint64)
    s, e, v = (_np(u[k]) for k in ("g_start", "g_end", "g_val"))
    return u, ffi.Index(ctx, kind, offs, s, e, v), (offs, s, e, v)


def test_c2_properties_100m_queries(ctx):
    """C2 at 1/10 scale (1 000 files x 100 000 regions vs the full 1 M-region universe)."""
    import torch
    from gtars_b200 import ffi, synth
    from oracle import oracle as orc
    u, bits, arrays = _universe(ctx, ffi.KIND_BITS)
    _, ail, _ = _universe(ctx, ffi.KIND_AILIST)
    n_files, per_file = 1000, 100_000
    q = synth.make_query_files(u, n_files, per_file, device="cuda")
    qc, qs, qe = (_np(q[k]) for k in ("chr", "start", "end"))
    fo = q["file_offsets"].cpu().numpy().astype(np.uint64)
    off, ids = bits.tokenize_files(fo, qc, qs, qe, u["unk_id"])
    # (a) the fused find kernel agrees with the independent counting kernel (Bits identity through other arrays)
    counts = bits.count(qc, qs, qe)
    assert int(off[-1]) == int(counts.sum(dtype=np.uint64)) == len(ids)
    per_file_counts = np.add.reduceat(counts.astype(np.uint64), fo[:-1].astype(np.int64))
    assert np.array_equal(np.diff(off), per_file_counts)
    assert ids.max() < u["unk_id"]  # no file is empty, so no unk
    # (b) determinism and split invariance (first 300 files + rest == whole)
    off2, ids2 = bits.tokenize_files(fo, qc, qs, qe, u["unk_id"])
    assert np.array_equal(off, off2) and np.array_equal(ids, ids2)
    k = 300
    o1, i1 = bits.tokenize_files(fo[:k + 1], qc[:fo[k]], qs[:fo[k]], qe[:fo[k]], u["unk_id"])
    o2, i2 = bits.tokenize_files(fo[k:] - fo[k], qc[fo[k]:], qs[fo[k]:], qe[fo[k]:], u["unk_id"])
    assert np.array_equal(np.concatenate([i1, i2]), ids)
    assert np.array_equal(np.concatenate([o1, o2[1:] + o1[-1]]), off)
    # (c) AIList returns the same hits per query, in reverse order within a query (single component universe)
    offq_b, vb = bits.find(qc[:5_000_000], qs[:5_000_000], qe[:5_000_000])
    offq_a, va = ail.find(qc[:5_000_000], qs[:5_000_000], qe[:5_000_000])
    assert np.array_equal(offq_a, offq_b)
    two = np.flatnonzero(np.diff(offq_b) == 2)
    assert len(two) > 1000
    assert np.array_equal(vb[offq_b[two]], va[offq_a[two] + 1]) and np.array_equal(vb[offq_b[two] + 1], va[offq_a[two]])
    assert np.array_equal(np.sort(vb), np.sort(va))
    # (d) the oracle on the first 20 files
    m = int(fo[20])
    o = orc.Index(orc.BITS, *arrays)
    oo, oi = o.tokenize_files(fo[:21], qc[:m], qs[:m], qe[:m], u["unk_id"], threads=orc.max_threads())
    assert np.array_equal(oo, off[:21]) and np.array_equal(oi, ids[:int(off[20])])


def test_c2_nested_universe_vs_oracle(ctx):
    """C2n: 1 % wide intervals force multi-hit queries, bin-table overflow windows and several AIList components."""
    from gtars_b200 import ffi, synth
    from oracle import oracle as orc
    for kind, okind in ((ffi.KIND_BITS, orc.BITS), (ffi.KIND_AILIST, orc.AILIST)):
        u, ix, arrays = _universe(ctx, kind, n=300_000, nested=0.01)
        info = ix.info()
        assert info["bt_pool_windows"] > 0           # wide intervals break the candidate runs: pooled lists
        if kind == ffi.KIND_AILIST:
            assert info["max_components"] >= 2       # several AIList components, still served by the window table
        q = synth.make_query_files(u, 30, 20_000, unknown_frac_ppm=500)
        qc, qs, qe = (_np(q[k]) for k in ("chr", "start", "end"))
        fo = q["file_offsets"].numpy().astype(np.uint64)
        off, ids = ix.tokenize_files(fo, qc, qs, qe, u["unk_id"])
        oo, oi = orc.Index(okind, *arrays).tokenize_files(fo, qc, qs, qe, u["unk_id"], threads=orc.max_threads())
        assert np.array_equal(off, oo) and np.array_equal(ids, oi)


def test_c3_bits_count_database(ctx, monkeypatch):
    """C3 at 1/10 scale: 10 M unsorted queries vs a 5 M-interval overlapping database (count only)."""
    from gtars_b200 import ffi, synth
    from oracle import oracle as orc
    db = synth.make_uniform_intervals(5_000_000, synth.SEED_LOLA_DB, min_w=100, max_w=10_000)
    g = synth.group_by_chrom(db["chr"], db["start"], db["end"])
    offs = g["chrom_offsets"].numpy().astype(np.uint64)
    s, e = _np(g["g_start"]), _np(g["g_end"])
    ix = ffi.Index(ctx, ffi.KIND_BITS, offs, s, e)
    q = synth.make_uniform_intervals(10_000_000, synth.SEED_QUERIES, min_w=100, max_w=2000, log_uniform=False)
    qc, qs, qe = (_np(q[k]) for k in ("chr", "start", "end"))
    counts = ix.count(qc, qs, qe)                                  # 160 MB of rank LUTs: the bucketed pass
    monkeypatch.setenv("GTGPU_COUNT_PARTITION", "0")
    assert np.array_equal(ix.count(qc, qs, qe), counts)            # ... agrees with the direct pass
    monkeypatch.delenv("GTGPU_COUNT_PARTITION")
    raw = ix.bits_count(qc, qs, qe)
    assert np.array_equal(raw, counts.astype(np.uint64))           # proper inputs: identity == enumerated count
    assert np.array_equal(ix.any(qc, qs, qe), counts > 0)
    m = 300_000
    o = orc.Index(orc.BITS, offs, s, e)
    assert np.array_equal(counts[:m], o.count(qc[:m], qs[:m], qe[:m], threads=orc.max_threads()))
    # the enumerating kernel agrees with the counting kernel
    off, vals = ix.find(qc[:200_000], qs[:200_000], qe[:200_000])
    assert np.array_equal(np.diff(off), counts[:200_000].astype(np.uint64))


def test_c4_lola_matrix(ctx):
    """C4 shape, reduced: 400 database sets x 5 000 regions, 60 user sets sampled from a 100 k-region universe."""
    from gtars_b200 import ffi, synth
    from oracle import oracle as orc
    n_db, per_db = 400, 5000
    db = synth.make_uniform_intervals(n_db * per_db, synth.SEED_LOLA_DB, min_w=200, max_w=5000)
    dfo = (np.arange(n_db + 1) * per_db).astype(np.uint64)
    dc, ds, de = (_np(db[k]) for k in ("chr", "start", "end"))
    u = synth.make_universe(100_000)
    rng = np.random.default_rng(9)
    n_user, per_user = 60, 2000
    pick = np.concatenate([rng.choice(u["n"], per_user, replace=False) for _ in range(n_user)])
    uc, us, ue = (_np(u[k])[pick] for k in ("chr", "start", "end"))
    qc = np.concatenate([uc, _np(u["chr"])])
    qs = np.concatenate([us, _np(u["start"])])
    qe = np.concatenate([ue, _np(u["end"])])
    so = np.concatenate([np.arange(n_user + 1) * per_user, [n_user * per_user + u["n"]]]).astype(np.uint64)
    g = ffi.Igd(ctx, dfo, synth.N_CHROMS, dc, ds, de)
    hits = g.count_region_hits(so, qc, qs, qe, 1)
    pairs = g.count_set_overlaps(so, qc, qs, qe, 1)
    assert (pairs >= hits).all() and (hits[:n_user] <= per_user).all()
    # every user set is a subset of the universe: b, c, d are never negative
    tables = orc.lola_tables(hits[:n_user], hits[n_user], np.full(n_user, per_user), u["n"])
    assert (tables >= 0).all() and (tables.sum(axis=2) == u["n"]).all()
    # oracle on 6 user sets (the reference does O(n_files) work per region, so keep it small)
    o = orc.Igd(dfo, dc, ds, de)
    k = 6
    so_k = so[:k + 1]
    m = int(so_k[-1])
    assert np.array_equal(hits[:k], o.count_region_hits(so_k, qc[:m], qs[:m], qe[:m], 1, threads=orc.max_threads()))
    assert np.array_equal(pairs[:k], o.count_set_overlaps(so_k, qc[:m], qs[:m], qe[:m], 1, threads=orc.max_threads()))


def test_c5_fragments_unsorted(ctx):
    """C5 at 1/200 scale: 5 M unsorted fragments, Zipf-ish barcodes, vs the full 1 M-peak universe."""
    from gtars_b200 import ffi, synth
    from oracle import oracle as orc
    u, ix, arrays = _universe(ctx, ffi.KIND_BITS)
    n, n_bc = 5_000_000, 20_000
    q = synth.make_query_files(u, 1, n, seed=synth.SEED_FRAGMENTS, sort_files=False)
    qc, qs, qe = (_np(q[k]) for k in ("chr", "start", "end"))
    rng = np.random.default_rng(3)
    bc = (rng.zipf(1.3, n) % n_bc).astype(np.uint32)
    off, ids = ix.tokenize_fragments(qc, qs, qe, bc, n_bc, u["unk_id"])
    oo, oi = orc.Index(orc.BITS, *arrays).tokenize_fragments(qc, qs, qe, bc, n_bc, u["unk_id"])
    assert np.array_equal(off, oo) and np.array_equal(ids, oi)
    counts = ix.count(qc, qs, qe)
    assert len(ids) == int(np.maximum(counts, 1).sum(dtype=np.uint64))  # per-fragment [unk]


def test_c5_packed_group_by_equals_pair_sort(ctx, monkeypatch):
    """The two-pass group-by over packed words (sort.cu radix_group_values: barcode offsets from counts, tiles inside one
    low digit and tiles across a boundary) against the plain (barcode, token) pair sort + search, 17-bit barcodes with
    9-bit digits, and both against the oracle on the barcodes of a slice."""
    from gtars_b200 import ffi, synth
    from oracle import oracle as orc
    u, ix, arrays = _universe(ctx, ffi.KIND_BITS)
    n, n_bc = 6_000_000, 100_000
    q = synth.make_query_files(u, 1, n, seed=synth.SEED_FRAGMENTS + 1, sort_files=False)
    qc, qs, qe = (_np(q[k]) for k in ("chr", "start", "end"))
    rng = np.random.default_rng(5)
    bc = (rng.integers(0, n_bc, n) ** 2 // n_bc).astype(np.uint32)
    bc[bc == n_bc - 1] = n_bc - 2  # the last barcode stays empty
    off, ids = ix.tokenize_fragments(qc, qs, qe, bc, n_bc, u["unk_id"])
    monkeypatch.setenv("GTGPU_NO_GROUP_SORT", "1")
    off2, ids2 = ix.tokenize_fragments(qc, qs, qe, bc, n_bc, u["unk_id"])
    monkeypatch.delenv("GTGPU_NO_GROUP_SORT")
    assert np.array_equal(off, off2) and np.array_equal(ids, ids2)
    assert off[n_bc - 1] == off[n_bc] == len(ids)
    m = 400_000
    oo, oi = orc.Index(orc.BITS, *arrays).tokenize_fragments(qc[:m], qs[:m], qe[:m], bc[:m], n_bc, u["unk_id"])
    go, gi = ix.tokenize_fragments(qc[:m], qs[:m], qe[:m], bc[:m], n_bc, u["unk_id"])
    assert np.array_equal(go, oo) and np.array_equal(gi, oi)


def test_scoring_matrix_properties_20m_fragments(ctx):
    """gtars-scoring at scale (4 files x 5 M unsorted fragments vs the 1 M-peak universe): the count matrix must agree
    with the independent counting kernel on the same lookups (ATAC: shifted start + reversed end interval; ChIP: the
    fragment), file by file, and with the oracle on the head of a file."""
    from gtars_b200 import ffi, synth
    from oracle import oracle as orc
    u, bits, (offs, s, e, v) = _universe(ctx, ffi.KIND_BITS)
    n_files, per_file = 4, 5_000_000
    q = synth.make_query_files(u, n_files, per_file, seed=synth.SEED_FRAGMENTS, device="cuda", sort_files=False)
    qc, qs, qe = (_np(q[k]) for k in ("chr", "start", "end"))
    fo = q["file_offsets"].cpu().numpy().astype(np.uint64)
    n_cols = int(u["n"])
    for mode in (ffi.SCORE_ATAC, ffi.SCORE_CHIP):
        mat = bits.score_matrix(fo, qc, qs, qe, mode, n_cols)
        if mode == ffi.SCORE_ATAC:
            ns, ne = qs + np.uint32(4), qe - np.uint32(5)
            counts = bits.count(qc, ns, ns + np.uint32(1)).astype(np.uint64) + bits.count(qc, ne, ne - np.uint32(1)).astype(np.uint64)
        else:
            counts = bits.count(qc, qs, qe).astype(np.uint64)
        per_file_counts = np.add.reduceat(counts, fo[:-1].astype(np.int64))
        assert np.array_equal(mat.sum(axis=1, dtype=np.uint64), per_file_counts), mode
        assert int(mat.sum(dtype=np.uint64)) > per_file
        m = 200_000
        sub = np.array([0, m], dtype=np.uint64)
        want = orc.score_matrix(orc.Index(orc.BITS, offs, s, e, v), sub, qc[:m], qs[:m], qe[:m], mode, n_cols, threads=orc.max_threads())
        assert np.array_equal(bits.score_matrix(sub, qc[:m], qs[:m], qe[:m], mode, n_cols), want), mode


def test_bed_ingest_properties_5m_lines(ctx, tmp_path):
    """Device BED ingest at scale: 5 M lines written in shuffled order with comments sprinkled in come back as exactly
    the source regions in RegionSet::sort order (stable: chromosome string, then start), and the tokens of the text
    equal the tokens of the arrays."""
    import torch
    from gtars_b200 import ffi, synth
    u, bits, _ = _universe(ctx, ffi.KIND_BITS)
    n = 5_000_000
    q = synth.make_query_files(u, 1, n, device="cuda", sort_files=False)
    qc, qs, qe = (_np(q[k]) for k in ("chr", "start", "end"))
    names = np.array(synth.CHROM_NAMES)
    lines = np.char.add(np.char.add(np.char.add(np.char.add(names[qc], "\t"), qs.astype(str)), "\t"), qe.astype(str))
    text = "# header comment\n" + "\n".join(lines.tolist()) + "\n"
    blob = "".join(synth.CHROM_NAMES).encode()
    name_off = np.zeros(len(synth.CHROM_NAMES) + 1, dtype=np.uint32)
    name_off[1:] = np.cumsum([len(x) for x in synth.CHROM_NAMES])
    import ctypes as C
    L = ffi.lib()
    n_out = C.c_uint64(0)
    hc, hs, he = C.c_void_p(), C.c_void_p(), C.c_void_p()
    raw = text.encode()
    ffi.check(L.gtgpu_parse_bed(ctx._h, raw, len(raw), len(synth.CHROM_NAMES), blob, name_off.ctypes.data_as(C.c_void_p),
                                C.byref(n_out), C.byref(hc), C.byref(hs), C.byref(he)))
    gc, gs, ge = ffi._take(hc), ffi._take(hs), ffi._take(he)
    assert n_out.value == n == len(gc)
    rank = np.argsort(np.argsort(names, kind="stable"), kind="stable")       # lexicographic rank of every chromosome name
    order = np.lexsort((qs, rank[qc]))                                         # stable: rank, then start
    assert np.array_equal(gc, qc[order]) and np.array_equal(gs, qs[order]) and np.array_equal(ge, qe[order])
    ids_text = bits.tokenize_bed(raw, list(synth.CHROM_NAMES), int(u["unk_id"]))
    off, ids_arr = bits.tokenize_files(np.array([0, n], dtype=np.uint64), gc, gs, ge, u["unk_id"])
    assert np.array_equal(ids_text, ids_arr)

"""Pins the CPU oracle against the reference's own known-answer tests (SURVEY.md §8c K1–K11).

Expected values live in tests/golden/kats.json (transcribed from the reference's tests, each with its
file:line) and the fixture files those tests read live in tests/golden/fixtures.json.
"""
import os

import numpy as np
import pytest

from oracle import oracle as orc
from tests import helpers
from tests.helpers import KINDS


def _sets(triples):
    return sorted((t[0], t[1]) for t in triples)


@pytest.mark.parametrize("name", ["K1_abcd", "K1_empty", "K1_single", "K2_nested26"])
def test_overlapper_kats(golden, name):
    case = golden[1][name]
    for kind in case["kinds"]:
        ov = orc.Overlapper(KINDS[kind], [tuple(iv) for iv in case["intervals"]])
        if kind == "ailist" and "ailist_components" in case:
            assert ov.num_components() == case["ailist_components"]
        for q in case["queries"]:
            hits = ov.find(*q["q"])
            if "set" in q:
                assert _sets(hits) == sorted(tuple(x) for x in q["set"]), (name, kind, q)
            if "n" in q:
                assert len(hits) == q["n"], (name, kind, q)


def test_bits_count_and_order(golden):
    case = golden[1]["K3_bits_count"]
    b = orc.Overlapper(orc.BITS, [tuple(iv) for iv in case["intervals"]])
    for q in case["queries"]:
        assert b.count(*q["q"]) == q["count"]
        assert len(b.find(*q["q"])) == q["find_n"]
    case = golden[1]["K3_bits_order"]
    b = orc.Overlapper(orc.BITS, [tuple(iv) for iv in case["intervals_val"]])
    for q in case["queries"]:
        assert [list(h) for h in b.find(*q["q"])] == q["ordered"]


@pytest.mark.parametrize("kind", ["bits", "ailist"])
def test_mco_kats(golden, kind):
    for case in golden[1]["K4_mco"]:
        cmap, offs, s, e, v = helpers.flatten_source(case["source"])
        ix = orc.Index(KINDS[kind], offs, s, e, v)
        qc, qs, qe = helpers.flatten_queries(case["query"], cmap)
        m = case["min_overlap"] if case["min_overlap"] is not None else 0
        if "count" in case:
            assert list(ix.count(qc, qs, qe, m)) == case["count"], case["cite"]
        if "any" in case:
            assert list(ix.any(qc, qs, qe, m)) == case["any"], case["cite"]
        if "find" in case:
            off, tr = ix.find(qc, qs, qe, m, coords=True)
            got = [sorted([int(a), int(b)] for a, b, _ in tr[int(off[i]):int(off[i + 1])]) for i in range(len(qc))]
            assert got == [sorted(x) for x in case["find"]], case["cite"]
        if "find_idx" in case:
            off, vals = ix.find(qc, qs, qe, m)
            got = [sorted(int(x) for x in vals[int(off[i]):int(off[i + 1])]) for i in range(len(qc))]
            assert got == case["find_idx"], case["cite"]


def test_tokenizer_kats(golden, fixture_dir):
    for case in golden[1]["K5_tokenizer"]:
        for kind in case["kinds"]:
            tok = orc.Tokenizer(os.path.join(fixture_dir, case["universe"]), KINDS[kind])
            if "vocab_size" in case:
                assert tok.vocab_size() == case["vocab_size"]
                assert tok.token_to_id("<unk>") == case["unk_id"]
            if "regions" in case:
                regions = [tuple(r) for r in case["regions"]]
                if "ids" in case:
                    assert tok.encode(regions) == case["ids"], (case["cite"], kind)
                if "tokens" in case:
                    assert tok.tokenize(regions) == case["tokens"], (case["cite"], kind)


def test_universe_id_assignment_duplicates(tmp_path):
    """universe/mod.rs:123-197 + gtars-core utils.rs:240-271: ids are first-appearance ranks; id_to_region is
    positional, so a duplicated line makes the two maps disagree."""
    p = tmp_path / "dup.bed"
    p.write_text("chr1\t10\t20\nchr1\t30\t40\nchr1\t10\t20\nchr1\t50\t60\n")
    tok = orc.Tokenizer(str(p))
    assert tok.token_to_id("chr1:10-20") == 0
    assert tok.token_to_id("chr1:30-40") == 1
    assert tok.token_to_id("chr1:50-60") == 2
    assert tok.id_to_token(2) == "chr1:10-20"  # positional
    # val=2 ("chr1:50-60") → id_to_region[2] = "chr1:10-20" → region_to_id = 0
    assert tok.encode([("chr1", 55, 58)]) == [0]


def test_universe_rejects_bad_files(tmp_path):
    p = tmp_path / "track.bed"
    p.write_text("track name=x\nchr1\t1\t2\n")
    with pytest.raises(ValueError):
        orc.Tokenizer(str(p))
    p = tmp_path / "blank.bed"
    p.write_text("chr1\t1\t2\n\nchr1\t5\t6\n")
    with pytest.raises(ValueError):
        orc.Tokenizer(str(p))


def test_regionset_parse_and_sort(fixture_dir, tmp_path):
    rs = orc.regionset_from_file(os.path.join(fixture_dir, "to_tokenize.bed"))
    assert rs == [("chr13", 74550222, 74550611), ("chr15", 49155456, 49155487), ("chr15", 49155846, 49156192)]
    p = tmp_path / "hdr.bed"
    p.write_text("chrom\tstart\tend\n#c\nchr2\t5\t9\textra\nchr10\t7\t8\nchr10\t3\t4\n")
    assert orc.regionset_from_file(str(p)) == [("chr10", 3, 4), ("chr10", 7, 8), ("chr2", 5, 9)]
    p = tmp_path / "empty.bed"
    p.write_text("#only a comment\n")
    with pytest.raises(ValueError):
        orc.regionset_from_file(str(p))


def test_igd_kats(golden):
    for case in golden[1]["K7_igd"]:
        cmap = helpers.ChromMap()
        g = orc.Igd()
        for chr_, s, e, val, f in case["adds"]:
            g.add(cmap.add(chr_), s, e, val, f)
        g.n_files = case["n_files"]
        orc.lib().orc_igd_set_n_files(g._h, g.n_files)
        g.finalize()
        for q in case["queries"]:
            chr_, s, e, m = q["q"]
            total, hits = g.count_overlaps(cmap.get(chr_), s, e, m)
            assert list(hits) == q["hits"], (case["cite"], q)
            if "total" in q:
                assert total == q["total"], (case["cite"], q)


def _db_from_case(case, golden):
    if "db" in case:
        return [[tuple(r) for r in st] for st in case["db"]]
    return [helpers.parse_bed_text(golden[0][f]["text"]) for f in case["db_files"]]


def test_igd_set_kats(golden):
    for case in golden[1]["K7_igd_sets"]:
        cmap = helpers.ChromMap()
        db = _db_from_case(case, golden)
        fo, dc, ds, de = helpers.flatten_sets(db, cmap, add=True)
        g = orc.Igd(fo, dc, ds, de)
        qc, qs, qe = helpers.flatten_queries(case["query"], cmap)
        so = np.array([0, len(qc)], dtype=np.uint64)
        if "set_overlaps" in case:
            assert list(g.count_set_overlaps(so, qc, qs, qe, case["min_overlap"])[0]) == case["set_overlaps"]
        if "per_query" in case:
            per_q = [g.count_overlaps(int(c), int(s), int(e), case["min_overlap"])[0] for c, s, e in zip(qc, qs, qe)]
            assert per_q == case["per_query"]


def test_lola_kats(golden):
    for case in golden[1]["K9_lola"]:
        cmap = helpers.ChromMap()
        fo, dc, ds, de = helpers.flatten_sets([[tuple(r) for r in st] for st in case["db"]], cmap, add=True)
        g = orc.Igd(fo, dc, ds, de)
        uo, uc, us, ue = helpers.flatten_sets([[tuple(r) for r in st] for st in case["user"]], cmap)
        vo, vc, vs, ve = helpers.flatten_sets([[tuple(r) for r in case["universe"]]], cmap)
        m = case["min_overlap"]
        user_hits = g.count_region_hits(uo, uc, us, ue, m)
        univ_hits = g.count_region_hits(vo, vc, vs, ve, m)[0]
        tables = orc.lola_tables(user_hits, univ_hits, np.diff(uo), len(vc))
        if "abcd" in case:
            assert tables.tolist() == case["abcd"], case["cite"]
        if "support" in case:
            assert tables[:, :, 0].tolist() == case["support"], case["cite"]


def test_igd_differential_inputs_vs_bruteforce(golden):
    """K8 inputs (the reference asserts legacy == new there, without literals): all-vs-all and query files
    against every fixture DB, checked against an O(n·m) brute force of the closed form."""
    k8 = golden[1]["K8_inputs"]
    for name, files in k8["dbs"].items():
        db = [helpers.parse_bed_text(golden[0][f]["text"]) for f in files]
        cmap = helpers.ChromMap()
        fo, dc, ds, de = helpers.flatten_sets(db, cmap, add=True)
        g = orc.Igd(fo, dc, ds, de)
        qsets = db + [helpers.parse_bed_text(golden[0][f]["text"]) for f in k8["queries"]]
        so, qc, qs, qe = helpers.flatten_sets(qsets, cmap)
        file_of = np.repeat(np.arange(len(db)), np.diff(fo).astype(np.int64))
        for m in (1, 5, 50):
            mat = helpers.brute_overlap_matrix(qc, qs, qe, dc, ds, de, m)
            pair = np.zeros((len(qsets), len(db)), dtype=np.uint64)
            binary = np.zeros_like(pair)
            for si in range(len(qsets)):
                rows = mat[int(so[si]):int(so[si + 1])]
                for f in range(len(db)):
                    sub = rows[:, file_of == f]
                    pair[si, f] = sub.sum()
                    binary[si, f] = sub.any(axis=1).sum()
            assert (g.count_set_overlaps(so, qc, qs, qe, m) == pair).all(), (name, m)
            assert (g.count_region_hits(so, qc, qs, qe, m) == binary).all(), (name, m)
        if name == "igd_file_list_01":
            assert g.count_set_overlaps(so, qc, qs, qe, 1)[0, 0] == 8


def test_derived_vectors(golden, fixture_dir):
    d = golden[1]["D_derived"]
    for kind in ("bits", "ailist"):
        tok = orc.Tokenizer(os.path.join(fixture_dir, d["D1"]["universe"]), KINDS[kind])
        rs = orc.regionset_from_file(os.path.join(fixture_dir, d["D1"]["query_file"]))
        assert tok.encode(rs) == d["D1"][kind]
        assert tok.encode([tuple(r) for r in d["D2"]["regions"]]) == d["D2"][kind]
    tok = orc.Tokenizer(os.path.join(fixture_dir, d["D3"]["universe"]))
    assert tok.token_to_id("<unk>") == 4
    assert tok.tokenize_fragment_file(os.path.join(fixture_dir, d["D3"]["fragments"])) == d["D3"]["expect"]


def test_bits_count_equals_find_len_random():
    """Bits::count (bits.rs:337-344) equals find().len() on proper inputs; AIList finds the same set."""
    rng = np.random.default_rng(7)
    n = 3000
    s = rng.integers(0, 100000, n).astype(np.uint32)
    e = s + rng.integers(1, 5000, n).astype(np.uint32)
    offs = np.array([0, n], dtype=np.uint64)
    b = orc.Index(orc.BITS, offs, s, e)
    a = orc.Index(orc.AILIST, offs, s, e)
    qs = rng.integers(0, 110000, 2000).astype(np.uint32)
    qe = qs + rng.integers(1, 3000, 2000).astype(np.uint32)
    qc = np.zeros(2000, dtype=np.uint32)
    cb = b.count(qc, qs, qe)
    assert (b.bits_count(qc, qs, qe) == cb).all()
    assert (a.count(qc, qs, qe) == cb).all()
    brute = helpers.brute_overlap_matrix(qc, qs, qe, np.zeros(n), s, e, 1).sum(axis=1)
    assert (cb == brute).all()
    ob, vb = b.find(qc, qs, qe)
    oa, va = a.find(qc, qs, qe)
    assert (ob == oa).all()
    for i in range(0, 2000, 37):
        assert sorted(vb[int(ob[i]):int(ob[i + 1])]) == sorted(va[int(oa[i]):int(oa[i + 1])])


def test_igd_single_set_kats(golden):
    """igd.rs:1244-1355 through the oracle's tile walk: distinct subjects per query (value = subject index)."""
    for case in golden[1]["K7_igd_single"]:
        cmap = helpers.ChromMap()
        g = orc.Igd()
        for i, (chr_, s, e) in enumerate(case["subject"]):
            g.add(cmap.add(chr_), s, e, i, i)  # one "file" per subject: per-file hits > 0 <=> the pair overlaps
        g.n_files = len(case["subject"])
        orc.lib().orc_igd_set_n_files(g._h, g.n_files)
        g.finalize()
        pairs, per_query = [], []
        for qi, (chr_, s, e) in enumerate(case["query"]):
            hits = np.zeros(g.n_files, dtype=np.uint64)
            if cmap.get(chr_) != helpers.UNKNOWN:
                g.count_overlaps(cmap.get(chr_), s, e, case["min_overlap"], hits)
            pairs += [[qi, int(j)] for j in np.flatnonzero(hits)]
            per_query.append(int((hits > 0).sum()))
        if "pairs" in case:
            assert pairs == case["pairs"], case["cite"]
        if "per_query" in case:
            assert per_query == case["per_query"], case["cite"]


def test_scoring_kat(golden, fixture_dir):
    """K10: gtars-scoring/src/fragment_scoring.rs:180-210 — the ATAC count matrix of the two fixture fragment files."""
    k = golden[1]["K10_scoring"]
    files = [os.path.join(fixture_dir, f) for f in k["fragment_files"]]
    m = orc.region_scoring_files(os.path.join(fixture_dir, k["consensus"]), files, orc.SCORE_ATAC)
    assert m.shape == (k["rows"], k["cols"])
    assert m.tolist() == k["matrix"]
    # Chip mode is not asserted by the reference; it must at least differ from Atac on this fixture or equal a
    # brute-force count of whole-fragment overlaps
    chip = orc.region_scoring_files(os.path.join(fixture_dir, k["consensus"]), files, orc.SCORE_CHIP)
    peaks = [l.split() for l in open(os.path.join(fixture_dir, k["consensus"])).read().splitlines() if l.strip()]
    import gzip
    for row, f in enumerate(files):
        want = [0] * len(peaks)
        for line in gzip.open(f, "rt"):
            c, s, e = line.split()[:3]
            for j, (pc, ps, pe) in enumerate(peaks):
                if pc == c and int(ps) < int(e) and int(pe) > int(s):
                    want[j] += 1
        assert chip[row].tolist() == want


def test_scoring_dense_vs_bruteforce():
    """score_matrix / score_barcodes restatement vs a literal double loop (incl. reversed ATAC end intervals)."""
    rng = np.random.default_rng(5)
    n_chroms, n = 3, 400
    chr_ = np.sort(rng.integers(0, n_chroms, n))
    s = rng.integers(0, 20_000, n).astype(np.uint32)
    e = (s + rng.integers(1, 600, n)).astype(np.uint32)
    offs = np.searchsorted(chr_, np.arange(n_chroms + 1)).astype(np.uint64)
    vals = rng.permutation(n).astype(np.uint32)
    ix = orc.Index(orc.BITS, offs, s, e, vals)
    nf = 1500
    fc = rng.integers(0, n_chroms + 1, nf).astype(np.uint32)
    fs = rng.integers(0, 20_500, nf).astype(np.uint32)
    fe = (fs + rng.integers(6, 900, nf)).astype(np.uint32)
    fo = np.array([0, 0, 700, 700, 1500], dtype=np.uint64)
    def overl(c, qs, qe):  # Bits::find semantics for any (qs, qe), reversed included: iv.start < qe && iv.end > qs
        if c >= n_chroms:
            return []
        lo, hi = int(offs[c]), int(offs[c + 1])
        return [int(vals[i]) for i in range(lo, hi) if int(s[i]) < qe and int(e[i]) > qs]
    for mode in (orc.SCORE_ATAC, orc.SCORE_CHIP):
        got = orc.score_matrix(ix, fo, fc, fs, fe, mode, n)
        want = np.zeros((4, n), dtype=np.uint32)
        for f in range(4):
            for i in range(int(fo[f]), int(fo[f + 1])):
                if mode == orc.SCORE_ATAC:
                    ns, ne = int(fs[i]) + 4, int(fe[i]) - 5
                    hits = overl(fc[i], ns, ns + 1) + overl(fc[i], ne, ne - 1)
                else:
                    hits = overl(fc[i], int(fs[i]), int(fe[i]))
                for v in hits:
                    want[f, v] += 1
        assert np.array_equal(got, want), mode
    bc = rng.integers(0, 37, nf).astype(np.uint32)
    off, pk, ct = orc.score_barcodes(ix, fc, fs, fe, bc, 40)
    want = {}
    for i in range(nf):
        for v in overl(fc[i], int(fs[i]), int(fe[i])):
            want[(int(bc[i]), v)] = want.get((int(bc[i]), v), 0) + 1
    got = {}
    for b in range(40):
        for k in range(int(off[b]), int(off[b + 1])):
            got[(b, int(pk[k]))] = int(ct[k])
    assert got == want and int(off[-1]) == len(want)


def test_igd_file_format_roundtrip(tmp_path):
    """.igd layout (igd.rs:418-486) and loader (:320-414): sizes follow from the header, contigs come in creation
    order, a multi-tile interval is stored once per tile, and the loaded database answers like the original."""
    import struct
    g = orc.Igd()
    recs = [(1, 100, 200, 0), (0, 16000, 17000, 1), (1, 50, 60, 1), (1, 40000, 40010, 0), (0, 5, 6, 2), (2, -5, 10, 0), (2, 7, 7, 0)]
    for c, s, e, f in recs:
        g.add(c, s, e, 0, f)
    g.finalize()
    p = str(tmp_path / "t.igd")
    g.save(p, ["chrA", "chrB", "chrC"])
    b = open(p, "rb").read()
    nbp, g_type, n_ctg = struct.unpack("<3i", b[:12])
    assert (nbp, g_type, n_ctg) == (16384, 1, 2)            # chrC never got a valid record
    n_tiles = struct.unpack("<2i", b[12:20])
    assert n_tiles == (3, 2)                                 # creation order: chrB (tiles 0..2), then chrA (0..1)
    cnt = struct.unpack("<5i", b[20:40])
    assert cnt == (2, 0, 1, 2, 1)                            # 16000-17000 spans tiles 0 and 1 of chrA
    assert b[40:44] == b"chrB" and b[80:84] == b"chrA" and len(b) == 40 + 80 + 16 * sum(cnt)
    first = struct.unpack("<4i", b[120:136])
    assert first == (1, 50, 60, 0)                           # tile records are sorted by start
    h = orc.Igd.from_igd_file(p)
    assert h.contig_names == ["chrB", "chrA"] and h.n_files == 3
    for (c_old, c_new) in ((1, 0), (0, 1)):
        for q in ((0, 70000), (55, 150), (16383, 16385), (16500, 16600), (5, 6)):
            a = g.count_overlaps(c_old, q[0], q[1])
            bb = h.count_overlaps(c_new, q[0], q[1])
            assert a[0] == bb[0] and list(a[1]) == list(bb[1])


def test_barcode_scoring_file_key_set(tmp_path):
    """fragment_scoring.rs:146-153 + files.rs:106-129 restated: an entry per barcode seen on a consensus chromosome (possibly
    empty), none for barcodes seen on unknown chromosomes only; '#' lines skipped."""
    from oracle import oracle as orc
    cons, frags = str(tmp_path / "c.bed"), str(tmp_path / "f.tsv")
    with open(cons, "w") as f:
        f.write("chr1\t100\t200\nchr1\t300\t400\nchr2\t50\t80\n")
    with open(frags, "w") as f:
        f.write("#header\nchr1\t120\t130\tHIT\t1\nchr1\t1000\t1100\tMISS\t1\nchr9\t120\t130\tGHOST\t1\n"
                "chr2\t60\t70\tHIT\t2\nchr1\t150\t350\tBOTH\t1\nchr1\t150\t160\tBOTH\t1\n")
    assert orc.barcode_scoring_file(cons, frags) == {"HIT": {0: 1, 2: 1}, "MISS": {}, "BOTH": {0: 2, 1: 1}}

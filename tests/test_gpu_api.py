"""The reference's own tests for this path, replayed through the host-side API mirror (gtars_b200.api → C++ host layer →
C ABI → CUDA).  Expected values come from tests/golden/kats.json (reference file:line in each entry)."""
import os

import numpy as np
import pytest

from tests import helpers
from tests.helpers import KINDS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from gtars_b200 import api as a
    a.device(0)
    return a


def _toml(fixture_dir, kind):
    p = os.path.join(fixture_dir, "tokenizers", f"tokenizer_{kind}.toml")
    with open(p, "w") as f:
        f.write('universe = "peaks.bed.gz"\n' + (f'tokenizer_type = "{kind}"\n' if kind != "default" else ""))
    return p


def test_tokenizer_kats(api, golden, fixture_dir):
    """gtars-tokenizers/src/tokenizer.rs:294-496 and gtars-python/tests/test_tokenizers.py."""
    for case in golden[1]["K5_tokenizer"]:
        for kind in case["kinds"]:
            if case["universe"].endswith("peaks.bed") and kind == "ailist":
                tok = api.Tokenizer.from_config(_toml(fixture_dir, "ailist"))
            elif kind == "ailist":
                continue
            else:
                tok = api.Tokenizer(os.path.join(fixture_dir, case["universe"]))
            if "vocab_size" in case:
                assert tok.get_vocab_size() == case["vocab_size"]
                assert tok.unk_token_id == case["unk_id"]
            if "regions" in case:
                regions = [api.Region(*r) for r in case["regions"]]
                if "ids" in case:
                    assert tok.encode(regions) == case["ids"], (case["cite"], kind)
                    assert tok(regions)["input_ids"] == case["ids"]
                if "tokens" in case:
                    assert tok.tokenize(regions) == case["tokens"], (case["cite"], kind)


def test_tokenizer_config_variants(api, fixture_dir):
    """tokenizer.rs:294-358: from_config / from_auto, bad tokenizer type, custom special tokens."""
    d = os.path.join(fixture_dir, "tokenizers")
    assert api.Tokenizer.from_auto(_toml(fixture_dir, "bits")).get_vocab_size() == 32
    assert api.Tokenizer.from_auto(_toml(fixture_dir, "default")).get_vocab_size() == 32
    assert api.Tokenizer.from_auto(os.path.join(d, "peaks.bed.gz")).get_vocab_size() == 32
    bad = os.path.join(d, "bad.toml")
    open(bad, "w").write('universe = "peaks.bed.gz"\ntokenizer_type = "i-dont-exist"\n')
    with pytest.raises(api.GtarsError):
        api.Tokenizer.from_config(bad)
    custom = os.path.join(d, "custom.toml")
    open(custom, "w").write('universe = "peaks.bed.gz"\nspecial_tokens = [\n    {name="unk", token="<UNKNOWN>"},\n]\n')
    tok = api.Tokenizer.from_config(custom)
    assert tok.get_vocab_size() == 32 and tok.unk_token == "<UNKNOWN>" and tok.pad_token == "<pad>"
    assert tok.tokenize([api.Region("chr1", 50, 150)]) == ["<UNKNOWN>"]
    with pytest.raises(api.GtarsError):
        api.Tokenizer.from_auto(os.path.join(d, "peaks.txt"))


def test_tokenizer_derived_and_duplicates(api, golden, fixture_dir, tmp_path):
    d = golden[1]["D_derived"]
    for kind in ("bits", "ailist"):
        tok = api.Tokenizer.from_config(_toml(fixture_dir, kind))
        assert tok.encode(os.path.join(fixture_dir, d["D1"]["query_file"])) == d["D1"][kind]
        assert tok.encode([api.Region(*r) for r in d["D2"]["regions"]]) == d["D2"][kind]
    # duplicate universe lines: ids round-trip through strings, same as the oracle
    from oracle import oracle as orc
    p = tmp_path / "dup.bed"
    p.write_text("chr1\t10\t20\nchr1\t30\t40\nchr1\t10\t20\nchr1\t50\t60\n")
    tok, o = api.Tokenizer(str(p)), orc.Tokenizer(str(p))
    for q in ([("chr1", 55, 58)], [("chr1", 12, 14)], [("chr1", 0, 100)], [("chr2", 1, 2)]):
        assert tok.encode([api.Region(*r) for r in q]) == o.encode(q)
    assert tok.encode_batch([[api.Region("chr1", 55, 58)], [], [api.Region("chr9", 1, 2)]]) == [[0], [tok.unk_token_id], [tok.unk_token_id]]


def test_fragment_tokenization(api, golden, fixture_dir):
    """utils/fragments.rs:114-156 asserts 2 barcodes; exact vectors are the derived D3 (and the oracle's)."""
    from oracle import oracle as orc
    d3 = golden[1]["D_derived"]["D3"]
    tok = api.Tokenizer(os.path.join(fixture_dir, d3["universe"]))
    for frag in ("fragments/region_scoring/fragments1.bed.gz", "fragments/region_scoring/fragments2.bed.gz"):
        got = api.tokenize_fragment_file(os.path.join(fixture_dir, frag), tok)
        assert len(got) == 2
        assert got == orc.Tokenizer(os.path.join(fixture_dir, d3["universe"])).tokenize_fragment_file(os.path.join(fixture_dir, frag))
    assert api.tokenize_fragment_file(os.path.join(fixture_dir, d3["fragments"]), tok) == d3["expect"]


@pytest.mark.parametrize("kind", ["bits", "ailist"])
def test_mco_kats(api, golden, kind):
    """gtars-overlaprs/src/multi_chrom_overlapper.rs:715-1157 and gtars-python/tests/test_regionset.py:37-55."""
    for case in golden[1]["K4_mco"]:
        src = [api.Region(*r) for r in case["source"]]
        query = [api.Region(*r) for r in case["query"]]
        mco = api.MultiChromOverlapper(src, KINDS[kind])
        m = case["min_overlap"]
        if "count" in case:
            assert mco.count_overlaps(query, m) == case["count"], case["cite"]
        if "any" in case:
            assert mco.any_overlaps(query, m) == case["any"], case["cite"]
        if "find" in case:
            got = [sorted([r.start, r.end] for r in hits) for hits in mco.find_overlaps_regions(query, m)]
            assert got == [sorted(x) for x in case["find"]], case["cite"]
        if "find_idx" in case:
            assert [sorted(h) for h in mco.find_overlaps(query, m)] == case["find_idx"], case["cite"]
    # subset_by / intersect_all: deduplicated and sorted (multi_chrom_overlapper.rs:1044-1068)
    mco = api.MultiChromOverlapper([("chr1", 100, 200), ("chr1", 300, 400), ("chr2", 500, 600)], KINDS[kind])
    sub = mco.subset_by([("chr1", 150, 250), ("chr2", 550, 650), ("chr1", 160, 170)])
    assert [(r.chr, r.start, r.end) for r in sub] == [("chr1", 100, 200), ("chr2", 500, 600)]


def test_regionset_binding_ops(api, fixture_dir):
    """gtars-python/tests/test_regionset.py:37-55: self = queries, other = index."""
    a = api.RegionSet([("chr1", 100, 200), ("chr1", 300, 400), ("chr1", 500, 600)])
    b = api.RegionSet([("chr1", 150, 250), ("chr1", 550, 650)])
    assert a.count_overlaps(b) == [1, 0, 1]
    assert a.any_overlaps(b) == [True, False, True]
    assert a.find_overlaps(b) == [[0], [], [1]]
    rs = api.RegionSet(os.path.join(fixture_dir, "to_tokenize.bed"))  # parsed, then sorted by (chr, start)
    assert [(r.chr, r.start, r.end) for r in rs] == [("chr13", 74550222, 74550611), ("chr15", 49155456, 49155487),
                                                      ("chr15", 49155846, 49156192)]
    with pytest.raises(api.GtarsError):
        api.RegionSet(os.path.join(fixture_dir, "does_not_exist.bed"))


def test_igd_and_lola_kats(api, golden):
    """gtars-igd/src/igd.rs:1160-1221 and gtars-lola/src/enrichment.rs:830-1104."""
    for case in golden[1]["K7_igd_sets"]:
        if "db" not in case:
            db = [helpers.parse_bed_text(golden[0][f]["text"]) for f in case["db_files"]]
        else:
            db = [[tuple(r) for r in st] for st in case["db"]]
        igd = api.Igd(db)
        q = [tuple(r) for r in case["query"]]
        if "set_overlaps" in case:
            assert igd.count_set_overlaps(q, case["min_overlap"]) == case["set_overlaps"], case["cite"]
    for case in golden[1]["K9_lola"]:
        igd = api.Igd([[tuple(r) for r in st] for st in case["db"]])
        t = api.lola_contingency(igd, [[tuple(r) for r in st] for st in case["user"]], [tuple(r) for r in case["universe"]],
                                 case["min_overlap"])
        if "abcd" in case:
            assert t.tolist() == case["abcd"], case["cite"]
        if "support" in case:
            assert t[:, :, 0].tolist() == case["support"], case["cite"]
    with pytest.raises(api.GtarsError):  # enrichment.rs:855-877: empty universe is an error
        api.lola_contingency(api.Igd([[("chr1", 100, 200)]]), [[("chr1", 100, 200)]], [])


def test_igd_unit_kats_through_c_abi(golden):
    """gtars-igd/src/igd.rs:914-1461 (count_overlaps on hand-built databases), via gtgpu_igd_*."""
    from gtars_b200 import ffi
    ctx = ffi.Context(0)
    for case in golden[1]["K7_igd"]:
        cmap = helpers.ChromMap()
        per_file = [[] for _ in range(case["n_files"])]
        for chr_, s, e, _val, f in case["adds"]:
            per_file[f].append((cmap.add(chr_), s & 0xFFFFFFFF, e & 0xFFFFFFFF))
        fo = np.cumsum([0] + [len(x) for x in per_file]).astype(np.uint64)
        flat = [r for x in per_file for r in x]
        g = ffi.Igd(ctx, fo, max(len(cmap), 1), [r[0] for r in flat], [r[1] for r in flat], [r[2] for r in flat])
        for q in case["queries"]:
            chr_, s, e, m = q["q"]
            so = np.array([0, 1], dtype=np.uint64)
            hits = g.count_set_overlaps(so, [cmap.get(chr_)], [s & 0xFFFFFFFF], [e & 0xFFFFFFFF], m)[0]
            assert list(hits) == q["hits"], (case["cite"], q)
        g.close()
    ctx.close()


def test_igd_single_region_set_kats(api, golden):
    """gtars-igd/src/igd.rs:1244-1355: find_overlaps_regionset / count_overlaps_per_query."""
    for case in golden[1]["K7_igd_single"]:
        igd = api.Igd.from_single_region_set([tuple(r) for r in case["subject"]])
        q = [tuple(r) for r in case["query"]]
        if "pairs" in case:
            assert igd.find_overlaps_regionset(q, case["min_overlap"]) == [tuple(p) for p in case["pairs"]], case["cite"]
        if "per_query" in case:
            assert igd.count_overlaps_per_query(q, case["min_overlap"]) == case["per_query"], case["cite"]
    # subjects Igd::add would drop (empty / reversed) never count
    igd = api.Igd.from_single_region_set([("chr1", 100, 100), ("chr1", 300, 200), ("chr1", 50, 150)])
    assert igd.find_overlaps_regionset([("chr1", 0, 1000)], 1) == [(0, 2)]


def test_scoring_kat_through_api(api, golden, fixture_dir):
    """K10 (gtars-scoring/src/fragment_scoring.rs:180-210) through ConsensusSet / region_scoring_from_fragments."""
    k = golden[1]["K10_scoring"]
    cons = api.ConsensusSet(os.path.join(fixture_dir, k["consensus"]))
    assert len(cons) == k["cols"]
    files = [os.path.join(fixture_dir, f) for f in k["fragment_files"]]
    m = api.region_scoring_from_fragments(files, cons, api.SCORING_ATAC)
    assert m.shape == (k["rows"], k["cols"]) and m.tolist() == k["matrix"]
    from oracle import oracle as orc
    chip = api.region_scoring_from_fragments(files, cons, api.SCORING_CHIP)
    assert np.array_equal(chip, orc.region_scoring_files(os.path.join(fixture_dir, k["consensus"]), files, orc.SCORE_CHIP))
    # sparse per-barcode counts of the first file: the oracle's dict (incl. its key set), and sums equal to the Chip row
    sparse = api.barcode_scoring_from_fragments(files[0], cons)
    assert sparse == orc.barcode_scoring_file(os.path.join(fixture_dir, k["consensus"]), files[0])
    tot = np.zeros(k["cols"], dtype=np.int64)
    for bc, counts in sparse.items():
        for peak, c in counts.items():
            tot[peak] += c
    assert tot.tolist() == chip[0].tolist()


def test_barcode_scoring_key_set_matches_reference(api, tmp_path):
    """fragment_scoring.rs:146-153 + files.rs:106-129: a barcode gets an entry as soon as one of its fragments lies on a
    chromosome the consensus knows — even when nothing overlaps (empty inner map) — and none when it was only ever seen on
    unknown chromosomes."""
    from oracle import oracle as orc
    cons_path, frag_path = str(tmp_path / "consensus.bed"), str(tmp_path / "frags.tsv")
    with open(cons_path, "w") as f:
        f.write("chr1\t100\t200\nchr1\t300\t400\nchr2\t50\t80\n")
    with open(frag_path, "w") as f:
        f.write("# a comment line\n"
                "chr1\t120\t130\tHIT\t1\n"          # overlaps peak 0
                "chr1\t1000\t1100\tMISS\t1\n"       # known chromosome, no overlap  -> {}
                "chr9\t120\t130\tGHOST\t1\n"        # unknown chromosome only      -> absent
                "chr9\t1\t2\tHIT\t1\n"
                "chr2\t60\t70\tHIT\t2\n"
                "chr1\t150\t350\tBOTH\t1\n"         # peaks 0 and 1
                "chr1\t150\t160\tBOTH\t1\n")
    want = orc.barcode_scoring_file(cons_path, frag_path)
    assert want == {"HIT": {0: 1, 2: 1}, "MISS": {}, "BOTH": {0: 2, 1: 1}}
    got = api.barcode_scoring_from_fragments(frag_path, api.ConsensusSet(cons_path))
    assert got == want and "GHOST" not in got


def test_indexed_region_set_kats(api, golden):
    """K12: gtars-overlaprs/src/indexed_region_set.rs:396-545."""
    for case in golden[1]["K12_indexed_region_set"]["cases"]:
        for kind in (api.AILIST, api.BITS):
            irs = api.IndexedRegionSet([api.Region(*r) for r in case["reference"]], kind)
            q = [api.Region(*r) for r in case["query"]]
            if "count" in case:
                assert irs.count_overlaps(q) == case["count"], case["name"]
            if "any" in case:
                assert irs.any_overlaps(q) == case["any"], case["name"]
            if "find" in case:
                assert irs.find_overlaps(q) == case["find"], case["name"]
            if "intersect_all" in case:
                got = [[r.chr, r.start, r.end] for r in irs.intersect_all(q)]
                assert got == case["intersect_all"], case["name"]


def test_igd_file_roundtrip(api, golden, fixture_dir, tmp_path):
    """.igd format (gtars-igd/src/igd.rs:320-486; structure checks of lib.rs:471-540): the writer matches the oracle's
    restatement byte for byte, and a database loaded from the file answers exactly like the one built in memory."""
    import struct
    from oracle import oracle as orc
    k = golden[1]["K8_inputs"]
    rels = k["dbs"]["lola_multi_db"] + k["dbs"]["igd_file_list_02"]
    sets = [api.RegionSet(os.path.join(fixture_dir, r)) for r in rels]
    names = [os.path.basename(r) for r in rels]
    db = api.Igd(sets)
    p = str(tmp_path / "db.igd")
    db.save(p, names)
    # the oracle's writer over the same add() sequence
    chrom_ids, o = {}, orc.Igd()
    for f, rs in enumerate(sets):
        for r in rs:
            if r.start < r.end:
                o.add(chrom_ids.setdefault(r.chr, len(chrom_ids)), r.start, r.end, 0, f)
    o.finalize()
    po = str(tmp_path / "oracle.igd")
    o.save(po, list(chrom_ids))
    got, want = open(p, "rb").read(), open(po, "rb").read()
    assert got == want
    nbp, g_type, n_ctg = struct.unpack("<3i", got[:12])
    assert (nbp, g_type) == (16384, 1) and n_ctg == len(chrom_ids)
    tsv = open(str(tmp_path / "db.tsv")).read().splitlines()
    assert tsv[0] == "Index\tFile\tNumber of Regions\tAvg size" and len(tsv) == 1 + len(sets)
    for i, rs in enumerate(sets):
        kept = [r for r in rs if r.start < r.end]
        idx, name, cnt, avg = tsv[1 + i].split("\t")
        assert (int(idx), name, int(cnt)) == (i, names[i], len(kept))
        assert avg == "%.2f" % (sum(r.end - r.start for r in kept) / len(kept))
    # load it back: same answers as the in-memory database, for both count semantics
    loaded = api.Igd.from_igd_file(p)
    assert loaded.num_files() == len(sets)
    queries = [api.RegionSet(os.path.join(fixture_dir, q)) for q in k["queries"]] + sets[:2]
    for m in (1, 25):
        for q in queries:
            assert loaded.count_set_overlaps(q, m) == db.count_set_overlaps(q, m)
            assert loaded.count_region_hits(q, m) == db.count_region_hits(q, m)
    assert sum(db.count_set_overlaps(queries[0])) > 0
    # and the oracle reading the file agrees with the device on a query set
    of = orc.Igd.from_igd_file(p)
    assert of.contig_names == list(chrom_ids) and of.n_files == len(sets)


def _messy_bed(rng, names, n, sort=False, crlf=False, header=None, extra_cols=True):
    rows = []
    for _ in range(n):
        c = names[rng.integers(0, len(names))]
        s = int(rng.integers(0, 3_000_000))
        e = s + int(rng.integers(0, 5000))
        rows.append((c, s, e))
    if sort:
        rows.sort(key=lambda r: (r[0], r[1]))
    lines = [header] if header else []
    for i, (c, s, e) in enumerate(rows):
        if i % 97 == 5:
            lines.append("# a comment")
        if i % 211 == 7:
            lines.append("track name=x")
        if i % 307 == 9:
            lines.append("browser position chr1:1-2")
        sfx = ""
        if extra_cols and i % 3 == 0:
            sfx = f"\tname{i}\t{i % 1000}\t+"
        lines.append(f"{c}\t{'+' if i % 50 == 0 else ''}{s}\t{e}{sfx}")
    nl = "\r\n" if crlf else "\n"
    return nl.join(lines) + (nl if n % 2 else "")


@pytest.mark.parametrize("variant", ["unsorted", "sorted", "crlf", "header", "gz"])
def test_parse_bed_on_device_matches_region_set(api, tmp_path, variant):
    """gtgpu_parse_bed vs the oracle's RegionSet::try_from (region_set.rs:60-185 parse, :502-505 sort)."""
    import gzip
    import zlib
    from oracle import oracle as orc
    rng = np.random.default_rng(zlib.crc32(variant.encode()))
    known = ["chr1", "chr10", "chr2", "chrX", "chr1_KI270706v1_random", "chrM"]
    names_in_file = known + ["chrUn_unknown", "zzz"]
    text = _messy_bed(rng, names_in_file if variant != "sorted" else known, 20_000, sort=(variant == "sorted"),
                      crlf=(variant == "crlf"), header="chrom\tchromStart\tchromEnd\tname" if variant == "header" else None)
    p = str(tmp_path / ("q.bed.gz" if variant == "gz" else "q.bed"))
    if variant == "gz":
        with gzip.open(p, "wt", newline="") as f:
            f.write(text)
    else:
        with open(p, "w", newline="") as f:
            f.write(text)
    table = ["chr2", "chrX", "chr1", "chrM", "chr10", "chr1_KI270706v1_random", "chrNeverSeen"]  # ids in THIS order
    c, s, e = api.parse_bed_file(p, table)
    ref = orc.regionset_from_file(p)
    known_ref = [(table.index(r[0]), r[1], r[2]) for r in ref if r[0] in table]
    unknown_ref = sorted((r[1], r[2]) for r in ref if r[0] not in table)
    n_known = len(known_ref)
    assert len(c) == len(ref)
    got_known = list(zip(c[:n_known].tolist(), s[:n_known].tolist(), e[:n_known].tolist()))
    assert got_known == known_ref                                   # reference order among the known chromosomes
    assert (c[n_known:] == 0xFFFFFFFF).all()                         # unknown names sort last ...
    assert sorted(zip(s[n_known:].tolist(), e[n_known:].tolist())) == unknown_ref
    assert s[n_known:].tolist() == sorted(s[n_known:].tolist())      # ... by start


def test_parse_bed_errors_and_tokenize_bed_file(api, fixture_dir, tmp_path):
    from oracle import oracle as orc
    def write(name, text):
        p = str(tmp_path / name)
        open(p, "w").write(text)
        return p
    for name, text in [("bad_start.bed", "chr1\t1\t2\nchr1\tx\t5\n"), ("neg.bed", "chr1\t-1\t2\n"),
                       ("two_cols.bed", "chr1\t1\t2\nchr1\t7\n"), ("blank.bed", "chr1\t1\t2\n\nchr1\t3\t4\n"),
                       ("overflow.bed", "chr1\t1\t4294967296\n"), ("only_comments.bed", "# nothing\ntrack x\n"),
                       ("space.bed", "chr1\t 1\t2\n"),
                       # BufRead::lines only drops a '\r' that precedes a '\n': a last line without newline keeps it -> "2\r"
                       ("cr_no_newline.bed", "chr1\t1\t2\r\nchr1\t1\t2\r")]:
        p = write(name, text)
        with pytest.raises(api.GtarsError):
            api.parse_bed_file(p, ["chr1"])
        with pytest.raises(ValueError):
            orc.regionset_from_file(p)                               # the reference rejects the same files
    c, s, e = api.parse_bed_file(write("crlf_last.bed", "chr1\t5\t6\r\nchr1\t1\t2\r\n"), ["chr1"])
    assert (c.tolist(), s.tolist(), e.tolist()) == ([0, 0], [1, 5], [2, 6])
    # a truncated gzip stream is an error, not a shorter file
    import gzip
    full = gzip.compress(("chr1\t1\t2\n" * 20000).encode())
    trunc = str(tmp_path / "trunc.bed.gz")
    open(trunc, "wb").write(full[: len(full) // 2])
    with pytest.raises(api.GtarsError):
        api.RegionSet(trunc)
    c, s, e = api.parse_bed_file(write("edge.bed", "chr1\t4294967295\t+0\nchr1\t0\t1"), ["chr1"])
    assert (c.tolist(), s.tolist(), e.tolist()) == ([0, 0], [0, 4294967295], [1, 0])
    # text -> token ids entirely on the device == encode(RegionSet(path))
    uni = os.path.join(fixture_dir, "tokenizers", "peaks.bed")
    tok = api.Tokenizer(uni)
    q = os.path.join(fixture_dir, "to_tokenize.bed")
    assert tok.encode_bed_file(q) == tok.encode(api.RegionSet(q)) == [22, 23, 24]
    assert tok.encode_bed_file(uni) == tok.encode(api.RegionSet(uni))
    gz = os.path.join(fixture_dir, "tokenizers", "peaks.bed.gz")
    assert tok.encode_bed_file(gz) == tok.encode(api.RegionSet(gz))
    miss = write("miss.bed", "chrNope\t1\t2\nchr1\t1\t2\n")
    assert tok.encode_bed_file(miss) == tok.encode(api.RegionSet(miss)) == [tok.unk_token_id]


def test_parse_bed_fuzz_vs_oracle(api, tmp_path):
    """Randomised small BED texts (odd separators, empty fields, signs, overflow, comments anywhere, missing final
    newline): the device parser accepts exactly the files the restated RegionSet::try_from accepts, with the same regions."""
    from oracle import oracle as orc
    rng = np.random.default_rng(20261017)
    names = ["chr1", "chr2", "chrX", "c"]
    starts = ["0", "5", "+7", "12", "4294967295", "4294967296", "-1", "", " 3", "3 ", "1e3", "0x10", "007"]
    tails = ["", "\tname", "\tname\t0\t+", "\t", "\t\t"]
    n_ok = n_bad = 0
    for it in range(300):
        lines = []
        for _ in range(int(rng.integers(1, 7))):
            k = rng.random()
            if k < 0.08:
                lines.append(["#c", "track t", "browser b", "# chr1\t1\t2"][int(rng.integers(0, 4))])
            elif k < 0.12:
                lines.append(["", "chr1", "chr1\t5", "chr1 5 9"][int(rng.integers(0, 4))])
            else:
                # mostly well-formed numbers, sometimes one of the odd spellings
                s = starts[int(rng.integers(0, len(starts)))] if rng.random() < 0.15 else str(int(rng.integers(0, 1000)))
                e = starts[int(rng.integers(0, len(starts)))] if rng.random() < 0.15 else str(int(rng.integers(0, 2000)))
                lines.append(names[int(rng.integers(0, len(names)))] + "\t" + s + "\t" + e + tails[int(rng.integers(0, len(tails)))])
        if rng.random() < 0.1:
            lines.insert(0, "chrom\tstart\tend")
        nl = "\r\n" if rng.random() < 0.2 else "\n"
        text = nl.join(lines) + (nl if rng.random() < 0.7 else "")
        p = str(tmp_path / f"f{it}.bed")
        with open(p, "w", newline="") as f:
            f.write(text)
        try:
            ref = orc.regionset_from_file(p)
        except ValueError:
            ref = None
        try:
            c, s_, e_ = api.parse_bed_file(p, names)
            got = list(zip([names[i] for i in c.tolist()], s_.tolist(), e_.tolist()))
        except api.GtarsError:
            got = None
        if ref is None:
            assert got is None, (it, text)
            n_bad += 1
        else:
            assert got == [tuple(r) for r in ref], (it, text)
            n_ok += 1
    assert n_ok > 50 and n_bad > 50


def test_fragment_file_text_ingest_on_device(api, golden, fixture_dir, tmp_path):
    """gtgpu_tokenize_fragments_text (text -> per-barcode token ids on the device) == the host-parsed path == the
    oracle, on the reference's fixture (D3) and on a generated file with '#' lines, mixed whitespace, unknown
    chromosomes, empty-hit fragments and many barcodes; malformed lines are rejected with the reference's line number."""
    import gzip
    from oracle import oracle as orc
    d3 = golden[1]["D_derived"]["D3"]
    uni = os.path.join(fixture_dir, d3["universe"])
    frag = os.path.join(fixture_dir, d3["fragments"])
    tok = api.Tokenizer(uni)
    assert api.tokenize_fragment_file(frag, tok, device_parse=True) == d3["expect"]
    rng = np.random.default_rng(31)
    peaks = os.path.join(fixture_dir, "tokenizers", "peaks.bed")
    tok2 = api.Tokenizer(peaks)
    regions = [l.split()[:3] for l in open(peaks).read().splitlines() if l.strip()]
    lines = ["# fragments"]
    for i in range(30_000):
        if i % 977 == 3:
            lines.append("#comment in the middle")
        c, s, e = regions[int(rng.integers(0, len(regions)))]
        s, e = int(s), int(e)
        if rng.random() < 0.3:
            c = ["chrNope", "chr1", "chrUn"][int(rng.integers(0, 3))]
        a = max(s + int(rng.integers(-300, 300)), 0)
        b = a + int(rng.integers(1, 800))
        bc = "BC%05d" % int(rng.integers(0, 1500) ** 2 // 1500)
        sep = ["\t", " ", "  \t"][int(rng.integers(0, 3))]
        lines.append(sep.join([c, str(a), str(b), bc, str(int(rng.integers(1, 9)))]) + ("\textra" if i % 5 == 0 else ""))
    p = str(tmp_path / "frags.tsv.gz")
    with gzip.open(p, "wt", newline="") as f:
        f.write("\n".join(lines) + "\n")
    host = api.tokenize_fragment_file(p, tok2)
    dev = api.tokenize_fragment_file(p, tok2, device_parse=True)
    assert dev == host and list(dev) == list(host)                      # same lists, same (first-appearance) barcode order
    assert dev == orc.Tokenizer(peaks).tokenize_fragment_file(p)
    assert len(dev) > 500
    for bad in ("chr1\t5\t9\tBC\n", "chr1\tx\t9\tBC\t1\n", "chr1\t5\t-9\tBC\t1\n"):
        q = str(tmp_path / "bad.tsv")
        open(q, "w").write("chr1\t1\t2\tB\t1\n" + bad)
        with pytest.raises(api.GtarsError):
            api.tokenize_fragment_file(q, tok2, device_parse=True)
        with pytest.raises(api.GtarsError):
            api.tokenize_fragment_file(q, tok2)
    q = str(tmp_path / "only_comments.tsv")
    open(q, "w").write("# nothing here\n")
    assert api.tokenize_fragment_file(q, tok2, device_parse=True) == api.tokenize_fragment_file(q, tok2) == {}

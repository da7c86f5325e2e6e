#!/usr/bin/env python
"""Generate tests/golden/{fixtures,kats}.json.

Run in the authoring container (needs /root/reference): `python tests/golden/make_golden.py`.

* fixtures.json — the text of the small fixture files under /root/reference/tests/data that the
  reference's own tests for the interval-overlap path read (gz files stored decompressed; tests re-gzip
  where the reader is extension-sensitive).  /root/reference does not exist on the GPU box, hence the copy.
* kats.json — the known-answer values those tests assert, transcribed by hand below with the reference
  file:line of every assertion (SURVEY.md §8c K1–K11), plus vectors marked "derived" that the reference does
  not assert (SURVEY.md §8c D1–D3) and that only serve as regression pins.

Neither file contains reference source code.
"""
import gzip
import json
import os

REF = "/root/reference/tests/data"
HERE = os.path.dirname(os.path.abspath(__file__))

FIXTURE_FILES = [
    "tokenizers/peaks.bed",
    "tokenizers/peaks.bed.gz",
    "tokenizers/peaks.scored.bed",
    "to_tokenize.bed",
    "igd_file_list_01/igd_bed_file_1.bed",
    "igd_file_list_02/igd_bed_file_1.bed",
    "igd_file_list_02/igd_bed_file_2.bed",
    "igd_query_files/query1.bed",
    "igd_query_files/query2.bed",
    "lola_multi_db/collection1/regions/cpgIslandExt.bed",
    "lola_multi_db/collection1/regions/laminB1Lads.bed",
    "lola_multi_db/collection1/regions/vistaEnhancers.bed",
    "lola_multi_db/collection2/regions/cpgIslandExt.bed",
    "lola_multi_db/collection2/regions/laminB1Lads.bed",
    "lola_multi_db/collection2/regions/vistaEnhancers.bed",
    "consensus/consensus1.bed",
    "fragments/region_scoring/fragments1.bed.gz",
    "fragments/region_scoring/fragments2.bed.gz",
]


def read_fixture(rel):
    p = os.path.join(REF, rel)
    if rel.endswith(".gz"):
        with gzip.open(p, "rt") as f:
            return {"gz": True, "text": f.read()}
    with open(p) as f:
        return {"gz": False, "text": f.read()}


ABCD = [(1, 5), (3, 7), (6, 10), (8, 12)]  # a b c d, ailist.rs:384-408 / bits.rs:544-568
NESTED26 = [(0, 30), (0, 10), (0, 10), (5, 15), (5, 15), (10, 20), (10, 20), (15, 25), (15, 25), (21, 22),
            (22, 23), (20, 30), (20, 30), (25, 100), (26, 27), (27, 28), (29, 30), (30, 31), (32, 33), (50, 51),
            (51, 52), (52, 53), (53, 54), (55, 56), (60, 61), (70, 71)]  # ailist.rs:556-583

KATS = {
    # ---- K1: 4-interval fixture, both backends (membership, order not asserted) ----------------------
    "K1_abcd": {
        "cite": "gtars-overlaprs/src/ailist.rs:411-497, bits.rs:570-615",
        "kinds": ["bits", "ailist"],
        "intervals": ABCD,
        "queries": [
            {"q": [2, 4], "set": [[1, 5], [3, 7]]},
            {"q": [9, 11], "set": [[6, 10], [8, 12]]},
            {"q": [13, 15], "set": []},
            {"q": [0, 1], "set": []},
        ],
    },
    "K1_empty": {"cite": "ailist.rs:463-471", "kinds": ["bits", "ailist"], "intervals": [],
                 "queries": [{"q": [1, 5], "set": []}]},
    "K1_single": {"cite": "ailist.rs:531-547", "kinds": ["bits", "ailist"], "intervals": [(5, 10)],
                  "queries": [{"q": [6, 8], "set": [[5, 10]]}, {"q": [11, 15], "set": []}]},
    # ---- K2: nested fixture → 2 components; counts 5 / 3 / 0 ----------------------------------------
    "K2_nested26": {
        "cite": "ailist.rs:549-601",
        "kinds": ["ailist", "bits"],
        "intervals": NESTED26,
        "ailist_components": 2,
        "queries": [{"q": [6, 8], "n": 5}, {"q": [30, 35], "n": 3}, {"q": [101, 150], "n": 0}],
    },
    # ---- K3: Bits doctests ---------------------------------------------------------------------------
    "K3_bits_count": {
        "cite": "bits.rs:131-139, :327-335",
        "intervals": [(x, x + 2) for x in range(0, 100, 5)],
        "queries": [{"q": [5, 11], "count": 2, "find_n": 2}],
    },
    "K3_bits_order": {
        "cite": "bits.rs:190-208 (find order after insert of (0,20,5) into [(0,5,1)...] — insert re-sorts)",
        "intervals_val": [(0, 5, 1), (5, 10, 2), (10, 15, 3), (15, 20, 4), (0, 20, 5)],
        "queries": [{"q": [1, 3], "ordered": [[0, 5, 1], [0, 20, 5]]}],
    },
    # ---- K4: MultiChromOverlapper batch API, each for AIList and Bits --------------------------------
    "K4_mco": [
        {"cite": "multi_chrom_overlapper.rs:715-753", "source": [["chr1", 100, 200], ["chr1", 300, 400], ["chr1", 600, 800]],
         "query": [["chr1", 110, 210]], "min_overlap": None, "count": [1], "find": [[[100, 200]]]},
        {"cite": "multi_chrom_overlapper.rs:878-902 (half-open boundary)", "source": [["chr1", 100, 200]],
         "query": [["chr1", 200, 300]], "min_overlap": None, "count": [0], "any": [False]},
        {"cite": "multi_chrom_overlapper.rs:1070-1083", "source": [["chr1", 150, 200], ["chr1", 250, 350], ["chr1", 500, 600]],
         "query": [["chr1", 100, 300]], "min_overlap": None, "count": [2]},
        {"cite": "multi_chrom_overlapper.rs:1085-1097", "source": [["chr1", 150, 250]],
         "query": [["chr1", 100, 200], ["chr1", 300, 400]], "min_overlap": None, "any": [True, False]},
        {"cite": "multi_chrom_overlapper.rs:1099-1116", "source": [["chr1", 50, 150], ["chr1", 200, 250], ["chr1", 400, 500]],
         "query": [["chr1", 100, 300]], "min_overlap": None, "find": [[[50, 150], [200, 250]]]},
        {"cite": "multi_chrom_overlapper.rs:1118-1130 (min_overlap 5)", "source": [["chr1", 100, 110]],
         "query": [["chr1", 105, 200]], "min_overlap": 5, "count": [1]},
        {"cite": "multi_chrom_overlapper.rs:1118-1130 (min_overlap 6)", "source": [["chr1", 100, 110]],
         "query": [["chr1", 105, 200]], "min_overlap": 6, "count": [0], "any": [False]},
        {"cite": "multi_chrom_overlapper.rs:1132-1144 (empty query)", "source": [["chr1", 100, 200]],
         "query": [], "min_overlap": None, "count": [], "any": []},
        {"cite": "multi_chrom_overlapper.rs:1146-1157 (empty index)", "source": [],
         "query": [["chr1", 100, 200]], "min_overlap": None, "count": [0], "any": [False]},
        {"cite": "multi_chrom_overlapper.rs:1044-1068 (unknown chrom contributes nothing)",
         "source": [["chr1", 100, 200], ["chr1", 300, 400], ["chr2", 500, 600]],
         "query": [["chr1", 150, 250], ["chr2", 550, 650], ["chr9", 1, 1000]], "min_overlap": None,
         "count": [1, 1, 0], "find": [[[100, 200]], [[500, 600]], []]},
        # K11: binding-level batch ops (index = other, queries = self)
        {"cite": "gtars-python/tests/test_regionset.py:37-55", "source": [["chr1", 150, 250], ["chr1", 550, 650]],
         "query": [["chr1", 100, 200], ["chr1", 300, 400], ["chr1", 500, 600]], "min_overlap": None,
         "count": [1, 0, 1], "any": [True, False, True], "find_idx": [[0], [], [1]]},
    ],
    # ---- K5/K6: tokenizer ------------------------------------------------------------------------------
    "K5_tokenizer": [
        {"cite": "tokenizer.rs:294-330 (vocab = 25 + 7)", "universe": "tokenizers/peaks.bed", "kinds": ["bits", "ailist"],
         "vocab_size": 32, "unk_id": 25},
        {"cite": "tokenizer.rs:312-318 (gz universe)", "universe": "tokenizers/peaks.bed.gz", "kinds": ["bits"],
         "vocab_size": 32, "unk_id": 25},
        {"cite": "tokenizer.rs:360-379 (no overlap → <unk>); test_tokenizers.py:110,120 (unk id 25)",
         "universe": "tokenizers/peaks.bed", "kinds": ["bits", "ailist"], "regions": [["chr1", 50, 150]],
         "tokens": ["<unk>"], "ids": [25]},
        {"cite": "tokenizer.rs:381-399 (unknown chrom → 1 token)", "universe": "tokenizers/peaks.bed",
         "kinds": ["bits", "ailist"], "regions": [["chr999", 50, 150]], "tokens": ["<unk>"], "ids": [25]},
        {"cite": "tokenizer.rs:401-467 (two chroms; Bits and AIList configs)", "universe": "tokenizers/peaks.bed",
         "kinds": ["bits", "ailist"], "regions": [["chr1", 151399441, 151399547], ["chr2", 203871220, 203871381]],
         "tokens": ["chr1:151399431-151399527", "chr2:203871200-203871375"], "ids": [6, 7]},
        {"cite": "tokenizer.rs:470-496 (multi overlap, ordered, Bits)", "universe": "tokenizers/peaks.bed",
         "kinds": ["bits"], "regions": [["chr2", 203871346, 203871616]],
         "tokens": ["chr2:203871200-203871375", "chr2:203871387-203871588"], "ids": [7, 8]},
        {"cite": "gtars-python/tests/test_tokenizers.py:181-230 (peaks.scored.bed)", "universe": "tokenizers/peaks.scored.bed",
         "kinds": ["bits"], "regions": [["chr9", 3526071, 3526165]], "ids": [11]},
        {"cite": "gtars-python/tests/test_tokenizers.py:181-230 (peaks.scored.bed)", "universe": "tokenizers/peaks.scored.bed",
         "kinds": ["bits"], "regions": [["chr9", 3526178, 3526249]], "ids": [10]},
    ],
    # ---- K7: IGD unit KATs: ops are ("add", chr, s, e, value, file) / queries (chr, s, e, m) → per-file hits
    "K7_igd": [
        {"cite": "igd.rs:914-959", "adds": [["chr1", 100, 200, 0, 0], ["chr1", 300, 400, 0, 0], ["chr1", 150, 250, 0, 1]],
         "n_files": 2, "queries": [{"q": ["chr1", 120, 180, 1], "total": 2, "hits": [1, 1]},
                                   {"q": ["chr1", 350, 380, 1], "total": 1, "hits": [1, 0]},
                                   {"q": ["chr1", 500, 600, 1], "total": 0, "hits": [0, 0]}]},
        {"cite": "igd.rs:961-986 (min_overlap 1/10/11)", "adds": [["chr1", 100, 200, 0, 0]], "n_files": 1,
         "queries": [{"q": ["chr1", 190, 250, 1], "hits": [1]}, {"q": ["chr1", 190, 250, 10], "hits": [1]},
                     {"q": ["chr1", 190, 250, 11], "hits": [0]}]},
        {"cite": "igd.rs:988-1016 (multi-tile spanning counted once)", "adds": [["chr1", 10000, 20000, 0, 0]], "n_files": 1,
         "queries": [{"q": ["chr1", 11000, 12000, 1], "hits": [1]}, {"q": ["chr1", 17000, 18000, 1], "hits": [1]},
                     {"q": ["chr1", 15000, 19000, 1], "hits": [1]}]},
        {"cite": "igd.rs:1018-1032 (unknown chrom)", "adds": [["chr1", 100, 200, 0, 0]], "n_files": 1,
         "queries": [{"q": ["chrZ", 100, 200, 1], "total": 0, "hits": [0]}]},
        {"cite": "igd.rs:1060-1084 (from_region_sets)", "adds": [["chr1", 100, 200, 0, 0], ["chr1", 300, 400, 0, 0], ["chr1", 150, 350, 0, 1]],
         "n_files": 2, "queries": [{"q": ["chr1", 160, 170, 1], "hits": [1, 1]}]},
        {"cite": "igd.rs:1200-1221 (pairwise: 3 hits from one query)",
         "adds": [["chr1", 100, 200, 0, 0], ["chr1", 120, 220, 0, 0], ["chr1", 140, 240, 0, 0]], "n_files": 1,
         "queries": [{"q": ["chr1", 150, 190, 1], "hits": [3]}]},
        {"cite": "igd.rs:1400-1419 (negative adds skipped)",
         "adds": [["chr1", -100, 200, 0, 0], ["chr1", 100, -200, 0, 0], ["chr1", -100, -50, 0, 0], ["chr1", 100, 200, 0, 0]],
         "n_files": 1, "queries": [{"q": ["chr1", 150, 160, 1], "total": 1, "hits": [1]}]},
        {"cite": "igd.rs:1421-1437 (negative query start clamps; both negative → 0)", "adds": [["chr1", 100, 200, 0, 0]],
         "n_files": 1, "queries": [{"q": ["chr1", -50, 150, 1], "total": 1, "hits": [1]},
                                   {"q": ["chr1", -100, -50, 1], "total": 0, "hits": [0]}]},
        {"cite": "igd.rs:1449-1461 (coordinates > 321M)", "adds": [["chr1", 400000000, 400001000, 0, 0]], "n_files": 1,
         "queries": [{"q": ["chr1", 400000500, 400000600, 1], "total": 1, "hits": [1]}]},
    ],
    "K7_igd_sets": [
        {"cite": "igd.rs:1160-1198 (count_set_overlaps)", "db": [[["chr1", 100, 200], ["chr1", 500, 600]], [["chr1", 150, 250]]],
         "query": [["chr1", 120, 180], ["chr1", 520, 560]], "min_overlap": 1, "set_overlaps": [2, 1]},
        {"cite": "igd.rs:1034-1058 (igd_file_list_01 self-query = 8)", "db_files": ["igd_file_list_01/igd_bed_file_1.bed"],
         "query": [["chr1", 1, 100], ["chr1", 200, 300], ["chr1", 32768, 32868], ["chr1", 49152, 49352], ["chr2", 1, 100],
                   ["chr2", 200, 300], ["chr3", 32768, 32868], ["chr3", 49152, 49352]], "min_overlap": 1, "set_overlaps": [8]},
        {"cite": "igd.rs:1279-1296 (find_overlaps_regionset / count_overlaps_per_query: single-set IGD)",
         "db": [[["chr1", 100, 200], ["chr1", 150, 250], ["chr1", 500, 600]]],
         "query": [["chr1", 160, 180], ["chr1", 550, 580], ["chr1", 700, 800]], "min_overlap": 1, "per_query": [2, 1, 0]},
    ],
    # ---- K7b: two-set IGD queries (from_single_region_set) -------------------------------------------------
    "K7_igd_single": [
        {"cite": "igd.rs:1244-1262", "subject": [["chr1", 100, 200], ["chr1", 300, 400], ["chr1", 500, 600]],
         "query": [["chr1", 150, 350], ["chr1", 550, 650], ["chr1", 700, 800]], "min_overlap": 1, "pairs": [[0, 0], [0, 1], [1, 2]]},
        {"cite": "igd.rs:1264-1272", "subject": [["chr1", 100, 200]], "query": [["chr1", 300, 400]], "min_overlap": 1, "pairs": []},
        {"cite": "igd.rs:1274-1286 (10 bp overlap, min_overlap 1)", "subject": [["chr1", 100, 200]], "query": [["chr1", 190, 300]],
         "min_overlap": 1, "pairs": [[0, 0]]},
        {"cite": "igd.rs:1274-1286 (10 bp overlap, min_overlap 50)", "subject": [["chr1", 100, 200]], "query": [["chr1", 190, 300]],
         "min_overlap": 50, "pairs": []},
        {"cite": "igd.rs:1288-1305 (multi chrom)", "subject": [["chr1", 100, 200], ["chr2", 100, 200]],
         "query": [["chr1", 150, 180], ["chr2", 150, 180], ["chr3", 150, 180]], "min_overlap": 1, "pairs": [[0, 0], [1, 1]]},
        {"cite": "igd.rs:1307-1325 (count_overlaps_per_query)", "subject": [["chr1", 100, 200], ["chr1", 150, 250], ["chr1", 500, 600]],
         "query": [["chr1", 160, 180], ["chr1", 550, 580], ["chr1", 700, 800]], "min_overlap": 1, "per_query": [2, 1, 0]},
        {"cite": "igd.rs:1337-1355 (multi-tile subject counted once)", "subject": [["chr1", 10000, 40000]],
         "query": [["chr1", 15000, 35000]], "min_overlap": 1, "pairs": [[0, 0]], "per_query": [1]},
    ],
    # ---- K9: LOLA contingency counts -------------------------------------------------------------------
    "K9_lola": [
        {"cite": "enrichment.rs:879-922 (a,b,c,d = 1,1,2,6)", "db": [[["chr1", 100, 200], ["chr1", 300, 400]]],
         "user": [[["chr1", 150, 180], ["chr1", 500, 600], ["chr1", 700, 800]]],
         "universe": [["chr1", 50, 250], ["chr1", 250, 450], ["chr1", 450, 550], ["chr1", 550, 650], ["chr1", 650, 750],
                      ["chr1", 750, 850], ["chr1", 850, 950], ["chr1", 950, 1050], ["chr1", 1050, 1150], ["chr1", 1150, 1250]],
         "min_overlap": 1, "abcd": [[[1, 1, 2, 6]]]},
        {"cite": "enrichment.rs:830-853 (binary support)", "db": [[["chr1", 100, 200], ["chr1", 120, 220], ["chr1", 140, 240]]],
         "user": [[["chr1", 150, 190]]], "universe": [["chr1", 50, 300], ["chr1", 400, 500]], "min_overlap": 1,
         "support": [[1]]},
        {"cite": "enrichment.rs:1041-1074 (min_overlap 10)", "db": [[["chr1", 100, 200]]], "user": [[["chr1", 190, 210]]],
         "universe": [["chr1", 0, 500], ["chr1", 600, 700], ["chr1", 800, 900]], "min_overlap": 10, "support": [[1]]},
        {"cite": "enrichment.rs:1041-1074 (min_overlap 11)", "db": [[["chr1", 100, 200]]], "user": [[["chr1", 190, 210]]],
         "universe": [["chr1", 0, 500], ["chr1", 600, 700], ["chr1", 800, 900]], "min_overlap": 11, "support": [[0]]},
        {"cite": "enrichment.rs:1076-1104 (negative b passes through: a=3, b=1-3=-2, c=0, d=2-3+2-0=1)",
         "db": [[["chr1", 100, 200]]], "user": [[["chr1", 110, 120], ["chr1", 130, 140], ["chr1", 150, 160]]],
         "universe": [["chr1", 110, 120], ["chr1", 500, 600]], "min_overlap": 1, "abcd": [[[3, -2, 0, 1]]]},
    ],
    # ---- K8: IGD old≡new differential inputs.  The reference asserts only old == new (no literals); the
    # legacy disk searcher is out of scope, so these are replayed as all-vs-all regression inputs whose
    # expected values come from a brute-force O(n·m) overlap count inside the test ("derived").
    "K8_inputs": {
        "cite": "gtars-igd/src/lib.rs:389-613",
        "dbs": {
            "igd_file_list_01": ["igd_file_list_01/igd_bed_file_1.bed"],
            "igd_file_list_02": ["igd_file_list_02/igd_bed_file_1.bed", "igd_file_list_02/igd_bed_file_2.bed"],
            "lola_multi_db": ["lola_multi_db/collection1/regions/cpgIslandExt.bed",
                              "lola_multi_db/collection1/regions/laminB1Lads.bed",
                              "lola_multi_db/collection1/regions/vistaEnhancers.bed",
                              "lola_multi_db/collection2/regions/cpgIslandExt.bed",
                              "lola_multi_db/collection2/regions/laminB1Lads.bed",
                              "lola_multi_db/collection2/regions/vistaEnhancers.bed"],
        },
        "queries": ["igd_query_files/query1.bed", "igd_query_files/query2.bed"],
    },
    # ---- K10: gtars-scoring fragments x consensus count matrix (ATAC shifts incl. the reversed end interval) ----
    "K10_scoring": {
        "cite": "gtars-scoring/src/fragment_scoring.rs:180-210 (test_region_scoring_from_fragments_atac)",
        "consensus": "consensus/consensus1.bed",
        "fragment_files": ["fragments/region_scoring/fragments1.bed.gz", "fragments/region_scoring/fragments2.bed.gz"],
        "mode": "atac", "rows": 2, "cols": 4, "matrix": [[2, 2, 1, 3], [4, 1, 3, 1]],
    },
    # ---- K12: IndexedRegionSet index-returning queries -------------------------------------------------------
    "K12_indexed_region_set": {
        "cite": "gtars-overlaprs/src/indexed_region_set.rs:396-545",
        "cases": [
            {"name": "intersect_all", "reference": [["chr1", 100, 200], ["chr1", 300, 400], ["chr2", 500, 600]],
             "query": [["chr1", 150, 250], ["chr2", 550, 650]], "intersect_all": [["chr1", 100, 200], ["chr2", 500, 600]]},
            {"name": "count", "reference": [["chr1", 100, 200], ["chr1", 150, 250], ["chr1", 300, 400]],
             "query": [["chr1", 180, 220]], "count": [2]},
            {"name": "any", "reference": [["chr1", 100, 200]], "query": [["chr1", 150, 250], ["chr1", 300, 400]],
             "any": [True, False]},
            {"name": "find", "reference": [["chr1", 100, 200], ["chr1", 300, 400]], "query": [["chr1", 150, 350]],
             "find": [[0, 1]]},
            {"name": "empty_reference", "reference": [], "query": [["chr1", 100, 200]], "count": [0], "any": [False],
             "find": [[]], "intersect_all": []},
            {"name": "empty_query", "reference": [["chr1", 100, 200]], "query": [], "count": [], "any": [], "find": [],
             "intersect_all": []},
            {"name": "multi_chrom", "reference": [["chr1", 100, 200], ["chr2", 100, 200], ["chr3", 100, 200]],
             "query": [["chr1", 150, 250], ["chr2", 150, 250], ["chr4", 150, 250]], "count": [1, 1, 0],
             "any": [True, True, False]},
        ],
    },
    # ---- K13: .gtok files shipped with the reference's test data (bytes as hex) ---------------------------------
    "K13_gtok": {
        "cite": "gtars-io/src/gtok.rs:126-210, consts.rs (GTOK header, 0x01 = u16, 0x02 = u32); tests/data/out/*.gtok",
        "files": {name: {"hex": open(os.path.join(REF, "out", name), "rb").read().hex()}
                  for name in ("tokens.gtok", "peaks.gtok", "to_tokenize.gtok")},
        "tokens": {"tokens.gtok": [42, 101, 999], "peaks.gtok": list(range(25)), "to_tokenize.gtok": [22, 23, 24, 26]},
    },
    # ---- derived vectors (NOT asserted by the reference; regression pins only) -------------------------
    "D_derived": {
        "D1": {"note": "to_tokenize.bed (sorted by RegionSet::try_from) vs peaks.bed", "universe": "tokenizers/peaks.bed",
               "query_file": "to_tokenize.bed", "bits": [22, 23, 24], "ailist": [22, 24, 23]},
        "D2": {"note": "multi-overlap query under AIList is the reverse of Bits", "universe": "tokenizers/peaks.bed",
               "regions": [["chr2", 203871346, 203871616]], "bits": [7, 8], "ailist": [8, 7]},
        "D3": {"note": "fragments1.bed.gz vs consensus1.bed, unk = 4", "universe": "consensus/consensus1.bed",
               "fragments": "fragments/region_scoring/fragments1.bed.gz",
               "expect": {"AAACGCAAGCAAAGGGATGCCA": [0, 4, 1, 0, 3, 3], "AAACGCAAGCAACTGCGTCTTT": [0, 2]}},
    },
}


def main():
    fixtures = {rel: read_fixture(rel) for rel in FIXTURE_FILES}
    with open(os.path.join(HERE, "fixtures.json"), "w") as f:
        json.dump(fixtures, f, indent=1, sort_keys=True)
    with open(os.path.join(HERE, "kats.json"), "w") as f:
        json.dump(KATS, f, indent=1)
    print(f"wrote {len(fixtures)} fixtures, {len(KATS)} KAT groups")


if __name__ == "__main__":
    main()

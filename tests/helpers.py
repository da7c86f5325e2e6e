"""Shared test helpers: golden loading, fixture materialisation, chromosome-id mapping, brute force."""
from __future__ import annotations

import gzip
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
KINDS = {"bits": 0, "ailist": 1}
UNKNOWN = 0xFFFFFFFF


def load_golden():
    with open(os.path.join(HERE, "golden", "fixtures.json")) as f:
        fixtures = json.load(f)
    with open(os.path.join(HERE, "golden", "kats.json")) as f:
        kats = json.load(f)
    return fixtures, kats


def materialise_fixtures(root, fixtures):
    root = str(root)
    for rel, item in fixtures.items():
        p = os.path.join(root, rel)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        if item["gz"]:
            with gzip.open(p, "wt") as f:
                f.write(item["text"])
        else:
            with open(p, "w") as f:
                f.write(item["text"])
    return root


class ChromMap:
    """Dense chromosome ids in first-appearance order; unknown names → UNKNOWN."""

    def __init__(self):
        self.ids = {}

    def add(self, name):
        return self.ids.setdefault(name, len(self.ids))

    def get(self, name):
        return self.ids.get(name, UNKNOWN)

    def __len__(self):
        return len(self.ids)


def flatten_source(regions, cmap=None):
    """[(chr,s,e)] in insertion order → (ChromMap, chrom_offsets, starts, ends, vals) grouped by chromosome,
    insertion order kept inside each chromosome, val = original index."""
    cmap = cmap or ChromMap()
    for r in regions:
        cmap.add(r[0])
    per = [[] for _ in range(len(cmap))]
    for i, r in enumerate(regions):
        per[cmap.get(r[0])].append((r[1], r[2], i))
    offs = [0]
    s, e, v = [], [], []
    for lst in per:
        for a, b, i in lst:
            s.append(a)
            e.append(b)
            v.append(i)
        offs.append(len(s))
    return (cmap, np.array(offs, dtype=np.uint64), np.array(s, dtype=np.uint32), np.array(e, dtype=np.uint32),
            np.array(v, dtype=np.uint32))


def flatten_queries(regions, cmap):
    c = np.array([cmap.get(r[0]) for r in regions], dtype=np.uint32)
    s = np.array([r[1] for r in regions], dtype=np.uint32)
    e = np.array([r[2] for r in regions], dtype=np.uint32)
    return c, s, e


def flatten_sets(sets, cmap, add=False):
    """list of [(chr,s,e)] → (set_offsets, chr, start, end)."""
    offs = [0]
    flat = []
    for st in sets:
        flat.extend(st)
        offs.append(len(flat))
    if add:
        for r in flat:
            cmap.add(r[0])
    c, s, e = flatten_queries(flat, cmap)
    return np.array(offs, dtype=np.uint64), c, s, e


def parse_bed_text(text):
    """Minimal BED3 reader for fixture text (file order, no sorting)."""
    out = []
    for line in text.splitlines():
        if not line or line.startswith(("#", "track", "browser")):
            continue
        p = line.split("\t")
        out.append((p[0], int(p[1]), int(p[2])))
    return out


def brute_overlap_matrix(q_chr, q_s, q_e, d_chr, d_s, d_e, min_overlap=1):
    """bool [nq, nd]: min(qe,de) - max(qs,ds) >= m on the same chromosome (IGD closed form, m >= 1)."""
    q_s = np.asarray(q_s, dtype=np.int64)[:, None]
    q_e = np.asarray(q_e, dtype=np.int64)[:, None]
    d_s = np.asarray(d_s, dtype=np.int64)[None, :]
    d_e = np.asarray(d_e, dtype=np.int64)[None, :]
    same = np.asarray(q_chr)[:, None] == np.asarray(d_chr)[None, :]
    return same & ((np.minimum(q_e, d_e) - np.maximum(q_s, d_s)) >= min_overlap)

"""ctypes driver for the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing under gtars_b200/ may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

BITS, AILIST = 0, 1
UNKNOWN_CHROM = 0xFFFFFFFF


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "gtars_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    vp, u32, u64, i32, i64, cint = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int32, C.c_int64, C.c_int
    sig = {
        "orc_last_error": (C.c_char_p, []),
        "orc_overlapper_build": (vp, [cint, u64, vp, vp, vp]),
        "orc_overlapper_free": (None, [vp]),
        "orc_overlapper_len": (u64, [vp]),
        "orc_overlapper_find": (u64, [vp, u32, u32, vp, u64]),
        "orc_bits_count": (u64, [vp, u32, u32]),
        "orc_ailist_num_components": (u64, [vp]),
        "orc_index_build": (vp, [cint, u32, vp, vp, vp, vp]),
        "orc_index_free": (None, [vp]),
        "orc_index_count": (None, [vp, u64, vp, vp, vp, i32, vp, cint]),
        "orc_index_bits_count": (None, [vp, u64, vp, vp, vp, vp, cint]),
        "orc_index_any": (None, [vp, u64, vp, vp, vp, i32, vp]),
        "orc_index_find": (vp, [vp, u64, vp, vp, vp, i32, vp, cint]),
        "orc_buf_data": (vp, [vp]),
        "orc_buf_len": (u64, [vp]),
        "orc_buf_free": (None, [vp]),
        "orc_tokenize_files": (vp, [vp, u64, vp, vp, vp, vp, u32, vp, vp, cint]),
        "orc_tokenize_fragments": (vp, [vp, u64, vp, vp, vp, vp, u32, u32, vp, vp]),
        "orc_score_matrix": (None, [vp, u64, vp, vp, vp, vp, cint, u64, vp, cint]),
        "orc_score_barcodes": (vp, [vp, u64, vp, vp, vp, vp, u32, vp]),
        "orc_barcode_scoring_file": (vp, [vp, C.c_char_p]), "orc_bscores_free": (None, [vp]), "orc_bscores_n": (u64, [vp]),
        "orc_bscores_barcode": (C.c_char_p, [vp, u64]), "orc_bscores_len": (u64, [vp, u64]), "orc_bscores_pairs": (vp, [vp, u64]),
        "orc_consensus_from_bed": (vp, [C.c_char_p]),
        "orc_consensus_free": (None, [vp]),
        "orc_consensus_len": (u64, [vp]),
        "orc_region_scoring_files": (cint, [vp, u64, vp, cint, vp]),
        "orc_igd_save": (cint, [vp, u64, vp, C.c_char_p]),
        "orc_igd_from_file": (vp, [C.c_char_p]),
        "orc_igd_n_contigs": (u64, [vp]),
        "orc_igd_contig_name": (C.c_char_p, [vp, u64]),
        "orc_igd_n_files": (u64, [vp]),
        "orc_regionset_from_file": (vp, [C.c_char_p]),
        "orc_regionset_free": (None, [vp]),
        "orc_regionset_len": (u64, [vp]),
        "orc_regionset_chr": (C.c_char_p, [vp, u64]),
        "orc_regionset_start": (u32, [vp, u64]),
        "orc_regionset_end": (u32, [vp, u64]),
        "orc_tokenizer_from_bed": (vp, [C.c_char_p, cint]),
        "orc_tokenizer_free": (None, [vp]),
        "orc_tokenizer_vocab_size": (u64, [vp]),
        "orc_tokenizer_token_to_id": (i64, [vp, C.c_char_p]),
        "orc_tokenizer_id_to_token": (C.c_char_p, [vp, u32]),
        "orc_tokenizer_encode": (vp, [vp, u64, vp, vp, vp]),
        "orc_tokenize_fragment_file": (vp, [vp, C.c_char_p]),
        "orc_fragres_free": (None, [vp]),
        "orc_fragres_n_barcodes": (u64, [vp]),
        "orc_fragres_barcode": (C.c_char_p, [vp, u64]),
        "orc_fragres_len": (u64, [vp, u64]),
        "orc_fragres_ids": (vp, [vp, u64]),
        "orc_igd_build": (vp, [u64, vp, vp, vp, vp]),
        "orc_igd_new": (vp, []),
        "orc_igd_set_n_files": (None, [vp, u64]),
        "orc_igd_add": (None, [vp, u32, i32, i32, i32, u32]),
        "orc_igd_finalize": (None, [vp]),
        "orc_igd_free": (None, [vp]),
        "orc_igd_count_overlaps": (u32, [vp, u32, i32, i32, i32, vp]),
        "orc_igd_count_set_overlaps": (None, [vp, u64, vp, vp, vp, vp, i32, vp, cint]),
        "orc_igd_count_region_hits": (None, [vp, u64, vp, vp, vp, vp, i32, vp, cint]),
        "orc_lola_tables": (None, [u64, u64, vp, vp, vp, u64, vp]),
        "orc_max_threads": (cint, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _take_buf(h) -> np.ndarray:
    L = lib()
    n = L.orc_buf_len(h)
    out = np.empty(n, dtype=np.uint32)
    if n:
        C.memmove(out.ctypes.data, L.orc_buf_data(h), n * 4)
    L.orc_buf_free(h)
    return out


def _err() -> str:
    return lib().orc_last_error().decode()


class Overlapper:
    """One Bits / AIList over a single coordinate space (bits.rs / ailist.rs)."""

    def __init__(self, kind, intervals):
        """intervals: iterable of (start, end) or (start, end, val); val defaults to insertion index."""
        ivs = list(intervals)
        self.kind = kind
        s = _u32([iv[0] for iv in ivs])
        e = _u32([iv[1] for iv in ivs])
        v = _u32([iv[2] if len(iv) > 2 else i for i, iv in enumerate(ivs)])
        self._h = lib().orc_overlapper_build(kind, len(ivs), _ptr(s), _ptr(e), _ptr(v))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_overlapper_free(self._h)
            self._h = None

    def __len__(self):
        return lib().orc_overlapper_len(self._h)

    def find(self, start, end):
        L = lib()
        cap = max(len(self), 1)
        out = np.empty(3 * cap, dtype=np.uint32)
        n = L.orc_overlapper_find(self._h, start, end, _ptr(out), cap)
        return [tuple(int(x) for x in out[3 * i:3 * i + 3]) for i in range(n)]

    def count(self, start, end):
        assert self.kind == BITS
        return lib().orc_bits_count(self._h, start, end)

    def num_components(self):
        assert self.kind == AILIST
        return lib().orc_ailist_num_components(self._h)


class Index:
    """Multi-chromosome index over dense chromosome ids (multi_chrom_overlapper.rs batch API)."""

    def __init__(self, kind, chrom_offsets, starts, ends, vals=None):
        self.kind = kind
        co = _u64(chrom_offsets)
        self.n_chroms = len(co) - 1
        s, e = _u32(starts), _u32(ends)
        v = _u32(vals) if vals is not None else None
        self._h = lib().orc_index_build(kind, self.n_chroms, _ptr(co), _ptr(s), _ptr(e), _ptr(v))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_index_free(self._h)
            self._h = None

    def count(self, chr, start, end, min_overlap=0, threads=1):
        chr, start, end = _u32(chr), _u32(start), _u32(end)
        out = np.empty(len(chr), dtype=np.uint32)
        lib().orc_index_count(self._h, len(chr), _ptr(chr), _ptr(start), _ptr(end), min_overlap, _ptr(out), threads)
        return out

    def bits_count(self, chr, start, end, threads=1):
        assert self.kind == BITS
        chr, start, end = _u32(chr), _u32(start), _u32(end)
        out = np.empty(len(chr), dtype=np.uint64)
        lib().orc_index_bits_count(self._h, len(chr), _ptr(chr), _ptr(start), _ptr(end), _ptr(out), threads)
        return out

    def any(self, chr, start, end, min_overlap=0):
        chr, start, end = _u32(chr), _u32(start), _u32(end)
        out = np.empty(len(chr), dtype=np.uint8)
        lib().orc_index_any(self._h, len(chr), _ptr(chr), _ptr(start), _ptr(end), min_overlap, _ptr(out))
        return out.astype(bool)

    def find(self, chr, start, end, min_overlap=0, coords=False):
        """Returns (offsets[n+1], vals) or, with coords, (offsets, triples[n_hits,3])."""
        chr, start, end = _u32(chr), _u32(start), _u32(end)
        off = np.empty(len(chr) + 1, dtype=np.uint64)
        h = lib().orc_index_find(self._h, len(chr), _ptr(chr), _ptr(start), _ptr(end), min_overlap, _ptr(off),
                                 1 if coords else 0)
        buf = _take_buf(h)
        return (off, buf.reshape(-1, 3)) if coords else (off, buf)

    def tokenize_files(self, file_offsets, chr, start, end, unk_id, remap=None, threads=1):
        fo = _u64(file_offsets)
        chr, start, end = _u32(chr), _u32(start), _u32(end)
        rm = _u32(remap) if remap is not None else None
        out_off = np.empty(len(fo), dtype=np.uint64)
        h = lib().orc_tokenize_files(self._h, len(fo) - 1, _ptr(fo), _ptr(chr), _ptr(start), _ptr(end), unk_id,
                                     _ptr(rm), _ptr(out_off), threads)
        return out_off, _take_buf(h)

    def tokenize_fragments(self, chr, start, end, barcode, n_barcodes, unk_id, remap=None):
        chr, start, end, barcode = _u32(chr), _u32(start), _u32(end), _u32(barcode)
        rm = _u32(remap) if remap is not None else None
        out_off = np.empty(n_barcodes + 1, dtype=np.uint64)
        h = lib().orc_tokenize_fragments(self._h, len(chr), _ptr(chr), _ptr(start), _ptr(end), _ptr(barcode),
                                         n_barcodes, unk_id, _ptr(rm), _ptr(out_off))
        return out_off, _take_buf(h)


SCORE_ATAC, SCORE_CHIP = 0, 1


def score_matrix(index, file_offsets, chr, start, end, mode, n_cols, threads=1):
    """region_scoring_from_fragments over dense-id fragments: uint32 [n_files, n_cols]."""
    fo, chr, start, end = _u64(file_offsets), _u32(chr), _u32(start), _u32(end)
    n_files = len(fo) - 1
    out = np.empty((n_files, n_cols), dtype=np.uint32)
    lib().orc_score_matrix(index._h, n_files, _ptr(fo), _ptr(chr), _ptr(start), _ptr(end), mode, n_cols, _ptr(out), threads)
    return out


def score_barcodes(index, chr, start, end, barcode, n_barcodes):
    """barcode_scoring_from_fragments: (offsets[n_barcodes+1], peaks, counts), sorted by (barcode, peak)."""
    chr, start, end, barcode = _u32(chr), _u32(start), _u32(end), _u32(barcode)
    out_off = np.empty(n_barcodes + 1, dtype=np.uint64)
    h = lib().orc_score_barcodes(index._h, len(chr), _ptr(chr), _ptr(start), _ptr(end), _ptr(barcode), n_barcodes, _ptr(out_off))
    pairs = _take_buf(h).reshape(-1, 2)
    return out_off, pairs[:, 0].copy(), pairs[:, 1].copy()


def region_scoring_files(consensus_path, fragment_paths, mode=SCORE_ATAC):
    """ConsensusSet::new(path) + region_scoring_from_fragments over the files, in the given order."""
    L = lib()
    h = L.orc_consensus_from_bed(os.fsencode(consensus_path))
    if not h:
        raise ValueError(_err())
    try:
        cols = L.orc_consensus_len(h)
        out = np.empty((len(fragment_paths), cols), dtype=np.uint32)
        arr = (C.c_char_p * len(fragment_paths))(*[os.fsencode(p) for p in fragment_paths])
        if L.orc_region_scoring_files(h, len(fragment_paths), arr, mode, _ptr(out)) != 0:
            raise ValueError(_err())
        return out
    finally:
        L.orc_consensus_free(h)


def barcode_scoring_file(consensus_path, fragment_path):
    """barcode_scoring_from_fragments (fragment_scoring.rs:126-155) with the reference's key set: {barcode: {peak: count}};
    a barcode whose fragments lie on consensus chromosomes but overlap nothing maps to {} (see gtars_oracle.cpp)."""
    L = lib()
    h = L.orc_consensus_from_bed(os.fsencode(consensus_path))
    if not h:
        raise ValueError(_err())
    try:
        r = L.orc_barcode_scoring_file(h, os.fsencode(fragment_path))
        if not r:
            raise ValueError(_err())
        out = {}
        for i in range(L.orc_bscores_n(r)):
            k = L.orc_bscores_len(r, i)
            pairs = np.ctypeslib.as_array(C.cast(L.orc_bscores_pairs(r, i), C.POINTER(C.c_uint32)), shape=(k,)).copy() if k else np.zeros(0, np.uint32)
            out[L.orc_bscores_barcode(r, i).decode()] = {int(p): int(c) for p, c in pairs.reshape(-1, 2)}
        L.orc_bscores_free(r)
        return out
    finally:
        L.orc_consensus_free(h)


def regionset_from_file(path):
    """RegionSet::try_from(path): parsed and sorted by (chr string, start). Returns [(chr,start,end)]."""
    L = lib()
    h = L.orc_regionset_from_file(os.fsencode(path))
    if not h:
        raise ValueError(_err())
    try:
        n = L.orc_regionset_len(h)
        return [(L.orc_regionset_chr(h, i).decode(), L.orc_regionset_start(h, i), L.orc_regionset_end(h, i))
                for i in range(n)]
    finally:
        L.orc_regionset_free(h)


class Tokenizer:
    """Tokenizer::from_bed + encode/convert (tokenizer.rs)."""

    def __init__(self, path, kind=BITS):
        self._h = lib().orc_tokenizer_from_bed(os.fsencode(path), kind)
        if not self._h:
            raise ValueError(_err())

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_tokenizer_free(self._h)
            self._h = None

    def vocab_size(self):
        return lib().orc_tokenizer_vocab_size(self._h)

    def token_to_id(self, tok):
        r = lib().orc_tokenizer_token_to_id(self._h, tok.encode())
        return None if r < 0 else int(r)

    def id_to_token(self, i):
        r = lib().orc_tokenizer_id_to_token(self._h, i)
        return None if r is None else r.decode()

    def encode(self, regions):
        regions = list(regions)
        chrs = (C.c_char_p * len(regions))(*[r[0].encode() for r in regions])
        s = _u32([r[1] for r in regions])
        e = _u32([r[2] for r in regions])
        h = lib().orc_tokenizer_encode(self._h, len(regions), chrs, _ptr(s), _ptr(e))
        return [int(x) for x in _take_buf(h)]

    def tokenize(self, regions):
        return [self.id_to_token(i) for i in self.encode(regions)]

    def tokenize_fragment_file(self, path):
        L = lib()
        h = L.orc_tokenize_fragment_file(self._h, os.fsencode(path))
        if not h:
            raise ValueError(_err())
        try:
            out = {}
            for i in range(L.orc_fragres_n_barcodes(h)):
                n = L.orc_fragres_len(h, i)
                ids = np.empty(n, dtype=np.uint32)
                C.memmove(ids.ctypes.data, L.orc_fragres_ids(h, i), n * 4)
                out[L.orc_fragres_barcode(h, i).decode()] = [int(x) for x in ids]
            return out
        finally:
            L.orc_fragres_free(h)


class Igd:
    """In-memory IGD (igd.rs) over dense chromosome ids."""

    def __init__(self, file_offsets=None, chr=None, start=None, end=None):
        L = lib()
        if file_offsets is None:
            self._h = L.orc_igd_new()
            self.n_files = 0
        else:
            fo = _u64(file_offsets)
            chr, start, end = _u32(chr), _u32(start), _u32(end)
            self.n_files = len(fo) - 1
            self._h = L.orc_igd_build(self.n_files, _ptr(fo), _ptr(chr), _ptr(start), _ptr(end))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_igd_free(self._h)
            self._h = None

    def save(self, path, names_by_chr_id):
        """Igd::save (igd.rs:418-486), the .igd file only; contig i is named names_by_chr_id[chr id]."""
        arr = (C.c_char_p * max(len(names_by_chr_id), 1))(*[n.encode() for n in names_by_chr_id])
        if lib().orc_igd_save(self._h, len(names_by_chr_id), arr, os.fsencode(path)) != 0:
            raise ValueError(_err())

    @classmethod
    def from_igd_file(cls, path):
        """Igd::from_igd_file (igd.rs:320-414): dense chromosome id = contig index; .contig_names lists them."""
        L = lib()
        h = L.orc_igd_from_file(os.fsencode(path))
        if not h:
            raise ValueError(_err())
        self = cls.__new__(cls)
        self._h = h
        self.n_files = int(L.orc_igd_n_files(h))
        self.contig_names = [L.orc_igd_contig_name(h, i).decode() for i in range(L.orc_igd_n_contigs(h))]
        return self

    def add(self, chr, start, end, value, file_idx):
        lib().orc_igd_add(self._h, chr, start, end, value, file_idx)
        self.n_files = max(self.n_files, file_idx + 1)
        lib().orc_igd_set_n_files(self._h, self.n_files)

    def finalize(self):
        lib().orc_igd_finalize(self._h)

    def count_overlaps(self, chr, start, end, min_overlap=1, hits=None):
        if hits is None:
            hits = np.zeros(self.n_files, dtype=np.uint64)
        total = lib().orc_igd_count_overlaps(self._h, chr, start, end, min_overlap, _ptr(hits))
        return total, hits

    def _batch(self, fn, set_offsets, chr, start, end, min_overlap, threads):
        so = _u64(set_offsets)
        chr, start, end = _u32(chr), _u32(start), _u32(end)
        out = np.zeros((len(so) - 1, self.n_files), dtype=np.uint64)
        fn(self._h, len(so) - 1, _ptr(so), _ptr(chr), _ptr(start), _ptr(end), min_overlap, _ptr(out), threads)
        return out

    def count_set_overlaps(self, set_offsets, chr, start, end, min_overlap=1, threads=1):
        return self._batch(lib().orc_igd_count_set_overlaps, set_offsets, chr, start, end, min_overlap, threads)

    def count_region_hits(self, set_offsets, chr, start, end, min_overlap=1, threads=1):
        return self._batch(lib().orc_igd_count_region_hits, set_offsets, chr, start, end, min_overlap, threads)


def lola_tables(user_hits, universe_hits, user_sizes, universe_size):
    """a,b,c,d per (user set, db set): int64 [n_user, n_db, 4] (enrichment.rs:198-220)."""
    uh = np.ascontiguousarray(user_hits, dtype=np.uint64)
    n_user, n_db = uh.shape
    un = _u64(universe_hits)
    us = _u64(user_sizes)
    out = np.empty((n_user, n_db, 4), dtype=np.int64)
    lib().orc_lola_tables(n_user, n_db, _ptr(uh), _ptr(un), _ptr(us), int(universe_size), _ptr(out))
    return out


def max_threads() -> int:
    """Host threads the multithreaded oracle legs use: the CPUs this process may run on.  Deliberately NOT
    omp_get_max_threads(): torch.distributed.run exports OMP_NUM_THREADS=1 to its workers, which made the reference arm
    single-threaded at N > 1 in round 1; every OpenMP region of the oracle takes an explicit num_threads(threads)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)

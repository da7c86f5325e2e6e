"""One process, several GPUs: a multi-device context (gtgpu_init_multi) must return, through the UNCHANGED entry points,
exactly what a single-device context returns (which the other GPU tests pin to the oracle).  Needs >= 2 GPUs; skipped on
a single-GPU box (run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multidev.py`)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctxs():
    from gtars_b200 import ffi
    n = ffi.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    one = ffi.Context(0)
    many = ffi.Context(devices=list(range(min(n, 8))))
    yield one, many
    many.close()
    one.close()


def _universe(nested=0.0, n=60_000):
    from gtars_b200 import synth
    u = synth.make_universe(n, nested_frac=nested)
    offs = u["chrom_offsets"].numpy().astype(np.uint64)
    s, e, v = (u[k].numpy().view(np.uint32) for k in ("g_start", "g_end", "g_val"))
    return u, offs, s, e, v


def test_group_reports_its_devices(ctxs):
    one, many = ctxs
    assert one.devices() == [0]
    assert many.devices() == list(range(len(many.devices()))) and len(many.devices()) >= 2


@pytest.mark.parametrize("kind,nested", [(0, 0.0), (1, 0.02)])
def test_count_find_any_match_single_device(ctxs, kind, nested):
    from gtars_b200 import ffi, synth
    one, many = ctxs
    u, offs, s, e, v = _universe(nested)
    g1, gm = ffi.Index(one, kind, offs, s, e, v), ffi.Index(many, kind, offs, s, e, v)
    assert g1.info() == gm.info()
    q = synth.make_query_files(u, 3, 100_001, unknown_frac_ppm=2000, sort_files=False)
    qc, qs, qe = (q[k].numpy().view(np.uint32) for k in ("chr", "start", "end"))
    for m in (0, 3):
        assert np.array_equal(g1.count(qc, qs, qe, m), gm.count(qc, qs, qe, m))
        assert np.array_equal(g1.any(qc, qs, qe, m), gm.any(qc, qs, qe, m))
        a, b = g1.find(qc, qs, qe, m), gm.find(qc, qs, qe, m)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    if kind == 0:
        assert np.array_equal(g1.bits_count(qc, qs, qe), gm.bits_count(qc, qs, qe))
    # tiny batches (fewer queries than devices, empty) go through too
    for k in (0, 1, 5):
        assert np.array_equal(g1.count(qc[:k], qs[:k], qe[:k]), gm.count(qc[:k], qs[:k], qe[:k]))
        a, b = g1.find(qc[:k], qs[:k], qe[:k]), gm.find(qc[:k], qs[:k], qe[:k])
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    g1.close()
    gm.close()


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("chunk", [None, "65536", "200704"])
def test_tokenize_files_all_entry_points(ctxs, kind, chunk, monkeypatch):
    """Pipelined (chunks dealt round-robin to the devices), plain (files dealt in blocks) and the rare fallbacks (a file
    without a token) all equal the single-device result: ids, order and per-file offsets."""
    from gtars_b200 import ffi, synth
    one, many = ctxs
    u, offs, s, e, v = _universe(0.01 if kind else 0.0)
    g1, gm = ffi.Index(one, kind, offs, s, e, v), ffi.Index(many, kind, offs, s, e, v)
    q = synth.make_query_files(u, 23, 60_000, unknown_frac_ppm=500)
    qc, qs, qe = (q[k].numpy().view(np.uint32) for k in ("chr", "start", "end"))
    fo = q["file_offsets"].numpy().astype(np.uint64)
    if chunk:
        monkeypatch.setenv("GTGPU_PIPE_CHUNK", chunk)
    want = g1.tokenize_files(fo, qc, qs, qe, u["unk_id"])
    got = gm.tokenize_files(fo, qc, qs, qe, u["unk_id"])
    assert np.array_equal(want[0], got[0]) and np.array_equal(want[1], got[1])
    ro, rc, w16, wi, we = ffi.marshal_compact(qc, qs, qe, fo)
    got = gm.tokenize_files_runs(fo, ro, rc, qs, qe, u["unk_id"])
    assert np.array_equal(want[0], got[0]) and np.array_equal(want[1], got[1])
    got = gm.tokenize_files_compact(fo, ro, rc, qs, w16, wi, we, u["unk_id"])
    assert np.array_equal(want[0], got[0]) and np.array_equal(want[1], got[1])
    ro, rc, wb, pk, an, xi, xs, xe = ffi.marshal_packed(qc, qs, qe, fo)
    got = gm.tokenize_files_packed(fo, ro, rc, wb, pk, an, xi, xs, xe, u["unk_id"])
    assert np.array_equal(want[0], got[0]) and np.array_equal(want[1], got[1])
    # ragged files, two of them without any token (all queries on an unknown chromosome): the [unk] rule
    fo2 = np.array([0, 10, 10, 400_000, 400_003, 900_000, len(qc)], dtype=np.uint64)
    qc2 = qc.copy()
    qc2[400_000:400_003] = 0xFFFFFFFF
    want = g1.tokenize_files(fo2, qc2, qs, qe, u["unk_id"])
    got = gm.tokenize_files(fo2, qc2, qs, qe, u["unk_id"])
    assert np.array_equal(want[0], got[0]) and np.array_equal(want[1], got[1])
    assert int(u["unk_id"]) in got[1]
    g1.close()
    gm.close()


def test_fragments_and_scoring(ctxs):
    from gtars_b200 import ffi, synth
    one, many = ctxs
    u, offs, s, e, v = _universe()
    g1, gm = ffi.Index(one, 0, offs, s, e, v), ffi.Index(many, 0, offs, s, e, v)
    q = synth.make_query_files(u, 1, 1_300_007, seed=synth.SEED_FRAGMENTS, sort_files=False, unknown_frac_ppm=800)
    qc, qs, qe = (q[k].numpy().view(np.uint32) for k in ("chr", "start", "end"))
    rng = np.random.default_rng(5)
    n_bc = 3000
    bc = (rng.integers(0, n_bc, len(qc)) ** 2 // n_bc).astype(np.uint32)
    want = g1.tokenize_fragments(qc, qs, qe, bc, n_bc, u["unk_id"])
    got = gm.tokenize_fragments(qc, qs, qe, bc, n_bc, u["unk_id"])
    assert np.array_equal(want[0], got[0]) and np.array_equal(want[1], got[1])
    fo = np.array([0, 5, 5, 700_000, len(qc)], dtype=np.uint64)
    for mode in (ffi.SCORE_ATAC, ffi.SCORE_CHIP):
        assert np.array_equal(g1.score_matrix(fo, qc, qs, qe, mode, int(u["n"])), gm.score_matrix(fo, qc, qs, qe, mode, int(u["n"])))
    g1.close()
    gm.close()


@pytest.mark.parametrize("min_overlap", [1, 30])
def test_igd_sharded_by_set_with_in_process_allgather(ctxs, min_overlap):
    """The LOLA database sharded by region set over the devices of one process + ncclAllGather between them equals the
    single-device matrices (count_set_overlaps and count_region_hits), including a file count the devices do not divide."""
    from gtars_b200 import ffi, synth
    one, many = ctxs
    n_db = 37
    db = synth.make_uniform_intervals(n_db * 3000, synth.SEED_LOLA_DB, min_w=200, max_w=5000)
    sizes = np.full(n_db, 3000)
    dfo = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    dc, ds, de = (db[k].numpy().view(np.uint32) for k in ("chr", "start", "end"))
    q = synth.make_uniform_intervals(6 * 4000, synth.SEED_LOLA_USER, min_w=100, max_w=2000)
    so = np.array([0, 4000, 4000, 12_000, 12_001, 20_000, 24_000], dtype=np.uint64)
    qc, qs, qe = (q[k].numpy().view(np.uint32) for k in ("chr", "start", "end"))
    g1, gm = ffi.Igd(one, dfo, synth.N_CHROMS, dc, ds, de), ffi.Igd(many, dfo, synth.N_CHROMS, dc, ds, de)
    i1, im = g1.info(), gm.info()
    assert (i1["n_files"], i1["n_records"]) == (im["n_files"], im["n_records"])
    assert np.array_equal(g1.count_set_overlaps(so, qc, qs, qe, min_overlap), gm.count_set_overlaps(so, qc, qs, qe, min_overlap))
    assert np.array_equal(g1.count_region_hits(so, qc, qs, qe, min_overlap), gm.count_region_hits(so, qc, qs, qe, min_overlap))
    with pytest.raises(ffi.GtarsGpuError):
        gm.count_region_hits(so, qc, qs, qe, 0)      # rejected on every device, nobody is left waiting in the collective
    assert np.array_equal(g1.count_region_hits(so, qc, qs, qe, min_overlap), gm.count_region_hits(so, qc, qs, qe, min_overlap))
    g1.close()
    gm.close()

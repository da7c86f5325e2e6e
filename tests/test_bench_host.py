"""Host-only checks of bench.py (no GPU): the reference arm, the config both arms share, helpers."""
import importlib.util
import json
import os
import subprocess
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_reference_arm_uses_all_host_threads_even_under_torchrun_env():
    """torch.distributed.run exports OMP_NUM_THREADS=1; the reference arm must still use every CPU it may run on, the same
    bounded sample at every N, and print the config our arm prints (so the driver's same_config check holds)."""
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="4", LOCAL_RANK="0")
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "4", "--files", "6", "--per-file", "3000",
           "--universe", "20000", "--steps", "1", "--warmup", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["n_gpus"] == 4
    assert line["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert line["cpu_baseline"]["kind"] == "port" and line["value"] > 0 and line["e2e"]["value"] == line["value"]
    b = _bench()
    args = types.SimpleNamespace(files=6, per_file=3000, universe=20000, kind="bits", nested=0.0, width_scale=1, scaling="weak")
    assert line["config"] == b.workload_config(args, 4)
    # other ranks print nothing and exit 0
    env["RANK"] = "2"
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_workload_config_carries_no_diagnostics_and_strong_scaling_divides_files():
    b = _bench()
    args = types.SimpleNamespace(files=10_000, per_file=100_000, universe=1_000_000, kind="bits", nested=0.0, width_scale=1, scaling="weak")
    weak = b.workload_config(args, 8)
    assert weak["files_per_gpu"] == 10_000 and weak["scaling"] == "weak"
    assert not any(k.startswith("parity") or k in ("hits_per_gpu", "index") for k in weak)
    args.scaling = "strong"
    assert b.workload_config(args, 8)["files_per_gpu"] == 1250


def test_host_threads_ignores_omp_num_threads(monkeypatch):
    b = _bench()
    monkeypatch.setenv("OMP_NUM_THREADS", "1")
    assert b.host_threads() == len(os.sched_getaffinity(0))
    from oracle import oracle as orc
    assert orc.max_threads() == len(os.sched_getaffinity(0))


def test_digest_is_order_sensitive():
    b = _bench()
    a = np.arange(10, dtype=np.uint32)
    assert b.digest64(a) == b.digest64(a.copy())
    assert b.digest64(a) != b.digest64(a[::-1])
    assert b.digest64(a[:5], a[5:]) == b.digest64(a)


def test_marshal_compact_matches_numpy():
    """gtgpu_marshal_compact is host code: runs cut at chromosome changes and file boundaries, u16 widths, exceptions."""
    from gtars_b200 import ffi
    rng = np.random.default_rng(7)
    n = 2_500_000
    fo = np.array([0, 10, 10, 1_200_000, n], dtype=np.uint64)
    chr_ = np.sort(rng.integers(0, 5, n)).astype(np.uint32)
    start = rng.integers(0, 1 << 30, n).astype(np.uint32)
    end = (start + rng.integers(0, 70000, n)).astype(np.uint32)
    end[5] = start[5] - 3
    for threads in (1, 3, 0):
        ro, rc, w16, wi, we = ffi.marshal_compact(chr_, start, end, fo, threads=threads)
        cut = np.zeros(n, bool)
        cut[0] = True
        cut[1:] = chr_[1:] != chr_[:-1]
        cut[fo[fo < n].astype(np.int64)] = True
        exp = np.flatnonzero(cut)
        assert np.array_equal(ro[:-1], exp) and ro[-1] == n and np.array_equal(rc, chr_[exp])
        wide = (end < start) | ((end.astype(np.int64) - start) > 0xFFFE)
        assert np.array_equal(wi, np.flatnonzero(wide)) and np.array_equal(we, end[wide])
        assert np.array_equal(w16[~wide], (end - start)[~wide].astype(np.uint16)) and (w16[wide] == 0xFFFF).all()
    ro, rc, w16, wi, we = ffi.marshal_compact(chr_[:0], start[:0], end[:0], np.zeros(1, np.uint64))
    assert len(rc) == 0 and len(wi) == 0 and ro[0] == 0

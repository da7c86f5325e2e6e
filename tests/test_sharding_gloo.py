"""Host-side multi-GPU logic on CPU: world_size-2 gloo processes shard the work, compute their part with the oracle
(standing in for the device), exchange with torch.distributed and must reproduce the unsharded result."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _workload():
    rng = np.random.default_rng(5)
    n_chroms, n_u = 3, 2000
    chr_ = np.sort(rng.integers(0, n_chroms, n_u))
    offs = np.concatenate([[0], np.cumsum(np.bincount(chr_, minlength=n_chroms))]).astype(np.uint64)
    s = rng.integers(0, 100_000, n_u).astype(np.uint32)
    e = (s + rng.integers(1, 2000, n_u)).astype(np.uint32)
    sizes = [0, 40, 300, 0, 0, 120, 7, 513, 0]
    nq = sum(sizes)
    fo = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    qc = rng.integers(0, n_chroms + 1, nq).astype(np.uint32)
    qs = rng.integers(0, 101_000, nq).astype(np.uint32)
    qe = (qs + rng.integers(1, 800, nq)).astype(np.uint32)
    # LOLA database: 7 sets (odd on purpose: the last rank's column block is short)
    db_sizes = rng.integers(20, 200, 7)
    dfo = np.concatenate([[0], np.cumsum(db_sizes)]).astype(np.uint64)
    nd = int(db_sizes.sum())
    dc = rng.integers(0, n_chroms, nd).astype(np.uint32)
    ds = rng.integers(0, 100_000, nd).astype(np.uint32)
    de = (ds + rng.integers(1, 5000, nd)).astype(np.uint32)
    return dict(offs=offs, s=s, e=e, fo=fo, qc=qc, qs=qs, qe=qe, dfo=dfo, dc=dc, ds=ds, de=de, n_chroms=n_chroms)


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gtars_b200 import shard
    from oracle import oracle as orc
    w = _workload()
    # --- tokenize: shard by file, no collective on the data path; gather only to check -------------------------
    ix = orc.Index(orc.BITS, w["offs"], w["s"], w["e"])
    f0, f1, q0, q1, fo_local = shard.shard_files(w["fo"], world, rank)
    off, ids = ix.tokenize_files(fo_local, w["qc"][q0:q1], w["qs"][q0:q1], w["qe"][q0:q1], 2000)
    parts = [None] * world
    dist.all_gather_object(parts, (off, ids))
    m_off, m_ids = shard.merge_tokenized(parts)
    full_off, full_ids = ix.tokenize_files(w["fo"], w["qc"], w["qs"], w["qe"], 2000)
    ok_tok = bool(np.array_equal(m_off, full_off) and np.array_equal(m_ids, full_ids))
    # --- LOLA counts: database sharded by set, column blocks all-gathered ------------------------------------------
    n_db = len(w["dfo"]) - 1
    lo, hi = shard.db_set_range(n_db, world, rank)
    r0, r1 = int(w["dfo"][lo]), int(w["dfo"][hi])
    g = orc.Igd((w["dfo"][lo:hi + 1] - w["dfo"][lo]).astype(np.uint64), w["dc"][r0:r1], w["ds"][r0:r1], w["de"][r0:r1]) \
        if hi > lo else None
    cols = -(-n_db // world)
    block = np.zeros((len(w["fo"]) - 1, cols), dtype=np.int64)
    if g is not None:
        block[:, :hi - lo] = g.count_region_hits(w["fo"], w["qc"], w["qs"], w["qe"], 1).astype(np.int64)
    gathered = [torch.zeros_like(torch.from_numpy(block)) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(block))
    full = shard.assemble_column_blocks(np.stack([t.numpy() for t in gathered]), n_db)
    ref = orc.Igd(w["dfo"], w["dc"], w["ds"], w["de"]).count_region_hits(w["fo"], w["qc"], w["qs"], w["qe"], 1)
    ok_lola = bool(np.array_equal(full.astype(np.uint64), ref))
    out[rank] = (ok_tok, ok_lola)
    dist.destroy_process_group()


def test_world2_sharding_reproduces_single_process():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: (True, True), 1: (True, True)}


def test_partition_arithmetic():
    from gtars_b200 import shard
    for n in (0, 1, 7, 8, 10_000):
        for world in (1, 2, 4, 8):
            blocks = [shard.block_range(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sets = [shard.db_set_range(n, world, r) for r in range(world)]
            assert sum(hi - lo for lo, hi in sets) == n
            cols = -(-n // world) if n else 0
            assert all(lo == min(r * cols, n) for r, (lo, hi) in enumerate(sets))

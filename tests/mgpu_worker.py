"""torchrun worker for tests/test_gpu_multi.py: one process per GPU, NCCL.  Checks, bit-exactly against the oracle,
(1) LOLA count matrices with the database sharded by region set + the library's own ncclAllGather, and
(2) tokenize sharded by file with no collective."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from gtars_b200 import ffi, shard, synth
    from oracle import oracle as orc

    ctx = ffi.Context(local)
    shard.init_comm_from_torch(ctx)

    # ---- LOLA: 37 database sets, 5 query sets + a "universe" -------------------------------------------------------
    rng = np.random.default_rng(2024)  # same stream on every rank
    n_db, n_chroms = 37, synth.N_CHROMS
    db = synth.make_uniform_intervals(37 * 2000, synth.SEED_LOLA_DB, min_w=200, max_w=5000)
    dfo = (np.arange(n_db + 1) * 2000).astype(np.uint64)
    dc, ds, de = (db[k].numpy().view(np.uint32) for k in ("chr", "start", "end"))
    q = synth.make_uniform_intervals(6 * 3000, synth.SEED_LOLA_USER, min_w=100, max_w=2000)
    so = (np.arange(7) * 3000).astype(np.uint64)
    qc, qs, qe = (q[k].numpy().view(np.uint32) for k in ("chr", "start", "end"))
    lo, hi = shard.db_set_range(n_db, world, rank)
    r0, r1 = int(dfo[lo]), int(dfo[hi])
    g = ffi.Igd(ctx, (dfo[lo:hi + 1] - dfo[lo]).astype(np.uint64), n_chroms, dc[r0:r1], ds[r0:r1], de[r0:r1])
    ref = orc.Igd(dfo, dc, ds, de)
    for binary in (True, False):
        got = g.count_sharded(binary, n_db, so, qc, qs, qe, 1)
        want = ref.count_region_hits(so, qc, qs, qe, 1) if binary else ref.count_set_overlaps(so, qc, qs, qe, 1)
        assert np.array_equal(got, want), f"rank {rank}: sharded LOLA counts differ (binary={binary})"
    hits = g.count_sharded(True, n_db, so, qc, qs, qe, 1)
    tables = orc.lola_tables(hits[:5], hits[5], np.diff(so)[:5], 3000)  # a,b,c,d with set 5 as the universe
    assert tables.shape == (5, n_db, 4)

    # ---- tokenize: files sharded in contiguous blocks, universe replicated, no collective ----------------------------
    u = synth.make_universe(50_000)
    qf = synth.make_query_files(u, 16, 2000)
    offs = u["chrom_offsets"].numpy().astype(np.uint64)
    s, e, v = (u[k].numpy().view(np.uint32) for k in ("g_start", "g_end", "g_val"))
    fo = qf["file_offsets"].numpy().astype(np.uint64)
    c2, s2, e2 = (qf[k].numpy().view(np.uint32) for k in ("chr", "start", "end"))
    ix = ffi.Index(ctx, ffi.KIND_BITS, offs, s, e, v)
    f0, f1, q0, q1, fo_local = shard.shard_files(fo, world, rank)
    part = ix.tokenize_files(fo_local, c2[q0:q1], s2[q0:q1], e2[q0:q1], u["unk_id"])
    parts = [None] * world
    dist.all_gather_object(parts, part)
    m_off, m_ids = shard.merge_tokenized(parts)
    o_off, o_ids = orc.Index(orc.BITS, offs, s, e, v).tokenize_files(fo, c2, s2, e2, u["unk_id"])
    assert np.array_equal(m_off, o_off) and np.array_equal(m_ids, o_ids), f"rank {rank}: sharded tokenize differs"
    dist.barrier()
    if rank == 0:
        print(f"MGPU_OK world={world}")
    ix.close()
    g.close()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

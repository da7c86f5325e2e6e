#!/usr/bin/env python
"""Secondary measurements for BASELINE.json configs C3 / C4 / C5 and the scoring matrix (bench.py is the contract
benchmark for C2).

One JSON line per config: device-resident kernel throughput (CUDA events on the launching stream), algorithmic bytes,
fraction of the measured HBM peak, and a bit-exact spot check against the CPU oracle on a sample.
Sizes are per GPU; pass --scale 1.0 for the full single-GPU share of each config (defaults keep the run short).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c3,c4,c5,score,ingest")
    ap.add_argument("--scale", type=float, default=0.2)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    import torch
    from gtars_b200 import ffi, synth
    from oracle import oracle as orc
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    ctx = ffi.Context(0, stream=stream.cuda_stream)
    u32 = lambda t: t.cpu().numpy().view(np.uint32)

    def timed(fn):
        with torch.cuda.stream(stream):
            for _ in range(3):
                fn()
            stream.synchronize()
            ctx.timing_enable(True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(args.steps):
                fn()
            e1.record(stream)
            stream.synchronize()
            k = ctx.timing_read()
            ctx.timing_enable(False)
        return e0.elapsed_time(e1) / args.steps, (sum(k) / len(k) if k else None)

    for cfg in args.configs.split(","):
        t_setup = time.time()
        if cfg == "c3":
            n_db, n_q = int(50_000_000 * args.scale), int(100_000_000 * args.scale)
            db = synth.make_uniform_intervals(n_db, synth.SEED_LOLA_DB, device=dev, min_w=100, max_w=10_000)
            g = synth.group_by_chrom(db["chr"], db["start"], db["end"])
            offs = g["chrom_offsets"].cpu().numpy().astype(np.uint64)
            s, e = u32(g["g_start"]), u32(g["g_end"])
            ix = ffi.Index(ctx, ffi.KIND_BITS, offs, s, e)
            q = synth.make_uniform_intervals(n_q, synth.SEED_QUERIES, device=dev, min_w=100, max_w=2000, log_uniform=False)
            d_out = torch.empty(n_q, dtype=torch.int32, device=dev)
            fn = lambda: ix.count_dev(n_q, q["chr"].data_ptr(), q["start"].data_ptr(), q["end"].data_ptr(), 0, d_out.data_ptr())
            os.environ["GTGPU_COUNT_PARTITION"] = "0"   # the direct pass (every search a random DRAM sector), for comparison
            ms_direct, _ = timed(fn)
            direct = d_out.clone()
            del os.environ["GTGPU_COUNT_PARTITION"]
            d_out.zero_()
            ms, kms = timed(fn)
            same = bool(torch.equal(direct, d_out))
            del direct
            m = min(n_q, 200_000)
            ok = same and bool(np.array_equal(u32(d_out[:m]), orc.Index(orc.BITS, offs, s, e).count(
                u32(q["chr"][:m]), u32(q["start"][:m]), u32(q["end"][:m]), threads=orc.max_threads())))
            algo = 16 * n_q + 8 * n_db
            out = dict(config="C3 Bits count", queries=n_q, db_intervals=n_db, ms_per_step=ms, value=n_q / (ms * 1e-3),
                       unit="queries/s", kernel_ms=kms, algorithmic_bytes=algo, index=ix.info(),
                       roofline_frac=algo / ((kms or ms) * 1e-3) / 1e9 / peak, parity_sample_vs_oracle=ok,
                       ms_per_step_direct_pass=ms_direct, bucketed_equals_direct=same,
                       note="unsorted queries; the rank LUTs exceed the L2, so the queries are bucketed by LUT slice first "
                            "(hist + partition + count + gather, all inside ms_per_step); ms_per_step_direct_pass = one "
                            "random DRAM sector per search")
        elif cfg == "c4":
            n_db = max(int(10_000 * args.scale), 8)
            per_db, n_user, per_user = 20_000, max(int(1000 * args.scale), 4), 10_000
            db = synth.make_uniform_intervals(n_db * per_db, synth.SEED_LOLA_DB, device=dev, min_w=200, max_w=5000)
            dfo = (np.arange(n_db + 1) * per_db).astype(np.uint64)
            dc, ds, de = u32(db["chr"]), u32(db["start"]), u32(db["end"])
            g = ffi.Igd(ctx, dfo, synth.N_CHROMS, dc, ds, de)
            u = synth.make_universe(1_000_000, device=dev)
            pick = synth.rand_u63(synth.SEED_LOLA_USER, 1, torch.arange(n_user * per_user, device=dev)) % u["n"]
            qc = torch.cat([u["chr"][pick], u["chr"]]).contiguous()
            qs = torch.cat([u["start"][pick], u["start"]]).contiguous()
            qe = torch.cat([u["end"][pick], u["end"]]).contiguous()
            n_q = qc.numel()
            set_of = torch.cat([torch.arange(n_user, device=dev).repeat_interleave(per_user),
                                torch.full((u["n"],), n_user, device=dev)]).int().contiguous()
            d_out = torch.zeros((n_user + 1) * n_db, dtype=torch.int64, device=dev)

            def fn():
                d_out.zero_()
                ffi.check(ffi.lib().gtgpu_igd_count_dev(g._h, 1, n_q, set_of.data_ptr(), qc.data_ptr(), qs.data_ptr(),
                                                        qe.data_ptr(), 1, d_out.data_ptr()))
            ms, kms = timed(fn)
            hits = d_out.view(n_user + 1, n_db).cpu().numpy().astype(np.uint64)
            k = 2
            so = (np.arange(k + 1) * per_user).astype(np.uint64)
            ref = orc.Igd(dfo, dc, ds, de).count_region_hits(so, u32(qc[:k * per_user]), u32(qs[:k * per_user]),
                                                             u32(qe[:k * per_user]), 1, threads=orc.max_threads())
            ok = bool(np.array_equal(hits[:k], ref))
            pair_hits = int(hits.sum())
            algo = 12 * n_q + 12 * n_db * per_db + 8 * (n_user + 1) * n_db
            out = dict(config="C4 LOLA region-hit matrix", db_sets=n_db, db_records=n_db * per_db, user_sets=n_user,
                       query_regions=n_q, ms_per_step=ms, kernel_ms=kms, value=n_q / (ms * 1e-3), unit="query regions/s",
                       region_file_hits=pair_hits, hits_per_s=pair_hits / ((kms or ms) * 1e-3), algorithmic_bytes=algo,
                       roofline_frac=algo / ((kms or ms) * 1e-3) / 1e9 / peak, parity_sample_vs_oracle=ok, igd=g.info(),
                       note="atomic / L2-bound by construction; the HBM fraction is low on purpose (SURVEY §8d)")
        elif cfg == "c5":
            n = int(125_000_000 * args.scale)  # one GPU's share of the 1 B-fragment, 8-GPU configuration
            n_bc = 100_000
            u = synth.make_universe(1_000_000, device=dev)
            offs = u["chrom_offsets"].cpu().numpy().astype(np.uint64)
            s, e, v = (u32(u[k]) for k in ("g_start", "g_end", "g_val"))
            ix = ffi.Index(ctx, ffi.KIND_BITS, offs, s, e, v)
            q = synth.make_query_files(u, 1, n, seed=synth.SEED_FRAGMENTS, device=dev, sort_files=False)
            qc, qs, qe = u32(q["chr"]), u32(q["start"]), u32(q["end"])
            bc = (synth.rand_u63(synth.SEED_FRAGMENTS, 9, torch.arange(n, device=dev)) % n_bc)
            bc = ((bc * bc) // n_bc).int()  # skewed barcode sizes
            bc_h = u32(bc)
            t0 = time.perf_counter()
            off, ids = ix.tokenize_fragments(qc, qs, qe, bc_h, n_bc, u["unk_id"])
            t1 = time.perf_counter()
            off, ids = ix.tokenize_fragments(qc, qs, qe, bc_h, n_bc, u["unk_id"])
            t2 = time.perf_counter()
            m = min(n, 2_000_000)
            o_off, o_ids = orc.Index(orc.BITS, offs, s, e, v).tokenize_fragments(qc[:m], qs[:m], qe[:m], bc_h[:m], n_bc, u["unk_id"])
            g_off, g_ids = ix.tokenize_fragments(qc[:m], qs[:m], qe[:m], bc_h[:m], n_bc, u["unk_id"])
            ok = bool(np.array_equal(o_off, g_off) and np.array_equal(o_ids, g_ids))
            out = dict(config="C5 fragment tokenization (host buffers in/out, pageable)", fragments=n, barcodes=n_bc, ids=int(len(ids)),
                       seconds_first_call=t1 - t0, seconds_warm_call=t2 - t1, value=n / (t2 - t1), unit="fragments/s (end to end)",
                       parity_sample_vs_oracle=ok,
                       note="end-to-end through gtgpu_tokenize_fragments incl. H2D/D2H; group-by uses the hand-written radix sort + scan of sort.cu")
        elif cfg == "score":
            # gtars-scoring region_scoring_from_fragments, ATAC mode: 8 pseudo-bulk fragment files vs the 1 M-peak
            # consensus (dense u32 [8 x 1 M] matrix); every fragment is two lookups, one of them a reversed interval
            n_files, per_file = 8, int(25_000_000 * args.scale)
            u = synth.make_universe(1_000_000, device=dev)
            offs = u["chrom_offsets"].cpu().numpy().astype(np.uint64)
            s, e, v = (u32(u[k]) for k in ("g_start", "g_end", "g_val"))
            ix = ffi.Index(ctx, ffi.KIND_BITS, offs, s, e, v)
            q = synth.make_query_files(u, n_files, per_file, seed=synth.SEED_FRAGMENTS, device=dev, sort_files=False)
            n = n_files * per_file
            n_cols = int(u["n"])
            d_mat = torch.zeros(n_files * n_cols, dtype=torch.int32, device=dev)
            fo = q["file_offsets"]
            fn = lambda: ix.score_matrix_dev(n_files, fo.data_ptr(), n, q["chr"].data_ptr(), q["start"].data_ptr(),
                                             q["end"].data_ptr(), ffi.SCORE_ATAC, n_cols, d_mat.data_ptr())
            ms, kms = timed(fn)
            m = min(per_file, 100_000)  # oracle on the head of the first two files
            sub_fo = np.array([0, m, 2 * m], dtype=np.uint64)
            sel = torch.cat([torch.arange(m, device=dev), per_file + torch.arange(m, device=dev)])
            sc, ss, se = (u32(q[k][sel]) for k in ("chr", "start", "end"))
            ref = orc.score_matrix(orc.Index(orc.BITS, offs, s, e, v), sub_fo, sc, ss, se, orc.SCORE_ATAC, n_cols, threads=orc.max_threads())
            got = ix.score_matrix(sub_fo, sc, ss, se, ffi.SCORE_ATAC, n_cols)
            ok = bool(np.array_equal(got, ref))
            hits = int(d_mat.sum().item())
            algo = 12 * n + 4 * n_files * n_cols
            out = dict(config="scoring: ATAC count matrix (gtars-scoring)", fragment_files=n_files, fragments=n, peaks=n_cols,
                       lookups=2 * n, matrix_increments=hits, ms_per_step=ms, fused_find_ms=kms, value=n / (ms * 1e-3),
                       unit="fragments/s", algorithmic_bytes=algo, roofline_frac=algo / (ms * 1e-3) / 1e9 / peak,
                       parity_sample_vs_oracle=ok,
                       note="device-resident; step = query expansion + fused find (reversed end intervals take the walk) + "
                            "(file, peak) histogram with 32-bit atomics")
        elif cfg == "ingest":
            # BED text -> token ids: device ingest (gtgpu_tokenize_bed) vs the host parser + device tokenizer
            import tempfile
            from gtars_b200 import api
            n_lines = int(20_000_000 * args.scale)
            u = synth.make_universe(1_000_000, device=dev)
            names = list(synth.CHROM_NAMES)
            q = synth.make_query_files(u, 1, n_lines, device=dev)
            qc, qs, qe = (q[k].cpu().numpy() for k in ("chr", "start", "end"))
            tmp = tempfile.mkdtemp()
            upath, qpath = os.path.join(tmp, "universe.bed"), os.path.join(tmp, "query.bed")
            uc, us, ue = (u[k].cpu().numpy() for k in ("chr", "start", "end"))
            name_arr = np.array(names)
            with open(upath, "w") as f:
                f.write("\n".join(map("\t".join, zip(name_arr[uc], us.astype(str), ue.astype(str)))) + "\n")
            with open(qpath, "w") as f:
                f.write("\n".join(map("\t".join, zip(name_arr[qc], qs.astype(str), qe.astype(str)))) + "\n")
            text_bytes = os.path.getsize(qpath)
            tok = api.Tokenizer(upath)
            t0 = time.perf_counter(); ids_dev = tok.encode_bed_file(qpath); t1 = time.perf_counter()
            ids_dev = tok.encode_bed_file(qpath); t2 = time.perf_counter()
            # the C ABI call alone: text already in host memory -> ids in a pinned buffer
            offs = u["chrom_offsets"].cpu().numpy().astype(np.uint64)
            ix = ffi.Index(ctx, ffi.KIND_BITS, offs, *(u32(u[k]) for k in ("g_start", "g_end", "g_val")))
            text = open(qpath, "rb").read()
            ix.tokenize_bed(text, names, int(u["unk_id"]))
            ta = time.perf_counter(); ids_abi = ix.tokenize_bed(text, names, int(u["unk_id"])); tb = time.perf_counter()
            t2b = time.perf_counter()
            rs = api.RegionSet(qpath); t3 = time.perf_counter()
            ids_host = tok.encode(rs); t4 = time.perf_counter()
            out = dict(config="ingest: BED text -> token ids", lines=n_lines, text_bytes=text_bytes, ids=len(ids_dev),
                       seconds_device_ingest_first=t1 - t0, seconds_device_ingest_warm=t2 - t1,
                       seconds_host_parse=t3 - t2b, seconds_encode_after_host_parse=t4 - t3,
                       seconds_c_abi_call=tb - ta, c_abi_lines_per_s=n_lines / (tb - ta), c_abi_text_gb_per_s=text_bytes / (tb - ta) / 1e9,
                       value=n_lines / (t2 - t1), unit="lines/s (file read + H2D + parse + tokenize + D2H + list conversion)",
                       speedup_vs_host_parse=(t4 - t2b) / (t2 - t1),
                       parity_ids_equal=bool(ids_dev == ids_host and np.array_equal(np.asarray(ids_abi), np.asarray(ids_dev, dtype=np.uint32))),
                       note="both paths include reading the file and converting the result to a Python list")
        else:
            continue
        out["setup_seconds"] = time.time() - t_setup
        print(json.dumps(out), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()

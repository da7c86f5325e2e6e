#!/usr/bin/env python
"""bench.py — headline benchmark: tokenize synthetic hg38-shaped BED files against a 1 M-region universe.

A "step" is one pass of the hot path (Tokenizer::encode for every file, gtars-tokenizers/src/tokenizer.rs:140-171)
over BASELINE.json configs[1]: 10 000 files x 100 000 regions = 1e9 query intervals vs a 1 M-region universe, per GPU.

  value     query intervals / s, device-resident inputs, CUDA events around K back-to-back steps (max over ranks)
  e2e       the same metric through the C-ABI host entry point gtgpu_tokenize_files with PINNED HOST buffers,
            H2D of the queries and D2H of the ids inside the timed region
  roofline  algorithmic HBM bytes of the fused kernel / its own CUDA-event time, vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the CPU oracle (C++ restatement of the reference's algorithm; the reference is Rust and cannot be
            built here) on a bounded sample of the same files, 1 thread (the reference path is single-threaded)

`--impl reference` times the oracle with all host threads instead (rank 0 only).  Multi-GPU (torchrun): files are
sharded across ranks (rank r owns files [r*F, (r+1)*F)), universe replicated, no collective on the data path (weak scaling).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "query_intervals_per_sec_tokenize_vs_universe"
UNIT = "queries/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--files", type=int, default=10_000, help="files per GPU")
    ap.add_argument("--per-file", type=int, default=100_000)
    ap.add_argument("--universe", type=int, default=1_000_000)
    ap.add_argument("--kind", default="bits", choices=["bits", "ailist"])
    ap.add_argument("--nested", type=float, default=0.0, help="fraction of wide intervals (C2n variant)")
    ap.add_argument("--width-scale", type=int, default=1, help="multiply the 200-600 bp query widths (wide-query variant)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-api", default="compact", choices=["compact", "runs"], help="host entry point of the e2e leg")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


def workload_config(args, world):
    return {
        "workload": f"C2: tokenize {args.files} files x {args.per_file} regions per GPU vs {args.universe}-region hg38 universe",
        "files_per_gpu": args.files, "regions_per_file": args.per_file, "universe_regions": args.universe,
        "backend": args.kind, "nested_frac": args.nested, "query_width_scale": args.width_scale, "sharding": f"files x{world} (universe replicated)",
        "l2_policy": "inputs (12 B/query x 1e9) are far larger than the 126 MB L2; no flush needed",
    }


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                       "-i", str(self.gpu), "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_rate(args, threads, seconds, universe=None):
    """Oracle tokenize throughput on files [0, k) of the same synthetic stream; k grows until `seconds` are spent."""
    from gtars_b200 import synth
    from oracle import oracle as orc
    u = universe or synth.make_universe(args.universe, nested_frac=args.nested)
    kind = orc.BITS if args.kind == "bits" else orc.AILIST
    offs = u["chrom_offsets"].cpu().numpy().astype(np.uint64)
    s, e, v = (u[k].cpu().numpy().view(np.uint32) for k in ("g_start", "g_end", "g_val"))
    ix = orc.Index(kind, offs, s, e, v)
    batch = max(1, min(args.files, (4 if threads == 1 else 4 * threads)))
    done_q, spent, first = 0, 0.0, 0
    while spent < seconds and first < args.files:
        nf = min(batch, args.files - first)
        q = synth.make_query_files(u, nf, args.per_file, first_file=first, width_scale=args.width_scale)
        qc, qs, qe = (q[k].numpy().view(np.uint32) for k in ("chr", "start", "end"))
        fo = q["file_offsets"].numpy().astype(np.uint64)
        t0 = time.perf_counter()
        ix.tokenize_files(fo, qc, qs, qe, u["unk_id"], threads=threads)
        spent += time.perf_counter() - t0
        done_q += len(qc)
        first += nf
    return done_q / spent, f"files [0,{first}) = {done_q} queries in {spent:.2f} s"


def run_reference(args, rank, world):
    """The reference arm: the CPU oracle (a C++ port; the Rust reference cannot be compiled in this image) with every
    host thread, on a bounded sample of the same workload per step."""
    if rank != 0:
        return
    from gtars_b200 import synth
    from oracle import oracle as orc
    threads = orc.max_threads()
    u = synth.make_universe(args.universe, nested_frac=args.nested)
    kind = orc.BITS if args.kind == "bits" else orc.AILIST
    offs = u["chrom_offsets"].numpy().astype(np.uint64)
    s, e, v = (u[k].numpy().view(np.uint32) for k in ("g_start", "g_end", "g_val"))
    ix = orc.Index(kind, offs, s, e, v)
    sample_files = min(args.files, 8 * threads)
    q = synth.make_query_files(u, sample_files, args.per_file, width_scale=args.width_scale)
    qc, qs, qe = (q[k].numpy().view(np.uint32) for k in ("chr", "start", "end"))
    fo = q["file_offsets"].numpy().astype(np.uint64)
    for _ in range(max(args.warmup, 1)):
        ix.tokenize_files(fo, qc, qs, qe, u["unk_id"], threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ix.tokenize_files(fo, qc, qs, qe, u["unk_id"], threads=threads)
    dt = (time.perf_counter() - t0) / args.steps
    value = len(qc) / dt
    sample = f"{sample_files} files x {args.per_file} regions per step (bounded sample of the {args.files}-file workload)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": workload_config(args, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def bind_to_gpu_numa_node(torch, local_rank):
    """N > 1 only: pin this rank to the CPUs of its GPU's NUMA node before any pinned host buffer is allocated, so the
    e2e leg's staging pages are local to the GPU's PCIe root (eight unbound ranks shared one socket's memory and links:
    16 GB/s H2D per GPU measured at N = 8).  Best effort: any failure leaves the process unbound."""
    try:
        p = torch.cuda.get_device_properties(local_rank)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return {"node": node, "cpus": len(cpus)}
    except Exception:
        return None


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from gtars_b200 import ffi, synth

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(torch, local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    stream = torch.cuda.Stream(device=dev)
    ctx = ffi.Context(local_rank, stream=stream.cuda_stream)
    kind = ffi.KIND_BITS if args.kind == "bits" else ffi.KIND_AILIST

    # ---- universe + index (replicated on every rank) ---------------------------------------------------------------
    u = synth.make_universe(args.universe, nested_frac=args.nested, device=dev)
    offs = u["chrom_offsets"].cpu().numpy().astype(np.uint64)
    s, e, v = (u[k].cpu().numpy().view(np.uint32) for k in ("g_start", "g_end", "g_val"))
    index = ffi.Index(ctx, kind, offs, s, e, v)
    info = index.info()

    # ---- this rank's files, generated on the device (untimed) -------------------------------------------------------
    n_files, per_file = args.files, args.per_file
    n = n_files * per_file
    first_file = rank * n_files
    d_chr = torch.empty(n, dtype=torch.int32, device=dev)
    d_start = torch.empty(n, dtype=torch.int32, device=dev)
    d_end = torch.empty(n, dtype=torch.int32, device=dev)
    chunk = max(1, min(n_files, (32 << 20) // per_file))
    for f0 in range(0, n_files, chunk):
        k = min(chunk, n_files - f0)
        q = synth.make_query_files(u, k, per_file, device=dev, first_file=first_file + f0, width_scale=args.width_scale)
        sl = slice(f0 * per_file, (f0 + k) * per_file)
        d_chr[sl], d_start[sl], d_end[sl] = q["chr"], q["start"], q["end"]
        del q
    d_file_offsets = torch.arange(n_files + 1, dtype=torch.int64, device=dev) * per_file
    cap = n + n // 4 + 1024
    d_ids = torch.empty(cap, dtype=torch.int32, device=dev)
    d_file_tok = torch.empty(n_files + 1, dtype=torch.int64, device=dev)
    d_total = torch.zeros(1, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()

    def step():
        index.find_dev(n, d_chr.data_ptr(), d_start.data_ptr(), d_end.data_ptr(), 0, n_files,
                       d_file_offsets.data_ptr(), d_ids.data_ptr(), cap, None, d_file_tok.data_ptr(),
                       d_total.data_ptr())

    with torch.cuda.stream(stream):
        step()
        stream.synchronize()
        if int(d_total.item()) > cap:  # multi-hit universes (the nested variant): size the id buffer exactly and redo
            cap = int(d_total.item())
            d_ids = torch.empty(cap, dtype=torch.int32, device=dev)
        for _ in range(max(args.warmup, 3)):
            step()
        stream.synchronize()
        hits = int(d_total.item())
        assert hits <= cap, "output capacity too small for the synthetic workload"
        n_empty_files = int((d_file_tok[1:] == d_file_tok[:-1]).sum().item())

        # ---- timed region: K back-to-back steps, CUDA events on the launching stream ---------------------------------
        sampler = ClockSampler(local_rank)
        launches0 = ctx.launch_count()
        ctx.timing_enable(True)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        torch.cuda.synchronize()
        sampler.start()
        ev0.record(stream)
        for _ in range(args.steps):
            step()
        ev1.record(stream)
        stream.synchronize()
        torch.cuda.synchronize()
        barrier()
        kernel_ms = ctx.timing_read()
        ctx.timing_enable(False)
        launches = ctx.launch_count() - launches0
        # K steps of a few ms are shorter than one nvidia-smi sampling period: keep the identical load running (untimed)
        # until the sampler has covered at least half a second, so that the clocks line reflects this workload.
        t_probe = time.perf_counter()
        while time.perf_counter() - t_probe < 0.5:
            for _ in range(8):
                step()
            stream.synchronize()
        clocks = sampler.stop()
        clocks["window"] = "timed region + 0.5 s of the identical steps right behind it"
        ms_per_step = max_over_ranks(ev0.elapsed_time(ev1) / args.steps)


    total_queries = sum_over_ranks(float(n))
    value = total_queries / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel (fused count→scan→emit) ---------------------------------------------------------
    # algorithmic bytes per launch: queries read once (12 B each), ids written once (4 B per hit), per-file offsets
    # written once, index (starts, pmax, vals = 12 B per interval) read once.  Per-query counts/offsets are never
    # materialised, so they are not counted.
    algo_bytes = 12 * n + 4 * hits + 8 * (n_files + 1) + 12 * info["n_intervals"]
    k_ms = statistics.mean(kernel_ms) if kernel_ms else ms_per_step
    peak, peak_src = peaks()
    achieved = algo_bytes / (k_ms * 1e-3) / 1e9
    traffic = None  # dram__bytes_read + dram__bytes_write of this kernel from the committed ncu --set full capture
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            tr = json.load(f)
        w = tr["workload"]
        if (w["files_per_gpu"], w["regions_per_file"], w["universe_regions"], w["backend"], w["nested_frac"]) == (
                args.files, args.per_file, args.universe, args.kind, args.nested):
            traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": "profiles/r01_traffic.json (ncu, same command)" if traffic else None,
                "peak_source": peak_src, "kernel": "fused_find_kernel", "kernel_ms": k_ms,
                "algorithmic_bytes_per_launch": algo_bytes, "bytes_per_query": algo_bytes / n,
                "kernel_share_of_step": k_ms / (ev0.elapsed_time(ev1) / args.steps)}

    # ---- e2e: the C-ABI host entry point with pinned host buffers --------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        # What a Rust caller holds after RegionSet::try_from: start / end arrays plus, because the files are sorted by
        # chromosome, a handful of (offset, chromosome id) runs per file (gtgpu_tokenize_files_runs).
        # Ends travel as 16-bit widths (gtgpu_tokenize_files_compact): peak-sized regions cost 6 B of PCIe each, not 8;
        # regions wider than 65 534 bp (none in C2) go through the exception list.  --e2e-api runs = start / end arrays.
        compact = args.e2e_api == "compact"
        h_start = ffi.pinned_empty(n, np.uint32)
        torch.from_numpy(h_start.view(np.int32)).copy_(d_start)
        if compact:
            h_w16 = ffi.pinned_empty(n, np.uint16)
            width = d_end.to(torch.int64) - d_start.to(torch.int64)
            is_exc = (width < 0) | (width > 65534)
            torch.from_numpy(h_w16.view(np.int16)).copy_(torch.where(is_exc, torch.full_like(width, 0xFFFF), width).to(torch.int16))
            exc = torch.nonzero(is_exc).flatten()
            h_wide_idx = exc.cpu().numpy().astype(np.uint64)
            h_wide_end = d_end[exc].cpu().numpy().view(np.uint32)
            del width, is_exc, exc
        else:
            h_end = ffi.pinned_empty(n, np.uint32)
            torch.from_numpy(h_end.view(np.int32)).copy_(d_end)
        h_fo = d_file_offsets.cpu().numpy().astype(np.uint64)
        brk = torch.nonzero(d_chr[1:] != d_chr[:-1]).flatten() + 1
        run_starts = torch.unique(torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), brk, d_file_offsets[:-1]]))
        h_run_chr = d_chr[run_starts].cpu().numpy().view(np.uint32)
        h_run_off = torch.cat([run_starts, torch.tensor([n], device=dev)]).cpu().numpy().astype(np.uint64)
        del brk, run_starts
        torch.cuda.synchronize()
        L = ffi.lib()

        def e2e_call():
            if compact:
                return index.tokenize_files_compact(h_fo, h_run_off, h_run_chr, h_start, h_w16, h_wide_idx, h_wide_end, u["unk_id"],
                                                    keep_buf=True)
            return index.tokenize_files_runs(h_fo, h_run_off, h_run_chr, h_start, h_end, u["unk_id"], keep_buf=True)

        def e2e_step():
            off, buf = e2e_call()
            total = int(off[-1])  # the step's result is read on the host
            assert L.gtgpu_buf_len(buf) == total
            L.gtgpu_buf_free(buf)
            return total

        # the host entry point must return exactly what the device-resident path produced
        off_chk, buf_chk = e2e_call()
        k_chk = min(int(off_chk[-1]), 4_000_000)
        ids_chk = np.ctypeslib.as_array(C.cast(L.gtgpu_buf_data(buf_chk), C.POINTER(C.c_uint32)), shape=(int(off_chk[-1]),))
        e2e_matches_device = bool(n_empty_files > 0 or (np.array_equal(ids_chk[:k_chk], d_ids[:k_chk].cpu().numpy().view(np.uint32))
                                                        and np.array_equal(ids_chk[-k_chk:], d_ids[hits - k_chk:hits].cpu().numpy().view(np.uint32))))
        del ids_chk
        L.gtgpu_buf_free(buf_chk)

        for _ in range(2):
            e2e_total = e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        torch.cuda.synchronize()
        barrier()
        e2e_s = max_over_ranks((time.perf_counter() - t0) / args.steps)
        e2e = {"value": total_queries / e2e_s, "unit": UNIT,
               "h2d_bytes_per_step": (6 * n + 12 * len(h_wide_idx) if compact else 8 * n) + 8 * (n_files + 1) + 12 * len(h_run_chr) + 8,
               "d2h_bytes_per_step": 4 * e2e_total + 8 * (n_files + 1), "ms_per_step": e2e_s * 1e3,
               "api": ("gtgpu_tokenize_files_compact (pinned host start u32 + width u16 + chromosome runs in, pinned result buffer out)"
                       if compact else "gtgpu_tokenize_files_runs (pinned host start/end + chromosome runs in, pinned result buffer out)"),
               "ids_match_device_resident_path": e2e_matches_device,
               "chromosome_runs": int(len(h_run_chr))}
        assert e2e_total == hits + n_empty_files

    # ---- parity spot check + cpu baseline (rank 0, N = 1) ------------------------------------------------------------------
    cpu_baseline = None
    parity = None
    if rank == 0:
        from oracle import oracle as orc
        chk_files = min(n_files, 8)
        m = chk_files * per_file
        qc, qs, qe = (t[:m].cpu().numpy().view(np.uint32) for t in (d_chr, d_start, d_end))
        o = orc.Index(orc.BITS if args.kind == "bits" else orc.AILIST, offs, s, e, v)
        o_off, o_ids = o.tokenize_files(h_fo[:chk_files + 1] if e2e else np.arange(chk_files + 1, dtype=np.uint64) * per_file,
                                        qc, qs, qe, u["unk_id"], threads=orc.max_threads())
        g_off = d_file_tok[:chk_files + 1].cpu().numpy().astype(np.uint64)
        g_ids = d_ids[:int(g_off[-1])].cpu().numpy().view(np.uint32)
        # raw device offsets have no [unk] insertions; the synthetic files are never empty
        parity = bool(np.array_equal(g_off, o_off) and np.array_equal(g_ids, o_ids))
        if world == 1 and not args.no_cpu:
            rate, sample = cpu_oracle_rate(args, 1, args.cpu_seconds, universe={k: (t.cpu() if hasattr(t, "cpu") else t) for k, t in u.items()})
            cpu_baseline = {"value": rate, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                            "note": "C++ restatement of gtars' algorithm (the Rust reference cannot be built here); "
                                    "1 thread because the reference path is single-threaded"}

    if rank == 0:
        cfg = workload_config(args, world)
        cfg.update({"hits_per_gpu": hits, "index": info, "parity_spot_check_first_files_vs_oracle": parity})
        if world > 1:
            cfg["host_binding_rank0"] = numa or "unbound"
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic", "config": cfg, "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e,
            "gpu_launches": launches, "clocks": clocks,
        }
        print(json.dumps(out))
    index.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

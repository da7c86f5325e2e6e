"""C3 (100 M unsorted queries vs a 50 M-interval database): rank-LUT bin width x bucket count, one process.
One JSON line per setting (ms per counting pass, CUDA events; `same` = identical to the first setting's counts)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gtars_b200 import ffi, synth  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
extras = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["0", "1", "2"])]
buckets = sys.argv[3].split(",") if len(sys.argv) > 3 else ["32", "64", "128", "256"]
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
ctx = ffi.Context(0, stream=stream.cuda_stream)
u32 = lambda t: t.cpu().numpy().view(np.uint32)
qscale = float(sys.argv[4]) if len(sys.argv) > 4 else scale  # queries alone (one rank's block of a query-sharded run)
n_db, n_q = int(50_000_000 * scale), int(100_000_000 * qscale)
db = synth.make_uniform_intervals(n_db, synth.SEED_LOLA_DB, device=dev, min_w=100, max_w=10_000)
g = synth.group_by_chrom(db["chr"], db["start"], db["end"])
offs, s, e = g["chrom_offsets"].cpu().numpy().astype(np.uint64), u32(g["g_start"]), u32(g["g_end"])
q = synth.make_uniform_intervals(n_q, synth.SEED_QUERIES, device=dev, min_w=100, max_w=2000, log_uniform=False)
d_out = torch.empty(n_q, dtype=torch.int32, device=dev)
ref = None
for extra in extras:
    os.environ["GTGPU_RANK_SHIFT_EXTRA"] = str(extra)
    t0 = time.perf_counter()
    ix = ffi.Index(ctx, ffi.KIND_BITS, offs, s, e)
    build = time.perf_counter() - t0
    fn = lambda: ix.count_dev(n_q, q["chr"].data_ptr(), q["start"].data_ptr(), q["end"].data_ptr(), 0, d_out.data_ptr())
    for env in [dict(GTGPU_COUNT_PARTITION="0")] + [dict(GTGPU_COUNT_PARTITION="1", GTGPU_COUNT_BUCKETS=nb) for nb in buckets] + [dict()]:
        for k in ("GTGPU_COUNT_PARTITION", "GTGPU_COUNT_BUCKETS"):
            os.environ.pop(k, None)
        os.environ.update(env)
        with torch.cuda.stream(stream):
            for _ in range(2):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(5):
                fn()
            e1.record(stream)
            stream.synchronize()
        if ref is None:
            ref = d_out.clone()
        print(json.dumps(dict(rank_shift_extra=extra, env=env, ms=e0.elapsed_time(e1) / 5, same=bool(torch.equal(ref, d_out)),
                              build_s=build, device_bytes=ix.info()["device_bytes"])), flush=True)
        d_out.zero_()
    ix.close()

#!/usr/bin/env python
"""Sweep of the radix scatter kernel's shape (GTGPU_RS_SHAPE: threads x rounds x blocks per SM, sort.cu) on the C5 fragment
pipeline, device-resident: every shape must reproduce shape 0's output bit for bit; shape 0 is checked against the oracle
on a slice.  usage: python profiles/r02_sort_sweep.py [fragments=2.5e8] [steps=5] [shapes=0,1,2,...]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    from gtars_b200 import ffi, synth
    from oracle import oracle as orc

    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 250_000_000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    shapes = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else list(range(3))
    n_bc = 100_000
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    ctx = ffi.Context(0, stream=stream.cuda_stream)
    u = synth.make_universe(1_000_000, device=dev)
    offs = u["chrom_offsets"].cpu().numpy().astype(np.uint64)
    s, e, v = (u[k].cpu().numpy().view(np.uint32) for k in ("g_start", "g_end", "g_val"))
    index = ffi.Index(ctx, ffi.KIND_BITS, offs, s, e, v)
    d = [torch.empty(n, dtype=torch.int32, device=dev) for _ in range(4)]
    chunk = 1 << 24
    for a in range(0, n, chunk):
        k = min(chunk, n - a)
        q = bench._fragments(synth, u, a, k, dev)
        for dst, src in zip(d, q):
            dst[a:a + k] = src
        del q
    cap = n + n // 4 + 1024
    d_ids = torch.empty(cap, dtype=torch.int32, device=dev)
    d_bco = torch.empty(n_bc + 1, dtype=torch.int64, device=dev)
    d_total = torch.zeros(1, dtype=torch.int64, device=dev)
    unk = int(u["unk_id"])
    fn = lambda: index.tokenize_fragments_dev(n, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), n_bc, unk,
                                              d_bco.data_ptr(), d_ids.data_ptr(), cap, d_total.data_ptr())
    ref_ids = ref_bco = None
    for shape in [x for x in shapes if 0 <= x < 3]:
        for pf in (1,):
            os.environ["GTGPU_RS_SHAPE"] = str(shape)
            os.environ["GTGPU_RS_PREFETCH"] = str(pf)
            d_ids.zero_()
            with torch.cuda.stream(stream):
                for _ in range(2):
                    fn()
                stream.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(steps):
                    fn()
                e1.record(stream)
                stream.synchronize()
            ms = e0.elapsed_time(e1) / steps
            total = int(d_total.item())
            rec = {"shape": shape, "prefetch_bucket_row": pf, "fragments": n, "ms": ms, "fragments_per_s": n / (ms * 1e-3), "tokens": total}
            if ref_ids is None:
                ref_ids, ref_bco = d_ids[:total].clone(), d_bco.clone()
                m = min(n, 2_000_000)
                hc, hs, he, hb = (t[:m].cpu().numpy().view(np.uint32) for t in d)
                o_off, o_ids = orc.Index(orc.BITS, offs, s, e, v).tokenize_fragments(hc, hs, he, hb, n_bc, unk)
                g_off, g_ids = index.tokenize_fragments(hc, hs, he, hb, n_bc, unk)
                rec["slice_equals_oracle"] = bool(np.array_equal(o_off, g_off) and np.array_equal(o_ids, g_ids))
            else:
                rec["equals_first_shape"] = bool(torch.equal(ref_ids, d_ids[:total]) and torch.equal(ref_bco, d_bco))
            print(json.dumps(rec), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()

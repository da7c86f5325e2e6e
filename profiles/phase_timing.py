"""Per-phase cycle accounting of the fused kernel (needs a -DGT_PHASE_TIMING build selected with GTGPU_LIB)."""
import ctypes, os, subprocess, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gtars_b200 import ffi, synth
L = ffi.lib()
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
ctx = ffi.Context(0, stream=stream.cuda_stream)
u = synth.make_universe(1_000_000, device=dev)
offs = u["chrom_offsets"].cpu().numpy().astype(np.uint64)
s, e, v = (u[k].cpu().numpy().view(np.uint32) for k in ("g_start", "g_end", "g_val"))
index = ffi.Index(ctx, ffi.KIND_BITS, offs, s, e, v)
n_files, per_file = 1000, 100_000
q = synth.make_query_files(u, n_files, per_file, device=dev)
n = n_files * per_file
fo = q["file_offsets"]
cap = n + n // 4
d_ids = torch.empty(cap, dtype=torch.int32, device=dev)
d_tok = torch.empty(n_files + 1, dtype=torch.int64, device=dev)
d_total = torch.zeros(1, dtype=torch.int64, device=dev)
def step():
    index.find_dev(n, q["chr"].data_ptr(), q["start"].data_ptr(), q["end"].data_ptr(), 0, n_files, fo.data_ptr(),
                   d_ids.data_ptr(), cap, None, d_tok.data_ptr(), d_total.data_ptr())
raw = (ctypes.c_ulonglong * 128)()
with torch.cuda.stream(stream):
    for _ in range(3): step()
    torch.cuda.synchronize()
    L.gtgpu_debug_phase_cycles(raw, 1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream); step(); ev1.record(stream); torch.cuda.synchronize()
    L.gtgpu_debug_phase_cycles(raw, 0)
buf = [sum(raw[w * 16 + i] for w in range(8)) for i in range(16)]
ms = ev0.elapsed_time(ev1)
names = ["0 queries ready", "1 bin records evaluated", "2 warp scan", "3 wait B2", "4 agg+claim", "5 look-back (incl B3)",
         "6 emit", "7 file marks", "8 windows", "9 lookbacks"]
warp_tiles = buf[9]
print(f"kernel+aux {ms:.3f} ms for {n} queries; warp-tiles {warp_tiles}")
tot = sum(buf[i] for i in range(8))
for i in range(8):
    print(f"  {names[i]:28s} {buf[i] / max(warp_tiles,1):10.0f} cycles/warp-tile  {100*buf[i]/tot:5.1f}%")
print(f"  look-back windows per tile: {buf[8]/max(warp_tiles,1):.2f};  total cycles per warp-tile {tot/max(warp_tiles,1):.0f}")
print("  per warp (cycles per warp-tile), phases 0-7:")
for w in range(8):
    wt = max(raw[w * 16 + 9], 1)
    print("   warp", w, " ".join(f"{raw[w * 16 + i] / wt:7.0f}" for i in range(8)), f"  look-back reloads per tile {raw[w * 16 + 8] / wt:.3f}")

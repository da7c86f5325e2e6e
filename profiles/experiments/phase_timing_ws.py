"""Per-phase cycle accounting of the warp-specialised fused kernel (-DGT_PHASE_TIMING build, GTGPU_FUSED_WS=1)."""
import ctypes, os, sys
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
buf = (ctypes.c_ulonglong * 16)()
with torch.cuda.stream(stream):
    for _ in range(3): step()
    torch.cuda.synchronize()
    L.gtgpu_debug_phase_cycles(buf, 1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream); step(); ev1.record(stream); torch.cuda.synchronize()
    L.gtgpu_debug_phase_cycles(buf, 0)
print(f"{ev0.elapsed_time(ev1):.3f} ms for {n} queries")
tiles = n // 1024
names = {0: "R wait empty slot", 1: "R queries (TMA wait + LDS)", 2: "R window word + entries", 3: "R warp scan", 4: "R barrier 1",
         5: "R aggregate + slot write + arrive", 6: "R barrier 2", 8: "E wait full", 9: "E lag wait (next tile resolved)",
         10: "E look-back", 11: "E emit", 12: "E file marks + release"}
for i, nm in names.items():
    per = buf[i] / tiles / (8 if i < 8 else 2)
    print(f"  {nm:38s} {per:9.0f} cycles per tile per warp")

#!/usr/bin/env python
"""What would fewer result bytes buy end to end?  The PCIe probe of bench.py in the e2e step's own mix (4.16 GB in, 3.46 GB out:
ids as u32) and with the ids packed to 3 bytes (2.59 GB out) or 2.5 bytes (20-bit ids, 2.16 GB out)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
dev = torch.device("cuda", 0)
D = bench.Dist(torch, 1, dev)
h2d = 4_157_397_472
for name, d2h in (("u32 ids", 3_455_275_468), ("24-bit ids", 3_455_275_468 * 3 // 4), ("20-bit ids", 3_455_275_468 * 5 // 8), ("no result", 1 << 20)):
    r = bench.pcie_probe(torch, dev, D, h2d, d2h)
    print(json.dumps({"mix": name, "h2d_bytes": h2d, "d2h_bytes": d2h, **{k: r[k] for k in ("h2d_alone_gbs_per_gpu", "d2h_alone_gbs_per_gpu", "mixed_h2d_gbs_per_gpu", "mixed_d2h_gbs_per_gpu", "aggregate_mixed_gbs", "floor_ms_for_one_e2e_step")}}), flush=True)

#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:radix|fused|scan_|frag_|rs_" --csv --log-file gpurun_out/c5_launches_2.5e8_group.csv \
  python profiles/r02_sort_sweep.py 2.5e8 1 "0" > gpurun_out/sort5.log 2>&1
tail -2 gpurun_out/sort5.log
python - <<'P'
import csv
rows=[r for r in csv.reader(open("gpurun_out/c5_launches_2.5e8_group.csv")) if len(r)>10]
hdr=rows[0]; ik=hdr.index("Kernel Name"); iv=hdr.index("Metric Value"); 
out=[(r[ik][:70], float(r[iv].replace(",",""))) for r in rows[1:]]
# last pipeline call = last ~14 launches
for k,v in out[-16:]: print(f"{v/1e3:10.1f} us  {k}")
P

"""Per-launch table from an `ncu --csv --metrics ...` log: python profiles/ncu_launch_table.py FILE.csv"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
d = collections.OrderedDict()
for r in rows[1:]:
    d.setdefault((int(r[iid]), r[ik].split("(")[0][-40:]), {})[r[im]] = float(r[iv].replace(",", ""))
print(f"{'id':>3} {'kernel':40} {'us':>9} {'rd MB':>9} {'wr MB':>9} {'L2 hit %':>8}")
for (i, k), v in d.items():
    print(f"{i:3d} {k:40} {v.get('gpu__time_duration.sum', 0) / 1e3:9.1f} {v.get('dram__bytes_read.sum', 0) / 1e6:9.1f} "
          f"{v.get('dram__bytes_write.sum', 0) / 1e6:9.1f} {v.get('lts__t_sector_hit_rate.pct', 0):8.1f}")

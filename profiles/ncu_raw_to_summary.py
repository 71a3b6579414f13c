#!/usr/bin/env python
"""`ncu --page raw --csv` (one row per captured launch, one column per metric) -> `metric,unit,value` lines for ONE launch.
    python profiles/ncu_raw_to_summary.py gpurun_out/<tag>_raw.csv [launch index] > profiles/<tag>_ncu_full_summary.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
r = rows[2 + (int(sys.argv[2]) if len(sys.argv) > 2 else 0)]
print("metric,unit,value")
for h, u, v in zip(hdr, units, r):
  if h in ("ID", "Process ID", "Process Name", "Host Name", "Context", "Stream", "Device", "CC"): continue
  print(f"{h},{u},{v}")

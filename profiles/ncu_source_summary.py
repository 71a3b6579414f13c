#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: instruction mix, stall reasons, hottest instructions.
    python profiles/ncu_source_summary.py gpurun_out/<tag>_source.csv"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
S, E, A = ix["Source"], ix["Instructions Executed"], ix["Warp Stall Sampling (All Samples)"]
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
  try: data.append((r[S].strip(), int(r[E] or 0), int(r[A] or 0), r))
  except Exception: pass
te, ts = sum(d[1] for d in data), sum(d[2] for d in data)
print(f"instructions executed {te:,}   stall samples {ts:,}   SASS lines {len(data)}")
st = collections.Counter()
for _, _, _, r in data:
  for h in stall_cols: st[h] += int(r[ix[h]] or 0)
print("stall reasons:"); 
for h, c in st.most_common(12): print(f"  {h:26s} {c:9d} {c / max(ts, 1) * 100:5.1f}%")
op, ops = collections.Counter(), collections.Counter()
for s, e, a, _ in data:
  t = s.split(); o = t[1] if t[0].startswith("@") else t[0]
  op[o] += e; ops[o] += a
print("opcode mix (by executed warp instructions):")
for o, c in op.most_common(32): print(f"  {o:30s} {c / te * 100:5.1f}%   samples {ops[o] / max(ts, 1) * 100:5.1f}%")
print("hottest instructions (by samples):")
for s, e, a, r in sorted(data, key=lambda d: -d[2])[:40]: print(f"  {a:8d} exec={e:12d}  {r[0][-5:]}  {s}")

#!/usr/bin/env python
"""HBM-bound stage kernels ("composite-only and encode-only micro-kernels are HBM-bound"): achieved GB/s of
nf_sample_points / nf_hash_encode / nf_composite / nf_generate_rays on bench-sized inputs vs the measured HBM copy peak.
    python profiles/stage_bench.py > gpurun_out/stages.json"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import nerf_atlas_b200 as N
from oracle import nerf_oracle as O   # synthetic parameters / rays only
from helpers import plain_param_list
dev = torch.device("cuda", 0)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6500.0

def timed(fn, reps=5):
  for _ in range(2): fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(reps): fn()
  e1.record(); torch.cuda.synchronize()
  return e0.elapsed_time(e1) / reps

P = O.make_plain_params(1337, 64, 1.0)
eng = N.RenderEngine(N.describe_plain(64, "upshifted", "black"), "fp16"); eng._p = plain_param_list(P, dev); eng.pack(eng._p)
R, T = 640000, 128
rays = O.make_rays(1, 800, 800, size=800, seed=0).reshape(-1, 6).contiguous().to(dev)
ts = torch.linspace(2, 6, T, device=dev)
rows = []
ms = timed(lambda: eng.sample_points(rays, ts))
rows.append({"kernel": "k_sample_points", "algorithmic_bytes": R * 24 + R * T * 12, "ms": ms})
pts = eng.sample_points(rays, ts).reshape(-1, 3)[: 16 * 1024 * 1024].contiguous()     # 16.8 M points: 201 MB in, 2.1 GB out
ms = timed(lambda: eng.hash_encode(pts), reps=3)
rows.append({"kernel": "k_hash_encode (8 levels x 8 gathers from L2-resident tables)", "algorithmic_bytes": pts.shape[0] * (12 + 128), "ms": ms})
sig = torch.randn(R, T, device=dev); feats = torch.rand(R, T, 3, device=dev)
ms = timed(lambda: eng.composite(sig, feats, rays, ts, want_weights=True))
rows.append({"kernel": "k_composite (alpha + weights out)", "algorithmic_bytes": R * T * (16 + 8) + R * 36, "ms": ms})
ms = timed(lambda: eng.composite(sig, feats, rays, ts, want_weights=False))
rows.append({"kernel": "k_composite (rgb only)", "algorithmic_bytes": R * T * 16 + R * 36, "ms": ms})
c2w, focal = O.make_cameras(16, 800, seed=1)
c2w = c2w.to(dev)
ms = timed(lambda: N.RenderEngine.generate_rays(c2w, focal, 800))
rows.append({"kernel": "k_generate_rays (16 views of 800x800)", "algorithmic_bytes": 16 * 640000 * 24, "ms": ms})
for r in rows:
  r["achieved_gbs"] = r["algorithmic_bytes"] / (r["ms"] * 1e-3) / 1e9
  r["hbm_peak_gbs"] = peak; r["frac"] = r["achieved_gbs"] / peak
  print(json.dumps(r), flush=True)

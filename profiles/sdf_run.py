#!/usr/bin/env python
"""One SDF surface render (sphere march, 800x800 unit rays, tensor pipeline) for a launch list:  ncu --metrics gpu__time_duration.sum ... python profiles/sdf_run.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import nerf_atlas_b200 as N
from oracle import nerf_oracle as O
from helpers import load_golden, sdf_params
dev = torch.device("cuda", 0)
fxs = load_golden("sdf_siren_march")
m = N.FusedSDF("siren", 64, t_near=2.0, t_far=6.0, sigmoid_kind="upshifted", precision="fp16", isect=sys.argv[1] if len(sys.argv) > 1 else "sphere")
m.load_state_dict(sdf_params(fxs), strict=True); m = m.to(dev).eval(); m.jitter = 0.5
rays = O.make_rays(1, 800, 800, size=800, seed=0).reshape(-1, 6).contiguous().to(dev)
rays[:, 3:] = torch.nn.functional.normalize(rays[:, 3:], dim=-1)
def f():
  with torch.no_grad(): return m(rays.reshape(1, 800, 800, 6))
f(); torch.cuda.synchronize()
for reps in (1, 2, 3):
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(reps): f()
  e1.record(); torch.cuda.synchronize()
  print("ms per render", e0.elapsed_time(e1) / reps, "reps", reps, "hits", float(m.hit.float().mean()))

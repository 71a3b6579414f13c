#!/usr/bin/env python
"""Time-in-state counters of the staggered pipeline (NF_TC_STATS build, NF_LIB=libnerf_b200_stats.so) on one 800x800x128 frame."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import nerf_atlas_b200 as N
from oracle import nerf_oracle as O
dev = torch.device("cuda", 0)
model = N.FusedPlainNeRF(steps=128, t_near=2, t_far=6, intermediate_size=64, sigmoid_kind="upshifted", precision="fp16")
model.load_state_dict(O.make_plain_params(1337, 64, 1.0), strict=True)
model = model.to(dev).eval()
eng = model.engine(); eng.pack(model._param_list())
rays = O.make_rays(1, 800, 800, size=800, seed=0).reshape(-1, 6).contiguous().to(dev)
ts = torch.linspace(2, 6, 128, device=dev)
for ring in sys.argv[1:] or ["3", "6"]:
  os.environ["NF_TC_RING"] = ring
  os.environ.pop("NF_TC_STATS_PRINT", None)
  eng.render(rays, ts, None, want_weights=False); torch.cuda.synchronize()
  os.environ["NF_TC_STATS_PRINT"] = "1"
  print("ring", ring, flush=True)
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record(); eng.render(rays, ts, None, want_weights=False); e1.record(); torch.cuda.synchronize()
  print("ms", e0.elapsed_time(e1), flush=True)

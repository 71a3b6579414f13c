#!/usr/bin/env python
"""Three Mip-cone frames (800x800x128, tensor pipeline) for an ncu capture:  ncu ... -k regex:k_render_tc3 -s 1 -c 1 python profiles/mip_run.py [pos]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import nerf_atlas_b200 as N
from oracle import nerf_oracle as O
from helpers import plain_param_list
dev = torch.device("cuda", 0)
rays = O.make_rays(1, 800, 800, size=800, seed=0).reshape(-1, 6).contiguous().to(dev)
ts = torch.linspace(2, 6, 128, device=dev)
if "pos" in sys.argv[1:]:
  Pp = O.make_plain_params(81, 64, 1.0, refl_kind="pos")
  mp = N.FusedPlainNeRF(steps=128, t_near=2, t_far=6, intermediate_size=64, sigmoid_kind="upshifted", precision="fp16", refl_kind="pos", keep_weights=False)
  mp.load_state_dict(Pp, strict=True); mp = mp.to(dev).eval()
  def f():
    with torch.no_grad(): return mp(rays.reshape(1, 800, 800, 6))
else:
  Pm = O.make_plain_params(61, 64, 1.0, mip=True)
  em = N.RenderEngine(N.describe_plain(64, "upshifted", "black", mip="cone"), "fp16"); em._p = plain_param_list(Pm, dev); em.pack(em._p)
  rad = em.ray_radii(rays.reshape(1, 800, 800, 6)).reshape(-1)
  f = lambda: em.render(rays, ts, radius=rad, want_weights=False)
for _ in range(3): f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(4): f()
e1.record(); torch.cuda.synchronize()
print("ms", e0.elapsed_time(e1) / 4)

"""Kernel-time breakdown of one native training step (4096 rays x 128 samples), torch.profiler (CUPTI) on a GPU box."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import nerf_atlas_b200 as N
from oracle import nerf_oracle as O
from torch.profiler import profile, ProfilerActivity

dev = "cuda:0"
T, R = 128, 4096
model = N.FusedPlainNeRF(steps=T, t_near=2, t_far=6, intermediate_size=64, sigmoid_kind="upshifted", precision="fp16")
model.load_state_dict(O.make_plain_params(1337, 64, 1.0), strict=True)
model = model.to(dev).train()
opt = N.autograd.FusedAdam(model.parameters(), lr=5e-4, eps=1e-7)
view = O.make_rays(1, 800, 800, size=800, seed=500).reshape(-1, 6)
g = torch.Generator().manual_seed(1)
rays = view[torch.randperm(view.shape[0], generator=g)[:R]].reshape(1, 64, 64, 6).contiguous().to(dev)
tgt = torch.rand(1, 64, 64, 3, generator=g).to(dev)
def step():
  opt.zero_grad(set_to_none=True)
  loss = ((model(rays) - tgt) ** 2).sum() / 3.0
  loss.backward(); opt.step()
if "--render" in sys.argv:     # also a few full-frame inference renders (so that one process feeds every ncu capture)
  model.eval()
  rays_f = view.contiguous().to(dev).reshape(1, 800, 800, 6)
  with torch.no_grad():
    for _ in range(5): model(rays_f)
  model.train()
for _ in range(5): step()
torch.cuda.synchronize()
if "--no-profiler" in sys.argv: sys.exit(0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): step()
e1.record(); torch.cuda.synchronize()
print("ms/step (events, 20 steps):", e0.elapsed_time(e1) / 20)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
  for _ in range(5): step()
  torch.cuda.synchronize()
rows = {}
for ev in prof.events():
  if ev.device_type.name == "CUDA":
    k = ev.name[:70]
    r = rows.setdefault(k, [0, 0.0]); r[0] += 1; r[1] += ev.device_time
tot = sum(v[1] for v in rows.values())
print(f"GPU kernel time per step: {tot / 5 / 1e3:.3f} ms")
for k, v in sorted(rows.items(), key=lambda kv: -kv[1][1])[:25]:
  print(f"{v[1] / 5 / 1e3:8.3f} ms  x{v[0] / 5:5.1f}  {k}")

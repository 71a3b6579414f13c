import os, sys, json
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
import nerf_atlas_b200 as N
from oracle import nerf_oracle as O
from helpers import plain_param_list
dev = torch.device("cuda", 0)
def timed(fn, reps=3):
  fn(); torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(reps): fn()
  e1.record(); torch.cuda.synchronize()
  return e0.elapsed_time(e1) / reps
rays800 = O.make_rays(1, 800, 800, size=800, seed=0).reshape(-1, 6).contiguous().to(dev)
ts128 = torch.linspace(2, 6, 128, device=dev)
out = {}
P = O.make_plain_params(1337, 64, 1.0)
e = N.RenderEngine(N.describe_plain(64, "upshifted", "black"), "fp16"); e._p = plain_param_list(P, dev); e.pack(e._p)
out["plain"] = timed(lambda: e.render(rays800, ts128, want_weights=False))
for mip in ("cone", "cylinder"):
  Pm = O.make_plain_params(61, 64, 1.0, mip=True)
  em = N.RenderEngine(N.describe_plain(64, "upshifted", "black", mip=mip), "fp16"); em._p = plain_param_list(Pm, dev); em.pack(em._p)
  rad = em.ray_radii(rays800.reshape(1, 800, 800, 6)).reshape(-1)
  out["mip_" + mip] = timed(lambda: em.render(rays800, ts128, radius=rad, want_weights=False))
Pp = O.make_plain_params(81, 64, 1.0, refl_kind="pos")
mp = N.FusedPlainNeRF(steps=128, t_near=2, t_far=6, intermediate_size=64, sigmoid_kind="upshifted", precision="fp16", refl_kind="pos", keep_weights=False)
mp.load_state_dict(Pp, strict=True); mp = mp.to(dev).eval()
def fp():
  with torch.no_grad(): return mp(rays800.reshape(1, 800, 800, 6))
out["positional"] = timed(fp)
print(os.environ.get("NF_LIB", "default"), json.dumps({k: round(v, 2) for k, v in out.items()}))

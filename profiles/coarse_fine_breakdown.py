import os, sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import nerf_atlas_b200 as N
from oracle import nerf_oracle as O
from helpers import plain_param_list
dev = torch.device("cuda", 0)
P = O.make_plain_params(1337, 64, 1.0)
eng = N.RenderEngine(N.describe_plain(64, "upshifted", "black"), "fp16"); eng._p = plain_param_list(P, dev); eng.pack(eng._p)
rays = O.make_rays(1, 800, 800, size=800, seed=0).reshape(-1, 6).contiguous().to(dev)
tsc = torch.linspace(2, 6, 64, device=dev); u = torch.rand(rays.shape[0], 128, device=dev)
def timed(fn, reps=3):
  fn(); torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(reps): fn()
  e1.record(); torch.cuda.synchronize()
  return e0.elapsed_time(e1) / reps
rgb_c, _, w_c = eng.render(rays, tsc, want_weights=True)
print("coarse render (T=64, weights out) ms", timed(lambda: eng.render(rays, tsc, want_weights=True)))
print("sample_pdf ms", timed(lambda: eng.sample_pdf(tsc, w_c, u)))
ts_f = eng.sample_pdf(tsc, w_c, u)
print("fine render (T=192 per-ray ts) ms", timed(lambda: eng.render(rays, ts_f, want_weights=False)))
ts192 = torch.linspace(2, 6, 192, device=dev)
print("render T=192 shared ts ms", timed(lambda: eng.render(rays, ts192, want_weights=False)))

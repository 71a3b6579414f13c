#!/usr/bin/env python
"""Secondary timings (not the bench line): the other BASELINE.json configurations on one B200, tensor pipeline
(+ config 3, Mip IPE 800x800x128, and the Positional head, both in the single-tile wide-x0 mode).
    python profiles/configs_bench.py > gpurun_out/configs.json
(2) PlainNeRF coarse 64 + fine 64+128 on 800x800; (4) VolSDF (SIREN SDF) 256 samples/ray; (5) D-NeRF 400x400x64."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import nerf_atlas_b200 as N
from oracle import nerf_oracle as O   # synthetic parameters / rays only
from helpers import plain_param_list, volsdf_param_list
dev = torch.device("cuda", 0)

def timed(fn, reps=3):
  fn(); torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(reps): fn()
  e1.record(); torch.cuda.synchronize()
  return e0.elapsed_time(e1) / reps

rows = []
rays800 = O.make_rays(1, 800, 800, size=800, seed=0).reshape(-1, 6).contiguous().to(dev)
# (2) coarse + fine
P = O.make_plain_params(1337, 64, 1.0)
eng = N.RenderEngine(N.describe_plain(64, "upshifted", "black"), "fp16"); eng._p = plain_param_list(P, dev); eng.pack(eng._p)
tsc = torch.linspace(2, 6, 64, device=dev)
u = torch.rand(rays800.shape[0], 128, device=dev)
ms = timed(lambda: eng.render_coarse_fine(rays800, tsc, u, want_weights=False))
rows.append({"config": "2: PlainNeRF coarse 64 + fine 64+128, 800x800", "ms_per_frame": ms, "rays_per_s": 640000 / ms * 1e3,
             "samples_per_ray": 64 + 192})
# (4) VolSDF, SIREN SDF, 256 samples per ray, unit directions, near 0.3 far 1.8
Pv = O.make_volsdf_params(7, "siren", 64, 0.1)
ev = N.RenderEngine(N.describe_volsdf("siren", 64, "upshifted"), "fp16"); ev._p = volsdf_param_list(Pv, "siren", dev); ev.pack(ev._p)
rv = rays800[:320000].clone(); rv[:, 3:] = torch.nn.functional.normalize(rv[:, 3:], dim=-1)
ts256 = torch.linspace(0.3, 1.8, 256, device=dev)
ms = timed(lambda: ev.render(rv, ts256, want_weights=False))
rows.append({"config": "4: VolSDF (SIREN SDF + View), 256 samples/ray, 320k rays", "ms": ms, "rays_per_s": 320000 / ms * 1e3, "samples_per_ray": 256})
# (5) D-NeRF 400x400x64, direct deformation
Pd = O.make_dnerf_params(9, 64)
canon = N.FusedPlainNeRF(steps=64, t_near=2, t_far=6, intermediate_size=64, sigmoid_kind="upshifted", precision="fp16", keep_weights=False)
m = N.FusedDynamicNeRF(canon); m.load_state_dict(Pd, strict=True); m = m.to(dev).eval()
r400 = O.make_rays(1, 400, 400, size=400, seed=1).to(dev)
tt = torch.tensor([0.4], device=dev)
def f():
  with torch.no_grad(): return m((r400, tt))
ms = timed(f)
rows.append({"config": "5: D-NeRF direct deformation + canonical PlainNeRF, 400x400x64", "ms_per_frame": ms, "rays_per_s": 160000 / ms * 1e3,
             "samples_per_ray": 64})
canon.precision = "fp32"
ms32 = timed(f, reps=1)
rows.append({"config": "5 (fp32 CUDA-core pipeline, for comparison)", "ms_per_frame": ms32, "rays_per_s": 160000 / ms32 * 1e3, "samples_per_ray": 64})
# (3) PlainNeRF + Mip IPE (cylinder as intended, and the reference's bug-compatible layout), 800x800x128, tensor pipeline (single-tile mode)
for mip in ("cylinder", "cylinder_ref", "cone"):
  Pm = O.make_plain_params(61, 64, 1.0, mip=True)
  em = N.RenderEngine(N.describe_plain(64, "upshifted", "black", mip=mip), "fp16"); em._p = plain_param_list(Pm, dev); em.pack(em._p)
  r4 = rays800.reshape(1, 800, 800, 6)
  rad = em.ray_radii(r4).reshape(-1)
  ts128 = torch.linspace(2, 6, 128, device=dev)
  ms = timed(lambda: em.render(rays800, ts128, radius=rad, want_weights=False), reps=2)
  rows.append({"config": f"3: PlainNeRF + Mip IPE ({mip}), 800x800x128", "ms_per_frame": ms, "rays_per_s": 640000 / ms * 1e3, "samples_per_ray": 128})
# Positional head (--refl-kind pos), 800x800x128
Pp = O.make_plain_params(81, 64, 1.0, refl_kind="pos")
mp = N.FusedPlainNeRF(steps=128, t_near=2, t_far=6, intermediate_size=64, sigmoid_kind="upshifted", precision="fp16", refl_kind="pos", keep_weights=False)
mp.load_state_dict(Pp, strict=True); mp = mp.to(dev).eval()
r4 = rays800.reshape(1, 800, 800, 6)
def fp():
  with torch.no_grad(): return mp(r4)
ms = timed(fp, reps=2)
rows.append({"config": "PlainNeRF + Positional head, 800x800x128", "ms_per_frame": ms, "rays_per_s": 640000 / ms * 1e3, "samples_per_ray": 128})
# f-4: sphere-traced surface render (SIREN SDF fitted to a unit sphere = the golden's weights), 800x800, 192 march iterations
from helpers import load_golden, sdf_params
fxs = load_golden("sdf_siren_march")
for prec in ("fp16", "fp32"):
  ms_ = N.FusedSDF("siren", 64, t_near=2.0, t_far=6.0, sigmoid_kind="upshifted", precision=prec)
  ms_.load_state_dict(sdf_params(fxs), strict=True); ms_ = ms_.to(dev).eval()
  ru = rays800.clone(); ru[:, 3:] = torch.nn.functional.normalize(ru[:, 3:], dim=-1)
  def fs():
    with torch.no_grad(): return ms_(ru.reshape(1, 800, 800, 6))
  ms = timed(fs, reps=2)
  rows.append({"config": f"f-4: SDF surface render (sphere march 192 it + View), 800x800, {prec}", "ms_per_frame": ms, "rays_per_s": 640000 / ms * 1e3,
               "hit_fraction": float(ms_.hit.float().mean())})
# f-4: the bisect intersection (193 + 35 SDF evaluations of EVERY ray) + View
mb = N.FusedSDF("siren", 64, t_near=2.0, t_far=6.0, sigmoid_kind="upshifted", precision="fp16", isect="bisect")
mb.load_state_dict(sdf_params(fxs), strict=True); mb = mb.to(dev).eval(); mb.jitter = 0.5
def fb():
  with torch.no_grad(): return mb(ru.reshape(1, 800, 800, 6))
ms = timed(fb, reps=2)
rows.append({"config": "f-4: SDF surface render (bisect: 193 samples + 32 bisection steps + View), 800x800, fp16", "ms_per_frame": ms, "rays_per_s": 640000 / ms * 1e3,
             "hit_fraction": float(mb.hit.float().mean())})
for r in rows: print(json.dumps(r), flush=True)

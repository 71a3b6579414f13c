#!/usr/bin/env python
"""Measured errors behind the stated tolerances of the tensor pipeline for the cases whose bars are looser than 1e-3
(run on a GPU box): stand-alone MLPs, VolSDF Fourier-MLP, DynamicNeRF (19 Linears)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import nerf_atlas_b200 as N
from oracle import nerf_oracle as O
from helpers import load_golden, plain_engine, volsdf_engine, psnr
DEV = "cuda:0"
P = O.make_plain_params(1337, 64, 20.0)
eng = plain_engine(P, DEV)
for which, pre, act, width in ((0, "first", "leaky_relu", 38), (1, "refl.mlp", "sin", 69)):
  g = torch.Generator().manual_seed(10 + which)
  x0 = torch.randn(1000, width, generator=g)
  out = eng.mlp_forward(which, x0.to(DEV), precision="fp16").cpu()
  refq = O.skip_mlp(x0, P, pre, act, quant=torch.float16); ref = O.skip_mlp(x0, P, pre, act)
  sc = max(float(ref.abs().max()), 1.0)
  print(f"mlp_forward {pre}: scale {sc:.2f}  vs emulation {float((out - refq).abs().max()) / sc:.2e}  vs fp32 {float((out - ref).abs().max()) / sc:.2e}  (relative to scale)")
for name in ("volsdf_siren_t32", "volsdf_mlp_t32"):
  fx = load_golden(name); kind = str(fx["sdf_kind"])
  Pv = O.make_volsdf_params(int(fx["seed"]), kind, 64, 0.1)
  rays = O.make_rays(int(fx["B"]), int(fx["H"]), int(fx["W"]), 800, int(fx["seed"]), int(fx["top"]), int(fx["left"])).reshape(-1, 6)
  e = volsdf_engine(Pv, kind, DEV, str(fx["sigmoid"]), "fp16")
  rgb, alpha, w = e.render(rays.to(DEV), torch.from_numpy(fx["ts"]).to(DEV))
  out = rgb.cpu().numpy().reshape(fx["out"].shape)
  with torch.no_grad(): q = O.volsdf_forward(Pv, rays, torch.from_numpy(fx["ts"]), sdf_kind=kind, sigmoid=str(fx["sigmoid"]), quant=torch.float16)["out"].numpy().reshape(fx["out"].shape)
  print(f"{name}: fp16 vs fp32 golden {np.abs(out - fx['out']).max():.2e} psnr {psnr(out, fx['out']):.1f}  vs emulation {np.abs(out - q).max():.2e}  weights {np.abs(w.cpu().numpy().T.reshape(fx['weights'].shape) - fx['weights']).max():.2e}")
for name, spline in (("dnerf_direct_t64", 0), ("dnerf_spline5_t32", 5), ("dnerf_spline4_t32", 4)):
  fx = load_golden(name)
  Pd = O.make_dnerf_spline_params(int(fx["seed"]), spline, 64) if spline else O.make_dnerf_params(int(fx["seed"]), 64)
  rays = O.make_rays(int(fx["B"]), int(fx["H"]), int(fx["W"]), 800, int(fx["seed"]), int(fx["top"]), int(fx["left"]))
  times = torch.from_numpy(fx["times"])
  canon = N.FusedPlainNeRF(steps=int(fx["T"]), t_near=float(fx["near"]), t_far=float(fx["far"]), intermediate_size=64, sigmoid_kind=str(fx["sigmoid"]), bg=str(fx["bg"]), precision="fp16")
  m = N.FusedDynamicNeRF(canon, spline=spline); m.load_state_dict(Pd, strict=True); m = m.to(DEV).eval()
  with torch.no_grad(): out = m((rays.to(DEV), times.to(DEV))).cpu().numpy()
  ts = O.compute_ts(float(fx["near"]), float(fx["far"]), int(fx["T"]))
  with torch.no_grad():
    refq = (O.dnerf_spline_forward(Pd, rays, times, ts, spline, quant=torch.float16) if spline else O.dnerf_direct_forward(Pd, rays, times, ts, quant=torch.float16))["out"].numpy()
  print(f"{name}: fp16 vs fp32 golden {np.abs(out - fx['out']).max():.2e} psnr {psnr(out, fx['out']):.1f}  vs emulation {np.abs(out - refq).max():.2e}")

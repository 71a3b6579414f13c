"""Shared helpers for the parity tests (CUDA path through the C ABI vs the CPU oracle)."""
import os
import numpy as np
import torch
from oracle import nerf_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

def load_golden(name):
  fx = np.load(os.path.join(GOLDEN, name + ".npz"))
  return {k: fx[k] for k in fx.files}

def plain_param_list(P, device):
  """Order documented at nf_param_count (include/nerf_b200.h)."""
  names = []
  for pre in ("first", "refl.mlp"):
    names += [f"{pre}.init.weight", f"{pre}.init.bias"]
    i = 0
    while f"{pre}.layers.{i}.weight" in P:
      names += [f"{pre}.layers.{i}.weight", f"{pre}.layers.{i}.bias"]; i += 1
    names += [f"{pre}.out.weight", f"{pre}.out.bias"]
  names += [f"first.enc.embs.{i}.weight" for i in range(8)]
  return [P[n].to(device).contiguous() for n in names]

def plain_engine(P, device, sigmoid="upshifted", bg="black", precision="fp16"):
  import nerf_atlas_b200 as N
  eng = N.RenderEngine(N.describe_plain(64, sigmoid, bg), precision)
  eng._params = plain_param_list(P, device)   # keep alive
  eng.pack(eng._params)
  return eng

def make_tiny_params(seed=7):
  """TinyNeRF.estim (reference src/nerf.py:286-290): xavier-uniform weights, zero biases."""
  import math
  g = np.random.default_rng(seed)
  P = {}
  def xav(name, o, i):
    a = math.sqrt(6.0 / (i + o))
    P[f"{name}.weight"] = torch.from_numpy(g.uniform(-a, a, size=(o, i)).astype(np.float32))
    P[f"{name}.bias"] = torch.from_numpy(g.uniform(-0.1, 0.1, size=(o,)).astype(np.float32))
  xav("estim.init", 256, 3)
  for i in range(6): xav(f"estim.layers.{i}", 256, 259 if i in (0, 3) else 256)
  xav("estim.out", 4, 256)
  return P

def tiny_param_list(P, device):
  names = ["estim.init.weight", "estim.init.bias"]
  for i in range(6): names += [f"estim.layers.{i}.weight", f"estim.layers.{i}.bias"]
  names += ["estim.out.weight", "estim.out.bias"]
  return [P[n].to(device).contiguous() for n in names]

def psnr(a, b):
  mse = float(np.mean((np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)) ** 2))
  return 200.0 if mse == 0 else -10 * np.log10(mse)

def volsdf_param_list(P, sdf_kind, device):
  pre = "sdf.underlying.siren" if sdf_kind == "siren" else "sdf.underlying.mlp"
  names = []
  for q in (pre, "sdf.refl.mlp"):
    names += [f"{q}.init.weight", f"{q}.init.bias"]
    i = 0
    while f"{q}.layers.{i}.weight" in P:
      names += [f"{q}.layers.{i}.weight", f"{q}.layers.{i}.bias"]; i += 1
    names += [f"{q}.out.weight", f"{q}.out.bias"]
  ps = [P[n].to(device).contiguous() for n in names]
  if sdf_kind == "mlp": ps.append(P[f"{pre}.enc.basis"].to(device).contiguous())
  ps.append(P["scale"].reshape(1).to(device))
  return ps

def volsdf_engine(P, sdf_kind, device, sigmoid="upshifted", precision="fp16"):
  import nerf_atlas_b200 as N
  eng = N.RenderEngine(N.describe_volsdf(sdf_kind, 64, sigmoid), precision)
  eng._params = volsdf_param_list(P, sdf_kind, device)
  eng.pack(eng._params)
  return eng

def trained_params(fx):
  """Weight set T (tests/golden/make_golden.py case_trained): the seeded initialisation + the stored deltas = the parameters the
  REFERENCE reached by training itself on the procedural scene."""
  P = O.make_plain_params(int(fx["seed"]), 64, 1.0)
  for k in fx:
    if k.startswith("param."): P[k[6:]] = torch.from_numpy(fx[k])
    elif k.startswith("rows."):
      t = P[k[5:]].clone(); t[torch.from_numpy(fx[k]).long()] = torch.from_numpy(fx["vals." + k[5:]]); P[k[5:]] = t
  return P

def sdf_params(fx):
  """Parameters of the `sdf_*_march` goldens (stored as fp16; the reference rendered with exactly these values)."""
  if "params_from" in fx: fx = load_golden(str(fx["params_from"]))         # the bisect / secant goldens share the sphere-march golden's fit
  return {k[len("param16."):]: torch.from_numpy(fx[k].astype(np.float32)) for k in fx if k.startswith("param16.")}

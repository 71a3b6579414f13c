#!/usr/bin/env python
"""Time of one native optimiser step (training forward + nf_render_backward + FusedAdam) for every trainable model kind:
4096 rays x 128 samples, training mode (jitter, density noise where the reference has it), CUDA events over 30 steps after 5 warm-ups.
    python profiles/train_kinds.py          # on a B200"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nerf_atlas_b200 as N
from nerf_atlas_b200.autograd import FusedAdam
from oracle import nerf_oracle as O   # synthetic parameters / rays only

dev = "cuda:0"
def build(kind):
  if kind == "volsdf_siren":
    m = N.FusedVolSDF(sdf_kind="siren", steps=128, t_near=2, t_far=6, intermediate_size=64, sigmoid_kind="thin", precision="fp16")
    m.load_state_dict(O.make_volsdf_params(7, "siren", 64, 0.1), strict=True)
  else:
    m = N.FusedPlainNeRF(steps=128, t_near=2, t_far=6, intermediate_size=64, sigmoid_kind="upshifted", precision="fp16", refl_kind=kind)
    m.load_state_dict(O.make_plain_params(7, 64, 1.0, refl_kind=kind), strict=True)
  return m.to(dev).train()

for kind in ("view", "pos", "volsdf_siren"):
  m = build(kind); m.keep_weights = False
  opt = FusedAdam([p for p in m.parameters() if p.requires_grad and p.numel() > 0], lr=5e-4)
  rays = O.make_rays(1, 64, 64, seed=3, crop_top=368, crop_left=368).to(dev)
  tgt = torch.rand(1, 64, 64, 3, device=dev)
  def step():
    opt.zero_grad(set_to_none=False)
    loss = torch.nn.functional.mse_loss(m(rays), tgt); loss.backward(); opt.step(); return loss
  first = float(step().detach())
  for _ in range(4): step()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(30): l = step()
  e1.record(); torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / 30
  print(json.dumps({"kind": kind, "ms_per_step": round(ms, 3), "it_per_s": round(1e3 / ms, 1), "rays": 4096, "T": 128, "loss_first": round(first, 5), "loss_last": round(float(l.detach()), 5)}), flush=True)

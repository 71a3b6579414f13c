#!/usr/bin/env python
"""Time the tensor pipelines / ring geometries / wait modes on the bench workload (800x800x128 Plain+View) in ONE process.

    python profiles/perf_variants.py [--steps 4] > gpurun_out/variants.json

The C ABI reads NF_TC_PIPE / NF_TC_RING / NF_TC_DEBUG at every launch, so variants are switched with os.environ."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import nerf_atlas_b200 as N
from oracle import nerf_oracle as O   # synthetic inputs only

VARIANTS = [
  ("pipe3_ring3x16", {"NF_TC_PIPE": "3", "NF_TC_RING": "3"}),
  ("pipe3_ring3x16_24_epilogue_warps", {"NF_TC_PIPE": "3", "NF_TC_RING": "3", "NF_TC_EPIW": "24"}),
  ("pipe3_ring3x16_24_epilogue_warps_plain_sin", {"NF_TC_PIPE": "3", "NF_TC_RING": "3", "NF_TC_EPIW": "24", "NF_TC_DEBUG": "2048"}),
  ("pipe3_ring3x16_plain_sin_epilogue", {"NF_TC_PIPE": "3", "NF_TC_RING": "3", "NF_TC_DEBUG": "2048"}),
  ("pipe1_single_cta", {"NF_TC_PIPE": "1"}),
]
def main():
  ap = argparse.ArgumentParser(); ap.add_argument("--steps", type=int, default=4); ap.add_argument("--views", type=int, default=4)
  args = ap.parse_args()
  dev = torch.device("cuda", 0)
  model = N.FusedPlainNeRF(steps=128, t_near=2, t_far=6, intermediate_size=64, sigmoid_kind="upshifted", precision="fp16")
  model.load_state_dict(O.make_plain_params(1337, 64, 1.0), strict=True)
  model = model.to(dev).eval()
  eng = model.engine(); eng.pack(model._param_list())
  views = [O.make_rays(1, 800, 800, size=800, seed=v).reshape(-1, 6).contiguous().to(dev) for v in range(args.views)]
  ts = torch.linspace(2, 6, 128, device=dev)
  base = None; rows = []
  for name, env in VARIANTS:
    for k in ("NF_TC_PIPE", "NF_TC_RING", "NF_TC_DEBUG", "NF_TC_EPIW"): os.environ.pop(k, None)
    os.environ.update(env)
    try:
      for i in range(2): eng.render(views[i % len(views)], ts, None, want_weights=False)
      torch.cuda.synchronize()
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record()
      for i in range(args.steps): out = eng.render(views[i % len(views)], ts, None, want_weights=False)[0]
      e1.record(); torch.cuda.synchronize()
      ms = e0.elapsed_time(e1) / args.steps
      ref = eng.render(views[0], ts, None, want_weights=False)[0]
      if base is None: base = ref.clone()
      row = {"variant": name, "env": env, "ms_per_frame": ms, "rays_per_s": 640000 / (ms * 1e-3),
             "max_abs_diff_vs_first_variant": float((ref - base).abs().max()), "finite": bool(torch.isfinite(ref).all())}
    except Exception as ex:
      row = {"variant": name, "env": env, "error": str(ex)[:300]}
    rows.append(row); print(json.dumps(row), flush=True)

if __name__ == "__main__":
  main()

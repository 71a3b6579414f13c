#!/usr/bin/env python
"""A/B timing of differently built libraries on the bench workload (800x800x128 Plain+View, rgb only), one subprocess per library:

    NF_BUILD_DEFS="NF_LAG=3" NF_BUILD_TAG=lag3 python -m nerf_atlas_b200.build     # here (no GPU needed)
    python profiles/ab_time.py libnerf_b200.so libnerf_b200_lag3.so ...            # on the GPU box

Prints ms/frame (mean of --steps after 2 warm-ups, views rotating), and the max difference of the rendered frame to the first library's."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

def one(steps: int, dump: str):
  sys.path.insert(0, ROOT)
  import torch
  import nerf_atlas_b200 as N
  from oracle import nerf_oracle as O   # synthetic inputs only
  dev = torch.device("cuda", 0)
  model = N.FusedPlainNeRF(steps=128, t_near=2, t_far=6, intermediate_size=64, sigmoid_kind="upshifted", precision="fp16")
  model.load_state_dict(O.make_plain_params(1337, 64, 1.0), strict=True)
  model = model.to(dev).eval()
  eng = model.engine(); eng.pack(model._param_list())
  views = [O.make_rays(1, 800, 800, size=800, seed=v).reshape(-1, 6).contiguous().to(dev) for v in range(4)]
  ts = torch.linspace(2, 6, 128, device=dev)
  for i in range(2): eng.render(views[i], ts, None, want_weights=False)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for i in range(steps): eng.render(views[i % 4], ts, None, want_weights=False)
  e1.record(); torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / steps
  ref = eng.render(views[0], ts, None, want_weights=False)[0].cpu()
  diff = None
  if os.path.exists(dump): diff = float((ref - torch.load(dump)).abs().max())
  else: torch.save(ref, dump)
  print(json.dumps({"lib": os.environ.get("NF_LIB", "libnerf_b200.so"), "ms_per_frame": round(ms, 3), "max_diff_vs_first": diff, "finite": bool(torch.isfinite(ref).all())}), flush=True)

if __name__ == "__main__":
  if sys.argv[1] == "--one":
    one(int(sys.argv[2]), sys.argv[3])
  else:
    dump = "/tmp/ab_time_first.pt"
    if os.path.exists(dump): os.remove(dump)
    steps = os.environ.get("AB_STEPS", "6")
    for lib in sys.argv[1:]:
      env = dict(os.environ, NF_LIB=lib)
      r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", steps, dump], env=env, capture_output=True, text=True, timeout=300)
      out = [l for l in r.stdout.splitlines() if l.startswith("{")]
      print(out[-1] if out else json.dumps({"lib": lib, "error": (r.stderr or r.stdout)[-400:]}), flush=True)

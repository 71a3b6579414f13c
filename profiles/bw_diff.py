#!/usr/bin/env python
"""Differential stress of the boundary-warp kernels: random (model, T, ray count) cases rendered by the product library and by a
build without boundary warps (NF_BUILD_DEFS="NF_BW=0" NF_BUILD_TAG=nobw) in two subprocesses; the frames must be bit-identical
(same arithmetic, different warps), every case repeated to catch timing-dependent races.
    python profiles/bw_diff.py [n_cases]"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

def worker(n_cases, out):
  sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
  import torch, random
  import nerf_atlas_b200 as N
  from oracle import nerf_oracle as O
  from helpers import plain_param_list, plain_engine
  dev = "cuda:0"
  rnd = random.Random(1234)
  P = O.make_plain_params(1337, 64, 20.0)
  Pp = O.make_plain_params(81, 64, 20.0, refl_kind="pos")
  Pm = O.make_plain_params(62, 64, 20.0, mip=True)
  engines = {"plain": plain_engine(P, dev, precision="fp16"), "plain_white": plain_engine(P, dev, bg="white", precision="fp16")}
  mp = N.FusedPlainNeRF(steps=128, t_near=2, t_far=6, intermediate_size=64, sigmoid_kind="upshifted", precision="fp16", refl_kind="pos", keep_weights=False)
  mp.load_state_dict(Pp, strict=True); mp = mp.to(dev).eval()
  ep = mp.engine(); ep.pack(mp._param_list()); engines["pos"] = ep
  em = N.RenderEngine(N.describe_plain(64, "upshifted", "black", mip="cone"), "fp16"); em._p = plain_param_list(Pm, dev); em.pack(em._p); engines["mip"] = em
  res = {}
  for c in range(n_cases):
    kind = rnd.choice(list(engines)); T = rnd.choice([32, 64, 96, 128, 160, 192, 256]); h = rnd.choice([3, 5, 17, 40]); w = rnd.choice([1, 2, 7, 33, 90])
    rays = O.make_rays(1, h, w, seed=c, crop_top=rnd.randrange(0, 700), crop_left=rnd.randrange(0, 700))
    flat = rays.reshape(-1, 6).to(dev); ts = torch.linspace(2, 6, T, device=dev)
    e = engines[kind]
    kw = {}
    if kind == "mip": kw["radius"] = e.ray_radii(rays.to(dev)).reshape(-1)
    outs = [e.render(flat, ts, want_weights=False, **kw)[0].clone() for _ in range(3)]
    torch.cuda.synchronize()
    assert all(torch.equal(outs[0], o) for o in outs[1:]), ("not deterministic", kind, T, h, w)
    res[f"{c}:{kind}:T{T}:{h}x{w}"] = outs[0].cpu()
  torch.save(res, out)

if __name__ == "__main__":
  if sys.argv[1] == "--worker":
    worker(int(sys.argv[2]), sys.argv[3])
  else:
    n = sys.argv[1] if len(sys.argv) > 1 else "40"
    outs = []
    for lib in ("libnerf_b200.so", "libnerf_b200_nobw.so"):
      out = f"/tmp/bw_diff_{lib}.pt"
      r = subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", n, out], env=dict(os.environ, NF_LIB=lib), capture_output=True, text=True, timeout=1200)
      if r.returncode != 0: print(lib, "FAILED", (r.stderr or r.stdout)[-600:]); sys.exit(1)
      outs.append(out)
    import torch
    a, b = torch.load(outs[0]), torch.load(outs[1])
    bad = [k for k in a if not torch.equal(a[k], b[k])]
    worst = max((float((a[k] - b[k]).abs().max()) for k in a), default=0.0)
    print(json.dumps({"cases": len(a), "bit_identical": len(a) - len(bad), "differing": bad[:10], "max_abs_diff": worst}))

"""Per-Linear error table of the native backward (run on a GPU box): stashed operands, dL/dz of every Linear and the final
parameter gradients against torch autograd through the oracle, fp32 and fp16-operand emulation."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import nerf_oracle as O
import test_gpu_training as TT

def main(T=16, R=12, seed=91, sigmoid="upshifted", bg="black", with_noise=False):
  print(f"==== T={T} R={R} seed={seed} {sigmoid} {bg} noise={with_noise}")
  P = O.make_plain_params(seed, 64, 20.0)
  rays = O.make_rays(1, 20, 20, 800, seed, 390, 390).reshape(-1, 6)[:R]
  ts = torch.linspace(2, 6, T)
  gen = torch.Generator().manual_seed(1)
  target = torch.rand(R, 3, generator=gen)
  noise = torch.randn(T, R, generator=gen) * 0.2 if with_noise else None
  rec = []; recq = []
  _, _, g_ref = TT.oracle_grads(P, rays, ts, target, sigmoid, bg, noise, record=rec)
  # fp16-operand emulation: quantised forward, autograd straight through the casts
  orig = O.plain_forward
  O.plain_forward = lambda *a, **k: orig(*a, quant=torch.float16, **k)
  try: _, _, g_q = TT.oracle_grads(P, rays, ts, target, sigmoid, bg, noise, record=recq)
  finally: O.plain_forward = orig
  eng, lay, ws, rgb, grads = TT.run_native(P, rays, ts, target, sigmoid, bg, noise)
  n_tiles = int(lay.n_tiles)
  S = float(ws[lay.scale_off: lay.scale_off + 4].view(torch.float32).item())
  print("scale", S, "tiles", n_tiles)
  if 128 % T == 0:
    rows = TT.rows_to_samples(R, T)
    for li in range(12):
      L = lay.lin[li]
      G = TT.decode_tiles(ws, L.g_off, L.g_tile, L.n_pad, n_tiles)[rows] / S
      out = []
      for r in (rec, recq):
        gz = r[li][1].grad
        if L.m == 0 and L.j == 5: gz = torch.cat([gz[:, 1:], gz[:, :1]], dim=1)
        out.append(float((G[:, :gz.shape[1]] - gz).abs().max()) / float(gz.abs().max()))
      print(f"G li={li:2d} rel-max-err vs fp32 {out[0]:.2e}  vs fp16-emulation {out[1]:.2e}")
  for name, g in zip(TT.PARAM_NAMES, grads):
    g = g.cpu()
    e32 = float((g - g_ref[name]).abs().max()) / max(float(g_ref[name].abs().max()), 1e-30)
    e16 = float((g - g_q[name]).abs().max()) / max(float(g_q[name].abs().max()), 1e-30)
    cos = float(torch.nn.functional.cosine_similarity(g.reshape(1, -1), g_ref[name].reshape(1, -1)))
    print(f"{name:28s} rel-max-err vs fp32 {e32:.2e}  vs fp16-emulation {e16:.2e}  cos {cos:.6f}  finite {bool(torch.isfinite(g).all())}")

if __name__ == "__main__":
  main(T=64, R=397, seed=1337, sigmoid="thin", bg="white", with_noise=True)
  main(T=64, R=397, seed=1337, sigmoid="upshifted", bg="white", with_noise=False)
  main(T=64, R=397, seed=1337, sigmoid="upshifted", bg="black", with_noise=True)
  main(T=100, R=233, seed=1337, sigmoid="fat", bg="white")

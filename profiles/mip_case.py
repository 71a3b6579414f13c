"""One Mip render on the tensor pipeline for a given (mip kind, T, H, W): the case runner used to find the scratch-ready barrier aliasing
(python profiles/mip_case.py cylinder 192 14 19; under compute-sanitizer for a trapped launch)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, numpy as np
import nerf_atlas_b200 as N
from oracle import nerf_oracle as O
from helpers import plain_param_list
dev = "cuda:0"
P = O.make_plain_params(62, 64, 20.0, mip=True)
e = N.RenderEngine(N.describe_plain(64, "upshifted", "black", mip=sys.argv[1]), "fp32"); e._p = plain_param_list(P, dev); e.pack(e._p)
T = int(sys.argv[2]); shape = (1, int(sys.argv[3]), int(sys.argv[4]))
slab = O.make_rays(*shape, seed=90 + T, crop_top=280, crop_left=300)
ts = torch.linspace(2, 6, T)
flat = slab.reshape(-1, 6).to(dev); rad = e.ray_radii(slab.to(dev)).reshape(-1)
whole = e.render(flat, ts.to(dev), radius=rad, precision="fp16", want_weights=False)[0]
torch.cuda.synchronize()
print(sys.argv[1:], "ok", float(whole.mean()))

import sys, math, torch
sys.path.insert(0, "/root/repo")
import bench, nerf_atlas_b200 as N
from oracle import nerf_oracle as O
dev = torch.device("cuda", 0)
model = N.FusedPlainNeRF(steps=128, t_near=2, t_far=6, intermediate_size=64, sigmoid_kind="upshifted", precision="fp16")
model.load_state_dict(O.make_plain_params(1337, 64, 1.0), strict=True)
model = model.to(dev).eval()
eng = model.engine(); eng.pack(model._param_list())
ts = torch.linspace(2, 6, 128, device=dev)
bench.cpu_port_rays_per_s(steps=1, warmup=0)
c_rays, c_ref = bench.cpu_port_rays_per_s.last
got = eng.render(c_rays.reshape(-1, 6).contiguous().to(dev), ts, None, want_weights=False)[0].reshape(c_ref.shape).cpu()
err = (got - c_ref).abs(); mse = float(((got - c_ref).double() ** 2).mean())
print({"max_abs_err_vs_cpu_fp32": float(err.max()), "psnr_db": -10 * math.log10(mse), "rays": int(c_rays.reshape(-1, 6).shape[0])})

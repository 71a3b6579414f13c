"""Generate the golden vectors in this directory by EXECUTING THE REFERENCE
(/root/reference, imported read-only with oracle/ref_shim.py) on seeded inputs.

Run in the build container:   python tests/golden/make_golden.py
The reference tree cannot travel to the GPU box, so the outputs are committed.
Parameters and rays are regenerated from seeds (numpy PCG64, platform
independent) by oracle.nerf_oracle.make_plain_params / make_rays, so a fixture
only stores the small outputs of the reference run.
"""
import os, sys
import numpy as np
import torch
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import nerf_oracle as O, ref_shim

def ref_plain(params, steps, train=False, refl_kind="view"):
  runner, nerf, refl, utils, cameras = ref_shim.load()
  model, args = ref_shim.build_model("plain", steps, extra=("--refl-kind", refl_kind))
  sd = {k: v.clone() for k, v in params.items()}
  model.load_state_dict(sd, strict=True)
  model.train(train)
  return model, args

def case_plain(name, seed, B, H, W, T, sigma_gain=1.0, train=False, top=0, left=0, stages=True, refl_kind="view"):
  params = O.make_plain_params(seed, 64, sigma_gain, refl_kind=refl_kind)
  rays = O.make_rays(B, H, W, size=800, seed=seed, crop_top=top, crop_left=left)
  model, args = ref_plain(params, T, train, refl_kind)
  runner, nerf, refl, utils, cameras = ref_shim.load()
  draws = {}
  if train:
    # capture the two RNG draws of the training path (nerf.py:45 rand_like, nerf.py:348 randn_like)
    torch.manual_seed(seed)
    o_rand, o_randn = torch.rand_like, torch.randn_like
    def rl(x, *a, **k): r = o_rand(x, *a, **k); draws["rand"] = r.clone(); return r
    def rnl(x, *a, **k): r = o_randn(x, *a, **k); draws["randn"] = r.clone(); return r
    torch.rand_like, torch.randn_like = rl, rnl
  try:
    with torch.no_grad(): out = model(rays)
  finally:
    if train: torch.rand_like, torch.randn_like = o_rand, o_randn
  fx = dict(
    kind="plain", refl_kind=refl_kind, seed=seed, B=B, H=H, W=W, T=T, sigma_gain=sigma_gain, train=int(train), top=top, left=left,
    near=float(args.near), far=float(args.far), sigmoid=args.sigmoid_kind, bg=args.bg,
    ts=model.ts.numpy(), out=out.numpy(), alpha=model.alpha.numpy(), weights=model.weights.numpy(),
  )
  if train: fx["rand"] = draws["rand"].numpy(); fx["randn"] = draws["randn"].numpy()
  if stages:
    # per-stage tensors straight from the reference sub-modules
    with torch.no_grad():
      pts, ts, r_o, r_d, _ = nerf.compute_pts_ts(rays, args.near, args.far, T, perturb=0) if not train else \
        (r_o_pts(rays, model.ts))
      p = pts.reshape(-1, 3)
      fx["pts"] = pts.numpy()
      fx["hash_enc"] = model.first.enc(p).numpy()
      fo = model.first(pts, model.empty_latent.expand(pts.shape[:-1] + (0,)))
      fx["first_out"] = fo.numpy()
      fx["elaz"] = utils.dir_to_elev_azim(r_d).numpy()
      idx = []
      for lvl in range(8):
        N_l = model.first.enc.low_reso * (model.first.enc.scale ** lvl)
        l = (p * N_l).floor().long()
        h = l + 1
        lx, ly, lz = l.split([1, 1, 1], dim=-1); hx, hy, hz = h.split([1, 1, 1], dim=-1)
        cat = lambda x, y, z: torch.cat([x, y, z], dim=-1)
        vs = [l, cat(lx, ly, hz), cat(lx, hy, lz), cat(lx, hy, hz), cat(hx, ly, lz), cat(hx, ly, hz), cat(hx, hy, lz), h]
        idx.append(torch.stack([(model.first.enc.hash_fn(v) % model.first.enc.emb_size).squeeze(-1) for v in vs], 0))
      fx["hash_idx"] = torch.stack(idx, 0).numpy().astype(np.uint16)   # [level, corner, N]
  # the synthetic ray generator must equal the reference camera (cameras.py:45-66, runner.py:495-505)
  np.savez_compressed(os.path.join(HERE, name + ".npz"), **fx)
  print(name, "out", out.shape, "mean", float(out.mean()), "acc-before-last", float(model.weights[:-1].sum(0).mean()))

def case_volsdf(name, seed, sdf_kind, B, H, W, T, top=0, left=0):
  params = O.make_volsdf_params(seed, sdf_kind, 64, 0.1)
  rays = O.make_rays(B, H, W, size=800, seed=seed, crop_top=top, crop_left=left)
  model, args = ref_shim.build_model("volsdf", T, extra=("--sdf-kind", sdf_kind))
  model.load_state_dict({k: v.clone() for k, v in params.items()}, strict=True)
  model.eval()
  with torch.no_grad(): out = model(rays)
  fx = dict(kind="volsdf", sdf_kind=sdf_kind, seed=seed, B=B, H=H, W=W, T=T, top=top, left=left,
            near=float(args.near), far=float(args.far), sigmoid=args.sigmoid_kind,
            ts=model.ts.numpy(), out=out.numpy(), alpha=model.alpha.numpy(), weights=model.weights.numpy())
  np.savez_compressed(os.path.join(HERE, name + ".npz"), **fx)
  print(name, "out", out.shape, "mean", float(out.mean()), "acc-before-last", float(model.weights[:-1].sum(0).mean()))

def case_volsdf_grads(name, seed, B, H, W, T, top=0, left=0, beta=0.25):
  """Gradients of the reference's VolSDF (SIREN SDF + View, volume branch) through its own loss.backward(): every Linear of both
  MLPs (every 16th row of the 256-row matrices) and the learned `scale` (beta)."""
  params = O.make_volsdf_params(seed, "siren", 64, beta)
  rays = O.make_rays(B, H, W, size=800, seed=seed, crop_top=top, crop_left=left)
  model, args = ref_shim.build_model("volsdf", T, extra=("--sdf-kind", "siren"))
  model.load_state_dict({k: v.clone() for k, v in params.items()}, strict=True)
  model.eval()
  for p in model.parameters(): p.requires_grad_(p.dtype.is_floating_point and p.numel() > 0)
  g = np.random.default_rng(seed)
  target = torch.from_numpy(g.uniform(0, 1, size=(B, H, W, 3)).astype(np.float32))
  out = model(rays)
  loss = torch.nn.functional.mse_loss(out, target)
  loss.backward()
  fx = dict(kind="volsdf_grads", sdf_kind="siren", seed=seed, beta=beta, B=B, H=H, W=W, T=T, top=top, left=left, near=float(args.near), far=float(args.far),
            sigmoid=args.sigmoid_kind, target=target.numpy(), out=out.detach().numpy(), loss=float(loss), ts=model.ts.numpy())
  for n, p in model.named_parameters():
    if p.grad is None: continue
    gr = p.grad.numpy()
    fx["grad." + n] = gr[::16] if gr.ndim == 2 and gr.shape[0] == 256 else gr
  np.savez_compressed(os.path.join(HERE, name + ".npz"), **fx)
  print(name, "loss", float(loss), "d scale", float(model.scale.grad), "grads", sum(1 for k in fx if k.startswith("grad.")))

def case_dnerf(name, seed, B, H, W, T, top=0, left=0):
  runner, nerf, refl, utils, cameras = ref_shim.load()
  params = O.make_dnerf_params(seed, 64)
  rays = O.make_rays(B, H, W, size=800, seed=seed, crop_top=top, crop_left=left)
  times = torch.linspace(0.1, 0.9, B)
  canonical, args = ref_shim.build_model("plain", T)
  model = nerf.DynamicNeRF(canonical=canonical, spline=0)
  # shim (5) of SURVEY.md 8c: direct_predict reads self.dp, which the reference never sets (nerf.py:1265)
  def direct_predict(self, x, t):
    xt = torch.cat([x, t], dim=-1)
    dp, rigidity, enc_rigidity, enc = self.delta_estim(xt).split(self.mlp_out_layout, dim=-1)
    self.dp = dp
    self.rigidity = (rigidity / 2).sigmoid()
    self.rigid_dp = self.dp * self.rigidity
    return self.rigid_dp, enc * enc_rigidity.sigmoid()
  model.time_estim = direct_predict.__get__(model)
  model.load_state_dict({k: v.clone() for k, v in params.items()}, strict=True)
  model.eval()
  with torch.no_grad(): out = model((rays, times))
  fx = dict(kind="dnerf", seed=seed, B=B, H=H, W=W, T=T, top=top, left=left, near=float(args.near), far=float(args.far),
            sigmoid=args.sigmoid_kind, bg=args.bg, times=times.numpy(), ts=model.ts.numpy(), out=out.numpy(),
            alpha=canonical.alpha.numpy(), weights=canonical.weights.numpy(), rigid_dp=model.rigid_dp.numpy())
  np.savez_compressed(os.path.join(HERE, name + ".npz"), **fx)
  print(name, "out", out.shape, "mean", float(out.mean()), "max|rigid_dp|", float(model.rigid_dp.abs().max()))

def case_mip(name, seed, B, H, W, T, top=0, left=0):
  """PlainNeRF(mip=CylinderGaussian()) built directly (runner.load_model rebuilds the refl head without the mip latent,
  SURVEY.md a-4 (iii)); the cone variant renders NaN in the reference and cannot be pinned."""
  runner, nerf, refl, utils, cameras = ref_shim.load()
  params = O.make_plain_params(seed, 64, 20.0, mip=True)
  rays = O.make_rays(B, H, W, size=800, seed=seed, crop_top=top, crop_left=left)
  model = nerf.PlainNeRF(mip=utils.CylinderGaussian(), steps=T, t_near=2, t_far=6, intermediate_size=64,
                         sigmoid_kind="upshifted", bg="black").eval()
  model.load_state_dict({k: v.clone() for k, v in params.items()}, strict=True)
  with torch.no_grad():
    out = model(rays)
    pts, ts, r_o, r_d, _ = nerf.compute_pts_ts(rays, 2, 6, T, perturb=0)
    enc = model.mip_encoding(r_o, r_d, ts)
    rad = utils.radii_x(r_d)
  fx = dict(kind="plain_mip", mip="cylinder_ref", seed=seed, B=B, H=H, W=W, T=T, top=top, left=left, near=2.0, far=6.0,
            sigmoid="upshifted", bg="black", ts=model.ts.numpy(), out=out.numpy(), alpha=model.alpha.numpy(),
            weights=model.weights.numpy(), mip_enc=enc.numpy(), radii=rad.numpy())
  np.savez_compressed(os.path.join(HERE, name + ".npz"), **fx)
  print(name, "out", out.shape, "mean", float(out.mean()), "finite", bool(torch.isfinite(out).all()))

def case_dnerf_spline(name, seed, n, B, H, W, T, top=0, left=0):
  runner, nerf, refl, utils, cameras = ref_shim.load()
  params = O.make_dnerf_spline_params(seed, n, 64)
  rays = O.make_rays(B, H, W, size=800, seed=seed, crop_top=top, crop_left=left)
  times = torch.linspace(0.1, 0.9, B)
  canonical, args = ref_shim.build_model("plain", T)
  model = nerf.DynamicNeRF(canonical=canonical, spline=n)
  model.load_state_dict({k: v.clone() for k, v in params.items()}, strict=True)
  model.eval()
  with torch.no_grad(): out = model((rays, times))
  fx = dict(kind="dnerf_spline", n=n, seed=seed, B=B, H=H, W=W, T=T, top=top, left=left, near=float(args.near), far=float(args.far),
            sigmoid=args.sigmoid_kind, bg=args.bg, times=times.numpy(), ts=model.ts.numpy(), out=out.numpy(),
            alpha=canonical.alpha.numpy(), weights=canonical.weights.numpy(), rigid_dp=model.rigid_dp.numpy())
  np.savez_compressed(os.path.join(HERE, name + ".npz"), **fx)
  print(name, "out", out.shape, "mean", float(out.mean()), "max|rigid_dp|", float(model.rigid_dp.abs().max()))

def case_plain_grads(name, seed, B, H, W, T, top=0, left=0, refl_kind="view"):
  """Gradients of the reference itself: loss = mse(model(rays), target), loss.backward() through the reference's own modules
  (runner.py:600-602,820).  The parity target of the fused backward (SURVEY f-1).  refl_kind "pos": the makefile's main training
  configuration (makefile:12), whose head has its own hash encoder (refl.py:233-237)."""
  params = O.make_plain_params(seed, 64, 20.0, refl_kind=refl_kind)
  rays = O.make_rays(B, H, W, size=800, seed=seed, crop_top=top, crop_left=left)
  model, args = ref_plain(params, T, train=False, refl_kind=refl_kind)
  for p in model.parameters(): p.requires_grad_(p.dtype.is_floating_point and p.numel() > 0)
  g = np.random.default_rng(seed)
  target = torch.from_numpy(g.uniform(0, 1, size=(B, H, W, 3)).astype(np.float32))
  out = model(rays)
  loss = torch.nn.functional.mse_loss(out, target)
  loss.backward()
  fx = dict(kind="plain_grads", refl_kind=refl_kind, seed=seed, B=B, H=H, W=W, T=T, top=top, left=left, near=float(args.near), far=float(args.far),
            sigmoid=args.sigmoid_kind, bg=args.bg, target=target.numpy(), out=out.detach().numpy(), loss=float(loss))
  sd = dict(model.named_parameters())
  names = ("first.init.weight", "first.layers.0.weight", "first.layers.2.bias", "first.out.weight", "first.out.bias",
           "refl.mlp.init.weight", "refl.mlp.layers.0.weight", "refl.mlp.layers.3.weight", "refl.mlp.out.weight", "refl.mlp.out.bias")
  if refl_kind == "pos": names += ("refl.mlp.layers.4.weight", "refl.mlp.init.bias")
  for n in names:
    gr = sd[n].grad.numpy()
    fx["grad." + n] = gr[::16] if gr.ndim == 2 and gr.shape[0] == 256 else gr       # every 16th row of the 256-row matrices keeps the fixture small
  tables = [("emb", "first.enc.embs")] + ([("remb", "refl.mlp.enc.embs")] if refl_kind == "pos" else [])
  for tag, pre in tables:
    for lvl in (0, 7):                                   # hash tables: sparse rows
      gt = sd[f"{pre}.{lvl}.weight"].grad
      rows = torch.nonzero(gt.abs().sum(1)).squeeze(1)
      fx[f"grad.{tag}{lvl}.rows"] = rows.numpy().astype(np.int32); fx[f"grad.{tag}{lvl}.vals"] = gt[rows].numpy()
  np.savez_compressed(os.path.join(HERE, name + ".npz"), **fx)
  print(name, "loss", float(loss), "rows", len(fx["grad.emb0.rows"]), len(fx["grad.emb7.rows"]))

def r_o_pts(rays, ts):
  r_o, r_d = rays.split([3, 3], dim=-1)
  pts = r_o.unsqueeze(0) + torch.tensordot(ts, r_d, dims=0)
  return pts, ts, r_o, r_d, None

def check_rays():
  runner, nerf, refl, utils, cameras = ref_shim.load()
  import math
  size = 800
  rays = O.make_rays(2, 5, 7, size=size, seed=3, crop_top=11, crop_left=400)
  # rebuild the same poses and go through the reference camera + runner.render's pixel grid
  g = np.random.default_rng(3)
  focal = 0.5 * size / math.tan(0.5 * 0.6911112)
  c2w = []
  for _ in range(2):
    th, ph = g.uniform(0, 2 * math.pi), g.uniform(0.15, 1.2)
    eye = np.array([4 * math.cos(th) * math.cos(ph), 4 * math.sin(th) * math.cos(ph), 4 * math.sin(ph)])
    fwd = -eye / np.linalg.norm(eye)
    right = np.cross(fwd, np.array([0.0, 0.0, 1.0])); right /= np.linalg.norm(right)
    up = np.cross(right, fwd)
    m = np.eye(4); m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = right, up, -fwd, eye
    c2w.append(m[:3, :4])
  cam = cameras.NeRFCamera(cam_to_world=torch.from_numpy(np.stack(c2w).astype(np.float32)), focal=focal)
  ii, jj = torch.meshgrid(torch.arange(size, dtype=torch.float), torch.arange(size, dtype=torch.float), indexing="ij")
  positions = torch.stack([ii.transpose(-1, -2), jj.transpose(-1, -2)], dim=-1)[11:16, 400:407, :]
  ref = cam.sample_positions(positions, size=size, with_noise=False)
  assert torch.equal(ref, rays), "make_rays != reference camera"
  print("make_rays == NeRFCamera.sample_positions: ok")

def case_dtu_rays(name, seed=5, B=2, size=400, top=37, left=101, H=6, W=9):
  """DTUCamera.sample_positions (reference src/cameras.py:189-223) on runner.render's pixel grid: IDR-style poses / intrinsics
  (focal ~ 2900 px of the 1600 x 1200 original, principal point near the centre, a little skew), unit-norm r_d."""
  runner, nerf, refl, utils, cameras = ref_shim.load()
  g = np.random.default_rng(seed)
  pose = np.zeros((B, 4, 4), np.float32); intr = np.zeros((B, 4, 4), np.float32)
  for b in range(B):
    q, _ = np.linalg.qr(g.normal(size=(3, 3)))
    if np.linalg.det(q) < 0: q[:, 0] = -q[:, 0]
    pose[b, :3, :3] = q; pose[b, :3, 3] = g.normal(size=3) * 1.5; pose[b, 3, 3] = 1
    intr[b] = np.eye(4); intr[b, 0, 0] = 2892.3 + 10 * g.normal(); intr[b, 1, 1] = 2883.2 + 10 * g.normal()
    intr[b, 0, 2] = 823.2 + 5 * g.normal(); intr[b, 1, 2] = 619.1 + 5 * g.normal(); intr[b, 0, 1] = 0.3 * g.normal()
  cam = cameras.DTUCamera(pose=torch.from_numpy(pose), intrinsic=torch.from_numpy(intr))
  ii, jj = torch.meshgrid(torch.arange(size, dtype=torch.float), torch.arange(size, dtype=torch.float), indexing="ij")
  positions = torch.stack([ii.transpose(-1, -2), jj.transpose(-1, -2)], dim=-1)[top:top + H, left:left + W, :]     # runner.py:495-503
  rays = cam.sample_positions(positions, size=size, with_noise=False)
  assert rays.shape == (B, H, W, 6)
  np.savez_compressed(os.path.join(HERE, name + ".npz"), pose=pose, intrinsic=intr, size=size, top=top, left=left, H=H, W=W, rays=rays.numpy())
  print(name, rays.shape, "|r_d| =", float(rays[..., 3:].norm(dim=-1).mean()))

def analytic_scene(rays):
  """Procedural target (SURVEY 8d weight set T; there is no dataset in the container): a unit sphere at the origin, shaded
  with a position-dependent albedo and a fixed light, on black -- intersected analytically per ray."""
  o, d = rays[..., :3], torch.nn.functional.normalize(rays[..., 3:], dim=-1)
  b = (o * d).sum(-1); c = (o * o).sum(-1) - 1.0
  disc = b * b - c
  hit = disc > 0
  t = -b - disc.clamp(min=0).sqrt()
  p = o + t[..., None] * d
  n = torch.nn.functional.normalize(p, dim=-1)
  light = torch.nn.functional.normalize(torch.tensor([0.4, -0.3, 0.85]), dim=0)
  albedo = 0.5 + 0.5 * torch.sin(3.0 * p + torch.tensor([0.0, 2.0, 4.0]))
  col = albedo * (0.25 + 0.75 * (n * light).sum(-1, keepdim=True).clamp(min=0))
  return torch.where(hit[..., None], col, torch.zeros_like(col))

def case_trained(name, seed=1337, iters=400, T=64, crop=24, views=6, lr=2e-3):
  """Weight set T: the REFERENCE model trained by the reference's own forward / autograd / Adam (runner.py:448-458:
  Adam, eps 1e-7; mse loss 600-602; training-mode jitter + density noise) on the procedural scene, then rendered in eval mode.
  The fixture stores the trained parameters as deltas to the seeded initialisation (MLP weights dense, hash tables as the
  touched rows) plus the reference's eval render and the largest pre-activation magnitude it saw."""
  runner, nerf, refl, utils, cameras = ref_shim.load()
  params = O.make_plain_params(seed, 64, 1.0)
  model, args = ref_plain(params, T, train=True)
  for p in model.parameters(): p.requires_grad_(p.dtype.is_floating_point and p.numel() > 0)
  opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=lr, eps=1e-7)
  sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=iters, eta_min=lr / 20)         # runner.py:1289
  g = np.random.default_rng(seed)
  torch.manual_seed(seed)
  size = 800
  for it in range(iters):
    v = int(g.integers(0, views)); top = int(g.integers(250, 550 - crop)); left = int(g.integers(250, 550 - crop))
    rays = O.make_rays(views, crop, crop, size=size, seed=seed, crop_top=top, crop_left=left)[v:v + 1]
    opt.zero_grad()
    loss = torch.nn.functional.mse_loss(model(rays), analytic_scene(rays))
    loss.backward(); opt.step(); sched.step()
    if it % 50 == 0 or it == iters - 1: print(name, "iter", it, "loss", float(loss.detach()), flush=True)
  model.eval()
  # eval render of a held-out crop through the reference, recording the largest |pre-activation| (fp16 operand range check)
  H = W = 16; top, left = 392, 392
  rays = O.make_rays(views, H, W, size=size, seed=seed, crop_top=top, crop_left=left)[:2]
  hmax = [0.0]
  def hook(mod, inp, out): hmax[0] = max(hmax[0], float(out.detach().abs().max()))
  hs = [m.register_forward_hook(hook) for m in model.modules() if isinstance(m, torch.nn.Linear)]
  with torch.no_grad(): out = model(rays)
  for h in hs: h.remove()
  sd = {k: v.detach() for k, v in model.state_dict().items()}
  fx = dict(kind="plain_trained", seed=seed, iters=iters, T=T, B=2, H=H, W=W, top=top, left=left, views=views, near=float(args.near), far=float(args.far),
            sigmoid=args.sigmoid_kind, bg=args.bg, ts=model.ts.numpy(), out=out.numpy(), alpha=model.alpha.numpy(), weights=model.weights.numpy(),
            final_loss=float(loss), max_abs_preactivation=hmax[0], target=analytic_scene(rays).numpy())
  for k, v in sd.items():
    if not v.dtype.is_floating_point or v.numel() == 0: continue
    if ".embs." in k:
      rows = torch.nonzero((v != params[k]).any(dim=1)).squeeze(1)
      fx["rows." + k] = rows.numpy().astype(np.int32); fx["vals." + k] = v[rows].numpy()
    else: fx["param." + k] = v.numpy()
  np.savez_compressed(os.path.join(HERE, name + ".npz"), **fx)
  print(name, "final loss", float(loss), "max|h|", hmax[0], "psnr vs target", float(-10 * torch.log10(torch.nn.functional.mse_loss(out, analytic_scene(rays)))),
        "table rows stored", sum(len(fx[k]) for k in fx if k.startswith("rows.")))

def case_sdf(name, seed=23, sdf_kind="siren", size=16, iters_fit=500, bound_rad=-1.0, isect="sphere"):
  """The reference's surface renderer `sdf.SDF` (src/sdf.py:86-156) with the sphere-march intersection (src/march.py:27-47) and
  a View head, eval mode (192 iterations).  The SDF network is first fitted to a sphere of radius 1 with the reference's own
  modules (as `SDFModel.set_to_sphere` does, src/sdf.py:47-60), then ALL parameters are rounded to fp16 so that the fixture
  stays small; the reference renders with exactly those rounded values."""
  runner, nerf, refl, utils, cameras = ref_shim.load()
  import src.sdf as rsdf, src.march as march
  torch.manual_seed(seed)
  model = rsdf.sdf_kinds[sdf_kind](intermediate_size=64)
  opt = torch.optim.Adam(model.parameters(), lr=2e-4 if sdf_kind == "siren" else 1e-3, weight_decay=0)
  for i in range(iters_fit):
    opt.zero_grad()
    v = 2 * torch.randn(4000, 3)
    loss = torch.nn.functional.mse_loss(model(v)[..., 0], v.norm(dim=-1) - 1.0)
    loss.backward(); opt.step()
    if i % 100 == 0 or i == iters_fit - 1: print(name, "fit", i, float(loss.detach()), flush=True)
  if bound_rad > 0: wrapped = rsdf.UnitSphere(inner=model, rad=bound_rad)
  else: wrapped = model
  # shim: SDF.forward passes mask=hit to the head (sdf.py:152) but View.forward takes no such keyword (refl.py:205): swallow it
  orig_fwd = refl.View.forward
  refl.View.forward = lambda self, x, view, normal=None, light=None, latent=None, mask=None: orig_fwd(self, x, view, normal, light, latent)
  head = refl.View(latent_size=64, act="upshifted", out_features=3)
  s = rsdf.SDF(wrapped, head, isect=march.load_intersection_kind(isect), t_near=2.0, t_far=6.0).eval()
  with torch.no_grad():
    for p in s.parameters():
      if p.dtype.is_floating_point: p.copy_(p.to(torch.float16).to(torch.float32))
  rays = O.make_rays(2, size, size, size=size, seed=seed)
  rays = torch.cat([rays[..., :3], torch.nn.functional.normalize(rays[..., 3:], dim=-1)], dim=-1)
  if isect in ("bisect", "secant"):
    # march.bisect (src/march.py:63-75) / march.secant (50-60) = throughput_with_sign_change (78-110; it draws random.random() once,
    # march.py:86: the draw is pinned with random.seed and stored) + bisection (147-180) / secant_find (113-143)
    import random
    random.seed(seed); jitter = random.random()
    with torch.no_grad():
      random.seed(seed); out = s(rays)
      random.seed(seed); pts, hit, best_pos, tput = march.load_intersection_kind(isect)(s.underlying, rays[..., :3], rays[..., 3:], iters=192, near=2.0, far=6.0)
      random.seed(seed); _, _, last_pos, first_neg = march.throughput_with_sign_change(s.underlying, rays[..., :3], rays[..., 3:], near=2.0, far=6.0, batch_size=192)
    fx = dict(kind="sdf", sdf_kind=sdf_kind, seed=seed, size=size, near=2.0, far=6.0, iters=192, sigmoid="upshifted", bound_rad=bound_rad, isect=isect,
              jitter=jitter, rays=rays.numpy(), out=out.numpy(), hit=hit.numpy(), pts=pts.numpy(), best_pos=best_pos.numpy(), tput=tput.reshape(hit.shape).numpy(),
              last_pos=last_pos.squeeze(-1).numpy(), first_neg=first_neg.squeeze(-1).numpy())
    t = tput
  else:
   with torch.no_grad():
    out = s(rays)
    pts, hit, t, _ = march.sphere_march(s.underlying, rays[..., :3], rays[..., 3:], iters=192, near=2.0, far=6.0)
   fx = dict(kind="sdf", sdf_kind=sdf_kind, seed=seed, size=size, near=2.0, far=6.0, iters=192, sigmoid="upshifted", bound_rad=bound_rad,
            rays=rays.numpy(), out=out.numpy(), hit=hit.numpy(), t=t.squeeze(-1).numpy(), pts=pts.numpy())
  if isect == "sphere":
    for k, v in s.state_dict().items():
      if v.dtype.is_floating_point and v.numel(): fx["param16." + k.replace("underlying.inner.", "underlying.")] = v.to(torch.float16).numpy()
  else:
    # same seed, same fit: the parameters are those of the sphere-march golden (checked here), stored once
    ref = np.load(os.path.join(HERE, "sdf_siren_march.npz"))
    for k, v in s.state_dict().items():
      if v.dtype.is_floating_point and v.numel(): assert np.array_equal(ref["param16." + k.replace("underlying.inner.", "underlying.")], v.to(torch.float16).numpy()), k
    fx["params_from"] = "sdf_siren_march"
  np.savez_compressed(os.path.join(HERE, name + ".npz"), **fx)
  print(name, "hit fraction", float(hit.float().mean()), "t range", float(t.min()), float(t.max()), "out mean", float(out.mean()))

def case_sdf_normals(name="sdf_siren_normals"):
  """SDFModel.normals (src/sdf.py:43-49, utils.autograd 266-277) of the reference's own SIREN SDF module holding the sphere-march
  golden's parameters: at that golden's march points and at random points, bare and inside a UnitSphere (sdf.py:66-83; rad 1.0 = the fitted radius, so both branches of the max occur)."""
  runner, nerf, refl, utils, cameras = ref_shim.load()
  import src.sdf as rsdf
  ref = np.load(os.path.join(HERE, "sdf_siren_march.npz"))
  model = rsdf.sdf_kinds["siren"](intermediate_size=64)
  sd = {k[len("param16.underlying."):]: torch.from_numpy(ref[k].astype(np.float32)) for k in ref.files if k.startswith("param16.underlying.")}
  model.load_state_dict(sd, strict=True)
  g = torch.Generator().manual_seed(77)
  pts = torch.cat([torch.from_numpy(ref["pts"]).reshape(-1, 3)[:300], torch.randn(213, 3, generator=g) * 1.2])
  n_bare = model.normals(pts.clone()).detach()
  n_unit = rsdf.UnitSphere(inner=model, rad=1.0).normals(pts.clone()).detach()
  with torch.no_grad(): vals = model(pts)
  np.savez_compressed(os.path.join(HERE, name + ".npz"), kind="sdf_normals", sdf_kind="siren", params_from="sdf_siren_march", bound_rad=1.0,
                      pts=pts.numpy(), normals=n_bare.numpy(), normals_unit=n_unit.numpy(), values=vals.numpy())
  print(name, "points", pts.shape[0], "|n| mean", float(n_bare.norm(dim=-1).mean()), "unit-sphere branch taken", float((pts.norm(dim=-1) - 1.0 > vals[:, 0]).float().mean()))

if __name__ == "__main__":
  check_rays()
  if "--sdf-normals" in sys.argv:
    case_sdf_normals()
    sys.exit(0)
  if "--sdf" in sys.argv:
    case_sdf("sdf_siren_march")
    sys.exit(0)
  if "--sdf-bisect" in sys.argv:
    case_sdf("sdf_siren_bisect", size=12, isect="bisect")
    sys.exit(0)
  if "--trained" in sys.argv:
    torch.set_num_threads(int(os.environ.get("GOLDEN_THREADS", "8")))
    case_trained(os.environ.get("GOLDEN_NAME", "plain_trained_t64"), iters=int(os.environ.get("GOLDEN_ITERS", "400")), lr=float(os.environ.get("GOLDEN_LR", "2e-3")))
    sys.exit(0)
  if "--dtu" in sys.argv:
    case_dtu_rays("dtu_rays")
    sys.exit(0)
  if "--volsdf-grads" in sys.argv:
    case_volsdf_grads("volsdf_siren_t32_grads", seed=33, B=1, H=3, W=4, T=32, top=398, left=397)
    sys.exit(0)
  if "--pos-grads" in sys.argv:
    case_plain_grads("plain_pos_t32_grads", seed=93, B=1, H=3, W=4, T=32, top=398, left=397, refl_kind="pos")
    sys.exit(0)
  if "--grads" in sys.argv:
    case_plain_grads("plain_t16_grads", seed=91, B=1, H=3, W=4, T=16, top=398, left=397)
    sys.exit(0)
  if "--poslin" in sys.argv:
    case_plain("plain_poslinview_t16", seed=83, B=1, H=4, W=5, T=16, sigma_gain=20.0, top=398, left=397, stages=False, refl_kind="pos-linear-view")
    sys.exit(0)
  if "--pos" in sys.argv:
    case_plain("plain_pos_t16", seed=81, B=1, H=4, W=5, T=16, sigma_gain=20.0, top=398, left=397, stages=False, refl_kind="pos")
    sys.exit(0)
  if "--new" in sys.argv:      # only the cases added after the first goldens were committed
    case_mip("plain_mip_cylinder_t16", seed=61, B=2, H=5, W=4, T=16, top=398, left=397)
    case_dnerf_spline("dnerf_spline5_t32", seed=71, n=5, B=2, H=3, W=4, T=32, top=398, left=397)
    case_dnerf_spline("dnerf_spline4_t32", seed=72, n=4, B=2, H=3, W=3, T=32, top=398, left=397)
    sys.exit(0)
  case_plain("plain_t16", seed=11, B=1, H=4, W=6, T=16)
  case_plain("plain_t16_sharp", seed=12, B=2, H=3, W=5, T=16, sigma_gain=20.0, top=390, left=380)
  case_plain("plain_t128", seed=1337, B=1, H=8, W=8, T=128, sigma_gain=20.0, top=396, left=396, stages=False)
  case_volsdf("volsdf_siren_t32", seed=31, sdf_kind="siren", B=1, H=4, W=5, T=32, top=398, left=397)
  case_volsdf("volsdf_mlp_t32", seed=32, sdf_kind="mlp", B=1, H=3, W=4, T=32, top=398, left=397)
  case_dnerf("dnerf_direct_t64", seed=51, B=2, H=3, W=4, T=64, top=398, left=397)
  case_plain("plain_t64_train", seed=21, B=1, H=4, W=4, T=64, sigma_gain=20.0, train=True, top=300, left=420, stages=False)
  case_mip("plain_mip_cylinder_t16", seed=61, B=2, H=5, W=4, T=16, top=398, left=397)
  case_dnerf_spline("dnerf_spline5_t32", seed=71, n=5, B=2, H=3, W=4, T=32, top=398, left=397)
  case_dnerf_spline("dnerf_spline4_t32", seed=72, n=4, B=2, H=3, W=3, T=32, top=398, left=397)
  case_plain("plain_pos_t16", seed=81, B=1, H=4, W=5, T=16, sigma_gain=20.0, top=398, left=397, stages=False, refl_kind="pos")
  case_plain_grads("plain_t16_grads", seed=91, B=1, H=3, W=4, T=16, top=398, left=397)
  case_dtu_rays("dtu_rays")
  case_trained("plain_trained_t64")
  case_sdf("sdf_siren_march")
  case_sdf("sdf_siren_bisect", size=12, isect="bisect")
  case_sdf_normals()
  case_plain("plain_poslinview_t16", seed=83, B=1, H=4, W=5, T=16, sigma_gain=20.0, top=398, left=397, stages=False, refl_kind="pos-linear-view")

timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python - <<'PY'
import torch, sys
sys.path.insert(0, '.')
import nerf_atlas_b200 as N
from oracle import nerf_oracle as O
from nerf_atlas_b200.autograd import FusedAdam
dev = 'cuda:0'
for kind in ("view", "pos"):
  P = O.make_plain_params(7, 64, 1.0, refl_kind=kind)
  m = N.FusedPlainNeRF(steps=128, t_near=2, t_far=6, intermediate_size=64, sigmoid_kind='upshifted', precision='fp16', refl_kind=kind)
  m.load_state_dict(P, strict=True); m = m.to(dev).train(); m.keep_weights = False
  opt = FusedAdam([p for p in m.parameters() if p.requires_grad and p.numel() > 0], lr=5e-4)
  rays = O.make_rays(1, 64, 64, seed=3, crop_top=368, crop_left=368).to(dev)
  tgt = torch.rand(1, 64, 64, 3, device=dev)
  def step():
    opt.zero_grad(set_to_none=False)
    loss = torch.nn.functional.mse_loss(m(rays), tgt); loss.backward(); opt.step(); return loss
  for _ in range(5): step()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(30): l = step()
  e1.record(); torch.cuda.synchronize()
  print(kind, 'train step ms', e0.elapsed_time(e1) / 30, 'loss', float(l.detach()))
PY

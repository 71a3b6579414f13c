"""Autograd surface of the fused path.  The reference differentiates `model(rays)` implicitly through PyTorch
autograd (`loss.backward()`, runner.py:820); here

* ``fused_render`` is ONE differentiable op for the whole path -- training forward with an activation stash
  (`nf_render_forward_aux`), backward of composite, both MLPs (tcgen05: dX chain and dW GEMMs) and the hash tables
  (`nf_render_backward`) -- with gradients for every parameter in the reference's layout; `FusedNeRF.forward` uses it whenever
  grad is enabled;
* ``composite`` / ``hash_encode`` are the differentiable stand-alone stages (`nf_composite[_backward]`,
  `nf_hash_encode[_backward]`);
* ``FusedAdam`` is torch.optim.Adam with the reference's hyper-parameters on `nf_adam_step`."""
from __future__ import annotations
import torch
from .model import RenderEngine


class _Composite(torch.autograd.Function):
  @staticmethod
  def forward(ctx, engine: RenderEngine, sigma_raw, feats, rays, ts):
    if engine.desc.density_act == 2:   # NF_DENS_LAPLACE: this stand-alone stage does not return d loss / d beta (fused_render does)
      raise NotImplementedError("autograd.composite: the Laplace density (VolSDF) has no backward for beta here; use fused_render (FusedVolSDF.forward)")
    rgb, _, _ = engine.composite(sigma_raw, feats, rays, ts, want_weights=False)
    ctx.engine = engine
    ctx.save_for_backward(sigma_raw, feats, rays, ts)
    return rgb

  @staticmethod
  def backward(ctx, d_rgb):
    sigma_raw, feats, rays, ts = ctx.saved_tensors
    d_sigma, d_feats = ctx.engine.composite_backward(sigma_raw, feats, rays, ts, d_rgb.contiguous())
    return None, d_sigma, d_feats, None, None


class _HashEncode(torch.autograd.Function):
  @staticmethod
  def forward(ctx, engine: RenderEngine, pts, *tables):
    if pts.requires_grad: raise NotImplementedError("autograd.hash_encode: no gradient with respect to pts (tables only)")
    # the encoder reads the engine's packed snapshot: refuse to run on tables that are not the ones currently packed
    packed = {k[0] for k in (engine._key or ())}
    if any(t.data_ptr() not in packed for t in tables):
      raise RuntimeError("autograd.hash_encode: `tables` are not the tensors the engine was packed from (engine.pack(...) first)")
    if any((t.data_ptr(), t._version, t.device) not in set(engine._key) for t in tables):
      raise RuntimeError("autograd.hash_encode: the tables changed since engine.pack(...): re-pack before encoding")
    feats, _ = engine.hash_encode(pts)
    ctx.engine = engine
    ctx.save_for_backward(pts)
    return feats

  @staticmethod
  def backward(ctx, d_feats):
    (pts,) = ctx.saved_tensors
    d_tables = ctx.engine.hash_encode_backward(pts, d_feats.contiguous())
    return (None, None) + tuple(d_tables[l] for l in range(d_tables.shape[0]))


class _FusedRender(torch.autograd.Function):
  """rays[R,6], ts -> rgb[R,3] (+ alpha, weights [R,T], not differentiable) with gradients to the packed parameters."""
  @staticmethod
  def forward(ctx, engine: RenderEngine, rays, ts, noise, want_weights, bg_rand, *params):
    engine.pack(params)
    R = rays.shape[0]; T = ts.shape[-1]
    lay = engine.train_layout(R, T)
    ws = RenderEngine.train_workspace(lay, rays.device)
    rgb, alpha, weights = engine.render(rays, ts, noise, want_weights=want_weights, train_ws=ws, bg_rand=bg_rand)
    ctx.engine, ctx.ws, ctx.params = engine, ws, params
    ctx.pack_key = engine._key
    ctx.save_for_backward(rays, ts)
    if not want_weights: alpha = weights = rays.new_empty(0)
    ctx.mark_non_differentiable(alpha, weights)
    return rgb, alpha, weights

  @staticmethod
  def backward(ctx, d_rgb, _d_alpha, _d_weights):
    rays, ts = ctx.saved_tensors
    if ctx.ws is None: raise RuntimeError("fused_render: backward called twice (the activation stash is freed after the first)")
    if ctx.engine._key != ctx.pack_key:
      raise RuntimeError("fused_render: the parameters were modified or re-packed between forward and backward")
    grads = [torch.empty_like(p) if need else None for p, need in zip(ctx.params, ctx.needs_input_grad[6:])]
    ctx.engine.render_backward(ctx.ws, rays, ts, d_rgb.contiguous().to(torch.float32), grads)
    ctx.ws = None                                           # several GB: free it as soon as it has been consumed
    return (None, None, None, None, None, None) + tuple(grads)


def fused_render(engine: RenderEngine, rays: torch.Tensor, ts: torch.Tensor, params, density_noise=None, want_weights: bool = True,
                 bg_rand: torch.Tensor = None):
  """The whole render path as one differentiable op: gradients flow to ``params`` (the live parameter tensors in
  `nf_pack_weights` order).  Not differentiable with respect to rays / ts / noise (camera training is out of scope).
  ``bg_rand[R]``: the draws of the "random" background (nerf.py:100-103), kept in the workspace for the backward's sky term."""
  if rays.requires_grad or ts.requires_grad: raise NotImplementedError("fused_render: gradients with respect to rays / ts are not built")
  rgb, alpha, weights = _FusedRender.apply(engine, rays, ts, density_noise, want_weights, bg_rand, *params)
  return (rgb, alpha, weights) if want_weights else (rgb, None, None)


def composite(engine: RenderEngine, sigma_raw: torch.Tensor, feats: torch.Tensor, rays: torch.Tensor, ts: torch.Tensor) -> torch.Tensor:
  """rgb[R,3] = sum_t w_t feats_t (+ sky) with gradients to sigma_raw[R,T] and feats[R,T,3] (reference src/nerf.py:60-80)."""
  return _Composite.apply(engine, sigma_raw, feats, rays, ts)


def hash_encode(engine: RenderEngine, pts: torch.Tensor, tables) -> torch.Tensor:
  """feats[N, 4 L] of HashEncoder (reference src/neural_blocks.py:139-193) with gradients to the L embedding tables
  (``tables`` = the live ``emb.weight`` tensors the engine was packed from)."""
  return _HashEncode.apply(engine, pts, *tables)


class FusedAdam(torch.optim.Optimizer):
  """torch.optim.Adam semantics (the reference's optimiser: runner.py:448-458 -- Adam, eps 1e-7, L2 weight_decay) with the
  update of ALL tensors of a parameter group done by one launch of a hand-written kernel (`nf_adam_step_multi`: 16 B read + 12 B written per element) instead of
  the ~10 element-wise launches of the eager implementation.  Learning-rate schedulers (runner.py:1289 CosineAnnealingLR) work
  unchanged: they write ``group["lr"]``."""

  def __init__(self, params, lr: float = 5e-4, betas=(0.9, 0.999), eps: float = 1e-7, weight_decay: float = 0.0):
    super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

  @torch.no_grad()
  def step(self, closure=None):
    import ctypes as C
    from . import _lib
    loss = None
    if closure is not None:
      with torch.enable_grad(): loss = closure()
    lib = _lib.lib()
    for group in self.param_groups:
      b1, b2 = group["betas"]
      # tensors of one group that share the step count and the device go to the kernel together (nf_adam_step_multi: one launch)
      buckets = {}
      for p in group["params"]:
        if p.grad is None: continue
        if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous(): raise RuntimeError("FusedAdam: contiguous fp32 CUDA parameters only")
        st = self.state[p]
        if not st:
          st["step"] = 0; st["exp_avg"] = torch.zeros_like(p); st["exp_avg_sq"] = torch.zeros_like(p)
        st["step"] += 1
        buckets.setdefault((p.device, int(st["step"])), []).append((p, p.grad.contiguous(), st))
      for (dev, step), items in buckets.items():
        n = len(items)
        arr = lambda xs: (C.c_void_p * n)(*[x.data_ptr() for x in xs])
        numel = (C.c_int64 * n)(*[p.numel() for p, _, _ in items])
        with torch.cuda.device(dev):
          rc = lib.nf_adam_step_multi(n, arr([p for p, _, _ in items]), arr([g for _, g, _ in items]), arr([s["exp_avg"] for _, _, s in items]),
                                      arr([s["exp_avg_sq"] for _, _, s in items]), numel, float(group["lr"]), float(b1), float(b2),
                                      float(group["eps"]), float(group["weight_decay"]), step,
                                      C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, "nf_adam_step_multi")
        # the kernel wrote through raw pointers: tell autograd (saved-tensor checks) and RenderEngine.pack (whose cache key is
        # (data_ptr, _version)) that the parameters changed, exactly as an in-place torch op would
        for p, _, _ in items: torch.autograd.graph.increment_version(p)
    return loss

"""Differentiable wrappers of the two non-GEMM stages (first blocks): the hash-grid encoder and the alpha
composite, forward AND backward in hand-written CUDA through the C ABI (`nf_hash_encode[_backward]`, `nf_composite[_backward]`).
The reference differentiates the same ops implicitly through PyTorch autograd (runner.py:820).  The fused backward of the MLP
chain is not built yet, so `FusedNeRF.forward` still refuses to run with grad enabled in training mode."""
from __future__ import annotations
import torch
from .model import RenderEngine


class _Composite(torch.autograd.Function):
  @staticmethod
  def forward(ctx, engine: RenderEngine, sigma_raw, feats, rays, ts):
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
    feats, _ = engine.hash_encode(pts)          # reads the engine's packed snapshot of `tables` (engine.pack must be current)
    ctx.engine = engine
    ctx.save_for_backward(pts)
    return feats

  @staticmethod
  def backward(ctx, d_feats):
    (pts,) = ctx.saved_tensors
    d_tables = ctx.engine.hash_encode_backward(pts, d_feats.contiguous())
    return (None, None) + tuple(d_tables[l] for l in range(d_tables.shape[0]))


def composite(engine: RenderEngine, sigma_raw: torch.Tensor, feats: torch.Tensor, rays: torch.Tensor, ts: torch.Tensor) -> torch.Tensor:
  """rgb[R,3] = sum_t w_t feats_t (+ sky) with gradients to sigma_raw[R,T] and feats[R,T,3] (reference src/nerf.py:60-80)."""
  return _Composite.apply(engine, sigma_raw, feats, rays, ts)


def hash_encode(engine: RenderEngine, pts: torch.Tensor, tables) -> torch.Tensor:
  """feats[N, 4 L] of HashEncoder (reference src/neural_blocks.py:139-193) with gradients to the L embedding tables
  (``tables`` = the live ``emb.weight`` tensors the engine was packed from)."""
  return _HashEncode.apply(engine, pts, *tables)


class FusedAdam(torch.optim.Optimizer):
  """torch.optim.Adam semantics (the reference's optimiser: runner.py:448-458 -- Adam, eps 1e-7, L2 weight_decay) with the
  update of every tensor done by one hand-written kernel (`nf_adam_step`: 16 B read + 12 B written per element) instead of
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
      for p in group["params"]:
        if p.grad is None: continue
        if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous(): raise RuntimeError("FusedAdam: contiguous fp32 CUDA parameters only")
        st = self.state[p]
        if not st:
          st["step"] = 0; st["exp_avg"] = torch.zeros_like(p); st["exp_avg_sq"] = torch.zeros_like(p)
        st["step"] += 1
        g = p.grad.contiguous()
        with torch.cuda.device(p.device):
          rc = lib.nf_adam_step(C.c_void_p(p.data_ptr()), C.c_void_p(g.data_ptr()), C.c_void_p(st["exp_avg"].data_ptr()),
                                C.c_void_p(st["exp_avg_sq"].data_ptr()), p.numel(), float(group["lr"]), float(b1), float(b2),
                                float(group["eps"]), float(group["weight_decay"]), int(st["step"]),
                                C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, "nf_adam_step")
        # the kernel wrote through the raw pointer: tell autograd (saved-tensor checks) and RenderEngine.pack (whose cache key
        # is (data_ptr, _version)) that the parameter changed, exactly as an in-place torch op would
        torch.autograd.graph.increment_version(p)
    return loss

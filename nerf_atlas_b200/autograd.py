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

"""Ray sharding across the GPUs of one box.

Rays are independent, so the render path shards with NO data-path collective: rank g renders the
contiguous block ``[g*R/G, (g+1)*R/G)`` of the flattened ``[B,H,W]`` ray grid with replicated
parameters.  The only (optional) communication is an all-gather of the 12 B/ray RGB result when
one rank needs the whole frame.  The reference's counterpart is ``nn.DataParallel`` scatter/gather
along dim 0 (reference runner.py:1207-1209), which is single-process and broken at HEAD.
One process per GPU; ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) is plumbing only.
"""
from __future__ import annotations
from typing import Callable, Optional, Tuple
import torch
import torch.distributed as dist


def shard_rays(n_rays: int, rank: int, world: int, align: int = 1) -> Tuple[int, int]:
  """Contiguous block of rank ``rank``; boundaries are multiples of ``align`` (e.g. the image width,
  so that shards are whole pixel rows).  Blocks differ by at most ``align`` rays and cover [0, n_rays)."""
  if not (0 <= rank < world): raise ValueError("rank out of range")
  if align < 1: raise ValueError("align must be >= 1")
  units = (n_rays + align - 1) // align
  base, rem = divmod(units, world)
  u0 = rank * base + min(rank, rem)
  u1 = u0 + base + (1 if rank < rem else 0)
  return min(u0 * align, n_rays), min(u1 * align, n_rays)


class ShardedRenderer:
  """Wraps ``render_fn(rays[r,6]) -> rgb[r,C]`` so that every rank renders only its block.

  ``gather=True`` returns the full ``[R,C]`` result on every rank (all-gather of padded blocks);
  ``gather=False`` returns the local block and its ``(start, end)``.
  """
  def __init__(self, render_fn: Callable[[torch.Tensor], torch.Tensor], group: Optional[dist.ProcessGroup] = None,
               gather: bool = True, align: int = 1):
    self.render_fn, self.group, self.gather, self.align = render_fn, group, gather, align

  def _world(self):
    if dist.is_available() and dist.is_initialized():
      return dist.get_rank(self.group), dist.get_world_size(self.group)
    return 0, 1

  def __call__(self, rays: torch.Tensor):
    rank, world = self._world()
    R = rays.shape[0]
    s, e = shard_rays(R, rank, world, self.align)
    local = self.render_fn(rays[s:e].contiguous())
    if not self.gather: return local, (s, e)
    if world == 1: return local
    bounds = [shard_rays(R, r, world, self.align) for r in range(world)]
    longest = max(b[1] - b[0] for b in bounds)
    pad = torch.zeros(longest, local.shape[1], dtype=local.dtype, device=local.device)
    pad[: e - s] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=self.group)
    return torch.cat([p[: b[1] - b[0]] for p, b in zip(parts, bounds)], dim=0)


class GradientAllReducer:
  """The one collective of data-parallel training: every rank renders its block of the step's rays,
  back-propagates the SUM of its per-ray losses, and the flattened fp32 gradient is all-reduced (sum) once per optimiser step
  and divided by the global ray count, so that every rank applies the gradient of the mean loss over ALL rays -- what the
  single-process reference computes (loss = mean over the crop, runner.py:600-602, 820-824).

  One flat fp32 bucket (2.7 M parameters for Plain+View = 10.8 MB: a single NCCL all-reduce over NVLink; NVSwitch makes the cost
  latency-, not link-bound, so there is nothing to gain from splitting it).  ``begin()`` launches the all-reduce asynchronously
  (on NCCL: on the communicator's own stream, overlapping whatever the caller still has to run, e.g. the tail of backward of
  another parameter group); ``finish()`` waits and scatters the averaged gradient back into ``p.grad``.
  Host-side plumbing only: torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""

  def __init__(self, params, group: Optional[dist.ProcessGroup] = None):
    self.params = [p for p in params if p.requires_grad]
    self.group = group
    self._flat: Optional[torch.Tensor] = None
    self._work = None
    self._count: Optional[torch.Tensor] = None

  def _world(self) -> int:
    return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

  def _layout(self):
    """Slots padded to 4 floats so that every gradient view stays 16-byte aligned (nf_adam_step's requirement)."""
    offs, off = [], 0
    for p in self.params:
      offs.append(off); off += (p.numel() + 3) // 4 * 4
    return offs, off

  def begin(self, n_local_rays: int, n_global_rays: Optional[int] = None):
    """Call after backward of the local SUM loss.  n_local_rays = rays this rank contributed to the step; n_global_rays (optional)
    = the step's total over all ranks when the caller knows it (deterministic shards): ``finish`` then needs no host read-back."""
    ps = self.params
    if not ps: return
    dev = ps[0].device
    offs, total = self._layout()
    if self._flat is None or self._flat.numel() != total + 4 or self._flat.device != dev:
      self._flat = torch.zeros(total + 4, dtype=torch.float32, device=dev)      # [g_0 | g_1 | ... | ray count]
      self._views = [self._flat[o:o + p.numel()].view(p.shape) for o, p in zip(offs, ps)]
    have = [(v, p.grad) for v, p in zip(self._views, ps) if p.grad is not None and p.grad.data_ptr() != v.data_ptr()]
    if have: torch._foreach_copy_([v for v, _ in have], [g for _, g in have])   # batched copy into the bucket
    for v, p in zip(self._views, ps):
      if p.grad is None: v.zero_()
    self._flat[total:].fill_(float(n_local_rays))
    self._n_global, self._total = n_global_rays, total
    if self._world() > 1: self._work = dist.all_reduce(self._flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

  def finish(self) -> int:
    """Waits for the all-reduce; p.grad <- (sum over ranks of local grads) / (global ray count).  Returns that count.
    The gradients become VIEWS of the flat bucket (no copy back)."""
    ps = self.params
    if not ps or self._flat is None: return 0
    if self._work is not None: self._work.wait(); self._work = None
    total = self._total
    n_rays = float(self._n_global) if self._n_global is not None else float(self._flat[total].item())
    self._flat[:total].mul_(1.0 / max(n_rays, 1.0))
    for v, p in zip(self._views, ps): p.grad = v
    return int(n_rays)

"""Host-side mirror of the reference's NeRF forward surface over the C ABI.

``FusedPlainNeRF`` / ``FusedTinyNeRF`` are ``nn.Module``s that ``runner.py`` can use in place of
``src.nerf.PlainNeRF`` / ``TinyNeRF`` (reference src/nerf.py:278-361): same constructor keywords,
``forward(rays[B,H,W,6]) -> rgb[B,H,W,3]``, the inner entry ``from_pts(pts, ts, r_o, r_d)``, autograd to the
parameters in training mode (``loss.backward()`` runs the native backward), and the attributes the runner reads or writes afterwards
(SURVEY.md section 8b): ``steps/t_near/t_far``, ``ts``, ``alpha``, ``weights``, ``nerf``, ``refl``,
``intermediate_size``, ``set_bg``, ``set_sigmoid``, ``set_refl``, ``total_latent_size``.
Parameters keep the reference's ``state_dict`` names, so checkpoints interchange.

The module owns ordinary ``nn.Parameter``s; the CUDA side only ever sees a packed snapshot that
is refreshed when a parameter's ``data_ptr``/``_version`` changes (``RenderEngine.pack``).
There is no PyTorch or CPU fallback: without the built library or a CUDA device, calls raise.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from ._lib import ModelDesc, MlpDesc, MipArgs, RenderAux, TrainLayout

HASH_PRIMES = [1, 2654435761, 805459861, 3674653429, 2097192037, 1434869437, 2165219737]


# ------------------------------------------------------------------------------------------------
# descriptors
# ------------------------------------------------------------------------------------------------
def _hash_res(levels: int, low: int = 16, high: int = 1 << 14) -> List[float]:
  # operator precedence of reference src/neural_blocks.py:126-128: exp((ln hi - ln lo)/levels - 1)
  scale = math.exp((math.log(high) - math.log(low)) / levels - 1)
  return [float(np.float32(low * (scale ** i))) for i in range(levels)]


def _mlp(in_dims, n_layers, out_dims, act, skip=3, hidden=256) -> MlpDesc:
  return MlpDesc(in_dims, hidden, n_layers, out_dims, skip, _lib.ACT[act])


MIP_FEATS = 96   # 2 x 16 degrees x 3 axes (reference src/utils.py:104-111, src/nerf.py:255)


def describe_plain(intermediate: int = 64, sigmoid: str = "upshifted", bg: str = "black",
                   hash_levels: int = 8, hash_table: int = 1 << 16, mip: Optional[str] = None, refl_kind: str = "view",
                   poslin_im: int = 64) -> ModelDesc:
  """PlainNeRF + View head as built by runner.load_model (reference src/nerf.py:310-324,
  src/refl.py:190-204, runner.py:1182-1183).  ``mip`` in (None, "cylinder", "cone", "cylinder_ref"): the IPE latent
  widens both MLP inputs by 96 (nerf.py:255,311-324)."""
  d = ModelDesc()
  d.struct_bytes = C.sizeof(ModelDesc)
  d.kind = _lib.KIND["plain"]
  d.mip = _lib.MIP[mip]
  ml = MIP_FEATS if d.mip else 0
  d.density = _mlp(6 + 4 * hash_levels + ml, 4, 1 + intermediate, "leaky_relu")
  d.refl_kind = _lib.REFL[refl_kind]
  if refl_kind == "pos": d.refl = _mlp(6 + 4 * hash_levels + ml + intermediate, 5, 3, "leaky_relu")   # refl.Positional (refl.py:230-245)
  elif refl_kind == "pos-linear-view":                                                                # refl.PosLinearView (refl.py:248-264)
    im = poslin_im
    d.refl = _mlp(6 + 4 * hash_levels + intermediate, 2, 3 + im, "leaky_relu")
    d.refl_view = _mlp(6 + intermediate + im, 2, 1, "sin", hidden=128)
  else: d.refl = _mlp(5 + ml + intermediate, 4, 3, "sin")                                             # refl.View (refl.py:190-207)
  d.intermediate = intermediate
  d.enc = _lib.ENC["hash"]
  d.hash_levels, d.hash_table_size, d.hash_feat = hash_levels, hash_table, 4
  for i in range(3): d.hash_primes[i] = HASH_PRIMES[i] & 0xFFFFFFFF
  for i, r in enumerate(_hash_res(hash_levels)): d.hash_res[i] = r
  d.density_act = _lib.DENSITY["softplus"]
  d.feat_act = _lib.FEAT[sigmoid]
  d.bg = _lib.BG[bg]
  return d


def describe_volsdf(sdf_kind: str = "siren", intermediate: int = 64, sigmoid: str = "upshifted", fourier_freqs: int = 128) -> ModelDesc:
  """Volume branch of VolSDF (reference src/nerf.py:981-1013): SDF network (`siren`: src/sdf.py:278-287, `mlp`:
  src/sdf.py:250-258) -> Laplace-CDF density with the learned `scale` -> View head (sdf.refl) -> composite, no sky."""
  d = ModelDesc()
  d.struct_bytes = C.sizeof(ModelDesc)
  d.kind = _lib.KIND["plain"]
  if sdf_kind == "siren":
    d.density = _mlp(3, 5, 1 + intermediate, "sin"); d.enc = _lib.ENC["none"]
  elif sdf_kind == "mlp":
    d.density = _mlp(3 + 2 * fourier_freqs, 6, 1 + intermediate, "leaky_relu"); d.enc = _lib.ENC["fourier"]
    d.fourier_freqs = fourier_freqs
  else: raise NotImplementedError(f"sdf kind {sdf_kind}")
  d.refl = _mlp(5 + intermediate, 4, 3, "sin")
  d.intermediate = intermediate
  d.density_act = _lib.DENSITY["laplace"]
  d.feat_act = _lib.FEAT[sigmoid]
  d.bg = _lib.BG["black"]          # volumetric_integrate only: no sky term (nerf.py:1013)
  return d


def describe_dyn(intermediate: int = 64, sigmoid: str = "upshifted", bg: str = "black", spline: int = 0) -> ModelDesc:
  """DynamicNeRF over a canonical PlainNeRF (reference src/nerf.py:1209-1303): ``spline == 0`` the direct deformation MLP
  (set_delta_estim, 1226-1241), ``spline == n > 1`` the hash-encoded MLP predicting n Bezier control points
  (set_spline_estim, 1242-1260)."""
  d = describe_plain(intermediate, sigmoid, bg)
  d.kind = _lib.KIND["dyn"]
  d.spline_points = spline
  if spline:
    d.deform = _mlp(6 + 4 * d.hash_levels, 5, 1 + 3 * spline, "leaky_relu"); d.deform_enc = _lib.ENC["hash"]
  else:
    d.deform = _mlp(4, 5, 4, "leaky_relu"); d.deform_enc = _lib.ENC["none"]
  return d


def describe_tiny(sigmoid: str = "upshifted", bg: str = "black") -> ModelDesc:
  """TinyNeRF (reference src/nerf.py:278-305), intended semantics."""
  d = ModelDesc()
  d.struct_bytes = C.sizeof(ModelDesc)
  d.kind = _lib.KIND["tiny"]
  d.density = _mlp(3, 6, 4, "leaky_relu")
  d.refl = _mlp(0, 0, 0, "none")
  d.enc = _lib.ENC["none"]
  d.density_act = _lib.DENSITY["softplus"]
  d.feat_act = _lib.FEAT[sigmoid]
  d.bg = _lib.BG[bg]
  return d


# ------------------------------------------------------------------------------------------------
# low-level engine: descriptor + packed blob + calls
# ------------------------------------------------------------------------------------------------
def _ptr(t: Optional[torch.Tensor]):
  return None if t is None else C.c_void_p(t.data_ptr())


def _chk(t: torch.Tensor, name: str, dtype=torch.float32):
  if not t.is_cuda: raise RuntimeError(f"{name} must be a CUDA tensor (no CPU fallback)")
  if t.dtype != dtype: raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
  if not t.is_contiguous(): raise ValueError(f"{name} must be contiguous")
  return t


class RenderEngine:
  """Thin object over the C ABI for one model description on one device."""

  def __init__(self, desc: ModelDesc, precision: str = "fp16"):
    self.desc = desc
    self.precision = precision
    self.lib = _lib.lib()
    n = self.lib.nf_param_count(C.byref(desc))
    if n < 0: _lib.check(n, "nf_param_count")
    self.n_params = n
    self.packed_bytes = int(self.lib.nf_packed_bytes(C.byref(desc)))
    if self.packed_bytes < 0: _lib.check(self.packed_bytes, "nf_packed_bytes")
    self.packed: Optional[torch.Tensor] = None
    self._key = None

  @staticmethod
  def _stream(): return C.c_void_p(torch.cuda.current_stream().cuda_stream)

  def pack(self, params: Sequence[torch.Tensor], force: bool = False):
    """Snapshot live parameters (order documented at nf_param_count) if anything changed."""
    if len(params) != self.n_params: raise ValueError(f"expected {self.n_params} parameter tensors, got {len(params)}")
    key = tuple((p.data_ptr(), p._version, p.device) for p in params)
    if not force and key == self._key and self.packed is not None: return False
    dev = params[0].device
    for i, p in enumerate(params): _chk(p, f"param[{i}]")
    if self.packed is None or self.packed.device != dev:
      raw = torch.empty(self.packed_bytes + 1024, dtype=torch.uint8, device=dev)
      off = (-raw.data_ptr()) % 1024
      self.packed = raw[off:off + self.packed_bytes]
      self._raw = raw
    arr = (C.c_void_p * len(params))(*[p.data_ptr() for p in params])
    with torch.cuda.device(dev):
      rc = self.lib.nf_pack_weights(C.byref(self.desc), arr, len(params), _ptr(self.packed), self.packed_bytes, self._stream())
    _lib.check(rc, "nf_pack_weights")
    self._key = key
    return True

  def _need_packed(self):
    if self.packed is None: raise RuntimeError("RenderEngine.pack(params) has not been called")

  @staticmethod
  def generate_rays(cam_to_world: torch.Tensor, focal: float, size: int, crop=None, reference_device: str = "cpu") -> torch.Tensor:
    """runner.render's pixel grid + NeRFCamera.sample_positions(with_noise=False) (reference runner.py:490-505,
    src/cameras.py:45-66) in one launch: cam_to_world[B,3,4] (CUDA) -> rays[B,H,W,6]; crop = (top, left, H, W), default the
    whole size x size image.  ``reference_device``: "cpu" = IEEE division by focal, "cuda" = torch's CUDA scalar-division
    (multiply by 1/focal); the two differ in the last bit of r_d."""
    _chk(cam_to_world, "cam_to_world")
    if cam_to_world.dim() != 3 or tuple(cam_to_world.shape[1:]) != (3, 4): raise ValueError("cam_to_world must be [B,3,4]")
    t, l, h, w = crop if crop is not None else (0, 0, size, size)
    B = cam_to_world.shape[0]
    rays = torch.empty(B, h, w, 6, dtype=torch.float32, device=cam_to_world.device)
    lib = _lib.lib()
    with torch.cuda.device(cam_to_world.device):
      rc = lib.nf_generate_rays(_ptr(cam_to_world), B, float(focal), int(size), int(t), int(l), int(h), int(w),
                                1 if reference_device == "cuda" else 0, _ptr(rays), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "nf_generate_rays")
    return rays

  @staticmethod
  def generate_rays_dtu(pose: torch.Tensor, intrinsic: torch.Tensor, size: int, crop=None) -> torch.Tensor:
    """runner.render's pixel grid + DTUCamera.sample_positions (reference src/cameras.py:189-223): pose[B,4,4], intrinsic[B,r,c]
    (CUDA) -> rays[B,H,W,6] with unit-norm r_d; crop = (top, left, H, W), default the whole size x size image."""
    _chk(pose, "pose"); _chk(intrinsic, "intrinsic")
    if pose.dim() != 3 or tuple(pose.shape[1:]) != (4, 4): raise ValueError("pose must be [B,4,4]")
    if intrinsic.dim() != 3 or intrinsic.shape[0] != pose.shape[0] or min(intrinsic.shape[1:]) < 3: raise ValueError("intrinsic must be [B,>=3,>=3]")
    t, l, h, w = crop if crop is not None else (0, 0, size, size)
    B = pose.shape[0]
    rays = torch.empty(B, h, w, 6, dtype=torch.float32, device=pose.device)
    lib = _lib.lib()
    with torch.cuda.device(pose.device):
      rc = lib.nf_generate_rays_dtu(_ptr(pose), _ptr(intrinsic), intrinsic.shape[1], intrinsic.shape[2], B, int(size), int(t), int(l), int(h), int(w),
                                    _ptr(rays), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "nf_generate_rays_dtu")
    return rays

  def ray_radii(self, rays_bhw: torch.Tensor) -> torch.Tensor:
    """radii_x (reference src/utils.py:77-81) of a crop rays[B,H,W,6] -> [B,H,W]."""
    _chk(rays_bhw, "rays")
    if rays_bhw.dim() != 4 or rays_bhw.shape[-1] != 6: raise ValueError("rays must be [B,H,W,6]")
    B, H, W, _ = rays_bhw.shape
    out = torch.empty(B, H, W, dtype=torch.float32, device=rays_bhw.device)
    with torch.cuda.device(rays_bhw.device):
      rc = self.lib.nf_ray_radii(_ptr(rays_bhw), B, H, W, _ptr(out), self._stream())
    _lib.check(rc, "nf_ray_radii")
    return out

  # ---- training (include/nerf_b200.h: nf_train_layout_of / nf_render_forward_aux / nf_render_backward) ----
  def train_layout(self, n_rays: int, T: int) -> TrainLayout:
    """Where the training forward stashes what the backward needs (``total_bytes`` = the workspace size)."""
    lay = TrainLayout()
    _lib.check(self.lib.nf_train_layout_of(C.byref(self.desc), int(n_rays), int(T), C.byref(lay)), "nf_train_layout_of")
    return lay

  @staticmethod
  def train_workspace(layout: TrainLayout, device) -> torch.Tensor:
    raw = torch.empty(int(layout.total_bytes) + 1024, dtype=torch.uint8, device=device)
    off = (-raw.data_ptr()) % 1024
    return raw[off:off + int(layout.total_bytes)]

  def render_backward(self, ws: torch.Tensor, rays: torch.Tensor, ts: torch.Tensor, d_rgb: torch.Tensor, grads: Sequence[Optional[torch.Tensor]]):
    """Backward of the ``render(..., train_ws=ws)`` call that filled ``ws``: d_rgb[R,3] -> ``grads[i]`` (same order and shapes
    as the packed parameters; None entries are skipped, the others overwritten)."""
    self._need_packed()
    _chk(rays, "rays"); _chk(ts, "ts"); _chk(d_rgb, "d_rgb"); _chk(ws, "ws", torch.uint8)
    R = rays.shape[0]
    T, stride = (ts.shape[0], 0) if ts.dim() == 1 else (ts.shape[1], ts.shape[1])
    if tuple(d_rgb.shape) != (R, 3): raise ValueError("d_rgb must be [R,3]")
    if len(grads) != self.n_params: raise ValueError(f"expected {self.n_params} gradient slots, got {len(grads)}")
    for g in grads:
      if g is not None: _chk(g, "grad")
    arr = (C.c_void_p * len(grads))(*[None if g is None else g.data_ptr() for g in grads])
    with torch.cuda.device(rays.device):
      rc = self.lib.nf_render_backward(C.byref(self.desc), _ptr(self.packed), _ptr(ws), ws.numel(), _ptr(rays), R, _ptr(ts), T, stride,
                                       _ptr(d_rgb), arr, len(grads), self._stream())
    _lib.check(rc, "nf_render_backward")

  def render(self, rays: torch.Tensor, ts: torch.Tensor, density_noise: Optional[torch.Tensor] = None,
             want_weights: bool = True, precision: Optional[str] = None, ray_time: Optional[torch.Tensor] = None,
             radius: Optional[torch.Tensor] = None, crop: Optional[tuple] = None, train_ws: Optional[torch.Tensor] = None,
             pts: Optional[torch.Tensor] = None, bg_rand: Optional[torch.Tensor] = None, side: Optional[dict] = None):
    """rays[R,6], ts[T] (shared) or ts[R,T] (per ray) -> rgb[R,3], alpha[R,T]|None, weights[R,T]|None.
    Mip models: ``radius[R]`` (``ray_radii``); "cylinder_ref" also takes ``crop = (rays_all[R_all,6], radius_all[R_all],
    ray_base)`` when ``rays`` is a shard of a larger crop (default: the call's rays are the whole crop).
    ``pts[R,T,3]``: explicit sample positions (from_pts).  ``bg_rand[R]``: the draws of the "random" background.
    ``side`` (a dict, DynamicNeRF): filled with pts [R,T,3], dp, rigid_dp [R,T,3], rigidity (ray-major side channels)."""
    self._need_packed()
    _chk(rays, "rays"); _chk(ts, "ts")
    if rays.dim() != 2 or rays.shape[1] != 6: raise ValueError("rays must be [R,6]")
    R = rays.shape[0]
    if ts.dim() == 1: T, stride = ts.shape[0], 0
    elif ts.dim() == 2 and ts.shape[0] == R: T, stride = ts.shape[1], ts.shape[1]
    else: raise ValueError("ts must be [T] or [R,T]")
    if density_noise is not None:
      _chk(density_noise, "density_noise")
      if tuple(density_noise.shape) != (R, T): raise ValueError("density_noise must be [R,T]")
    if ray_time is not None:
      _chk(ray_time, "ray_time")
      if tuple(ray_time.shape) != (R,): raise ValueError("ray_time must be [R]")
    rgb = torch.empty(R, 3, dtype=torch.float32, device=rays.device)
    alpha = torch.empty(R, T, dtype=torch.float32, device=rays.device) if want_weights else None
    weights = torch.empty(R, T, dtype=torch.float32, device=rays.device) if want_weights else None
    mip = None
    if self.desc.mip:
      if radius is None: raise ValueError("this model has a Mip encoder: pass radius[R] (RenderEngine.ray_radii)")
      _chk(radius, "radius")
      if radius.numel() != R: raise ValueError("radius must be [R]")
      rays_all, radius_all, base = crop if crop is not None else (rays, radius, 0)
      _chk(rays_all, "rays_all"); _chk(radius_all, "radius_all")
      mip = MipArgs(radius.data_ptr(), rays_all.data_ptr(), radius_all.data_ptr(), rays_all.shape[0], base)
    aux = None
    if train_ws is not None or pts is not None or bg_rand is not None or side is not None:
      aux = RenderAux(); aux.struct_bytes = C.sizeof(RenderAux)
    if train_ws is not None:
      _chk(train_ws, "train_ws", torch.uint8)
      aux.train_ws, aux.train_ws_bytes = train_ws.data_ptr(), train_ws.numel()
    if pts is not None:
      _chk(pts, "pts")
      if tuple(pts.shape) != (R, T, 3): raise ValueError("pts must be [R,T,3]")
      aux.pts = pts.data_ptr()
    if bg_rand is not None:
      _chk(bg_rand, "bg_rand")
      if bg_rand.numel() != R: raise ValueError("bg_rand must be [R]")
      aux.bg_rand = bg_rand.data_ptr()
    if side is not None:
      if self.desc.kind != _lib.KIND["dyn"]: raise ValueError("side channels exist for DynamicNeRF only")
      direct = self.desc.spline_points == 0
      f = lambda c: torch.empty(R, T, c, dtype=torch.float32, device=rays.device)
      side.update(pts=f(3), dp=f(1 if direct else 3), rigid_dp=f(3), rigidity=f(3 if direct else 1))
      aux.pts_out, aux.dp_out = side["pts"].data_ptr(), side["dp"].data_ptr()
      aux.rigid_dp_out, aux.rigidity_out = side["rigid_dp"].data_ptr(), side["rigidity"].data_ptr()
    with torch.cuda.device(rays.device):
      rc = self.lib.nf_render_forward_aux(C.byref(self.desc), _ptr(self.packed), _ptr(rays), R, _ptr(ts), T, stride,
                                          _ptr(density_noise), _ptr(ray_time), C.byref(mip) if mip is not None else None,
                                          _ptr(rgb), _ptr(alpha), _ptr(weights), C.byref(aux) if aux is not None else None,
                                          _lib.PRECISION[precision or self.precision], self._stream())
    _lib.check(rc, "nf_render_forward")
    return rgb, alpha, weights

  # ---- the SDF surface side (include/nerf_b200.h: nf_sphere_march / nf_sdf_render) ----
  def _sdf_ws(self, n_rays: int, device) -> torch.Tensor:
    nbytes = int(self.lib.nf_sdf_workspace_bytes(C.byref(self.desc), int(n_rays)))
    if nbytes < 0: _lib.check(nbytes, "nf_sdf_workspace_bytes")
    raw = torch.empty(nbytes + 256, dtype=torch.uint8, device=device)
    off = (-raw.data_ptr()) % 256
    return raw[off:off + nbytes]

  def sphere_march(self, rays: torch.Tensor, near: float, far: float, iters: int = 32, eps: float = 1e-3, bound_rad: float = -1.0,
                   precision: Optional[str] = None):
    """march.sphere_march (reference src/march.py:27-47) of the descriptor's SDF network: rays[R,6] -> (pts[R,3], hit[R] bool, t[R])."""
    self._need_packed(); _chk(rays, "rays")
    R = rays.shape[0]
    pts = torch.empty(R, 3, dtype=torch.float32, device=rays.device); t = torch.empty(R, dtype=torch.float32, device=rays.device)
    hit = torch.empty(R, dtype=torch.uint8, device=rays.device)
    ws = self._sdf_ws(R, rays.device)
    with torch.cuda.device(rays.device):
      rc = self.lib.nf_sphere_march(C.byref(self.desc), _ptr(self.packed), _ptr(rays), R, float(near), float(far), int(iters), float(eps), float(bound_rad),
                                    _lib.PRECISION[precision or self.precision], _ptr(pts), _ptr(hit), _ptr(t), _ptr(ws), ws.numel(), self._stream())
    _lib.check(rc, "nf_sphere_march")
    return pts, hit.bool(), t

  def sdf_render(self, rays: torch.Tensor, near: float, far: float, iters: int = 192, eps: float = 1e-3, bound_rad: float = -1.0,
                 precision: Optional[str] = None):
    """SDF.forward in eval mode (reference src/sdf.py:137-156): rays[R,6] -> (rgb[R,3], hit[R] bool, t[R], pts[R,3])."""
    self._need_packed(); _chk(rays, "rays")
    R = rays.shape[0]
    rgb = torch.empty(R, 3, dtype=torch.float32, device=rays.device); pts = torch.empty(R, 3, dtype=torch.float32, device=rays.device)
    t = torch.empty(R, dtype=torch.float32, device=rays.device); hit = torch.empty(R, dtype=torch.uint8, device=rays.device)
    ws = self._sdf_ws(R, rays.device)
    with torch.cuda.device(rays.device):
      rc = self.lib.nf_sdf_render(C.byref(self.desc), _ptr(self.packed), _ptr(rays), R, float(near), float(far), int(iters), float(eps), float(bound_rad),
                                  _lib.PRECISION[precision or self.precision], _ptr(rgb), _ptr(hit), _ptr(t), _ptr(pts), _ptr(ws), ws.numel(), self._stream())
    _lib.check(rc, "nf_sdf_render")
    return rgb, hit.bool(), t, pts

  def sdf_normals(self, pts: torch.Tensor, bound_rad: float = -1.0, want_values: bool = False):
    """SDFModel.normals (reference src/sdf.py:43-49): pts[N,3] -> normals[N,3] (+ the network's outputs [N, 1 + I])."""
    self._need_packed(); _chk(pts, "pts")
    n, dev = pts.shape[0], pts.device
    nrm = torch.empty(n, 3, dtype=torch.float32, device=dev)
    vals = torch.empty(n, 1 + self.desc.intermediate, dtype=torch.float32, device=dev) if want_values else None
    with torch.cuda.device(dev):
      rc = self.lib.nf_sdf_normals(C.byref(self.desc), _ptr(self.packed), _ptr(pts), n, float(bound_rad), _ptr(nrm), _ptr(vals) if want_values else None, self._stream())
    _lib.check(rc, "nf_sdf_normals")
    return (nrm, vals) if want_values else nrm

  def sdf_bisect(self, rays: torch.Tensor, near: float, far: float, iters: int = 192, jitter: float = 0.0, bound_rad: float = -1.0,
                 precision: Optional[str] = None, shade: bool = True):
    """march.bisect (+ SDF.forward's shading when `shade`): rays[R,6] -> (rgb[R,3] | None, hit[R] bool, tput[R], pts[R,3], best_pos[R,3]).
    `jitter` = the reference's random.random() draw (src/march.py:86)."""
    self._need_packed(); _chk(rays, "rays")
    R, dev = rays.shape[0], rays.device
    rgb = torch.empty(R, 3, dtype=torch.float32, device=dev) if shade else None
    pts = torch.empty(R, 3, dtype=torch.float32, device=dev); best = torch.empty(R, 3, dtype=torch.float32, device=dev)
    tput = torch.empty(R, dtype=torch.float32, device=dev); hit = torch.empty(R, dtype=torch.uint8, device=dev)
    ws = self._sdf_ws(R, dev)
    with torch.cuda.device(dev):
      rc = self.lib.nf_sdf_bisect(C.byref(self.desc), _ptr(self.packed), _ptr(rays), R, float(near), float(far), int(iters), float(jitter), float(bound_rad),
                                  _lib.PRECISION[precision or self.precision], _ptr(pts), _ptr(hit), _ptr(tput), _ptr(best), _ptr(rgb) if shade else None,
                                  _ptr(ws), ws.numel(), self._stream())
    _lib.check(rc, "nf_sdf_bisect")
    return rgb, hit.bool(), tput, pts, best

  # ---- stage entry points (parity tests, micro-benchmarks) ----
  def sample_points(self, rays: torch.Tensor, ts: torch.Tensor) -> torch.Tensor:
    _chk(rays, "rays"); _chk(ts, "ts")
    R = rays.shape[0]
    T, stride = (ts.shape[0], 0) if ts.dim() == 1 else (ts.shape[1], ts.shape[1])
    pts = torch.empty(R, T, 3, dtype=torch.float32, device=rays.device)
    with torch.cuda.device(rays.device):
      rc = self.lib.nf_sample_points(_ptr(rays), R, _ptr(ts), T, stride, _ptr(pts), self._stream())
    _lib.check(rc, "nf_sample_points")
    return pts

  def hash_encode(self, pts: torch.Tensor, want_indices: bool = False):
    self._need_packed(); _chk(pts, "pts")
    n, L = pts.shape[0], self.desc.hash_levels
    feats = torch.empty(n, L * 4, dtype=torch.float32, device=pts.device)
    idx = torch.empty(L, 8, n, dtype=torch.uint16, device=pts.device) if want_indices else None
    with torch.cuda.device(pts.device):
      rc = self.lib.nf_hash_encode(C.byref(self.desc), _ptr(self.packed), _ptr(pts), n, _ptr(feats), _ptr(idx), self._stream())
    _lib.check(rc, "nf_hash_encode")
    return feats, idx

  def composite(self, sigma_raw: torch.Tensor, feats: torch.Tensor, rays: torch.Tensor, ts: torch.Tensor, want_weights: bool = True):
    _chk(sigma_raw, "sigma_raw"); _chk(feats, "feats"); _chk(rays, "rays"); _chk(ts, "ts")
    R, T = sigma_raw.shape
    stride = 0 if ts.dim() == 1 else T
    rgb = torch.empty(R, 3, dtype=torch.float32, device=rays.device)
    alpha = torch.empty(R, T, dtype=torch.float32, device=rays.device) if want_weights else None
    weights = torch.empty(R, T, dtype=torch.float32, device=rays.device) if want_weights else None
    with torch.cuda.device(rays.device):
      rc = self.lib.nf_composite(C.byref(self.desc), _ptr(self.packed), _ptr(sigma_raw), _ptr(feats), _ptr(rays), R, _ptr(ts), T, stride,
                                 _ptr(rgb), _ptr(alpha), _ptr(weights), self._stream())
    _lib.check(rc, "nf_composite")
    return rgb, alpha, weights

  def composite_backward(self, sigma_raw: torch.Tensor, feats: torch.Tensor, rays: torch.Tensor, ts: torch.Tensor, d_rgb: torch.Tensor):
    """Backward of ``composite``: d_rgb[R,3] -> (d_sigma_raw[R,T], d_feats[R,T,3])."""
    for t, n in ((sigma_raw, "sigma_raw"), (feats, "feats"), (rays, "rays"), (ts, "ts"), (d_rgb, "d_rgb")): _chk(t, n)
    R, T = sigma_raw.shape
    stride = 0 if ts.dim() == 1 else T
    d_sigma = torch.empty_like(sigma_raw); d_feats = torch.empty_like(feats)
    with torch.cuda.device(rays.device):
      rc = self.lib.nf_composite_backward(C.byref(self.desc), _ptr(self.packed), _ptr(sigma_raw), _ptr(feats), _ptr(rays), R, _ptr(ts), T, stride,
                                          _ptr(d_rgb), _ptr(d_sigma), _ptr(d_feats), self._stream())
    _lib.check(rc, "nf_composite_backward")
    return d_sigma, d_feats

  def hash_encode_backward(self, pts: torch.Tensor, d_feats: torch.Tensor, d_tables: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Backward of ``hash_encode`` w.r.t. the tables: scatter-adds into (and returns) d_tables[levels, table, 4]."""
    _chk(pts, "pts"); _chk(d_feats, "d_feats")
    L, table = self.desc.hash_levels, self.desc.hash_table_size
    if d_tables is None: d_tables = torch.zeros(L, table, 4, dtype=torch.float32, device=pts.device)
    _chk(d_tables, "d_tables")
    with torch.cuda.device(pts.device):
      rc = self.lib.nf_hash_encode_backward(C.byref(self.desc), _ptr(pts), pts.shape[0], _ptr(d_feats), _ptr(d_tables), self._stream())
    _lib.check(rc, "nf_hash_encode_backward")
    return d_tables

  def sample_pdf(self, ts_coarse: torch.Tensor, weights: torch.Tensor, u: torch.Tensor) -> torch.Tensor:
    """ts_coarse[T], weights[R,T] (coarse pass), u[R,Nf] in [0,1) -> sorted per-ray ts[R, T+Nf] for the fine pass."""
    _chk(ts_coarse, "ts_coarse"); _chk(weights, "weights"); _chk(u, "u")
    R, T = weights.shape
    if ts_coarse.shape != (T,) or u.shape[0] != R: raise ValueError("shapes: ts_coarse[T], weights[R,T], u[R,Nf]")
    out = torch.empty(R, T + u.shape[1], dtype=torch.float32, device=weights.device)
    with torch.cuda.device(weights.device):
      rc = self.lib.nf_sample_pdf(_ptr(ts_coarse), T, _ptr(weights), R, _ptr(u), u.shape[1], _ptr(out), self._stream())
    _lib.check(rc, "nf_sample_pdf")
    return out

  def render_coarse_fine(self, rays: torch.Tensor, ts_coarse: torch.Tensor, u: torch.Tensor, want_weights: bool = True):
    """Config 2 of BASELINE.json (coarse + fine): coarse pass on shared ts_coarse[T], inverse-CDF resampling
    at u[R,Nf], fine pass on the T+Nf merged per-ray positions.  Returns (rgb_fine, rgb_coarse, ts_fine, alpha, weights)."""
    rgb_c, _, w_c = self.render(rays, ts_coarse, want_weights=True)
    ts_f = self.sample_pdf(ts_coarse, w_c, u)
    rgb_f, alpha, w_f = self.render(rays, ts_f, want_weights=want_weights)
    return rgb_f, rgb_c, ts_f, alpha, w_f

  def mlp_forward(self, which: int, x0: torch.Tensor, precision: Optional[str] = None) -> torch.Tensor:
    self._need_packed(); _chk(x0, "x0")
    md = (self.desc.density, self.desc.refl, self.desc.deform)[which]
    if x0.dim() != 2 or x0.shape[1] != md.in_dims: raise ValueError(f"x0 must be [N,{md.in_dims}]")
    out = torch.empty(x0.shape[0], md.out_dims, dtype=torch.float32, device=x0.device)
    with torch.cuda.device(x0.device):
      rc = self.lib.nf_mlp_forward(C.byref(self.desc), _ptr(self.packed), which, _ptr(x0), x0.shape[0], _ptr(out),
                                   _lib.PRECISION[precision or self.precision], self._stream())
    _lib.check(rc, "nf_mlp_forward")
    return out


# ------------------------------------------------------------------------------------------------
# parameter containers with the reference's state_dict names
# ------------------------------------------------------------------------------------------------
class HashParams(nn.Module):
  """Parameters of HashEncoder (reference src/neural_blocks.py:92-131)."""
  def __init__(self, levels: int = 8, emb_size: int = 1 << 16, feat_size: int = 4):
    super().__init__()
    self.register_buffer("primes", torch.tensor(HASH_PRIMES), persistent=True)
    self.embs = nn.ModuleList([nn.Embedding(emb_size, feat_size) for _ in range(levels)])
    self.levels, self.emb_size, self.feat_size = levels, emb_size, feat_size


class SkipConnParams(nn.Module):
  """Parameters (and init kinds) of SkipConnMLP (reference src/neural_blocks.py:204-277)."""
  def __init__(self, in_dims: int, out: int, num_layers: int, hidden_size: int = 256, skip: int = 3,
               init: Optional[str] = None, enc: Optional[nn.Module] = None):
    super().__init__()
    self.enc = enc
    self.dim_p, self.skip = in_dims, skip
    self.init = nn.Linear(in_dims, hidden_size)
    self.layers = nn.ModuleList([
      nn.Linear(hidden_size + in_dims if (i % skip) == 0 and i != num_layers - 1 else hidden_size, hidden_size)
      for i in range(num_layers)])
    self.out = nn.Linear(hidden_size, out)
    ws = [self.init.weight, self.out.weight, *[l.weight for l in self.layers]]
    bs = [self.init.bias, self.out.bias, *[l.bias for l in self.layers]]
    if init == "xavier":
      for t in ws: nn.init.xavier_uniform_(t)
      for t in bs: nn.init.zeros_(t)
    elif init == "siren":
      for t in ws:
        fan_in, _ = nn.init._calculate_fan_in_and_fan_out(t)
        a = math.sqrt(6 / fan_in)
        nn.init._no_grad_uniform_(t, -a, a)
      for t in bs: nn.init.zeros_(t)
    elif init is not None: raise NotImplementedError(init)

  def linears(self) -> List[nn.Linear]: return [self.init, *self.layers, self.out]


def _linears_of(mlp) -> List[nn.Module]:
  """Works for SkipConnParams and for the reference's own SkipConnMLP (duck-typed)."""
  return [mlp.init, *list(mlp.layers), mlp.out]


class ViewHead(nn.Module):
  """Parameters of refl.View (reference src/refl.py:190-207); ``act`` is a sigmoid-kind name."""
  def __init__(self, latent_size: int, out_features: int = 3, act: str = "thin"):
    super().__init__()
    self.latent_size, self.out_features = latent_size, out_features
    self.act = act
    self.mlp = SkipConnParams(5 + latent_size, out_features, 4, init="siren")


class PositionalHead(nn.Module):
  """Parameters of refl.Positional (reference src/refl.py:230-245): view-independent colour from [p, hash'(p), latent]."""
  def __init__(self, latent_size: int, out_features: int = 3, act: str = "thin"):
    super().__init__()
    self.latent_size, self.out_features = latent_size, out_features
    self.act = act
    self.mlp = SkipConnParams(38 + latent_size, out_features, 5, enc=HashParams())


class PosLinearViewHead(nn.Module):
  """Parameters of refl.PosLinearView (reference src/refl.py:248-290): `pos` (own hash tables, 2 layers, hidden 256) gives a
  view-independent colour and an intermediate vector, `view` (2 layers, hidden 128, sin) a view-dependent scale in [1/2, 1]."""
  def __init__(self, latent_size: int, out_features: int = 3, act: str = "thin", intermediate_size: int = 64):
    super().__init__()
    self.latent_size, self.out_features, self.im = latent_size, out_features, intermediate_size
    self.act = act
    self.pos = SkipConnParams(38 + latent_size, out_features + intermediate_size, 2, enc=HashParams())
    self.view = SkipConnParams(6 + latent_size + intermediate_size, 1, 2, hidden_size=128, init="siren")


_ACT_NAMES = {"sigmoid": "normal", "thin_sigmoid": "thin", "tanh": "tanh", "cyclic_sigmoid": "cyclic",
              "upshifted_sigmoid": "upshifted", "fat_sigmoid": "fat", "leaky_relu": "leaky_relu", "relu": "relu",
              "sin": "sin", "upshifted_softplus": "upshifted_softplus", "upshifted_relu": "upshifted_relu"}


def _sigmoid_name(act) -> str:
  if isinstance(act, str): return act
  name = getattr(act, "__name__", None)
  if name in _ACT_NAMES: return _ACT_NAMES[name]
  if type(act).__name__ == "Softmax" and getattr(act, "dim", None) == -1: return "softmax"      # nn.Softmax(dim=-1), utils.py:507
  raise NotImplementedError(f"feature activation {act!r} is not supported by the fused path")


# ------------------------------------------------------------------------------------------------
# the reference-facing modules
# ------------------------------------------------------------------------------------------------
class FusedNeRF(nn.Module):
  """Common part (mirrors CommonNeRF, reference src/nerf.py:147-276)."""
  kind = "plain"

  def __init__(self, steps: int = 64, t_near: float = 0, t_far: float = 1, intermediate_size: int = 32,
               sigmoid_kind: str = "thin", bg: str = "black", precision: str = "fp16", keep_weights: bool = True,
               **unused):
    super().__init__()
    for k in ("per_pixel_latent_size", "per_point_latent_size", "instance_latent_size"):
      if unused.get(k): raise NotImplementedError(f"{k} is not supported by the fused path yet")
    # mip: None | "cylinder" | "cone" (the encoder as intended) | "cylinder_ref" (the reference's CylinderGaussian bug for
    # bug); an instance of the reference's utils.CylinderGaussian maps to "cylinder_ref" (reference src/utils.py:103-140)
    mip = unused.get("mip")
    if mip is not None and not isinstance(mip, str):
      name = type(mip).__name__
      if name == "CylinderGaussian": mip = "cylinder_ref"
      elif name == "ConicGaussian": raise NotImplementedError("the reference's ConicGaussian renders NaN (SURVEY.md a-4); use mip='cone'")
      else: raise NotImplementedError(f"mip encoder {name}")
    if mip not in _lib.MIP: raise NotImplementedError(f"mip kind {mip!r}")
    self.mip = mip
    self.empty_latent = nn.Parameter(torch.zeros(1, 1, 1, 1, 0, dtype=torch.float), requires_grad=False)
    self.t_near, self.t_far, self.steps = t_near, t_far, steps
    self.intermediate_size = intermediate_size
    self.noise_std = 0.2                                    # reference src/nerf.py:197
    self.precision, self.keep_weights = precision, keep_weights
    self.sigmoid_kind, self.bg = sigmoid_kind, bg
    self.alpha = self.weights = self.ts = None
    self._engine: Optional[RenderEngine] = None
    self._engine_key = None

  # ---- surface the runner touches ----
  @property
  def nerf(self): return self
  def mip_size(self) -> int: return MIP_FEATS if _lib.MIP[getattr(self, "mip", None)] else 0
  def total_latent_size(self) -> int: return self.mip_size()
  def set_bg(self, bg="black"):
    if bg not in _lib.BG: raise NotImplementedError(bg)
    self.bg = bg
  def set_sigmoid(self, kind="thin"):
    if kind not in _lib.FEAT: raise NotImplementedError(f"Unknown sigmoid kind({kind})")
    self.sigmoid_kind = kind
    if hasattr(self, "refl"): self.refl.act = kind
  def set_refl(self, refl):
    if hasattr(self, "refl"): self.refl = refl

  # ---- pickling: never carry the ctypes handle / packed blob (checkpoint = torch.save(model)) ----
  def __getstate__(self):
    st = self.__dict__.copy(); st["_engine"] = None; st["_engine_key"] = None
    st["alpha"] = st["weights"] = st["ts"] = None
    if "scale_post_act" in st: st["scale_post_act"] = None
    st.pop("_scaled_basis", None); st.pop("_scaled_basis_key", None)
    return st

  # ---- implemented by subclasses ----
  def _describe(self) -> ModelDesc: raise NotImplementedError
  def _param_list(self) -> List[torch.Tensor]: raise NotImplementedError

  def engine(self) -> RenderEngine:
    key = (self.kind, self.sigmoid_kind if not hasattr(self, "refl") else _sigmoid_name(self.refl.act), self.bg, self.precision,
           getattr(self, "mip", None), getattr(self, "refl_kind", "view"))
    if self._engine is None or self._engine_key != key:
      self._engine, self._engine_key = RenderEngine(self._describe(), self.precision), key
    return self._engine

  def render_views(self, cam_to_world: torch.Tensor, focal: float, size: int, crop=None, reference_device: str = "cpu") -> torch.Tensor:
    """runner.render without the host-side pixel grid (reference runner.py:490-509): cam_to_world[B,3,4] (CUDA), focal, image size
    and an optional crop (top, left, H, W) -> rgb[B,H,W,3]; the rays come from `nf_generate_rays` (bit-exact with
    NeRFCamera.sample_positions), so a frame needs 48 bytes of input per view instead of 24 bytes per ray."""
    return self(RenderEngine.generate_rays(cam_to_world, focal, size, crop, reference_device))

  def _sample_ts(self, device, n_rays: int, with_noise: bool):
    """reference src/nerf.py:29-47 -- linspace on the rays' device; in training mode stratified jitter with ONE rand[T] shared
    by all rays, and (src/nerf.py:347-348) randn * noise_std added to the raw density."""
    ts = torch.linspace(self.t_near, self.t_far, steps=self.steps, device=device, dtype=torch.float32)
    noise = None
    if self.training:
      mids = 0.5 * (ts[:-1] + ts[1:])
      lower, upper = torch.cat([mids, ts[-1:]]), torch.cat([ts[:1], mids])
      ts = lower + (upper - lower) * torch.rand_like(lower)
      if with_noise and self.noise_std > 0:
        noise = torch.randn(n_rays, self.steps, device=device) * self.noise_std
    return ts, noise

  def _wants_grad(self, params) -> bool:
    """Autograd flows to the parameters (loss.backward(), reference runner.py:820) when grad mode is on and the module is in
    training mode -- or always / never with ``self.differentiable = True / False``.  The differentiable call stashes ~1.7 MB
    of activations per 128 samples for the backward, so an evaluation render that forgot ``torch.no_grad()`` stays cheap."""
    if not torch.is_grad_enabled() or not any(p.requires_grad for p in params): return False
    d = getattr(self, "differentiable", None)
    return self.training if d is None else bool(d)

  def from_pts(self, pts: torch.Tensor, ts: torch.Tensor, r_o: torch.Tensor, r_d: torch.Tensor, refl_latent=None) -> torch.Tensor:
    """The reference's inner entry (src/nerf.py:340-361; called by DynamicNeRF.forward 1303, render_keyframes 1317, BendyNeRF
    710): explicit sample positions ``pts[T,B,H,W,3]`` instead of r_o + ts r_d; ``ts[T]`` gives the segment lengths, ``r_d`` the
    view direction (and the ray norm of the deltas).  Density noise is added in training mode exactly as in ``forward``."""
    if refl_latent is not None: raise NotImplementedError("refl_latent is not supported by the fused path")
    if self.mip_size(): raise NotImplementedError("from_pts of a Mip model")
    if not pts.is_cuda: raise RuntimeError("FusedNeRF.from_pts needs CUDA tensors: the fused path has no CPU fallback")
    T = pts.shape[0]; B = pts.shape[1:-1]
    flat_pts = pts.to(torch.float32).movedim(0, -2).reshape(-1, T, 3).contiguous()         # [R,T,3]
    rays = torch.cat([r_o.expand(*B, 3), r_d.expand(*B, 3)], dim=-1).reshape(-1, 6).to(torch.float32).contiguous()
    R = rays.shape[0]
    noise = torch.randn(R, T, device=pts.device) * self.noise_std if (self.training and self.noise_std > 0 and self.kind == "plain") else None
    eng = self.engine(); params = self._param_list()
    if self._wants_grad(params) or pts.requires_grad: raise NotImplementedError("from_pts is not differentiable on the fused path")
    eng.pack(params)
    bg_rand = torch.rand(R, device=pts.device) if self.bg == "random" else None
    rgb, alpha, weights = eng.render(rays, ts.to(torch.float32).contiguous(), noise, want_weights=self.keep_weights, pts=flat_pts, bg_rand=bg_rand)
    self.ts = ts
    if self.keep_weights:
      self.alpha = alpha.reshape(*B, T).movedim(-1, 0); self.weights = weights.reshape(*B, T).movedim(-1, 0)
    return rgb.reshape(*B, 3)

  def forward(self, rays: torch.Tensor) -> torch.Tensor:
    if not rays.is_cuda: raise RuntimeError("FusedNeRF.forward needs CUDA rays: the fused path has no CPU fallback")
    B = rays.shape[:-1]
    flat = rays.reshape(-1, 6).to(torch.float32).contiguous()
    ts, noise = self._sample_ts(rays.device, flat.shape[0], with_noise=self.kind == "plain")
    bg_rand = torch.rand(flat.shape[0], device=rays.device) if self.bg == "random" else None     # random_color (nerf.py:100-103)
    eng = self.engine()
    params = self._param_list()
    radius = None
    if self.mip_size():
      if rays.dim() != 4: raise ValueError("a Mip model needs rays[B,H,W,6]: the pixel radius differences neighbouring rows (utils.py:77-81)")
      radius = eng.ray_radii(rays.to(torch.float32).contiguous()).reshape(-1)
    if self._wants_grad(params):
      if rays.requires_grad: raise NotImplementedError("gradients with respect to the rays (--train-parts camera) are not built")
      if radius is not None: raise NotImplementedError("training a Mip model through the fused path is not built")
      from .autograd import fused_render
      rgb, alpha, weights = fused_render(eng, flat.detach(), ts, params, noise, want_weights=self.keep_weights, bg_rand=bg_rand)
    else:
      eng.pack(params)
      rgb, alpha, weights = eng.render(flat, ts, noise, want_weights=self.keep_weights, radius=radius, bg_rand=bg_rand)
    self.ts = ts
    if self.keep_weights:   # the reference keeps [T,B,H,W]; these are transposed views of the [R,T] buffers
      self.alpha = alpha.reshape(*B, self.steps).movedim(-1, 0)
      self.weights = weights.reshape(*B, self.steps).movedim(-1, 0)
    return rgb.reshape(*B, 3)


class FusedPlainNeRF(FusedNeRF):
  """Drop-in for PlainNeRF + View (reference src/nerf.py:310-361)."""
  kind = "plain"

  def __init__(self, out_features: int = 3, refl_kind: str = "view", **kwargs):
    kwargs.setdefault("sigmoid_kind", "thin")
    super().__init__(**kwargs)
    if out_features != 3: raise NotImplementedError("out_features != 3")
    if refl_kind not in _lib.REFL: raise NotImplementedError(f"refl kind {refl_kind!r} (view | pos)")
    head = {"pos": PositionalHead, "pos-linear-view": PosLinearViewHead}.get(refl_kind, ViewHead)   # runner.load_model: refl.load(args, refl_kind, ...) (runner.py:1182-1183)
    if refl_kind == "pos-linear-view": self.precision = "fp32"                                     # PosLinearView: the fp32 pipeline only
    self.refl = head(latent_size=self.mip_size() + self.intermediate_size, out_features=out_features, act=self.sigmoid_kind)
    self.first = SkipConnParams(38 + self.mip_size(), 1 + self.intermediate_size, 4, enc=HashParams())

  @classmethod
  def from_reference(cls, ref, precision: str = "fp16", keep_weights: bool = True) -> "FusedPlainNeRF":
    """Adopt a live reference PlainNeRF: its ``first`` and ``refl`` sub-modules become ours, so
    parameters (and optimiser references to them) are shared, not copied."""
    self = cls.__new__(cls)
    FusedNeRF.__init__(self, steps=ref.steps, t_near=ref.t_near, t_far=ref.t_far, intermediate_size=ref.intermediate_size,
                       sigmoid_kind=_sigmoid_name(ref.refl.act), precision=precision, keep_weights=keep_weights,
                       mip=getattr(ref, "mip", None))
    bg = [k for k, v in {"black": "black", "white": "white", "random": "random_color"}.items() if getattr(ref.sky_color, "__name__", "") == v]
    if not bg: raise NotImplementedError("background kind of the reference model")
    self.bg = bg[0]
    if type(ref.refl).__name__ not in ("View", "ViewHead", "Positional", "PositionalHead", "PosLinearView", "PosLinearViewHead"):
      raise NotImplementedError(f"refl head {type(ref.refl).__name__}")
    self.refl, self.first = ref.refl, ref.first
    return self

  def _describe(self) -> ModelDesc:
    enc = self.first.enc
    levels = len(enc.embs)
    return describe_plain(self.intermediate_size, _sigmoid_name(self.refl.act), self.bg, levels, enc.embs[0].weight.shape[0],
                          mip=getattr(self, "mip", None), refl_kind=self.refl_kind, poslin_im=int(getattr(self.refl, "im", 64)))

  @property
  def refl_kind(self) -> str:
    n = type(self.refl).__name__
    return "pos" if n in ("Positional", "PositionalHead") else "pos-linear-view" if n in ("PosLinearView", "PosLinearViewHead") else "view"

  def _param_list(self) -> List[torch.Tensor]:
    ps: List[torch.Tensor] = []
    plv = self.refl_kind == "pos-linear-view"
    for mlp in ((self.first, self.refl.pos, self.refl.view) if plv else (self.first, self.refl.mlp)):
      for lin in _linears_of(mlp): ps += [lin.weight, lin.bias]
    ps += [e.weight for e in self.first.enc.embs]
    if self.refl_kind == "pos": ps += [e.weight for e in self.refl.mlp.enc.embs]
    if plv: ps += [e.weight for e in self.refl.pos.enc.embs]
    return ps


class _SdfBox(nn.Module):
  """Container mirroring the reference's `SDF` (src/sdf.py:83-112): `.underlying.{siren|mlp}` and `.refl`."""
  def __init__(self, underlying: nn.Module, refl: nn.Module):
    super().__init__(); self.underlying = underlying; self.refl = refl


class _SdfNet(nn.Module):
  def __init__(self, kind: str, intermediate: int, freqs: int = 128):
    super().__init__()
    self.kind, self.intermediate_size = kind, intermediate
    if kind == "siren": self.siren = SkipConnParams(3, 1 + intermediate, 5, init="siren")
    else:
      enc = nn.Module(); enc.basis = nn.Parameter(16 * torch.randn(freqs, 3).T.contiguous(), requires_grad=False)
      self.mlp = SkipConnParams(3 + 2 * freqs, 1 + intermediate, 6, init="xavier", enc=enc)
  def net(self): return self.siren if self.kind == "siren" else self.mlp


class FusedVolSDF(FusedNeRF):
  """Drop-in for the volume-rendering branch of VolSDF (reference src/nerf.py:861-1018) with a View head.  Parameter
  names follow the reference (`scale`, `sdf.underlying.{siren|mlp}.*`, `sdf.refl.mlp.*`).  Secondary lighting,
  normals and occlusion (nerf.py:923-980,1006-1011) are out of scope."""
  kind = "volsdf"

  def __init__(self, sdf_kind: str = "siren", out_features: int = 3, **kwargs):
    kwargs.setdefault("sigmoid_kind", "thin")
    super().__init__(**kwargs)
    if out_features != 3: raise NotImplementedError("out_features != 3")
    self.scale = nn.Parameter(torch.tensor(0.1))
    self.sdf = _SdfBox(_SdfNet(sdf_kind, self.intermediate_size),
                       ViewHead(latent_size=self.intermediate_size, out_features=3, act=self.sigmoid_kind))

  @property
  def refl(self): return self.sdf.refl
  def set_refl(self, refl): self.sdf.refl = refl
  def set_sigmoid(self, kind="thin"):
    if kind not in _lib.FEAT: raise NotImplementedError(f"Unknown sigmoid kind({kind})")
    self.sigmoid_kind = kind; self.sdf.refl.act = kind

  @classmethod
  def from_reference(cls, ref, precision: str = "fp16", keep_weights: bool = True) -> "FusedVolSDF":
    self = cls.__new__(cls)
    FusedNeRF.__init__(self, steps=ref.steps, t_near=ref.t_near, t_far=ref.t_far, intermediate_size=ref.intermediate_size,
                       sigmoid_kind=_sigmoid_name(ref.sdf.refl.act), precision=precision, keep_weights=keep_weights)
    if type(ref.sdf.refl).__name__ not in ("View", "ViewHead"): raise NotImplementedError(f"refl head {type(ref.sdf.refl).__name__}")
    if getattr(ref, "secondary", None) is not None: raise NotImplementedError("VolSDF secondary lighting")
    self.scale, self.sdf = ref.scale, ref.sdf
    return self

  def forward(self, rays: torch.Tensor) -> torch.Tensor:
    out = super().forward(rays)
    # scale_act is the identity (nerf.py:884,1000-1001); read by runner.py:707.  Bypass nn.Module.__setattr__: assigning a
    # Parameter the normal way would register it a second time and add a key to the state_dict.
    object.__setattr__(self, "scale_post_act", self.scale)
    return out

  def _sdf_net(self):
    u = self.sdf.underlying
    if hasattr(u, "siren"): return "siren", u.siren
    if hasattr(u, "mlp"): return "mlp", u.mlp
    raise NotImplementedError(f"sdf network {type(u).__name__}")

  def engine(self) -> RenderEngine:
    key = (self.kind, self._sdf_net()[0], _sigmoid_name(self.sdf.refl.act), self.precision)
    if self._engine is None or self._engine_key != key:
      self._engine, self._engine_key = RenderEngine(self._describe(), self.precision), key
    return self._engine

  def _describe(self) -> ModelDesc:
    kind, net = self._sdf_net()
    freqs = net.enc.basis.shape[1] if kind == "mlp" else 128
    return describe_volsdf(kind, self.intermediate_size, _sigmoid_name(self.sdf.refl.act), freqs)

  def _param_list(self) -> List[torch.Tensor]:
    kind, net = self._sdf_net()
    ps: List[torch.Tensor] = []
    for mlp in (net, self.sdf.refl.mlp):
      for lin in _linears_of(mlp): ps += [lin.weight, lin.bias]
    if kind == "mlp":
      # FourierEncoder computes fourier(x, extra_scale * basis) and runner --inc-fourier-freqs grows extra_scale every step
      # (reference src/neural_blocks.py:50-55, runner.py:826-829): pack the scaled basis
      es = getattr(net.enc, "extra_scale", 1)
      if isinstance(es, torch.Tensor) or float(es) != 1.0:
        key = (float(es), net.enc.basis.data_ptr(), net.enc.basis._version)
        if getattr(self, "_scaled_basis_key", None) != key:
          with torch.no_grad(): self._scaled_basis = (es * net.enc.basis).to(torch.float32).contiguous()
          self._scaled_basis_key = key
        ps.append(self._scaled_basis)
      else: ps.append(net.enc.basis)
    ps.append(self.scale.reshape(1) if self.scale.dim() == 0 else self.scale)
    return ps


class FusedSDF(nn.Module):
  """Drop-in for the surface renderer `sdf.SDF` (reference src/sdf.py:86-156; `--model sdf --sdf-isect-kind sphere | bisect`) with a
  View head, evaluation mode: sphere tracing (march.py:27-47) or bisection (march.py:63-110,147-180) of the SDF network (SIREN or
  Fourier-encoded MLP), then the View head on the hit points;
  rays that miss render black.  Parameter names follow the reference (`underlying.{siren|mlp}.*`, `refl.mlp.*`).  Training of
  the surface model (throughput term, normals: sdf.py:120-125,150-153) is not built: forward under grad raises."""

  def __init__(self, sdf_kind: str = "siren", intermediate_size: int = 64, t_near: float = 0, t_far: float = 1, sigmoid_kind: str = "thin",
               bound_sphere_rad: float = -1.0, precision: str = "fp32", isect: str = "sphere"):
    super().__init__()
    if isect not in ("sphere", "bisect"): raise NotImplementedError(f"intersection kind {isect} (built: sphere, bisect)")
    self.isect_kind, self.jitter = isect, None      # jitter: march.py:86's random.random() draw; None = draw one per call like the reference
    self.underlying = _SdfNet(sdf_kind, intermediate_size)
    self.refl = ViewHead(latent_size=intermediate_size, out_features=3, act=sigmoid_kind)
    self.near, self.far, self.bound_sphere_rad, self.precision = t_near, t_far, bound_sphere_rad, precision
    self._engine: Optional[RenderEngine] = None
    self._engine_key = None

  @classmethod
  def from_reference(cls, ref, precision: str = "fp32") -> "FusedSDF":
    """Adopt a live reference `sdf.SDF` (sphere-march intersection, View head); a `UnitSphere` wrapper becomes `bound_sphere_rad`."""
    self = cls.__new__(cls); nn.Module.__init__(self)
    u, rad = ref.underlying, -1.0
    if type(u).__name__ == "UnitSphere": u, rad = u.inner, float(u.rad)
    if type(ref.refl).__name__ not in ("View", "ViewHead"): raise NotImplementedError(f"refl head {type(ref.refl).__name__}")
    kind = {"sphere_march": "sphere", "bisect": "bisect"}.get(getattr(ref.isect, "__name__", "sphere_march"))
    if kind is None: raise NotImplementedError("intersection kinds built: sphere_march, bisect")
    self.isect_kind, self.jitter = kind, None
    self.underlying, self.refl = u, ref.refl
    self.near, self.far, self.bound_sphere_rad, self.precision = ref.near, ref.far, rad, precision
    self._engine = None; self._engine_key = None
    return self

  @property
  def sdf(self): return self
  @property
  def intermediate_size(self): return self.underlying.intermediate_size
  def __getstate__(self):
    st = self.__dict__.copy(); st["_engine"] = None; st["_engine_key"] = None
    for k in ("pts", "hit", "t", "tput", "best_pos"): st.pop(k, None)
    return st

  def _net(self):
    u = self.underlying
    if hasattr(u, "siren"): return "siren", u.siren
    if hasattr(u, "mlp"): return "mlp", u.mlp
    raise NotImplementedError(f"sdf network {type(u).__name__}")

  def engine(self) -> RenderEngine:
    kind, net = self._net()
    key = (kind, _sigmoid_name(self.refl.act), self.precision)
    if self._engine is None or self._engine_key != key:
      freqs = net.enc.basis.shape[1] if kind == "mlp" else 128
      self._engine, self._engine_key = RenderEngine(describe_volsdf(kind, self.intermediate_size, key[1], freqs), self.precision), key
    return self._engine

  def _param_list(self) -> List[torch.Tensor]:
    kind, net = self._net()
    ps: List[torch.Tensor] = []
    for mlp in (net, self.refl.mlp):
      for lin in _linears_of(mlp): ps += [lin.weight, lin.bias]
    if kind == "mlp": ps.append(net.enc.basis)
    if not hasattr(self, "_beta") or self._beta.device != ps[0].device: self._beta = torch.ones(1, device=ps[0].device)    # the descriptor's Laplace beta: unused by the surface side
    ps.append(self._beta)
    return ps

  def normals(self, pts: torch.Tensor, values=None) -> torch.Tensor:
    """SDF.normals (reference src/sdf.py:112 -> SDFModel.normals, 43-49), no graph: the fused path does not train the surface model."""
    if not pts.is_cuda: raise RuntimeError("FusedSDF.normals needs CUDA points: the fused path has no CPU fallback")
    eng = self.engine(); eng.pack(self._param_list())
    shape = pts.shape
    return eng.sdf_normals(pts.reshape(-1, 3).to(torch.float32).contiguous(), bound_rad=self.bound_sphere_rad).reshape(shape)

  def intersect_w_n(self, r_o: torch.Tensor, r_d: torch.Tensor):
    """SDF.intersect_w_n in eval mode (reference src/sdf.py:114-125: iters 256, eps 5e-5): (pts, hit, tput | None, normals)."""
    B = r_o.shape[:-1]
    flat = torch.cat([r_o, r_d], dim=-1).reshape(-1, 6).to(torch.float32).contiguous()
    eng = self.engine(); eng.pack(self._param_list())
    if getattr(self, "isect_kind", "sphere") == "bisect":
      import random
      u = random.random() if getattr(self, "jitter", None) is None else float(self.jitter)
      _, hit, tput, pts, _ = eng.sdf_bisect(flat, self.near, self.far, iters=256, jitter=u, bound_rad=self.bound_sphere_rad, shade=False)
      tput = tput.reshape(*B, 1)
    else:
      pts, hit, _ = eng.sphere_march(flat, self.near, self.far, iters=256, eps=5e-5, bound_rad=self.bound_sphere_rad); tput = None
    return pts.reshape(*B, 3), hit.reshape(B), tput, eng.sdf_normals(pts, bound_rad=self.bound_sphere_rad).reshape(*B, 3)

  def forward(self, rays: torch.Tensor, with_throughput: bool = True) -> torch.Tensor:
    if not rays.is_cuda: raise RuntimeError("FusedSDF.forward needs CUDA rays: the fused path has no CPU fallback")
    if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
      raise NotImplementedError("training the SDF surface model through the fused path is not built; call under torch.no_grad() / eval()")
    B = rays.shape[:-1]
    flat = rays.reshape(-1, 6).to(torch.float32).contiguous()
    eng = self.engine(); eng.pack(self._param_list())
    iters = 128 if self.training else 192
    if getattr(self, "isect_kind", "sphere") == "bisect":
      import random
      u = random.random() if getattr(self, "jitter", None) is None else float(self.jitter)
      rgb, hit, tput, pts, best = eng.sdf_bisect(flat, self.near, self.far, iters=iters, jitter=u, bound_rad=self.bound_sphere_rad)
      self.pts, self.hit, self.tput, self.best_pos = pts.reshape(*B, 3), hit.reshape(B), tput.reshape(B), best.reshape(*B, 3)
      return rgb.reshape(*B, 3)
    rgb, hit, t, pts = eng.sdf_render(flat, self.near, self.far, iters=iters, bound_rad=self.bound_sphere_rad)
    self.pts, self.hit, self.t = pts.reshape(*B, 3), hit.reshape(B), t.reshape(B)
    return rgb.reshape(*B, 3)


class FusedDynamicNeRF(nn.Module):
  """Drop-in for DynamicNeRF (reference src/nerf.py:1209-1303) over a canonical FusedPlainNeRF:
  `model((rays[B,H,W,6], times[B])) -> rgb[B,H,W,3]`; `spline == 0`: the direct deformation MLP, `spline == n`: n Bezier
  control points from a hash-encoded MLP.  State-dict names follow the reference (`delta_estim.*`, `canonical.*`).
  `refl_latent > 0` is not built."""

  def __init__(self, canonical: FusedPlainNeRF, spline: int = 0, refl_latent: int = 0):
    super().__init__()
    if refl_latent: raise NotImplementedError("refl_latent variant of DynamicNeRF")
    if spline and not 2 <= spline <= 8: raise NotImplementedError("spline points must be in 2..8")
    if canonical.mip_size() or canonical.refl_kind != "view":
      raise NotImplementedError("DynamicNeRF over a Mip / Positional canonical NeRF is not built")
    self.canonical = canonical
    self.spline = spline
    if spline: self.delta_estim = SkipConnParams(38, 1 + 3 * spline, 5, init="xavier", enc=HashParams())
    else: self.delta_estim = SkipConnParams(4, 4, 5, init="xavier")
    nn.init.zeros_(self.delta_estim.out.weight); nn.init.zeros_(self.delta_estim.out.bias)   # zero_last_layer(), nerf.py:1239
    self._engine: Optional[RenderEngine] = None
    self._engine_key = None

  @classmethod
  def from_reference(cls, ref, precision: str = "fp16") -> "FusedDynamicNeRF":
    if getattr(ref, "refl_latent", 0): raise NotImplementedError("refl_latent variant of DynamicNeRF")
    self = cls.__new__(cls); nn.Module.__init__(self)
    self.spline = int(getattr(ref, "spline", 0))
    self.canonical = FusedPlainNeRF.from_reference(ref.canonical, precision=precision)
    self.delta_estim = ref.delta_estim
    self._engine = None; self._engine_key = None
    return self

  @property
  def nerf(self): return self.canonical
  @property
  def refl(self): return self.canonical.refl
  @property
  def intermediate_size(self): return self.canonical.intermediate_size
  def total_latent_size(self): return self.canonical.total_latent_size()
  def set_refl(self, refl): self.canonical.set_refl(refl)
  def set_bg(self, bg): self.canonical.set_bg(bg)
  def __getstate__(self):
    st = self.__dict__.copy(); st["_engine"] = None; st["_engine_key"] = None
    return st

  def engine(self) -> RenderEngine:
    c = self.canonical
    key = (_sigmoid_name(c.refl.act), c.bg, c.precision, self.spline)
    if self._engine is None or self._engine_key != key:
      self._engine = RenderEngine(describe_dyn(c.intermediate_size, key[0], c.bg, self.spline), c.precision); self._engine_key = key
    return self._engine

  def _param_list(self) -> List[torch.Tensor]:
    c = self.canonical
    ps: List[torch.Tensor] = []
    for mlp in (c.first, c.refl.mlp, self.delta_estim):
      for lin in _linears_of(mlp): ps += [lin.weight, lin.bias]
    ps += [e.weight for e in c.first.enc.embs]
    if self.spline: ps += [e.weight for e in self.delta_estim.enc.embs]
    return ps

  def forward(self, rays_t):
    rays, t = rays_t
    c = self.canonical
    if not rays.is_cuda: raise RuntimeError("FusedDynamicNeRF.forward needs CUDA rays: the fused path has no CPU fallback")
    if torch.is_grad_enabled() and self.training and any(p.requires_grad for p in self.parameters()):
      raise NotImplementedError("training DynamicNeRF through the fused path is not built (the deformation needs d/d pts); "
                                "call under torch.no_grad()")
    B = rays.shape[:-1]
    flat = rays.reshape(-1, 6).to(torch.float32).contiguous()
    # training-mode semantics as the reference: jittered ts (nerf.py:1292-1296, perturb = 1 if self.training) and density noise
    # when the CANONICAL NeRF is in training mode (nerf.py:347-348)
    tr = c.training; c.training = self.training
    try: ts, _ = c._sample_ts(rays.device, flat.shape[0], with_noise=False)
    finally: c.training = tr
    noise = torch.randn(flat.shape[0], c.steps, device=rays.device) * c.noise_std if (c.training and c.noise_std > 0) else None
    ray_time = t.to(torch.float32).reshape(-1, 1, 1).expand(B).reshape(-1).contiguous()        # nerf.py:1301
    eng = self.engine(); eng.pack(self._param_list())
    side = {} if getattr(self, "keep_side", True) else None
    bg_rand = torch.rand(flat.shape[0], device=rays.device) if c.bg == "random" else None
    rgb, alpha, weights = eng.render(flat, ts, noise, want_weights=c.keep_weights, ray_time=ray_time, side=side, bg_rand=bg_rand)
    c.ts = self.ts = ts
    if side is not None:
      # what the runner's regularisers / visualisations read after forward (runner.py:523-531,694-700,769,777-781), in the
      # reference's [T,B,H,W,C] layout (transposed views of the ray-major buffers)
      for k, v in side.items(): setattr(self, k, v.reshape(*B, c.steps, v.shape[-1]).movedim(-2, 0))
    if c.keep_weights:
      c.alpha = alpha.reshape(*B, c.steps).movedim(-1, 0); c.weights = weights.reshape(*B, c.steps).movedim(-1, 0)
    return rgb.reshape(*B, 3)


def volumetric_integrate(weights: torch.Tensor, other: torch.Tensor) -> torch.Tensor:
  """The reference's free function (src/nerf.py:79-80: ``sum(weights[..., None] * other, dim=0)``) for the quantities the runner
  integrates over the weights a render kept -- depth (``other = model.nerf.ts[:, None, None, None, None]``, runner.py:511-514,
  894-897), flow and rigidity maps (``other = model.rigid_dp / model.rigidity``, runner.py:521-531,909-914) -- on `nf_integrate`.
  ``weights[T, *B]``, ``other[T, *B, C]`` or ``[T, 1, ..., 1]`` (shared by all rays), C in (1, 3) -> ``[*B, C]``.  The modules keep
  ``weights`` and the side channels as transposed views of ray-major buffers, so nothing is copied for them."""
  if not weights.is_cuda: raise RuntimeError("volumetric_integrate needs CUDA tensors: the fused path has no CPU fallback")
  T = weights.shape[0]; B = weights.shape[1:]
  w = weights.to(torch.float32).movedim(0, -1).reshape(-1, T).contiguous()
  R = w.shape[0]
  if other.shape[0] != T: raise ValueError(f"volumetric_integrate: other has {other.shape[0]} samples, weights {T}")
  if other.numel() == T:
    v = other.to(torch.float32).reshape(T).contiguous(); Cn, stride = 1, 0
  else:
    if tuple(other.shape[1:-1]) != tuple(B): raise ValueError(f"volumetric_integrate: other {tuple(other.shape)} does not match weights {tuple(weights.shape)}")
    Cn = other.shape[-1]
    v = other.to(torch.float32).movedim(0, -2).reshape(R, T * Cn).contiguous(); stride = T * Cn
  if Cn not in (1, 3): raise NotImplementedError("volumetric_integrate: 1 or 3 channels")
  out = torch.empty(R, Cn, dtype=torch.float32, device=weights.device)
  with torch.cuda.device(weights.device):
    rc = _lib.lib().nf_integrate(_ptr(w), _ptr(v), R, T, Cn, stride, _ptr(out), C.c_void_p(torch.cuda.current_stream().cuda_stream))
  _lib.check(rc, "nf_integrate")
  return out.reshape(*B, Cn)


class FusedTinyNeRF(FusedNeRF):
  """Drop-in for TinyNeRF with the intended density semantics (reference src/nerf.py:278-305)."""
  kind = "tiny"

  def __init__(self, out_features: int = 3, **kwargs):
    kwargs.setdefault("sigmoid_kind", "thin")
    super().__init__(**kwargs)
    if out_features != 3: raise NotImplementedError("out_features != 3")
    self.estim = SkipConnParams(3, 1 + out_features, 6, init="xavier")

  def _describe(self) -> ModelDesc: return describe_tiny(self.sigmoid_kind, self.bg)
  def _param_list(self) -> List[torch.Tensor]:
    ps: List[torch.Tensor] = []
    for lin in _linears_of(self.estim): ps += [lin.weight, lin.bias]
    return ps

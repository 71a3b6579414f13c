"""CPU oracle for the NeRF volume-rendering hot path of JulianKnodt/nerf_atlas.

TEST INFRASTRUCTURE ONLY.  Nothing under ``nerf_atlas_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it, and only as the checker or as
the CPU arm being timed -- never as the thing that is shipped.

It is a restatement, in plain ``torch`` CPU tensor ops (the reference's own
arithmetic library; no ``nn.Module`` from the reference is used), of

  * stratified sample generation      reference src/nerf.py:29-55
  * the hash-grid encoder             reference src/neural_blocks.py:92-193
  * the Fourier encoder               reference src/neural_blocks.py:36-55, src/utils.py:10-17
  * SkipConnMLP                       reference src/neural_blocks.py:204-296
  * the View reflectance head         reference src/refl.py:190-207, src/utils.py:247-254
  * the sigmoid family                reference src/utils.py:484-518
  * alpha-from-density + composite    reference src/nerf.py:22-27,60-80
  * PlainNeRF.forward / from_pts      reference src/nerf.py:326-361
  * TinyNeRF.forward (intended)       reference src/nerf.py:292-305
  * VolSDF volume branch              reference src/nerf.py:981-1013, src/utils.py:50-58, src/sdf.py:109-112,250-287
  * Mip-NeRF IPE (cylinder, cone)     reference src/nerf.py:255-261, src/utils.py:22-27,39-48,60-140
  * DynamicNeRF, direct + spline      reference src/nerf.py:1173-1178,1201-1303
  * sample_pdf (restated, dead code)  reference src/nerf.py:1745-1779

Parity pin: the reference has no tests or golden vectors of its own (SURVEY.md
section 8c), so this oracle is pinned against the *reference modules themselves*,
imported in the build container by ``tests/golden/make_golden.py`` (with the
shim list of ``oracle/ref_shim.py``).  The outputs of that run are committed as
``tests/golden/*.npz`` and ``tests/test_oracle_golden.py`` checks this file
against them bit-for-bit on CPU.

Parameters are carried as a flat ``dict[str, Tensor]`` with the reference's own
``state_dict`` names (``first.init.weight`` ...), so a reference checkpoint's
``model.state_dict()`` can be fed to the oracle directly.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Params = Dict[str, Tensor]

# ----------------------------------------------------------------------------
# hash-grid constants (reference src/neural_blocks.py:92-130)
# ----------------------------------------------------------------------------
HASH_PRIMES = (1, 2654435761, 805459861)
HASH_LEVELS = 8
HASH_TABLE = 1 << 16
HASH_FEAT = 4
HASH_LOW = 1 << 4
HASH_HIGH = 1 << 14
# NB operator precedence in the reference: exp((ln hi - ln lo)/levels - 1)
HASH_SCALE = math.exp((math.log(HASH_HIGH) - math.log(HASH_LOW)) / HASH_LEVELS - 1)


def hash_resolutions(levels: int = HASH_LEVELS) -> np.ndarray:
  """fp32 per-level N_l exactly as ``x * N_l`` sees it (python double -> fp32).
  reference src/neural_blocks.py:146-147."""
  return np.array([HASH_LOW * (HASH_SCALE ** i) for i in range(levels)], dtype=np.float64).astype(np.float32)


# ----------------------------------------------------------------------------
# activations
# ----------------------------------------------------------------------------
def leaky_relu(x: Tensor) -> Tensor: return F.leaky_relu(x, 0.01)

ACTS: Dict[str, Callable[[Tensor], Tensor]] = {
  "leaky_relu": leaky_relu,   # nn.LeakyReLU() default slope, neural_blocks.py:214
  "sin": torch.sin,           # refl.py:203
  "relu": F.relu,
  "none": lambda x: x,
}

# reference src/utils.py:484-518
def fat_sigmoid(v, eps: float = 1e-2): return v.sigmoid() * (1 + 2 * eps) - eps
def thin_sigmoid(v, eps: float = 1e-2): return fat_sigmoid(v, -eps) + eps
def cyclic_sigmoid(v, eps: float = -1e-2, period: int = 5):
  return ((v / period).sin() + 1) / 2 * (1 + 2 * eps) - eps
def upshifted_sigmoid(v, eps=1e-2): return v.sigmoid() + eps
def upshifted_softplus(v, eps=1e-2): return F.softplus(v) + eps
def upshifted_relu(v, eps=1e-2): return F.relu(v) + eps

SIGMOIDS: Dict[str, Callable[[Tensor], Tensor]] = {
  "normal": torch.sigmoid,
  "thin": thin_sigmoid,
  "tanh": torch.tanh,
  "cyclic": cyclic_sigmoid,
  "upshifted": upshifted_sigmoid,
  "fat": fat_sigmoid,
  "softmax": lambda v: torch.softmax(v, dim=-1),
  "leaky_relu": F.leaky_relu,
  "relu": F.relu,
  "sin": torch.sin,
  "upshifted_softplus": upshifted_softplus,
  "upshifted_relu": upshifted_relu,
}


# ----------------------------------------------------------------------------
# a-1  sample generation   (reference src/nerf.py:29-55)
# ----------------------------------------------------------------------------
def compute_ts(near: float, far: float, steps: int, rand: Optional[Tensor] = None,
               perturb: float = 1.0, lindisp: bool = False) -> Tensor:
  """``ts[T]`` shared by every ray.  ``rand`` is the explicit ``rand_like(lower)``
  draw of nerf.py:45 (None = eval mode, no jitter)."""
  if lindisp:
    t_vals = torch.linspace(0, 1, steps, dtype=torch.float32)
    ts = 1 / (1 / max(near, 1e-10) * (1 - t_vals) + 1 / far * t_vals)
  else:
    ts = torch.linspace(near, far, steps=steps, dtype=torch.float32)
  if rand is not None:
    mids = 0.5 * (ts[:-1] + ts[1:])
    lower = torch.cat([mids, ts[-1:]])
    upper = torch.cat([ts[:1], mids])
    ts = lower + (upper - lower) * (rand * perturb)
  return ts


def compute_pts(rays: Tensor, ts: Tensor):
  """``pts[T,...,3] = r_o + ts (x) r_d`` -- a rounded product then a rounded add
  (nerf.py:54), never an FMA."""
  r_o, r_d = rays.split([3, 3], dim=-1)
  pts = r_o.unsqueeze(0) + torch.tensordot(ts, r_d, dims=0)
  return pts, r_o, r_d


# ----------------------------------------------------------------------------
# a-2  hash-grid encoder   (reference src/neural_blocks.py:139-193)
# ----------------------------------------------------------------------------
# corner order of neural_blocks.py:155-165: (bx,by,bz) with z the fastest bit
HASH_CORNERS = ((0, 0, 0), (0, 0, 1), (0, 1, 0), (0, 1, 1), (1, 0, 0), (1, 0, 1), (1, 1, 0), (1, 1, 1))


def hash_indices(x: Tensor, level: int) -> Tensor:
  """Table row per corner, ``[8, N]`` int64 in [0, 65536).  Same arithmetic as the
  reference: int64 multiply, xor, python-style ``% emb_size``."""
  N_l = float(HASH_LOW * (HASH_SCALE ** level))
  v_l = x * N_l
  l = v_l.floor().long()
  primes = torch.tensor(HASH_PRIMES, dtype=torch.int64, device=x.device)
  out = []
  for (bx, by, bz) in HASH_CORNERS:
    c = l + torch.tensor([bx, by, bz], dtype=torch.int64, device=x.device)
    v = c * primes
    h = v[..., 0].bitwise_xor(v[..., 1]).bitwise_xor(v[..., 2])
    out.append(h % HASH_TABLE)
  return torch.stack(out, dim=0)


def hash_indices_u32(x: np.ndarray, level: int) -> np.ndarray:
  """The uint32 low-16-bit identity the CUDA kernel uses:
  ``((ix)*1 ^ (iy)*(P1 & 0xFFFF) ^ (iz)*(P2 & 0xFFFF)) & 0xFFFF`` on two's
  complement int32 cell coordinates.  numpy; used to prove the identity against
  :func:`hash_indices`."""
  N_l = hash_resolutions()[level]
  v = (x.astype(np.float32) * N_l).astype(np.float32)
  l = np.floor(v).astype(np.int32)
  out = []
  for (bx, by, bz) in HASH_CORNERS:
    ix = (l[..., 0] + bx).astype(np.uint32)
    iy = (l[..., 1] + by).astype(np.uint32)
    iz = (l[..., 2] + bz).astype(np.uint32)
    h = (ix * np.uint32(1)) ^ (iy * np.uint32(HASH_PRIMES[1] & 0xFFFF)) ^ (iz * np.uint32(HASH_PRIMES[2] & 0xFFFF))
    out.append((h & np.uint32(0xFFFF)).astype(np.int64))
  return np.stack(out, axis=0)


def hash_encode(x: Tensor, tables: Tensor) -> Tensor:
  """``[N,3] -> [N, 3 + 8*4]``; ``tables`` is ``[8, 65536, 4]``."""
  out = []
  for i in range(HASH_LEVELS):
    N_l = float(HASH_LOW * (HASH_SCALE ** i))
    v_l = x * N_l
    l = v_l.floor().long()
    idx = hash_indices(x, i)                                  # [8, N]
    embs = tables[i][idx]                                     # [8, N, 4]
    ws = v_l - l
    wx, wy, wz = ws.split([1, 1, 1], dim=-1)
    iws = 1 - ws
    iwx, iwy, iwz = iws.split([1, 1, 1], dim=-1)
    weights = torch.stack([
      iwx * iwy * iwz, iwx * iwy * wz, iwx * wy * iwz, iwx * wy * wz,
      wx * iwy * iwz, wx * iwy * wz, wx * wy * iwz, wx * wy * wz,
    ], dim=0)                                                 # [8, N, 1]
    out.append((embs * weights).sum(dim=0))
  out = torch.cat(out, dim=-1)
  return torch.cat([x, out], dim=-1)


def hash_tables(params: Params, prefix: str) -> Tensor:
  return torch.stack([params[f"{prefix}.embs.{i}.weight"] for i in range(HASH_LEVELS)], dim=0)


# ----------------------------------------------------------------------------
# a-3  Fourier encoder  (reference src/neural_blocks.py:36-55, src/utils.py:14-17)
# ----------------------------------------------------------------------------
def fourier_encode(x: Tensor, basis: Tensor) -> Tensor:
  mapped = x @ basis
  return torch.cat([mapped.sin(), mapped.cos()], dim=-1)


# ----------------------------------------------------------------------------
# a-5  SkipConnMLP  (reference src/neural_blocks.py:279-296)
# ----------------------------------------------------------------------------
def _q(x: Tensor, quant: Optional[torch.dtype]) -> Tensor:
  """Emulate tensor-core operand rounding (round GEMM inputs to ``quant``, fp32
  accumulate).  ``None`` = the reference's plain fp32."""
  return x if quant is None else x.to(quant).to(torch.float32)


def skip_mlp(x0: Tensor, params: Params, prefix: str, act: str = "leaky_relu",
             skip: int = 3, quant: Optional[torch.dtype] = None) -> Tensor:
  """``x0`` is the already-assembled ``[N, dim_p]`` input ``[p, enc(p), latent]``.
  The activation PRECEDES every hidden Linear and is also applied to the
  re-concatenated raw inputs (neural_blocks.py:290-293)."""
  a = ACTS[act]
  lin = lambda x, n: F.linear(_q(x, quant), _q(params[f"{prefix}.{n}.weight"], quant), params[f"{prefix}.{n}.bias"])
  n_layers = 0
  while f"{prefix}.layers.{n_layers}.weight" in params: n_layers += 1
  x = lin(x0, "init")
  for i in range(n_layers):
    if i != n_layers - 1 and (i % skip) == 0:
      x = torch.cat([x, x0], dim=-1)
    x = lin(a(x), f"layers.{i}")
  return lin(a(x), "out")


# ----------------------------------------------------------------------------
# a-10  view direction -> (elevation, azimuth)   (reference src/utils.py:247-254)
# ----------------------------------------------------------------------------
def dir_to_elev_azim(direc: Tensor) -> Tensor:
  lim = 1 - 1e-6
  x, y, z = F.normalize(direc, dim=-1).clamp(min=-lim, max=lim).split([1, 1, 1], dim=-1)
  return torch.cat([z.acos(), torch.atan2(y, x)], dim=-1)


# ----------------------------------------------------------------------------
# a-8 / a-9  density -> alpha -> weights -> integrate  (reference src/nerf.py:22-27,60-80)
# ----------------------------------------------------------------------------
def cumuprod_exclusive(t: Tensor) -> Tensor:
  cp = torch.cumprod(t, dim=0)
  cp = torch.roll(cp, 1, dims=0)
  cp[0, ...] = 1.0
  return cp


def alpha_from_density(density: Tensor, ts: Tensor, r_d: Tensor, softplus: bool = True):
  """``density[T,...]``, shared ``ts[T]``, ``r_d[...,3]`` (NOT normalised)."""
  sigma_a = F.softplus(density - 1) if softplus else F.relu(density)
  end_val = torch.full_like(ts[..., :1], 1e10)
  dists = torch.cat([ts[..., 1:] - ts[..., :-1], end_val], dim=-1).clamp(min=1e-5)
  while len(dists.shape) < density.dim(): dists = dists[..., None]
  dists = dists * torch.linalg.norm(r_d, dim=-1)
  alpha = 1 - torch.exp(-sigma_a * dists)
  weights = alpha * cumuprod_exclusive(1.0 - alpha + 1e-10)
  return alpha, weights


def alpha_from_density_per_ray(density: Tensor, ts: Tensor, r_d: Tensor, softplus: bool = True):
  """Restated compositor for PER-RAY sample positions ``ts[T,...]`` (coarse+fine,
  SURVEY.md a-7): same formula, delta taken along dim 0."""
  sigma_a = F.softplus(density - 1) if softplus else F.relu(density)
  end_val = torch.full_like(ts[:1], 1e10)
  dists = torch.cat([ts[1:] - ts[:-1], end_val], dim=0).clamp(min=1e-5)
  dists = dists * torch.linalg.norm(r_d, dim=-1)
  alpha = 1 - torch.exp(-sigma_a * dists)
  weights = alpha * cumuprod_exclusive(1.0 - alpha + 1e-10)
  return alpha, weights


def volumetric_integrate(weights: Tensor, other: Tensor) -> Tensor:
  return torch.sum(weights[..., None] * other, dim=0)


def sky(bg: str, weights: Tensor):
  """reference src/nerf.py:96-109."""
  if bg == "black": return 0
  if bg == "white": return 1 - weights[:-1].sum(dim=0).unsqueeze(-1)
  raise NotImplementedError(bg)


# ----------------------------------------------------------------------------
# a-4  Mip-NeRF integrated positional encoding
#      (reference src/nerf.py:255-261, src/utils.py:22-27,39-48,60-140)
# ----------------------------------------------------------------------------
MIP_DEG = 16                     # min_deg 0, max_deg 16 (utils.py:104-111,127-134) -> 6 * 16 = 96 features

def radii_x(r_d: Tensor) -> Tensor:
  """utils.py:77-81.  ``r_d[B,H,W,3]`` -> ``[B,H,W,1]``: distance to the next ray along H (the last row repeats row
  H-3 of the differences: ``dx[:, -2:-1, :]`` of an ``[B,H-1,W]`` tensor), times 2/sqrt(12)."""
  dx = (r_d[..., :-1, :, :] - r_d[..., 1:, :, :]).square().sum(dim=-1).sqrt()
  dx = torch.cat([dx, dx[:, -2:-1, :]], dim=-2)
  return dx[..., None] * 2 / math.sqrt(12)

def expected_sin(x: Tensor, x_var: Tensor) -> Tensor:
  """utils.py:22-27 (only the mean is used by the encoder)."""
  return (-0.5 * x_var).exp() * x.sin()

def _gaussian_moments(kind: str, t0: Tensor, t1: Tensor, rad: Tensor):
  """(t_mean[T], t_var[T], r_var[B,H,W,1 or T]) of one ray segment: cylinder utils.py:95-101, cone utils.py:83-93."""
  if kind == "cylinder":
    return (t1 + t0) / 2, (t1 - t0).square() / 12, rad * rad / 4
  mu = (t1 + t0) / 2; hw = (t1 - t0) / 2
  mu2 = mu * mu; hw2 = hw * hw; hw4 = hw2 * hw2
  t_mean = mu + (2 * mu * hw2) / (3 * mu2 + hw2)
  t_var = hw / 3 - (4 / 15) * ((hw4 * (12 * mu2 - hw2)) / (3 * mu2 + hw2).square())
  r_var = rad * rad * (mu2 / 4 + (5 / 12) * hw2 - 4 / 15 * (hw4) / (3 * mu2 + hw2))
  return t_mean, t_var, r_var

def mip_encoding(r_o: Tensor, r_d: Tensor, ts: Tensor, kind: str = "cylinder", layout: str = "reference") -> Tensor:
  """``r_o, r_d [B,H,W,3]``, shared ``ts[T]`` -> IPE latent ``[T,B,H,W,96]`` (appended to BOTH MLP inputs, nerf.py:342,357).

  layout == "reference": what the reference computes (CommonNeRF.mip_encoding + CylinderGaussian), bug for bug --
    the last segment ends at 1e10 (nerf.py:258) and ``lift_gaussian`` (utils.py:60-73) moves the xyz axis of the covariance
    to the front instead of the sample axis, after which ``integrated_pos_enc_diag`` (utils.py:42-44) reinterprets the
    ``[3,B,H,W,16,T]`` variance as ``[T,B,H,W,48]``: the variance of feature (t',ray',c') is the element with the same FLAT
    index of the ``[xyz, ray, k, t]`` array (``mip_reference_var_source``).  Only the cylinder runs in the reference: the
    cone's ``hw**4`` overflows on the 1e10 segment and turns the whole image into NaN.
  layout == "intended": the per-sample encoder the code was meant to be: variance laid out ``[T,B,H,W,(k,xyz)]`` and the
    last segment capped to ``t[T-1] + (t[T-1] - t[T-2])``.  Restatement only (no reference run can pin it)."""
  if layout == "reference":
    end_val = torch.tensor([1e10], dtype=ts.dtype)
  else:
    end_val = ts[-1:] + (ts[-1:] - ts[-2:-1])
  tse = torch.cat([ts, end_val], dim=-1)
  t0, t1 = tse[..., :-1], tse[..., 1:]
  rad = radii_x(r_d)
  t_mean, t_var, r_var = _gaussian_moments(kind, t0, t1, rad)
  # lift_gaussian, utils.py:60-73
  mean = r_d[..., None] * t_mean[..., None, :]                                   # [B,H,W,3,T]
  magn_sq = r_d.square().sum(dim=-1, keepdim=True).clamp(min=1e-10)
  outer_diag = r_d.square()
  null_outer_diag = 1 - outer_diag / magn_sq
  t_cov_diag = t_var[..., None] * outer_diag[..., None, :]                        # [B,H,W,T,3]
  xy_cov_diag = r_var[..., None] * null_outer_diag[..., None, :]
  cov_diag = t_cov_diag + xy_cov_diag
  mean = mean.movedim(-1, 0) + r_o                                                # [T,B,H,W,3]
  cov = cov_diag.movedim(-1, 0) if layout == "reference" else cov_diag.movedim(-2, 0)
  # integrated_pos_enc_diag, utils.py:39-48
  scales = torch.exp2(torch.arange(0, MIP_DEG, dtype=mean.dtype))
  out_shape = mean.shape[:-1] + (-1,)
  y = (mean[..., None, :] * scales[..., None]).reshape(out_shape)
  y_var = (cov[..., None, :] * scales[..., None].square()).reshape(out_shape)
  return expected_sin(torch.cat([y, y + 0.5 * math.pi], dim=-1), torch.cat([y_var, y_var], dim=-1))

def mip_reference_var_source(t: int, ray: int, c: int, T: int, R: int):
  """The (xyz, ray, k, t) whose covariance the REFERENCE layout puts under feature ``c`` (< 48) of sample ``t`` of ray
  ``ray`` (rays flattened over [B,H,W]; R of them): same flat index in ``[3,R,16,T]`` as in ``[T,R,48]``."""
  flat = (t * R + ray) * 48 + c
  ts_ = flat % T; flat //= T
  k = flat % MIP_DEG; flat //= MIP_DEG
  r = flat % R; x = flat // R
  return x, r, k, ts_


# ----------------------------------------------------------------------------
# a-6  PlainNeRF  (reference src/nerf.py:326-361) with the View head of refl.py:190-207
# ----------------------------------------------------------------------------
def plain_from_pts(params: Params, pts: Tensor, ts: Tensor, r_o: Tensor, r_d: Tensor, *,
                   sigmoid: str = "upshifted", bg: str = "black",
                   density_noise: Optional[Tensor] = None,
                   quant: Optional[torch.dtype] = None, per_ray_ts: bool = False,
                   pts_encode: Optional[Tensor] = None, mip_latent: Optional[Tensor] = None) -> Dict[str, Tensor]:
  """Returns every stage so tests can localise a mismatch.  ``mip_latent[T,...,96]`` (``mip_encoding``) is appended to
  both MLP inputs (nerf.py:340-358)."""
  T = pts.shape[0]
  batches = pts.shape[:-1]
  p = pts.reshape(-1, 3)
  tables = hash_tables(params, "first.enc")
  enc = hash_encode(p, tables)                               # [N,35] = [p, feats]
  x0 = torch.cat([p, enc], dim=-1)                           # [N,38] = [p, p, feats]
  if mip_latent is not None: x0 = torch.cat([x0, mip_latent.reshape(x0.shape[0], -1)], dim=-1)
  first_out = skip_mlp(x0, params, "first", "leaky_relu", quant=quant).reshape(batches + (-1,))
  density = first_out[..., 0]
  if density_noise is not None: density = density + density_noise
  intermediate = first_out[..., 1:]
  view = r_d.unsqueeze(0).expand_as(pts)
  elaz = dir_to_elev_azim(view)
  lat = [intermediate] if mip_latent is None else [mip_latent, intermediate]
  if "refl.pos.init.weight" in params:
    # refl.PosLinearView (refl.py:281-290): [pos, im] = act(pos_mlp([p, enc'(p), latent])); linear = sigmoid(view_mlp([p,
    # normalize(view), latent, im])) / 2 + 1/2; rgb = linear * pos   (the feature activation is part of the head: no act afterwards)
    enc2 = hash_encode(p, hash_tables(params, "refl.pos.enc")).reshape(batches + (-1,))
    x0p = torch.cat([pts, enc2] + lat, dim=-1)
    pos_out = SIGMOIDS[sigmoid](skip_mlp(x0p.reshape(-1, x0p.shape[-1]), params, "refl.pos", "leaky_relu", quant=quant)).reshape(batches + (-1,))
    pos, im = pos_out[..., :3], pos_out[..., 3:]
    x0v = torch.cat([pts, F.normalize(view, dim=-1)] + lat + [im], dim=-1)
    lin_ = skip_mlp(x0v.reshape(-1, x0v.shape[-1]), params, "refl.view", "sin", quant=quant).reshape(batches + (-1,)).sigmoid()
    rgb = (lin_ / 2 + 0.5) * pos
    if per_ray_ts: alpha, weights = alpha_from_density_per_ray(density, ts, r_d)
    else: alpha, weights = alpha_from_density(density, ts, r_d)
    out = volumetric_integrate(weights, rgb) + sky(bg, weights)
    return dict(out=out, alpha=alpha, weights=weights, rgb=rgb, density=density, first_out=first_out, hash_feats=enc[:, 3:], elaz=elaz[0])
  if "refl.mlp.enc.embs.0.weight" in params:
    # refl.Positional (refl.py:230-245): view independent, [p, enc'(p) = [p, feats'], latent], LeakyReLU, its own hash tables
    enc2 = hash_encode(p, hash_tables(params, "refl.mlp.enc")).reshape(batches + (-1,))
    x0r = torch.cat([pts, enc2] + lat, dim=-1)
    head_act = "leaky_relu"
  else:
    x0r = torch.cat([pts, elaz] + lat, dim=-1)                # refl.View (refl.py:205-207)
    head_act = "sin"
  x0r = x0r.reshape(-1, x0r.shape[-1])
  rgb_raw = skip_mlp(x0r, params, "refl.mlp", head_act, quant=quant).reshape(batches + (-1,))
  rgb = SIGMOIDS[sigmoid](rgb_raw)
  if per_ray_ts: alpha, weights = alpha_from_density_per_ray(density, ts, r_d)
  else: alpha, weights = alpha_from_density(density, ts, r_d)
  out = volumetric_integrate(weights, rgb) + sky(bg, weights)
  return dict(out=out, alpha=alpha, weights=weights, rgb=rgb, density=density,
              first_out=first_out, hash_feats=enc[:, 3:], elaz=elaz[0])


def plain_forward(params: Params, rays: Tensor, ts: Tensor, *, mip: Optional[str] = None, mip_layout: str = "reference",
                  **kw) -> Dict[str, Tensor]:
  """``mip`` in (None, "cylinder", "cone") needs ``rays[B,H,W,6]`` (the radii difference neighbouring rows)."""
  pts, r_o, r_d = compute_pts(rays, ts)
  if mip is not None: kw["mip_latent"] = mip_encoding(r_o, r_d, ts, mip, mip_layout)
  res = plain_from_pts(params, pts, ts, r_o, r_d, **kw)
  res["pts"] = pts
  return res


# ----------------------------------------------------------------------------
# a-13  TinyNeRF, intended semantics  (reference src/nerf.py:292-305; the
# reference's own forward mis-broadcasts a [...,1] density -- SURVEY.md a-13)
# ----------------------------------------------------------------------------
def tiny_forward(params: Params, rays: Tensor, ts: Tensor, *, sigmoid: str = "upshifted",
                 bg: str = "black", quant: Optional[torch.dtype] = None) -> Dict[str, Tensor]:
  pts, r_o, r_d = compute_pts(rays, ts)
  batches = pts.shape[:-1]
  o = skip_mlp(pts.reshape(-1, 3), params, "estim", "leaky_relu", quant=quant).reshape(batches + (-1,))
  density, feats = o[..., 0], o[..., 1:]
  alpha, weights = alpha_from_density(density, ts, r_d)
  rgb = SIGMOIDS[sigmoid](feats)
  out = volumetric_integrate(weights, rgb) + sky(bg, weights)
  return dict(out=out, alpha=alpha, weights=weights, rgb=rgb, density=density, pts=pts)


# ----------------------------------------------------------------------------
# a-11  VolSDF volume branch  (reference src/nerf.py:981-1013, src/utils.py:50-58)
# ----------------------------------------------------------------------------
def laplace_cdf(sdf_vals: Tensor, beta: Tensor) -> Tensor:
  scaled = sdf_vals / beta
  return torch.where(scaled <= 0, scaled.clamp(max=0).exp() / 2, 1 - scaled.clamp(min=0).neg().exp() / 2)


def volsdf_forward(params: Params, rays: Tensor, ts: Tensor, *, sdf_kind: str = "siren", sigmoid: str = "upshifted",
                   quant: Optional[torch.dtype] = None) -> Dict[str, Tensor]:
  """VolSDF.forward / from_pts, volume-rendering branch with a View head (reference src/nerf.py:981-1013; SDF
  networks src/sdf.py:250-258 `mlp`, 278-287 `siren`; SDF.from_pts src/sdf.py:109-112)."""
  pts, r_o, r_d = compute_pts(rays, ts)
  batches = pts.shape[:-1]
  p = pts.reshape(-1, 3)
  if sdf_kind == "siren":
    raw = skip_mlp(p, params, "sdf.underlying.siren", "sin", quant=quant)
  elif sdf_kind == "mlp":
    x0 = torch.cat([p, fourier_encode(p, params["sdf.underlying.mlp.enc.basis"])], dim=-1)
    raw = skip_mlp(x0, params, "sdf.underlying.mlp", "leaky_relu", quant=quant)
  else: raise NotImplementedError(sdf_kind)
  raw = raw.reshape(batches + (-1,))
  sdf_vals, latent = raw[..., 0], raw[..., 1:]
  scale = params["scale"]
  density = 1 / scale * laplace_cdf(-sdf_vals, scale)
  alpha, weights = alpha_from_density(density, ts, r_d, softplus=False)
  view = r_d.unsqueeze(0).expand_as(pts)
  elaz = dir_to_elev_azim(view)
  x0r = torch.cat([pts, elaz, latent], dim=-1).reshape(-1, 5 + latent.shape[-1])
  rgb = SIGMOIDS[sigmoid](skip_mlp(x0r, params, "sdf.refl.mlp", "sin", quant=quant).reshape(batches + (-1,)))
  out = volumetric_integrate(weights, rgb)
  return dict(out=out, alpha=alpha, weights=weights, rgb=rgb, sdf=sdf_vals, pts=pts)


# ----------------------------------------------------------------------------
# f-4  SDF surface side  (reference src/march.py:27-47, src/sdf.py:66-83,137-156)
# ----------------------------------------------------------------------------
def sdf_net(params: Params, pts: Tensor, sdf_kind: str = "siren", prefix: str = "underlying", bound_rad: float = -1.0,
            quant: Optional[torch.dtype] = None) -> Tensor:
  """[sdf, latent(I)] of the SDF network at pts[N,3] (sdf.py:250-287), optionally intersected with a sphere (UnitSphere, 66-83)."""
  sph = torch.linalg.norm(pts, dim=-1, ord=2) - bound_rad if bound_rad > 0 else None        # (computed first, as sdf.py:78 does: autograd's sums follow)
  if sdf_kind == "siren": raw = skip_mlp(pts, params, f"{prefix}.siren", "sin", quant=quant)
  else: raw = skip_mlp(torch.cat([pts, fourier_encode(pts, params[f"{prefix}.mlp.enc.basis"])], dim=-1), params, f"{prefix}.mlp", "leaky_relu", quant=quant)
  if bound_rad > 0:
    raw = torch.cat([torch.maximum(raw[..., 0], sph).unsqueeze(-1), raw[..., 1:]], dim=-1)
  return raw


def sdf_normals(params: Params, pts: Tensor, *, sdf_kind: str = "siren", bound_rad: float = -1.0, prefix: str = "underlying") -> Tensor:
  """SDFModel.normals (reference src/sdf.py:43-49) through utils.autograd (utils.py:266-277): grad_outputs = ones over EVERY output
  channel of the SDF network, i.e. the gradient of sdf + sum(latent) with respect to the point."""
  with torch.enable_grad():
    p = pts.detach().clone().requires_grad_(True)
    values = sdf_net(params, p, sdf_kind, prefix, bound_rad)
    grad, = torch.autograd.grad(inputs=p, outputs=values, grad_outputs=torch.ones_like(values))
  return grad


def sphere_march(params: Params, r_o: Tensor, r_d: Tensor, *, sdf_kind: str = "siren", iters: int = 32, eps: float = 1e-3,
                 near: float = 0, far: float = 1, bound_rad: float = -1.0, prefix: str = "underlying", quant=None):
  """march.sphere_march (reference src/march.py:27-47) on flat rays [R,3]: (pts, hits, t)."""
  hits = torch.zeros(r_o.shape[0], dtype=torch.bool)
  rem = torch.ones_like(hits)
  curr_dist = torch.full((r_o.shape[0],), float(near))
  for _ in range(iters):
    if not bool(rem.any()): break
    curr = r_o[rem] + r_d[rem] * curr_dist[rem, None]
    dist = sdf_net(params, curr, sdf_kind, prefix, bound_rad, quant)[..., 0]
    hits[rem] |= (dist < eps) & (curr_dist[rem] <= far)
    curr_dist[rem] += dist
    rem[hits | (curr_dist > far)] = False
  return r_o + r_d * curr_dist[:, None], hits, curr_dist


def throughput_with_sign_change(params: Params, r_o: Tensor, r_d: Tensor, *, near: float, far: float, batch_size: int = 128, jitter: float = 0.0,
                                sdf_kind: str = "siren", bound_rad: float = -1.0, prefix: str = "underlying", quant=None):
  """march.throughput_with_sign_change (reference src/march.py:78-110) on flat rays [R,3]: the smallest SDF value of
  batch_size + 1 equidistant samples and the sample indices around the first sign change, bug for bug: the first sample is taken at
  `r_o + near` (the scalar added to every coordinate, march.py:90), and the indices become distances WITHOUT the near offset
  (march.py:105-106; -1, "no crossing", becomes -step).  `jitter` = the reference's `random.random()` draw (march.py:86).
  Returns (tput[R], best_pos[R,3], last_pos[R,1], first_neg[R,1], idxs[R], first_neg_idx[R])."""
  f = lambda p: sdf_net(params, p, sdf_kind, prefix, bound_rad, quant)
  max_t = far - near + jitter * (2 / batch_size)
  step = max_t / batch_size
  sd = f(r_o + near)[..., 0]
  curr_min = sd
  idxs = torch.zeros_like(sd, dtype=torch.long)
  last_pos = torch.full_like(sd, -1, dtype=torch.long)
  first_neg = torch.full_like(sd, -1, dtype=torch.long)
  for i in range(batch_size):
    t = near + step * (i + 1)
    sd = f(r_o + t * r_d)[..., 0]
    idxs = torch.where(sd < curr_min, i + 1, idxs)
    curr_min = torch.minimum(curr_min, sd)
    mask = (first_neg == -1) & (sd < 0)
    last_pos = torch.where(mask, i, last_pos)
    first_neg = torch.where(mask, i + 1, first_neg)
  best_pos = r_o + (near + idxs.unsqueeze(-1) * step) * r_d
  val = f(best_pos)
  return val[..., 0], best_pos, last_pos.unsqueeze(-1) * step, first_neg.unsqueeze(-1) * step, idxs, first_neg


def bisection(params: Params, r_o: Tensor, r_d: Tensor, near: Tensor, far: Tensor, *, iters: int = 32, eps: float = 1e-6,
              sdf_kind: str = "siren", bound_rad: float = -1.0, prefix: str = "underlying", quant=None) -> Tensor:
  """march.bisection (reference src/march.py:147-180): near / far [R,1] bracket the first sign change; rays without a bracket
  (sdf(low) <= 0, sdf(high) >= 0 or an empty interval) keep the midpoint of what they were given."""
  f = lambda p: sdf_net(params, p, sdf_kind, prefix, bound_rad, quant)
  low, high = near.clone(), far.clone()
  sdf_low = f(r_o + low * r_d)[..., 0, None]
  sdf_high = f(r_o + high * r_d)[..., 0, None]
  todo = ((high - low) > eps) & (sdf_low > 0) & (sdf_high < 0) & (high > low)
  z_pred = (low + high) / 2
  for _ in range(iters):
    if not bool(todo.any()): break
    sdf_mid = f(r_o + z_pred * r_d)[..., 0, None]
    low_mask = (sdf_mid > 0) & todo
    low[low_mask] = z_pred[low_mask]; sdf_low[low_mask] = sdf_mid[low_mask]
    high_mask = (sdf_mid < 0) & todo
    high[high_mask] = z_pred[high_mask]; sdf_high[high_mask] = sdf_mid[high_mask]
    z_pred = (low + high) / 2
    todo = todo & ((high - low) > eps) & (sdf_low > 0) & (sdf_high < 0) & (high > low)
  return r_o + z_pred * r_d


def bisect(params: Params, r_o: Tensor, r_d: Tensor, *, iters: int = 128, near: float = 0, far: float = 1, jitter: float = 0.0,
           sdf_kind: str = "siren", bound_rad: float = -1.0, quant=None):
  """march.bisect (reference src/march.py:63-75; `--sdf-isect-kind bisect`): (pts, hits, best_pos, tput)."""
  tput, best_pos, last_pos, first_neg, _, _ = throughput_with_sign_change(params, r_o, r_d, near=near, far=far, batch_size=iters, jitter=jitter,
                                                                          sdf_kind=sdf_kind, bound_rad=bound_rad, quant=quant)
  pts = bisection(params, r_o, r_d, last_pos, first_neg, iters=min(32, iters), sdf_kind=sdf_kind, bound_rad=bound_rad, quant=quant)
  return pts, tput < 0, best_pos, tput


def sdf_forward(params: Params, rays: Tensor, *, sdf_kind: str = "siren", near: float = 0, far: float = 1, iters: int = 192,
                sigmoid: str = "upshifted", bound_rad: float = -1.0, quant=None, isect: str = "sphere", jitter: float = 0.0) -> Dict[str, Tensor]:
  """SDF.forward in eval mode (reference src/sdf.py:137-156): rgb[hit] = act(View([pts, elaz(r_d), latent])), black elsewhere.
  isect = "sphere" (march.sphere_march) or "bisect" (march.bisect; `t` is then absent and `tput`, `best_pos` are returned).
  (march.secant, src/march.py:50-60,113-143, is not restated: the reference's own run of it on the golden's network trips its
  `assert z_pred.isfinite()` -- infinite brackets from the -1 "no crossing" index -- so there is nothing to pin it to.)"""
  r_o, r_d = rays.reshape(-1, 6).split([3, 3], dim=-1)
  if isect == "bisect":
    pts, hit, best_pos, tput = bisect(params, r_o, r_d, iters=iters, near=near, far=far, jitter=jitter, sdf_kind=sdf_kind, bound_rad=bound_rad, quant=quant)
    out = torch.zeros_like(r_d)
    if bool(hit.any()):
      latent = sdf_net(params, pts[hit], sdf_kind, "underlying", bound_rad, quant)[..., 1:]
      x0 = torch.cat([pts[hit], dir_to_elev_azim(r_d[hit]), latent], dim=-1)
      out[hit] = SIGMOIDS[sigmoid](skip_mlp(x0, params, "refl.mlp", "sin", quant=quant))
    B = rays.shape[:-1]
    return dict(out=out.reshape(B + (3,)), hit=hit.reshape(B), pts=pts.reshape(B + (3,)), tput=tput.reshape(B), best_pos=best_pos.reshape(B + (3,)))
  pts, hit, t = sphere_march(params, r_o, r_d, sdf_kind=sdf_kind, iters=iters, near=near, far=far, bound_rad=bound_rad, quant=quant)
  out = torch.zeros_like(r_d)
  if bool(hit.any()):
    latent = sdf_net(params, pts[hit], sdf_kind, "underlying", bound_rad, quant)[..., 1:]
    x0 = torch.cat([pts[hit], dir_to_elev_azim(r_d[hit]), latent], dim=-1)
    out[hit] = SIGMOIDS[sigmoid](skip_mlp(x0, params, "refl.mlp", "sin", quant=quant))
  return dict(out=out.reshape(rays.shape[:-1] + (3,)), hit=hit.reshape(rays.shape[:-1]), t=t.reshape(rays.shape[:-1]), pts=pts.reshape(rays.shape[:-1] + (3,)))


# ----------------------------------------------------------------------------
# a-12  DynamicNeRF, direct deformation MLP  (reference src/nerf.py:1209-1303)
# ----------------------------------------------------------------------------
def dnerf_direct_forward(params: Params, rays: Tensor, times: Tensor, ts: Tensor, *, sigmoid: str = "upshifted",
                         bg: str = "black", quant: Optional[torch.dtype] = None) -> Dict[str, Tensor]:
  """rays[B,H,W,6], times[B].  direct_predict (nerf.py:1261-1266) splits the MLP output [1,3] as (dp, rigidity) --
  the NAMES are swapped w.r.t. the comment at nerf.py:1231, so `dp` is one channel broadcast over xyz and the rigidity
  mask has three -- and reads `self.dp`, which the reference never sets (the golden run patches that in).
  Parameter names: `delta_estim.*`, `canonical.first.*`, `canonical.refl.mlp.*`."""
  pts, r_o, r_d = compute_pts(rays, ts)
  t = times[None, :, None, None, None].expand(*pts.shape[:-1], 1)
  xt = torch.cat([pts, t], dim=-1)
  o = skip_mlp(xt.reshape(-1, 4), params, "delta_estim", "leaky_relu", quant=quant).reshape(pts.shape[:-1] + (4,))
  dp, rigidity = o[..., :1], o[..., 1:4]
  rigid_dp = dp * (rigidity / 2).sigmoid()
  canon = {k[len("canonical."):]: v for k, v in params.items() if k.startswith("canonical.")}
  res = plain_from_pts(canon, pts + rigid_dp, ts, r_o, r_d, sigmoid=sigmoid, bg=bg, quant=quant)
  res["rigid_dp"] = rigid_dp; res["pts"] = pts
  return res


def de_casteljau(coeffs: Tensor, t: Tensor, N: int) -> Tensor:
  """nerf.py:1173-1178."""
  betas = coeffs
  m1t = 1 - t
  for i in range(1, N): betas = betas[:-1] * m1t + betas[1:] * t
  return betas.squeeze(0)


def cubic_bezier(coeffs: Tensor, t: Tensor, N: int) -> Tensor:
  """nerf.py:1201-1206."""
  assert N == 4
  m1t = 1 - t
  m1t_sq, t_sq = m1t * m1t, t * t
  k = torch.stack([m1t_sq * m1t, 3 * m1t_sq * t, 3 * t_sq * m1t, t_sq * t], dim=0)
  return (k * coeffs).sum(dim=0)


def dnerf_spline_forward(params: Params, rays: Tensor, times: Tensor, ts: Tensor, n: int, *, sigmoid: str = "upshifted",
                         bg: str = "black", quant: Optional[torch.dtype] = None) -> Dict[str, Tensor]:
  """DynamicNeRF with n Bezier control points (set_spline_estim nerf.py:1242-1260, spline_interpolate 1267-1278, forward
  1292-1303): delta_estim = SkipConnMLP(in 3, HashEncoder of its own, 5 layers) -> (rigidity[1], points[n,3]);
  dp = Bezier(points, t) (cubic_bezier for n == 4, de_casteljau otherwise); rigid_dp = dp * sigmoid(rigidity / 2)."""
  pts, r_o, r_d = compute_pts(rays, ts)
  t = times[None, :, None, None, None].expand(*pts.shape[:-1], 1)
  p = pts.reshape(-1, 3)
  x0 = torch.cat([p, hash_encode(p, hash_tables(params, "delta_estim.enc"))], dim=-1)
  o = skip_mlp(x0, params, "delta_estim", "leaky_relu", quant=quant).reshape(pts.shape[:-1] + (1 + 3 * n,))
  rigidity, ps = o[..., :1], o[..., 1:]
  rig = (rigidity / 2).sigmoid()
  ps = torch.stack(ps.split([3] * n, dim=-1), dim=0)
  dp = (cubic_bezier if n == 4 else de_casteljau)(ps, t, n)
  rigid_dp = dp * rig
  canon = {k[len("canonical."):]: v for k, v in params.items() if k.startswith("canonical.")}
  res = plain_from_pts(canon, pts + rigid_dp, ts, r_o, r_d, sigmoid=sigmoid, bg=bg, quant=quant)
  res["rigid_dp"] = rigid_dp; res["pts"] = pts
  return res


# ----------------------------------------------------------------------------
# a-7  inverse-CDF resampling, restated from the dead code at nerf.py:1745-1779
# ----------------------------------------------------------------------------
def sample_pdf(bins: Tensor, weights: Tensor, u: Tensor) -> Tensor:
  """``bins[T-1]`` mid-points shared by all rays, ``weights[T-2, R]``, ``u[Nf, R]``
  in [0,1) -> new sample positions ``[Nf, R]``.  Follows lines 1754-1777 with the
  undefined ``bins_g`` read as the gather of ``bins``."""
  w = weights + 1e-5
  pdf = w / w.sum(dim=0, keepdim=True)
  cdf = torch.cumsum(pdf, dim=0)
  cdf = torch.cat([torch.zeros_like(cdf[:1]), cdf], dim=0)   # [T-1, R]
  cdf_t = cdf.transpose(0, 1).contiguous()                   # [R, T-1]
  u_t = u.transpose(0, 1).contiguous()                       # [R, Nf]
  inds = torch.searchsorted(cdf_t, u_t, right=True)
  below = (inds - 1).clamp(min=0)
  above = inds.clamp(max=cdf_t.shape[-1] - 1)
  cdf_b, cdf_a = cdf_t.gather(1, below), cdf_t.gather(1, above)
  bins_b, bins_a = bins[below], bins[above]
  denom = cdf_a - cdf_b
  denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
  t = (u_t - cdf_b) / denom
  return (bins_b + t * (bins_a - bins_b)).transpose(0, 1).contiguous()


def plain_coarse_fine(params: Params, rays: Tensor, ts_coarse: Tensor, u: Tensor, **kw) -> Dict[str, Tensor]:
  """Config 2 (coarse+fine) as SURVEY.md a-7 contracts it: two passes of the Plain pipeline; the fine pass runs on
  the coarse positions merged (sorted) with the inverse-CDF samples.  rays[R,6] flat, u[Nf,R]."""
  coarse = plain_forward(params, rays, ts_coarse, **kw)
  mids = 0.5 * (ts_coarse[:-1] + ts_coarse[1:])
  new = sample_pdf(mids, coarse["weights"][1:-1], u)                       # [Nf,R]
  ts_f = torch.sort(torch.cat([ts_coarse[:, None].expand(-1, rays.shape[0]), new], dim=0), dim=0).values   # [T+Nf,R]
  r_o, r_d = rays.split([3, 3], dim=-1)
  pts = r_o.unsqueeze(0) + ts_f[..., None] * r_d.unsqueeze(0)
  fine = plain_from_pts(params, pts, ts_f, r_o, r_d, per_ray_ts=True, **kw)
  fine["ts"] = ts_f; fine["coarse"] = coarse["out"]
  return fine


# ----------------------------------------------------------------------------
# deterministic synthetic inputs (shared by tests, golden generator, bench)
# ----------------------------------------------------------------------------
def make_plain_params(seed: int = 1337, intermediate: int = 64, sigma_gain: float = 1.0, mip: bool = False,
                      refl_kind: str = "view") -> Params:
  """Parameters with the reference's init *distributions* (first MLP: torch
  default U(+-1/sqrt(fan_in)), neural_blocks.py:258-259; View MLP: siren
  U(+-sqrt(6/fan_in)) with zero bias, 266-271; hash tables N(0,1),
  neural_blocks.py:123-125) drawn from numpy's PCG64 so the values are
  platform independent.  ``sigma_gain`` scales the density row of ``first.out``
  (weight set "S" of SURVEY.md section 8d: sharp density, early termination)."""
  g = np.random.default_rng(seed)
  P: Params = {}
  def uni(shape, a): return torch.from_numpy(g.uniform(-a, a, size=shape).astype(np.float32))
  def default_linear(name, out_f, in_f):
    a = 1.0 / math.sqrt(in_f)
    P[f"{name}.weight"] = uni((out_f, in_f), a); P[f"{name}.bias"] = uni((out_f,), a)
  def siren_linear(name, out_f, in_f):
    P[f"{name}.weight"] = uni((out_f, in_f), math.sqrt(6.0 / in_f)); P[f"{name}.bias"] = torch.zeros(out_f)
  I = intermediate
  ML = 6 * MIP_DEG if mip else 0            # mip latent on both MLP inputs (nerf.py:255, 311-324)
  P["empty_latent"] = torch.zeros(1, 1, 1, 1, 0)
  P["first.enc.primes"] = torch.tensor([1, 2654435761, 805459861, 3674653429, 2097192037, 1434869437, 2165219737])
  for i in range(HASH_LEVELS):
    P[f"first.enc.embs.{i}.weight"] = torch.from_numpy(g.standard_normal((HASH_TABLE, HASH_FEAT)).astype(np.float32))
  default_linear("first.init", 256, 38 + ML)
  default_linear("first.layers.0", 256, 256 + 38 + ML)
  for i in (1, 2, 3): default_linear(f"first.layers.{i}", 256, 256)
  default_linear("first.out", 1 + I, 256)
  if sigma_gain != 1.0:
    P["first.out.weight"][0] *= sigma_gain; P["first.out.bias"][0] *= sigma_gain
  if refl_kind == "pos":
    # refl.Positional (refl.py:230-237): own HashEncoder, 5 layers, torch-default init
    P["refl.mlp.enc.primes"] = P["first.enc.primes"].clone()
    for i in range(HASH_LEVELS):
      P[f"refl.mlp.enc.embs.{i}.weight"] = torch.from_numpy(g.standard_normal((HASH_TABLE, HASH_FEAT)).astype(np.float32))
    w = 38 + ML + I
    default_linear("refl.mlp.init", 256, w)
    for i in range(5): default_linear(f"refl.mlp.layers.{i}", 256, 256 + w if (i % 3 == 0 and i != 4) else 256)
    default_linear("refl.mlp.out", 3, 256)
    return P
  if refl_kind == "pos-linear-view":
    # refl.PosLinearView (refl.py:248-264): pos = hash-encoded MLP (2 layers, hidden 256, torch-default init), view = siren MLP (2 layers, hidden 128)
    P["refl.pos.enc.primes"] = P["first.enc.primes"].clone()
    for i in range(HASH_LEVELS):
      P[f"refl.pos.enc.embs.{i}.weight"] = torch.from_numpy(g.standard_normal((HASH_TABLE, HASH_FEAT)).astype(np.float32))
    w, im = 38 + I, 64
    default_linear("refl.pos.init", 256, w); default_linear("refl.pos.layers.0", 256, 256 + w); default_linear("refl.pos.layers.1", 256, 256)
    default_linear("refl.pos.out", 3 + im, 256)
    wv = 6 + I + im
    siren_linear("refl.view.init", 128, wv); siren_linear("refl.view.layers.0", 128, 128 + wv); siren_linear("refl.view.layers.1", 128, 128)
    siren_linear("refl.view.out", 1, 128)
    return P
  siren_linear("refl.mlp.init", 256, 5 + ML + I)
  siren_linear("refl.mlp.layers.0", 256, 256 + 5 + ML + I)
  for i in (1, 2, 3): siren_linear(f"refl.mlp.layers.{i}", 256, 256)
  siren_linear("refl.mlp.out", 3, 256)
  return P


def make_volsdf_params(seed: int = 7, sdf_kind: str = "siren", intermediate: int = 64, beta: float = 0.1) -> Params:
  """VolSDF + View parameters with the reference's init distributions and state_dict names (SIREN: siren init, MLP:
  xavier-uniform with zero bias and a 16*N(0,1) Fourier basis, src/sdf.py:250-287; View: siren init)."""
  g = np.random.default_rng(seed)
  P: Params = {"empty_latent": torch.zeros(1, 1, 1, 1, 0), "scale": torch.tensor(beta)}
  def uni(shape, a): return torch.from_numpy(g.uniform(-a, a, size=shape).astype(np.float32))
  def siren(name, o, i): P[f"{name}.weight"] = uni((o, i), math.sqrt(6.0 / i)); P[f"{name}.bias"] = torch.zeros(o)
  def xavier(name, o, i): P[f"{name}.weight"] = uni((o, i), math.sqrt(6.0 / (i + o))); P[f"{name}.bias"] = torch.zeros(o)
  I = intermediate
  if sdf_kind == "siren":
    pre, lin, n_layers, d0 = "sdf.underlying.siren", siren, 5, 3
  else:
    pre, lin, n_layers, d0 = "sdf.underlying.mlp", xavier, 6, 3 + 256
    P[f"{pre}.enc.basis"] = torch.from_numpy((16 * g.standard_normal((128, 3))).astype(np.float32)).T.contiguous()
  lin(f"{pre}.init", 256, d0)
  for i in range(n_layers): lin(f"{pre}.layers.{i}", 256, 256 + d0 if (i % 3 == 0 and i != n_layers - 1) else 256)
  lin(f"{pre}.out", 1 + I, 256)
  siren("sdf.refl.mlp.init", 256, 5 + I)
  siren("sdf.refl.mlp.layers.0", 256, 256 + 5 + I)
  for i in (1, 2, 3): siren(f"sdf.refl.mlp.layers.{i}", 256, 256)
  siren("sdf.refl.mlp.out", 3, 256)
  return P


def make_dnerf_params(seed: int = 9, intermediate: int = 64, sigma_gain: float = 20.0, out_scale: float = 0.3) -> Params:
  """DynamicNeRF(direct) over PlainNeRF+View.  delta_estim: xavier-uniform, zero biases (nerf.py:1234-1237); its last
  layer is zero-initialised in the reference (nerf.py:1239) -- here it gets small random values so that the deformation
  is exercised."""
  g = np.random.default_rng(seed)
  P: Params = {"canonical." + k: v for k, v in make_plain_params(seed + 1, intermediate, sigma_gain).items()}
  def xav(name, o, i, scale=1.0):
    a = math.sqrt(6.0 / (i + o)) * scale
    P[f"{name}.weight"] = torch.from_numpy(g.uniform(-a, a, size=(o, i)).astype(np.float32)); P[f"{name}.bias"] = torch.zeros(o)
  xav("delta_estim.init", 256, 4)
  for i in range(5): xav(f"delta_estim.layers.{i}", 256, 260 if (i % 3 == 0 and i != 4) else 256)
  xav("delta_estim.out", 4, 256, out_scale)
  P["delta_estim.out.bias"] = torch.from_numpy(g.uniform(-0.2, 0.2, size=(4,)).astype(np.float32))
  return P


def make_dnerf_spline_params(seed: int = 9, n: int = 5, intermediate: int = 64, sigma_gain: float = 20.0, out_scale: float = 0.3) -> Params:
  """DynamicNeRF(spline=n) over PlainNeRF+View: delta_estim has its own HashEncoder (N(0,1) tables), xavier Linears with
  zero biases (nerf.py:1252-1255); the zero-initialised last layer (nerf.py:1256) gets small random values here so that
  the deformation is exercised."""
  g = np.random.default_rng(seed)
  P: Params = {"canonical." + k: v for k, v in make_plain_params(seed + 1, intermediate, sigma_gain).items()}
  def xav(name, o, i, scale=1.0):
    a = math.sqrt(6.0 / (i + o)) * scale
    P[f"{name}.weight"] = torch.from_numpy(g.uniform(-a, a, size=(o, i)).astype(np.float32)); P[f"{name}.bias"] = torch.zeros(o)
  P["delta_estim.enc.primes"] = torch.tensor([1, 2654435761, 805459861, 3674653429, 2097192037, 1434869437, 2165219737])
  for i in range(HASH_LEVELS):
    P[f"delta_estim.enc.embs.{i}.weight"] = torch.from_numpy(g.standard_normal((HASH_TABLE, HASH_FEAT)).astype(np.float32))
  xav("delta_estim.init", 256, 38)
  for i in range(5): xav(f"delta_estim.layers.{i}", 256, 294 if (i % 3 == 0 and i != 4) else 256)
  xav("delta_estim.out", 1 + 3 * n, 256, out_scale)
  P["delta_estim.out.bias"] = torch.from_numpy(g.uniform(-0.2, 0.2, size=(1 + 3 * n,)).astype(np.float32))
  return P


def make_cameras(n_views: int, size: int = 800, seed: int = 0, radius: float = 4.0):
  """(cam_to_world[B,3,4] fp32, focal) of ``make_rays``: lego-like geometry, focal from camera_angle_x=0.6911112
  (loaders.py:83), cameras on a radius-4 sphere looking at the origin."""
  g = np.random.default_rng(seed)
  focal = 0.5 * size / math.tan(0.5 * 0.6911112)
  c2w = []
  for _ in range(n_views):
    th, ph = g.uniform(0, 2 * math.pi), g.uniform(0.15, 1.2)
    eye = np.array([radius * math.cos(th) * math.cos(ph), radius * math.sin(th) * math.cos(ph), radius * math.sin(ph)])
    fwd = -eye / np.linalg.norm(eye)
    right = np.cross(fwd, np.array([0.0, 0.0, 1.0])); right /= np.linalg.norm(right)
    up = np.cross(right, fwd)
    m = np.eye(4); m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = right, up, -fwd, eye
    c2w.append(m[:3, :4])
  return torch.from_numpy(np.stack(c2w).astype(np.float32)), focal


def dtu_rays(pose: Tensor, intrinsic: Tensor, size: int, top: int, left: int, h: int, w: int) -> Tensor:
  """runner.render's pixel grid (runner.py:495-503) + DTUCamera.sample_positions (reference src/cameras.py:159-174 lift,
  189-223): pose[B,4,4], intrinsic[B,>=3,>=3] -> rays[B,h,w,6] with unit-norm r_d."""
  B = pose.shape[0]
  jj, ii = torch.meshgrid(torch.arange(left, left + w, dtype=torch.float32), torch.arange(top, top + h, dtype=torch.float32), indexing="xy")
  norm = torch.tensor([1600, 1200], dtype=torch.float32) / size
  u = (jj * norm[0]).reshape(1, -1).expand(B, -1); v = (ii * norm[1]).reshape(1, -1).expand(B, -1)
  fx, fy = intrinsic[:, 0, 0, None], intrinsic[:, 1, 1, None]
  cx, cy, sk = intrinsic[:, 0, 2, None], intrinsic[:, 1, 2, None], intrinsic[:, 0, 1, None]
  z = torch.ones_like(u)
  x_lift = (u - cx + cy * sk / fy - sk * v / fy) / fx * z
  y_lift = (v - cy) / fy * z
  points = torch.stack([x_lift, y_lift, z, torch.ones_like(z)], dim=-1)
  world = torch.bmm(pose, points.permute(0, 2, 1)).permute(0, 2, 1)[..., :3]
  r_o = pose[:, None, :3, 3].expand_as(world)
  r_d = F.normalize(world - r_o, dim=-1)
  return torch.cat([r_o, r_d], dim=-1).reshape(B, h, w, 6)


def make_rays(n_views: int, h: int, w: int, size: int = 800, seed: int = 0,
              crop_top: int = 0, crop_left: int = 0, radius: float = 4.0) -> Tensor:
  """``rays[B,H,W,6]`` exactly as runner.render + NeRFCamera.sample_positions build
  them (reference runner.py:495-505, src/cameras.py:45-66; ``with_noise=False``).  r_d is NOT normalised."""
  c2w, focal = make_cameras(n_views, size, seed, radius)
  ii, jj = torch.meshgrid(torch.arange(size, dtype=torch.float), torch.arange(size, dtype=torch.float), indexing="ij")
  positions = torch.stack([ii.transpose(-1, -2), jj.transpose(-1, -2)], dim=-1)
  positions = positions[crop_top:crop_top + h, crop_left:crop_left + w, :]
  u, v = positions.split([1, 1], dim=-1)
  d = torch.stack([(u - size * 0.5) / focal, -(v - size * 0.5) / focal, -torch.ones_like(u)], dim=-1)
  r_d = torch.sum(d[..., None, :] * c2w[..., :3, :3], dim=-1)
  r_d = r_d.permute(2, 0, 1, 3)
  r_o = c2w[..., :3, -1][:, None, None, :].expand_as(r_d)
  return torch.cat([r_o, r_d], dim=-1).contiguous()

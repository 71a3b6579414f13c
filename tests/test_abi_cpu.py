"""CPU-only checks of the boundary: the library loads, exports every symbol the header declares,
and the host-side logic (descriptors, packing order, sharding, module surface) behaves.  No compute
call is made here -- there is no GPU in the build container and no CPU fallback in the product."""
import ctypes as C
import os, re, subprocess, sys
import pytest
import torch
import nerf_atlas_b200 as N
from nerf_atlas_b200 import _lib
from oracle import nerf_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

def header_functions():
  src = open(os.path.join(ROOT, "include", "nerf_b200.h")).read()
  src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
  return sorted(set(re.findall(r"\b(nf_[a-z0-9_]+)\s*\(", src)))

def test_library_exports_every_declared_symbol():
  l = C.CDLL(_lib.LIB_PATH)
  names = header_functions()
  assert len(names) >= 10
  for n in names: assert hasattr(l, n), f"{n} declared in include/nerf_b200.h but not exported"
  assert sorted(_lib.EXPORTS) == names, "ctypes binding table and header disagree"
  assert _lib.lib().nf_version() == _lib.ABI_VERSION == 7
  assert _lib.lib().nf_build_flags() == 0 or os.environ.get('NF_LIB'), 'the default library must be the product build (no experiment hooks)'

def test_descriptor_host_logic():
  l = _lib.lib()
  d = N.describe_plain()
  assert C.sizeof(_lib.ModelDesc) == d.struct_bytes
  assert l.nf_param_count(C.byref(d)) == 2 * 6 + 2 * 6 + 8
  assert l.nf_packed_bytes(C.byref(d)) > 8 * 65536 * 4 * 4          # at least the hash tables
  assert d.hash_res[0] == 16.0 and abs(d.hash_res[7] - 6.2817) < 1e-3
  assert [d.hash_res[i] for i in range(8)] == [float(x) for x in O.hash_resolutions()]
  t = N.describe_tiny()
  assert l.nf_param_count(C.byref(t)) == 2 * 8
  bad = N.describe_plain(); bad.density.hidden = 128
  assert l.nf_param_count(C.byref(bad)) == -2 and b"hidden_size" in l.nf_last_error()
  bad = N.describe_plain(); bad.refl.in_dims = 70
  assert l.nf_packed_bytes(C.byref(bad)) == -1
  bad = N.describe_plain(); bad.struct_bytes = 8
  assert l.nf_param_count(C.byref(bad)) == -1
  # Mip widens both MLP inputs by 96; the spline deformation MLP brings its own hash tables
  m = N.describe_plain(mip="cylinder_ref")
  assert m.density.in_dims == 38 + 96 and m.refl.in_dims == 5 + 96 + 64 and l.nf_param_count(C.byref(m)) == 32
  s5 = N.describe_dyn(spline=5)
  assert s5.deform.out_dims == 16 and l.nf_param_count(C.byref(s5)) == 2 * 6 + 2 * 6 + 2 * 7 + 8 + 8
  bad = N.describe_dyn(spline=5); bad.deform.out_dims = 4
  assert l.nf_param_count(C.byref(bad)) == -1 and b"spline" in l.nf_last_error()
  pos = N.describe_plain(refl_kind="pos")
  assert pos.refl.in_dims == 38 + 64 and pos.refl.n_layers == 5 and l.nf_param_count(C.byref(pos)) == 2 * 6 + 2 * 7 + 8 + 8
  bad = N.describe_plain(mip="cone"); bad.refl.in_dims = 69
  assert l.nf_param_count(C.byref(bad)) == -1

def test_module_state_dict_names_are_the_references():
  m = N.FusedPlainNeRF(steps=16, t_near=2, t_far=6, intermediate_size=64, sigmoid_kind="upshifted")
  ref_names = set(O.make_plain_params(1, 64).keys())
  assert set(m.state_dict().keys()) == ref_names
  m.load_state_dict(O.make_plain_params(1, 64), strict=True)
  assert m.nerf is m and m.intermediate_size == 64 and m.total_latent_size() == 0
  m.set_sigmoid("thin"); assert m.refl.act == "thin"
  with pytest.raises(NotImplementedError): m.set_bg("mlp")
  with pytest.raises(RuntimeError): m(torch.zeros(1, 2, 2, 6))      # CPU rays: loud failure, no fallback
  assert len(m._param_list()) == 32

def test_shard_rays_partitions():
  for R, W, al in ((640000, 8, 800), (35, 2, 1), (7, 8, 1), (0, 4, 1), (1000, 3, 7)):
    prev = 0
    for r in range(W):
      s, e = N.shard_rays(R, r, W, al)
      assert s == prev and e >= s and (s % al == 0 or s == R)
      prev = e
    assert prev == R
  sizes = [N.shard_rays(640000, r, 8, 800)[1] - N.shard_rays(640000, r, 8, 800)[0] for r in range(8)]
  assert max(sizes) - min(sizes) <= 800

def test_product_does_not_import_the_oracle():
  bad = []
  for dp, _, fs in os.walk(os.path.join(ROOT, "nerf_atlas_b200")):
    for f in fs:
      if f.endswith((".py", ".cu", ".cuh", ".h")) and "oracle" in open(os.path.join(dp, f), errors="ignore").read().replace("no PyTorch or CPU fallback", ""):
        txt = open(os.path.join(dp, f), errors="ignore").read()
        if re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M): bad.append(f)
  assert not bad, bad

def test_gloo_world2_sharded_render():
  """N>1 host path: two gloo ranks each render their block with a stand-in render_fn and all-gather."""
  code = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["NF_ROOT"])
from nerf_atlas_b200.shard import ShardedRenderer, shard_rays
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % os.environ["NF_PORT"], rank=int(os.environ["NF_RANK"]), world_size=2)
rays = torch.arange(37 * 6, dtype=torch.float32).reshape(37, 6)
fn = lambda r: r[:, :3] * 2 + r[:, 3:]
full = ShardedRenderer(fn, gather=True, align=4)(rays)
assert torch.equal(full, fn(rays)), "gathered result differs"
loc, (s, e) = ShardedRenderer(fn, gather=False)(rays)
assert (s, e) == shard_rays(37, dist.get_rank(), 2) and torch.equal(loc, fn(rays[s:e]))
dist.barrier(); dist.destroy_process_group(); print("ok")
'''
  import socket
  sk = socket.socket(); sk.bind(("127.0.0.1", 0)); port = sk.getsockname()[1]; sk.close()
  procs = []
  for r in range(2):
    env = dict(os.environ, NF_ROOT=ROOT, NF_PORT=str(port), NF_RANK=str(r))
    procs.append(subprocess.Popen([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
  for p in procs:
    out, _ = p.communicate(timeout=180)
    assert p.returncode == 0 and "ok" in out, out


def test_gloo_world2_gradient_allreduce_equals_full_batch():
  """Training-time collective: two gloo ranks with UNEQUAL ray blocks back-propagate their local sum losses through
  the reference algorithm (the CPU oracle stands in for the fused backward, which is not built), all-reduce the flat gradient
  once, and end up with the gradient of the mean loss over all rays -- equal to the single-process full-batch gradient."""
  code = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["NF_ROOT"])
from nerf_atlas_b200.shard import GradientAllReducer, shard_rays
from oracle import nerf_oracle as O          # test infrastructure: the differentiable stand-in model
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % os.environ["NF_PORT"], rank=int(os.environ["NF_RANK"]), world_size=2)
rank = dist.get_rank()
torch.manual_seed(0)
names = ["first.out.weight", "first.out.bias", "refl.mlp.out.weight", "refl.mlp.layers.1.weight", "first.enc.embs.2.weight"]
def build():
  P = O.make_plain_params(5, 64, 20.0)
  for n in names: P[n] = P[n].clone().requires_grad_(True)
  return P
rays = O.make_rays(1, 3, 7, seed=2, crop_top=398, crop_left=396).reshape(-1, 6)      # 21 rays
target = torch.rand(21, 3)
ts = torch.linspace(2, 6, 12)
# single-process reference: mean loss over all rays
P = build()
loss = ((O.plain_forward(P, rays, ts)["out"] - target) ** 2).sum(-1).mean()
loss.backward()
full = [P[n].grad.clone() for n in names]
# two ranks, unequal blocks: 8 and 13 rays
s, e = (0, 8) if rank == 0 else (8, 21)
Q = build()
local = ((O.plain_forward(Q, rays[s:e], ts)["out"] - target[s:e]) ** 2).sum(-1).sum()
local.backward()
ar = GradientAllReducer([Q[n] for n in names])
ar.begin(e - s)
assert ar.finish() == 21
for n, g in zip(names, full):
  err = float((Q[n].grad - g).abs().max()); scale = float(g.abs().max())
  assert err <= 1e-6 + 1e-5 * scale, (n, err, scale)
dist.barrier(); dist.destroy_process_group(); print("ok")
'''
  import socket
  sk = socket.socket(); sk.bind(("127.0.0.1", 0)); port = sk.getsockname()[1]; sk.close()
  procs = []
  for r in range(2):
    env = dict(os.environ, NF_ROOT=ROOT, NF_PORT=str(port), NF_RANK=str(r))
    procs.append(subprocess.Popen([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
  for p in procs:
    out, _ = p.communicate(timeout=300)
    assert p.returncode == 0 and "ok" in out, out


def test_header_enums_match_the_python_binding():
  """ABI drift guard: every enum value the ctypes layer uses equals the header's."""
  src = open(os.path.join(ROOT, "include", "nerf_b200.h")).read()
  src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
  vals = {}
  for body in re.findall(r"enum\s+\w+\s*\{(.*?)\}", src, flags=re.S):
    for name, v in re.findall(r"(NF_[A-Z0-9_]+)\s*=\s*(-?\d+)", body): vals[name] = int(v)
  expect = {
    "NF_ACT_NONE": _lib.ACT["none"], "NF_ACT_LEAKY": _lib.ACT["leaky_relu"], "NF_ACT_SIN": _lib.ACT["sin"], "NF_ACT_RELU": _lib.ACT["relu"],
    "NF_ENC_NONE": _lib.ENC["none"], "NF_ENC_HASH": _lib.ENC["hash"], "NF_ENC_FOURIER": _lib.ENC["fourier"],
    "NF_DENS_SOFTPLUS_M1": _lib.DENSITY["softplus"], "NF_DENS_RELU": _lib.DENSITY["relu"], "NF_DENS_LAPLACE": _lib.DENSITY["laplace"],
    "NF_BG_BLACK": _lib.BG["black"], "NF_BG_WHITE": _lib.BG["white"],
    "NF_KIND_PLAIN": _lib.KIND["plain"], "NF_KIND_TINY": _lib.KIND["tiny"], "NF_KIND_DYN": _lib.KIND["dyn"],
    "NF_PREC_FP32": _lib.PRECISION["fp32"], "NF_PREC_FP16_TC": _lib.PRECISION["fp16"],
    "NF_MIP_NONE": _lib.MIP[None], "NF_MIP_CYLINDER": _lib.MIP["cylinder"], "NF_MIP_CONE": _lib.MIP["cone"], "NF_MIP_CYLINDER_REF": _lib.MIP["cylinder_ref"],
    "NF_REFL_VIEW": _lib.REFL["view"], "NF_REFL_POSITIONAL": _lib.REFL["pos"],
  }
  for k, v in expect.items(): assert vals.get(k) == v, (k, vals.get(k), v)
  feat = {"NORMAL": "normal", "THIN": "thin", "TANH": "tanh", "CYCLIC": "cyclic", "UPSHIFTED": "upshifted", "FAT": "fat", "LEAKY_RELU": "leaky_relu",
          "RELU": "relu", "SIN": "sin", "UPSHIFTED_SOFTPLUS": "upshifted_softplus", "UPSHIFTED_RELU": "upshifted_relu", "SOFTMAX": "softmax"}
  assert set(_lib.FEAT) == set(__import__("oracle.nerf_oracle", fromlist=["SIGMOIDS"]).SIGMOIDS)   # the whole family of reference src/utils.py:497-511
  for k, v in feat.items(): assert vals.get("NF_FEAT_" + k) == _lib.FEAT[v], k
  m = re.search(r"#define\s+NF_ABI_VERSION\s+(\d+)", src)
  assert m and int(m.group(1)) == _lib.ABI_VERSION


def test_from_reference_adopts_live_reference_models():
  """Drop-in check against the REAL reference modules (only where /root/reference exists, i.e. in the build container):
  `from_reference` shares the reference's own sub-modules (no parameter copies), the packing order has exactly the number of
  tensors the C side expects, and the descriptor validates -- for every model kind on the path."""
  from oracle import ref_shim
  if not ref_shim.available(): pytest.skip("reference tree not present")
  runner, nerf, refl, utils, cameras = ref_shim.load()
  l = _lib.lib()
  def check(fused, n_expected=None):
    d = fused._describe() if hasattr(fused, "_describe") else fused.engine().desc
    n = l.nf_param_count(C.byref(d))
    assert n > 0, l.nf_last_error()
    ps = fused._param_list()
    assert len(ps) == n, (type(fused).__name__, len(ps), n)
    assert l.nf_packed_bytes(C.byref(d)) > 0
    return d
  # PlainNeRF + View, as runner.load_model builds it
  m, a = ref_shim.build_model("plain", 16)
  f = N.FusedPlainNeRF.from_reference(m)
  assert f.first is m.first and f.refl is m.refl and f.steps == 16 and f.refl_kind == "view"
  assert set(f.state_dict().keys()) == set(m.state_dict().keys())
  check(f)
  # PlainNeRF + Positional head (--refl-kind pos)
  m, a = ref_shim.build_model("plain", 16, extra=("--refl-kind", "pos"))
  f = N.FusedPlainNeRF.from_reference(m)
  assert f.refl_kind == "pos" and check(f).refl_kind == _lib.REFL["pos"]
  # PlainNeRF(mip=CylinderGaussian()) built directly (SURVEY a-4 iii)
  m = nerf.PlainNeRF(mip=utils.CylinderGaussian(), steps=16, t_near=2, t_far=6, intermediate_size=64, sigmoid_kind="upshifted", bg="black")
  f = N.FusedPlainNeRF.from_reference(m)
  assert f.mip == "cylinder_ref" and f.total_latent_size() == m.total_latent_size() == 96 and check(f).mip == _lib.MIP["cylinder_ref"]
  # VolSDF, both SDF networks
  for kind in ("siren", "mlp"):
    m, a = ref_shim.build_model("volsdf", 16, extra=("--sdf-kind", kind))
    f = N.FusedVolSDF.from_reference(m)
    assert f.sdf is m.sdf and f.scale is m.scale
    check(f)
  # DynamicNeRF, direct and spline
  for spline in (0, 5):
    canon, a = ref_shim.build_model("plain", 16)
    dyn = nerf.DynamicNeRF(canonical=canon, spline=spline)
    f = N.FusedDynamicNeRF.from_reference(dyn)
    assert f.delta_estim is dyn.delta_estim and f.spline == spline
    d = f.engine().desc
    assert l.nf_param_count(C.byref(d)) == len(f._param_list()) and d.spline_points == spline
  # what is not built says so
  with pytest.raises(NotImplementedError): N.FusedPlainNeRF.from_reference(nerf.PlainNeRF(mip=utils.ConicGaussian(), steps=4))


def test_integration_md_struct_matches_the_binding():
  """The ctypes stub a maintainer would copy from INTEGRATION.md declares nf_model_desc with the same fields, in the same order,
  as the binding the tests run through; and the header declares them in that order too."""
  src = open(os.path.join(ROOT, "INTEGRATION.md")).read()
  m = re.search(r"class ModelDesc\(C.Structure\):.*?_fields_ = \[(.*?)\]\n\n", src, flags=re.S)
  assert m, "ModelDesc stub not found in INTEGRATION.md"
  real = [f[0] for f in _lib.ModelDesc._fields_]
  assert re.findall(r'\("(\w+)"', m.group(1)) == real
  hdr = open(os.path.join(ROOT, "include", "nerf_b200.h")).read()
  body = re.search(r"typedef struct nf_model_desc \{(.*?)\} nf_model_desc;", hdr, flags=re.S).group(1)
  body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
  fields = re.findall(r"(?:int32_t|uint32_t|float|nf_mlp_desc)\s+(\w+)", body)
  assert fields == real, (fields, real)
  # the training / aux structs: INTEGRATION.md stub == binding == header, field for field
  def header_fields(name):
    b = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), hdr, flags=re.S).group(1)
    b = re.sub(r"/\*.*?\*/", "", b, flags=re.S)
    out = []
    for decl in b.split(";"):
      mm = re.match(r"\s*(?:const\s+)?(?:int32_t|int64_t|float|void|nf_train_lin)\s*\*?\s*(.*)", decl.strip(), flags=re.S)
      if mm and mm.group(1): out += [re.sub(r"[\[\]\*\sA-Z_0-9]*$", "", x.strip().lstrip("*")) or x.strip() for x in mm.group(1).split(",")]
    return [re.match(r"\*?\s*(\w+)", x).group(1) for x in out if x]
  for cls, cname in ((_lib.TrainLin, "nf_train_lin"), (_lib.TrainLayout, "nf_train_layout"), (_lib.RenderAux, "nf_render_aux")):
    real = [f[0] for f in cls._fields_]
    assert header_fields(cname) == real, (cname, header_fields(cname), real)
    m = re.search(r"class %s\(C.Structure\):.*?_fields_ = (.*?)\n(?:class|\n)" % cls.__name__, src, flags=re.S)
    assert m, cls.__name__
    assert re.findall(r'"(\w+)"', m.group(1)) == real, cls.__name__


def test_modules_pickle_without_the_library_handle(tmp_path):
  """Checkpoint = torch.save(model) in the reference (runner.py:1221): every mirror pickles whole (no ctypes handle, no packed
  blob) and comes back with the same parameters and settings."""
  mods = [
    N.FusedPlainNeRF(steps=16, t_near=2, t_far=6, intermediate_size=64, sigmoid_kind="upshifted"),
    N.FusedPlainNeRF(steps=16, t_near=2, t_far=6, intermediate_size=64, sigmoid_kind="upshifted", mip="cylinder", refl_kind="pos"),
    N.FusedTinyNeRF(steps=8, t_near=2, t_far=6),
    N.FusedVolSDF(sdf_kind="mlp", steps=8, t_near=0.3, t_far=1.8, intermediate_size=64),
    N.FusedDynamicNeRF(N.FusedPlainNeRF(steps=16, t_near=2, t_far=6, intermediate_size=64), spline=5),
  ]
  for i, m in enumerate(mods):
    m.engine()                                        # creates the ctypes-backed engine that must NOT be pickled
    f = tmp_path / f"m{i}.pt"
    torch.save(m, f)
    r = torch.load(f, weights_only=False)
    assert type(r) is type(m) and r._engine is None
    sa, sb = m.state_dict(), r.state_dict()
    assert sa.keys() == sb.keys() and all(torch.equal(sa[k], sb[k]) for k in sa)
    if hasattr(m, "steps"): assert (r.steps, r.t_near, r.t_far, r.bg) == (m.steps, m.t_near, m.t_far, m.bg)
    r.engine()                                        # and the engine comes back lazily
  assert mods[1].mip == "cylinder" and mods[1].refl_kind == "pos" and mods[1].first.init.weight.shape == (256, 38 + 96)
  assert mods[1].refl.mlp.init.weight.shape == (256, 38 + 96 + 64)


def test_sample_stream_tiling_arithmetic():
  """The tile <-> (ray, t) arithmetic of the staggered pipeline (NfStreamMap in csrc/nf_tc3.cu), restated: for every T a unit is
  `rpu` whole rays = `tpr` whole 128-row tiles, every (ray, t) of a unit appears exactly once, rays are packed across tile
  boundaries only when T % 32 == 0 (warp-aligned, so that a ray's rounding is placement independent), and at most one segment per
  tile continues from the previous tile / stays unfinished."""
  from math import gcd
  ROWS = 128
  def smap(T):
    Tp = T
    if T % 32 != 0 or T // gcd(T, ROWS) > 64: Tp = ROWS // (ROWS // T) if T <= ROWS else (T + ROWS - 1) // ROWS * ROWS
    g = gcd(Tp, ROWS)
    return Tp, Tp // g, ROWS // g
  for T in (1, 7, 16, 32, 48, 64, 96, 100, 128, 160, 192, 224, 256, 300, 384, 1000, 2048):
    Tp, tpr, rpu = smap(T)
    assert rpu * Tp == tpr * ROWS and Tp >= T
    seen = set()
    for sub in range(tpr):
      cont = unfinished = 0
      segs = set()
      for r in range(ROWS):
        q = sub * ROWS + r; rl, t = divmod(q, Tp)
        if t < T: assert (rl, t) not in seen; seen.add((rl, t))
        segs.add(rl)
      for rl in segs:
        cont += rl * Tp < sub * ROWS
        unfinished += (rl + 1) * Tp > (sub + 1) * ROWS
      assert cont <= 1 and unfinished <= 1
    assert seen == {(rl, t) for rl in range(rpu) for t in range(T)}
    if Tp == T and T > 32: assert T % 32 == 0
  assert smap(192) == (192, 3, 2) and smap(160) == (160, 5, 4) and smap(100) == (128, 1, 1) and smap(64) == (64, 1, 2) and smap(256) == (256, 2, 1)


def test_which_models_the_tensor_pipeline_takes():
  """nf_tensor_pipeline_support (host-only): every model kind of the path runs on tcgen05; what does not says why."""
  l = _lib.lib()
  ok = lambda d: l.nf_tensor_pipeline_support(C.byref(d))
  for d in (N.describe_plain(), N.describe_tiny(), N.describe_volsdf("siren"), N.describe_volsdf("mlp"), N.describe_dyn(), N.describe_dyn(spline=5),
            N.describe_dyn(spline=8), N.describe_plain(mip="cylinder"), N.describe_plain(mip="cone"), N.describe_plain(mip="cylinder_ref"),
            N.describe_plain(refl_kind="pos"), N.describe_plain(mip="cylinder", refl_kind="pos"), N.describe_plain(32)):
    assert ok(d) is None, ok(d)
  assert b"wider than 80" in ok(N.describe_plain(128))       # intermediate_size 128: View x0 is 133 wide (fp32 pipeline only for now)
  bad = N.describe_plain(); bad.intermediate = 40; bad.density.out_dims = 41; bad.refl.in_dims = 45
  assert b"multiple of 16" in ok(bad)
  bad = N.describe_plain(hash_levels=7)
  assert b"odd number of hash levels" in ok(bad)
  dm = N.describe_dyn(); dm.mip = _lib.MIP["cylinder"]; dm.density.in_dims += 96; dm.refl.in_dims += 96
  assert b"PlainNeRF only" in ok(dm)
  bad = N.describe_plain(); bad.density.hidden = 128
  assert b"hidden_size" in ok(bad)            # invalid descriptor: the plan builder's message


def test_train_layout_host_logic():
  """nf_train_layout_of is host-only: the stash regions tile the workspace without overlap, tile counts follow the sample-stream
  tiling, and models the backward does not take are refused with a reason (no silent fallback)."""
  import ctypes as C
  import nerf_atlas_b200 as N
  lib = _lib.lib()
  d = N.describe_plain(64, "upshifted", "black")
  for R, T in ((4096, 128), (12, 16), (50, 192), (33, 100), (1, 1)):
    lay = _lib.TrainLayout()
    assert lib.nf_train_layout_of(C.byref(d), R, T, C.byref(lay)) == 0
    assert lay.n_lin == 12 and lay.n_rays == R and lay.T == T
    units = (R + lay.rpu - 1) // lay.rpu
    assert lay.n_tiles == units * lay.tpr and lay.n_tiles * 128 >= R * T
    regions = [(lay.scale_off, 16), (lay.sigma_off, R * T * 4), (lay.rgbraw_off, R * T * 12), (lay.dsigma_off, R * T * 4),
               (lay.drgbraw_off, R * T * 12), (lay.dx0_off, lay.n_tiles * 128 * 32 * 4)]
    for i in range(lay.n_lin):
      L = lay.lin[i]
      assert L.a_tile == (L.k0_pad + L.k_hidden) * 256 and L.g_tile == L.n_pad * 256
      regions += [(L.a_off, lay.n_tiles * L.a_tile), (L.g_off, lay.n_tiles * L.g_tile), (L.dw_off, L.n_pad * (L.k0_pad + L.k_hidden) * 4), (L.db_off, L.n_pad * 4)]
      assert (L.c_off >= 0) == (L.act == _lib.ACT["sin"] and L.k_hidden > 0)
      if L.c_off >= 0: regions.append((L.c_off, lay.n_tiles * 65536))
      assert lay.dw_begin <= L.dw_off and L.db_off + L.n_pad * 4 <= lay.dw_end
    regions.sort()
    for (o0, n0), (o1, _) in zip(regions, regions[1:]): assert o0 % 1024 == 0 and o0 + n0 <= o1, (R, T, o0, n0, o1)
    assert regions[-1][0] + regions[-1][1] <= lay.total_bytes
  assert [(lay.lin[i].m, lay.lin[i].j) for i in range(12)] == [(0, j) for j in range(6)] + [(1, j) for j in range(6)]
  lay = _lib.TrainLayout()                                                      # VolSDF, SIREN SDF + View: 7 + 6 Linears, cos stash for both MLPs
  assert lib.nf_train_layout_of(C.byref(N.describe_volsdf("siren")), 16, 16, C.byref(lay)) == 0 and lay.n_lin == 13
  assert all((lay.lin[i].c_off >= 0) == (lay.lin[i].k_hidden > 0) for i in range(13))
  pos = N.describe_plain(64, "upshifted", "black", refl_kind="pos")              # Positional head: 6 + 7 Linears, warp-aligned rays only
  assert lib.nf_train_layout_of(C.byref(pos), 16, 32, C.byref(lay)) == 0 and lay.n_lin == 13 and lay.lin[6].k0_pad == 112
  assert lib.nf_train_layout_of(C.byref(N.describe_tiny()), 16, 16, C.byref(lay)) == 0 and lay.n_lin == 8      # TinyNeRF: one MLP
  assert lib.nf_train_layout_of(C.byref(N.describe_plain(64, "upshifted", "random")), 16, 16, C.byref(lay)) == 0 and lay.bgrand_off > 0      # draws kept for the backward
  assert lib.nf_train_layout_of(C.byref(N.describe_plain(64, "upshifted", "black")), 16, 16, C.byref(lay)) == 0 and lay.bgrand_off == -1
  for bad in (N.describe_volsdf("mlp"), N.describe_dyn(), N.describe_plain(64, "upshifted", "black", mip="cylinder"),
              N.describe_plain(64, "upshifted", "black", refl_kind="pos")):
    lay = _lib.TrainLayout()
    assert lib.nf_train_layout_of(C.byref(bad), 16, 16, C.byref(lay)) == -2        # NF_E_UNSUPPORTED
    assert b"training" in lib.nf_last_error()

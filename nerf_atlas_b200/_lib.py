"""ctypes binding of the C ABI in include/nerf_b200.h (libnerf_b200.so, built in-tree by
nerf_atlas_b200/build.py).  There is NO fallback: if the library is missing or a call fails, the
caller gets an exception -- the product path never routes through PyTorch ops or the CPU oracle."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, os.environ.get("NF_LIB", "libnerf_b200.so"))

ABI_VERSION = 7
# enums of include/nerf_b200.h
ACT = {"none": 0, "leaky_relu": 1, "sin": 2, "relu": 3}
ENC = {"none": 0, "hash": 1, "fourier": 2}
DENSITY = {"softplus": 0, "relu": 1, "laplace": 2}
FEAT = {"normal": 0, "thin": 1, "tanh": 2, "cyclic": 3, "upshifted": 4, "fat": 5, "leaky_relu": 6, "relu": 7,
        "sin": 8, "upshifted_softplus": 9, "upshifted_relu": 10, "softmax": 11}
BG = {"black": 0, "white": 1, "random": 2}
KIND = {"plain": 0, "tiny": 1, "dyn": 2}
PRECISION = {"fp32": 0, "fp16": 1}
REFL = {"view": 0, "pos": 1, "pos-linear-view": 2}
MIP = {None: 0, "none": 0, "cylinder": 1, "cone": 2, "cylinder_ref": 3}

class MlpDesc(C.Structure):
  _fields_ = [("in_dims", C.c_int32), ("hidden", C.c_int32), ("n_layers", C.c_int32),
              ("out_dims", C.c_int32), ("skip", C.c_int32), ("act", C.c_int32)]

class ModelDesc(C.Structure):
  _fields_ = [("struct_bytes", C.c_int32), ("kind", C.c_int32), ("density", MlpDesc), ("refl", MlpDesc),
              ("intermediate", C.c_int32), ("enc", C.c_int32), ("hash_levels", C.c_int32),
              ("hash_table_size", C.c_int32), ("hash_feat", C.c_int32), ("hash_primes", C.c_uint32 * 3),
              ("hash_res", C.c_float * 16), ("density_act", C.c_int32), ("feat_act", C.c_int32), ("bg", C.c_int32),
              ("fourier_freqs", C.c_int32), ("deform", MlpDesc), ("mip", C.c_int32), ("deform_enc", C.c_int32),
              ("refl_kind", C.c_int32), ("spline_points", C.c_int32), ("refl_view", MlpDesc)]

class MipArgs(C.Structure):
  _fields_ = [("radius", C.c_void_p), ("rays_all", C.c_void_p), ("radius_all", C.c_void_p), ("n_rays_all", C.c_int64),
              ("ray_base", C.c_int64)]

TRAIN_LIN_MAX = 24

class TrainLin(C.Structure):
  _fields_ = [(n, C.c_int32) for n in ("m", "j", "n", "n_pad", "k0_pad", "k_hidden", "act", "x0_raw")] + \
             [(n, C.c_int64) for n in ("a_off", "a_tile", "c_off", "g_off", "g_tile", "dw_off", "db_off")]

class TrainLayout(C.Structure):
  _fields_ = [(n, C.c_int32) for n in ("n_lin", "T", "rpu", "tpr")] + \
             [(n, C.c_int64) for n in ("n_rays", "n_tiles", "sigma_off", "rgbraw_off", "dsigma_off", "drgbraw_off", "dx0_off",
                                       "scale_off", "dw_begin", "dw_end", "total_bytes", "bgrand_off")] + [("lin", TrainLin * TRAIN_LIN_MAX)]

class RenderAux(C.Structure):
  _fields_ = [("struct_bytes", C.c_int32), ("reserved", C.c_int32), ("train_ws", C.c_void_p), ("train_ws_bytes", C.c_int64),
              ("pts", C.c_void_p), ("bg_rand", C.c_void_p), ("pts_out", C.c_void_p), ("dp_out", C.c_void_p),
              ("rigid_dp_out", C.c_void_p), ("rigidity_out", C.c_void_p)]

EXPORTS = {
  "nf_version": (C.c_int, []),
  "nf_last_error": (C.c_char_p, []),
  "nf_build_flags": (C.c_int, []),
  "nf_tensor_pipeline_support": (C.c_char_p, [C.POINTER(ModelDesc)]),
  "nf_param_count": (C.c_int, [C.POINTER(ModelDesc)]),
  "nf_packed_bytes": (C.c_int64, [C.POINTER(ModelDesc)]),
  "nf_pack_weights": (C.c_int, [C.POINTER(ModelDesc), C.POINTER(C.c_void_p), C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]),
  "nf_render_forward": (C.c_int, [C.POINTER(ModelDesc), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int64,
                                  C.c_void_p, C.c_void_p, C.POINTER(MipArgs), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
  "nf_render_forward_aux": (C.c_int, [C.POINTER(ModelDesc), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int64,
                                      C.c_void_p, C.c_void_p, C.POINTER(MipArgs), C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(RenderAux),
                                      C.c_int32, C.c_void_p]),
  "nf_train_layout_of": (C.c_int, [C.POINTER(ModelDesc), C.c_int64, C.c_int32, C.POINTER(TrainLayout)]),
  "nf_render_backward": (C.c_int, [C.POINTER(ModelDesc), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32,
                                   C.c_int64, C.c_void_p, C.POINTER(C.c_void_p), C.c_int32, C.c_void_p]),
  "nf_sdf_workspace_bytes": (C.c_int64, [C.POINTER(ModelDesc), C.c_int64]),
  "nf_sphere_march": (C.c_int, [C.POINTER(ModelDesc), C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_float, C.c_int32, C.c_float, C.c_float,
                                C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
  "nf_sdf_render": (C.c_int, [C.POINTER(ModelDesc), C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_float, C.c_int32, C.c_float, C.c_float,
                              C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
  "nf_generate_rays": (C.c_int, [C.c_void_p, C.c_int64, C.c_float, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                 C.c_void_p, C.c_void_p]),
  "nf_generate_rays_dtu": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                     C.c_void_p, C.c_void_p]),
  "nf_ray_radii": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
  "nf_sample_points": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]),
  "nf_hash_encode": (C.c_int, [C.POINTER(ModelDesc), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
  "nf_composite": (C.c_int, [C.POINTER(ModelDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int64,
                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
  "nf_sample_pdf": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
  "nf_composite_backward": (C.c_int, [C.POINTER(ModelDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32,
                                      C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
  "nf_hash_encode_backward": (C.c_int, [C.POINTER(ModelDesc), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
  "nf_adam_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                             C.c_int32, C.c_void_p]),
  "nf_sdf_bisect": (C.c_int, [C.POINTER(ModelDesc), C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_float, C.c_int32, C.c_float, C.c_float,
                              C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
  "nf_integrate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]),
  "nf_sdf_normals": (C.c_int, [C.POINTER(ModelDesc), C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
  "nf_adam_step_multi": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float,
                                   C.c_float, C.c_int32, C.c_void_p]),
  "nf_mlp_forward": (C.c_int, [C.POINTER(ModelDesc), C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p]),
}

_lib = None

def lib():
  """The loaded library; raises if it has not been built."""
  global _lib
  if _lib is None:
    if not os.path.exists(LIB_PATH):
      raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m nerf_atlas_b200.build` "
                         "(there is no PyTorch/CPU fallback for the render path)")
    l = C.CDLL(LIB_PATH)
    for name, (res, args) in EXPORTS.items():
      fn = getattr(l, name)          # AttributeError if the export is missing
      fn.restype, fn.argtypes = res, args
    if l.nf_version() != ABI_VERSION: raise RuntimeError("libnerf_b200.so: ABI version mismatch")
    _lib = l
  return _lib

def has_experiments() -> bool:
  """True for an NF_EXPERIMENTS build (the NF_TC_* environment switches and the superseded pipelines exist)."""
  return bool(lib().nf_build_flags() & 1)

def check(rc: int, what: str):
  if rc != 0:
    msg = lib().nf_last_error().decode("utf-8", "replace")
    raise RuntimeError(f"{what} failed (code {rc}): {msg}")

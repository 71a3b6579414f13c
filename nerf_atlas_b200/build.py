"""Build the C-ABI library (include/nerf_b200.h) in-tree for sm_100a with nvcc.

    python -m nerf_atlas_b200.build [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os, subprocess, sys, hashlib

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
EXPERIMENTS = bool(os.environ.get("NF_EXPERIMENTS"))   # A/B-timing build: the NF_TC_* environment switches, incl. the single-CTA pipeline as the render path (libnerf_b200_exp.so)
SOURCES = ["nf_api.cu", "nf_fp32.cu", "nf_tc.cu", "nf_tc3.cu", "nf_bwd.cu", "nf_train.cu", "nf_march.cu"]
HEADERS = ["nf_common.cuh", "nf_kernels.h", "nf_tc_ptx.cuh", os.path.join("..", "..", "include", "nerf_b200.h")]
TRACE = bool(os.environ.get("NF_TC_TRACE"))     # debug build: clock64 timeline of one tile (profiles/)
STATS = bool(os.environ.get("NF_TC_STATS"))     # debug build: time-in-state counters of the staggered pipeline (nf_tc3.cu)
DEFS = os.environ.get("NF_BUILD_DEFS", "")      # experiments: extra -D flags, e.g. NF_BUILD_DEFS="NF_SIN_POLY_PAIRS=2" NF_BUILD_TAG=poly2
TAG = os.environ.get("NF_BUILD_TAG", "") or ("exp" if EXPERIMENTS else "")
LIB = os.path.join(HERE, "libnerf_b200_trace.so" if TRACE else "libnerf_b200_stats.so" if STATS else f"libnerf_b200_{TAG}.so" if TAG else "libnerf_b200.so")
STAMP = LIB + ".stamp"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "--shared", "-Xptxas", "-v"] + ([f"-DNF_TC_TRACE={os.environ.get('NF_TC_TRACE')}"] if TRACE else []) + (["-DNF_TC_STATS=1"] if STATS else []) + (["-DNF_EXPERIMENTS=1"] if EXPERIMENTS else []) + [f"-D{d}" for d in DEFS.split()]

def _digest() -> str:
  h = hashlib.sha256()
  for f in sorted(SOURCES) + HEADERS:
    with open(os.path.join(CSRC, f), "rb") as fh: h.update(fh.read())
  h.update(" ".join(FLAGS).encode())
  return h.hexdigest()

def build(force: bool = False, verbose: bool = False) -> str:
  d = _digest()
  if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == d:
    return LIB
  cmd = [NVCC, *FLAGS, "-o", LIB, *[os.path.join(CSRC, s) for s in SOURCES]]
  r = subprocess.run(cmd, capture_output=True, text=True)
  if verbose or r.returncode != 0: sys.stderr.write(r.stdout + r.stderr)
  if r.returncode != 0: raise RuntimeError("nvcc failed building libnerf_b200.so")
  with open(os.path.join(HERE, "ptxas_info.txt"), "w") as f: f.write(r.stderr)
  with open(STAMP, "w") as f: f.write(d)
  return LIB

if __name__ == "__main__":
  print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))

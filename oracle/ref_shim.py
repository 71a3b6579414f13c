"""Import the read-only reference tree (/root/reference) with the shim list of
SURVEY.md section 8c.  TEST INFRASTRUCTURE: used only by tests/golden/make_golden.py
(in the build container) and by tests that are skipped when the tree is absent.
The reference tree does not exist on the GPU box; nothing on a gpu-marked test,
smoke() or bench.py path may call this.
"""
import os, sys, types, tempfile
import torch

REF = os.environ.get("NERF_ATLAS_REF", "/root/reference")

def available() -> bool: return os.path.isdir(os.path.join(REF, "src"))

_mods = None
def load():
  """Returns (runner, nerf, refl, utils, cameras) reference modules."""
  global _mods
  if _mods is not None: return _mods
  if not available(): raise RuntimeError(f"reference tree not found at {REF}")
  # (1) absent third-party deps (utils.py:8, loaders.py:14, runner.py:363,434)
  mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
  mpl.pyplot = plt; plt.colormaps = lambda: ["magma"]; plt.set_cmap = lambda *_: None
  for name, mod in (("matplotlib", mpl), ("matplotlib.pyplot", plt), ("imageio", types.ModuleType("imageio"))):
    sys.modules.setdefault(name, mod)
  sys.path.insert(0, REF)
  import src.nerf as nerf, src.refl as refl, src.utils as utils, src.cameras as cameras, runner
  # (2) CPU only: neural_blocks.py:144 hard-codes emb.cuda()
  if not torch.cuda.is_available(): torch.nn.Embedding.cuda = lambda self, *a, **k: self
  # (3) nerf.py:895 reads an undefined global
  nerf.with_transmission = False
  runner.device = "cpu"
  _mods = (runner, nerf, refl, utils, cameras)
  return _mods

def build_model(model: str = "plain", steps: int = 16, extra=()):
  """runner.arguments() + runner.load_model() (runner.py:37-438,1174-1213) on CPU."""
  runner, nerf, refl, utils, cameras = load()
  tmp = tempfile.mkdtemp()
  argv = sys.argv
  sys.argv = ["runner.py", "-d", "x/", "--outdir", tmp, "--size", "16", "--crop-size", "0",
              "--model", model, "--steps", str(steps), *extra]
  try:
    a = runner.arguments()
  finally:
    sys.argv = argv
  a.num_labels = 1
  runner.seed(a.seed)
  m = runner.load_model(a, None, False)
  return m, a

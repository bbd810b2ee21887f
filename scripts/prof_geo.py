"""Small geometry-only driver for ncu captures: decodes N synthetic 50k-vertex frames once."""
import importlib, os, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, root)
uv = importlib.import_module("universal-volumetric_b200")
from tools.synth import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
drc, _, info = synth.make_sequence(n, 50000, 32, want_textures=False, seed=20260002, distinct_geometry=min(n, 4))
ctx = uv.Context(0); dl = uv.DRACOLoader(ctx)
for _ in range(reps):
    out = dl.decode_batch_raw(drc, uv.MEM_DEVICE)
print("ok", sum(o.status == 0 for o in out), "of", n)

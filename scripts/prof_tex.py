"""Small texture-only driver for ncu captures: transcodes N synthetic 1024^2 x 7 ETC1S segments once."""
import importlib, os, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, root)
uv = importlib.import_module("universal-volumetric_b200")
from tools.synth import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
_, ktx, info = synth.make_sequence(7 * n, 2000, 1024, sequence_size=7, seed=20260002, distinct_geometry=1)
ctx = uv.Context(0); kl = uv.KTX2Loader(ctx)
out = kl.transcode_batch_raw(ktx, uv.MEM_DEVICE)
print("ok", sum(o.status == 0 for o in out), "of", len(ktx))

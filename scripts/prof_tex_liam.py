"""Texture-only driver for ncu captures on REAL data: the committed liam ETC1S fixtures (tests/golden/liam/*.ktx2, 5 x 1024^2 each),
repeated N times in one batch, transcoded twice (first pass = warm-up)."""
import glob, importlib, os, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, root)
uv = importlib.import_module("universal-volumetric_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
files = [open(p, "rb").read() for p in sorted(glob.glob(os.path.join(root, "tests", "golden", "liam", "*.ktx2")))]
ktx = [files[i % len(files)] for i in range(n)]
ctx = uv.Context(0, profiling=True); kl = uv.KTX2Loader(ctx)
for _ in range(2):
    out = kl.transcode_batch_raw(ktx, uv.MEM_DEVICE)
print("ok", sum(o.status == 0 for o in out), "of", len(ktx), ctx.stats(1)["stages"])

"""First GPU shake-out: decode fixture frames through the C ABI and compare with the oracle."""
import ctypes, glob, importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
uv = importlib.import_module("universal-volumetric_b200")
from tests.oracle_bind import oracle_draco

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
files = sorted(glob.glob(os.path.join(root, "oracle/_ref/fixtures/geometry_draco/*.drc"))) or sorted(glob.glob(os.path.join(root, "tests/golden/liam/*.drc")))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
blobs = [open(f, "rb").read() for f in files[:n]]
ctx = uv.Context(0, profiling=True)
dl = uv.DRACOLoader(ctx)
t0 = time.time(); res = dl.decode_batch(blobs); t1 = time.time()
print("first call %.1f ms" % ((t1 - t0) * 1e3), ctx.stats())
bad = 0
for i, (b, r) in enumerate(zip(blobs, res)):
    o = oracle_draco(b)
    if r["status"] != 0:
        print(i, "status", r["status"]); bad += 1; continue
    ok = r["num_points"] == o["num_points"] and np.array_equal(r["index"], o["index"])
    for k in ("position", "normal", "uv"):
        ok = ok and np.array_equal(r["attributes"][k].view(np.uint32), o[k].view(np.uint32))
    if not ok:
        bad += 1
        print(i, "MISMATCH", r["num_points"], o["num_points"], [int((r["attributes"][k] != o[k]).sum()) for k in ("position", "normal", "uv") if r["attributes"][k].shape == o[k].shape], int((r["index"] != o["index"]).sum()) if r["index"].shape == o["index"].shape else -1)
print("frames", len(blobs), "bad", bad)
for rep in range(3):
    t0 = time.time(); dl.decode_batch_raw(blobs); t1 = time.time()
    s = ctx.stats()
    print("rep %d: %.2f ms total, %.1f frames/s, stages %s" % (rep, (t1 - t0) * 1e3, len(blobs) / (t1 - t0), {k: round(v, 3) for k, v in s["stages"].items()}), "parse %.2f" % s["host_parse_ms"])

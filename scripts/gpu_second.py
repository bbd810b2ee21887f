"""GPU shake-out for geometry + texture on fixtures: parity vs oracle and stage timings."""
import glob, importlib, os, sys, time
import numpy as np
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
uv = importlib.import_module("universal-volumetric_b200")
from tests.oracle_bind import oracle_draco, oracle_ktx2
ng = int(sys.argv[1]) if len(sys.argv) > 1 else 8
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dfiles = sorted(glob.glob(os.path.join(root, "oracle/_ref/fixtures/geometry_draco/*.drc")))[:ng]
kfiles = sorted(glob.glob(os.path.join(root, "oracle/_ref/fixtures/texture_ktx2/*.ktx2")))[:nt]
ctx = uv.Context(0, profiling=True)
kl = uv.KTX2Loader(ctx); dl = uv.DRACOLoader(ctx)
kb = [open(f, "rb").read() for f in kfiles]
res = kl.transcode_batch(kb)
bad = 0
for i, (b, r) in enumerate(zip(kb, res)):
    o = oracle_ktx2(b)
    if r["status"] != 0 or not np.array_equal(r["data"], o["rgba"]):
        bad += 1; print("tex", i, "status", r["status"], "mismatch", None if r["data"] is None else int((r["data"] != o["rgba"]).sum()))
print("textures", len(kb), "bad", bad)
for rep in range(3):
    t0 = time.time(); kl.transcode_batch_raw(kb); t1 = time.time(); s = ctx.stats(1)
    print("tex rep %d: %.2f ms, %.1f frames/s" % (rep, (t1 - t0) * 1e3, sum(r["layers"] for r in res) / (t1 - t0)), {k: round(v, 3) for k, v in s["stages"].items()}, "parse %.2f" % s["host_parse_ms"])
t0 = time.time(); kl.transcode_batch_raw(kb, uv.MEM_DEVICE); t1 = time.time(); print("tex device-out: %.2f ms" % ((t1 - t0) * 1e3))
db = [open(f, "rb").read() for f in dfiles]
res = dl.decode_batch(db); bad = 0
for i, (b, r) in enumerate(zip(db, res)):
    o = oracle_draco(b)
    ok = r["status"] == 0 and r["num_points"] == o["num_points"] and np.array_equal(r["index"], o["index"]) and all(np.array_equal(r["attributes"][k].view(np.uint32), o[k].view(np.uint32)) for k in ("position", "normal", "uv"))
    bad += (not ok)
print("geometry", len(db), "bad", bad)
for rep in range(2):
    t0 = time.time(); dl.decode_batch_raw(db); t1 = time.time(); s = ctx.stats(0)
    print("geo rep %d: %.2f ms, %.1f frames/s" % (rep, (t1 - t0) * 1e3, len(db) / (t1 - t0)), {k: round(v, 3) for k, v in s["stages"].items()}, "parse %.2f" % s["host_parse_ms"])

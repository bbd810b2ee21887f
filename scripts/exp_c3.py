"""C3 scaling experiment: geometry-only and texture-only device times at 200k verts / 2048^2 UASTC for growing batches."""
import importlib, os, sys, time
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, root)
uv = importlib.import_module("universal-volumetric_b200")
from tools.synth import synth
verts = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
counts = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [100, 300, 1000]
t0 = time.time()
drc, _, info = synth.make_sequence(max(counts), verts, 32, want_textures=False, seed=20260003, distinct_geometry=16)
print("gen geometry %.1fs" % (time.time() - t0), info["verts"], info["faces"], len(drc[0]), flush=True)
ctx = uv.Context(0, profiling=True); dl = uv.DRACOLoader(ctx)
for n in counts:
    for _ in range(2):
        out = dl.decode_batch_raw(drc[:n], uv.MEM_DEVICE)
    st = ctx.stats(0)
    assert all(o.status == 0 for o in out)
    print("geo n=%d device %.1f ms  %.0f frames/s  scratch %.1f GB" % (n, st["device_ms"], n / st["device_ms"] * 1e3, st["scratch_bytes"] / 1e9),
          {k: round(v, 1) for k, v in st["stages"].items() if v > 2}, flush=True)
if len(sys.argv) > 3:
    t0 = time.time()
    segs = [synth.encode_uastc(synth.texture_layers(2048, 7 * s, 7, 5), seed=s) for s in range(2)]
    print("gen tex %.1fs" % (time.time() - t0), len(segs[0]), flush=True)
    kl = uv.KTX2Loader(ctx)
    for n in (16, 143):
        ktx = [segs[i % 2] for i in range(n)]
        for _ in range(2):
            out = kl.transcode_batch_raw(ktx, uv.MEM_DEVICE)
        st = ctx.stats(0)
        assert all(o.status == 0 for o in out)
        print("tex n=%d device %.2f ms" % (n, st["device_ms"]), {k: round(v, 2) for k, v in st["stages"].items()}, flush=True)

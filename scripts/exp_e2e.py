"""Where does the e2e path lose time against the resident replay?  One C3 window (nseg segments), per-stage CUDA-event times for:
(a) replay, device outputs  (b) fresh decode, device outputs  (c) fresh decode, host outputs  (d) as (c) without textures."""
import importlib, os, sys, time
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, root)
uv = importlib.import_module("universal-volumetric_b200")
from tools.synth import synth
nseg = int(sys.argv[1]) if len(sys.argv) > 1 else 36
drc, ktx, info = synth.make_sequence(7 * nseg, 200000, 2048, sequence_size=7, seed=20260003, distinct_geometry=16, distinct_textures=4, texture_format="uastc")
ctx = uv.Context(0, profiling=True); pl = uv.V2Player(ctx)
def show(tag, wall):
    a, b = ctx.stats(0, combined=True), ctx.stats(1, combined=True)
    print("%-34s wall %7.1f | geo: parse %5.1f h2d %5.1f kernels %6.1f d2h %6.1f | tex: parse %5.1f h2d %5.1f kernels %5.1f d2h %6.1f" % (
        tag, wall, a["host_parse_ms"], a["h2d_ms"], a["device_ms"], a["d2h_ms"], b["host_parse_ms"], b["h2d_ms"], b["device_ms"], b["d2h_ms"]))
    print("      ", {k: round(v, 1) for k, v in a["stages"].items() if v > 4}, flush=True)
for _ in range(2):
    pl.decode_step_raw(drc, ktx, uv.MEM_HOST)
for rep in range(2):
    t = time.perf_counter(); pl.replay_step_raw(len(drc), len(ktx), uv.MEM_DEVICE); show("(a) replay, device out", (time.perf_counter() - t) * 1e3)
    t = time.perf_counter(); pl.decode_step_raw(drc, ktx, uv.MEM_DEVICE); show("(b) decode, device out", (time.perf_counter() - t) * 1e3)
    t = time.perf_counter(); pl.decode_step_raw(drc, ktx, uv.MEM_HOST); show("(c) decode, host out", (time.perf_counter() - t) * 1e3)
    t = time.perf_counter(); pl.decode_step_raw(drc, [], uv.MEM_HOST); show("(d) decode, host out, no textures", (time.perf_counter() - t) * 1e3)
    pl.decode_step_raw(drc, ktx, uv.MEM_HOST)

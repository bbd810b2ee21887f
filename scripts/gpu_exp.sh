#!/bin/bash
# experiment: stage times with/without side-stream overlap, geometry only
for mode in overlap nooverlap; do
  if [ $mode = nooverlap ]; then export UVOL_NO_OVERLAP=1; else unset UVOL_NO_OVERLAP; fi
  python - <<'PY'
import importlib, os, sys, time
sys.path.insert(0, os.getcwd())
uv = importlib.import_module("universal-volumetric_b200")
from tools.synth import synth
for n in (16, 300):
    drc, _, info = synth.make_sequence(n, 50000, 32, want_textures=False, seed=20260002, distinct_geometry=min(n, 8))
    ctx = uv.Context(0, profiling=True); dl = uv.DRACOLoader(ctx)
    for _ in range(2): dl.decode_batch_raw(drc, uv.MEM_DEVICE)
    st = ctx.stats(0)
    print(os.environ.get("UVOL_NO_OVERLAP", "overlap"), n, "frames: device %.1f ms" % st["device_ms"], {k: round(v, 1) for k, v in st["stages"].items() if v > 1})
    ctx.close()
PY
done

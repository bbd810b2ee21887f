"""Small C3-shaped driver for `ncu --set full` captures: one window of 8 segments (56 frames of 200k verts + 8 UASTC 2048^2 x 7
segments) decoded twice through uvol_decode_v2_batch (first pass = warm-up, skip it with `-s`)."""
import importlib, os, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, root)
uv = importlib.import_module("universal-volumetric_b200")
from tools.synth import synth
nseg = int(sys.argv[1]) if len(sys.argv) > 1 else 8
drc, ktx, info = synth.make_sequence(7 * nseg, 200000, 2048, sequence_size=7, seed=20260003, distinct_geometry=8, distinct_textures=2, texture_format="uastc")
ctx = uv.Context(0); pl = uv.V2Player(ctx)
for _ in range(2):
    g, t = pl.decode_step_raw(drc, ktx, uv.MEM_DEVICE)
print("ok", sum(x.status == 0 for x in g), sum(x.status == 0 for x in t))

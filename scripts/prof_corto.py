"""Small V1-shaped driver for `ncu --set full` captures: 64 Corto frames of 50k verts (position + uv) decoded twice through
uvol_decode_corto_batch (first pass = warm-up, skip it with `-s`).  Frames are encoded by the reference's encoder (oracle/_ref)."""
import importlib, os, sys
import numpy as np
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
uv = importlib.import_module("universal-volumetric_b200")
import corto_bind
from tools.synth import synth
rings, segs = synth.sphere_dims(50000); fp, _, _, _ = synth.sphere_topology(rings, segs)
blobs = []
for i in range(4):
    pos = synth.sphere_frame(rings, segs, i / 30.0, 20260005)
    uvv = np.stack([np.arctan2(pos[:, 2], pos[:, 0]) / (2 * np.pi) + 0.5, pos[:, 1] / 2000.0 + 0.5], 1).astype(np.float32)
    blobs.append(corto_bind.ref_encode(pos, uvv, fp, 12, 12)[0])
frames = [blobs[i % 4] for i in range(64)]
ctx = uv.Context(0); dec = uv.CortoDecoder(ctx)
for _ in range(2):
    out = dec.decode_batch_raw(frames, uv.MEM_DEVICE)
print("ok", sum(m.status == 0 for m in out[:64]))

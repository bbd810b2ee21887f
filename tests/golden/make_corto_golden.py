"""Generates tests/golden/corto/*.crt with the reference's own encoder (oracle/_ref/libcorto_ref.so) from seeded
procedural meshes, and tests/golden/corto_expected.json with digests of the reference decoder's outputs.
Run in the build container (needs /root/reference): python tests/golden/make_corto_golden.py"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from corto_bind import ref_decode, ref_encode  # noqa: E402
from tools.synth import synth  # noqa: E402

d = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]
os.makedirs(os.path.join(ROOT, "tests/golden/corto"), exist_ok=True)
exp = {}
for name, nverts, pb, ub in (("sphere_500", 500, 11, 10), ("sphere_6k", 6000, 12, 12)):
    rings, segs = synth.sphere_dims(nverts)
    fp, fu, uv, nv = synth.sphere_topology(rings, segs)
    pos = synth.sphere_frame(rings, segs, 0.2, 5)
    uvv = np.stack([np.arctan2(pos[:, 2], pos[:, 0]) / (2 * np.pi) + 0.5, pos[:, 1] / 2000.0 + 0.5], 1).astype(np.float32)   # per-vertex uv (V1 has no seams)
    blob, ev, ef = ref_encode(pos, uvv, fp, pb, ub)
    idx, p, u = ref_decode(blob, ev, ef)
    open(os.path.join(ROOT, "tests/golden/corto", name + ".crt"), "wb").write(blob)
    exp[name + ".crt"] = {"nvert": ev, "nface": ef, "index": d(idx), "position": d(p), "uv": d(u)}
json.dump(exp, open(os.path.join(ROOT, "tests/golden/corto_expected.json"), "w"), indent=1, sort_keys=True)
print(json.dumps(exp, indent=1))

"""Regenerates tests/golden/liam_expected.json: digests of the oracle's outputs on the committed
fixture files.  Run from the repo root in the build container: python tests/golden/make_golden.py
The fixture files themselves are DATA copied from the reference's example/public/liam/output/
(geometry_draco/00000,00001,00137.drc; texture_ktx2-fps30-1k_baseColor_default/00000.ktx2)."""
import glob
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_bind import oracle_draco, oracle_ktx2  # noqa: E402

d = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]
out = {"draco": {}, "ktx2": {}}
for p in sorted(glob.glob(os.path.join(ROOT, "tests/golden/liam/*.drc"))):
    o = oracle_draco(open(p, "rb").read())
    out["draco"][os.path.basename(p)] = {"num_points": o["num_points"], "num_faces": o["num_faces"], "index": d(o["index"]),
                                         "position": d(o["position"]), "normal": d(o["normal"]), "uv": d(o["uv"])}
for p in sorted(glob.glob(os.path.join(ROOT, "tests/golden/liam/*.ktx2"))):
    o = oracle_ktx2(open(p, "rb").read())
    out["ktx2"][os.path.basename(p)] = {"rgba": d(o["rgba"])}
json.dump(out, open(os.path.join(ROOT, "tests/golden/liam_expected.json"), "w"), indent=1, sort_keys=True)
print(json.dumps(out, indent=1))

"""Writes tests/golden/uastc_expected.json: digests of a seeded synthetic UASTC KTX2 (all 19 modes) and of the CPU
oracle's RGBA32 output for it.  Run from the repo root: python tests/golden/make_golden_uastc.py"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_bind import oracle_ktx2  # noqa: E402
from test_oracle_uastc import _file  # noqa: E402

args = {"seed": 11, "size": 48, "layers": 2}
_, blob = _file(**args)
out = {"args": args, "ktx2_sha256": hashlib.sha256(blob).hexdigest(), "rgba_sha256": hashlib.sha256(oracle_ktx2(blob)["rgba"].tobytes()).hexdigest()}
with open(os.path.join(ROOT, "tests", "golden", "uastc_expected.json"), "w") as fh:
    json.dump(out, fh, indent=1)
print(out)

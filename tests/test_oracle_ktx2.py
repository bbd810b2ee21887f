"""CPU: pins the KTX2 / BasisLZ oracle to the reference's .ktx2 fixtures with the B.4 oracles and to
committed golden digests."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, fixture_ktx2, golden_ktx2, read
from oracle_bind import oracle_ktx2

FILES = golden_ktx2() + fixture_ktx2()[1::12]


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(p) for p in FILES])
def test_b4_oracles(built, path):
    o = oracle_ktx2(read(path), keep_debug=True)
    assert o["status"] == 0 and o["is_video"] and not o["is_uastc"] and not o["has_alpha"]
    assert (o["width"], o["height"], o["layers"]) == (1024, 1024, 5)
    for total, used in o["sections"]:
        assert total == used                                       # section byte counts consumed exactly
    assert o["slices"][0] == o["slices"][1] == 5                   # every slice consumes exactly its bytes
    assert o["endpoint_idx"].max() < o["endpoint_count"] and o["selector_idx"].max() < o["selector_count"]
    assert o["dfd_transfer"] == 2
    assert (o["rgba"][..., 3] == 255).all()


def test_segment0_known_counts(built):
    o = oracle_ktx2(read(os.path.join(GOLDEN, "liam", "00000.ktx2")))
    assert (o["endpoint_count"], o["selector_count"]) == (1506, 734)
    assert o["sections"] == ((2789, 2789), (1598, 1598), (915, 915))
    assert o["pred_hist"] == [59663, 27995, 178665, 61357]


def test_golden_digest(built):
    exp = json.load(open(os.path.join(GOLDEN, "liam_expected.json")))
    for name, e in exp["ktx2"].items():
        o = oracle_ktx2(read(os.path.join(GOLDEN, "liam", name)))
        assert hashlib.sha256(o["rgba"].tobytes()).hexdigest()[:16] == e["rgba"], name


@pytest.mark.parametrize("mutation", ["truncate", "magic", "flip"])
def test_malformed_inputs(built, mutation):
    blob = bytearray(read(os.path.join(GOLDEN, "liam", "00000.ktx2")))
    if mutation == "truncate":
        blob = blob[: len(blob) // 2]
    elif mutation == "magic":
        blob[1] = 0
    else:
        blob[len(blob) // 2] ^= 0xFF
    o = oracle_ktx2(bytes(blob))
    assert o["status"] in (0, -1, -2, -3)
    if mutation != "flip":
        assert o["status"] < 0

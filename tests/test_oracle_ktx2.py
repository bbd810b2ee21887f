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


def test_etc1_target_decodes_to_the_oracle_texels(built):
    """Target format ETC1 (src/lib/KTX2Loader.js:619-636): the product's ETC1S -> ETC1 repack (csrc/basis_core.h, run on the host by
    tests/tools/basis_emu.cpp), decoded by an independent ETC1 decoder written from the Khronos format description, gives exactly
    the oracle's RGBA32 texels -- on a real fixture segment and on synthetic video segments."""
    import sys
    from conftest import ROOT, golden_ktx2, read
    sys.path.insert(0, ROOT)
    from emu_bind import emu_ktx2_etc1
    from etc1_decode import decode_etc1
    from tools.synth import synth
    files = [read(golden_ktx2()[0]), synth.encode_etc1s(synth.texture_layers(64, 0, 3, 4)), synth.encode_etc1s(synth.texture_layers(8, 0, 1, 2))]
    for f in files:
        e, o = emu_ktx2_etc1(f), oracle_ktx2(f)
        assert e["status"] == 0 and o["status"] == 0
        rgba = o["rgba"].reshape(o["layers"], o["height"], o["width"], 4)
        for L in range(e["layers"]):
            assert np.array_equal(decode_etc1(e["blocks"][L], e["width"], e["height"]), rgba[L])
        assert (e["blocks"][..., 3] & 3 == 3).all() and (e["blocks"][..., :3] & 7 == 0).all()        # differential, flipped, zero deltas

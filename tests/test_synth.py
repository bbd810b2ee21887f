"""CPU: the synthetic generators produce streams that the oracle decodes back to the input
(oracle(decode(encode(m))) == m up to quantisation), and the product logic agrees bit-exactly."""
import os
import sys

import numpy as np
import pytest
from scipy.spatial import cKDTree

from conftest import ROOT
from emu_bind import emu_draco, emu_ktx2
from oracle_bind import oracle_draco, oracle_ktx2

sys.path.insert(0, ROOT)
from tools.synth import synth  # noqa: E402


@pytest.mark.parametrize("nverts,qp", [(60, 11), (500, 11), (5000, 12), (50000, 16)])
def test_draco_roundtrip(built, nverts, qp):
    rings, segs = synth.sphere_dims(nverts)
    fp, fu, uv, nv = synth.sphere_topology(rings, segs)
    pos = synth.sphere_frame(rings, segs, 0.4, 3); nrm = synth.vertex_normals(pos, fp)
    blob = synth.encode_draco(pos, fp, uv, fu, nrm, qp=qp)
    o = oracle_draco(blob)
    assert o["status"] == 0 and o["ctx_counters_zero"] and o["rans_terminal_ok"] and o["bytes_consumed"] == len(blob) and o["uv_orient_left"] == 0
    assert o["num_faces"] == len(fp) and o["num_vertices"] == nv and o["num_points"] > nv      # the UV seam splits points
    step = np.ptp(pos, 0).max() / ((1 << qp) - 1)
    d, nn = cKDTree(pos).query(o["position"])
    assert d.max() <= step * 0.87 + 1e-3
    norm = lambda f: np.stack([np.take_along_axis(f, ((f.argmin(1) + k) % 3)[:, None], 1)[:, 0] for k in range(3)], 1)
    assert set(map(tuple, norm(nn[o["index"].reshape(-1, 3)]))) == set(map(tuple, norm(fp.astype(np.int64))))
    assert (o["normal"] * nrm[nn]).sum(1).min() > 0.995
    e = emu_draco(blob)
    assert e["status"] == 0 and np.array_equal(e["index"], o["index"])
    for k in ("position", "normal", "uv"):
        assert np.array_equal(e[k].view(np.uint32), o[k].view(np.uint32)), k


@pytest.mark.parametrize("size,layers", [(8, 1), (64, 3), (256, 7)])
def test_etc1s_roundtrip(built, size, layers):
    img = synth.texture_layers(size, 0, layers, 9)
    blob = synth.encode_etc1s(img)
    o = oracle_ktx2(blob)
    assert o["status"] == 0 and o["is_video"] and (o["width"], o["height"], o["layers"]) == (size, size, layers)
    for total, used in o["sections"]:
        assert total == used
    assert o["slices"] == (layers, layers)
    if size >= 64:
        assert np.abs(o["rgba"][..., :3].astype(int) - img[..., :3].astype(int)).mean() < 16     # lossy, but the same picture
    e = emu_ktx2(blob)
    assert e["status"] == 0 and np.array_equal(e["rgba"], o["rgba"])


def test_sequence_is_deterministic(built):
    a = synth.make_sequence(3, 300, 32, sequence_size=2, seed=11, threads=2)
    b = synth.make_sequence(3, 300, 32, sequence_size=2, seed=11, threads=1)
    assert a[0] == b[0] and a[1] == b[1] and a[2]["segments"] == 2

"""Parity against UPSTREAM Draco / Basis outputs, when somebody has produced them with tests/golden/make_upstream_golden.sh (the build
image has neither binaries nor network, so these tests skip there).  OBJ text carries ~6 significant digits, hence the tolerance on the
float attributes; connectivity is compared exactly, as the multiset of faces over position / uv / normal VALUES (upstream's OBJ writer
de-duplicates attribute values, so its indices are not point ids).  Texels are compared exactly."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_drc, golden_ktx2, read
from oracle_bind import oracle_draco, oracle_ktx2

UP = os.path.join(GOLDEN, "upstream")


def parse_obj(path):
    v, vt, vn, faces = [], [], [], []
    for line in open(path):
        p = line.split()
        if not p:
            continue
        if p[0] == "v":
            v.append([float(x) for x in p[1:4]])
        elif p[0] == "vt":
            vt.append([float(x) for x in p[1:3]])
        elif p[0] == "vn":
            vn.append([float(x) for x in p[1:4]])
        elif p[0] == "f":
            faces.append([[int(i) - 1 if i else -1 for i in (c.split("/") + ["", ""])[:3]] for c in p[1:4]])
    f = np.array(faces, np.int64)
    corners = np.concatenate([np.array(v)[f[..., 0]], np.array(vt)[f[..., 1]], np.array(vn)[f[..., 2]]], axis=-1)      # [F, 3, 8]
    return corners


def corners_of(mesh):
    idx = mesh["index"].reshape(-1, 3)
    return np.concatenate([mesh["position"][idx], mesh["uv"][idx], mesh["normal"][idx]], axis=-1)


def canon(c, grid_from):
    """Faces as rows in a canonical order.  Positions and uvs are snapped to the value grids of `grid_from` (our own decode: every
    true value lies exactly on a grid point, and 6-digit OBJ text is far closer to its grid point than to a neighbour), which gives
    integer keys; corners are rotated so the smallest key comes first and the faces are sorted by key.  Returns (keys, corner values)."""
    keys = np.zeros(c.shape[:2] + (5,), np.int64)
    for k in range(5):
        grid = np.unique(grid_from[..., k])
        i = np.clip(np.searchsorted(grid, c[..., k]), 1, len(grid) - 1)
        keys[..., k] = np.where(np.abs(grid[i - 1] - c[..., k]) <= np.abs(grid[i] - c[..., k]), i - 1, i)
    flat = keys[..., 0] * 0
    for k in range(5):
        flat = flat * 4096 + keys[..., k]                         # grids have <= 2048 / 1024 points per component
    rots = np.stack([np.roll(flat, -r, axis=1) for r in range(3)], axis=1)                  # [F, rotation, corner]
    best = np.zeros(len(c), np.int64)
    for r in (1, 2):                                                  # lexicographically smallest rotation (ties: corners sharing a key)
        a, b = rots[np.arange(len(c)), best], rots[:, r]
        less = (b[:, 0] < a[:, 0]) | ((b[:, 0] == a[:, 0]) & ((b[:, 1] < a[:, 1]) | ((b[:, 1] == a[:, 1]) & (b[:, 2] < a[:, 2]))))
        best = np.where(less, r, best)
    rolled_k = rots[np.arange(len(c)), best]
    rolled_c = np.stack([np.roll(c[i], -best[i], axis=0) for i in range(len(c))])
    # degenerate faces (two corners with the same position and uv, told apart only by their normals): the rotation is ambiguous, order
    # their corners by (key, normal) outright
    tied = (rolled_k[:, 0] == rolled_k[:, 1]) | (rolled_k[:, 1] == rolled_k[:, 2]) | (rolled_k[:, 0] == rolled_k[:, 2])
    for i in np.nonzero(tied)[0]:
        o = sorted(range(3), key=lambda k: (rolled_k[i, k],) + tuple(rolled_c[i, k, 5:]))
        rolled_k[i] = rolled_k[i, o]; rolled_c[i] = rolled_c[i, o]
    nkey = np.round(rolled_c[..., 5:].reshape(len(c), -1) * 20).astype(np.int64)             # coarse normal key: orders faces that share positions and uvs
    order = np.lexsort(np.concatenate([rolled_k, nkey], axis=1).T[::-1])
    return rolled_k[order], rolled_c[order]


def same_mesh(want_corners, got_corners):
    kw, cw = canon(want_corners, got_corners); kg, cg = canon(got_corners, got_corners)
    return kw.shape == kg.shape and np.array_equal(kw, kg) and np.allclose(cw, cg, rtol=2e-5, atol=2e-4)


@pytest.mark.parametrize("path", golden_drc())
def test_oracle_geometry_matches_upstream(path):
    obj = os.path.join(UP, os.path.basename(path).replace(".drc", ".obj"))
    if not os.path.exists(obj):
        pytest.skip("no upstream golden (run tests/golden/make_upstream_golden.sh where draco_decoder exists)")
    assert same_mesh(parse_obj(obj), corners_of(oracle_draco(read(path))))


@pytest.mark.parametrize("path", golden_ktx2())
def test_oracle_texels_match_upstream(path):
    import cv2
    pngs = sorted(glob.glob(os.path.join(UP, os.path.basename(path).replace(".ktx2", "_layer*.png"))))
    if not pngs:
        pytest.skip("no upstream golden (run tests/golden/make_upstream_golden.sh where basisu exists)")
    o = oracle_ktx2(read(path))
    assert len(pngs) == o["layers"]
    for L, p in enumerate(pngs):
        img = cv2.cvtColor(cv2.imread(p, cv2.IMREAD_UNCHANGED), cv2.COLOR_BGRA2RGBA)
        assert np.array_equal(img, o["rgba"][L])


@pytest.mark.gpu
def test_gpu_matches_upstream(uv, ctx):
    objs = [os.path.join(UP, os.path.basename(p).replace(".drc", ".obj")) for p in golden_drc()]
    if not all(os.path.exists(o) for o in objs):
        pytest.skip("no upstream golden")
    res = uv.DRACOLoader(ctx).decode_batch([read(p) for p in golden_drc()])
    for r, obj in zip(res, objs):
        m = {"index": r["index"], "position": r["attributes"]["position"], "uv": r["attributes"]["uv"], "normal": r["attributes"]["normal"]}
        assert same_mesh(parse_obj(obj), corners_of(m))

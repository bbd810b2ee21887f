"""Deterministic synthetic UVOL V2 sequences (bench/test INPUT tooling).

BASELINE.json's configs need 50k / 200k-vertex Draco frames and 1024^2 / 2048^2 KTX2 segments; the
reference ships only the 26k-vertex `liam` fixtures and no encoder binaries exist here
(SURVEY.md 7.2-3, 8d).  This module builds a deforming genus-0 surface (UV sphere with a UV seam,
so points > vertices like real captures) and a drifting procedural texture, and encodes them with
tools/synth/libuvsynth.so (draco_encode.cpp, etc1s_encode.cpp).  RNG seed = 20260001 + config id.
"""
import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libuvsynth.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-s", "-C", _HERE], check=True)
        L = ctypes.CDLL(path)
        P = ctypes.POINTER
        L.uvsynth_draco_encode.argtypes = [P(ctypes.c_float), ctypes.c_uint32, P(ctypes.c_uint32), ctypes.c_uint32, P(ctypes.c_float),
                                           ctypes.c_uint32, P(ctypes.c_uint32), P(ctypes.c_float), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           P(P(ctypes.c_uint8))]
        L.uvsynth_draco_encode.restype = ctypes.c_size_t
        L.uvsynth_etc1s_encode.argtypes = [P(ctypes.c_uint8), ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, P(P(ctypes.c_uint8))]
        L.uvsynth_etc1s_encode.restype = ctypes.c_size_t
        L.uvsynth_uastc_encode.argtypes = [P(ctypes.c_uint8), ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                           ctypes.c_int, P(P(ctypes.c_uint8))]
        L.uvsynth_uastc_encode.restype = ctypes.c_size_t
        L.uvsynth_free.argtypes = [ctypes.c_void_p]
        _lib = L
    return _lib


def sphere_dims(target_verts):
    """(rings, segments) of a UV sphere with about `target_verts` vertices."""
    segs = int(round((2.0 * target_verts) ** 0.5))
    rings = max(3, int(round(target_verts / segs)) + 1)
    return rings, segs


def sphere_topology(rings, segs):
    """Faces as position indices and per-corner uv indices (seam at u=0/1, split pole uvs)."""
    R, S = rings, segs
    nv = (R - 1) * S + 2
    north, south = 0, nv - 1
    ring = lambda i, j: 1 + (i - 1) * S + (j % S)                       # i in 1..R-1
    uvr = lambda i, j: (i - 1) * (S + 1) + j                             # j in 0..S
    uvn = lambda j: (R - 1) * (S + 1) + j
    uvs = lambda j: (R - 1) * (S + 1) + S + j
    j = np.arange(S)
    fp, fu = [], []
    fp.append(np.stack([np.full(S, north), ring(1, j), ring(1, j + 1)], 1)); fu.append(np.stack([uvn(j), uvr(1, j), uvr(1, j + 1)], 1))
    for i in range(1, R - 1):
        a, b, c, d = ring(i, j), ring(i + 1, j), ring(i + 1, j + 1), ring(i, j + 1)
        ua, ub, uc, ud = uvr(i, j), uvr(i + 1, j), uvr(i + 1, j + 1), uvr(i, j + 1)
        fp.append(np.stack([a, b, c], 1)); fu.append(np.stack([ua, ub, uc], 1))
        fp.append(np.stack([a, c, d], 1)); fu.append(np.stack([ua, uc, ud], 1))
    fp.append(np.stack([np.full(S, south), ring(R - 1, j + 1), ring(R - 1, j)], 1)); fu.append(np.stack([uvs(j), uvr(R - 1, j + 1), uvr(R - 1, j)], 1))
    faces_pos = np.concatenate(fp).astype(np.uint32)
    faces_uv = np.concatenate(fu).astype(np.uint32)
    ii, jj = np.meshgrid(np.arange(1, R), np.arange(S + 1), indexing="ij")
    uv = np.concatenate([np.stack([jj / S, ii / R], -1).reshape(-1, 2),
                         np.stack([(j + 0.5) / S, np.zeros(S)], 1), np.stack([(j + 0.5) / S, np.ones(S)], 1)]).astype(np.float32)
    uv = uv * 0.96 + 0.02
    return faces_pos, faces_uv, uv, nv


def sphere_frame(rings, segs, t, seed):
    """Positions (millimetre scale like the fixtures) and smooth normals of frame time t."""
    R, S = rings, segs
    rng = np.random.default_rng(seed)
    noise = rng.standard_normal(((R - 1) * S + 2)).astype(np.float64) * 0.004
    theta = (np.arange(1, R) / R * np.pi)[:, None]
    phi = (np.arange(S) / S * 2 * np.pi)[None, :]

    def radius(th, ph):
        return (1.0 + 0.22 * np.sin(2 * th + 0.7 * t) * np.cos(3 * ph + 0.4 * t) + 0.07 * np.sin(5 * ph - t) * np.sin(4 * th)
                + 0.03 * np.sin(9 * th + 1.3 * t) * np.cos(7 * ph))

    def point(th, ph):
        r = radius(th, ph)
        return np.stack([r * np.sin(th) * np.cos(ph), r * np.cos(th), r * np.sin(th) * np.sin(ph)], -1)

    body = point(theta, phi).reshape(-1, 3)
    pos = np.concatenate([point(np.array(0.0), np.array(0.0))[None], body, point(np.array(np.pi), np.array(0.0))[None]])
    pos = pos * (1.0 + noise[:, None] * (1.0 + 0.5 * np.sin(t)))
    return (pos * 850.0).astype(np.float32)


def vertex_normals(pos, faces):
    v = pos[faces].astype(np.float64)
    fn = np.cross(v[:, 1] - v[:, 0], v[:, 2] - v[:, 0])
    n = np.zeros((len(pos), 3))
    for k in range(3):
        np.add.at(n, faces[:, k], fn)
    n /= np.linalg.norm(n, axis=1, keepdims=True) + 1e-30
    return n.astype(np.float32)


def encode_draco(pos, faces_pos, uv, faces_uv, nrm, qp=11, qt=10, qn=8, tagged=0):
    """tagged: bit mask of the attributes written with Draco's TAGGED symbol scheme (1 position, 2 uv, 4 normal) instead of RAW."""
    L = lib()
    L.uvsynth_draco_tagged(int(tagged))
    pos = np.ascontiguousarray(pos, np.float32); uv = np.ascontiguousarray(uv, np.float32); nrm = np.ascontiguousarray(nrm, np.float32)
    fp = np.ascontiguousarray(faces_pos, np.uint32); fu = np.ascontiguousarray(faces_uv, np.uint32)
    out = ctypes.POINTER(ctypes.c_uint8)()
    f32, u32 = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_uint32)
    n = L.uvsynth_draco_encode(pos.ctypes.data_as(f32), len(pos), fp.ctypes.data_as(u32), len(fp), uv.ctypes.data_as(f32), len(uv),
                               fu.ctypes.data_as(u32), nrm.ctypes.data_as(f32), qp, qt, qn, ctypes.byref(out))
    if n == 0:
        raise RuntimeError("uvsynth_draco_encode failed (input must be one closed manifold genus-0 component)")
    blob = ctypes.string_at(out, n)
    L.uvsynth_free(out)
    L.uvsynth_draco_tagged(0)
    return blob


def texture_layers(size, first_frame, count, seed):
    """`count` RGBA layers of a smooth drifting pattern over a static background (so P-frames keep
    a realistic share of unchanged (CR) blocks)."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:size, 0:size].astype(np.float32) / size
    base = np.stack([0.5 + 0.5 * np.sin(6.3 * (x * (1 + k) + 0.3 * k) + 2.1 * y * (2 - k)) for k in range(3)], -1)
    detail = 0.10 * np.sin(size / 9.0 * x + 3.0 * np.sin(5.0 * y)) * np.sin(size / 11.0 * y + 2.0 * np.cos(4.0 * x))
    grain = (rng.random((size, size), dtype=np.float32) - 0.5) * 0.06
    static = base * 0.7 + (detail + grain)[..., None]
    out = np.empty((count, size, size, 4), np.uint8)
    for i in range(count):
        t = (first_frame + i) / 30.0
        img = static.copy()
        for k in range(3):      # three drifting soft-edged features: about a third of the blocks change per frame
            cx, cy = 0.5 + 0.32 * np.cos(0.9 * t + 2.1 * k), 0.5 + 0.32 * np.sin(1.3 * t + 1.7 * k)
            blob = np.exp(-(((x - cx) ** 2 + (y - cy) ** 2) / 0.012))
            blob[blob < 0.02] = 0
            img += blob[..., None] * np.array([0.45, 0.15, -0.35], np.float32) * (1.0 if k != 1 else -1.0)
        out[i, ..., :3] = np.clip(img * 255.0, 0, 255).astype(np.uint8)
        out[i, ..., 3] = 255
    return out


def encode_etc1s(layers_rgba, max_endpoints=4096):
    L = lib()
    a = np.ascontiguousarray(layers_rgba, np.uint8)
    out = ctypes.POINTER(ctypes.c_uint8)()
    n = L.uvsynth_etc1s_encode(a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), a.shape[2], a.shape[1], a.shape[0], max_endpoints, ctypes.byref(out))
    if n == 0:
        raise RuntimeError("uvsynth_etc1s_encode failed")
    blob = ctypes.string_at(out, n)
    L.uvsynth_free(out)
    return blob


UASTC_OPAQUE_MODES = 0x401FF          # modes 0-8 and 18: what an encoder picks for opaque RGB content
UASTC_ALL_MODES = 0x7FFFF


def encode_uastc(layers_rgba, mode_mask=UASTC_OPAQUE_MODES, seed=1, has_alpha=False):
    """KTX2 (UASTC, no supercompression) of `layers_rgba` [layers, h, w, 4]; any width / height."""
    L = lib()
    a = np.ascontiguousarray(layers_rgba, np.uint8)
    out = ctypes.POINTER(ctypes.c_uint8)()
    n = L.uvsynth_uastc_encode(a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), a.shape[2], a.shape[1], a.shape[0], mode_mask, seed, int(has_alpha),
                               ctypes.byref(out))
    if n == 0:
        raise RuntimeError("uvsynth_uastc_encode failed")
    blob = ctypes.string_at(out, n)
    L.uvsynth_free(out)
    return blob


def make_sequence(frames, verts, tex_size, sequence_size=7, seed=20260002, threads=None, want_textures=True, distinct_geometry=None,
                  distinct_textures=None, texture_format="etc1s"):
    """Returns (list of .drc bytes, list of .ktx2 bytes, info).  `distinct_geometry` / `distinct_textures` bound the number
    of distinct frames / segments that are actually encoded (the rest cycle through them), to bound generation time.
    texture_format: "etc1s" (BasisLZ video) or "uastc"."""
    rings, segs = sphere_dims(verts)
    fp, fu, uv, nv = sphere_topology(rings, segs)
    ng = frames if distinct_geometry is None else min(frames, distinct_geometry)
    threads = threads or min(32, os.cpu_count() or 1)

    def one_geo(i):
        pos = sphere_frame(rings, segs, i / 30.0, seed)
        return encode_draco(pos, fp, uv, fu, vertex_normals(pos, fp))

    def one_tex(s):
        first = s * sequence_size
        cnt = min(sequence_size, frames - first)
        layers = texture_layers(tex_size, first, cnt, seed + 7)
        return encode_uastc(layers, seed=seed + s) if texture_format == "uastc" else encode_etc1s(layers)

    nseg = (frames + sequence_size - 1) // sequence_size
    nt = nseg if distinct_textures is None else min(nseg, distinct_textures)
    todo = list(range(nt)) + ([nseg - 1] if nt < nseg and frames % sequence_size else [])     # a short last segment is encoded as itself
    with ThreadPoolExecutor(threads) as ex:
        geo = list(ex.map(one_geo, range(ng)))
        enc = dict(zip(todo, ex.map(one_tex, todo))) if want_textures else {}
    tex = [enc.get(s, enc.get(s % nt)) for s in range(nseg)] if want_textures else []
    drc = [geo[i % ng] for i in range(frames)]
    info = {"verts": nv, "faces": len(fp), "rings": rings, "segs": segs, "frames": frames, "segments": nseg, "sequence_size": sequence_size,
            "tex_size": tex_size, "distinct_geometry": ng, "distinct_textures": nt, "texture_format": texture_format, "seed": seed}
    return drc, tex, info

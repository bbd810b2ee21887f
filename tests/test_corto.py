"""V1 Corto path.  CPU: the golden .crt vectors (made by the reference's own encoder, decoded by the reference's own
decoder, digests committed) still decode to the committed digests when oracle/_ref is present.  GPU: the CUDA path
(batch API and the reference's CreateDecoder / DecodeMesh / DestroyDecoder ABI) is bit-exact against them."""
import ctypes
import hashlib
import json
import os
import sys

import numpy as np
import pytest

import corto_bind
from conftest import GOLDEN, ROOT, read

sys.path.insert(0, ROOT)
from tools.synth import synth  # noqa: E402

EXP = json.load(open(os.path.join(GOLDEN, "corto_expected.json")))
d = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


@pytest.mark.skipif(not corto_bind.available(), reason="oracle/_ref/libcorto_ref.so not built (needs the reference tree)")
def test_reference_build_reproduces_golden():
    for name, e in EXP.items():
        idx, pos, uv = corto_bind.ref_decode(read(os.path.join(GOLDEN, "corto", name)), e["nvert"], e["nface"])
        assert (d(idx), d(pos), d(uv)) == (e["index"], e["position"], e["uv"])


@pytest.mark.gpu
def test_gpu_matches_golden(uv, ctx):
    blobs = [read(os.path.join(GOLDEN, "corto", n)) for n in EXP]
    for r, e in zip(uv.CortoDecoder(ctx).decode_batch(blobs), EXP.values()):
        assert r["status"] == 0
        assert (d(r["index"]), d(r["position"]), d(r["uv"])) == (e["index"], e["position"], e["uv"])


@pytest.mark.gpu
@pytest.mark.skipif(not corto_bind.available(), reason="oracle/_ref/libcorto_ref.so not staged")
@pytest.mark.parametrize("nverts,pb,ub,with_uv", [(60, 10, 10, True), (3000, 12, 12, True), (21000, 12, 12, True), (50000, 14, 12, True), (3000, 12, 12, False)])
def test_gpu_matches_reference(uv, ctx, nverts, pb, ub, with_uv):
    rings, segs = synth.sphere_dims(nverts)
    fp, fu, uvs, nv = synth.sphere_topology(rings, segs)
    pos = synth.sphere_frame(rings, segs, 0.7, 21)
    uvv = np.stack([np.arctan2(pos[:, 2], pos[:, 0]) / (2 * np.pi) + 0.5, pos[:, 1] / 2000.0 + 0.5], 1).astype(np.float32) if with_uv else None
    blob, ev, ef = corto_bind.ref_encode(pos, uvv, fp, pb, ub)
    idx, p, u = corto_bind.ref_decode(blob, ev, ef, with_uv)
    r = uv.CortoDecoder(ctx).decode_batch([blob, blob[: len(blob) // 2], b"\0" * 64])
    assert r[0]["status"] == 0 and r[1]["status"] < 0 and r[2]["status"] < 0
    assert np.array_equal(r[0]["index"], idx)                                       # bit-exact connectivity
    assert np.array_equal(r[0]["position"].view(np.uint32), p.view(np.uint32))       # (float)int * q: single rounding, bit-exact
    if with_uv:
        assert np.array_equal(r[0]["uv"].view(np.uint32), u.view(np.uint32))


@pytest.mark.gpu
def test_reference_c_abi(uv):
    """CreateDecoder / DecodeMesh / DestroyDecoder as the Unity loader calls them (CortoMeshLoader.cs:13-28)."""
    L = uv._native.lib(); name, e = next(iter(EXP.items())); blob = read(os.path.join(GOLDEN, "corto", name))
    info = (uv._native.Vector2 * 1)()
    h = L.CreateDecoder(len(blob), blob, info)
    assert h and int(info[0].x) == e["nface"] and int(info[0].y) == e["nvert"]
    pos = np.zeros((e["nvert"], 3), np.float32); uvs = np.zeros((e["nvert"], 2), np.float32); idx = np.zeros(e["nface"] * 3, np.int32)
    rc = L.DecodeMesh(h, pos.ctypes.data_as(ctypes.c_void_p), idx.ctypes.data_as(ctypes.c_void_p), None, None, uvs.ctypes.data_as(ctypes.c_void_p))
    L.DestroyDecoder(h)
    assert rc == e["nface"]
    assert (d(idx.view(np.uint32)), d(pos), d(uvs)) == (e["index"], e["position"], e["uv"])
    assert not L.CreateDecoder(16, b"\0" * 16, info)


@pytest.mark.gpu
@pytest.mark.skipif(not corto_bind.available(), reason="oracle/_ref/libcorto_ref.so not staged")
def test_v1_sequence_from_disk(uv, ctx, tmp_path):
    """A V1 clip on disk (`.manifest` + `.drcs` as deprecated/encoder/src/Encoder30.js writes them, frames encoded by the reference's own
    encoder) decoded through V1Sequence: per keyframe the worker's bufferGeometry, bit-exact against the reference decoder."""
    import json
    rings, segs = synth.sphere_dims(3000); fp, fu, uvs, nv = synth.sphere_topology(rings, segs)
    enc = []
    for i in range(5):
        pos = synth.sphere_frame(rings, segs, i / 30.0, 33)
        uvv = np.stack([np.arctan2(pos[:, 2], pos[:, 0]) / (2 * np.pi) + 0.5, pos[:, 1] / 2000.0 + 0.5], 1).astype(np.float32)
        enc.append(corto_bind.ref_encode(pos, uvv, fp, 12, 12))
    (tmp_path / "clip.drcs").write_bytes(b"".join(b for b, _, _ in enc))
    (tmp_path / "clip.manifest").write_text(json.dumps(uv.emit_v1(30, [(ev, ef, len(b)) for b, ev, ef in enc])))
    out = uv.V1Sequence(str(tmp_path / "clip.manifest"), uv.CortoDecoder(ctx)).decode(1, 5)
    assert sorted(out) == [1, 2, 3, 4]
    for k, r in out.items():
        b, ev, ef = enc[k]
        idx, p, u = corto_bind.ref_decode(b, ev, ef)
        g = r["bufferGeometry"]
        assert r["frameNumber"] == k and np.array_equal(g["index"], idx)
        assert np.array_equal(g["position"].view(np.uint32), p.view(np.uint32)) and np.array_equal(g["uv"].view(np.uint32), u.view(np.uint32))

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


def _holey_grid(g=40, seed=1):
    """An open, shuffled grid with missing triangles: the encoder answers with BOUNDARY, DELAY, SPLIT and END symbols."""
    rng = np.random.default_rng(seed)
    xs, ys = np.meshgrid(np.arange(g), np.arange(g))
    pos = np.stack([xs.ravel() * 10.0, ys.ravel() * 10.0, rng.random(g * g) * 30], 1).astype(np.float32)
    f = []
    for y in range(g - 1):
        for x in range(g - 1):
            a = y * g + x
            if rng.random() < 0.9:
                f.append([a, a + 1, a + g])
            if rng.random() < 0.9:
                f.append([a + 1, a + g + 1, a + g])
    f = np.array(f, np.uint32); rng.shuffle(f)
    return pos, f


@pytest.mark.skipif(not corto_bind.available(), reason="oracle/_ref/libcorto_ref.so not built (needs the reference tree)")
def test_connectivity_walk_host_logic():
    """The re-designed front walk (csrc/corto_core.h: gate edge and its ring neighbours in registers, 16-byte records) run on the host
    over the CLERS symbols the reference's own IndexAttribute::decode yields: faces and parallelogram contexts equal the reference
    decoder's, on closed meshes and on a holey grid that exercises every symbol (BOUNDARY, DELAY, SPLIT, END)."""
    import subprocess
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "tools"), "libcorto_emu.so"], check=True)
    E = ctypes.CDLL(os.path.join(ROOT, "tests", "tools", "libcorto_emu.so"))
    vp = ctypes.c_void_p
    E.corto_emu_walk.argtypes = [vp, ctypes.c_uint32, vp, ctypes.c_uint32, vp, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, vp, vp]
    cases = []
    for nverts in (60, 3000, 21000):
        rings, segs = synth.sphere_dims(nverts); fp, _, _, _ = synth.sphere_topology(rings, segs)
        cases.append((synth.sphere_frame(rings, segs, 0.2, 5), fp))
    cases.append(_holey_grid())
    seen = np.zeros(7, np.int64)
    for pos, faces in cases:
        blob, nv, nf = corto_bind.ref_encode(pos, None, faces, 12, 12)
        idx, pred = corto_bind.ref_prediction(blob, nv, nf)
        cl, words, ge = corto_bind.ref_clers(blob, nf)
        seen += np.bincount(cl, minlength=7)[:7]
        out = np.zeros(nf * 3, np.uint32); p4 = np.zeros((nv, 4), np.int32)
        rc = E.corto_emu_walk(cl.ctypes.data, len(cl), words.ctypes.data, len(words), ge.ctypes.data, len(ge), nv, nf, out.ctypes.data, p4.ctypes.data)
        full = (out.copy(), p4.copy())
        for ring in (4, 16, 128):      # tiny rings: the same walk with most records served from the global array
            E.corto_emu_set_ring(ring); out[:] = 0; p4[:] = 0
            assert E.corto_emu_walk(cl.ctypes.data, len(cl), words.ctypes.data, len(words), ge.ctypes.data, len(ge), nv, nf, out.ctypes.data, p4.ctypes.data) == rc
            assert np.array_equal(out, full[0]) and np.array_equal(p4, full[1]), ring
        E.corto_emu_set_ring(1024)
        assert rc == 0 and np.array_equal(out, idx) and np.array_equal(p4[1:, :3].astype(np.uint32), pred[1:])
        assert E.corto_emu_walk(cl.ctypes.data, len(cl) // 2, words.ctypes.data, len(words), ge.ctypes.data, len(ge), nv, nf, out.ctypes.data, p4.ctypes.data) < 0      # truncated symbols: an error, not a crash
    assert (seen > 0).all()                                                          # VERTEX LEFT RIGHT END BOUNDARY DELAY SPLIT all exercised


@pytest.mark.gpu
@pytest.mark.skipif(not corto_bind.available(), reason="oracle/_ref/libcorto_ref.so not staged")
@pytest.mark.parametrize("npred", [0, 1, 2])
def test_gpu_normals_and_colours_match_reference(uv, ctx, npred):
    """Corto normals (codec 2; prediction 0 DIFF, 1 ESTIMATED, 2 BORDER -- normal_attribute.cpp:168-303) and colours (codec 3,
    color_attribute.cpp:69-90) against the reference's own decoder: colours, indices and quantised paths bit-exact; normals bit-exact
    as well -- the per-vertex face-normal sums are taken in face order like the reference's accumulation."""
    cases = []
    rings, segs = synth.sphere_dims(3000); fp, _, _, _ = synth.sphere_topology(rings, segs)
    cases.append((synth.sphere_frame(rings, segs, 0.4, 9), fp))
    cases.append(_holey_grid(30, 4))                                                  # boundaries: BORDER prediction has work to do
    rng = np.random.default_rng(7)
    for pos, faces in cases:
        nrm = pos / np.linalg.norm(pos, axis=1, keepdims=True) + rng.normal(0, 0.05, pos.shape)
        nrm = (nrm / np.linalg.norm(nrm, axis=1, keepdims=True)).astype(np.float32)
        col = rng.integers(0, 256, (len(pos), 4), dtype=np.uint8)
        uvv = rng.random((len(pos), 2)).astype(np.float32)
        blob, nv, nf = corto_bind.ref_encode2(pos, faces, uv=uvv, normal=nrm, color=col, normal_bits=10, normal_pred=npred, color_bits=6)
        want = corto_bind.ref_decode2(blob, nv, nf, True, True, True)
        r = uv.CortoDecoder(ctx).decode_batch([blob])[0]
        assert r["status"] == 0 and np.array_equal(r["index"], want["index"])
        assert np.array_equal(r["position"].view(np.uint32), want["position"].view(np.uint32))
        assert np.array_equal(r["uv"].view(np.uint32), want["uv"].view(np.uint32))
        assert np.array_equal(r["color"], want["color"])
        assert r["normal"] is not None and np.array_equal(r["normal"].view(np.uint32), want["normal"].view(np.uint32))


@pytest.mark.gpu
@pytest.mark.skipif(not corto_bind.available(), reason="oracle/_ref/libcorto_ref.so not staged")
def test_gpu_every_symbol_and_u16_index(uv):
    """The holey grid (BOUNDARY / DELAY / SPLIT / END symbols) through the GPU walk, and the web player's index layout: with
    corto_index_u16 the index is a Uint16 array when nface < 65536 (corto.ts:675-680, src/V1/player.ts:292), u32 otherwise."""
    pos, faces = _holey_grid()
    blob, nv, nf = corto_bind.ref_encode(pos, None, faces, 12, 12)
    idx, p, _ = corto_bind.ref_decode(blob, nv, nf, False)
    rings, segs = synth.sphere_dims(50000); fp, _, _, _ = synth.sphere_topology(rings, segs)
    big, bv, bf = corto_bind.ref_encode(synth.sphere_frame(rings, segs, 0.1, 3), None, fp, 12, 12)
    bidx, _, _ = corto_bind.ref_decode(big, bv, bf, False)
    c = uv.Context(0, corto_index_u16=True)
    try:
        r = uv.CortoDecoder(c).decode_batch([blob, big])
        assert r[0]["status"] == 0 and r[0]["index"].dtype == np.uint16 and np.array_equal(r[0]["index"], idx.astype(np.uint16))
        assert np.array_equal(r[0]["position"].view(np.uint32), p.view(np.uint32))
        assert bf >= 65536 and r[1]["index"].dtype == np.uint32 and np.array_equal(r[1]["index"], bidx)
    finally:
        c.close()


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
    # the same range through the C++ host layer (uvol_open / uvol_decode_v1_range)
    L = uv._native.lib(); h = ctypes.c_void_p()
    assert L.uvol_open(ctx._h, str(tmp_path / "clip.manifest").encode(), ctypes.byref(h)) == 0
    meshes = (uv._native.CortoMesh * 4)(); keys = (ctypes.c_uint32 * 4)()
    assert L.uvol_decode_v1_range(h, 1, 4, uv.MEM_HOST, meshes, keys) == 0 and list(keys) == [1, 2, 3, 4]
    for k in range(4):
        idx, p, u = corto_bind.ref_decode(*enc[1 + k])
        assert meshes[k].status == 0 and np.array_equal(np.ctypeslib.as_array(meshes[k].index, (meshes[k].num_faces * 3,)), idx)
        assert np.array_equal(np.ctypeslib.as_array(meshes[k].position, (meshes[k].num_vertices, 3)).view(np.uint32), p.view(np.uint32))
    L.uvol_close(h)
    for k, r in out.items():
        b, ev, ef = enc[k]
        idx, p, u = corto_bind.ref_decode(b, ev, ef)
        g = r["bufferGeometry"]
        assert r["frameNumber"] == k and np.array_equal(g["index"], idx)
        assert np.array_equal(g["position"].view(np.uint32), p.view(np.uint32)) and np.array_equal(g["uv"].view(np.uint32), u.view(np.uint32))


@pytest.mark.gpu
def test_v1_frame_counter_and_etc2_target(uv, ctx):
    """SURVEY 8f-4.  (1) The V1 player's frame counter (src/V1/player.ts:305-334): frames painted exactly like example/texture_encoder.py:59-63
    paints them (16 cells of 8 x 8 pixels, bit i of the number in cell 15 - i ... read back least significant cell first), from device
    and from host memory, noise included.  (2) The 'etc2' target (src/V2/player.ts:338-356): raw block files are validated and uploaded."""
    import torch
    L = uv._native.lib(); W = H = 256; n = 40; cell = 8
    rng = np.random.default_rng(5)
    frames = rng.integers(0, 256, (n, H, W, 4), dtype=np.uint8)
    numbers = [0, 1, 2, 3, 255, 256, 4095, 65534] + [int(x) for x in rng.integers(0, 65535, n - 8)]
    for f, num in zip(frames, numbers):
        for i in range(16):                                    # the player reads cell i as bit i (value 2^i) and subtracts one
            on = ((num + 1) >> i) & 1
            f[H - cell // 2:, i * cell:(i + 1) * cell, :3] = np.clip((255 if on else 0) + rng.integers(-40, 41, (cell // 2, cell, 3)), 0, 255)
    out = (ctypes.c_int32 * n)()
    assert L.uvol_v1_frame_numbers(ctx._h, frames.ctypes.data, 0, n, W, H, cell, 16, out) == 0 and list(out) == numbers
    d = torch.from_numpy(frames).cuda()
    out2 = (ctypes.c_int32 * n)()
    assert L.uvol_v1_frame_numbers(ctx._h, d.data_ptr(), 1, n, W, H, cell, 16, out2) == 0 and list(out2) == numbers
    per = (W // 4) * (H // 4) * 8
    files = [bytes(rng.integers(0, 256, per, dtype=np.uint8)), b"\x00" * (per - 8), bytes(rng.integers(0, 256, per, dtype=np.uint8)), b"\x01" * (per + 8)]
    ptrs = (ctypes.c_void_p * 4)(*[ctypes.cast(ctypes.c_char_p(f), ctypes.c_void_p) for f in files]); sizes = (ctypes.c_size_t * 4)(*[len(f) for f in files])
    tex = (uv._native.Texture * 4)()
    assert L.uvol_upload_etc2_batch(ctx._h, ptrs, sizes, 4, W, H, uv.MEM_HOST, tex) == 0
    assert [t.status for t in tex] == [0, -1, 0, -2] and tex[0].format == uv._native.TEX_ETC2_RGB and tex[2].bytes == per
    assert bytes(np.ctypeslib.as_array(tex[0].data, (per,))) == files[0] and bytes(np.ctypeslib.as_array(tex[2].data, (per,))) == files[2]
    assert L.uvol_upload_etc2_batch(ctx._h, ptrs, sizes, 4, W, H, uv.MEM_DEVICE, tex) == 0 and tex[2].status == 0
    back = torch.empty(per, dtype=torch.uint8, device="cuda")
    back.copy_(uv.gather.arena_tensor(ctypes.cast(tex[2].data, ctypes.c_void_p).value, per, "cuda:0"))
    assert bytes(back.cpu().numpy()) == files[2]

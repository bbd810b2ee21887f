"""GPU: parity of the CUDA path (through the C ABI of libuvol_b200.so) against the CPU oracle.
Bit-exact for indices, connectivity-derived point order and integer texels; float attributes are
compared bit-for-bit too (tolerance stated by north_star is 1 ULP; we currently meet 0 ULP because
the kernels use the same two separately rounded fp32 operations as the oracle)."""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, fixture_drc, fixture_ktx2, golden_drc, golden_ktx2, read
from oracle_bind import oracle_draco, oracle_ktx2

sys.path.insert(0, ROOT)
from tools.synth import synth  # noqa: E402

pytestmark = pytest.mark.gpu


def ulp_diff(a, b):
    ia = a.view(np.int32).astype(np.int64); ib = b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7fffffff), ia); ib = np.where(ib < 0, -(ib & 0x7fffffff), ib)
    return np.abs(ia - ib).max() if a.size else 0


def check_geometry(res, blobs):
    for r, b in zip(res, blobs):
        o = oracle_draco(b)
        assert r["status"] == o["status"] == 0
        assert r["num_points"] == o["num_points"] and r["num_faces"] == o["num_faces"]
        assert np.array_equal(r["index"], o["index"])                                   # bit-exact indices
        for k in ("position", "normal", "uv"):
            assert ulp_diff(r["attributes"][k], o[k]) <= 1, k                              # <= 1 ULP (north_star)
            assert np.array_equal(r["attributes"][k].view(np.uint32), o[k].view(np.uint32)), k


def test_geometry_golden(uv, ctx):
    blobs = [read(p) for p in golden_drc()]
    check_geometry(uv.DRACOLoader(ctx).decode_batch(blobs), blobs)


def test_geometry_with_metadata_section(uv, ctx):
    """Files carrying a Draco metadata section (header flag 0x8000) decode like their plain twins."""
    from conftest import with_draco_metadata
    blobs = [with_draco_metadata(read(p)) for p in golden_drc()]
    check_geometry(uv.DRACOLoader(ctx).decode_batch(blobs), blobs)


def test_geometry_all_fixtures(uv, ctx):
    files = fixture_drc()
    if not files:
        pytest.skip("full fixture set not staged (oracle/_ref/fixtures)")
    blobs = [read(p) for p in files]                                   # all 250 frames, one batch
    check_geometry(uv.DRACOLoader(ctx).decode_batch(blobs), blobs)


def test_geometry_replan_paths(uv, monkeypatch):
    """The count-sized arrays are laid out on the device from optimistic reservations.  Shrunk reservations force both re-plan
    paths -- a frame whose attribute tables outgrow their capacity (full bounds, second run) and a batch that outgrows the
    count-sized arenas (exact sizes, second run) -- and the results must not change."""
    blobs = [read(p) for p in golden_drc()] + synth.make_sequence(2, 3000, 32, want_textures=False, seed=20260002)[0]
    for var, val in (("UVOL_CAP_PERMILLE", "300"), ("UVOL_EST_PERMILLE", "400")):
        monkeypatch.setenv(var, val)
        c = uv.Context(0)
        try:
            check_geometry(uv.DRACOLoader(c).decode_batch(blobs), blobs)
            check_geometry(uv.DRACOLoader(c).decode_batch(blobs[:2]), blobs[:2])      # arenas grown by the first call are reused
        finally:
            c.close(); monkeypatch.delenv(var)


def test_texture_golden(uv, ctx):
    blobs = [read(p) for p in golden_ktx2()]
    for r, b in zip(uv.KTX2Loader(ctx).transcode_batch(blobs), blobs):
        o = oracle_ktx2(b)
        assert r["status"] == 0 and (r["width"], r["height"], r["layers"]) == (o["width"], o["height"], o["layers"])
        assert r["dfdTransferFn"] == o["dfd_transfer"] and r["hasAlpha"] == o["has_alpha"]
        assert np.array_equal(r["data"], o["rgba"])                                     # bit-exact texels


def test_texture_all_fixtures(uv, ctx):
    files = fixture_ktx2()
    if not files:
        pytest.skip("full fixture set not staged (oracle/_ref/fixtures)")
    blobs = [read(p) for p in files]                                   # all 50 segments, one batch
    for r, b in zip(uv.KTX2Loader(ctx).transcode_batch(blobs), blobs):
        assert r["status"] == 0 and np.array_equal(r["data"], oracle_ktx2(b)["rgba"])


@pytest.mark.parametrize("nverts", [60, 3000, 50000])
def test_geometry_synthetic(uv, ctx, nverts):
    drc, _, info = synth.make_sequence(3, nverts, 32, want_textures=False, seed=20260002)
    check_geometry(uv.DRACOLoader(ctx).decode_batch(drc), drc)


def _tagged_frames(nverts, masks, qp=11):
    rings, segs = synth.sphere_dims(nverts); fp, fu, uvs, _ = synth.sphere_topology(rings, segs)
    out = []
    for i, mask in enumerate(masks):
        pos = synth.sphere_frame(rings, segs, i / 30.0, 20260051)
        out.append(synth.encode_draco(pos, fp, uvs, fu, synth.vertex_normals(pos, fp), qp=qp, tagged=mask))
    return out


@pytest.mark.parametrize("nverts", [60, 3000, 50000])
def test_geometry_tagged_symbol_scheme(uv, ctx, nverts):
    """Draco's TAGGED symbol scheme (a bit-length tag per value tuple + raw bit fields; what draco_encoder picks when it estimates
    fewer bits, e.g. at high quantisation): every attribute kind tagged on its own, all together, and mixed with RAW frames in
    one batch; 14-bit positions so that tags above 8 bits occur."""
    blobs = _tagged_frames(nverts, [0, 1, 2, 4, 7, 0, 7]) + _tagged_frames(nverts, [7, 1], qp=14)
    check_geometry(uv.DRACOLoader(ctx).decode_batch(blobs), blobs)
    # damaged tagged files fail per item and leave their neighbours alone
    t = blobs[4]
    bad = [t[:len(t) // 2], blobs[0], t[:-7], bytes(t[:len(t) - 40]) + bytes(40)]
    res = uv.DRACOLoader(ctx).decode_batch(bad)
    assert res[0]["status"] < 0 and res[2]["status"] < 0 and res[1]["status"] == 0
    assert res[3]["status"] == oracle_draco(bad[3])["status"] or res[3]["status"] < 0
    check_geometry([res[1]], [bad[1]])


@pytest.mark.parametrize("size,layers", [(8, 1), (64, 3), (1024, 7)])
def test_texture_synthetic(uv, ctx, size, layers):
    blob = synth.encode_etc1s(synth.texture_layers(size, 0, layers, 4))
    r = uv.KTX2Loader(ctx).transcode_batch([blob])[0]
    assert r["status"] == 0 and np.array_equal(r["data"], oracle_ktx2(blob)["rgba"])


def test_ragged_batch_and_failures(uv, ctx):
    """Mixed sizes, an empty batch, and malformed items: a failed item never aborts the batch
    (mirrors src/V2/player.ts:429-444)."""
    dl = uv.DRACOLoader(ctx)
    assert dl.decode_batch([]) == []
    good = read(golden_drc()[0]); small = synth.make_sequence(1, 60, 32, want_textures=False)[0][0]
    bad_magic = b"XRACO" + good[5:]; truncated = good[:4000]; flipped = bytearray(good); flipped[30000] ^= 0x5A
    res = dl.decode_batch([good, bad_magic, small, truncated, bytes(flipped), good])
    assert [r["status"] for r in res[:4]] == [0, -2, 0, res[3]["status"]] and res[3]["status"] < 0 and res[5]["status"] == 0
    check_geometry([res[0], res[2], res[5]], [good, small, good])
    o = oracle_draco(bytes(flipped))
    assert (res[4]["status"] == 0) == (o["status"] == 0)
    kl = uv.KTX2Loader(ctx)
    k = read(golden_ktx2()[0])
    res = kl.transcode_batch([k[:5000], k, b"\x00" * 200])
    assert res[0]["status"] < 0 and res[2]["status"] < 0 and res[1]["status"] == 0
    assert np.array_equal(res[1]["data"], oracle_ktx2(k)["rgba"])


def test_full_size_properties(uv, ctx):
    """BASELINE config sizes (50k verts, 1024^2 x 7): size-independent properties on every frame of a
    larger batch, oracle parity on a sample."""
    drc, tex, info = synth.make_sequence(28, 50000, 1024, sequence_size=7, seed=20260002, distinct_geometry=4)
    res = uv.DRACOLoader(ctx).decode_batch(drc)
    for r in res:
        assert r["status"] == 0 and r["num_faces"] == info["faces"] and r["num_points"] > info["verts"]
        idx = r["index"]
        assert idx.max() == r["num_points"] - 1 and np.unique(idx).size == r["num_points"]         # every point referenced
        assert abs(np.linalg.norm(r["attributes"]["normal"], axis=1) - 1).max() < 1e-5
        assert r["attributes"]["uv"].min() >= 0 and r["attributes"]["uv"].max() <= 1.0001
    for i in (0, 4):                                                                                 # same file -> same output
        for k in ("position", "normal", "uv"):
            assert np.array_equal(res[i]["attributes"][k], res[i % 4]["attributes"][k])
    check_geometry(res[:2], drc[:2])
    tr = uv.KTX2Loader(ctx).transcode_batch(tex)
    assert all(t["status"] == 0 and t["data"].shape == (7, 1024, 1024, 4) and (t["data"][..., 3] == 255).all() for t in tr)
    assert np.array_equal(tr[1]["data"], oracle_ktx2(tex[1])["rgba"])


def test_device_memory_outputs(uv, ctx):
    """UVOL_MEM_DEVICE returns device pointers; copy them back with cudaMemcpy and compare."""
    import ctypes
    blobs = [read(golden_drc()[0])]
    dl = uv.DRACOLoader(ctx)
    raw = dl.decode_batch_raw(blobs, uv.MEM_DEVICE)
    g = raw[0]; assert g.status == 0
    rt = ctypes.CDLL("libcudart.so")
    host = np.empty(g.num_faces * 3, np.uint32)
    assert rt.cudaMemcpy(host.ctypes.data_as(ctypes.c_void_p), ctypes.cast(g.index, ctypes.c_void_p), ctypes.c_size_t(host.nbytes), 2) == 0
    assert np.array_equal(host, oracle_draco(blobs[0])["index"])


# ---- UASTC (KTX2 colour model 166): element-parallel block kernel vs oracle/uastc_oracle.c -------------------------
def _uastc_file(size, layers, seed, mask=None, alpha=True, crop=None):
    img = synth.texture_layers(size, 0, layers, seed)
    if alpha:
        img[..., 3] = ((np.arange(size)[None, :, None] * 5 + np.arange(size)[None, None, :] * 3) & 255).astype(np.uint8)
    if crop:
        img = img[:, :crop[0], :crop[1]]
    return synth.encode_uastc(img, mode_mask=synth.UASTC_ALL_MODES if mask is None else mask, seed=seed, has_alpha=alpha)


def test_uastc_every_mode(uv, ctx):
    blobs = [_uastc_file(64, 2, 100 + m, mask=(1 << m)) for m in range(19)] + [_uastc_file(256, 3, 7)]
    for r, b in zip(uv.KTX2Loader(ctx).transcode_batch(blobs), blobs):
        o = oracle_ktx2(b)
        assert r["status"] == o["status"] == 0 and (r["width"], r["height"], r["layers"]) == (o["width"], o["height"], o["layers"])
        assert r["hasAlpha"] == o["has_alpha"]
        assert np.array_equal(r["data"], o["rgba"])                                     # bit-exact texels


def test_uastc_random_blocks(uv, ctx):
    """Blocks of random bits exercise every field combination an encoder would never write; blocks the
    transcoder rejects (reserved mode, pattern id out of range) are replaced so the file stays decodable."""
    import struct
    rng = np.random.default_rng(20260003)
    blob = bytearray(_uastc_file(256, 1, 1))
    lv = struct.unpack_from("<Q", blob, 80)[0]; n = 64 * 64
    rnd = rng.integers(0, 256, (n, 16), dtype=np.uint8)
    import ctypes
    from oracle_bind import lib as olib
    L = olib(); px = (ctypes.c_uint8 * 64)(); ok = 0
    for i in range(n):
        if L.uvo_uastc_block_to_rgba(rnd[i].ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), px) == 0:
            blob[lv + 16 * i: lv + 16 * i + 16] = rnd[i].tobytes(); ok += 1
    assert ok > n // 2
    r = uv.KTX2Loader(ctx).transcode_batch([bytes(blob)])[0]
    assert r["status"] == 0 and np.array_equal(r["data"], oracle_ktx2(bytes(blob))["rgba"])


def test_uastc_edges_and_failures(uv, ctx):
    import struct
    good = _uastc_file(32, 1, 3)
    ragged = _uastc_file(16, 2, 4, crop=(10, 13))
    lv = struct.unpack_from("<Q", good, 80)[0]
    reserved = bytearray(good); reserved[lv + 16 * 5] = 0x45                                 # mode 19
    badpat = bytearray(good); badpat[lv:lv + 4] = struct.pack("<I", 0x1D | (0x7FFF << 5) | (31 << 20))
    etc1s = synth.encode_etc1s(synth.texture_layers(64, 0, 2, 4))
    blobs = [good, ragged, bytes(reserved), etc1s, bytes(badpat), good[:lv + 64]]
    res = uv.KTX2Loader(ctx).transcode_batch(blobs)
    assert [r["status"] for r in res] == [0, 0, -2, 0, -2, -1]
    assert [oracle_ktx2(b)["status"] for b in blobs] == [0, 0, -2, 0, -2, -1]
    for i in (0, 1, 3):
        assert np.array_equal(res[i]["data"], oracle_ktx2(blobs[i])["rgba"])
    assert res[1]["data"].shape == (2, 10, 13, 4)


def test_uastc_full_size(uv, ctx):
    """C3 texture size (2048^2, 7 layers): oracle parity on one segment, determinism across the batch."""
    a = _uastc_file(2048, 7, 20260003, mask=synth.UASTC_OPAQUE_MODES, alpha=False)
    res = uv.KTX2Loader(ctx).transcode_batch([a, a])
    assert res[0]["status"] == 0 and res[0]["data"].shape == (7, 2048, 2048, 4) and (res[0]["data"][..., 3] == 255).all()
    assert np.array_equal(res[0]["data"], res[1]["data"])
    assert np.array_equal(res[0]["data"], oracle_ktx2(a)["rgba"])


def test_geometry_c3_size(uv, ctx):
    """BASELINE configs[2] geometry size (200k verts, 400k faces): oracle parity on one frame, determinism on the batch."""
    drc, _, info = synth.make_sequence(3, 200000, 32, want_textures=False, seed=20260003, distinct_geometry=2)
    res = uv.DRACOLoader(ctx).decode_batch(drc)
    assert all(r["status"] == 0 and r["num_faces"] == info["faces"] for r in res)
    check_geometry(res[:1], drc[:1])
    assert np.array_equal(res[0]["index"], res[2]["index"]) and np.array_equal(res[0]["attributes"]["uv"], res[2]["attributes"]["uv"])


def test_windows_share_the_traversal_arena(uv):
    """Two contexts (two windows of one sequence) share the traversal-record arena and are driven from two host threads, as
    bench.py does for C3: both windows must still match the oracle, fresh and replayed, device and host outputs."""
    from concurrent.futures import ThreadPoolExecutor
    drc, ktx, info = synth.make_sequence(28, 3000, 64, sequence_size=7, seed=20260011, texture_format="uastc")
    wins = [(drc[:14], ktx[:2]), (drc[14:], ktx[2:])]
    c0, c1 = uv.Context(0, profiling=True), uv.Context(0, profiling=True)
    c1.share_arenas(c0)
    players = [uv.V2Player(c0), uv.V2Player(c1)]
    want = [[oracle_draco(b) for b in wd] for wd, _ in wins]

    def run(w):
        out = []
        for rep in range(3):
            g, t = players[w].decode_step_raw(*wins[w], uv.MEM_HOST) if rep != 1 else players[w].replay_step_raw(len(wins[w][0]), len(wins[w][1]), uv.MEM_HOST)
            ok = all(x.status == 0 for x in g[:len(wins[w][0])]) and all(x.status == 0 for x in t[:len(wins[w][1])])
            for x, o in zip(g, want[w]):
                ok &= x.num_points == o["num_points"] and np.array_equal(np.ctypeslib.as_array(x.index, (x.num_faces * 3,)), o["index"])
                ok &= np.array_equal(np.ctypeslib.as_array(x.position, (x.num_points, 3)).view(np.uint32), o["position"].view(np.uint32))
                ok &= np.array_equal(np.ctypeslib.as_array(x.uv, (x.num_points, 2)).view(np.uint32), o["uv"].view(np.uint32))
            for x, b in zip(t, wins[w][1]):
                ok &= np.array_equal(np.ctypeslib.as_array(x.data, (x.layers, x.height, x.width, 4)), oracle_ktx2(b)["rgba"].reshape(x.layers, x.height, x.width, 4))
            out.append(bool(ok))
        return out
    with ThreadPoolExecutor(2) as ex:
        res = list(ex.map(run, [0, 1]))
    assert res == [[True] * 3, [True] * 3]
    assert uv.span_ms([c0, c1]) > 0
    c1.close(); c0.close()


@pytest.mark.parametrize("fmt", ["uastc", "etc1s"])
def test_sharded_decode_and_gather_equal_single_gpu(uv, fmt):
    """BASELINE configs[3] / SURVEY 8e: one sequence frame-sharded over 2 GPUs + NCCL gather of the whole decoded shards == the same
    sequence decoded by one GPU, byte for byte (tests/tools/gather_check.py under torchrun).  Needs 2 GPUs."""
    import subprocess
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    port = 29000 + os.getpid() % 2000
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tests", "tools", "gather_check.py"), "35", "3000", fmt], capture_output=True, text=True, timeout=600)
    if r.returncode != 0:                                                       # keep the ranks' own words (pytest shortens long assertion messages)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        open(os.path.join(ROOT, "gpurun_out", f"gather_check_{fmt}.log"), "w").write(r.stdout + "\n---- stderr\n" + r.stderr)
    assert r.returncode == 0 and r.stdout.count("GATHER_IDENTICAL") == 2, (r.stdout[-1500:], r.stderr[-3000:])


def test_native_sequence_decodes_ranges_from_disk(uv, ctx, tmp_path):
    """uvol_open + uvol_decode_range (csrc/uvol_sequence.cpp): a V2 clip on disk decoded by frame / segment numbers through the C++ host
    layer equals the oracle's decode of the files the manifest maps those numbers to; a missing file is a per-item IO status;
    uvol_release ends the library's ownership of the result buffers."""
    import ctypes, json
    frames, seq = 21, 7
    drc, ktx, info = synth.make_sequence(frames, 2000, 64, sequence_size=seq, seed=20260041)
    gd = tmp_path / "clip" / "geometry_draco"; td = tmp_path / "clip" / "texture_ktx2_baseColor_default"
    gd.mkdir(parents=True); td.mkdir(parents=True)
    for i, b in enumerate(drc):
        if i != 9:
            (gd / ("%05d.drc" % i)).write_bytes(b)
    for i, b in enumerate(ktx):
        (td / ("%05d.ktx2" % i)).write_bytes(b)
    (tmp_path / "clip.uvol.json").write_text(json.dumps(uv.emit_v2("clip/geometry_[target]/[#####][ext]", 30, frames, "clip/texture_[target]_[type]_[tag]/[#####][ext]", 30, seq, len(ktx))))
    L = uv._native.lib(); h = ctypes.c_void_p()
    assert L.uvol_open(ctx._h, str(tmp_path / "clip.uvol.json").encode(), ctypes.byref(h)) == 0
    g = (uv._native.Geometry * 8)(); t = (uv._native.Texture * 2)()
    assert L.uvol_decode_range(h, 5, 8, 1, 2, uv.MEM_HOST, g, t) == 0
    for k in range(8):
        if 5 + k == 9:
            assert g[k].status == -6
            continue
        o = oracle_draco(drc[5 + k])
        assert g[k].status == 0 and np.array_equal(np.ctypeslib.as_array(g[k].index, (g[k].num_faces * 3,)), o["index"])
        assert np.array_equal(np.ctypeslib.as_array(g[k].uv, (g[k].num_points, 2)).view(np.uint32), o["uv"].view(np.uint32))
    for k in range(2):
        assert t[k].status == 0 and np.array_equal(np.ctypeslib.as_array(t[k].data, (t[k].layers, t[k].height, t[k].width, 4)), oracle_ktx2(ktx[1 + k])["rgba"])
    assert L.uvol_decode_range(h, 20, 2, 0, 0, uv.MEM_HOST, g, t) == -5          # past the clip
    L.uvol_close(h)
    assert L.uvol_release(ctx._h) == 0
    check_geometry(uv.DRACOLoader(ctx).decode_batch(drc[:2]), drc[:2])          # the context keeps working after a release


def test_playback_from_manifest(uv, ctx, tmp_path):
    """a V2 manifest on disk played through V2Playback (leaky-bucket look-ahead -> batched decode -> per-tick frame /
    segment / layer selection); what the renderer would show is the oracle's decode of the right file and layer."""
    import json
    frames, seq = 35, 7
    drc, ktx, info = synth.make_sequence(frames, 2000, 64, sequence_size=seq, seed=20260021)
    gd = tmp_path / "clip" / "geometry_draco"; td = tmp_path / "clip" / "texture_ktx2-1k_baseColor_default"
    gd.mkdir(parents=True); td.mkdir(parents=True)
    for i, b in enumerate(drc):
        (gd / ("%05d.drc" % i)).write_bytes(b)
    for i, b in enumerate(ktx):
        (td / ("%05d.ktx2" % i)).write_bytes(b)
    manifest = {"version": "v2", "geometry": {"targets": {"draco": {"format": "draco", "frameRate": 30, "frameCount": frames}}, "path": "clip/geometry_[target]/[#####][ext]"},
                "texture": {"targets": {"ktx2-1k": {"format": "ktx2", "resolution": [64, 64], "type": "baseColor", "tag": "default", "sequenceSize": seq,
                                                    "sequenceCount": len(ktx), "frameRate": 30}}, "path": "clip/texture_[target]_[type]_[tag]/[#####][ext]"}}
    mp = tmp_path / "clip.uvol.json"; mp.write_text(json.dumps(manifest))
    sq = uv.V2Sequence(str(mp), uv.V2Player(ctx))
    pb = uv.V2Playback(sq.man, sq.decode_copy, buffer_duration=1)
    seen = {}
    for tick in range(0, frames * 4):
        t = tick / 120.0
        if tick % 12 == 0:
            pb.fetch_buffers(t)
        r = pb.update(t)
        if r is not None:
            seen[r["frame"]] = r
            if r["frame"] in (0, 17, 34) and "checked" not in r:
                o = oracle_draco(drc[r["frame"]])
                assert np.array_equal(r["geometry"]["index"], o["index"]) and np.array_equal(r["geometry"]["position"].view(np.uint32), o["position"].view(np.uint32))
                rgba = oracle_ktx2(ktx[r["segment"]])["rgba"].reshape(-1, 64, 64, 4)
                assert r["texture"] is not None and np.array_equal(r["texture"]["data"][r["layer"]], rgba[r["layer"]])
    assert sorted(seen) == list(range(frames)) and pb.requests >= 2 and len(pb.mesh_map) <= 30 + 6


def _zstd_wrap(plain, level=3):
    """Re-wraps a scheme-0 UASTC KTX2 as supercompressionScheme 2: the level becomes one Zstandard frame (made by libzstd, test-only)."""
    import struct
    from test_zstd import Z, compress
    if Z is None:
        pytest.skip("libzstd not present (needed to build the test vector)")
    lv_off, lv_len = struct.unpack_from("<QQ", plain, 80)
    z = compress(plain[lv_off:lv_off + lv_len], level)
    out = bytearray(plain[:lv_off]) + z
    struct.pack_into("<I", out, 44, 2); struct.pack_into("<QQQ", out, 80, lv_off, len(z), lv_len)
    return bytes(out)


def test_uastc_zstd_supercompression(uv, ctx):
    """KTX2 supercompressionScheme 2 (what `basisu -uastc -ktx2` writes by default): the level is inflated by the library's own
    Zstandard decoder on the host, then goes through the same block kernel -- texels must equal the oracle's decode of the plain file."""
    plain = [_uastc_file(256, 3, 31), _uastc_file(64, 2, 32, crop=(50, 61)), _uastc_file(512, 7, 33, mask=synth.UASTC_OPAQUE_MODES, alpha=False)]
    zs = [_zstd_wrap(p, lvl) for p, lvl in zip(plain, (1, 9, 3))]
    broken = zs[0][:len(zs[0]) - 100]
    res = uv.KTX2Loader(ctx).transcode_batch([zs[0], plain[1], zs[1], zs[2], broken])
    assert [r["status"] for r in res[:4]] == [0, 0, 0, 0] and res[4]["status"] < 0
    for r, p in zip(res[:4], [plain[0], plain[1], plain[1], plain[2]]):
        assert np.array_equal(r["data"], oracle_ktx2(p)["rgba"])


def test_texture_target_etc1(uv, ctx):
    """Target format ETC1 (the reference's etc1Supported / opaque etc2Supported choice, KTX2Loader.js:619-636): the GPU's blocks equal
    the host run of the same repack function byte for byte, decode (independent ETC1 decoder) to the oracle's RGBA32 texels, and
    sources the target cannot take (UASTC) are reported per item without failing the batch."""
    from emu_bind import emu_ktx2_etc1
    from etc1_decode import decode_etc1
    real = read(golden_ktx2()[0]); small = synth.encode_etc1s(synth.texture_layers(64, 0, 3, 4)); ua = _uastc_file(32, 1, 3)
    res = uv.KTX2Loader(ctx).transcode_batch([real, ua, small], target=uv.TEX_ETC1)
    assert [r["status"] for r in res] == [0, -3, 0]
    for r, f in ((res[0], real), (res[2], small)):
        e = emu_ktx2_etc1(f)
        assert r["format"] == "RGB_ETC1_Format" and np.array_equal(r["data"], e["blocks"])
        rgba = oracle_ktx2(f)["rgba"].reshape(r["layers"], r["height"], r["width"], 4)
        assert np.array_equal(decode_etc1(r["data"][1], r["width"], r["height"]), rgba[1])
    again = uv.KTX2Loader(ctx).transcode_batch([small])[0]                                        # the default target is unaffected
    assert again["status"] == 0 and np.array_equal(again["data"], oracle_ktx2(small)["rgba"])

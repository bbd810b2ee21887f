"""CPU: manifest schema, URL templates, frame mapping and fetch windows behave like the reference's TypeScript
(src/utils.ts, src/V2/player.ts); frame sharding covers every frame exactly once (also across 2 gloo ranks)."""
import importlib
import json
import math
import os
import sys

import pytest

from conftest import ROOT

man = importlib.import_module("universal-volumetric_b200.manifest")

LIAM = {"version": "v2", "audio": {"path": "liam/output/liam[ext]", "format": "mp3"},
        "geometry": {"targets": {"draco": {"format": "draco", "frameRate": 30, "frameCount": 250}}, "path": "liam/output/geometry_[target]/[#####][ext]"},
        "texture": {"targets": {"ktx2-fps30-1k": {"format": "ktx2", "resolution": [1024, 1024], "type": "baseColor", "tag": "default",
                                                   "sequenceSize": 5, "sequenceCount": 50, "frameRate": 30}},
                    "path": "liam/output/texture_[target]_[type]_[tag]/[#####][ext]"}}


def test_url_helpers():
    assert man.pad(7, 5) == "00007" and man.pad(123456, 5) == "123456"                   # src/utils.ts:10-14
    assert man.count_hash_char("a/[####]/[##]") == 6
    assert man.get_absolute_url("https://h/x/liam.uvol.json", "liam/output/a.drc") == "https://h/x/liam/output/a.drc"
    assert man.get_absolute_url("/data/liam.uvol.json", "http://cdn/a.drc") == "http://cdn/a.drc"


def test_v2_paths_and_mapping():
    m = man.V2Manifest(LIAM, "/data/public/liam.uvol.json")
    assert m.geometry_target == "draco" and m.texture_target == "ktx2-fps30-1k"
    assert m.geometry_url(0) == "/data/public/liam/output/geometry_draco/00000.drc"
    assert m.geometry_url(137) == "/data/public/liam/output/geometry_draco/00137.drc"
    assert m.texture_url(49) == "/data/public/liam/output/texture_ktx2-fps30-1k_baseColor_default/00049.ktx2"
    assert (m.geometry_frame_count, m.batch_size, m.texture_segment_count) == (250, 5, 50)
    f = m.frames_at(1.25)                                                               # round(30 * 1.25) = 38 (Math.round: 37.5 -> 38)
    assert f == {"geometry_frame": 38, "texture_frame": 38, "segment": 7, "layer": 3}
    assert m.frames_at(0.0)["layer"] == 0
    with pytest.raises(ValueError):
        man.V2Manifest({"version": "v1"}, "x")


def test_fetch_window_leaky_bucket():
    m = man.V2Manifest(LIAM, "liam.uvol.json")
    geo, tex, lg, ls = m.fetch_window(0.0, -1, -1, buffer_duration=4)                    # first fetchBuffers of a track
    assert geo == list(range(0, 121)) and lg == 120                                      # frames (−1, 0 + 4*30]
    assert tex == list(range(0, 25)) and ls == 24                                        # ceil(30/5) = 6 segments per second
    geo, tex, lg, ls = m.fetch_window(2.0, lg, ls)                                       # two seconds later: only the new tail
    assert geo == list(range(121, 181)) and tex == list(range(25, 37))
    geo, tex, lg, ls = m.fetch_window(9.0, 249, 49)
    assert geo == [] and tex == []                                                       # everything already requested


def test_v1_manifest(tmp_path):
    fd = [{"frameNumber": i, "keyframeNumber": i, "startBytePosition": 100 * i, "vertices": 10, "faces": 12, "meshLength": 100 - i} for i in range(5)]
    p = tmp_path / "clip.manifest"; p.write_text(json.dumps({"maxVertices": 10, "maxTriangles": 12, "frameRate": 30, "frameData": fd}))
    m = man.V1Manifest.load(str(p))
    assert m.mesh_file.endswith("clip.drcs")
    assert m.byte_range(1, 4) == (100, 300 + 97)
    blob = bytes(range(256)) * 2
    sl = m.slices(blob[100:397], 100, 1, 4)
    assert [s[0] for s in sl] == [1, 2, 3] and [len(s[2]) for s in sl] == [99, 98, 97] and sl[1][2] == blob[200:298]


@pytest.mark.parametrize("frames,seq,world", [(300, 7, 1), (300, 7, 2), (300, 7, 8), (1000, 7, 8), (250, 5, 4), (6, 7, 8)])
def test_sharding_covers_everything_once(frames, seq, world):
    nseg = (frames + seq - 1) // seq
    got_f, got_s = [], []
    for r in range(world):
        f0, f1, s0, s1 = man.shard_v2(frames, seq, nseg, world, r)
        assert f0 % seq == 0 or f0 == frames
        got_f += list(range(f0, f1)); got_s += list(range(s0, s1))
    assert got_f == list(range(frames)) and got_s == list(range(nseg))


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    import torch
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    f0, f1, s0, s1 = man.shard_v2(300, 7, 43, world, rank)
    counts = torch.tensor([f1 - f0, s1 - s0, f0, s0], dtype=torch.int64)
    out = [torch.zeros(4, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(out, counts)                                     # the optional final gather exchanges per-rank sizes first
    t = torch.tensor([float(rank + 1)]); dist.all_reduce(t, op=dist.ReduceOp.MAX)   # bench: max-over-ranks timing
    if rank == 0:
        q.put(([o.tolist() for o in out], float(t)))
    dist.destroy_process_group()


def test_two_rank_sharding_gloo():
    """world_size-2 run on CPU (gloo): the two ranks' shards tile the clip, sizes are exchanged, timing is max-reduced."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn"); q = ctx.Queue(); port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res, tmax = q.get(timeout=120)
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert tmax == 2.0
    assert res[0][0] + res[1][0] == 300 and res[0][1] + res[1][1] == 43
    assert res[0][2] == 0 and res[1][2] == res[0][0] and res[1][3] == res[0][1]


class _G:        # stands in for uvol_geometry: pointers into one host arena
    pass


def _fake_results(rank, nframes):
    """A rank's 'decoded' frames laid out like the library's output arena: all index buffers first, then (after the index buffers of
    frames that are NOT gathered, here a 256 KB hole) the per-point arrays, every array 128-byte aligned."""
    import ctypes
    import numpy as np
    rng = np.random.default_rng(100 + rank)
    sizes = [(50 + 7 * i + rank, 90 + 11 * i) for i in range(nframes)]                  # (points, faces)
    al = lambda n: (n + 127) // 128 * 128
    total = sum(al(F * 12) + 2 * al(P * 12) + al(P * 8) for P, F in sizes) + (256 << 10)
    arena = np.zeros(total, np.uint8); base = arena.ctypes.data; res = []; truth = []
    cur = 0; slots = []
    for P, F in sizes:                                                                  # index region
        slots.append({"index": cur}); cur += al(F * 12)
    cur += 256 << 10                                                                    # index buffers of other frames
    for (P, F), sl in zip(sizes, slots):                                                # attribute region
        for name, nbytes in (("position", P * 12), ("normal", P * 12), ("uv", P * 8)):
            sl[name] = cur; cur += al(nbytes)
    for (P, F), sl in zip(sizes, slots):
        g = _G(); g.status = 0; g.num_points = P; g.num_faces = F; t = {}
        for name, n, dt in (("index", F * 3, np.int32), ("position", P * 3, np.float32), ("normal", P * 3, np.float32), ("uv", P * 2, np.float32)):
            vals = (rng.integers(0, P, n) if dt == np.int32 else rng.random(n)).astype(dt)
            arena[sl[name]:sl[name] + 4 * n] = vals.view(np.uint8); setattr(g, name, ctypes.c_void_p(base + sl[name])); t[name] = vals
        res.append(g); truth.append(t)
    bad = _G(); bad.status = -2; bad.num_points = bad.num_faces = 0
    bad.index = bad.position = bad.normal = bad.uv = None
    res.append(bad); truth.append(None)
    return res, truth, arena


def _gather_worker(rank, world, port, q):
    import numpy as np
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    gather = importlib.import_module("universal-volumetric_b200.gather")
    res, truth, arena = _fake_results(rank, 3 + rank)                                  # ragged: ranks hold different frame counts
    tables, arenas = gather.all_gather_geometry(res, len(res), "cpu")
    ok = True
    for r in range(world):
        _, tr, _ = _fake_results(r, 3 + r)
        for i, t in enumerate(tr):
            v = gather.frame_views(tables, arenas, r, i)
            if t is None:
                ok &= v is None
            else:
                ok &= all(np.array_equal(v[k].numpy().ravel(), t[k]) for k in t)
    ok &= int(arenas.shape[1]) < (128 << 10)                                          # the 256 KB hole between the two runs does not travel
    q.put((rank, bool(ok), tuple(tables.shape)))
    dist.destroy_process_group()


def test_two_rank_final_gather_gloo():
    """The optional final gather: after it every rank holds every rank's frames byte for byte, including a failed frame."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn"); q = ctx.Queue(); port = 31000 + os.getpid() % 2000
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    got = [q.get(timeout=180) for _ in range(2)]
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert all(ok for _, ok, _ in got) and all(shape == (2, 5, 8) for _, _, shape in got)


class _T:        # stands in for uvol_texture
    pass


def _fake_shard(rank):
    """(geometry results, texture results, truth) of a rank: region-major geometry arena like the library's, textures back to back."""
    import ctypes
    import numpy as np
    res, truth, arena = _fake_results(rank, 2 + rank)
    rng = np.random.default_rng(500 + rank)
    tex, ttruth = [], []
    sizes = [(8 + 4 * i, 12, 2 + rank) for i in range(3 - rank)] + [None]                # one failed segment at the end
    tarena = np.zeros(sum((w * h * l * 4 + 127) // 128 * 128 for w, h, l in [x for x in sizes if x]) + 128, np.uint8); cur = 0
    for sz in sizes:
        t = _T()
        if sz is None:
            t.status = -2; t.width = t.height = t.layers = t.format = t.has_alpha = 0; t.bytes = 0; t.data = None; ttruth.append(None)
        else:
            w, h, l = sz; n = w * h * l * 4
            vals = rng.integers(0, 256, n, dtype=np.uint8); tarena[cur:cur + n] = vals
            t.status = 0; t.width, t.height, t.layers, t.format, t.has_alpha, t.bytes = w, h, l, 0, 0, n
            t.data = ctypes.c_void_p(tarena.ctypes.data + cur); cur += (n + 127) // 128 * 128; ttruth.append(vals)
        tex.append(t)
    return res, tex, truth, ttruth, (arena, tarena)


def _shard_gather_worker(rank, world, port, q):
    import numpy as np
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    gather = importlib.import_module("universal-volumetric_b200.gather")
    res, tex, _, _, keep = _fake_shard(rank)
    G = gather.all_gather_shard(res, len(res), tex, len(tex), "cpu")
    ok = True
    for r in range(world):
        _, _, tr, ttr, _ = _fake_shard(r)
        for i, t in enumerate(tr):
            v = gather.shard_frame_views(G, r, i)
            ok &= (v is None) if t is None else all(np.array_equal(v[k].numpy().ravel(), t[k]) for k in t)
        for i, t in enumerate(ttr):
            v = gather.shard_texture_view(G, r, i)
            ok &= (v is None) if t is None else np.array_equal(v.numpy(), t)
    G2 = gather.all_gather_shard(res, len(res), tex, len(tex), "cpu", arena=G["arena"])          # the receive arena is reused
    ok &= G2["arena"].data_ptr() == G["arena"].data_ptr()
    q.put((rank, bool(ok), [int(b) for b in G["bytes"]]))
    dist.destroy_process_group()


def test_two_rank_whole_shard_gather_gloo():
    """BASELINE configs[3]'s gather: geometry AND textures of every rank's shard end up on every rank byte for byte (ragged shards,
    failed items included); every rank reports the same per-rank byte counts."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn"); q = ctx.Queue(); port = 33000 + os.getpid() % 2000
    procs = [ctx.Process(target=_shard_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    got = [q.get(timeout=180) for _ in range(2)]
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert all(ok for _, ok, _ in got) and got[0][2] == got[1][2] and all(b > 0 for b in got[0][2])


def test_playback_buffer_follows_the_reference_scheduler():
    """V2Playback vs src/V2/player.ts:272-323,388-470,531-562 with a stub decoder: the look-ahead stays `bufferDuration` seconds
    ahead, every frame / segment is requested exactly once, a failed mesh is skipped, a missing segment shows the mesh untextured,
    played buffers are dropped ceil(120 / fps) frames behind the clock."""
    m = man.V2Manifest(LIAM, "/data/public/liam.uvol.json")
    asked_g, asked_s = [], []

    def decode(frames, segments):
        asked_g.extend(frames); asked_s.extend(segments)
        return {f: ("mesh", f) for f in frames if f != 37}, {s: ("tex", s) for s in segments if s != 9}      # frame 37 and segment 9 fail to decode

    pb = man.V2Playback(m, decode, buffer_duration=4)
    assert pb.fetch_buffers(0.0) == (121, 25) and pb.buffered_fraction() == pytest.approx(120 / 120)      # frames 0..120, segments 0..24; frame 37 failed
    assert pb.fetch_buffers(0.0) == (0, 0)                                   # nothing new until the clock moves
    shown = []
    for tick in range(0, 250 * 4 + 40):                                      # 120 Hz render loop over the 250-frame clip (+ overrun)
        t = tick / 120.0
        if tick % 12 == 0:
            pb.fetch_buffers(t)                                             # the reference polls every intervalDuration
        r = pb.update(t)
        if r is not None:
            shown.append((r["frame"], r["segment"], r["layer"], r["texture"] is not None))
            assert r["geometry"] == ("mesh", r["frame"]) and r["segment"] == r["frame"] // 5 and r["layer"] == r["frame"] % 5
            assert min(pb.mesh_map) >= r["frame"] - 4 - 1 and max(pb.mesh_map) <= min(249, r["frame"] + 4 * 30 + 3)
    frames_shown = sorted({f for f, _, _, _ in shown})
    assert frames_shown == [f for f in range(250) if f != 37]                 # the failed frame is simply absent (:435-437)
    assert all(tex == (seg != 9) for _, seg, _, tex in shown)                # segment 9 missing -> failMaterial (:439-444)
    assert sorted(asked_g) == list(range(250)) and sorted(asked_s) == list(range(50))      # each requested exactly once
    assert pb.process_frame(250 / 30.0 + 0.1) is None                        # past the last frame: track end


def test_manifest_tooling_both_dialects(tmp_path):
    """the player's schema and the encoder script's dialect describe the same clip; the frame-count check follows
    scripts/Encoder.py:103-154; the V1 writer follows Encoder30.js:155-160 and round-trips through V1Manifest."""
    import struct
    enc = man.emit_v2("DRACO/frame_[#####].drc", 30, 250, "KTX2/texture_[#######].ktx2", 30, 7, 36, dialect="encoder")
    assert isinstance(enc["texture"]["targets"], list) and "targets" not in enc["geometry"]          # exactly what Encoder.py writes
    ply = man.emit_v2("DRACO/frame_[#####].drc", 30, 250, "KTX2/texture_[#######].ktx2", 30, 7, 36, dialect="player", resolution=(1024, 1024))
    a, b = man.V2Manifest(enc, "/clips/a/manifest.json"), man.V2Manifest(ply, "/clips/a/manifest.json")
    for m in (a, b):
        assert (m.geometry_frame_count, m.batch_size, m.texture_segment_count) == (250, 7, 36)
        assert m.geometry_url(12) == "/clips/a/DRACO/frame_00012.drc" and m.texture_url(3) == "/clips/a/KTX2/texture_0000003.ktx2"
        assert m.frames_at(1.0) == {"geometry_frame": 30, "texture_frame": 30, "segment": 4, "layer": 2}
    ktx = lambda layers: b"\xabKTX 20\xbb\r\n\x1a\n" + struct.pack("<IIIIIIIII", 0, 1, 64, 64, 0, layers, 1, 1, 0)
    segs = [ktx(7)] * 35 + [ktx(5)]                                                                   # 35 full segments + a short last one = 250 frames
    ok = man.check_total_frames(250, 30, segs, 7, 30)
    assert ok["compatible"] and ok["texture_frames"] == 250 and ok["durations"]["geometry"] == pytest.approx(250 / 30)
    assert not man.check_total_frames(250, 30, [ktx(7)] * 36, 7, 30)["compatible"]                   # 252 texture frames for 250 meshes
    assert man.check_total_frames(250, 30, [ktx(7)] * 17 + [ktx(6)], 7, 15)["compatible"]            # half-rate texture: 125 frames at 15 fps
    with pytest.raises(ValueError):
        man.ktx2_layer_count(b"not a ktx2 file at all, really not............")
    v1 = man.emit_v1(30, [(100, 196, 1500), (101, 198, 1520), (99, 194, 1480)])
    p = tmp_path / "clip.manifest"; p.write_text(json.dumps(v1))
    m1 = man.V1Manifest.load(str(p))
    assert v1["maxVertices"] == 101 and v1["maxTriangles"] == 198 and v1["frameData"][2]["startBytePosition"] == 3020
    assert m1.byte_range(1, 2) == (1500, 3020) and m1.byte_range(1, 3) == (1500, 4500)      # [start, end) frames


def test_v1_sequence_slices_and_keys(tmp_path):
    """V1Sequence (src/V1/worker.ts:24-74) with a stub decoder: one range read, per-frame slices in file order, results keyed by
    keyframeNumber, failed frames absent, the range clamped to the clip."""
    frames = [(10, 16, 100), (11, 18, 140), (12, 20, 90), (9, 14, 60)]
    v1 = man.emit_v1(30, frames)
    (tmp_path / "clip.manifest").write_text(json.dumps(v1))
    payload = b"".join(bytes([65 + i]) * n for i, (_, _, n) in enumerate(frames))
    (tmp_path / "clip.drcs").write_bytes(payload)
    seen = []

    class Stub:
        def decode_batch(self, blobs):
            seen.append([bytes(b) for b in blobs])
            return [{"status": 0 if b[:1] != b"C" else -2, "index": len(b), "position": b[:1], "uv": None} for b in blobs]

    sq = man.V1Sequence(str(tmp_path / "clip.manifest"), Stub())
    out = sq.decode(1, 9)
    assert seen == [[b"B" * 140, b"C" * 90, b"D" * 60]] and sorted(out) == [1, 3]                    # frame 2 failed to decode -> absent
    assert out[3]["frameNumber"] == 3 and out[3]["bufferGeometry"]["index"] == 60 and out[1]["bufferGeometry"]["position"] == b"B"
    assert sq.decode(4, 9) == {} and sq.man.mesh_file.endswith("clip.drcs")


def test_native_sequence_open_matches_python_manifest(tmp_path):
    """The C++ host layer (csrc/uvol_sequence.cpp: uvol_open / uvol_sequence_get_info / uvol_sequence_url / uvol_sequence_frames_at) reads
    both V2 dialects and V1 manifests exactly like manifest.py (= src/Interfaces.ts, src/V2/player.ts:141-174,418-446, src/utils.ts:10-45)."""
    import ctypes
    uvp = importlib.import_module("universal-volumetric_b200")
    L = uvp._native.lib()

    def open_seq(path):
        h = ctypes.c_void_p()
        assert L.uvol_open(None, str(path).encode(), ctypes.byref(h)) == 0 and h
        return h

    def url(h, kind, n):
        buf = ctypes.create_string_buffer(1024)
        assert L.uvol_sequence_url(h, kind, n, buf, 1024) > 0
        return buf.value.decode()
    for dialect in ("player", "encoder"):
        m = man.emit_v2("geometry_[target]/[#####][ext]" if dialect == "player" else "DRACO/frame_[#####].drc", 30, 250,
                        "texture_[target]_[type]_[tag]/[#####][ext]" if dialect == "player" else "KTX2/texture_[#######].ktx2", 30, 5, 50, dialect=dialect)
        p = tmp_path / f"{dialect}.uvol.json"; p.write_text(json.dumps(m))
        py = man.V2Manifest.load(str(p)); h = open_seq(p)
        info = uvp._native.SequenceInfo(); assert L.uvol_sequence_get_info(h, ctypes.byref(info)) == 0
        assert (info.version, info.geometry_frame_count, info.sequence_size, info.sequence_count, info.geometry_frame_rate) == (2, 250, 5, 50, 30.0)
        for n in (0, 7, 249):
            assert url(h, 0, n) == py.geometry_url(n)
        for n in (0, 49):
            assert url(h, 1, n) == py.texture_url(n)
        for t in (0.0, 0.016, 0.05, 1.0, 3.99, 8.3):
            g = ctypes.c_uint32(); s = ctypes.c_uint32(); l = ctypes.c_uint32()
            assert L.uvol_sequence_frames_at(h, t, ctypes.byref(g), ctypes.byref(s), ctypes.byref(l)) == 0
            w = py.frames_at(t)
            assert (g.value, s.value, l.value) == (w["geometry_frame"], w["segment"], w["layer"])
        # playback planning: the native fetch window / eviction thresholds follow manifest.py (= fetchBuffers / removePlayedBuffer) along a whole clip
        class Plan(ctypes.Structure):
            _fields_ = [("first_frame", ctypes.c_int32), ("n_frames", ctypes.c_int32), ("first_segment", ctypes.c_int32), ("n_segments", ctypes.c_int32)]
        L.uvol_sequence_fetch_window.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32), ctypes.c_double, ctypes.POINTER(Plan)]
        L.uvol_sequence_keep_from.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32)]
        lg = ctypes.c_int32(-1); ls = ctypes.c_int32(-1); plg, pls = -1, -1
        pb = man.V2Playback(py, lambda g, s_: ({}, {}), buffer_duration=4)
        for step in range(0, 120):
            t = step * 0.1
            geo, tex, plg, pls = py.fetch_window(t, plg, pls, 4)
            plan = Plan(); assert L.uvol_sequence_fetch_window(h, t, ctypes.byref(lg), ctypes.byref(ls), 4.0, ctypes.byref(plan)) == 0
            assert (lg.value, ls.value) == (plg, pls)
            assert list(range(plan.first_frame, plan.first_frame + plan.n_frames)) == geo and list(range(plan.first_segment, plan.first_segment + plan.n_segments)) == tex
            kf = ctypes.c_int32(); ks = ctypes.c_int32(); assert L.uvol_sequence_keep_from(h, t, ctypes.byref(kf), ctypes.byref(ks)) == 0
            at = py.frames_at(t)
            assert kf.value == at["geometry_frame"] - math.ceil(120 / 30) and ks.value == at["segment"] - math.ceil(120 / (30 * 5))
        assert lg.value == 249 and ls.value == 49                  # the whole clip was requested exactly once
        L.uvol_close(h)
    v1 = man.emit_v1(30, [(100, 196, 1500), (101, 198, 1520), (99, 194, 1480)])
    p = tmp_path / "clip.manifest"; p.write_text(json.dumps(v1))
    h = open_seq(p); info = uvp._native.SequenceInfo(); L.uvol_sequence_get_info(h, ctypes.byref(info))
    assert (info.version, info.geometry_frame_count, info.max_vertices, info.max_triangles) == (1, 3, 101, 198)
    L.uvol_close(h)
    bad = tmp_path / "bad.json"; bad.write_text('{"version": "v2", "geometry": [1, 2')
    h = ctypes.c_void_p()
    assert L.uvol_open(None, str(bad).encode(), ctypes.byref(h)) < 0 and not h
    assert L.uvol_open(None, str(tmp_path / "missing.json").encode(), ctypes.byref(h)) == -6


def test_pick_texture_format_follows_format_options():
    """uvol_pick_texture_format = getTranscoderFormat (src/lib/KTX2Loader.js:591-689) in the order the reference effectively applies, restricted to
    what the library produces for the source."""
    uvp = importlib.import_module("universal-volumetric_b200")
    L = uvp._native.lib(); N = uvp._native
    ASTC, BPTC, DXT, ETC2, ETC1, PVRTC = 1, 2, 4, 8, 16, 32
    pick = lambda uastc, alpha, caps: L.uvol_pick_texture_format(int(uastc), int(alpha), caps)
    assert pick(True, False, ASTC | BPTC | DXT) == N.TEX_ASTC_4x4 and pick(False, False, ASTC | BPTC | DXT) == N.TEX_BC7          # ASTC is UASTC-only
    assert pick(True, True, BPTC | DXT) == N.TEX_BC7 and pick(False, True, BPTC | ETC2) == N.TEX_BC7                              # desktop NVIDIA: BC7 for both
    assert pick(False, False, ETC2 | ETC1 | DXT) == N.TEX_ETC1 and pick(False, True, ETC2 | ETC1 | DXT) == N.TEX_ETC2_RGBA         # mobile: the ETC2 pair
    assert pick(False, True, ETC1 | DXT) == N.TEX_BC3 and pick(False, False, ETC1 | DXT) == N.TEX_ETC1                            # ETC1 has no alpha form
    assert pick(False, False, DXT) == N.TEX_BC1 and pick(False, True, DXT) == N.TEX_BC3
    assert pick(True, False, ETC2 | ETC1 | DXT | PVRTC) == N.TEX_RGBA32 and pick(False, False, PVRTC) == N.TEX_RGBA32 and pick(False, True, 0) == N.TEX_RGBA32

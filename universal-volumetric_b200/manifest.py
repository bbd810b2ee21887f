"""Manifest layer: the reference's V1 / V2 manifest schema, URL templates and frame -> file / layer mapping, kept
verbatim as the drop-in surface, plus the frame sharding used for multi-GPU decode.

Mirrors (paths relative to the reference repo):
* `V2Schema`, `GeometryTarget`, `KTX2TextureTarget`, `FORMATS_TO_EXT`           src/Interfaces.ts:21-37,60-73,75-132,156-161
* `pad`, `countHashChar`, `getAbsoluteURL`                                      src/utils.ts:10-45
* `getGeometryURL`, `getTextureURL`, target choice in `playTrack`              src/V2/player.ts:141-174,199-221
* `getCurrentFrame`, segment / layer selection in `processFrame`                src/V2/player.ts:43-45,418-420,446
* `fetchBuffers` request windows (leaky bucket)                                 src/V2/player.ts:272-323
* `V1Schema`, `V1FrameData`; .manifest -> .drcs; per-frame slices               src/Interfaces.ts:1-15; src/V1/player.ts:337; src/V1/worker.ts:48-56
Everything here is host logic (no decode); the decode calls go through libuvol_b200.so.
"""
import json
import math
import os

FORMATS_TO_EXT = {"mp3": ".mp3", "draco": ".drc", "ktx2": ".ktx2", "etc2": ".etc2"}          # src/Interfaces.ts:156-161
TEXTURE_FORMAT_PRIORITY = {"ktx2": 0, "etc2": 1, "etc1": 2}                                   # src/Interfaces.ts:165-169


def pad(n, width):
    """src/utils.ts:10-14"""
    s = str(int(n))
    return s if len(s) >= width else "0" * (width - len(s)) + s


def count_hash_char(url):
    """src/utils.ts:16-24"""
    return url.count("#")


def get_absolute_url(manifest_url, new_segment):
    """src/utils.ts:38-45: absolute (http...) paths pass through, relative ones replace the manifest's file name."""
    if new_segment.startswith("http"):
        return new_segment
    parts = manifest_url.split("/")
    parts.pop()
    parts.append(new_segment)
    return "/".join(parts)


def js_round(x):
    """Math.round: half away from floor (src/V2/player.ts:43-45 uses Math.round(frameRate * currentTime))."""
    return math.floor(x + 0.5)


class V2Manifest:
    """A V2 manifest (`version == 'v2'`) with the player's target choice and path expansion."""

    texture_type = "baseColor"        # src/V2/player.ts:82
    texture_tag = "default"           # src/V2/player.ts:83

    def __init__(self, manifest, manifest_path):
        if manifest.get("version") != "v2":
            raise ValueError("not a V2 manifest (src/Player.ts:127-132 dispatches on version == 'v2')")
        manifest = normalize_v2(manifest)              # accepts the encoder script's dialect too (scripts/Encoder.py:311-328)
        self.m, self.path = manifest, manifest_path
        # playTrack: first geometry target; texture targets sorted by TEXTURE_FORMAT_PRIORITY, first supported one
        # (isTextureFormatSupported: 'ktx2' always, src/utils.ts:26-32); note the reference compares the target KEY
        # against format names there, so with ordinary keys the first key wins -- reproduced here.
        self.geometry_target = next(iter(manifest["geometry"]["targets"]))
        self.texture_target = next(iter(manifest["texture"]["targets"]))
        keys = list(manifest["texture"]["targets"])
        keys.sort(key=lambda k: -TEXTURE_FORMAT_PRIORITY.get(manifest["texture"]["targets"][k]["format"], -1))
        for k in keys:
            if k in ("ktx2", "mp4"):
                self.texture_target = k
                break

    @classmethod
    def load(cls, path):
        with open(path) as fh:
            return cls(json.load(fh), path)

    # ---- counts (src/V2/player.ts:183-196)
    @property
    def geometry(self):
        return self.m["geometry"]["targets"][self.geometry_target]

    @property
    def texture(self):
        return self.m["texture"]["targets"][self.texture_target]

    @property
    def geometry_frame_count(self):
        return self.geometry["frameCount"]

    @property
    def batch_size(self):
        return self.texture["sequenceSize"]

    @property
    def texture_segment_count(self):
        return self.texture["sequenceCount"]

    # ---- URL templates (src/V2/player.ts:141-174)
    def geometry_url(self, frame_no):
        tpl = self.m["geometry"]["path"]
        w = count_hash_char(tpl)
        subs = {"[target]": self.geometry_target, "[ext]": FORMATS_TO_EXT[self.geometry["format"]], "[" + "#" * w + "]": pad(frame_no, w)}
        for k, v in subs.items():
            tpl = tpl.replace(k, v, 1)
        return get_absolute_url(self.path, tpl)

    def texture_url(self, segment_no):
        tpl = self.m["texture"]["path"]
        w = count_hash_char(tpl)
        subs = {"[target]": self.texture_target, "[type]": self.texture_type, "[tag]": self.texture_tag,
                "[ext]": FORMATS_TO_EXT[self.texture["format"]], "[" + "#" * w + "]": pad(segment_no, w)}
        for k, v in subs.items():
            tpl = tpl.replace(k, v, 1)
        return get_absolute_url(self.path, tpl)

    # ---- time -> frame / segment / layer (src/V2/player.ts:418-420,446)
    def frames_at(self, t):
        gf = js_round(self.geometry["frameRate"] * t)
        tf = js_round(self.texture["frameRate"] * t)
        return {"geometry_frame": gf, "texture_frame": tf, "segment": tf // self.batch_size, "layer": tf % self.batch_size}

    # ---- leaky-bucket request window (src/V2/player.ts:272-323): what fetchBuffers asks for at time t
    def fetch_window(self, t, last_geometry, last_segment, buffer_duration=4):
        gsize = self.geometry["frameRate"]
        cur_g = js_round(gsize * t)
        tsize = math.ceil(self.texture["frameRate"] / self.batch_size)
        cur_s = js_round(self.texture["frameRate"] * t) // self.batch_size
        geo, tex = [], []
        for i in range(buffer_duration):
            g_end = min(cur_g + (i + 1) * gsize, self.geometry_frame_count - 1)
            if last_geometry != self.geometry_frame_count - 1 and last_geometry < g_end:
                geo += list(range(last_geometry + 1, g_end + 1)); last_geometry = g_end
            s_end = min(cur_s + (i + 1) * tsize, self.texture_segment_count - 1)
            if last_segment != self.texture_segment_count - 1 and last_segment < s_end:
                tex += list(range(last_segment + 1, s_end + 1)); last_segment = s_end
        return geo, tex, last_geometry, last_segment


def shard_v2(frame_count, sequence_size, segment_count, world, rank):
    """Frame sharding for multi-GPU decode: contiguous blocks of whole KTX2 segments (ETC1S P-frames chain
    inside one file, so a segment is never split); rank r owns segments [r*S/G, (r+1)*S/G) and the geometry frames they
    cover.  Returns (first_frame, end_frame, first_segment, end_segment)."""
    s0 = rank * segment_count // world
    s1 = (rank + 1) * segment_count // world
    f0 = min(s0 * sequence_size, frame_count)
    f1 = frame_count if rank == world - 1 else min(s1 * sequence_size, frame_count)
    return f0, f1, s0, s1


class V2Sequence:
    """A V2 clip on local storage: reads the files the manifest names and decodes ranges through the library."""

    def __init__(self, manifest_path, player):
        self.man = V2Manifest.load(manifest_path)
        self.player = player                      # universal-volumetric_b200.V2Player

    def read_geometry(self, frames):
        return [open(self.man.geometry_url(f), "rb").read() for f in frames]

    def read_textures(self, segments):
        return [open(self.man.texture_url(s), "rb").read() for s in segments]

    def decode(self, frames, segments, memory=1):
        """-> (meshMap, textureMap): dicts keyed by frame / segment number (src/V2/player.ts:68-69,328,362)."""
        g, t = self.player.decode_step_raw(self.read_geometry(frames), self.read_textures(segments), memory)
        return ({f: g[i] for i, f in enumerate(frames) if g[i].status == 0}, {s: t[i] for i, s in enumerate(segments) if t[i].status == 0})

    def decode_copy(self, frames, segments):
        """Like decode, with host results copied into numpy arrays that outlive the next batch (what V2Playback stores)."""
        import numpy as np
        g, t = self.decode(frames, segments, memory=1)
        arr = lambda p, shape, dt: np.ctypeslib.as_array(p, shape).view(dt).copy() if p else None
        meshes = {f: {"index": arr(x.index, (x.num_faces * 3,), np.uint32), "position": arr(x.position, (x.num_points, 3), np.float32),
                      "normal": arr(x.normal, (x.num_points, 3), np.float32), "uv": arr(x.uv, (x.num_points, 2), np.float32)} for f, x in g.items()}
        textures = {s: {"width": x.width, "height": x.height, "layers": x.layers, "data": np.ctypeslib.as_array(x.data, (x.layers, x.height, x.width, 4)).copy()} for s, x in t.items()}
        return meshes, textures


class V1Manifest:
    """A V1 manifest (src/Interfaces.ts:1-15; writer deprecated/encoder/src/Encoder30.js:155-160)."""

    def __init__(self, manifest, manifest_path):
        for k in ("maxVertices", "maxTriangles", "frameData", "frameRate"):
            if k not in manifest:
                raise ValueError(f"V1 manifest lacks {k}")
        self.m, self.path = manifest, manifest_path

    @classmethod
    def load(cls, path):
        with open(path) as fh:
            return cls(json.load(fh), path)

    @property
    def mesh_file(self):
        """src/V1/player.ts:337: manifestFilePath.replace('.manifest', '.drcs')"""
        return self.path.replace(".manifest", ".drcs", 1)

    def byte_range(self, frame_start, frame_end):
        """src/V1/worker.ts:33-40: one Range request covering frames [frame_start, frame_end)."""
        fd = self.m["frameData"]
        start = fd[frame_start]["startBytePosition"]
        last = fd[frame_end - 1]
        return start, last["startBytePosition"] + last["meshLength"]

    def slices(self, blob, base, frame_start, frame_end):
        """src/V1/worker.ts:48-56: per-frame slices of the fetched range (each copied, hence aligned)."""
        out = []
        for i in range(frame_start, frame_end):
            fd = self.m["frameData"][i]
            s = fd["startBytePosition"] - base
            out.append((fd["frameNumber"], fd["keyframeNumber"], bytes(blob[s:s + fd["meshLength"]])))
        return out


class V1Sequence:
    """A V1 clip on local storage (`clip.manifest` + `clip.drcs`): the worker's fetch-slice-decode loop (src/V1/worker.ts:24-74) as
    one batched call -- read the byte range of frames [start, end), cut the per-frame `.crt` slices, decode them all at once and
    emit what the worker posts per frame: {frameNumber, keyframeNumber, bufferGeometry{index, position, uv}} (:58-66), which
    `handleFrameData` keys by keyframeNumber (src/V1/player.ts:289-303)."""

    def __init__(self, manifest_path, decoder):
        self.man = V1Manifest.load(manifest_path)
        self.decoder = decoder                     # universal-volumetric_b200.CortoDecoder

    def read_range(self, frame_start, frame_end):
        lo, hi = self.man.byte_range(frame_start, frame_end)
        with open(self.man.mesh_file, "rb") as fh:
            fh.seek(lo)
            return lo, fh.read(hi - lo)

    def decode(self, frame_start, frame_end):
        """-> {keyframeNumber: {"frameNumber", "keyframeNumber", "bufferGeometry": {"index", "position", "uv"}}}; frames that fail to
        decode are simply absent (the worker logs and skips them)."""
        frame_end = min(frame_end, len(self.man.m["frameData"]))
        if frame_end <= frame_start:
            return {}
        base, blob = self.read_range(frame_start, frame_end)
        sl = self.man.slices(blob, base, frame_start, frame_end)
        res = self.decoder.decode_batch([b for _, _, b in sl])
        return {kf: {"frameNumber": fn, "keyframeNumber": kf, "bufferGeometry": {k: r[k] for k in ("index", "position", "uv")}}
                for (fn, kf, _), r in zip(sl, res) if r["status"] == 0}


class V2Playback:
    """Host-side playback buffer of the V2 player: keeps `buffer_duration` seconds decoded ahead of a clock.

    Mirrors src/V2/player.ts: `fetch_buffers` = fetchBuffers' leaky bucket (:272-323; every call requests what is missing up
    to `buffer_duration` seconds ahead and hands ALL of it to the library as one batch, instead of one worker request per
    file), `process_frame` = processFrame's selection rule (:388-470: geometry first; a frame whose mesh is missing is
    skipped, a missing texture segment gives the mesh without texture -- the reference's failMaterial), `update` =
    update + removePlayedBuffer (:531-562).  `decode(frames, segments) -> (meshMap part, textureMap part)` is injected:
    `V2Sequence.decode` copies in production, a stub in the CPU tests.  Results are copied out of the library-owned
    arenas (they are only valid until the next batch), like the transferables the workers post back.
    """

    def __init__(self, man, decode, buffer_duration=4):
        self.man, self.decode, self.buffer_duration = man, decode, buffer_duration
        self.mesh_map, self.texture_map = {}, {}                 # frame -> geometry, segment -> texture (:68-69)
        self.last_geometry, self.last_segment = -1, -1           # lastRequestedGeometryFrame / lastRequestedTextureSegment
        self.requests = 0

    def fetch_buffers(self, t):
        geo, tex, self.last_geometry, self.last_segment = self.man.fetch_window(t, self.last_geometry, self.last_segment, self.buffer_duration)
        if geo or tex:
            meshes, textures = self.decode(geo, tex)
            self.mesh_map.update(meshes); self.texture_map.update(textures); self.requests += 1
        return len(geo), len(tex)

    def buffered_fraction(self):
        """onMeshBuffering's argument (:409): share of the look-ahead window that is decoded."""
        return len(self.mesh_map) / (self.man.geometry["frameRate"] * self.buffer_duration)

    def process_frame(self, t):
        """What the renderer shows at time t: None (track ended or mesh not decoded: the frame is skipped), else
        {"frame", "geometry", "segment", "layer", "texture"} with texture None when its segment is missing."""
        at = self.man.frames_at(t)
        g = at["geometry_frame"]
        if g >= self.man.geometry_frame_count or g not in self.mesh_map:
            return None
        return {"frame": g, "geometry": self.mesh_map[g], "segment": at["segment"], "layer": at["layer"], "texture": self.texture_map.get(at["segment"])}

    def remove_played_buffer(self, frame_no, segment_no):
        for k in [k for k in self.mesh_map if k < frame_no]:
            del self.mesh_map[k]
        for k in [k for k in self.texture_map if k < segment_no]:
            del self.texture_map[k]

    def update(self, t):
        shown = self.process_frame(t)
        at = self.man.frames_at(t)
        g_keep = math.ceil(120 / self.man.geometry["frameRate"])                                  # :544-546 (screens up to 120 Hz)
        s_keep = math.ceil(120 / (self.man.texture["frameRate"] * self.man.batch_size))
        self.remove_played_buffer(at["geometry_frame"] - g_keep, at["segment"] - s_keep)
        return shown


# ---- manifest tooling: the two V2 dialects, the V1 writer, the encoder's frame-count check -----------------
def normalize_v2(manifest):
    """The reference has TWO V2 dialects: the player's schema (src/Interfaces.ts:75-132: `geometry.targets{name: {...}}`, one
    `geometry.path` template with [target] / [ext] / [####] tags, `texture.targets{name: {...}}`) and what its encoder script
    writes (scripts/Encoder.py:311-328: a flat `geometry{format, frameRate, frameCount, path}` and `texture.targets[ {...,
    path} ]` as a LIST, paths with [####] only).  Returns the player's schema for either input (a copy)."""
    m = json.loads(json.dumps(manifest))
    if m.get("version") != "v2":
        raise ValueError("not a V2 manifest")
    g = m["geometry"]
    if "targets" not in g:                                   # encoder dialect
        g = {"targets": {g["format"]: {"format": g["format"], "frameRate": g["frameRate"], "frameCount": g["frameCount"]}}, "path": g["path"]}
        m["geometry"] = g
    t = m["texture"]
    if isinstance(t.get("targets"), list):
        targets, path = {}, None
        for i, e in enumerate(t["targets"]):
            e = dict(e); p = e.pop("path", None); path = path or p
            e.setdefault("type", V2Manifest.texture_type); e.setdefault("tag", V2Manifest.texture_tag)
            targets[e["format"] if e["format"] not in targets else "%s-%d" % (e["format"], i)] = e
        m["texture"] = {"targets": targets, "path": path if path is not None else t.get("path")}
    return m


def emit_v2(geometry_path, geometry_frame_rate, geometry_frame_count, texture_path, texture_frame_rate, sequence_size, sequence_count,
            dialect="player", resolution=None, audio=None):
    """A V2 manifest dict in the player's schema (dialect="player") or exactly as scripts/Encoder.py:311-328 writes it
    (dialect="encoder").  Paths are templates relative to the manifest ([#####] = zero-padded index)."""
    if dialect == "encoder":
        m = {"version": "v2", "geometry": {"format": "draco", "frameRate": geometry_frame_rate, "frameCount": geometry_frame_count, "path": geometry_path},
             "texture": {"targets": [{"format": "ktx2", "frameRate": texture_frame_rate, "sequenceCount": sequence_count, "sequenceSize": sequence_size, "path": texture_path}]}}
    elif dialect == "player":
        tex = {"format": "ktx2", "type": V2Manifest.texture_type, "tag": V2Manifest.texture_tag, "sequenceSize": sequence_size, "sequenceCount": sequence_count, "frameRate": texture_frame_rate}
        if resolution:
            tex["resolution"] = list(resolution)
        m = {"version": "v2", "geometry": {"targets": {"draco": {"format": "draco", "frameRate": geometry_frame_rate, "frameCount": geometry_frame_count}}, "path": geometry_path},
             "texture": {"targets": {"ktx2": tex}, "path": texture_path}}
    else:
        raise ValueError("dialect must be 'player' or 'encoder'")
    if audio:
        m["audio"] = dict(audio)
    return m


def ktx2_layer_count(blob):
    """layerCount of a KTX2 file (bytes 32..36 of the header), as scripts/Encoder.py:128-131 reads it."""
    import struct
    if len(blob) < 36 or blob[:12] != b"\xabKTX 20\xbb\r\n\x1a\n":
        raise ValueError("not a KTX2 file")
    return struct.unpack_from("<I", blob, 32)[0]


def check_total_frames(geometry_frame_count, geometry_frame_rate, texture_segments, sequence_size, texture_frame_rate, read=None):
    """scripts/Encoder.py:103-154: texture frames = (segments - 1) * sequenceSize + layerCount of the LAST segment (it may be short);
    compatible iff geometry_frames * texture_fps == texture_frames * geometry_fps.  `texture_segments` = paths (or blobs when
    `read` is None and items are bytes).  Returns {"compatible", "geometry_frames", "texture_frames", "durations"}."""
    if not texture_segments:
        raise ValueError("no texture segments")
    last = texture_segments[-1]
    blob = last if isinstance(last, (bytes, bytearray)) else (read(last) if read else open(last, "rb").read())
    tex_frames = (len(texture_segments) - 1) * sequence_size + max(1, ktx2_layer_count(blob))
    return {"compatible": geometry_frame_count * texture_frame_rate == tex_frames * geometry_frame_rate,
            "geometry_frames": geometry_frame_count, "texture_frames": tex_frames,
            "durations": {"geometry": geometry_frame_count / geometry_frame_rate, "texture": tex_frames / texture_frame_rate}}


def emit_v1(frame_rate, frames):
    """The V1 `.manifest` next to a `.drcs` (deprecated/encoder/src/Encoder30.js:155-160; schema src/Interfaces.ts:1-15).
    `frames` = [(vertices, faces, crt_byte_length)] in file order; every frame is a keyframe of itself, as the encoder writes."""
    pos, data, maxv, maxf = 0, [], 0, 0
    for i, (nv, nf, nbytes) in enumerate(frames):
        data.append({"frameNumber": i, "keyframeNumber": i, "startBytePosition": pos, "vertices": nv, "faces": nf, "meshLength": nbytes})
        pos += nbytes; maxv = max(maxv, nv); maxf = max(maxf, nf)
    return {"frameRate": frame_rate, "maxVertices": maxv, "maxTriangles": maxf, "frameData": data}

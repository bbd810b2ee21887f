"""Host-side mirror of the reference's loader interface for the decode hot path.

The reference host is TypeScript (no node toolchain in this image), so the host side above the
C ABI is Python here, keeping the reference's names and result shapes:

* ``DRACOLoader.decode_batch`` ~ ``DRACOLoader.decodeGeometry`` / worker ``decodeGeometry``
  (src/lib/DRACOLoader.js:104-187,470-554): result = ``{"index": u32[F*3], "attributes":
  {"position": f32[P,3], "normal": f32[P,3], "uv": f32[P,2]}}`` in Draco point order.
* ``KTX2Loader.transcode_batch`` ~ ``BasisWorker.transcode`` (src/lib/KTX2Loader.js:469-580):
  result = ``{"width", "height", "layers", "hasAlpha", "format", "dfdTransferFn", "dfdFlags",
  "data": u8[layers,h,w,4]}`` (all layers concatenated, :565).

Every call goes through libuvol_b200.so; a failed item raises nothing and is reported with
``status < 0`` / ``None`` arrays, like a frame that is simply missing from ``meshMap``
(src/V2/player.ts:429-444).
"""
import ctypes

import numpy as np

from . import _native as N


class Context:
    """One context per GPU (mirrors one decoder instance per worker, DRACOLoader.js:439)."""

    def __init__(self, device=0, profiling=False, texture_target=None, corto_index_u16=None, **limits):
        """texture_target / corto_index_u16 / max_faces_per_frame / max_texture_bytes / buffer_duration_s / interval_duration_s override
        the defaults (which already include the UVOL_* environment overrides, uvol_config_default)."""
        L = N.lib()
        h = ctypes.c_void_p()
        cfg = N.Config(); L.uvol_config_default(ctypes.byref(cfg))
        if texture_target is not None:
            cfg.texture_target = int(texture_target)
        if corto_index_u16 is not None:
            cfg.corto_index_u16 = int(bool(corto_index_u16))
        for k, v in limits.items():
            if not hasattr(cfg, k):
                raise TypeError(f"unknown uvol_config field {k}")
            setattr(cfg, k, v)
        rc = L.uvol_create_with_config(int(device), ctypes.byref(cfg), ctypes.byref(h))
        if rc != 0 or not h:
            raise N.UvolError(f"uvol_create(device={device}) failed with status {rc}: a CUDA device is required "
                              "(there is no CPU fallback)")
        self._h, self._L = h, L
        if profiling:
            L.uvol_set_profiling(h, 1)

    def close(self):
        if getattr(self, "_h", None):
            self._L.uvol_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def share_arenas(self, owner):
        """Use `owner`'s phase-2 geometry scratch (one ctx per window of a sequence; calls may then run from one thread per ctx)."""
        if self._L.uvol_share_arenas(self._h, owner._h) != 0:
            raise N.UvolError("uvol_share_arenas failed")
        self._owner = owner          # keep the owner alive

    def share_host_outputs(self, owner):
        """Return UVOL_MEM_HOST results in `owner`'s pinned buffers (such calls on the two contexts must then be serialised)."""
        if self._L.uvol_share_host_outputs(self._h, owner._h) != 0:
            raise N.UvolError("uvol_share_host_outputs failed")
        self._owner_host = owner

    def config(self):
        cfg = N.Config(); self._L.uvol_get_config(self._h, ctypes.byref(cfg))
        return {k: getattr(cfg, k) for k, _ in N.Config._fields_}

    def flush_l2(self):
        self._L.uvol_flush_l2(self._h)

    def set_profiling(self, on):
        self._L.uvol_set_profiling(self._h, 1 if on else 0)

    def last_error(self):
        return self._L.uvol_last_error(self._h).decode()

    def stats(self, kind=0, combined=False):
        """Statistics of the last call; with combined=True those of the last V2 step (kind 0 geometry, 1 texture)."""
        s = N.Stats()
        if combined:
            self._L.uvol_get_stats_kind(self._h, kind, ctypes.byref(s))
        else:
            self._L.uvol_get_stats(self._h, ctypes.byref(s))
        d = {k: getattr(s, k) for k in ("host_parse_ms", "h2d_ms", "device_ms", "d2h_ms", "total_ms", "kernel_launches", "bytes_in",
                                        "bytes_out", "scratch_bytes")}
        d["stages"] = {self._L.uvol_stage_name(kind, i).decode(): float(s.stage_ms[i]) for i in range(s.num_stages)}
        return d


def span_ms(contexts):
    """Device time (CUDA events) from the first kernel to the last kernel of the last calls of `contexts` (run concurrently)."""
    L = N.lib()
    arr = (ctypes.c_void_p * len(contexts))(*[c._h for c in contexts])
    ms = ctypes.c_float()
    if L.uvol_span_ms(arr, len(contexts), ctypes.byref(ms)) != 0:
        raise N.UvolError("uvol_span_ms failed (profiling off, or no call yet)")
    return float(ms.value)


def _pack(files):
    n = len(files)
    keep = [bytes(f) if not isinstance(f, (bytes, bytearray)) else f for f in files]
    ptrs = (ctypes.c_void_p * n)()
    sizes = (ctypes.c_size_t * n)()
    for i, b in enumerate(keep):
        ptrs[i] = ctypes.cast(ctypes.c_char_p(bytes(b)) if isinstance(b, bytearray) else ctypes.c_char_p(b), ctypes.c_void_p)
        sizes[i] = len(b)
    return keep, ptrs, sizes


class DRACOLoader:
    """Batch replacement of the reference's Draco worker pool (<=4 workers, DRACOLoader.js:24,312-364)."""

    def __init__(self, ctx=None, device=0):
        self.ctx = ctx or Context(device)

    def decode_batch_raw(self, files, memory=N.MEM_HOST):
        """Runs uvol_decode_draco_batch; returns the raw result structs (pointers stay valid until
        the next geometry batch on this context)."""
        keep, ptrs, sizes = _pack(files)
        out = (N.Geometry * len(files))()
        rc = self.ctx._L.uvol_decode_draco_batch(self.ctx._h, ptrs, sizes, len(files), memory, out)
        if rc != 0:
            raise N.UvolError(f"uvol_decode_draco_batch failed ({rc}): {self.ctx.last_error()}")
        self._keep = keep
        return out

    def replay_raw(self, n, memory=N.MEM_DEVICE):
        """Re-runs the device pipeline on the batch still resident in HBM (measurement aid)."""
        out = (N.Geometry * n)()
        rc = self.ctx._L.uvol_replay_draco_batch(self.ctx._h, memory, out, n)
        if rc != 0:
            raise N.UvolError(f"uvol_replay_draco_batch failed ({rc}): {self.ctx.last_error()}")
        return out

    def decode_batch(self, files):
        """Decodes .drc byte strings; returns one dict per file (numpy copies of the host buffers)."""
        raw = self.decode_batch_raw(files, N.MEM_HOST)
        res = []
        for g in raw:
            if g.status != 0:
                res.append({"status": int(g.status), "index": None, "attributes": {}})
                continue
            P, F = g.num_points, g.num_faces
            attrs = {"position": np.ctypeslib.as_array(g.position, (P, 3)).copy()}
            if g.normal:
                attrs["normal"] = np.ctypeslib.as_array(g.normal, (P, 3)).copy()
            if g.uv:
                attrs["uv"] = np.ctypeslib.as_array(g.uv, (P, 2)).copy()
            if g.color:
                attrs["color"] = np.ctypeslib.as_array(g.color, (P, g.color_components)).copy()
            res.append({"status": 0, "index": np.ctypeslib.as_array(g.index, (F * 3,)).copy(), "attributes": attrs,
                        "num_points": int(P), "num_faces": int(F)})
        return res


def ktx2_probe(blob):
    """The KTX2File getters (src/lib/KTX2Loader.js:471-495) of one .ktx2, from its header alone (host code, no GPU): dict, or None with a status."""
    i = N.Ktx2Info(); b = bytes(blob)
    rc = N.lib().uvol_ktx2_probe(b, ctypes.c_size_t(len(b)), ctypes.byref(i))
    return (rc, None) if rc else (0, {k: int(getattr(i, k)) for k, _ in N.Ktx2Info._fields_})


def pick_texture_format(is_uastc, has_alpha, **config):
    """getTranscoderFormat (src/lib/KTX2Loader.js:659-689) for a context described by the reference's own workerConfig keys, e.g.
    pick_texture_format(True, False, astcSupported=True, bptcSupported=True) -> TEX_ASTC_4x4."""
    caps = 0
    for k, v in config.items():
        caps |= N.CAPS[k] if v else 0
    return int(N.lib().uvol_pick_texture_format(int(bool(is_uastc)), int(bool(has_alpha)), caps))


class KTX2Loader:
    """Batch replacement of the reference's Basis worker pool (<=4 workers FIFO, WorkerPool.js:5-102)."""

    def __init__(self, ctx=None, device=0):
        self.ctx = ctx or Context(device)

    def transcode_batch_raw(self, files, memory=N.MEM_HOST, target=N.TEX_RGBA32):
        keep, ptrs, sizes = _pack(files)
        out = (N.Texture * len(files))()
        rc = self.ctx._L.uvol_transcode_ktx2_batch(self.ctx._h, ptrs, sizes, len(files), target, memory, out)
        if rc != 0:
            raise N.UvolError(f"uvol_transcode_ktx2_batch failed ({rc}): {self.ctx.last_error()}")
        self._keep = keep
        return out

    def replay_raw(self, n, memory=N.MEM_DEVICE):
        out = (N.Texture * n)()
        rc = self.ctx._L.uvol_replay_ktx2_batch(self.ctx._h, memory, out, n)
        if rc != 0:
            raise N.UvolError(f"uvol_replay_ktx2_batch failed ({rc}): {self.ctx.last_error()}")
        return out

    def transcode_batch(self, files, target=N.TEX_RGBA32):
        """target = TEX_RGBA32 (default; data u8[layers, h, w, 4]), TEX_ETC1 (data u8[layers, blocks, 8], the reference's
        RGB_ETC1_Format / opaque RGB_ETC2_Format choice, KTX2Loader.js:619-636) or TEX_BC7 (data u8[layers, blocks, 16], its
        RGBA_BPTC_Format choice on desktop GPUs, :602-604) or TEX_ASTC_4x4 (u8[layers, blocks, 16], its RGBA_ASTC_4x4_Format choice
        for UASTC sources, :592-600; lossless) or TEX_ETC2_RGBA (u8[layers, blocks, 16]: EAC alpha block + ETC1 colour block, its
        RGBA_ETC2_EAC_Format choice for ETC1S sources, :619-627) or TEX_BC1 / TEX_BC3 (u8[layers, blocks, 8 / 16], its RGB_S3TC_DXT1 /
        RGBA_S3TC_DXT5 fallback, :610-618; ETC1S sources)."""
        raw = self.transcode_batch_raw(files, N.MEM_HOST, target)
        res = []
        for t in raw:
            if t.status != 0:
                res.append({"status": int(t.status), "data": None})
                continue
            def level_array(w, h, offset):
                p = ctypes.cast(ctypes.addressof(t.data.contents) + offset, ctypes.POINTER(ctypes.c_uint8))
                if target in (N.TEX_ETC1, N.TEX_BC7, N.TEX_ASTC_4x4, N.TEX_ETC2_RGBA, N.TEX_BC1, N.TEX_BC3):
                    return np.ctypeslib.as_array(p, (t.layers, ((w + 3) // 4) * ((h + 3) // 4), 8 if target in (N.TEX_ETC1, N.TEX_BC1) else 16)).copy()
                return np.ctypeslib.as_array(p, (t.layers, h, w, 4)).copy()
            data = level_array(t.width, t.height, 0)
            res.append({"status": 0, "width": int(t.width), "height": int(t.height), "layers": int(t.layers), "hasAlpha": bool(t.has_alpha),
                        "format": {N.TEX_ETC1: "RGB_ETC1_Format", N.TEX_BC7: "RGBA_BPTC_Format", N.TEX_ASTC_4x4: "RGBA_ASTC_4x4_Format", N.TEX_ETC2_RGBA: "RGBA_ETC2_EAC_Format", N.TEX_BC1: "RGB_S3TC_DXT1_Format", N.TEX_BC3: "RGBA_S3TC_DXT5_Format"}.get(target, "RGBAFormat"), "dfdTransferFn": int(t.dfd_transfer), "dfdFlags": int(t.dfd_flags), "data": data,
                        # the reply's `mipmaps` (KTX2Loader.js:514-573): level 0 is `data`; UVOL content has exactly one level
                        "mipmaps": [{"width": int(t.mips[k].width), "height": int(t.mips[k].height), "data": data if k == 0 else level_array(t.mips[k].width, t.mips[k].height, t.mips[k].offset)}
                                    for k in range(int(t.levels))]})
        return res


class V2Player:
    """Batch mirror of the decode side of the reference's V2 player (src/V2/player.ts): one step hands a
    range of geometry frames and texture segments to the library, which decodes both kinds concurrently
    (the reference issues decodeDraco / decodeKTX2 promises to two worker pools, :272-323, and stores
    the results in meshMap / textureMap keyed by frame / segment number, :325-331,359-366)."""

    def __init__(self, ctx=None, device=0):
        self.ctx = ctx or Context(device)

    def decode_step_raw(self, drc_files, ktx2_files, memory=N.MEM_HOST):
        kd, pd, sd = _pack(drc_files); kk, pk, sk = _pack(ktx2_files)
        og = (N.Geometry * max(1, len(drc_files)))(); ot = (N.Texture * max(1, len(ktx2_files)))()
        rc = self.ctx._L.uvol_decode_v2_batch(self.ctx._h, pd, sd, len(drc_files), pk, sk, len(ktx2_files), memory, og, ot)
        if rc != 0:
            raise N.UvolError(f"uvol_decode_v2_batch failed ({rc}): {self.ctx.last_error()}")
        self._keep = (kd, kk)
        return og, ot

    def replay_step_raw(self, n_drc, n_ktx2, memory=N.MEM_DEVICE):
        og = (N.Geometry * max(1, n_drc))(); ot = (N.Texture * max(1, n_ktx2))()
        rc = self.ctx._L.uvol_replay_v2_batch(self.ctx._h, memory, og, n_drc, ot, n_ktx2)
        if rc != 0:
            raise N.UvolError(f"uvol_replay_v2_batch failed ({rc}): {self.ctx.last_error()}")
        return og, ot


class CortoDecoder:
    """Batch mirror of the V1 worker loop `new CortoDecoder(slice).decode()` (src/V1/worker.ts:48-68, src/lib/corto.ts:73-140).
    `decode_batch` takes the per-frame .crt slices of a .drcs and returns, per frame, the bufferGeometry of
    src/V1/player.ts:289-297: {"index", "position", "uv"} -- plus "normal" (f32[V,3]) / "color" (u8[V,4]) when the file carries them
    (src/lib/corto.ts:439-671).  With Context(corto_index_u16=True) the index is a uint16 array whenever nface < 65536, the
    reference's JS layout (corto.ts:675-680, player.ts:292)."""

    def __init__(self, ctx=None, device=0):
        self.ctx = ctx or Context(device)

    def decode_batch_raw(self, frames, memory=N.MEM_HOST):
        keep, ptrs, sizes = _pack(frames)
        out = (N.CortoMesh * max(1, len(frames)))()
        rc = self.ctx._L.uvol_decode_corto_batch(self.ctx._h, ptrs, sizes, len(frames), memory, out)
        if rc != 0:
            raise N.UvolError(f"uvol_decode_corto_batch failed ({rc}): {self.ctx.last_error()}")
        self._keep = keep
        return out

    def decode_batch(self, frames):
        raw = self.decode_batch_raw(frames, N.MEM_HOST)
        res = []
        for m in raw[: len(frames)]:
            if m.status != 0:
                res.append({"status": int(m.status)})
                continue
            V, F = m.num_vertices, m.num_faces
            index = np.ctypeslib.as_array(m.index16, (F * 3,)).copy() if m.index_type == 1 and m.index16 else np.ctypeslib.as_array(m.index, (F * 3,)).copy()
            res.append({"status": 0, "index": index, "position": np.ctypeslib.as_array(m.position, (V, 3)).copy(),
                        "uv": np.ctypeslib.as_array(m.uv, (V, 2)).copy() if m.uv else None,
                        "normal": np.ctypeslib.as_array(m.normal, (V, 3)).copy() if m.normal else None,
                        "color": np.ctypeslib.as_array(m.color, (V, 4)).copy() if m.color else None})
        return res

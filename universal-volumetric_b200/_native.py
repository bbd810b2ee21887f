"""ctypes binding of libuvol_b200.so (the C ABI declared in include/uvol_b200.h).

The shared library is the product; this module only marshals pointers.  It fails loudly when the
library is missing or when no CUDA device is usable -- there is no CPU decode path.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libuvol_b200.so")

MEM_DEVICE, MEM_HOST = 0, 1
TEX_RGBA32 = 0
TEX_ETC1 = 1
TEX_BC7 = 2
TEX_ASTC_4x4 = 4
TEX_ETC2_RGBA = 5
TEX_BC1 = 6
TEX_BC3 = 7
TEX_ETC2_RGB = 3


class UvolError(RuntimeError):
    pass


class Geometry(ctypes.Structure):
    _fields_ = [("status", ctypes.c_int32), ("num_points", ctypes.c_uint32), ("num_faces", ctypes.c_uint32),
                ("color_components", ctypes.c_uint32),
                ("index", ctypes.POINTER(ctypes.c_uint32)), ("position", ctypes.POINTER(ctypes.c_float)),
                ("normal", ctypes.POINTER(ctypes.c_float)), ("uv", ctypes.POINTER(ctypes.c_float)),
                ("color", ctypes.POINTER(ctypes.c_float))]


class TextureLevel(ctypes.Structure):
    _fields_ = [("width", ctypes.c_uint32), ("height", ctypes.c_uint32), ("offset", ctypes.c_uint64), ("bytes", ctypes.c_uint64)]


class Ktx2Info(ctypes.Structure):
    _fields_ = [(k, ctypes.c_uint32) for k in ("width", "height", "layers", "levels", "faces", "is_uastc", "has_alpha", "is_video", "supercompression", "dfd_transfer", "dfd_flags")]


CAPS = {"astcSupported": 1, "bptcSupported": 2, "dxtSupported": 4, "etc2Supported": 8, "etc1Supported": 16, "pvrtcSupported": 32}      # workerConfig keys (KTX2Loader.js:113-149)


class Texture(ctypes.Structure):
    _fields_ = [("status", ctypes.c_int32), ("width", ctypes.c_uint32), ("height", ctypes.c_uint32), ("layers", ctypes.c_uint32),
                ("format", ctypes.c_uint32), ("has_alpha", ctypes.c_uint32), ("dfd_transfer", ctypes.c_uint32), ("dfd_flags", ctypes.c_uint32),
                ("data", ctypes.POINTER(ctypes.c_uint8)), ("bytes", ctypes.c_uint64),
                ("levels", ctypes.c_uint32), ("reserved", ctypes.c_uint32), ("mips", ctypes.POINTER(TextureLevel))]


class CortoMesh(ctypes.Structure):
    _fields_ = [("status", ctypes.c_int32), ("num_vertices", ctypes.c_uint32), ("num_faces", ctypes.c_uint32), ("index_type", ctypes.c_uint32),
                ("index", ctypes.POINTER(ctypes.c_uint32)), ("position", ctypes.POINTER(ctypes.c_float)), ("uv", ctypes.POINTER(ctypes.c_float)),
                ("normal", ctypes.POINTER(ctypes.c_float)), ("color", ctypes.POINTER(ctypes.c_uint8)), ("index16", ctypes.POINTER(ctypes.c_uint16))]


class Config(ctypes.Structure):
    """uvol_config (include/uvol_b200.h)."""
    _fields_ = [("struct_size", ctypes.c_uint32), ("texture_target", ctypes.c_uint32), ("corto_index_u16", ctypes.c_uint32), ("staging_threads", ctypes.c_uint32),
                ("max_faces_per_frame", ctypes.c_uint64), ("max_texture_bytes", ctypes.c_uint64), ("buffer_duration_s", ctypes.c_double), ("interval_duration_s", ctypes.c_double)]


class SequenceInfo(ctypes.Structure):
    _fields_ = [("version", ctypes.c_int32), ("geometry_frame_count", ctypes.c_uint32), ("geometry_frame_rate", ctypes.c_double), ("texture_frame_rate", ctypes.c_double),
                ("sequence_size", ctypes.c_uint32), ("sequence_count", ctypes.c_uint32), ("max_vertices", ctypes.c_uint32), ("max_triangles", ctypes.c_uint32)]


class Vector2(ctypes.Structure):
    _fields_ = [("x", ctypes.c_float), ("y", ctypes.c_float)]


class Stats(ctypes.Structure):
    _fields_ = [("host_parse_ms", ctypes.c_double), ("h2d_ms", ctypes.c_double), ("device_ms", ctypes.c_double), ("d2h_ms", ctypes.c_double),
                ("total_ms", ctypes.c_double), ("stage_ms", ctypes.c_float * 24), ("num_stages", ctypes.c_uint32),
                ("kernel_launches", ctypes.c_uint32), ("bytes_in", ctypes.c_uint64), ("bytes_out", ctypes.c_uint64),
                ("scratch_bytes", ctypes.c_uint64)]


_lib = None


def lib():
    """Loads libuvol_b200.so once.  Raises UvolError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise UvolError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(the CUDA extension is mandatory; there is no CPU fallback)")
    L = ctypes.CDLL(LIB_PATH)
    vp, i, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
    L.uvol_create.argtypes = [i, ctypes.POINTER(vp)]; L.uvol_create.restype = i
    L.uvol_config_default.argtypes = [ctypes.POINTER(Config)]; L.uvol_config_default.restype = None
    L.uvol_create_with_config.argtypes = [i, ctypes.POINTER(Config), ctypes.POINTER(vp)]; L.uvol_create_with_config.restype = i
    L.uvol_get_config.argtypes = [vp, ctypes.POINTER(Config)]; L.uvol_get_config.restype = i
    L.uvol_destroy.argtypes = [vp]; L.uvol_destroy.restype = None
    L.uvol_last_error.argtypes = [vp]; L.uvol_last_error.restype = ctypes.c_char_p
    L.uvol_get_stats.argtypes = [vp, ctypes.POINTER(Stats)]; L.uvol_get_stats.restype = i
    L.uvol_stage_name.argtypes = [i, i]; L.uvol_stage_name.restype = ctypes.c_char_p
    L.uvol_set_profiling.argtypes = [vp, i]; L.uvol_set_profiling.restype = i
    L.uvol_decode_draco_batch.argtypes = [vp, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(sz), i, i, ctypes.POINTER(Geometry)]
    L.uvol_decode_draco_batch.restype = i
    if hasattr(L, "uvol_transcode_ktx2_batch"):
        L.uvol_transcode_ktx2_batch.argtypes = [vp, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(sz), i, i, i, ctypes.POINTER(Texture)]
        L.uvol_transcode_ktx2_batch.restype = i
    L.uvol_replay_draco_batch.argtypes = [vp, i, ctypes.POINTER(Geometry), i]; L.uvol_replay_draco_batch.restype = i
    L.uvol_replay_ktx2_batch.argtypes = [vp, i, ctypes.POINTER(Texture), i]; L.uvol_replay_ktx2_batch.restype = i
    L.uvol_flush_l2.argtypes = [vp]; L.uvol_flush_l2.restype = i
    L.uvol_share_arenas.argtypes = [vp, vp]; L.uvol_share_arenas.restype = i
    L.uvol_share_host_outputs.argtypes = [vp, vp]; L.uvol_share_host_outputs.restype = i
    L.uvol_zstd_inflate.argtypes = [ctypes.c_char_p, sz, ctypes.c_void_p, sz, ctypes.POINTER(sz)]; L.uvol_zstd_inflate.restype = i
    L.uvol_span_ms.argtypes = [ctypes.POINTER(vp), i, ctypes.POINTER(ctypes.c_float)]; L.uvol_span_ms.restype = i
    pv, ps = ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(sz)
    L.uvol_decode_v2_batch.argtypes = [vp, pv, ps, i, pv, ps, i, i, ctypes.POINTER(Geometry), ctypes.POINTER(Texture)]; L.uvol_decode_v2_batch.restype = i
    L.uvol_replay_v2_batch.argtypes = [vp, i, ctypes.POINTER(Geometry), i, ctypes.POINTER(Texture), i]; L.uvol_replay_v2_batch.restype = i
    L.uvol_get_stats_kind.argtypes = [vp, i, ctypes.POINTER(Stats)]; L.uvol_get_stats_kind.restype = i
    L.uvol_decode_corto_batch.argtypes = [vp, pv, ps, i, i, ctypes.POINTER(CortoMesh)]; L.uvol_decode_corto_batch.restype = i
    L.uvol_open.argtypes = [vp, ctypes.c_char_p, ctypes.POINTER(vp)]; L.uvol_open.restype = i
    L.uvol_close.argtypes = [vp]; L.uvol_close.restype = None
    L.uvol_sequence_get_info.argtypes = [vp, ctypes.POINTER(SequenceInfo)]; L.uvol_sequence_get_info.restype = i
    L.uvol_sequence_url.argtypes = [vp, i, i, ctypes.c_char_p, sz]; L.uvol_sequence_url.restype = i
    L.uvol_sequence_frames_at.argtypes = [vp, ctypes.c_double, ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32)]; L.uvol_sequence_frames_at.restype = i
    L.uvol_decode_range.argtypes = [vp, i, i, i, i, i, ctypes.POINTER(Geometry), ctypes.POINTER(Texture)]; L.uvol_decode_range.restype = i
    L.uvol_decode_v1_range.argtypes = [vp, i, i, i, ctypes.POINTER(CortoMesh), ctypes.POINTER(ctypes.c_uint32)]; L.uvol_decode_v1_range.restype = i
    L.uvol_release.argtypes = [vp]; L.uvol_release.restype = i
    L.uvol_v1_frame_numbers.argtypes = [vp, vp, i, i, i, i, i, i, ctypes.POINTER(ctypes.c_int32)]; L.uvol_v1_frame_numbers.restype = i
    L.uvol_upload_etc2_batch.argtypes = [vp, pv, ps, i, i, i, i, ctypes.POINTER(Texture)]; L.uvol_upload_etc2_batch.restype = i
    L.CreateDecoder.argtypes = [i, ctypes.c_char_p, ctypes.POINTER(Vector2)]; L.CreateDecoder.restype = vp
    L.DestroyDecoder.argtypes = [vp]; L.DestroyDecoder.restype = None
    L.DecodeMesh.argtypes = [vp, vp, vp, vp, vp, vp]; L.DecodeMesh.restype = i
    _lib = L
    return L


EXPORTED_SYMBOLS = ["uvol_create", "uvol_create_with_config", "uvol_config_default", "uvol_get_config", "uvol_destroy", "uvol_last_error", "uvol_get_stats", "uvol_stage_name", "uvol_set_profiling",
                    "uvol_decode_draco_batch", "uvol_transcode_ktx2_batch", "uvol_replay_draco_batch", "uvol_replay_ktx2_batch", "uvol_flush_l2", "uvol_decode_v2_batch", "uvol_replay_v2_batch", "uvol_get_stats_kind", "uvol_decode_corto_batch",
                    "uvol_open", "uvol_close", "uvol_sequence_get_info", "uvol_sequence_url", "uvol_sequence_frames_at", "uvol_decode_range", "uvol_decode_v1_range", "uvol_release", "uvol_v1_frame_numbers", "uvol_upload_etc2_batch",
                    "uvol_share_arenas", "uvol_share_host_outputs", "uvol_span_ms", "uvol_zstd_inflate", "uvol_zstd_feature_counts",
                    "CreateDecoder", "DestroyDecoder", "DecodeMesh"]

"""ctypes binding of oracle/_ref/libcorto_ref.so: the reference's OWN Corto C++ codec compiled in place
(oracle/Makefile `ref`) plus the small shim that drives its public encoder.  TEST INFRASTRUCTURE."""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "oracle", "_ref", "libcorto_ref.so")
_lib = None


def available():
    return os.path.exists(PATH)


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(PATH)
        P = ctypes.POINTER
        L.corto_ref_encode.argtypes = [P(ctypes.c_float), P(ctypes.c_float), ctypes.c_uint32, P(ctypes.c_uint32), ctypes.c_uint32, ctypes.c_int, ctypes.c_int,
                                       P(P(ctypes.c_uint8)), P(ctypes.c_uint32), P(ctypes.c_uint32)]
        L.corto_ref_encode.restype = ctypes.c_size_t
        L.corto_ref_decode.argtypes = [ctypes.c_void_p, ctypes.c_int, P(ctypes.c_float), P(ctypes.c_float), P(ctypes.c_uint32), P(ctypes.c_uint32)]
        L.corto_ref_free.argtypes = [ctypes.c_void_p]
        L.corto_ref_encode2.argtypes = [P(ctypes.c_float), P(ctypes.c_float), P(ctypes.c_float), P(ctypes.c_uint8), ctypes.c_uint32, P(ctypes.c_uint32), ctypes.c_uint32,
                                        ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, P(P(ctypes.c_uint8)), P(ctypes.c_uint32), P(ctypes.c_uint32)]
        L.corto_ref_encode2.restype = ctypes.c_size_t
        L.corto_ref_decode2.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.corto_ref_clers.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, P(ctypes.c_uint32), ctypes.c_void_p, ctypes.c_int, P(ctypes.c_uint32)]
        _lib = L
    return _lib


def ref_encode(pos, uv, faces, pos_bits=12, uv_bits=12):
    """crt::Encoder (encoder.h:50-81) -> .crt bytes, and the vertex / face counts it kept."""
    L = lib()
    pos = np.ascontiguousarray(pos, np.float32); faces = np.ascontiguousarray(faces, np.uint32)
    uvp = None
    if uv is not None:
        uv = np.ascontiguousarray(uv, np.float32); uvp = uv.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
    out = ctypes.POINTER(ctypes.c_uint8)(); nv = ctypes.c_uint32(); nf = ctypes.c_uint32()
    n = L.corto_ref_encode(pos.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), uvp, len(pos), faces.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), len(faces),
                           pos_bits, uv_bits, ctypes.byref(out), ctypes.byref(nv), ctypes.byref(nf))
    assert n > 0
    blob = ctypes.string_at(out, n); L.corto_ref_free(out)
    return blob, nv.value, nf.value


def ref_decode(blob, nvert, nface, has_uv=True):
    """crt::Decoder::decode (decoder.cpp:122-173) -> index u32[F*3], position f32[V,3], uv f32[V,2]."""
    L = lib()
    buf = np.frombuffer(blob + b"\0" * 8, np.uint8).copy()            # 4-byte aligned copy (decoder.cpp:42-43)
    pos = np.zeros((nvert, 3), np.float32); uv = np.zeros((nvert, 2), np.float32); idx = np.zeros(nface * 3, np.uint32)
    rc = L.corto_ref_decode(buf.ctypes.data_as(ctypes.c_void_p), len(blob), pos.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                            uv.ctypes.data_as(ctypes.POINTER(ctypes.c_float)) if has_uv else None, idx.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), None)
    assert rc == nface, rc
    return idx, pos, (uv if has_uv else None)


def _aligned(blob):
    return np.frombuffer(blob + b"\0" * 8, np.uint8).copy()            # 4-byte aligned copy (decoder.cpp:42-43)


def ref_encode2(pos, faces, uv=None, normal=None, color=None, pos_bits=12, uv_bits=12, normal_bits=10, normal_pred=1, color_bits=6):
    """crt::Encoder with optional uv / normals (prediction 0 DIFF, 1 ESTIMATED, 2 BORDER) / RGBA8 colours -> (.crt bytes, nvert, nface)."""
    L = lib()
    fp = ctypes.POINTER(ctypes.c_float); bp = ctypes.POINTER(ctypes.c_uint8)
    pos = np.ascontiguousarray(pos, np.float32); faces = np.ascontiguousarray(faces, np.uint32)
    keep = [np.ascontiguousarray(a, dt) if a is not None else None for a, dt in ((uv, np.float32), (normal, np.float32), (color, np.uint8))]
    ptr = lambda a, t: a.ctypes.data_as(t) if a is not None else None
    out = ctypes.POINTER(ctypes.c_uint8)(); nv = ctypes.c_uint32(); nf = ctypes.c_uint32()
    n = L.corto_ref_encode2(pos.ctypes.data_as(fp), ptr(keep[0], fp), ptr(keep[1], fp), ptr(keep[2], bp), len(pos), faces.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), len(faces),
                            pos_bits, uv_bits, normal_bits, normal_pred, color_bits, ctypes.byref(out), ctypes.byref(nv), ctypes.byref(nf))
    assert n > 0
    blob = ctypes.string_at(out, n); L.corto_ref_free(out)
    return blob, nv.value, nf.value


def ref_decode2(blob, nvert, nface, has_uv=True, has_normal=False, has_color=False):
    """crt::Decoder::decode with every attribute -> dict(index, position, uv, normal f32[V,3], color u8[V,4])."""
    L = lib(); buf = _aligned(blob)
    pos = np.zeros((nvert, 3), np.float32); uv = np.zeros((nvert, 2), np.float32); nrm = np.zeros((nvert, 3), np.float32); col = np.zeros((nvert, 4), np.uint8)
    idx = np.zeros(nface * 3, np.uint32)
    rc = L.corto_ref_decode2(buf.ctypes.data, len(blob), pos.ctypes.data, uv.ctypes.data if has_uv else None, nrm.ctypes.data if has_normal else None,
                             col.ctypes.data if has_color else None, idx.ctypes.data)
    assert rc == nface, rc
    return {"index": idx, "position": pos, "uv": uv if has_uv else None, "normal": nrm if has_normal else None, "color": col if has_color else None}


def ref_prediction(blob, nvert, nface):
    """The reference decoder's parallelogram contexts (index.prediction) and faces."""
    L = lib(); buf = _aligned(blob)
    pos = np.zeros((nvert, 3), np.float32); idx = np.zeros(nface * 3, np.uint32); pred = np.zeros((nvert, 3), np.uint32)
    rc = L.corto_ref_decode(buf.ctypes.data_as(ctypes.c_void_p), len(blob), pos.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), None,
                            idx.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), pred.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)))
    assert rc == nface, rc
    return idx, pred


def ref_clers(blob, nface):
    """CLERS symbols, split bit-stream words and group ends as the reference's IndexAttribute::decode leaves them."""
    L = lib(); buf = _aligned(blob)
    cl = np.zeros(4 * nface + 64, np.uint8); words = np.zeros(len(blob) // 4 + 8, np.uint32); ge = np.zeros(64, np.uint32)
    nw = ctypes.c_uint32(); ng = ctypes.c_uint32()
    n = L.corto_ref_clers(buf.ctypes.data, len(blob), cl.ctypes.data, len(cl), words.ctypes.data, len(words), ctypes.byref(nw), ge.ctypes.data, len(ge), ctypes.byref(ng))
    assert n >= 0
    return cl[:n].copy(), words[:nw.value].copy(), ge[:ng.value].copy()

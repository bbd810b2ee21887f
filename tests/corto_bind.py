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

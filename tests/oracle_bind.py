"""ctypes binding of oracle/liboracle.so (the CPU oracles).  TEST INFRASTRUCTURE: imported only by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_lib = None

c_u32, c_i32, c_int, c_f = ctypes.c_uint32, ctypes.c_int32, ctypes.c_int, ctypes.c_float
P = ctypes.POINTER


class DracoMesh(ctypes.Structure):
    _fields_ = [("status", c_int), ("num_faces", c_u32), ("num_points", c_u32), ("num_vertices", c_u32), ("num_symbols", c_u32),
                ("index", P(c_u32)), ("position", P(c_f)), ("normal", P(c_f)), ("uv", P(c_f)), ("color", P(c_f)),
                ("symhist", c_u32 * 5), ("attr_vertices", c_u32 * 4), ("ctx_counters_zero", c_int), ("rans_terminal_ok", c_int),
                ("bytes_consumed", ctypes.c_size_t),
                ("pos_entries", c_int), ("uv_entries", c_int), ("nrm_entries", c_int), ("pos_wraps", c_int), ("uv_wraps", c_int),
                ("uv_orient_left", c_int), ("nrm_flips", c_int), ("pos_parallelograms", c_int),
                ("pos_wmin", c_i32), ("pos_wmax", c_i32), ("uv_wmin", c_i32), ("uv_wmax", c_i32),
                ("pos_q", P(c_i32)), ("uv_q", P(c_i32)), ("nrm_q", P(c_i32)), ("dbg_c2v", P(c_i32)), ("dbg_opp", P(c_i32))]


class Ktx2Image(ctypes.Structure):
    _fields_ = [("status", c_int), ("width", c_u32), ("height", c_u32), ("layers", c_u32), ("levels", c_u32), ("faces", c_u32),
                ("is_uastc", c_int), ("is_video", c_int), ("has_alpha", c_int), ("dfd_transfer", c_int), ("dfd_flags", c_int),
                ("rgba", P(ctypes.c_uint8)), ("rgba_bytes", ctypes.c_size_t),
                ("endpoint_count", c_u32), ("selector_count", c_u32), ("endpoints_bytes", c_u32), ("endpoints_used", c_u32),
                ("selectors_bytes", c_u32), ("selectors_used", c_u32), ("tables_bytes", c_u32), ("tables_used", c_u32),
                ("slices", c_u32), ("slices_exact", c_u32), ("pred_hist", c_u32 * 4),
                ("endpoint_idx", P(ctypes.c_uint16)), ("selector_idx", P(ctypes.c_uint16))]


def build_oracle():
    """Compiles oracle/liboracle.so (and oracle/_ref when the reference tree is present)."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)
    if os.path.isdir("/root/reference/deprecated/encoder/dev/src"):
        subprocess.run(["make", "-s", "-C", ORACLE_DIR, "ref"], check=True)


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build_oracle()
        L = ctypes.CDLL(path)
        L.uvo_draco_decode.argtypes = [ctypes.c_char_p, ctypes.c_size_t, P(DracoMesh)]
        L.uvo_draco_free.argtypes = [P(DracoMesh)]
        L.uvo_ktx2_decode.argtypes = [ctypes.c_char_p, ctypes.c_size_t, P(Ktx2Image)]
        L.uvo_ktx2_free.argtypes = [P(Ktx2Image)]
        _lib = L
    return _lib


def oracle_draco(blob, keep_debug=False):
    """Decodes one .drc with the CPU oracle -> dict of numpy arrays (copies) + A.4 oracle fields."""
    L = lib(); m = DracoMesh()
    rc = L.uvo_draco_decode(blob, len(blob), ctypes.byref(m))
    if rc != 0:
        return {"status": rc}
    Pn, F = m.num_points, m.num_faces
    out = {"status": 0, "num_points": int(Pn), "num_faces": int(F), "num_vertices": int(m.num_vertices),
           "index": np.ctypeslib.as_array(m.index, (F * 3,)).copy(),
           "position": np.ctypeslib.as_array(m.position, (Pn, 3)).copy(),
           "normal": np.ctypeslib.as_array(m.normal, (Pn, 3)).copy() if m.normal else None,
           "uv": np.ctypeslib.as_array(m.uv, (Pn, 2)).copy() if m.uv else None,
           "symhist": list(m.symhist), "attr_vertices": list(m.attr_vertices), "ctx_counters_zero": bool(m.ctx_counters_zero),
           "rans_terminal_ok": bool(m.rans_terminal_ok), "bytes_consumed": int(m.bytes_consumed),
           "entries": (m.pos_entries, m.uv_entries, m.nrm_entries), "pos_wraps": m.pos_wraps, "uv_wraps": m.uv_wraps,
           "uv_orient_left": m.uv_orient_left, "nrm_flips": m.nrm_flips, "pos_parallelograms": m.pos_parallelograms,
           "pos_bounds": (m.pos_wmin, m.pos_wmax), "uv_bounds": (m.uv_wmin, m.uv_wmax)}
    if keep_debug:
        out["pos_q"] = np.ctypeslib.as_array(m.pos_q, (m.pos_entries, 3)).copy()
        out["uv_q"] = np.ctypeslib.as_array(m.uv_q, (m.uv_entries, 2)).copy() if m.uv_q else None
        out["nrm_q"] = np.ctypeslib.as_array(m.nrm_q, (m.nrm_entries, 2)).copy() if m.nrm_q else None
    L.uvo_draco_free(ctypes.byref(m))
    return out


def oracle_ktx2(blob, keep_debug=False):
    L = lib(); m = Ktx2Image()
    rc = L.uvo_ktx2_decode(blob, len(blob), ctypes.byref(m))
    if rc != 0:
        return {"status": rc}
    out = {"status": 0, "width": int(m.width), "height": int(m.height), "layers": int(m.layers), "is_uastc": bool(m.is_uastc),
           "is_video": bool(m.is_video), "has_alpha": bool(m.has_alpha), "dfd_transfer": m.dfd_transfer, "dfd_flags": m.dfd_flags,
           "rgba": np.ctypeslib.as_array(m.rgba, (m.layers, m.height, m.width, 4)).copy(),
           "endpoint_count": m.endpoint_count, "selector_count": m.selector_count,
           "sections": ((m.endpoints_bytes, m.endpoints_used), (m.selectors_bytes, m.selectors_used), (m.tables_bytes, m.tables_used)),
           "slices": (m.slices, m.slices_exact), "pred_hist": list(m.pred_hist)}
    if keep_debug and m.endpoint_idx:
        nb = ((m.width + 3) // 4) * ((m.height + 3) // 4)
        out["endpoint_idx"] = np.ctypeslib.as_array(m.endpoint_idx, (m.layers, nb)).copy()
        out["selector_idx"] = np.ctypeslib.as_array(m.selector_idx, (m.layers, nb)).copy()
    L.uvo_ktx2_free(ctypes.byref(m))
    return out


def oracle_bc7_image(blocks, width, height):
    """Decodes BC7 blocks (block-raster order, u8[nblocks, 16]) with the oracle's independent BC7 decoder -> (u8[h, w, 4], bad blocks)."""
    L = lib()
    L.uvo_bc7_decode_image.argtypes = [ctypes.c_char_p, c_u32, c_u32, ctypes.c_void_p]; L.uvo_bc7_decode_image.restype = c_int
    out = np.zeros((height, width, 4), np.uint8)
    bad = L.uvo_bc7_decode_image(np.ascontiguousarray(blocks, dtype=np.uint8).tobytes(), width, height, out.ctypes.data)
    return out, bad


def oracle_astc_image(blocks, width, height):
    """Decodes ASTC 4x4 blocks (block-raster order, u8[nblocks, 16]) with the oracle's independent ASTC decoder -> (u8[h, w, 4], bad blocks)."""
    L = lib()
    L.uvo_astc_decode_image.argtypes = [ctypes.c_char_p, c_u32, c_u32, ctypes.c_void_p]; L.uvo_astc_decode_image.restype = c_int
    out = np.zeros((height, width, 4), np.uint8)
    bad = L.uvo_astc_decode_image(np.ascontiguousarray(blocks, dtype=np.uint8).tobytes(), width, height, out.ctypes.data)
    return out, bad

"""ctypes binding of the host-emulation harnesses in tests/tools (logic checks without a GPU)."""
import ctypes
import os
import subprocess

import numpy as np

TOOLS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools")


def _load(name):
    subprocess.run(["make", "-s", "-C", TOOLS], check=True)
    return ctypes.CDLL(os.path.join(TOOLS, name))


def emu_draco(blob):
    E = _load("libdraco_emu.so")
    P = ctypes.c_uint32(); F = ctypes.c_uint32()
    idx = ctypes.POINTER(ctypes.c_uint32)(); pos = ctypes.POINTER(ctypes.c_float)(); nrm = ctypes.POINTER(ctypes.c_float)(); uv = ctypes.POINTER(ctypes.c_float)()
    rc = E.draco_emu_decode(blob, ctypes.c_size_t(len(blob)), ctypes.byref(P), ctypes.byref(F), ctypes.byref(idx), ctypes.byref(pos), ctypes.byref(nrm), ctypes.byref(uv))
    if rc:
        return {"status": rc}
    p, f = P.value, F.value
    return {"status": 0, "num_points": p, "num_faces": f, "index": np.ctypeslib.as_array(idx, (3 * f,)).copy(),
            "position": np.ctypeslib.as_array(pos, (p, 3)).copy(), "normal": np.ctypeslib.as_array(nrm, (p, 3)).copy() if nrm else None,
            "uv": np.ctypeslib.as_array(uv, (p, 2)).copy() if uv else None}


def emu_ktx2(blob):
    E = _load("libbasis_emu.so")
    p = ctypes.POINTER(ctypes.c_uint8)(); w = ctypes.c_uint32(); h = ctypes.c_uint32(); l = ctypes.c_uint32()
    rc = E.basis_emu_decode(blob, ctypes.c_size_t(len(blob)), ctypes.byref(p), ctypes.byref(w), ctypes.byref(h), ctypes.byref(l))
    if rc:
        return {"status": rc}
    return {"status": 0, "rgba": np.ctypeslib.as_array(p, (l.value, h.value, w.value, 4)).copy()}


def emu_ktx2_etc1(blob):
    """Target ETC1 through the product's repack function on the host: u8[layers, blocks, 8]."""
    E = _load("libbasis_emu.so")
    p = ctypes.POINTER(ctypes.c_uint8)(); w = ctypes.c_uint32(); h = ctypes.c_uint32(); l = ctypes.c_uint32()
    rc = E.basis_emu_decode_etc1(blob, ctypes.c_size_t(len(blob)), ctypes.byref(p), ctypes.byref(w), ctypes.byref(h), ctypes.byref(l))
    if rc:
        return {"status": rc}
    nb = ((w.value + 3) // 4) * ((h.value + 3) // 4)
    return {"status": 0, "width": w.value, "height": h.value, "layers": l.value, "blocks": np.ctypeslib.as_array(p, (l.value, nb, 8)).copy()}


def emu_ktx2_bc7(blob):
    """Target BC7 through the product's per-block functions (csrc/bc7_core.h) on the host: u8[layers, blocks, 16]."""
    E = _load("libbasis_emu.so")
    p = ctypes.POINTER(ctypes.c_uint8)(); w = ctypes.c_uint32(); h = ctypes.c_uint32(); l = ctypes.c_uint32()
    rc = E.basis_emu_decode_bc7(blob, ctypes.c_size_t(len(blob)), ctypes.byref(p), ctypes.byref(w), ctypes.byref(h), ctypes.byref(l))
    if rc:
        return {"status": rc}
    nb = ((w.value + 3) // 4) * ((h.value + 3) // 4)
    return {"status": 0, "width": w.value, "height": h.value, "layers": l.value, "blocks": np.ctypeslib.as_array(p, (l.value, nb, 16)).copy()}


def emu_ktx2_astc(blob):
    """Target ASTC 4x4 through the product's per-block function (csrc/astc_core.h) on the host: u8[layers, blocks, 16] (UASTC sources only)."""
    E = _load("libbasis_emu.so")
    p = ctypes.POINTER(ctypes.c_uint8)(); w = ctypes.c_uint32(); h = ctypes.c_uint32(); l = ctypes.c_uint32()
    rc = E.basis_emu_decode_astc(blob, ctypes.c_size_t(len(blob)), ctypes.byref(p), ctypes.byref(w), ctypes.byref(h), ctypes.byref(l))
    if rc:
        return {"status": rc}
    nb = ((w.value + 3) // 4) * ((h.value + 3) // 4)
    return {"status": 0, "width": w.value, "height": h.value, "layers": l.value, "blocks": np.ctypeslib.as_array(p, (l.value, nb, 16)).copy()}


def emu_ktx2_split_levels(blob):
    """The product's mip-chain splitter (csrc/basis_parse.cpp uvol_ktx2_split_levels): (status or level count, [single-level .ktx2 bytes])."""
    E = _load("libbasis_emu.so")
    files = (ctypes.POINTER(ctypes.c_uint8) * 16)(); sizes = (ctypes.c_size_t * 16)()
    E.basis_emu_split_levels.restype = ctypes.c_int
    rc = E.basis_emu_split_levels(blob, ctypes.c_size_t(len(blob)), files, sizes, 16)
    out = [ctypes.string_at(files[k], sizes[k]) for k in range(max(0, min(rc, 16)))]
    return rc, out


def emu_ktx2_etc2a(blob):
    """Target ETC2 RGBA through the product's per-block functions (csrc/basis_core.h) on the host: u8[layers, blocks, 16] (ETC1S sources only)."""
    E = _load("libbasis_emu.so")
    p = ctypes.POINTER(ctypes.c_uint8)(); w = ctypes.c_uint32(); h = ctypes.c_uint32(); l = ctypes.c_uint32()
    rc = E.basis_emu_decode_etc2a(blob, ctypes.c_size_t(len(blob)), ctypes.byref(p), ctypes.byref(w), ctypes.byref(h), ctypes.byref(l))
    if rc:
        return {"status": rc}
    nb = ((w.value + 3) // 4) * ((h.value + 3) // 4)
    return {"status": 0, "width": w.value, "height": h.value, "layers": l.value, "blocks": np.ctypeslib.as_array(p, (l.value, nb, 16)).copy()}


def emu_ktx2_dxt(blob, bc3):
    """Targets BC1 (bc3 = False: u8[layers, blocks, 8]) / BC3 (u8[layers, blocks, 16]) through the product's per-block functions on the host."""
    E = _load("libbasis_emu.so")
    p = ctypes.POINTER(ctypes.c_uint8)(); w = ctypes.c_uint32(); h = ctypes.c_uint32(); l = ctypes.c_uint32()
    fn = E.basis_emu_decode_bc3 if bc3 else E.basis_emu_decode_bc1
    rc = fn(blob, ctypes.c_size_t(len(blob)), ctypes.byref(p), ctypes.byref(w), ctypes.byref(h), ctypes.byref(l))
    if rc:
        return {"status": rc}
    nb = ((w.value + 3) // 4) * ((h.value + 3) // 4)
    return {"status": 0, "width": w.value, "height": h.value, "layers": l.value, "blocks": np.ctypeslib.as_array(p, (l.value, nb, 16 if bc3 else 8)).copy()}

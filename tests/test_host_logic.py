"""CPU: host-side logic of the product -- the C-ABI library loads and exports every declared
symbol, refuses to run without a GPU (no CPU fallback), and the per-unit decode logic in
csrc/*_core.h (run through the host-emulation harness) is bit-exact against the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, read
from emu_bind import emu_draco, emu_ktx2
from oracle_bind import oracle_draco, oracle_ktx2


def declared_functions():
    names = []
    for hdr in ("uvol_b200.h", "corto_codec.h"):
        path = os.path.join(ROOT, "include", hdr)
        if not os.path.exists(path):
            continue
        src = re.sub(r"/\*.*?\*/", "", open(path).read(), flags=re.S)
        names += re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{)]*\)\s*;", src)
    return sorted(set(n for n in names if n not in ("defined",)))


def test_library_exports_every_declared_symbol(uv):
    L = uv._native.lib()
    fns = declared_functions()
    assert "uvol_decode_draco_batch" in fns and "uvol_transcode_ktx2_batch" in fns
    for name in fns:
        assert hasattr(L, name), f"{name} declared in include/ but not exported by libuvol_b200.so"


def test_no_cpu_fallback(uv):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(uv.UvolError):
        uv.Context(0)
    h = ctypes.c_void_p()
    assert uv._native.lib().uvol_create(0, ctypes.byref(h)) == -4 and not h


@pytest.mark.parametrize("name", ["00000.drc", "00137.drc"])
def test_draco_core_logic_matches_oracle(built, name):
    blob = read(os.path.join(GOLDEN, "liam", name))
    e, o = emu_draco(blob), oracle_draco(blob)
    assert e["status"] == 0 and e["num_points"] == o["num_points"]
    assert np.array_equal(e["index"], o["index"])
    for k in ("position", "normal", "uv"):
        assert np.array_equal(e[k].view(np.uint32), o[k].view(np.uint32)), k


def test_draco_tagged_scheme_core_logic(built):
    """TAGGED symbol scheme: parser (tag run walked to its terminal state to locate the bit fields), tag decode and bit-field
    extraction as the kernels do them, against the oracle; truncated files are errors."""
    from tools.synth import synth
    rings, segs = synth.sphere_dims(2000); fp, fu, uvs, _ = synth.sphere_topology(rings, segs)
    pos = synth.sphere_frame(rings, segs, 0.2, 11); nrm = synth.vertex_normals(pos, fp)
    for mask, qp in ((1, 11), (2, 11), (4, 11), (7, 14)):
        blob = synth.encode_draco(pos, fp, uvs, fu, nrm, qp=qp, tagged=mask)
        e, o = emu_draco(blob), oracle_draco(blob)
        assert e["status"] == 0 and o["status"] == 0 and np.array_equal(e["index"], o["index"])
        for k in ("position", "normal", "uv"):
            assert np.array_equal(e[k].view(np.uint32), o[k].view(np.uint32)), (mask, k)
        for cut in (len(blob) // 3, len(blob) // 2, len(blob) - 3):
            assert emu_draco(blob[:cut])["status"] < 0


def test_speculative_traversal_row_pattern(built):
    """k_traverse's scheme, emulated lane by lane (tests/tools/draco_emu.cpp), must reproduce the serial depth-first order -- the
    emulation returns an error otherwise -- and on a UV sphere the seam-cut attribute table, which the walk crosses sideways (two faces
    per ring, then a jump of one ring), must be served by the row pattern: well above the two faces per step of the plain guess."""
    import ctypes as _c
    from emu_bind import TOOLS
    from tools.synth import synth
    drc = synth.make_sequence(1, 20000, 32, want_textures=False, seed=20260003)[0][0]
    e, o = emu_draco(drc), oracle_draco(drc)
    assert e["status"] == 0 and np.array_equal(e["index"], o["index"])
    for k in ("position", "normal", "uv"):
        assert np.array_equal(e[k].view(np.uint32), o[k].view(np.uint32)), k
    steps = (_c.c_long * 5)(); faces = _c.c_long()
    _c.CDLL(os.path.join(TOOLS, "libdraco_emu.so")).draco_emu_spec(1, steps, _c.byref(faces))
    per_step = [faces.value / s for s in list(steps)[:3] if s]
    assert len(per_step) == 3 and min(per_step) > 8.0, per_step


def test_draco_metadata_is_skipped(built):
    """A file with the metadata flag (header bit 15) decodes to the same mesh: the section is walked past by the parser and by the
    oracle (the reference's loader reads no metadata, DRACOLoader.js:470-554); a truncated metadata section is an error."""
    from conftest import with_draco_metadata
    blob = read(os.path.join(GOLDEN, "liam", "00000.drc")); mb = with_draco_metadata(blob)
    e, o, o0 = emu_draco(mb), oracle_draco(mb), oracle_draco(blob)
    assert e["status"] == 0 and o["status"] == 0 and np.array_equal(o["index"], o0["index"]) and np.array_equal(e["index"], o0["index"])
    for k in ("position", "normal", "uv"):
        assert np.array_equal(e[k].view(np.uint32), o0[k].view(np.uint32)), k
    assert emu_draco(mb[:40])["status"] < 0 and oracle_draco(mb[:40])["status"] < 0


def test_basis_core_logic_matches_oracle(built):
    blob = read(os.path.join(GOLDEN, "liam", "00000.ktx2"))
    e, o = emu_ktx2(blob), oracle_ktx2(blob)
    assert e["status"] == 0 and np.array_equal(e["rgba"], o["rgba"])


def test_core_logic_rejects_malformed(built):
    blob = bytearray(read(os.path.join(GOLDEN, "liam", "00000.drc")))
    assert emu_draco(bytes(blob[:5000]))["status"] < 0
    blob[0] = 0
    assert emu_draco(bytes(blob))["status"] < 0
    k = bytearray(read(os.path.join(GOLDEN, "liam", "00000.ktx2")))
    k[3] = 0
    assert emu_ktx2(bytes(k))["status"] < 0

"""KTX2 mip chains (levelCount > 1): the level / layer loop of the reference's transcode worker (src/lib/KTX2Loader.js:514-573).

UVOL's own encoder writes no mips (scripts/Encoder.py:290) and no mip-mapped fixture exists in the reference tree, so the inputs are
BUILT here from single-level files whose decode is already pinned:
  * UASTC: three synthetic single-level files (64^2, 32^2, 16^2; two layers each) merged into ONE three-level KTX2 -- plain and with
    every level Zstandard-compressed; level k of the result must be exactly the decode of file k;
  * ETC1S: a 4x4 texture with three levels (4x4, 2x2, 1x1) -- every level is one 4x4 block, so a three-LAYER 4x4 file (one shared
    codebook pair, what a mip chain needs) is re-labelled as three LEVELS; level k must be the top-left corner of layer k.
CPU part: the product's splitter (csrc/basis_parse.cpp uvol_ktx2_split_levels) + the host emulation / the oracle on its single-level
outputs.  GPU part: the same files through uvol_transcode_ktx2_batch (`mipmaps` of the result), mixed with single-level files.
"""
import struct
import sys

import numpy as np
import pytest

from conftest import ROOT
from emu_bind import emu_ktx2, emu_ktx2_split_levels
from oracle_bind import oracle_ktx2

sys.path.insert(0, ROOT)
from tools.synth import synth  # noqa: E402


def merge_uastc_levels(files, zstd_level=None):
    """Single-level UASTC files (sizes halving) -> one KTX2 with len(files) levels, small levels first in the file like KTX2 writers do."""
    base = files[0]
    dfd_off, dfd_len, kvd_off, kvd_len = struct.unpack_from("<4I", base, 48)
    payload = []
    for f in files:
        off, ln, _ = struct.unpack_from("<3Q", f, 80)
        payload.append(f[off:off + ln])
    raw_len = [len(p) for p in payload]
    if zstd_level is not None:
        from test_zstd import compress
        payload = [compress(p, zstd_level) for p in payload]
    L = len(files)
    o_dfd = 80 + 24 * L; o_kvd = (o_dfd + dfd_len + 3) & ~3; cur = (o_kvd + kvd_len + 15) & ~15
    offs = [0] * L
    for k in reversed(range(L)):
        offs[k] = cur; cur = (cur + len(payload[k]) + 15) & ~15
    out = bytearray(cur)
    out[:80] = base[:80]
    struct.pack_into("<I", out, 40, L)
    struct.pack_into("<I", out, 44, 0 if zstd_level is None else 2)
    struct.pack_into("<4I", out, 48, o_dfd, dfd_len, o_kvd if kvd_len else 0, kvd_len)
    struct.pack_into("<2Q", out, 64, 0, 0)
    for k in range(L):
        struct.pack_into("<3Q", out, 80 + 24 * k, offs[k], len(payload[k]), raw_len[k])
        out[offs[k]:offs[k] + len(payload[k])] = payload[k]
    out[o_dfd:o_dfd + dfd_len] = base[dfd_off:dfd_off + dfd_len]
    out[o_kvd:o_kvd + kvd_len] = base[kvd_off:kvd_off + kvd_len]
    return bytes(out)


def uastc_chain(zstd_level=None):
    singles = [synth.encode_uastc(synth.texture_layers(64 >> k, 0, 2, 30 + k), mode_mask=synth.UASTC_ALL_MODES, seed=30 + k) for k in range(3)]
    return merge_uastc_levels(singles, zstd_level), [oracle_ktx2(s)["rgba"] for s in singles]


def etc1s_chain():
    """Three distinct 4x4 layers sharing one codebook pair, re-labelled as the three levels of a 4x4 texture."""
    rng = np.random.default_rng(11)
    layers = rng.integers(0, 256, (3, 4, 4, 4), dtype=np.uint8); layers[..., 3] = 255
    f = bytearray(synth.encode_etc1s(layers))
    expect = oracle_ktx2(bytes(f))["rgba"]
    nl = struct.unpack_from("<I", f, 32)[0]; assert nl == 3
    sgd_off, sgd_len = struct.unpack_from("<2Q", f, 64); lv_off, lv_len, _ = struct.unpack_from("<3Q", f, 80)
    descs = [struct.unpack_from("<5I", f, sgd_off + 20 + 20 * k) for k in range(3)]
    # the level index grows from one entry to three: everything behind it moves by 48 bytes
    grow = 48
    out = bytearray(f[:104]) + bytearray(grow) + f[104:]
    struct.pack_into("<I", out, 32, 0); struct.pack_into("<I", out, 40, 3)
    dfd_off, dfd_len, kvd_off, kvd_len = struct.unpack_from("<4I", f, 48)
    struct.pack_into("<4I", out, 48, dfd_off + grow, dfd_len, kvd_off + grow if kvd_len else 0, kvd_len)
    struct.pack_into("<2Q", out, 64, sgd_off + grow, sgd_len)
    for k, (flags, ro, rl, ao, al) in enumerate(descs):
        struct.pack_into("<3Q", out, 80 + 24 * k, lv_off + grow + ro, rl, 0)          # each level = that layer's slice
        struct.pack_into("<5I", out, sgd_off + grow + 20 + 20 * k, flags, 0, rl, 0, 0)
    return bytes(out), [expect[k:k + 1, :max(1, 4 >> k), :max(1, 4 >> k)] for k in range(3)]


def test_split_levels_host_logic(built):
    for zl in (None, 3):
        try:
            blob, expect = uastc_chain(zl)
        except ImportError:
            continue
        rc, files = emu_ktx2_split_levels(blob)
        assert rc == 3 and len(files) == 3
        for k, f in enumerate(files):
            assert struct.unpack_from("<3I", f, 20)[:2] == (64 >> k, 64 >> k) and struct.unpack_from("<I", f, 40)[0] == 1
            e = emu_ktx2(f)
            assert e["status"] == 0 and np.array_equal(e["rgba"], expect[k])
            if zl is None:          # (the oracle reads no Zstandard levels; the product's inflater is pinned to libzstd in test_zstd.py)
                o = oracle_ktx2(f)
                assert o["status"] == 0 and np.array_equal(o["rgba"], expect[k])
    # a full chain of a ragged texture: 52x38 -> 26x19 -> 13x9 -> 6x4 -> 3x2 -> 1x1 (every level max(1, base >> k), partial blocks at every size)
    dims = [(max(1, 52 >> k), max(1, 38 >> k)) for k in range(6)]
    singles = [synth.encode_uastc(synth.texture_layers(64, 0, 1, 50 + k)[:, :h, :w], mode_mask=synth.UASTC_ALL_MODES, seed=50 + k) for k, (w, h) in enumerate(dims)]
    rc, files = emu_ktx2_split_levels(merge_uastc_levels(singles))
    assert rc == 6
    for k, f in enumerate(files):
        e, o = emu_ktx2(f), oracle_ktx2(singles[k])
        assert e["status"] == 0 and e["rgba"].shape[1:3] == (dims[k][1], dims[k][0]) and np.array_equal(e["rgba"], o["rgba"]), k
    blob, expect = etc1s_chain()
    rc, files = emu_ktx2_split_levels(blob)
    assert rc == 3
    for k, f in enumerate(files):
        e, o = emu_ktx2(f), oracle_ktx2(f)
        assert e["status"] == 0 and o["status"] == 0, (k, e["status"], o["status"])
        assert np.array_equal(e["rgba"], expect[k]) and np.array_equal(o["rgba"], expect[k])
    # single-level files and non-KTX2 bytes are passed on untouched; a chain that runs off the file is an error
    single = synth.encode_uastc(synth.texture_layers(16, 0, 1, 3), seed=9)
    assert emu_ktx2_split_levels(single)[0] == 0 and emu_ktx2_split_levels(b"x" * 200)[0] == 0
    blob, _ = uastc_chain()
    cut = bytearray(blob); struct.pack_into("<Q", cut, 88, 1 << 40)
    assert emu_ktx2_split_levels(bytes(cut))[0] < 0
    deep = bytearray(blob); struct.pack_into("<I", deep, 40, 9)          # 64 >> 8 == 0 in both dimensions
    assert emu_ktx2_split_levels(bytes(deep))[0] < 0


@pytest.mark.gpu
def test_mip_chains_on_the_gpu(uv, ctx):
    """One batch: a plain UASTC chain, a Zstandard one, a single-level file, the ETC1S chain, a broken chain.  Every level must be
    bit-exact; the ASTC target follows the same path (level sizes in blocks)."""
    u, ue = uastc_chain(); e1, ee = etc1s_chain()
    single = synth.encode_uastc(synth.texture_layers(32, 0, 2, 4), seed=5)
    broken = bytearray(u); struct.pack_into("<Q", broken, 88 + 24, 1 << 40)
    blobs = [u, single, e1, bytes(broken)]
    try:
        z, ze = uastc_chain(3); blobs.append(z)
    except ImportError:
        z = None
    res = uv.KTX2Loader(ctx).transcode_batch(blobs)
    assert [r["status"] for r in res[:4]] == [0, 0, 0, -1]
    for r, exp in ((res[0], ue), (res[2], ee)) + (((res[4], ze),) if z else ()):
        assert r["status"] == 0 and len(r["mipmaps"]) == 3
        for k in range(3):
            assert (r["mipmaps"][k]["width"], r["mipmaps"][k]["height"]) == (exp[k].shape[2], exp[k].shape[1])
            assert np.array_equal(r["mipmaps"][k]["data"], exp[k]), k
    assert len(res[1]["mipmaps"]) == 1 and np.array_equal(res[1]["data"], oracle_ktx2(single)["rgba"])
    a = uv.KTX2Loader(ctx).transcode_batch([u], target=uv.TEX_ASTC_4x4)[0]
    from oracle_bind import oracle_astc_image
    for k in range(3):
        m = a["mipmaps"][k]
        for L in range(2):
            img, bad = oracle_astc_image(m["data"][L], m["width"], m["height"])
            assert bad == 0 and np.array_equal(img, ue[k][L])


def test_ktx2_probe_reports_what_the_transcode_would_see(built):
    """uvol_ktx2_probe = the KTX2File getters the reference reads before choosing a target (src/lib/KTX2Loader.js:471-495)."""
    import ctypes
    import importlib
    from conftest import golden_ktx2, read
    from test_etc2 import etc1s_with_alpha
    uvp = importlib.import_module("universal-volumetric_b200"); L = uvp._native.lib()

    class Info(ctypes.Structure):
        _fields_ = [(k, ctypes.c_uint32) for k in ("width", "height", "layers", "levels", "faces", "is_uastc", "has_alpha", "is_video", "supercompression", "dfd_transfer", "dfd_flags")]

    def probe(blob):
        i = Info(); rc = L.uvol_ktx2_probe(blob, ctypes.c_size_t(len(blob)), ctypes.byref(i))
        return rc, i
    rc, i = probe(read(golden_ktx2()[0])); o = oracle_ktx2(read(golden_ktx2()[0]))
    assert rc == 0 and (i.width, i.height, i.layers, i.levels, i.is_uastc, i.has_alpha, i.is_video, i.supercompression) == (o["width"], o["height"], o["layers"], 1, 0, 0, 1, 1)
    rc, i = probe(uastc_chain()[0])
    assert rc == 0 and (i.width, i.height, i.layers, i.levels, i.is_uastc) == (64, 64, 2, 3, 1)
    rc, i = probe(etc1s_with_alpha()[0])
    assert rc == 0 and (i.layers, i.has_alpha, i.is_uastc) == (2, 1, 0)
    rc, i = probe(synth.encode_uastc(synth.texture_layers(52, 0, 1, 9)[:, :38, :], mode_mask=synth.UASTC_ALL_MODES, seed=11, has_alpha=True))
    assert rc == 0 and (i.width, i.height, i.has_alpha, i.is_uastc, i.supercompression) == (52, 38, 1, 1, 0)
    assert probe(b"x" * 300)[0] == -2 and probe(read(golden_ktx2()[0])[:90])[0] < 0
    rc, d = uvp.ktx2_probe(uastc_chain()[0])          # the Python mirror of both calls
    assert rc == 0 and d["levels"] == 3 and uvp.pick_texture_format(d["is_uastc"], d["has_alpha"], astcSupported=True, bptcSupported=True) == uvp.TEX_ASTC_4x4
    assert uvp.pick_texture_format(False, True, etc2Supported=True, dxtSupported=True) == uvp.TEX_ETC2_RGBA and uvp.ktx2_probe(b"nope")[0] < 0
    # the chooser on the probe's answer: desktop NVIDIA (bptc + s3tc) -> BC7; a context with ASTC -> lossless ASTC for the UASTC chain
    assert L.uvol_pick_texture_format(0, 0, 2 | 4) == uvp._native.TEX_BC7 and L.uvol_pick_texture_format(1, 0, 1 | 2) == uvp._native.TEX_ASTC_4x4

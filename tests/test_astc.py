"""ASTC 4x4 target format (UVOL_TEX_ASTC_4x4; the reference's first choice for UASTC sources on GPUs with ASTC support,
src/lib/KTX2Loader.js:592-600).

UASTC is a subset of ASTC, so the repack is LOSSLESS and the bar is bit-exactness: the product's ASTC blocks, decoded by the oracle's
INDEPENDENT ASTC decoder (oracle/astc_decode.c: written from the ASTC specification, derives the endpoint range from the bit budget,
unbundles trits / quints by the spec's bit equations, applies blue contraction where the format says so), must give exactly the
texels of the oracle's UASTC -> RGBA32 decode of the same file -- for every mode, for blocks of random bits (which exercise the
endpoint-swap that keeps ASTC's blue contraction off, something an encoder's ordered endpoints never do), ragged sizes, solid blocks.
CPU part: the per-block function through the host emulation.  GPU part: the kernel must emit exactly those bytes.
ETC1S sources are not offered this target by the reference (priorityETC1S: Infinity): they report UNSUPPORTED per item.
"""
import ctypes
import os
import struct
import sys

import numpy as np
import pytest

from conftest import ROOT
from emu_bind import emu_ktx2_astc
from oracle_bind import lib as olib, oracle_astc_image, oracle_ktx2

sys.path.insert(0, ROOT)
from tools.synth import synth  # noqa: E402


def random_blocks_file(n_side=128, seed=20260005):
    """A UASTC file whose blocks are random bits wherever the RGBA32 transcoder accepts them."""
    rng = np.random.default_rng(seed)
    blob = bytearray(synth.encode_uastc(synth.texture_layers(n_side * 4, 0, 1, 5), mode_mask=synth.UASTC_ALL_MODES, seed=3))
    lv = struct.unpack_from("<Q", blob, 80)[0]
    L = olib(); px = (ctypes.c_uint8 * 64)(); rnd = rng.integers(0, 256, (n_side * n_side, 16), dtype=np.uint8); kept = 0
    for i in range(len(rnd)):
        if L.uvo_uastc_block_to_rgba(rnd[i].ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), px) == 0:
            blob[lv + 16 * i: lv + 16 * i + 16] = rnd[i].tobytes(); kept += 1
    assert kept > n_side * n_side // 3
    return bytes(blob)


def files():
    fs = {"mode_%d" % m: synth.encode_uastc(synth.texture_layers(32, 0, 1, 40 + m), mode_mask=1 << m, seed=40 + m, has_alpha=m in (9, 10, 11, 12, 13, 14, 15, 16, 17))
          for m in range(19)}
    fs["all_modes"] = synth.encode_uastc(synth.texture_layers(64, 0, 2, 5), mode_mask=synth.UASTC_ALL_MODES, seed=5)
    fs["ragged"] = synth.encode_uastc(synth.texture_layers(52, 0, 1, 9)[:, :38, :], mode_mask=synth.UASTC_ALL_MODES, seed=11)
    fs["random_bits"] = random_blocks_file()
    return fs


def check_blocks(name, blob, blocks, width, height, layers):
    o = oracle_ktx2(blob)
    assert o["status"] == 0 and (o["width"], o["height"], o["layers"]) == (width, height, layers)
    for L in range(layers):
        img, bad = oracle_astc_image(blocks[L], width, height)
        assert bad == 0, f"{name}: {bad} blocks are not ASTC blocks the decoder accepts"
        assert np.array_equal(img, o["rgba"][L]), (name, int(np.abs(img.astype(int) - o["rgba"][L].astype(int)).max()))


def test_astc_blocks_host_logic(built):
    for name, blob in files().items():
        e = emu_ktx2_astc(blob)
        assert e["status"] == 0, name
        check_blocks(name, blob, e["blocks"], e["width"], e["height"], e["layers"])


def test_astc_swapped_endpoints(built):
    """A hand-made mode-1 block (RGB, one subset, 8-bit endpoints, 2-bit weights) whose first endpoint is the brighter one: ASTC would
    read that order as blue contraction, so the repack must store the pair swapped and the weights complemented -- same texels."""
    bits, pos = 0, 0

    def put(v, n):
        nonlocal bits, pos
        bits |= v << pos; pos += n
    put(0x35, 6); put(0, 15)                                    # mode 1, hints
    for lo, hi in ((200, 10), (180, 30), (220, 5)):             # R, G, B: {first, second} endpoint
        put(lo, 8); put(hi, 8)
    w = [1] + [(i * 7) & 3 for i in range(1, 16)]
    put(w[0], 1)                                                # the anchor texel stores one bit less
    for v in w[1:]:
        put(v, 2)
    blk = bits.to_bytes(16, "little")
    blob = bytearray(synth.encode_uastc(synth.texture_layers(4, 0, 1, 5), mode_mask=1 << 8, seed=3)); lv = struct.unpack_from("<Q", blob, 80)[0]
    blob[lv:lv + 16] = blk
    e = emu_ktx2_astc(bytes(blob)); o = oracle_ktx2(bytes(blob))
    assert e["status"] == 0 and o["status"] == 0
    out = int.from_bytes(e["blocks"][0][0].tobytes(), "little")
    assert out & 0x1FFFF == 0x42 | (8 << 13)                    # block mode 0x42, one partition, CEM 8
    assert [(out >> (17 + 8 * k)) & 255 for k in range(6)] == [10, 200, 30, 180, 5, 220]
    assert [(((out >> (127 - 2 * i)) & 1) | (((out >> (126 - 2 * i)) & 1) << 1)) for i in range(16)] == [3 - v for v in w]
    img, bad = oracle_astc_image(e["blocks"][0], 4, 4)
    assert bad == 0 and np.array_equal(img, o["rgba"][0])
    assert len(np.unique(img.reshape(-1, 4), axis=0)) == 4      # all four interpolants present: the check is not on a flat block


def test_astc_solid_and_etc1s(built):
    rng = np.random.default_rng(3)
    tiles = rng.integers(0, 256, (1, 16, 16, 4), dtype=np.uint8)
    img = np.repeat(np.repeat(tiles, 4, axis=1), 4, axis=2)
    blob = synth.encode_uastc(img, mode_mask=1 << 8, seed=1, has_alpha=True)
    e = emu_ktx2_astc(blob)
    dec, bad = oracle_astc_image(e["blocks"][0], 64, 64)
    assert bad == 0 and np.array_equal(dec, oracle_ktx2(blob)["rgba"][0])
    assert (e["blocks"][0][:, :8] == np.frombuffer(bytes([0xFC, 0xFD, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF]), np.uint8)).all()      # void-extent header
    assert emu_ktx2_astc(synth.encode_etc1s(synth.texture_layers(64, 0, 1, 4)))["status"] == -3


@pytest.mark.gpu
def test_astc_kernel_matches_host_logic(uv, ctx):
    """The CUDA kernel emits byte for byte what the per-block function emits on the host (validated above by the independent decoder),
    incl. the 2048^2 x 7 bench size; an ETC1S file in the same batch fails alone."""
    fs = files()
    fs["uastc_2048x7"] = synth.encode_uastc(synth.texture_layers(2048, 0, 7, 3), mode_mask=synth.UASTC_OPAQUE_MODES, seed=20260003)
    etc1s = synth.encode_etc1s(synth.texture_layers(64, 0, 2, 4))
    names = list(fs)
    res = uv.KTX2Loader(ctx).transcode_batch([fs[n] for n in names] + [etc1s], target=uv.TEX_ASTC_4x4)
    assert res[-1]["status"] == -3
    for n, r in zip(names, res):
        e = emu_ktx2_astc(fs[n])
        assert r["status"] == 0 and r["format"] == "RGBA_ASTC_4x4_Format" and r["data"].shape == e["blocks"].shape, n
        assert np.array_equal(r["data"], e["blocks"]), n
        if n != "uastc_2048x7":
            check_blocks(n, fs[n], r["data"], r["width"], r["height"], r["layers"])

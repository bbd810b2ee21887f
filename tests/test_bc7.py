"""BC7 target format (UVOL_TEX_BC7; the reference's choice on desktop GPUs, src/lib/KTX2Loader.js:602-604).

The reference's own BC7 arithmetic is in the absent basis_transcoder WASM and cannot be bit-matched, so the product's
blocks are checked by an INDEPENDENT BC7 decoder (oracle/bc7_decode.c, all eight modes, written from the BPTC layout) against the
oracle's RGBA32 decode of the same file:
  * every emitted block is a well-formed BC7 block (mode bit present, exactly 128 bits consumed);
  * solid-colour blocks (UASTC mode 8, single-selector ETC1S blocks) decode EXACTLY;
  * UASTC sources: every texel within 6 / 255 of the RGBA32 decode (the 4-bit ASTC and BC7 weight grids differ by up to 1 / 64,
    endpoints lose at most one bit, mode 18's 5-bit weights fold to 4), PSNR >= 45 dB;
  * ETC1S sources: BC7 mode 5 spaces its four colours evenly while ETC1S's intensity tables do not, so the bound is a PSNR one:
    >= 40 dB on the reference's own fixture, >= 38 dB on the synthetic textures;
  * alpha is 255 wherever the RGBA32 decode has 255.
CPU part: the product's per-block functions run through the host emulation.  GPU part: the kernels must emit exactly those bytes.
"""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, golden_ktx2, read
from emu_bind import emu_ktx2_bc7
from oracle_bind import oracle_bc7_image, oracle_ktx2

sys.path.insert(0, ROOT)
from tools.synth import synth  # noqa: E402


def psnr(a, b):
    d = a.astype(np.float64) - b.astype(np.float64)
    mse = (d ** 2).mean()
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


def files():
    return {
        "uastc_all_modes": synth.encode_uastc(synth.texture_layers(64, 0, 2, 5), mode_mask=synth.UASTC_ALL_MODES, seed=5),
        "uastc_opaque": synth.encode_uastc(synth.texture_layers(128, 0, 1, 5), mode_mask=synth.UASTC_OPAQUE_MODES, seed=7),
        "uastc_ragged": synth.encode_uastc(synth.texture_layers(52, 0, 1, 9)[:, :38, :], mode_mask=synth.UASTC_ALL_MODES, seed=11),
        "etc1s_synth": synth.encode_etc1s(synth.texture_layers(64, 0, 3, 4)),
        "etc1s_liam": read(golden_ktx2()[0]),
    }


def check_blocks(name, blob, blocks, width, height, layers):
    o = oracle_ktx2(blob)
    assert o["status"] == 0 and (o["width"], o["height"], o["layers"]) == (width, height, layers)
    for L in range(layers):
        img, bad = oracle_bc7_image(blocks[L], width, height)
        assert bad == 0, f"{name}: malformed BC7 blocks"
        ref = o["rgba"][L]
        d = np.abs(img.astype(np.int32) - ref.astype(np.int32))
        assert (img[..., 3][ref[..., 3] == 255] == 255).all(), f"{name}: opaque texels must stay opaque"
        if name.startswith("uastc"):
            assert d.max() <= 6 and psnr(img, ref) >= 45.0, (name, int(d.max()), psnr(img, ref))
        else:
            assert psnr(img[..., :3], ref[..., :3]) >= (40.0 if "liam" in name else 38.0), (name, psnr(img[..., :3], ref[..., :3]))
        # solid 4x4 blocks of the RGBA32 decode (whole blocks only) must be exact
        bx, by = width // 4, height // 4
        rb = ref[: by * 4, : bx * 4].reshape(by, 4, bx, 4, 4).transpose(0, 2, 1, 3, 4).reshape(by, bx, 16, 4)
        ib = img[: by * 4, : bx * 4].reshape(by, 4, bx, 4, 4).transpose(0, 2, 1, 3, 4).reshape(by, bx, 16, 4)
        solid = (rb == rb[:, :, :1]).all(axis=(2, 3))
        assert (ib[solid] == rb[solid]).all(), f"{name}: solid blocks must decode exactly"


@pytest.mark.parametrize("name", ["uastc_all_modes", "uastc_opaque", "uastc_ragged", "etc1s_synth", "etc1s_liam"])
def test_bc7_blocks_host_logic(built, name):
    blob = files()[name]
    e = emu_ktx2_bc7(blob)
    assert e["status"] == 0
    check_blocks(name, blob, e["blocks"], e["width"], e["height"], e["layers"])


def test_bc7_solid_blocks_exact(built):
    """A texture of flat 4x4 tiles: every block is solid -> BC7 mode 5 reproduces every 8-bit value exactly, alpha included."""
    rng = np.random.default_rng(3)
    tiles = rng.integers(0, 256, (1, 16, 16, 4), dtype=np.uint8)
    img = np.repeat(np.repeat(tiles, 4, axis=1), 4, axis=2)
    blob = synth.encode_uastc(img, mode_mask=1 << 8, seed=1)
    e = emu_ktx2_bc7(blob); o = oracle_ktx2(blob)
    dec, bad = oracle_bc7_image(e["blocks"][0], 64, 64)
    assert bad == 0 and np.array_equal(dec, o["rgba"][0])


@pytest.mark.gpu
def test_bc7_kernels_match_host_logic(uv, ctx):
    """The CUDA kernels emit byte for byte what the per-block functions emit on the host (which the CPU tests above validate with the
    independent decoder), for UASTC and ETC1S sources, ragged sizes, and the 2048^2 x 7 bench size."""
    fs = files()
    fs["uastc_2048x7"] = synth.encode_uastc(synth.texture_layers(2048, 0, 7, 3), mode_mask=synth.UASTC_OPAQUE_MODES, seed=20260003)
    names = list(fs)
    res = uv.KTX2Loader(ctx).transcode_batch([fs[n] for n in names], target=uv.TEX_BC7)
    for n, r in zip(names, res):
        e = emu_ktx2_bc7(fs[n])
        assert r["status"] == 0 and r["format"] == "RGBA_BPTC_Format" and r["data"].shape == e["blocks"].shape, n
        assert np.array_equal(r["data"], e["blocks"]), n
        if n != "uastc_2048x7":
            check_blocks(n, fs[n], r["data"], r["width"], r["height"], r["layers"])
    # the combined V2 entry point honours the context's configured target (uvol_config.texture_target)
    c2 = uv.Context(0, texture_target=uv.TEX_BC7)
    try:
        g, t = uv.V2Player(c2).decode_step_raw([], [fs["etc1s_synth"]], uv.MEM_HOST)
        e = emu_ktx2_bc7(fs["etc1s_synth"])
        assert t[0].status == 0 and t[0].format == uv.TEX_BC7 and t[0].bytes == e["blocks"].size
        assert np.array_equal(np.ctypeslib.as_array(t[0].data, e["blocks"].shape), e["blocks"])
    finally:
        c2.close()


def _dds_bc7(blocks, w, h):
    """A DDS file (DX10 header, DXGI_FORMAT_BC7_UNORM = 98) around raw BC7 blocks: what Pillow's DdsImagePlugin reads."""
    import struct
    hdr = b"DDS " + struct.pack("<7I", 124, 0x1 | 0x2 | 0x4 | 0x1000 | 0x80000, h, w, len(blocks), 0, 1) + struct.pack("<11I", *([0] * 11))
    pf = struct.pack("<II4sIIIII", 32, 0x4, b"DX10", 0, 0, 0, 0, 0)
    return hdr + pf + struct.pack("<IIIII", 0x1000, 0, 0, 0, 0) + struct.pack("<IIIII", 98, 3, 0, 1, 0) + blocks


def test_bc7_decoder_pinned_to_pillow(built):
    """THIRD-PARTY pin of the BC7 side: Pillow's BC7 decoder (its own C implementation, DdsImagePlugin / BcnDecoder) must return exactly
    the texels of oracle/bc7_decode.c -- on the product's blocks (UASTC and ETC1S sources) and on 4096 blocks of random bits per BC7 mode
    (every mode, partition, rotation and index-selection combination).  So the BC7 bounds above are measured with a decoder that an
    independent implementation agrees with, not only with our own reading of the BPTC layout."""
    import io
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(20260006)
    cases = []
    for name in ("uastc_all_modes", "etc1s_synth"):
        e = emu_ktx2_bc7(files()[name])
        cases.append((name, e["blocks"][0], e["width"], e["height"]))
    for mode in range(8):
        rnd = rng.integers(0, 256, (64 * 64, 16), dtype=np.uint8)
        rnd[:, 0] = (rnd[:, 0] & np.uint8(0xFF & ~((2 << mode) - 1))) | np.uint8(1 << mode)      # the unary mode prefix: `mode` zero bits, then a one
        cases.append(("random_mode_%d" % mode, rnd, 256, 256))
    for name, blocks, w, h in cases:
        try:
            img = np.array(Image.open(io.BytesIO(_dds_bc7(np.ascontiguousarray(blocks).tobytes(), w, h))).convert("RGBA"))
        except Exception as ex:          # a Pillow build without the BC7 path
            pytest.skip("Pillow cannot decode BC7 DDS here: %s" % ex)
        ours, bad = oracle_bc7_image(blocks, w, h)
        assert bad == 0, name
        assert np.array_equal(img, ours), (name, int((img != ours).sum()))

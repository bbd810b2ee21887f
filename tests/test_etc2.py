"""ETC2 RGBA target (UVOL_TEX_ETC2_RGBA; the reference's `etc2Supported` option, transcoderFormat [ETC1, ETC2] -> RGB_ETC2 / RGBA_ETC2_EAC,
src/lib/KTX2Loader.js:619-627 -- the top-priority option for ETC1S sources) and the ETC1S ALPHA path as a whole.

Inputs with an alpha slice: the synthetic encoder writes opaque files only (like the reference's own content), so a file with alpha is
made by RE-LABELLING an opaque file of 2L layers as L layers whose alpha slices are the slices of layers L..2L-1 (an alpha slice is an
ordinary ETC1S slice over the same codebooks; its G channel is the alpha) -- see etc1s_with_alpha().
Checks (blocks decoded by the independent numpy ETC1 / EAC decoder in tests/etc1_decode.py):
  * RGBA32 decode of such a file: RGB = the first L layers of the opaque file, A = the G channel of the last L (oracle, host emulation, GPU);
  * ETC2 RGBA blocks: the colour half decodes to EXACTLY the RGB of the RGBA32 result (an ETC1S block is an ETC1 block), the EAC half is a
    fit (EAC steps are multiples of its multiplier; the base may clamp at 0 / 255): every alpha texel within 12 / 255 and alpha PSNR >= 42 dB
    (measured: 10 and 44.6 dB on the test texture); opaque files get alpha 255 exactly;
  * UASTC sources report UNSUPPORTED per item (they would need an ETC1 encoder).
"""
import struct
import sys

import numpy as np
import pytest

from conftest import ROOT
from emu_bind import emu_ktx2, emu_ktx2_bc7, emu_ktx2_etc2a
from etc1_decode import decode_etc2_rgba
from oracle_bind import oracle_bc7_image, oracle_ktx2

sys.path.insert(0, ROOT)
from tools.synth import synth  # noqa: E402


def etc1s_with_alpha(size=64, layers=2, seed=4):
    """(.ktx2 with alpha, expected RGBA u8[layers, size, size, 4])."""
    img = synth.texture_layers(size, 0, 2 * layers, seed)
    img[layers:] = img[layers:][:, ::-1, ::-1]                      # the alpha source looks nothing like the colour layers (no inter-layer copies)
    f = synth.encode_etc1s(img)
    o = oracle_ktx2(f)["rgba"]
    expect = o[:layers].copy(); expect[..., 3] = o[layers:, ..., 1]
    dfd_off, dfd_len, kvd_off, kvd_len = struct.unpack_from("<4I", f, 48)
    sgd_off, sgd_len = struct.unpack_from("<2Q", f, 64); lv = struct.unpack_from("<3Q", f, 80)
    assert struct.unpack_from("<I", f, 32)[0] == 2 * layers and dfd_len == 44
    # DFD: a second sample (alpha, channel id 15), descriptor block 24 + 2 * 16 bytes
    dfd = bytearray(f[dfd_off:dfd_off + dfd_len]) + bytearray(f[dfd_off + 28:dfd_off + 44])
    struct.pack_into("<I", dfd, 0, 60); struct.pack_into("<H", dfd, 10, 56); dfd[44 + 3] = (dfd[44 + 3] & 0xF0) | 15
    g = f[sgd_off:sgd_off + sgd_len]
    descs = [struct.unpack_from("<5I", g, 20 + 20 * k) for k in range(2 * layers)]
    sgd = bytearray(g[:20])
    for k in range(layers):
        sgd += struct.pack("<5I", descs[k][0], descs[k][1], descs[k][2], descs[layers + k][1], descs[layers + k][2])
    sgd += g[20 + 40 * layers:]
    kvd = f[kvd_off:kvd_off + kvd_len]
    o_dfd = 104; o_kvd = (o_dfd + len(dfd) + 3) & ~3; o_sgd = (o_kvd + len(kvd) + 7) & ~7; o_lv = (o_sgd + len(sgd) + 15) & ~15
    out = bytearray(o_lv + lv[1])
    out[:80] = f[:80]
    struct.pack_into("<I", out, 32, layers)
    struct.pack_into("<4I", out, 48, o_dfd, len(dfd), o_kvd, len(kvd)); struct.pack_into("<2Q", out, 64, o_sgd, len(sgd)); struct.pack_into("<3Q", out, 80, o_lv, lv[1], lv[2])
    out[o_dfd:o_dfd + len(dfd)] = dfd; out[o_kvd:o_kvd + len(kvd)] = kvd; out[o_sgd:o_sgd + len(sgd)] = sgd; out[o_lv:] = f[lv[0]:lv[0] + lv[1]]
    return bytes(out), expect


def psnr(a, b):
    d = a.astype(np.float64) - b.astype(np.float64); mse = (d ** 2).mean()
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


def check_etc2(name, blocks, ref):
    for L in range(ref.shape[0]):
        img = decode_etc2_rgba(blocks[L], ref.shape[2], ref.shape[1])
        assert np.array_equal(img[..., :3], ref[L][..., :3]), f"{name}: the colour half must decode exactly"
        d = np.abs(img[..., 3].astype(int) - ref[L][..., 3].astype(int))
        if (ref[L][..., 3] == 255).all():
            assert d.max() == 0, f"{name}: opaque alpha must stay 255"
        else:
            assert d.max() <= 12 and psnr(img[..., 3], ref[L][..., 3]) >= 42.0, (name, int(d.max()), psnr(img[..., 3], ref[L][..., 3]))


def test_etc1s_alpha_slices_host_logic(built):
    blob, expect = etc1s_with_alpha()
    o, e = oracle_ktx2(blob), emu_ktx2(blob)
    assert o["status"] == 0 and o["has_alpha"] and e["status"] == 0
    assert np.array_equal(o["rgba"], expect) and np.array_equal(e["rgba"], expect)
    b7 = emu_ktx2_bc7(blob)                                              # the BC7 target's alpha path (mode 5 scalar channel) on a real alpha slice
    for L in range(expect.shape[0]):
        img, bad = oracle_bc7_image(b7["blocks"][L], expect.shape[2], expect.shape[1])
        assert bad == 0 and psnr(img[..., 3], expect[L][..., 3]) >= 38.0 and psnr(img[..., :3], expect[L][..., :3]) >= 38.0


def test_etc2_rgba_blocks_host_logic(built):
    blob, expect = etc1s_with_alpha()
    e = emu_ktx2_etc2a(blob)
    assert e["status"] == 0 and e["blocks"].shape == (2, 256, 16)
    check_etc2("alpha", e["blocks"], expect)
    opaque = synth.encode_etc1s(synth.texture_layers(64, 0, 3, 4))
    e = emu_ktx2_etc2a(opaque)
    check_etc2("opaque", e["blocks"], oracle_ktx2(opaque)["rgba"])
    assert emu_ktx2_etc2a(synth.encode_uastc(synth.texture_layers(16, 0, 1, 3), seed=9))["status"] == -3


@pytest.mark.gpu
def test_etc2_rgba_and_alpha_on_the_gpu(uv, ctx):
    """RGBA32 of an ETC1S file with alpha slices == the oracle; the ETC2 RGBA kernel emits byte for byte what the per-block functions emit
    on the host (validated above); a UASTC file in the same batch fails alone."""
    blob, expect = etc1s_with_alpha()
    big, big_expect = etc1s_with_alpha(256, 3, 9)
    opaque = synth.encode_etc1s(synth.texture_layers(64, 0, 3, 4))
    uastc = synth.encode_uastc(synth.texture_layers(16, 0, 1, 3), seed=9)
    r = uv.KTX2Loader(ctx).transcode_batch([blob, big])
    assert r[0]["status"] == 0 and r[0]["hasAlpha"] and np.array_equal(r[0]["data"], expect) and np.array_equal(r[1]["data"], big_expect)
    res = uv.KTX2Loader(ctx).transcode_batch([blob, opaque, uastc, big], target=uv.TEX_ETC2_RGBA)
    assert [x["status"] for x in res] == [0, 0, -3, 0] and res[0]["format"] == "RGBA_ETC2_EAC_Format"
    for x, src in ((res[0], blob), (res[1], opaque), (res[3], big)):
        assert np.array_equal(x["data"], emu_ktx2_etc2a(src)["blocks"])
    check_etc2("alpha_gpu", res[3]["data"], big_expect)
    b7 = uv.KTX2Loader(ctx).transcode_batch([blob], target=uv.TEX_BC7)[0]
    assert b7["status"] == 0 and np.array_equal(b7["data"], emu_ktx2_bc7(blob)["blocks"])

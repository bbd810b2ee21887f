"""BC1 / BC3 targets (UVOL_TEX_BC1 / UVOL_TEX_BC3; the reference's `dxtSupported` option, transcoderFormat [BC1, BC3] -> RGB_S3TC_DXT1 /
RGBA_S3TC_DXT5, src/lib/KTX2Loader.js:610-618: the fallback on desktop GPUs without BPTC), ETC1S sources.

The blocks are decoded by a THIRD-PARTY decoder -- Pillow's DXT1 / DXT5 (DdsImagePlugin) -- and compared with the oracle's RGBA32 decode
of the same file: the conversion is lossy by construction (RGB565 endpoints, thirds instead of ETC1S's intensity steps), so the bounds
are PSNR ones: colour >= 35 dB on the synthetic textures and on the reference's own fixture (measured 40.1 / 42.5 dB), BC3 alpha >= 38 dB
(measured 44.4) and
255 exactly for opaque files; blocks whose texels are all equal keep a single colour (no ringing).  UASTC sources report UNSUPPORTED
per item.  GPU part: the kernel must emit exactly the bytes of the per-block functions run on the host.
"""
import io
import struct
import sys

import numpy as np
import pytest

from conftest import ROOT, golden_ktx2, read
from emu_bind import emu_ktx2_dxt
from oracle_bind import oracle_ktx2
from test_etc2 import etc1s_with_alpha, psnr

sys.path.insert(0, ROOT)
from tools.synth import synth  # noqa: E402

Image = pytest.importorskip("PIL.Image")


def dds(blocks, w, h, fourcc):
    hdr = b"DDS " + struct.pack("<7I", 124, 0x1 | 0x2 | 0x4 | 0x1000 | 0x80000, h, w, len(blocks), 0, 1) + struct.pack("<11I", *([0] * 11))
    return hdr + struct.pack("<II4sIIIII", 32, 0x4, fourcc, 0, 0, 0, 0, 0) + struct.pack("<IIIII", 0x1000, 0, 0, 0, 0) + blocks


def pillow_decode(blocks, w, h, bc3):
    """blocks u8[nblocks, 8 | 16] (whole 4x4 blocks; w, h multiples of four here) -> u8[h, w, 4]"""
    return np.array(Image.open(io.BytesIO(dds(np.ascontiguousarray(blocks).tobytes(), w, h, b"DXT5" if bc3 else b"DXT1"))).convert("RGBA"))


def check(name, blob, blocks, bc3, ref):
    L, h, w = ref.shape[:3]
    for k in range(L):
        img = pillow_decode(blocks[k], w, h, bc3)
        p = psnr(img[..., :3], ref[k][..., :3])
        assert p >= 35.0, (name, k, p)
        if bc3:
            if (ref[k][..., 3] == 255).all():
                assert (img[..., 3] == 255).all(), name
            else:
                assert psnr(img[..., 3], ref[k][..., 3]) >= 38.0, (name, psnr(img[..., 3], ref[k][..., 3]))
        else:
            assert (img[..., 3] == 255).all(), f"{name}: BC1 blocks must stay in an opaque mode"
        rb = ref[k][..., :3].reshape(h // 4, 4, w // 4, 4, 3).transpose(0, 2, 1, 3, 4).reshape(h // 4, w // 4, 16, 3)
        ib = img[..., :3].reshape(h // 4, 4, w // 4, 4, 3).transpose(0, 2, 1, 3, 4).reshape(h // 4, w // 4, 16, 3)
        solid = (rb == rb[:, :, :1]).all(axis=(2, 3))
        assert (ib[solid] == ib[solid][:, :1]).all(), f"{name}: a solid block must decode to one colour"
        assert np.abs(ib[solid].astype(int) - rb[solid].astype(int)).max(initial=0) <= 4, f"{name}: solid blocks within the RGB565 step"


def cases():
    alpha, aexp = etc1s_with_alpha()
    opaque = synth.encode_etc1s(synth.texture_layers(64, 0, 3, 4))
    liam = read(golden_ktx2()[0])
    return [("alpha", alpha, aexp), ("opaque", opaque, oracle_ktx2(opaque)["rgba"]), ("liam", liam, oracle_ktx2(liam)["rgba"])]


def test_dxt_blocks_host_logic(built):
    for name, blob, ref in cases():
        for bc3 in (False, True):
            e = emu_ktx2_dxt(blob, bc3)
            assert e["status"] == 0, name
            check(name, blob, e["blocks"], bc3, ref)
    assert emu_ktx2_dxt(synth.encode_uastc(synth.texture_layers(16, 0, 1, 3), seed=9), True)["status"] == -3


@pytest.mark.gpu
def test_dxt_kernels_match_host_logic(uv, ctx):
    cs = cases()
    uastc = synth.encode_uastc(synth.texture_layers(16, 0, 1, 3), seed=9)
    for target, bc3, fmt in ((uv.TEX_BC1, False, "RGB_S3TC_DXT1_Format"), (uv.TEX_BC3, True, "RGBA_S3TC_DXT5_Format")):
        res = uv.KTX2Loader(ctx).transcode_batch([c[1] for c in cs] + [uastc], target=target)
        assert res[-1]["status"] == -3
        for (name, blob, ref), r in zip(cs, res):
            assert r["status"] == 0 and r["format"] == fmt, name
            assert np.array_equal(r["data"], emu_ktx2_dxt(blob, bc3)["blocks"]), name

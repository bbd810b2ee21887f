"""CPU: what pins the UASTC restatement (oracle/uastc_oracle.c).  No UASTC fixture or encoder exists in the reference
tree ("fully unpinned"), so the pins are redundancy checks on the restated tables, a round trip through
the independent synthetic encoder (tools/synth/uastc_encode.cpp shares only the generated pattern tables), and a
committed digest of the oracle's output on a seeded file (tests/golden/uastc_expected.json, made by
tests/golden/make_golden_uastc.py) so a later edit of the oracle cannot drift silently."""
import hashlib
import json
import os
import struct
import sys

import numpy as np

from conftest import GOLDEN, ROOT
from oracle_bind import oracle_ktx2

sys.path.insert(0, ROOT)
from tools.gen import gen_uastc_tables as gen  # noqa: E402
from tools.synth import synth  # noqa: E402

MODE_CODE = [(0x01, 4), (0x35, 6), (0x1D, 5), (0x03, 5), (0x13, 5), (0x0B, 5), (0x1B, 5), (0x07, 5), (0x17, 5), (0x0F, 5),
             (0x02, 3), (0x00, 2), (0x06, 3), (0x1F, 5), (0x0D, 5), (0x05, 7), (0x15, 6), (0x25, 6), (0x09, 4), (0x45, 7)]
COMPS = [3, 3, 3, 3, 3, 3, 3, 3, 0, 4, 4, 4, 4, 4, 4, 2, 2, 2, 3]
SUBSETS = [1, 1, 2, 3, 2, 1, 1, 2, 0, 2, 1, 1, 1, 1, 1, 1, 2, 1, 1]
PLANES = [1, 1, 1, 1, 1, 1, 2, 1, 0, 1, 1, 2, 1, 2, 1, 1, 1, 2, 1]
WBITS = [4, 2, 3, 2, 2, 3, 2, 2, 0, 2, 4, 2, 3, 1, 2, 4, 2, 2, 5]
EPRANGE = [19, 20, 8, 7, 12, 20, 18, 12, 0, 8, 13, 13, 19, 20, 20, 20, 20, 20, 11]
HINTS = [15, 15, 15, 15, 15, 15, 15, 15, 0, 23, 17, 17, 17, 23, 23, 23, 23, 23, 15]
BISE = {7: (2, 1, 0), 8: (4, 0, 0), 11: (5, 0, 0), 12: (3, 0, 1), 13: (4, 1, 0), 18: (5, 0, 1), 19: (6, 1, 0), 20: (8, 0, 0)}


def test_mode_codes_form_a_complete_prefix_code():
    assert sum(2.0 ** -n for _, n in MODE_CODE) == 1.0
    for i, (a, na) in enumerate(MODE_CODE):
        for j, (b, nb) in enumerate(MODE_CODE):
            if i != j and na <= nb:
                assert (b & ((1 << na) - 1)) != a, (i, j)


def test_mode_bit_budgets():
    full = []
    for m in range(19):
        if m == 8:
            continue
        bits, tr, qu = BISE[EPRANGE[m]]
        n = COMPS[m] * 2 * SUBSETS[m]
        tq = 0
        if tr:
            tq = (n // 5) * 8 + [0, 2, 4, 5, 7][n % 5]
        elif qu:
            tq = (n // 3) * 7 + [0, 3, 5][n % 3]
        pat = 0 if SUBSETS[m] == 1 else (4 if m == 3 else 5)
        ccs = 2 if PLANES[m] == 2 and m != 17 else 0
        total = MODE_CODE[m][1] + HINTS[m] + pat + ccs + tq + n * bits + 16 * PLANES[m] * WBITS[m] - SUBSETS[m] * PLANES[m]
        assert total <= 128, (m, total)
        if total == 128:
            full.append(m)
    assert full == [0, 6, 10, 11, 12, 16, 18]      # seven modes use every bit of the block


def test_partition_tables_are_consistent_with_bc7_and_the_astc_hash():
    assert gen.check()


def _file(seed=11, size=48, layers=2, mask=synth.UASTC_ALL_MODES, alpha=True):
    img = synth.texture_layers(size, 0, layers, seed)
    if alpha:
        img[..., 3] = ((np.arange(size)[None, :, None] * 5 + np.arange(size)[None, None, :] * 3) & 255).astype(np.uint8)
    return img, synth.encode_uastc(img, mode_mask=mask, seed=seed, has_alpha=alpha)


def test_round_trip_every_mode(built):
    for m in range(19):
        img, blob = _file(seed=5, size=32, layers=1, mask=(1 << m) | (1 << 8) if m != 8 else 1 << 8)
        o = oracle_ktx2(blob)
        assert o["status"] == 0 and o["is_uastc"] and o["layers"] == 1
        d = o["rgba"].reshape(img.shape).astype(np.float64) - img
        if m == 8:
            continue                                        # every block flattened to its first texel
        if COMPS[m] == 2:                                   # luminance + alpha
            assert (o["rgba"].reshape(img.shape)[..., 0] == o["rgba"].reshape(img.shape)[..., 1]).all()
            err = np.sqrt((d[..., 3] ** 2).mean())
        else:
            err = np.sqrt((d[..., :COMPS[m]] ** 2).mean())
            if COMPS[m] == 3:
                assert (o["rgba"].reshape(img.shape)[..., 3] == 255).all()
        assert err < 30.0, (m, err)                         # min/max endpoints + projection on a grainy image: coarse, but far from garbage (> 60)


def test_ragged_sizes_and_rejected_blocks(built):
    img = synth.texture_layers(16, 0, 1, 3)[:, :10, :13]
    o = oracle_ktx2(synth.encode_uastc(img, seed=9))
    assert o["status"] == 0 and (o["width"], o["height"]) == (13, 10) and o["rgba"].size == 13 * 10 * 4
    _, blob = _file(size=16, layers=1)
    lv = struct.unpack_from("<Q", blob, 80)[0]
    bad = bytearray(blob); bad[lv] = 0x45                    # mode 19 (reserved) -> transcodeImage fails
    assert oracle_ktx2(bytes(bad))["status"] == -2
    bad = bytearray(blob); bad[lv:lv + 4] = struct.pack("<I", 0x1D | (0x7FFF << 5) | (31 << 20))    # mode 2, pattern 31 >= 30
    assert oracle_ktx2(bytes(bad))["status"] == -2
    assert oracle_ktx2(blob[:lv + 100])["status"] == -1      # truncated level


def test_golden_digest(built):
    with open(os.path.join(GOLDEN, "uastc_expected.json")) as fh:
        exp = json.load(fh)
    _, blob = _file(**exp["args"])
    assert hashlib.sha256(blob).hexdigest() == exp["ktx2_sha256"]
    assert hashlib.sha256(oracle_ktx2(blob)["rgba"].tobytes()).hexdigest() == exp["rgba_sha256"]


def test_kernel_block_logic_on_the_host(built):
    """The product's per-block function (csrc/uastc_core.h, the code the sm_100a kernel runs) compiled for the host by
    tests/tools/basis_emu.cpp: bit-exact against the oracle on every mode, on 16k random blocks, on ragged sizes, and on a
    Zstd-supercompressed file (which also runs the product's KTX2 parse + Zstandard decoder); rejected blocks are rejected."""
    import struct
    from emu_bind import emu_ktx2
    files = [_file(seed=40 + m, size=32, layers=1, mask=1 << m) [1] for m in range(19)] + [_file(seed=7, size=64, layers=2)[1]]
    files.append(synth.encode_uastc(synth.texture_layers(16, 0, 1, 3)[:, :10, :13], seed=9))
    rng = np.random.default_rng(20260004)
    blob = bytearray(_file(size=512, layers=1)[1]); lv = struct.unpack_from("<Q", blob, 80)[0]
    import ctypes
    from oracle_bind import lib as olib
    L = olib(); px = (ctypes.c_uint8 * 64)(); rnd = rng.integers(0, 256, (128 * 128, 16), dtype=np.uint8); kept = 0
    for i in range(len(rnd)):
        if L.uvo_uastc_block_to_rgba(rnd[i].ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), px) == 0:
            blob[lv + 16 * i: lv + 16 * i + 16] = rnd[i].tobytes(); kept += 1
    assert kept > 6000
    files.append(bytes(blob))
    for f in files:
        e, o = emu_ktx2(f), oracle_ktx2(f)
        assert e["status"] == o["status"] == 0 and np.array_equal(e["rgba"].ravel(), o["rgba"].ravel())
    bad = bytearray(files[0]); lv0 = struct.unpack_from("<Q", bad, 80)[0]; bad[lv0] = 0x45
    assert emu_ktx2(bytes(bad))["status"] == -2
    try:
        from test_zstd import Z, compress
    except ImportError:
        Z = None
    if Z is not None:
        plain = files[19]; lv_off, lv_len = struct.unpack_from("<QQ", plain, 80)
        z = compress(plain[lv_off:lv_off + lv_len], 3); wrapped = bytearray(plain[:lv_off]) + z
        struct.pack_into("<I", wrapped, 44, 2); struct.pack_into("<QQQ", wrapped, 80, lv_off, len(z), lv_len)
        e = emu_ktx2(bytes(wrapped))
        assert e["status"] == 0 and np.array_equal(e["rgba"].ravel(), oracle_ktx2(plain)["rgba"].ravel())
        assert emu_ktx2(bytes(wrapped[:len(wrapped) - 50]))["status"] < 0

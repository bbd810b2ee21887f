"""CPU: the library's Zstandard decoder (csrc/zstd_inflate.cpp, KTX2 supercompressionScheme 2) against libzstd 1.5.5 -- a REAL
reference implementation, loaded here through ctypes as the oracle only (the product never links it).  Vectors are produced by
ZSTD_compress at several levels over inputs that exercise raw / RLE / compressed blocks, Huffman literals with direct and
FSE-coded weights, 1 and 4 streams, treeless literals, predefined / RLE / FSE / repeat sequence tables and repeat offsets."""
import ctypes
import ctypes.util
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, ROOT)


def _libzstd():
    for name in ("libzstd.so.1", ctypes.util.find_library("zstd")):
        try:
            if name:
                return ctypes.CDLL(name)
        except OSError:
            pass
    return None


Z = _libzstd()
pytestmark = pytest.mark.skipif(Z is None, reason="libzstd not present (test oracle only)")


def compress(data, level):
    Z.ZSTD_compressBound.restype = ctypes.c_size_t; Z.ZSTD_compressBound.argtypes = [ctypes.c_size_t]
    Z.ZSTD_compress.restype = ctypes.c_size_t
    Z.ZSTD_compress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int]
    cap = Z.ZSTD_compressBound(len(data)); buf = ctypes.create_string_buffer(cap)
    n = Z.ZSTD_compress(buf, cap, data, len(data), level)
    assert not Z.ZSTD_isError(n)
    return buf.raw[:n]


def inflate(uv, blob, cap):
    L = uv._native.lib()
    out = ctypes.create_string_buffer(max(cap, 1)); n = ctypes.c_size_t()
    rc = L.uvol_zstd_inflate(blob, len(blob), out, cap, ctypes.byref(n))
    return rc, out.raw[:n.value]


def corpus():
    rng = np.random.default_rng(7)
    words = [bytes(rng.integers(97, 123, rng.integers(2, 9), dtype=np.uint8)) for _ in range(300)]
    text = b" ".join(words[i] for i in rng.integers(0, 300, 60000))
    skew = rng.choice(np.arange(256, dtype=np.uint8), 300000, p=np.r_[np.full(16, 0.05), np.full(240, 0.2 / 240)]).tobytes()
    structured = (np.arange(200000, dtype=np.uint32) * 2654435761 >> 7).astype(np.uint16).tobytes()
    from tools.synth import synth
    uastc = synth.encode_uastc(synth.texture_layers(256, 0, 2, 5), seed=3)[-2 * 64 * 64 * 16:]
    rle_lits = text[:140000] + b"".join(text[o:o + 40] + b"z" for o in rng.integers(0, 100000, 3000))      # second block: matches separated by one 'z'
    pr = np.array([2.0 ** -(1 + i // 2) for i in range(16)]); pr /= pr.sum()
    four = rng.choice(np.arange(16, dtype=np.uint8), 200000, p=pr).tobytes()              # byte values 0..15 only: the weights are cheaper as raw nibbles (direct)
    return {"rle_literals": rle_lits, "four_symbols": four, "empty": b"", "one": b"x", "tiny": b"hello hello hello hello", "zeros": bytes(500000), "random": rng.bytes(300000),
            "text": text, "skewed_bytes": skew, "structured_u16": structured, "uastc_level": uastc,
            "mixed": bytes(70000) + rng.bytes(40000) + text[:150000] + bytes(range(256)) * 300,
            "long_matches": (text[:5000] * 60) + rng.bytes(1000) + (text[:5000] * 10)}


FEATURES = ["raw blocks", "RLE blocks", "compressed blocks", "raw literals", "RLE literals", "Huffman literals", "treeless literals", "four-stream literals",
            "direct weights", "FSE-coded weights", "predefined tables", "RLE tables", "FSE tables", "repeat tables", "repeat offsets", "frames"]


def test_matches_libzstd_and_covers_the_format(uv):
    L = uv._native.lib(); counts = (ctypes.c_uint64 * 16)()
    L.uvol_zstd_feature_counts(counts, 1)
    for level in (1, 3, 9, 19):
        for name, data in corpus().items():
            blob = compress(data, level)
            rc, out = inflate(uv, blob, len(data))
            assert rc == 0 and out == data, (name, level, rc, len(out), len(data))
    L.uvol_zstd_feature_counts(counts, 1)
    missing = [FEATURES[i] for i in range(16) if counts[i] == 0]
    assert not missing, (missing, list(counts))


def test_frames_and_errors(uv):
    c = corpus()
    a, b = compress(c["text"], 3), compress(c["tiny"], 3)
    skippable = b"\x50\x2a\x4d\x18" + (5).to_bytes(4, "little") + b"abcde"
    rc, out = inflate(uv, a + skippable + b, len(c["text"]) + len(c["tiny"]))                # concatenated + skippable frames
    assert rc == 0 and out == c["text"] + c["tiny"]
    assert inflate(uv, a[:len(a) // 2], len(c["text"]))[0] < 0                                # truncated
    assert inflate(uv, a, len(c["text"]) - 1)[0] < 0                                         # output too small
    assert inflate(uv, b"\x00\x01\x02\x03" + a, len(c["text"]))[0] == -2                      # bad magic
    bad = bytearray(a); bad[len(bad) // 2] ^= 0x55
    rc, out = inflate(uv, bytes(bad), len(c["text"]))
    assert rc < 0 or out != c["text"] or True                                                 # must not crash; any result is acceptable (no checksum check)

"""CPU: mutation fuzzing of the product's host-side parsers -- the Draco header walk, the KTX2 container parse and the Zstandard
decoder see untrusted bytes before anything reaches the GPU.  tests/tools/fuzz_host places every mutated input (and the inflate
output) between inaccessible guard pages, so a read or write outside a buffer kills the process; any status code is acceptable."""
import os
import subprocess
import sys

import pytest

from conftest import GOLDEN, ROOT, golden_drc, golden_ktx2

sys.path.insert(0, ROOT)
TOOLS = os.path.join(ROOT, "tests", "tools")


def test_host_parsers_survive_mutations(built, tmp_path):
    subprocess.run(["make", "-s", "-C", TOOLS, "fuzz_host"], check=True)
    import numpy as np
    from tools.synth import synth
    seeds = list(golden_drc()[:2]) + list(golden_ktx2())
    plain = synth.encode_uastc(synth.texture_layers(32, 0, 2, 5), seed=3)
    p = tmp_path / "uastc_plain.ktx2"; p.write_bytes(plain); seeds.append(str(p))
    from test_ktx2_mips import etc1s_chain, uastc_chain          # mip chains: the splitter (uvol_ktx2_split_levels) reads the same untrusted bytes
    p = tmp_path / "uastc_mips.ktx2"; p.write_bytes(uastc_chain()[0]); seeds.append(str(p))
    p = tmp_path / "etc1s_mips.ktx2"; p.write_bytes(etc1s_chain()[0]); seeds.append(str(p))
    import glob
    seeds += sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "corto", "*.crt")))[:3]      # V1 Corto frames (made by the reference's encoder)
    small = synth.make_sequence(1, 500, 32, want_textures=False, seed=5)[0][0]
    p = tmp_path / "small.drc"; p.write_bytes(small); seeds.append(str(p))
    rings, segs = synth.sphere_dims(500); fp, fu, uvs, _ = synth.sphere_topology(rings, segs); pos = synth.sphere_frame(rings, segs, 0.1, 5)
    p = tmp_path / "tagged.drc"; p.write_bytes(synth.encode_draco(pos, fp, uvs, fu, synth.vertex_normals(pos, fp), tagged=7)); seeds.append(str(p))      # TAGGED symbol scheme: the parser walks the tag runs
    try:
        from test_zstd import Z, compress, corpus
        if Z is not None:
            import struct
            lv_off, lv_len = struct.unpack_from("<QQ", plain, 80)
            z = compress(plain[lv_off:lv_off + lv_len], 3)
            wrapped = bytearray(plain[:lv_off]) + z
            struct.pack_into("<I", wrapped, 44, 2); struct.pack_into("<QQQ", wrapped, 80, lv_off, len(z), lv_len)
            p = tmp_path / "uastc_zstd.ktx2"; p.write_bytes(bytes(wrapped)); seeds.append(str(p))
            c = corpus()
            for name, level in (("text", 3), ("skewed_bytes", 9), ("four_symbols", 1), ("rle_literals", 19), ("tiny", 3)):
                p = tmp_path / (name + ".zst"); p.write_bytes(compress(c[name][:150000], level)); seeds.append(str(p))
    except ImportError:
        pass
    r = subprocess.run([os.path.join(TOOLS, "fuzz_host"), "1200", "20260017"] + seeds, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.returncode, r.stdout[-500:], r.stderr[-500:])
    assert "no access outside the buffers" in r.stdout

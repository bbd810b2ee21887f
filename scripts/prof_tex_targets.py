"""Small texture-only driver for ncu captures of the 16-byte-per-block targets: N synthetic UASTC 2048^2 x 7 segments transcoded once to
RGBA32, BC7 and ASTC 4x4 (k_uastc_blocks, k_uastc_blocks_16<Bc7Shared>, k_uastc_blocks_16<AstcShared>)."""
import importlib, os, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, root)
uv = importlib.import_module("universal-volumetric_b200")
from tools.synth import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
_, ktx, info = synth.make_sequence(7 * n, 2000, 2048, sequence_size=7, seed=20260003, distinct_geometry=1, distinct_textures=2, texture_format="uastc")
ctx = uv.Context(0); kl = uv.KTX2Loader(ctx)
for target in (uv.TEX_RGBA32, uv.TEX_BC7, uv.TEX_ASTC_4x4):
    out = kl.transcode_batch_raw(ktx, uv.MEM_DEVICE, target)
    print("target", target, "ok", sum(o.status == 0 for o in out), "of", len(ktx))

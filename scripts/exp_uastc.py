"""UASTC block kernels: CUDA-event time and algorithmic bandwidth for different mode mixes, RGBA32 (80 B per block) and BC7 (32 B per block)."""
import importlib, os, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, root)
uv = importlib.import_module("universal-volumetric_b200")
from tools.synth import synth
ctx = uv.Context(0, profiling=True); kl = uv.KTX2Loader(ctx)
img = synth.texture_layers(2048, 0, 7, 5)
for name, mask in (("modes 0-8,18 hashed per block", synth.UASTC_OPAQUE_MODES), ("mode 0 only", 1 | 256), ("all 19 modes hashed per block", synth.UASTC_ALL_MODES)):
    seg = synth.encode_uastc(img, mode_mask=mask, seed=3)
    ktx = [seg] * 32
    nb = 32 * 7 * 512 * 512
    for tname, target, per in (("rgba32", uv.TEX_RGBA32, 80), ("bc7", uv.TEX_BC7, 32)):
        out = kl.transcode_batch_raw(ktx, uv.MEM_DEVICE, target)
        for _ in range(3):          # timed on the replay path (resident inputs, one launch over the whole batch; the fresh path is pipelined over chunks)
            out = kl.replay_raw(len(ktx), uv.MEM_DEVICE)
        st = ctx.stats(1)
        print(name, tname, "blocks %.3f ms -> %.0f GB/s (%d B/block), %.2f Gblocks/s" % (st["stages"]["blocks"], nb * per / st["stages"]["blocks"] / 1e6, per, nb / st["stages"]["blocks"] / 1e6), flush=True)

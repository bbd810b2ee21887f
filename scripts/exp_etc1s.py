"""ETC1S block kernel: CUDA-event time and algorithmic bandwidth (4 B of indices in + 64 B RGBA out per block)."""
import importlib, os, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, root)
uv = importlib.import_module("universal-volumetric_b200")
from tools.synth import synth
ctx = uv.Context(0, profiling=True); kl = uv.KTX2Loader(ctx)
for size, nseg in ((1024, 43), (2048, 16)):
    seg = synth.encode_etc1s(synth.texture_layers(size, 0, 7, 5))
    ktx = [seg] * nseg
    out = kl.transcode_batch_raw(ktx, uv.MEM_DEVICE)
    for _ in range(3):          # timed on the replay path (resident inputs, one launch over the whole batch)
        out = kl.replay_raw(len(ktx), uv.MEM_DEVICE)
    st = ctx.stats(1)
    nb = nseg * 7 * (size // 4) ** 2
    print("etc1s %d^2 x7 x%d: blocks %.3f ms -> %.0f GB/s (68 B/block); slices %.1f resolve %.1f" % (size, nseg, st["stages"]["blocks"], nb * 68 / st["stages"]["blocks"] / 1e6, st["stages"]["slices"], st["stages"]["resolve"]), flush=True)

"""Multi-GPU check of BASELINE configs[3] (run under torchrun, one rank per GPU): ONE sequence is frame-sharded with
manifest.shard_v2, every rank decodes only its shard on its GPU, the decoded shards (geometry and textures) are gathered on every
rank over NCCL (gather.all_gather_shard), and the gathered bytes -- walked frame by frame / segment by segment -- must equal what a
single GPU produces when it decodes the whole sequence alone (rank 0 does that as the reference run).  Prints one line per rank."""
import hashlib
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 35
    verts = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
    fmt = sys.argv[3] if len(sys.argv) > 3 else "uastc"
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    uv = importlib.import_module("universal-volumetric_b200")
    from tools.synth import synth
    seq = 7
    drc, ktx, info = synth.make_sequence(frames, verts, 64, sequence_size=seq, seed=20260031, texture_format=fmt)      # the same sequence on every rank
    f0, f1, s0, s1 = uv.shard_v2(frames, seq, len(ktx), world, rank)
    ctx = uv.Context(local)
    g, t = uv.V2Player(ctx).decode_step_raw(drc[f0:f1], ktx[s0:s1], uv.MEM_DEVICE)
    G = uv.gather.all_gather_shard(g, f1 - f0, t, s1 - s0, f"cuda:{local}")
    h = hashlib.sha256(); nbytes = 0
    for r in range(world):                         # rank order == frame order (contiguous shards): every geometry frame of the sequence ...
        for i in range(len(G["gtabs"][r])):
            v = uv.gather.shard_frame_views(G, r, i)
            assert v is not None, (r, i)
            for k in ("index", "position", "normal", "uv"):
                b = v[k].cpu().numpy().tobytes(); h.update(b); nbytes += len(b)
    for r in range(world):                         # ... then every texture segment, like the single-GPU walk below
        for i in range(len(G["ttabs"][r])):
            b = uv.gather.shard_texture_view(G, r, i).cpu().numpy().tobytes(); h.update(b); nbytes += len(b)
    mine = h.hexdigest()
    ref = None
    if rank == 0:                                  # the whole sequence on one GPU
        c1 = uv.Context(local)
        g1, t1 = uv.V2Player(c1).decode_step_raw(drc, ktx, uv.MEM_HOST)
        h1 = hashlib.sha256()
        for x in g1[:frames]:
            assert x.status == 0
            for p, n in ((x.index, x.num_faces * 12), (x.position, x.num_points * 12), (x.normal, x.num_points * 12), (x.uv, x.num_points * 8)):
                h1.update(np.ctypeslib.as_array(ctypes_u8(p), (n,)).tobytes())
        for x in t1[:len(ktx)]:
            assert x.status == 0
            h1.update(np.ctypeslib.as_array(x.data, (int(x.bytes),)).tobytes())
        ref = h1.hexdigest(); c1.close()
    box = [ref]
    dist.broadcast_object_list(box, src=0)
    ok = mine == box[0]
    print(f"rank {rank}/{world}: shard frames [{f0},{f1}) segments [{s0},{s1}), gathered {nbytes} bytes, sha256 {mine[:16]}, single-GPU {box[0][:16]} -> {'GATHER_IDENTICAL' if ok else 'MISMATCH'}", flush=True)
    ctx.close()
    dist.destroy_process_group()
    return 0 if ok else 1


def ctypes_u8(p):
    import ctypes
    return ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8))


if __name__ == "__main__":
    sys.exit(main())

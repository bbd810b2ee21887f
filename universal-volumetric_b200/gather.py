"""Optional final gather of a frame-sharded decode (BASELINE configs[3]).

The decode itself never communicates: geometry frames and KTX2 segments are independent units, rank r decodes its own
contiguous shard (manifest.shard_v2).  Only when the caller wants EVERY frame on EVERY rank (one renderer process fed by
all GPUs) are the decoded buffers exchanged -- one all_gather of a small per-frame table (status, counts, byte offsets
inside the rank's output arena) and one all_gather of the frames' bytes (contiguous runs of the output arena packed back to back), padded to the largest rank.  With the NCCL
backend the arenas are the library's device buffers (no host staging; NVLink / NVSwitch carries the payload); the same
code runs on CPU tensors under gloo for the tests.  `torch.distributed` is plumbing here, not the product.
"""
import ctypes

import numpy as np

COLS = 8        # status, num_points, num_faces, off_index, off_position, off_normal, off_uv, arena_bytes


class _RawCuda:
    """Exposes a raw device pointer through __cuda_array_interface__ so torch can wrap it without a copy."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def _addr(p):
    return ctypes.cast(p, ctypes.c_void_p).value or 0


def geometry_table(raw, n, gap=1 << 16):
    """(table int64[n, COLS], runs) of the first n results of a *_batch call (device or host pointers).  The arrays of those frames
    are grouped into contiguous RUNS of the library's output arena (the arena holds all index buffers first, then the per-point
    arrays, so a prefix of the frames is two runs); `runs` = [(address, bytes)], and the table's offsets are relative to the
    runs packed back to back -- only the bytes of the selected frames travel."""
    t = np.full((n, COLS), -1, np.int64)
    spans = []
    for i in range(n):
        g = raw[i]
        t[i, 0] = g.status
        if g.status != 0:
            continue
        t[i, 1], t[i, 2] = g.num_points, g.num_faces
        for col, (p, nbytes) in enumerate(((g.index, g.num_faces * 12), (g.position, g.num_points * 12), (g.normal, g.num_points * 12), (g.uv, g.num_points * 8)), start=3):
            a = _addr(p)
            if a:
                t[i, col] = a
                spans.append((a, a + nbytes))
    if not spans:
        return t, []
    spans.sort()
    runs = [list(spans[0])]
    for a, e in spans[1:]:
        if a <= runs[-1][1] + gap:
            runs[-1][1] = max(runs[-1][1], e)
        else:
            runs.append([a, e])
    packed, starts = 0, []
    for a, e in runs:
        starts.append((a, e, packed)); packed += (e - a + 127) // 128 * 128
    for col in range(3, 7):
        for i in range(n):
            a = t[i, col]
            if a >= 0:
                for ra, re_, off in starts:
                    if ra <= a < re_:
                        t[i, col] = a - ra + off
                        break
    t[:, 7] = packed
    return t, [(a, e - a) for a, e in runs]


def pack_runs(runs, device):
    """One contiguous uint8 tensor holding the runs back to back (128-byte aligned), copied on `device`."""
    import torch
    total = sum((nb + 127) // 128 * 128 for _, nb in runs)
    out = torch.zeros(total, dtype=torch.uint8, device=device)
    off = 0
    for a, nb in runs:
        out[off:off + nb] = arena_tensor(a, nb, device); off += (nb + 127) // 128 * 128
    return out


def arena_tensor(base, nbytes, device):
    """uint8 torch view of [base, base + nbytes): device memory when `device` is a CUDA device, else pinned / pageable host memory."""
    import torch
    if nbytes == 0:
        return torch.empty(0, dtype=torch.uint8, device=device)
    if torch.device(device).type == "cuda":
        return torch.as_tensor(_RawCuda(base, nbytes), device=device)
    return torch.from_numpy(np.ctypeslib.as_array((ctypes.c_uint8 * nbytes).from_address(base)))


def all_gather_geometry(raw, n, device, group=None):
    """Every rank contributes the n geometry frames of its last batch; returns (tables [world, max_n, COLS] int64 on CPU,
    arenas [world, max_bytes] uint8 on `device`).  Frame i of rank r: tables[r, i] / views from `frame_views`."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    table, runs = geometry_table(raw, n)
    mine_packed = pack_runs(runs, device); nbytes = int(mine_packed.numel())
    sizes = torch.tensor([n, nbytes], dtype=torch.int64, device=device)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    max_n = int(max(int(s[0]) for s in all_sizes)); max_b = int(max(int(s[1]) for s in all_sizes))
    tpad = torch.full((max_n, COLS), -1, dtype=torch.int64, device=device); tpad[:n] = torch.from_numpy(table).to(device)
    tables = torch.empty((world, max_n, COLS), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(tables, tpad, group=group) if hasattr(dist, "all_gather_into_tensor") and torch.device(device).type == "cuda" else \
        dist.all_gather(list(tables.unbind(0)), tpad, group=group)
    mine = torch.zeros(max_b, dtype=torch.uint8, device=device)
    if nbytes:
        mine[:nbytes] = mine_packed
    arenas = torch.empty((world, max_b), dtype=torch.uint8, device=device)
    if torch.device(device).type == "cuda":
        dist.all_gather_into_tensor(arenas, mine, group=group)
    else:
        dist.all_gather(list(arenas.unbind(0)), mine, group=group)
    return tables.cpu(), arenas


def frame_views(tables, arenas, rank, i):
    """Typed views (no copy) of frame i of `rank` inside the gathered arenas: dict(index, position, normal, uv) or None."""
    import torch
    row = tables[rank, i]
    if int(row[0]) != 0:
        return None
    P, F = int(row[1]), int(row[2]); a = arenas[rank]
    out = {}
    for name, col, dt, shape in (("index", 3, torch.int32, (F * 3,)), ("position", 4, torch.float32, (P, 3)), ("normal", 5, torch.float32, (P, 3)), ("uv", 6, torch.float32, (P, 2))):
        off = int(row[col])
        if off < 0:
            out[name] = None
            continue
        count = 1
        for s in shape:
            count *= s
        out[name] = a[off: off + 4 * count].view(dt).view(*shape)
    return out

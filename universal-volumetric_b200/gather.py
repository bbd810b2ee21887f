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


# ---------------------------------------------------------------------------------------------------------------------
# Whole-shard gather (BASELINE configs[3]: one sequence frame-sharded across G GPUs, then gathered): geometry AND textures.
# A rank's decoded shard is two contiguous spans of library-owned memory (the geometry output arena: every index buffer, then
# the per-point arrays; the texture output arena: the segments back to back), so nothing is packed or staged: after one
# all_gather of the small per-item tables, rank r copies its spans into its own slot of one receive arena and broadcasts the
# slot to every other rank (an all-gather with per-rank sizes; NVLink / NVSwitch carries the payload).
GCOLS = 8       # geometry rows: status, num_points, num_faces, off_index, off_position, off_normal, off_uv, span_bytes
TCOLS = 8       # texture rows:  status, width, height, layers, format, has_alpha, off_data, bytes


def _span(ptrs_and_sizes):
    lo, hi = None, 0
    for a, nb in ptrs_and_sizes:
        if a and nb:
            lo = a if lo is None else min(lo, a); hi = max(hi, a + nb)
    return (lo or 0), (hi - lo if lo else 0)


def shard_tables(geo, n_geo, tex, n_tex):
    """Tables + spans of one rank's decoded shard.  Offsets in the tables are relative to the start of the rank's geometry span /
    texture span.  -> (gtab int64[n_geo, GCOLS], ttab int64[n_tex, TCOLS], (geo_addr, geo_bytes), (tex_addr, tex_bytes))"""
    gt = np.full((max(n_geo, 0), GCOLS), -1, np.int64); tt = np.full((max(n_tex, 0), TCOLS), -1, np.int64)
    arrs = []
    for i in range(n_geo):
        g = geo[i]; gt[i, 0] = g.status
        if g.status == 0:
            arrs += [(_addr(g.index), g.num_faces * 12), (_addr(g.position), g.num_points * 12), (_addr(g.normal), g.num_points * 12), (_addr(g.uv), g.num_points * 8)]
    gbase, gbytes = _span(arrs)
    for i in range(n_geo):
        g = geo[i]
        if g.status == 0:
            gt[i, 1], gt[i, 2] = g.num_points, g.num_faces
            for col, p in enumerate((g.index, g.position, g.normal, g.uv), start=3):
                a = _addr(p); gt[i, col] = a - gbase if a else -1
    gt[:, 7] = gbytes
    tbase, tbytes = _span([(_addr(t.data), int(t.bytes)) for t in tex[:n_tex] if t.status == 0])
    for i in range(n_tex):
        t = tex[i]; tt[i, 0] = t.status
        if t.status == 0:
            tt[i, 1:6] = (t.width, t.height, t.layers, t.format, t.has_alpha); tt[i, 6] = _addr(t.data) - tbase; tt[i, 7] = int(t.bytes)
    return gt, tt, (gbase, gbytes), (tbase, tbytes)


def all_gather_shard(geo, n_geo, tex, n_tex, device, group=None, arena=None):
    """Gathers every rank's decoded shard on every rank.  Returns dict(gtabs=[world][n, GCOLS], ttabs=[world][n, TCOLS], arena=uint8 tensor
    on `device`, used=bytes of the arena in use, geo_off=[world], tex_off=[world], bytes=[world]): rank r's geometry span starts at
    arena[geo_off[r]], its texture span at arena[tex_off[r]].  `arena` (optional) is a caller-owned receive buffer that is reused when large enough."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    cuda = torch.device(device).type == "cuda"
    gt, tt, (gb, gn), (tb, tn) = shard_tables(geo, n_geo, tex, n_tex)
    meta = torch.tensor([n_geo, n_tex, gn, tn], dtype=torch.int64, device=device)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    metas = [[int(x) for x in m.cpu()] for m in metas]
    mg, mt = max(m[0] for m in metas), max(m[1] for m in metas)
    pad = torch.full((mg * GCOLS + mt * TCOLS,), -1, dtype=torch.int64, device=device)
    pad[:n_geo * GCOLS] = torch.from_numpy(gt.reshape(-1)).to(device); pad[mg * GCOLS: mg * GCOLS + n_tex * TCOLS] = torch.from_numpy(tt.reshape(-1)).to(device)
    tabs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(tabs, pad, group=group)
    tabs = [t.cpu().numpy() for t in tabs]
    gtabs = [tabs[r][: metas[r][0] * GCOLS].reshape(-1, GCOLS) for r in range(world)]
    ttabs = [tabs[r][mg * GCOLS: mg * GCOLS + metas[r][1] * TCOLS].reshape(-1, TCOLS) for r in range(world)]
    al = lambda v: (v + 255) // 256 * 256
    if cuda:
        # NCCL: ONE all-gather over equal slots (a slot = the largest rank's geometry span + texture span; shards of one sequence differ
        # by at most one segment, so the padding is under 1 %).  Every rank copies its two spans out of the library's buffers into its
        # own slot of the receive arena -- the collective then runs in place on torch-owned memory, all links busy in both directions.
        slot = max(al(m[2]) + al(m[3]) for m in metas)
        geo_off = [r * slot for r in range(world)]; tex_off = [r * slot + al(metas[r][2]) for r in range(world)]; cur = world * slot
        if arena is None or arena.numel() < cur:
            arena = torch.empty(cur, dtype=torch.uint8, device=device)
        for off, nb, base in ((geo_off[rank], gn, gb), (tex_off[rank], tn, tb)):
            if nb:
                arena[off: off + nb].copy_(arena_tensor(base, nb, device))
        if slot:
            dist.all_gather_into_tensor(arena[:cur], arena[rank * slot: (rank + 1) * slot], group=group)
    else:
        geo_off, tex_off, cur = [], [], 0
        for m in metas:
            geo_off.append(cur); cur += al(m[2]); tex_off.append(cur); cur += al(m[3])
        if arena is None or arena.numel() < cur:
            arena = torch.empty(cur, dtype=torch.uint8, device=device)
        for r in range(world):          # host backends (gloo, the CPU tests): per-rank sized broadcasts
            for off, nb, base in ((geo_off[r], metas[r][2], gb), (tex_off[r], metas[r][3], tb)):
                if nb == 0:
                    continue
                slot = arena[off: off + nb]
                if r == rank:
                    slot.copy_(arena_tensor(base, nb, device))
                dist.broadcast(slot, src=dist.get_global_rank(group, r) if group is not None else r, group=group)
    return {"gtabs": gtabs, "ttabs": ttabs, "arena": arena, "used": cur, "geo_off": geo_off, "tex_off": tex_off, "bytes": [m[2] + m[3] for m in metas], "cuda": cuda}


def shard_frame_views(G, rank, i):
    """Typed views of geometry frame i of `rank` in a gathered arena (dict(index, position, normal, uv)), or None if it failed."""
    import torch
    row = G["gtabs"][rank][i]
    if int(row[0]) != 0:
        return None
    P, F = int(row[1]), int(row[2]); a = G["arena"][G["geo_off"][rank]:]
    out = {}
    for name, col, dt, cnt in (("index", 3, torch.int32, F * 3), ("position", 4, torch.float32, P * 3), ("normal", 5, torch.float32, P * 3), ("uv", 6, torch.float32, P * 2)):
        off = int(row[col])
        out[name] = None if off < 0 else a[off: off + 4 * cnt].view(dt)
    return out


def shard_texture_view(G, rank, i):
    """uint8 view of texture segment i of `rank` in a gathered arena (all layers back to back), or None if it failed."""
    row = G["ttabs"][rank][i]
    if int(row[0]) != 0:
        return None
    o = G["tex_off"][rank] + int(row[6])
    return G["arena"][o: o + int(row[7])]

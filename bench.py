#!/usr/bin/env python
"""bench.py -- decoded frames/sec (geometry + texture) of the UVOL V2 decode hot path.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W [--impl reference]`
prints ONE JSON line on rank 0.

Workload at N=1 (default `--workload c3`): the configuration BASELINE.json's metric is quoted on, configs[2] -- a
synthetic 1000-frame V2 sequence, 200k verts/frame Draco geometry, 2048^2 UASTC KTX2 textures with sequenceSize 7
(143 segments), built by tools/synth (deterministic).  `--workload c2` is configs[1] (300 frames, 50k verts, 1024^2
ETC1S), `--workload liam` the reference's 250 real frames / 50 real segments cycled to 1000 frames, `--workload c5` configs[4]
(V1 Corto).  One "step" = one pass of the hot path over the whole sequence; C3 is ONE window of 1000 frames (82.5 MB of scratch
per frame).  A sequence whose scratch would not fit HBM at once can still be decoded in WINDOWS of whole segments (`window_segments`
of a workload; the prefetch window of src/V2/player.ts:272-323), one ctx per window.
N>1 (BASELINE configs[3]): ONE sequence (same seed on every rank) is frame-sharded with manifest.shard_v2 -- rank r decodes only
its contiguous block of whole KTX2 segments and the geometry frames they cover, no data-path collective -- and the decoded
shards (geometry AND textures) are then gathered on every rank over NCCL / NVLink (gather.all_gather_shard).  "scaling":
"strong"; a step = decode of the shard + the gather; value = frames of the sequence / max-over-ranks step time.  The old
weak-scaling figure (every rank decodes a whole sequence of its own) is kept under the extra key "weak".

  value     frames/s with the compressed inputs already resident in HBM (uvol_replay_v2_batch), device
            time from CUDA events on the library's streams (geometry and texture run concurrently, the
            step time is the longer of the two spans), outputs left in HBM.
  e2e       frames/s through the C ABI with HOST buffers: uvol_decode_v2_batch with UVOL_MEM_HOST (host parse, H2D of the
            compressed bytes, kernels, D2H of every decoded buffer into pinned host memory), wall clock around the calls.
  bc7_target  the same sequence with UVOL_TEX_BC7 output (what the reference picks on desktop GPUs): value and e2e.
  astc_target the same sequence with UVOL_TEX_ASTC_4x4 output (UASTC workloads; the reference's choice on GPUs with ASTC): lossless repack.
  roofline  for the kernel stage with the largest share of the step (ALGORITHMIC bytes of that stage /
            its CUDA-event duration vs the measured HBM peak), plus the same for every stage.
  cpu_baseline  the CPU oracle (oracle/liboracle.so, a restatement: "port") on all host cores over a
            bounded sample of the same workload.
`--impl reference` times that CPU oracle as the reference arm (the reference's own Draco/Basis WASM
cannot run here).
"""
import argparse
import ctypes
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

def _metric():
    """BASELINE.json's metric string (quoted on the 200k-vert / 2048^2 sequence = the default workload c3)."""
    try:
        return json.load(open(os.path.join(ROOT, "BASELINE.json")))["metric"]
    except Exception:
        return "decoded frames/sec (geom+tex) on 200k-vert/2048\u00b2 seq; Mverts/s + Mtexels/s"


METRIC = _metric()
WORKLOADS = {
    # window_segments: segments decoded per library call (None = the whole sequence at once); distinct_*: how many distinct
    # frames / segments are actually encoded by the generator (the rest cycle through them) to bound generation time.
    "c3": dict(frames=1000, verts=200000, tex=2048, seq=7, seed=20260003, fmt="uastc", window_segments=None, distinct_geo=32, distinct_tex=6,
               label="configs[2]: 1000-frame V2 seq, 200k verts/frame Draco, 2048^2 UASTC KTX2 batch=7"),
    "c2": dict(frames=300, verts=50000, tex=1024, seq=7, seed=20260002, fmt="etc1s", window_segments=None, distinct_geo=None, distinct_tex=None,
               label="configs[1]: 300-frame V2 seq, 50k verts/frame Draco, 1024^2 ETC1S KTX2 batch=7"),
    "c5": dict(frames=300, verts=50000, tex=1024, seq=1, seed=20260005, fmt="corto", window_segments=None, distinct_geo=8, distinct_tex=None,
               label="configs[4]: V1 manifest, 300-frame Corto .crt geometry (position 12 bit + uv 12 bit, u32 index); V1's mp4 texture leg is out of scope"),
    "tiny": dict(frames=28, verts=2000, tex=64, seq=7, seed=20260009, fmt="uastc", window_segments=None, distinct_geo=None, distinct_tex=None,
                 label="tiny smoke workload"),
    # the reference's own capture (example/public/liam/output): 250 real .drc frames and 50 real ETC1S segments of 5 x 1024^2, cycled to
    # 1000 frames / 200 segments -- real entropy statistics, real seams (SURVEY 8d "liam x N")
    "liam": dict(frames=1000, verts=26145, tex=1024, seq=5, seed=0, fmt="etc1s", window_segments=None, distinct_geo=250, distinct_tex=50,
                 label="liam x4: the reference's 250 real Draco frames (26k verts, 52k faces) + 50 real ETC1S KTX2 segments (5 x 1024^2), cycled to 1000 frames"),
}


def measured_peak():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = max([int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()] or [0])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons, "samples": len(sm)}


def bind_to_gpu_cpus(local, world):
    """One process per GPU: pin this rank to the CPUs NVML reports as local to its GPU (pinned staging / result buffers are then
    allocated and touched on the GPU's NUMA node) and split the staging threads between the ranks.  Returns the CPU count used."""
    ncpu = len(os.sched_getaffinity(0))
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        phys = int(vis.split(",")[local]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else local
        words = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(phys), (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1} & os.sched_getaffinity(0)
        if cpus and world > 1:
            os.sched_setaffinity(0, cpus); ncpu = len(cpus)
    except Exception:
        pass
    os.environ.setdefault("UVOL_STAGING_THREADS", str(max(1, min(8, (os.cpu_count() or 1) // (2 * max(1, world))))))
    return ncpu


def make_workload(name, rank=0):
    """The workload's files.  `rank` only varies the seed of the synthetic generators (the weak-scaling side figure); the headline
    decodes ONE sequence, rank 0's."""
    w = WORKLOADS[name]
    t0 = time.time()
    if name == "liam":
        import glob
        fx = os.path.join(ROOT, "oracle", "_ref", "fixtures")
        d = sorted(glob.glob(os.path.join(fx, "geometry_draco", "*.drc"))) or sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "liam", "*.drc")))
        k = sorted(glob.glob(os.path.join(fx, "texture_ktx2", "*.ktx2"))) or sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "liam", "*.ktx2")))
        db = [open(f, "rb").read() for f in d]; kb = [open(f, "rb").read() for f in k]
        nseg = (w["frames"] + w["seq"] - 1) // w["seq"]
        drc = [db[i % len(db)] for i in range(w["frames"])]; ktx = [kb[i % len(kb)] for i in range(nseg)]
        info = {"verts": w["verts"], "faces": 2 * w["verts"], "tex_size": w["tex"], "distinct_geometry": len(db), "distinct_textures": len(kb),
                "source": "oracle/_ref/fixtures" if len(db) > 3 else "tests/golden/liam (full fixture set not staged)"}
    else:
        from tools.synth import synth
        drc, ktx, info = synth.make_sequence(w["frames"], w["verts"], w["tex"], sequence_size=w["seq"], seed=w["seed"] + 1000 * rank,
                                             distinct_geometry=w["distinct_geo"], distinct_textures=w["distinct_tex"], texture_format=w["fmt"])
    info["gen_s"] = round(time.time() - t0, 2)
    return drc, ktx, info


def workload_config(name, W, info, n_gpus, n_segments, target):
    """`config` of the JSON line: the same keys and values in our arm and in the reference arm."""
    return {"workload": f"{W['label']} ({'real fixtures' if name == 'liam' else 'synthetic, tools/synth seed %d' % W['seed']})",
            "frames": W["frames"], "segments": n_segments, "sequence_size": W["seq"], "verts": info["verts"], "faces": info["faces"],
            "texture": f"{W['tex']}x{W['tex']} {W['fmt']}", "texture_target": target,
            "distinct_geometry_frames": info["distinct_geometry"], "distinct_texture_segments": info["distinct_textures"],
            "l2": "flushed between timed iterations (256 MiB memset); the working set is far larger than L2",
            "parallelism": f"one sequence frame-sharded over {n_gpus} GPU(s) (manifest.shard_v2), NCCL gather of the decoded shards when > 1"}


def make_windows(drc, ktx, seq, window_segments):
    """Splits the sequence into windows of whole segments: [(drc files, ktx2 files)]."""
    ws = window_segments or max(1, len(ktx))
    return [(drc[s * seq:(s + ws) * seq], ktx[s:s + ws]) for s in range(0, max(1, len(ktx)), ws)]


def cpu_oracle_run(drc, ktx, threads):
    """Times the CPU oracle over the given files; returns (frames/s, seconds)."""
    from oracle_bind import lib
    L = lib()
    L.uvo_draco_decode_batch.argtypes = [ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_size_t), ctypes.c_int, ctypes.c_int,
                                         ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64)]
    L.uvo_ktx2_decode_batch.argtypes = [ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_size_t), ctypes.c_int, ctypes.c_int,
                                        ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64)]

    def arr(bl):
        return (ctypes.c_char_p * len(bl))(*bl), (ctypes.c_size_t * len(bl))(*[len(b) for b in bl])
    da, dl = arr(drc); ka, kl = arr(ktx)
    a = ctypes.c_uint64(); b = ctypes.c_uint64(); c = ctypes.c_uint64()
    t0 = time.perf_counter()
    ok1 = L.uvo_draco_decode_batch(da, dl, len(drc), threads, None, ctypes.byref(a), ctypes.byref(b))
    t1 = time.perf_counter()
    ok2 = L.uvo_ktx2_decode_batch(ka, kl, len(ktx), threads, None, ctypes.byref(c))
    t2 = time.perf_counter()
    assert ok1 == len(drc) and ok2 == len(ktx), "CPU oracle failed on the bench workload"
    return t1 - t0, t2 - t1, a.value, c.value


def cpu_sample(drc, ktx, seq, nseg):
    """Bounded sample of the same workload: the first nseg FULL segments and the geometry frames they cover."""
    nseg = max(1, min(nseg, len(ktx) - 1 if len(drc) % seq else len(ktx)))
    return drc[: nseg * seq], ktx[:nseg]


TARGETS = {"rgba32": 0, "etc1": 1, "bc7": 2, "astc": 4, "etc2": 5}


def stage_bytes(info, P_total, frames, bytes_in_geo, bytes_in_tex, fmt, target="rgba32"):
    """ALGORITHMIC bytes per step for each kernel stage (DESIGN.md 'Kernels and rooflines')."""
    F, V = info["faces"], info["verts"]
    nblk = (info["tex_size"] // 4) ** 2
    P = P_total / frames
    g = {
        "edgebreaker": frames * (F * 1 + 2 * 3 * F * 4),                       # symbols in, corner table (opp + c2v) out
        "traverse": frames * 3 * (F * 32 + F + 2 * V * 4),                    # per table: 32 B face records in, visited bytes + entry maps out
        "face_records": frames * 3 * (2 * 3 * F * 4 + 3 * F + F * 32),        # per table: corner table + seam flags in, 32 B records out
        "rans_attr": bytes_in_geo + frames * (3 * V + 2 * P + 2 * V) * 4,      # compressed in, int32 corrections out
        "predict_wrap": frames * (V * 16 + 2 * V * 12),                       # parents + corrections in, values out
        "predict_uv": frames * P * (32 + 8 + 8),
        "normals": frames * V * (8 + 8 + 7 * 12),
        "expand_pnc": frames * (P * 4 + P * 2 * 8 + P * 20 + P * 24),          # positions + normals: p2c + (corner->vertex, vertex->entry) x 2 + 20 B of int values in, 24 B/point out
        "expand": frames * (P * 4 + P * 8 + P * 8 + P * 8),                    # uv: p2c + one table walk + 8 B of int values in, 8 B/point out
        "seams": frames * (3 * F * 4 + 2 * 3 * F),
        "attr_tables": frames * 2 * (3 * F * 4 + 3 * F * 4),
        "point_assign": frames * (3 * F * 4 * 3),
    }
    t = {
        "slices": bytes_in_tex + frames * nblk * 5,                            # VLC bits in, {pred u8, delta/selector u16} out
        "resolve": frames * nblk * (1 + 2 + 2 + 2),
        "blocks": frames * nblk * ((16 if fmt == "uastc" else 4) + {"rgba32": 64, "etc1": 8, "bc7": 16, "astc": 16, "etc2": 16}[target]),   # UASTC: 16 B block in; ETC1S: 2x u16 indices in; 64 B RGBA / 16 B BC7 or ASTC / 8 B ETC1 out
    }
    return g, t


def bench_v1(args, W, rank, world, local):
    """configs[4]: V1 Corto geometry.  Frames are encoded AND (for the CPU baseline / reference arm) decoded by the reference's own
    C++ (oracle/_ref/libcorto_ref.so, built from deprecated/encoder/dev/src) -- cpu_baseline.kind = "reference"."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    import corto_bind
    from tools.synth import synth
    frames, verts = W["frames"], W["verts"]; ncores = os.cpu_count() or 1
    workload_name = f"{W['label']} (synthetic sphere, reference encoder, seed {W['seed']})"
    if not corto_bind.available():
        if rank == 0:
            print(json.dumps({"impl": args.impl, "unavailable": "oracle/_ref/libcorto_ref.so not built (needs the reference tree at build time)"}))
        return 0
    rings, segs = synth.sphere_dims(verts); fp, fu, uvs, nv = synth.sphere_topology(rings, segs)
    t0 = time.time(); enc = []
    for i in range(W["distinct_geo"]):
        pos = synth.sphere_frame(rings, segs, i / 30.0, W["seed"] + 1000 * rank)
        uvv = np.stack([np.arctan2(pos[:, 2], pos[:, 0]) / (2 * np.pi) + 0.5, pos[:, 1] / 2000.0 + 0.5], 1).astype(np.float32)
        enc.append(corto_bind.ref_encode(pos, uvv, fp, 12, 12))
    gen_s = time.time() - t0
    crt = [enc[i % len(enc)] for i in range(frames)]
    blobs = [c[0] for c in crt]

    def cpu_pass(items, threads):
        t = time.perf_counter()
        with ThreadPoolExecutor(threads) as ex:          # ctypes releases the GIL inside the reference decoder
            list(ex.map(lambda c: corto_bind.ref_decode(c[0], c[1], c[2]), items))
        return time.perf_counter() - t

    if args.impl == "reference":
        if rank != 0:
            return 0
        sample = crt[: max(ncores * 4, 32)]
        cpu_pass(sample[:ncores], ncores)
        dt = sum(cpu_pass(sample, ncores) for _ in range(args.steps)); fps = len(sample) * args.steps / dt
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32+f32", "data": "synthetic",
                          "config": {"workload": workload_name, "frames_per_step": len(sample)},
                          "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": ncores, "kind": "reference",
                                           "sample": f"{len(sample)} frames per step, {ncores} threads, the reference's own crt::Decoder (oracle/_ref/libcorto_ref.so)"},
                          "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))
        return 0
    import torch
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    uv = importlib.import_module("universal-volumetric_b200")
    ctx = uv.Context(local, profiling=True); dec = uv.CortoDecoder(ctx)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(max(args.warmup, 1)):
        out = dec.decode_batch_raw(blobs, uv.MEM_HOST)
    assert all(m.status == 0 for m in out[:frames]), "corto decode failed on the bench workload"
    V_total = sum(m.num_vertices for m in out[:frames]); F_total = sum(m.num_faces for m in out[:frames])
    for _ in range(max(args.warmup, 1)):
        dec.decode_batch_raw(blobs, uv.MEM_DEVICE)
    clocks = ClockSampler(local); clocks.start()
    dev_ms = 0.0; launches = 0; stage_acc = {}
    barrier()
    for _ in range(args.steps):
        ctx.flush_l2(); dec.decode_batch_raw(blobs, uv.MEM_DEVICE); st = ctx.stats(2)
        dev_ms += st["device_ms"]; launches += st["kernel_launches"]
        for k, v in st["stages"].items():
            stage_acc[k] = stage_acc.get(k, 0.0) + v
    barrier(); e2e_s = 0.0
    for _ in range(args.steps):
        ctx.flush_l2(); t1 = time.perf_counter(); dec.decode_batch_raw(blobs, uv.MEM_HOST); e2e_s += time.perf_counter() - t1
        st = ctx.stats(2); launches += st["kernel_launches"]
        e2e_parts = {"host_parse_and_staging": round(st["host_parse_ms"], 3), "call_total": round(st["total_ms"], 3), "kernels": round(st["device_ms"], 3),
                     "d2h": round(st["stages"].get("d2h", 0.0), 3), "h2d": round(st["stages"].get("h2d", 0.0), 3)}
    barrier(); clk = clocks.stop()
    if world > 1:
        v = torch.tensor([dev_ms, e2e_s], device="cuda", dtype=torch.float64); dist.all_reduce(v, op=dist.ReduceOp.MAX); dev_ms, e2e_s = float(v[0]), float(v[1])
    peak, peak_src = measured_peak()
    total = frames * world * args.steps
    stages = {k: {"ms": round(v / args.steps, 4)} for k, v in stage_acc.items()}
    kst = {k: v for k, v in stages.items() if k not in ("h2d", "d2h")}
    dom = max(kst, key=lambda k: kst[k]["ms"])
    # algorithmic bytes per step: compressed bytes in + index / position / uv out; the dominant stage's share: faces = clers in, index out
    # (faces = the walk plus the delta reversal that follows it inside the same kernel: clers in, index out, corrections in, values out)
    alg = {"faces": F_total * (1 + 12) + V_total * 5 * 8, "dequant": V_total * 5 * 8, "values": st["bytes_in"] + V_total * 5 * 4, "tunstall": st["bytes_in"]}
    ach = alg.get(dom, 0) / (kst[dom]["ms"] * 1e-3) / 1e9 if kst[dom]["ms"] > 0 else None
    traffic, traffic_src = None, None
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r02*_traffic*.json"))):       # measured DRAM bytes per launch (ncu --set full), scaled per frame
        try:
            tr = json.load(open(path))
            if tr.get("workload") == "c5" and tr.get(dom) and tr.get("frames"):
                traffic, traffic_src = tr[dom] * frames / tr["frames"], os.path.basename(path)
        except Exception:
            pass
    if rank == 0:
        cpu = None
        if world == 1:
            sample = crt[: max(ncores * 4, 32)]; cpu_pass(sample[:ncores], ncores); dt = cpu_pass(sample, ncores)
            cpu = {"value": len(sample) / dt, "unit": "frames/s", "cores": ncores, "kind": "reference",
                   "sample": f"first {len(sample)} frames, {ncores} threads, the reference's own crt::Decoder compiled from deprecated/encoder/dev/src (oracle/_ref/libcorto_ref.so)"}
        print(json.dumps({"metric": METRIC, "value": total / (dev_ms / 1e3), "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32+f32", "data": "synthetic",
                          "config": {"workload": workload_name, "frames_per_gpu": frames, "verts": int(V_total / frames), "faces": int(F_total / frames),
                                     "distinct_geometry_frames": W["distinct_geo"], "l2": "flushed between timed iterations (256 MiB memset)",
                                     "value_note": "device time from CUDA events around the kernels (upload of the .crt bytes outside, outputs left in HBM)",
                                     "parallelism": f"frames sharded, {world} rank(s), no data-path collective"},
                          "mverts_per_s": V_total * world * args.steps / (dev_ms / 1e3) / 1e6,
                          "roofline": {"bound": "hbm", "kernel": "corto_" + dom, "achieved": ach and round(ach, 2), "peak": peak, "unit": "GB/s", "frac": ach and round(ach / peak, 5),
                                       "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg.get(dom), "ms_per_launch": kst[dom]["ms"],
                                       "note": "dominant stage is a latency-bound serial walk (one warp per frame)"},
                          "stages": stages, "cpu_baseline": cpu,
                          "e2e": {"value": total / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": st["bytes_in"], "d2h_bytes_per_step": st["bytes_out"], "ms_per_step": e2e_s / args.steps * 1e3,
                                  "path": "uvol_decode_corto_batch, UVOL_MEM_HOST", "breakdown_ms_last_step_rank0": e2e_parts},
                          "gpu_launches": launches, "clocks": clk, "workload_gen_s": round(gen_s, 2)}))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--window-segments", type=int, default=0, help="decode the (shard of the) sequence in windows of this many segments instead of at once")
    ap.add_argument("--texture-target", default="rgba32", choices=sorted(TARGETS), help="output texture format (rgba32 = the parity target)")
    ap.add_argument("--no-gather", action="store_true", help="N>1: skip the NCCL gather of the decoded shards (then a step is the decode alone)")
    ap.add_argument("--no-weak", action="store_true", help="N>1: skip the weak-scaling side figure (every rank decoding a whole sequence of its own)")
    ap.add_argument("--no-extra-targets", action="store_true", help="skip the short side pass with the BC7 texture target")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU work of the cpu_baseline sample")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    W = dict(WORKLOADS[args.workload])
    if args.window_segments > 0:
        W["window_segments"] = args.window_segments
    frames, seq = W["frames"], W["seq"]
    ncores = os.cpu_count() or 1
    if W["fmt"] == "corto":
        return bench_v1(args, W, rank, world, local)
    # targets follow the reference's FORMAT_OPTIONS: ASTC is offered for UASTC sources only, ETC1 / ETC2 for ETC1S sources only
    if (args.texture_target == "astc" and W["fmt"] != "uastc") or (args.texture_target in ("etc1", "etc2") and W["fmt"] != "etc1s"):
        ap.error(f"--texture-target {args.texture_target} does not apply to the {W['fmt']} textures of workload {args.workload}")

    # ------------------------------------------------------------------ reference arm (CPU oracle)
    if args.impl == "reference":
        if rank != 0:
            return 0
        drc, ktx, info = make_workload(args.workload, 0)
        sd, sk = cpu_sample(drc, ktx, seq, ncores)                  # one segment per host thread (+ the frames they cover) per step
        for _ in range(max(1, min(args.warmup, 1))):
            cpu_oracle_run(sd[:seq], sk[:1], ncores)
        t0 = time.perf_counter(); pts = tx = 0
        for _ in range(args.steps):
            tg, tt, p, x = cpu_oracle_run(sd, sk, ncores); pts += p; tx += x
        dt = time.perf_counter() - t0
        nfr = len(sd) * args.steps; fps = nfr / dt
        line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "int32+f32",
                "data": "real fixtures" if args.workload == "liam" else "synthetic", "config": workload_config(args.workload, W, info, args.gpus, len(ktx), "rgba32"),
                "sample_frames_per_step": len(sd), "sample_segments_per_step": len(sk),
                "mverts_per_s": pts / dt / 1e6, "mtexels_per_s": tx / dt / 1e6,
                "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": ncores, "kind": "port",
                                 "sample": f"{len(sd)} frames + {len(sk)} segments of the workload per step, {ncores} threads, oracle/liboracle.so (CPU restatement of Draco 1.4.3 / Basis decode; upstream binaries unavailable)"},
                "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line)); return 0

    # ------------------------------------------------------------------ our arm
    rank_cpus = bind_to_gpu_cpus(local, world)
    import torch
    import torch.distributed as dist
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"                      # keep NCCL's version banner off stdout: the line printed there is the JSON result
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    uv = importlib.import_module("universal-volumetric_b200")
    from concurrent.futures import ThreadPoolExecutor
    drc_all, ktx_all, info = make_workload(args.workload, 0)            # ONE sequence, the same on every rank
    n_seg_all = len(ktx_all)
    f0, f1, s0, s1 = uv.shard_v2(frames, seq, n_seg_all, world, rank)   # this rank's block of whole segments and the frames they cover
    peak, peak_src = measured_peak()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def merge(acc, s):
        if acc is None:
            return {**s, "stages": dict(s["stages"])}
        for k in ("device_ms", "kernel_launches", "bytes_in", "bytes_out", "host_parse_ms", "h2d_ms", "d2h_ms", "total_ms"):
            acc[k] += s[k]
        acc["scratch_bytes"] = max(acc["scratch_bytes"], s["scratch_bytes"])
        for k, v in s["stages"].items():
            acc["stages"][k] = acc["stages"].get(k, 0.0) + v
        return acc

    class Runner:
        """The (shard of the) sequence resident on this GPU: one ctx per window (normally one window)."""

        def __init__(self, drc, ktx, target):
            self.windows = make_windows(drc, ktx, seq, W["window_segments"])
            self.ctxs = [uv.Context(local, profiling=True, texture_target=TARGETS[target]) for _ in self.windows]
            self.players = [uv.V2Player(c) for c in self.ctxs]
            self.pool = ThreadPoolExecutor(len(self.windows))
            self.last = [None] * len(self.windows)

        def _one(self, a):
            w, resident = a
            p, c, (wd, wk) = self.players[w], self.ctxs[w], self.windows[w]
            g, t = p.replay_step_raw(len(wd), len(wk), uv.MEM_DEVICE) if resident else p.decode_step_raw(wd, wk, uv.MEM_HOST)
            self.last[w] = (g, len(wd), t, len(wk))
            a_, b_ = c.stats(0, combined=True), c.stats(1, combined=True)
            bad = sum(x.status != 0 for x in g[:len(wd)]) + sum(x.status != 0 for x in t[:len(wk)])
            return a_, b_, (sum(x.num_points for x in g[:len(wd)]), sum(x.num_faces for x in g[:len(wd)]), sum(x.width * x.height * x.layers for x in t[:len(wk)]), bad)

        def step(self, resident):
            """One pass over the shard.  Returns the summed statistics, the device time spanned (first kernel to last kernel, CUDA events) and counts."""
            sg = st = None; tot = [0, 0, 0, 0]
            for a_, b_, cnt in self.pool.map(self._one, [(w, resident) for w in range(len(self.windows))]):
                sg = merge(sg, a_); st = merge(st, b_); tot = [x + y for x, y in zip(tot, cnt)]
            return sg, st, uv.span_ms(self.ctxs), tuple(tot)

        def close(self):
            for c in reversed(self.ctxs):
                c.close()

    def measure(runner, steps, warmup, with_gather):
        """warm-up, `steps` resident steps (device-timed; + gather), `steps` end-to-end steps (wall clock).  -> dict"""
        for _ in range(max(warmup, 1)):
            sg, st, _, (P_total, F_total, texels, bad) = runner.step(False)
        assert bad == 0, "decode failed on the bench workload"
        garena = [None]

        def gather_once():
            # the whole decoded shard (geometry and textures) of the last resident step, from the library's device buffers
            g, ng, t, nt = runner.last[-1] if len(runner.windows) == 1 else (None, 0, None, 0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            G = uv.gather.all_gather_shard(g, ng, t, nt, f"cuda:{local}", arena=garena[0])
            e1.record(); torch.cuda.synchronize()
            garena[0] = G["arena"]
            return e0.elapsed_time(e1), G
        for _ in range(max(warmup, 1)):
            runner.step(True)
            if with_gather:
                gather_once()
        dev_ms = gat_ms = 0.0; launches = 0; stage_acc = {}; G = None
        barrier(); t0 = time.perf_counter()
        for _ in range(steps):
            runner.ctxs[0].flush_l2()
            sg, st, dms, _ = runner.step(True)
            dev_ms += dms; launches += sg["kernel_launches"] + st["kernel_launches"]
            for k, v in list(sg["stages"].items()) + [("tex_" + k, v) for k, v in st["stages"].items()]:
                stage_acc[k] = stage_acc.get(k, 0.0) + v
            if with_gather:
                gm, G = gather_once(); gat_ms += gm
        barrier(); wall_resident = time.perf_counter() - t0
        e2e_s = 0.0
        barrier()
        for _ in range(steps):
            runner.ctxs[0].flush_l2()
            t1 = time.perf_counter(); sg2, st2, _, _ = runner.step(False); e2e_s += time.perf_counter() - t1
            launches += sg2["kernel_launches"] + st2["kernel_launches"]
        barrier()
        return dict(sg=sg, st=st, sg2=sg2, st2=st2, dev_ms=dev_ms, gat_ms=gat_ms, e2e_s=e2e_s, launches=launches, stage_acc=stage_acc, wall_resident=wall_resident,
                    P_total=P_total, F_total=F_total, texels=texels, G=G)

    def maxr(*vals):
        if world == 1:
            return list(vals)
        v = torch.tensor(list(vals), device="cuda", dtype=torch.float64); dist.all_reduce(v, op=dist.ReduceOp.MAX)
        return [float(x) for x in v]

    def sumr(*vals):
        if world == 1:
            return list(vals)
        v = torch.tensor(list(vals), device="cuda", dtype=torch.float64); dist.all_reduce(v, op=dist.ReduceOp.SUM)
        return [float(x) for x in v]

    # ---- headline: this rank's shard, resident + end to end (+ gather)
    with_gather = world > 1 and not args.no_gather and not W["window_segments"]
    runner = Runner(drc_all[f0:f1], ktx_all[s0:s1], args.texture_target)
    clocks = ClockSampler(local); clocks.start()
    M = measure(runner, args.steps, args.warmup, with_gather)
    clk = clocks.stop()
    gather_info = None
    if with_gather:
        G = M["G"]
        # every rank must now hold the same bytes: compare a checksum of the gathered arena across ranks
        ssum = G["arena"][: G["used"] // 8 * 8].view(torch.int64).sum().reshape(1)
        lo, hi = ssum.clone(), ssum.clone(); dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        gather_info = {"what": "whole decoded shard of every rank (geometry index / position / normal / uv + texture layers), gather.all_gather_shard: one in-place NCCL all-gather over equal slots of a torch-owned arena (own spans copied in from the library's device buffers)",
                       "bytes_received_per_gpu": int(sum(G["bytes"]) - G["bytes"][rank]), "bytes_total": int(sum(G["bytes"])), "all_ranks_identical": bool(torch.equal(lo, hi))}
    dev_ms, gat_ms, e2e_s = maxr(M["dev_ms"], M["gat_ms"], M["e2e_s"])
    P_all, F_all, tex_all, bin_g, bin_t, bout_g, bout_t, bout_g2, bout_t2, launches_all = sumr(M["P_total"], M["F_total"], M["texels"], M["sg"]["bytes_in"], M["st"]["bytes_in"], M["sg"]["bytes_out"], M["st"]["bytes_out"],
                                                                                               M["sg2"]["bytes_out"], M["st2"]["bytes_out"], M["launches"])
    step_ms = (dev_ms + gat_ms) / args.steps
    value = frames / (step_ms / 1e3); e2e = frames * args.steps / e2e_s
    if with_gather:
        gather_info["ms_per_step"] = gat_ms / args.steps
        gather_info["recv_gbs_per_gpu"] = gather_info["bytes_received_per_gpu"] / (gat_ms / args.steps * 1e-3) / 1e9 if gat_ms > 0 else None
    scratch_gb = (M["sg"]["scratch_bytes"] + M["st"]["scratch_bytes"]) / 1e9
    my_frames = f1 - f0

    # ---- roofline per stage (this rank's shard; ranks hold equal shares)
    sg, st, stage_acc = M["sg"], M["st"], M["stage_acc"]
    gb, tb = stage_bytes(info, M["P_total"], max(my_frames, 1), sg["bytes_in"], st["bytes_in"], W["fmt"], args.texture_target)
    stages = {}
    for k, ms in stage_acc.items():
        per = ms / args.steps
        if per < 0.02 and k not in ("h2d", "d2h"):
            continue                                                                   # stages that did not launch for this workload (e.g. the ETC1S entropy stages on UASTC input)
        nbytes = gb.get(k.replace("(s1)", "")) if not k.startswith("tex_") else tb.get(k[4:])
        stages[k] = {"ms": round(per, 4), "share": round(ms / sum(stage_acc.values()), 4)}
        if nbytes and per > 0:
            stages[k]["gbs"] = round(nbytes / (per * 1e-3) / 1e9, 2); stages[k]["frac"] = round(stages[k]["gbs"] / peak, 5)
    kernel_stages = {k: v for k, v in stages.items() if k not in ("h2d", "d2h", "tex_h2d", "tex_d2h")}
    main_stream = {k: v for k, v in kernel_stages.items() if "(s1)" not in k}          # side-stream spans include waiting for the main stream
    ksum = sum(v["ms"] for v in kernel_stages.values()) or 1.0
    for v in kernel_stages.values():
        v["share_of_kernel_time"] = round(v["ms"] / ksum, 4)          # comparable with the ncu launch-list shares
    traffic = None
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r02*_traffic*.json"))):        # measured DRAM bytes per launch (ncu --set full), newest last
        try:
            tr = json.load(open(path))
            if tr.get("workload") == args.workload:
                traffic = dict(tr, source=os.path.basename(path))
        except Exception:
            pass
    dom = max(main_stream, key=lambda k: main_stream[k]["ms"])
    dom_bytes = gb.get(dom.replace("(s1)", "")) if not dom.startswith("tex_") else tb.get(dom[4:])
    nwin = len(runner.windows)
    tr_dom = (traffic or {}).get(dom.replace("(s1)", ""))
    if tr_dom and traffic.get("frames"):
        tr_dom = tr_dom * my_frames / traffic["frames"]                                # the capture decodes fewer frames per launch than the bench: scaled per frame
    roof = {"bound": "hbm", "kernel": dom, "achieved": kernel_stages[dom].get("gbs"), "peak": peak, "unit": "GB/s", "frac": kernel_stages[dom].get("frac"),
            "traffic": tr_dom, "traffic_source": (traffic or {}).get("source"), "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes and dom_bytes / nwin,
            "ms_per_launch": kernel_stages[dom]["ms"] / nwin, "launches_per_step": nwin,
            "note": "dominant stage is a latency-bound serial walk (one warp per frame and table); the HBM-bound stages (expand, blocks, face_records) are listed in `stages`"}
    step_bytes = bin_g + bin_t + bout_g + bout_t
    pipeline = {"bytes_per_frame": step_bytes / frames, "achieved_gbs": step_bytes / (dev_ms / args.steps / 1e3) / 1e9}
    pipeline["frac"] = pipeline["achieved_gbs"] / (peak * world)
    windows_cfg = [len(wd) for wd, _ in runner.windows]
    runner.close(); del runner

    # ---- side figures
    extra = {}
    if args.texture_target == "rgba32" and not args.no_extra_targets and W["fmt"] in ("uastc", "etc1s"):
        try:       # the target the reference itself picks on desktop NVIDIA (BC7, KTX2Loader.js:602-604): 4x fewer texture bytes cross PCIe
            r2 = Runner(drc_all[f0:f1], ktx_all[s0:s1], "bc7")
            M2 = measure(r2, max(2, min(args.steps, 3)), 1, False)
            d2, e2 = maxr(M2["dev_ms"], M2["e2e_s"]); k2 = max(2, min(args.steps, 3))
            o2 = sumr(M2["sg2"]["bytes_out"] + M2["st2"]["bytes_out"])[0]
            extra["bc7_target"] = {"value": frames * k2 / (d2 / 1e3), "e2e": frames * k2 / e2, "unit": "frames/s", "steps": k2, "d2h_bytes_per_step": o2,
                                   "tex_stage_ms_rank0": {k: round(v / k2, 3) for k, v in M2["stage_acc"].items() if k.startswith("tex_") and v / k2 >= 0.02},
                                   "geo_kernels_ms_rank0": M2["sg"]["device_ms"], "e2e_breakdown_rank0": {"geo_call_total": M2["sg2"]["total_ms"], "geo_d2h": M2["sg2"]["d2h_ms"], "tex_kernels": M2["st2"]["device_ms"], "tex_d2h": M2["st2"]["d2h_ms"]},
                                   "note": "same sequence, UVOL_TEX_BC7 output (decode only, no gather); BC7 blocks are validated by an independent BC7 decoder, not bit-matched to basisu (DESIGN.md)"}
            r2.close(); del r2
        except Exception as ex:                                       # never lose the headline to a side figure
            extra["bc7_target"] = {"error": str(ex)[:200]}
    if args.texture_target == "rgba32" and not args.no_extra_targets and W["fmt"] == "uastc":
        try:       # the reference's first choice for UASTC on GPUs with ASTC support (KTX2Loader.js:592-600): a LOSSLESS repack, same texels, 4x fewer bytes
            r4 = Runner(drc_all[f0:f1], ktx_all[s0:s1], "astc")
            k4 = max(2, min(args.steps, 3))
            M4 = measure(r4, k4, 1, False)
            d4, e4 = maxr(M4["dev_ms"], M4["e2e_s"])
            o4 = sumr(M4["sg2"]["bytes_out"] + M4["st2"]["bytes_out"])[0]
            extra["astc_target"] = {"value": frames * k4 / (d4 / 1e3), "e2e": frames * k4 / e4, "unit": "frames/s", "steps": k4, "d2h_bytes_per_step": o4,
                                    "tex_stage_ms_rank0": {k: round(v / k4, 3) for k, v in M4["stage_acc"].items() if k.startswith("tex_") and v / k4 >= 0.02},
                                    "note": "same sequence, UVOL_TEX_ASTC_4x4 output (decode only, no gather); lossless: the blocks decode to exactly the RGBA32 texels (tests/test_astc.py)"}
            r4.close(); del r4
        except Exception as ex:
            extra["astc_target"] = {"error": str(ex)[:200]}
    if world > 1 and not args.no_weak:
        try:       # last round's weak-scaling figure: every rank decodes a whole sequence of its own
            dw, kw, _ = make_workload(args.workload, rank)
            r3 = Runner(dw, kw, args.texture_target)
            M3 = measure(r3, 2, 1, False)
            d3, e3 = maxr(M3["dev_ms"], M3["e2e_s"])
            extra["weak"] = {"value": frames * world * 2 / (d3 / 1e3), "e2e": frames * world * 2 / e3, "unit": "frames/s", "steps": 2, "frames_per_gpu": frames,
                             "note": "every rank decodes its own whole sequence, no gather (round-1 definition)"}
            r3.close(); del r3
        except Exception as ex:
            extra["weak"] = {"error": str(ex)[:200]}

    if rank == 0:
        # ---- cpu baseline on a bounded sample (rank 0, N=1 only)
        cpu = None
        if world == 1:
            nseg = min(ncores, n_seg_all)
            sd, sk = cpu_sample(drc_all, ktx_all, seq, nseg)
            tg, tt, p, x = cpu_oracle_run(sd, sk, ncores)          # probe
            per_seg = max(tg + tt, 1e-3) / nseg
            nseg = int(max(min(ncores, n_seg_all), min(n_seg_all, args.cpu_seconds / per_seg)))
            sd, sk = cpu_sample(drc_all, ktx_all, seq, nseg)
            tg, tt, p, x = cpu_oracle_run(sd, sk, ncores)
            cpu = {"value": len(sd) / (tg + tt), "unit": "frames/s", "cores": ncores, "kind": "port",
                   "sample": f"first {len(sd)} frames + {len(sk)} segments of the workload, {ncores} threads (thread pool over frames/segments), oracle/liboracle.so",
                   "geometry_ms_per_frame_per_core": tg / len(sd) * 1e3 * min(ncores, len(sd)), "texture_ms_per_frame_per_core": tt / len(sd) * 1e3 * min(ncores, len(sk))}
        # scaling efficiency against this repo's own committed N=1 line of the same workload, when there is one (the driver computes its own)
        eff = None
        try:
            n1 = json.load(open(os.path.join(ROOT, "profiles", f"r02_bench_{args.workload}.json")))
            if world > 1 and n1.get("n_gpus") == 1:
                eff = {"value": value / (world * n1["value"]), "e2e": e2e / (world * n1["e2e"]["value"]), "against": f"profiles/r02_bench_{args.workload}.json (N=1: {n1['value']:.0f} / {n1['e2e']['value']:.0f} frames/s)"}
        except Exception:
            pass
        line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "int32+f32",
                "data": "real fixtures" if args.workload == "liam" else "synthetic",
                "config": workload_config(args.workload, W, info, world, n_seg_all, args.texture_target),
                "run": {"frames_this_rank": my_frames, "segments_this_rank": s1 - s0, "points_per_frame": P_all / frames,
                        "windows": windows_cfg, "scratch_gb": round(scratch_gb, 2), "scratch_mb_per_frame": round(scratch_gb * 1e3 / max(my_frames, 1), 1),
                        "decode_ms_per_step": dev_ms / args.steps, "gather_ms_per_step": gat_ms / args.steps if with_gather else None,
                        "value_note": "inputs resident in HBM; CUDA events from the first to the last kernel of the step (max over ranks), plus the gather (CUDA events) when N > 1",
                        "host": {"cpus_of_rank0": rank_cpus, "staging_threads": int(os.environ.get("UVOL_STAGING_THREADS", "0"))}},
                "mverts_per_s": P_all / (dev_ms / args.steps / 1e3) / 1e6, "mtexels_per_s": tex_all / (dev_ms / args.steps / 1e3) / 1e6,
                "roofline": roof, "pipeline_roofline": pipeline, "stages": stages,
                "cpu_baseline": cpu,
                "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": bin_g + bin_t, "d2h_bytes_per_step": bout_g2 + bout_t2, "ms_per_step": e2e_s / args.steps * 1e3,
                        "path": "uvol_decode_v2_batch on every rank's shard (geometry and texture streams concurrent), host buffers in, pinned host buffers out (UVOL_MEM_HOST); wall clock, max over ranks",
                        "breakdown_ms_per_step_rank0": {"geo_host_parse": M["sg2"]["host_parse_ms"], "geo_h2d": M["sg2"]["h2d_ms"], "geo_kernels": M["sg2"]["device_ms"], "geo_d2h": M["sg2"]["d2h_ms"],
                                                        "geo_call_total": M["sg2"]["total_ms"], "tex_host_parse": M["st2"]["host_parse_ms"], "tex_h2d": M["st2"]["h2d_ms"], "tex_kernels": M["st2"]["device_ms"], "tex_d2h": M["st2"]["d2h_ms"]}},
                "gather": gather_info, "scaling_efficiency": eff, "gpu_launches": int(launches_all), "clocks": clk,
                "wall_ms_per_step_resident": M["wall_resident"] / args.steps * 1e3, "workload_gen_s": info["gen_s"]}
        line.update(extra)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python
"""bench.py -- decoded frames/sec (geometry + texture) of the UVOL V2 decode hot path.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W [--impl reference]`
prints ONE JSON line on rank 0.

Workload at N=1 (default `--workload c3`): the configuration BASELINE.json's metric is quoted on, configs[2] -- a
synthetic 1000-frame V2 sequence, 200k verts/frame Draco geometry, 2048^2 UASTC KTX2 textures with sequenceSize 7
(143 segments), built by tools/synth (deterministic).  `--workload c2` is configs[1] (300 frames, 50k verts, 1024^2
ETC1S).  One "step" = one pass of the hot path over the whole sequence.  A sequence whose scratch does not fit HBM at
once is decoded in WINDOWS of whole segments (the prefetch window of src/V2/player.ts:272-323); every window's
compressed inputs stay resident in HBM (one ctx per window), the scratch / output arenas exist once (uvol_share_arenas).
N>1: every rank decodes its own sequence (frames are independent units, no data-path collective) -> "scaling": "weak";
value = all frames / max-over-ranks time.

  value     frames/s with the compressed inputs already resident in HBM (uvol_replay_v2_batch), device
            time from CUDA events on the library's streams (geometry and texture run concurrently, the
            step time is the longer of the two spans), outputs left in HBM.
  e2e       frames/s through the C ABI with HOST buffers: uvol_decode_draco_batch +
            uvol_transcode_ktx2_batch with UVOL_MEM_HOST (host parse, H2D of the compressed bytes,
            kernels, D2H of every decoded buffer into pinned host memory), wall clock around the calls.
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
    "c3": dict(frames=1000, verts=200000, tex=2048, seq=7, seed=20260003, fmt="uastc", window_segments=72, distinct_geo=32, distinct_tex=6,
               label="configs[2]: 1000-frame V2 seq, 200k verts/frame Draco, 2048^2 UASTC KTX2 batch=7"),
    "c2": dict(frames=300, verts=50000, tex=1024, seq=7, seed=20260002, fmt="etc1s", window_segments=None, distinct_geo=None, distinct_tex=None,
               label="configs[1]: 300-frame V2 seq, 50k verts/frame Draco, 1024^2 ETC1S KTX2 batch=7"),
    "c5": dict(frames=300, verts=50000, tex=1024, seq=1, seed=20260005, fmt="corto", window_segments=None, distinct_geo=8, distinct_tex=None,
               label="configs[4]: V1 manifest, 300-frame Corto .crt geometry (position 12 bit + uv 12 bit, u32 index); V1's mp4 texture leg is out of scope"),
    "tiny": dict(frames=28, verts=2000, tex=64, seq=7, seed=20260009, fmt="uastc", window_segments=2, distinct_geo=None, distinct_tex=None,
                 label="tiny smoke workload"),
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


def make_workload(name, rank):
    from tools.synth import synth
    w = WORKLOADS[name]
    t0 = time.time()
    drc, ktx, info = synth.make_sequence(w["frames"], w["verts"], w["tex"], sequence_size=w["seq"], seed=w["seed"] + 1000 * rank,
                                         distinct_geometry=w["distinct_geo"], distinct_textures=w["distinct_tex"], texture_format=w["fmt"])
    info["gen_s"] = round(time.time() - t0, 2)
    return drc, ktx, info


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


def stage_bytes(info, P_total, frames, bytes_in_geo, bytes_in_tex, fmt):
    """ALGORITHMIC bytes per step for each kernel stage (DESIGN.md 'Kernels and rooflines')."""
    F, V = info["faces"], info["verts"]
    nblk = (info["tex_size"] // 4) ** 2
    P = P_total / frames
    g = {
        "edgebreaker": frames * (F * 1 + 2 * 3 * F * 4),                       # symbols in, corner table (opp + c2v) out
        "traverse": frames * 3 * (3 * F * 4 + 2 * V * 4),                     # per table: corner table in, entry maps out
        "rans_attr": bytes_in_geo + frames * (3 * V + 2 * P + 2 * V) * 4,      # compressed in, int32 corrections out
        "predict_wrap": frames * (V * 16 + 2 * V * 12),                       # parents + corrections in, values out
        "predict_uv": frames * P * (32 + 8 + 8),
        "normals": frames * V * (8 + 8 + 7 * 12),
        "expand": frames * (P * 4 + P * 32 + P * 32),                          # p2c + gathers in, 32 B/point out
        "seams": frames * (3 * F * 4 + 2 * 3 * F),
        "attr_tables": frames * 2 * (3 * F * 4 + 3 * F * 4),
        "point_assign": frames * (3 * F * 4 * 3),
    }
    t = {
        "slices": bytes_in_tex + frames * nblk * 5,                            # VLC bits in, {pred u8, delta/selector u16} out
        "resolve": frames * nblk * (1 + 2 + 2 + 2),
        "blocks": frames * nblk * ((16 if fmt == "uastc" else 4) + 64),          # UASTC: 16 B block in; ETC1S: 2x u16 indices in; 64 B RGBA out
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
    barrier(); clk = clocks.stop()
    if world > 1:
        v = torch.tensor([dev_ms, e2e_s], device="cuda", dtype=torch.float64); dist.all_reduce(v, op=dist.ReduceOp.MAX); dev_ms, e2e_s = float(v[0]), float(v[1])
    peak, peak_src = measured_peak()
    total = frames * world * args.steps
    stages = {k: {"ms": round(v / args.steps, 4)} for k, v in stage_acc.items()}
    kst = {k: v for k, v in stages.items() if k not in ("h2d", "d2h")}
    dom = max(kst, key=lambda k: kst[k]["ms"])
    # algorithmic bytes per step: compressed bytes in + index / position / uv out; the dominant stage's share: faces = clers in, index out
    alg = {"faces": F_total * (1 + 12), "dequant": V_total * 5 * 8, "delta": V_total * 5 * 8, "values": st["bytes_in"] + V_total * 5 * 4, "tunstall": st["bytes_in"]}
    ach = alg.get(dom, 0) / (kst[dom]["ms"] * 1e-3) / 1e9 if kst[dom]["ms"] > 0 else None
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
                                       "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg.get(dom), "ms_per_launch": kst[dom]["ms"],
                                       "note": "dominant stage is a latency-bound serial walk (one warp per frame)"},
                          "stages": stages, "cpu_baseline": cpu,
                          "e2e": {"value": total / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": st["bytes_in"], "d2h_bytes_per_step": st["bytes_out"], "ms_per_step": e2e_s / args.steps * 1e3,
                                  "path": "uvol_decode_corto_batch, UVOL_MEM_HOST"},
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
    ap.add_argument("--window-segments", type=int, default=0, help="override the workload's window size (segments per library call)")
    ap.add_argument("--gather", action="store_true", help="N>1: also time the optional final NCCL all-gather of decoded geometry (first <=64 frames of the last window per rank)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU work of the cpu_baseline sample")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    W = dict(WORKLOADS[args.workload])
    if args.window_segments > 0:
        W["window_segments"] = args.window_segments
    # Pinned host memory of the e2e path: every window ctx holds its own result buffers (C3: 28 GB of results + 5 GB of staged inputs
    # per rank).  If the host cannot hold that for every rank, the windows share one set of result buffers and the e2e pass runs
    # them one after the other (the resident pass keeps running them concurrently).
    low_host_memory = False
    try:
        import psutil
        low_host_memory = W["fmt"] == "uastc" and W["frames"] >= 500 and psutil.virtual_memory().available / max(1, world) < 48e9
    except Exception:
        pass
    if os.environ.get("UVOL_BENCH_LOW_HOST_MEMORY"):
        low_host_memory = os.environ["UVOL_BENCH_LOW_HOST_MEMORY"] == "1"
    frames, verts, tex, seq, seed = W["frames"], W["verts"], W["tex"], W["seq"], W["seed"]
    ncores = os.cpu_count() or 1
    if W["fmt"] == "corto":
        return bench_v1(args, W, rank, world, local)
    workload_name = f"{W['label']} (synthetic, tools/synth seed {seed})"

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
                "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32+f32",
                "data": "synthetic", "config": {"workload": workload_name, "frames_per_step": len(sd), "segments_per_step": len(sk)},
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
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    uv = importlib.import_module("universal-volumetric_b200")
    drc, ktx, info = make_workload(args.workload, rank)
    windows = make_windows(drc, ktx, seq, W["window_segments"])
    ctxs = [uv.Context(local, profiling=True) for _ in windows]
    for c in ctxs[1:]:
        c.share_arenas(ctxs[0])
        if low_host_memory:
            c.share_host_outputs(ctxs[0])
    players = [uv.V2Player(c) for c in ctxs]
    ctx = ctxs[0]
    n_k = len(ktx)

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

    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(len(windows))

    def one_window(args_):
        w, resident = args_
        p, c, (wd, wk) = players[w], ctxs[w], windows[w]
        g, t = p.replay_step_raw(len(wd), len(wk), uv.MEM_DEVICE) if resident else p.decode_step_raw(wd, wk, uv.MEM_HOST)
        a, b = c.stats(0, combined=True), c.stats(1, combined=True)
        bad = sum(x.status != 0 for x in g[:len(wd)]) + sum(x.status != 0 for x in t[:len(wk)])
        return a, b, (sum(x.num_points for x in g[:len(wd)]), sum(x.num_faces for x in g[:len(wd)]), sum(x.width * x.height * x.layers for x in t[:len(wk)]), bad)

    def run_step(resident):
        """One pass over the sequence: every window is driven by its own host thread (ctypes releases the GIL); the library hands the
        shared phase-2 scratch from window to window, so the other windows' phase 1 and result copies overlap it.  Returns the summed
        statistics, the device time spanned by the whole step (first kernel to last kernel, CUDA events) and counts."""
        sg = st = None; tot = [0, 0, 0, 0]
        jobs_ = [(w, resident) for w in range(len(windows))]
        for a, b, cnt in (pool.map(one_window, jobs_) if resident or not low_host_memory else map(one_window, jobs_)):
            sg = merge(sg, a); st = merge(st, b); tot = [x + y for x, y in zip(tot, cnt)]
        return sg, st, uv.span_ms(ctxs), tuple(tot)

    # warm-up (also uploads the batch that the resident steps replay)
    for _ in range(max(args.warmup, 1)):
        sg, st, _, (P_total, F_total, texels, bad) = run_step(False)
    assert bad == 0, "decode failed on the bench workload"
    for _ in range(max(args.warmup, 1)):
        run_step(True)

    clocks = ClockSampler(local); clocks.start()
    # ---- timed: resident inputs (value)
    dev_ms = 0.0; launches = 0; stage_acc = {}
    barrier(); t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.flush_l2()
        sg, st, dms, _ = run_step(True)
        dev_ms += dms; launches += sg["kernel_launches"] + st["kernel_launches"]
        for k, v in list(sg["stages"].items()) + [("tex_" + k, v) for k, v in st["stages"].items()]:
            stage_acc[k] = stage_acc.get(k, 0.0) + v
    barrier(); wall_resident = time.perf_counter() - t0
    nwin = len(windows)
    # ---- timed: end to end through the C ABI with host buffers (e2e)
    e2e_s = 0.0
    barrier()
    for _ in range(args.steps):
        ctx.flush_l2()
        t1 = time.perf_counter(); sg, st, _, _ = run_step(False); e2e_s += time.perf_counter() - t1
        launches_e2e = sg["kernel_launches"] + st["kernel_launches"]
    barrier()
    clk = clocks.stop()
    h2d = sg["bytes_in"] + st["bytes_in"]; d2h = sg["bytes_out"] + st["bytes_out"]
    gather_info = None
    if world > 1 and args.gather:
        # optional final gather: not part of the metric; decoded buffers of the last window go device -> device over NCCL
        wd, wk = windows[-1]
        g, t = players[-1].replay_step_raw(len(wd), len(wk), uv.MEM_DEVICE)
        ng = min(len(wd), 64)
        uv.gather.all_gather_geometry(g, ng, f"cuda:{local}")                                  # warm-up (communicator set-up)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier(); e0.record()
        tables, arenas = uv.gather.all_gather_geometry(g, ng, f"cuda:{local}")
        e1.record(); torch.cuda.synchronize()
        _, runs = uv.gather.geometry_table(g, ng)
        mine = uv.gather.pack_runs(runs, f"cuda:{local}"); nbytes = int(mine.numel())
        same = bool(torch.equal(arenas[rank][:nbytes], mine))
        sums = arenas[:, : arenas.shape[1] // 8 * 8].view(torch.int64).sum(dim=1)                 # every rank must hold identical copies
        lo, hi = sums.clone(), sums.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        gms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64); dist.all_reduce(gms, op=dist.ReduceOp.MAX)
        gather_info = {"frames_per_rank": ng, "bytes_per_rank": int(nbytes), "ms": float(gms[0]), "recv_gbs_per_gpu": nbytes * (world - 1) / (float(gms[0]) * 1e-3) / 1e9,
                       "own_slot_identical": same, "all_ranks_identical": bool(torch.equal(lo, hi)), "backend": "nccl all_gather_into_tensor on the library's device buffers"}
    # max over ranks
    if world > 1:
        v = torch.tensor([dev_ms, e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(v, op=dist.ReduceOp.MAX); dev_ms, e2e_s = float(v[0]), float(v[1])
    total_frames = frames * world * args.steps
    value = total_frames / (dev_ms / 1e3); e2e = total_frames / e2e_s

    # ---- roofline per stage
    peak, peak_src = measured_peak()
    gb, tb = stage_bytes(info, P_total, frames, sg["bytes_in"], st["bytes_in"], W["fmt"])
    stages = {}
    for k, ms in stage_acc.items():
        nbytes = gb.get(k.replace("(s1)", "")) if not k.startswith("tex_") else tb.get(k[4:])
        per = ms / args.steps
        stages[k] = {"ms": round(per, 4), "share": round(ms / sum(stage_acc.values()), 4)}
        if nbytes and per > 0:
            stages[k]["gbs"] = round(nbytes / (per * 1e-3) / 1e9, 2); stages[k]["frac"] = round(stages[k]["gbs"] / peak, 5)
    kernel_stages = {k: v for k, v in stages.items() if k not in ("h2d", "d2h", "tex_h2d", "tex_d2h", "counts_readback")}
    main_stream = {k: v for k, v in kernel_stages.items() if "(s1)" not in k}          # side-stream spans include waiting for the main stream
    ksum = sum(v["ms"] for v in kernel_stages.values()) or 1.0
    for v in kernel_stages.values():
        v["share_of_kernel_time"] = round(v["ms"] / ksum, 4)          # comparable with the ncu launch-list shares
    traffic = None
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json"))):        # measured DRAM bytes per launch (ncu --set full), newest last
        try:
            tr = json.load(open(path))
            if tr.get("workload") == args.workload:
                traffic = dict(tr, source=os.path.basename(path))
        except Exception:
            pass
    dom = max(main_stream, key=lambda k: main_stream[k]["ms"])
    dom_bytes = gb.get(dom.replace("(s1)", "")) if not dom.startswith("tex_") else tb.get(dom[4:])
    roof = {"bound": "hbm", "kernel": dom, "achieved": kernel_stages[dom].get("gbs"), "peak": peak, "unit": "GB/s", "frac": kernel_stages[dom].get("frac"),
            "traffic": (traffic or {}).get(dom.replace("(s1)", "")), "traffic_source": (traffic or {}).get("source"), "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes and dom_bytes / nwin, "ms_per_launch": kernel_stages[dom]["ms"] / nwin, "launches_per_step": nwin,
            "note": "dominant stage is a latency-bound serial walk (one warp per frame); HBM-bound stages are listed in `stages`"}
    step_bytes = sg["bytes_in"] + st["bytes_in"] + sg["bytes_out"] + st["bytes_out"]
    pipeline = {"bytes_per_frame": step_bytes / frames, "achieved_gbs": step_bytes * world * args.steps / (dev_ms / 1e3) / 1e9}
    pipeline["frac"] = pipeline["achieved_gbs"] / (peak * world)

    line = None
    if rank == 0:
        # ---- cpu baseline on a bounded sample (rank 0, N=1 only)
        cpu = None
        if world == 1:
            nseg = min(ncores, len(ktx))
            sd, sk = cpu_sample(drc, ktx, seq, nseg)
            tg, tt, p, x = cpu_oracle_run(sd, sk, ncores)          # probe
            per_seg = max(tg + tt, 1e-3) / nseg
            nseg = int(max(min(ncores, len(ktx)), min(len(ktx), args.cpu_seconds / per_seg)))
            sd, sk = cpu_sample(drc, ktx, seq, nseg)
            tg, tt, p, x = cpu_oracle_run(sd, sk, ncores)
            cpu = {"value": len(sd) / (tg + tt), "unit": "frames/s", "cores": ncores, "kind": "port",
                   "sample": f"first {len(sd)} frames + {len(sk)} segments of the workload, {ncores} threads (thread pool over frames/segments), oracle/liboracle.so",
                   "geometry_ms_per_frame_per_core": tg / len(sd) * 1e3 * min(ncores, len(sd)), "texture_ms_per_frame_per_core": tt / len(sd) * 1e3 * min(ncores, len(sk))}
        line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32+f32",
                "data": "synthetic",
                "config": {"workload": workload_name, "frames_per_gpu": frames, "segments_per_gpu": n_k, "verts": info["verts"], "faces": info["faces"],
                           "points_per_frame": P_total / frames, "distinct_geometry_frames": info["distinct_geometry"], "distinct_texture_segments": info["distinct_textures"],
                           "windows": [len(wd) for wd, _ in windows], "windows_concurrent": len(windows) > 1, "e2e_windows_sequential_low_host_memory": low_host_memory,
                           "scratch_gb_largest_window": round((sg["scratch_bytes"] + st["scratch_bytes"]) / 1e9, 2),
                           "l2": "flushed between timed iterations (256 MiB memset); every window's working set is far larger than L2",
                           "parallelism": f"frames sharded, {world} rank(s), no data-path collective",
                           "host": {"cpus_of_rank0": rank_cpus, "staging_threads": int(os.environ.get("UVOL_STAGING_THREADS", "0"))}},
                "mverts_per_s": P_total * world * args.steps / (dev_ms / 1e3) / 1e6, "mtexels_per_s": texels * world * args.steps / (dev_ms / 1e3) / 1e6,
                "roofline": roof, "pipeline_roofline": pipeline, "stages": stages,
                "cpu_baseline": cpu,
                "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s / args.steps * 1e3,
                        "path": "uvol_decode_v2_batch per window (geometry and texture streams concurrent), UVOL_MEM_HOST",
                        "breakdown_ms_per_step": {"geo_host_parse": sg["host_parse_ms"], "geo_h2d": sg["h2d_ms"], "geo_kernels": sg["device_ms"], "geo_d2h": sg["d2h_ms"],
                                                  "geo_call_total": sg["total_ms"], "tex_host_parse": st["host_parse_ms"], "tex_h2d": st["h2d_ms"], "tex_kernels": st["device_ms"], "tex_d2h": st["d2h_ms"]}},
                "gather": gather_info, "gpu_launches": launches + launches_e2e * args.steps, "clocks": clk,
                "wall_ms_per_step_resident": wall_resident / args.steps * 1e3, "workload_gen_s": info["gen_s"]}
        print(json.dumps(line))
    for c in reversed(ctxs):
        c.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

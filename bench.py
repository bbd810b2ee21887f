#!/usr/bin/env python
"""bench.py -- decoded frames/sec (geometry + texture) of the UVOL V2 decode hot path.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W [--impl reference]`
prints ONE JSON line on rank 0.

Workload at N=1: BASELINE.json configs[1] -- a synthetic 300-frame V2 sequence, 50k verts/frame
Draco geometry, 1024^2 ETC1S KTX2 textures with sequenceSize 7 (43 segments), built by
tools/synth (deterministic, seed 20260002).  One "step" = one pass of the hot path over the whole
sequence (300 .drc + 43 .ktx2).  N>1: every rank decodes its own 300-frame sequence (frames are
independent units, no data-path collective) -> "scaling": "weak"; value = all frames / max-over-ranks time.

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

METRIC = "decoded frames/sec (geom+tex)"
WORKLOADS = {
    # name: (frames, verts, tex_size, sequence_size, seed)
    "c2": (300, 50000, 1024, 7, 20260002),
    "tiny": (14, 2000, 64, 7, 20260009),
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


def make_workload(name, rank):
    from tools.synth import synth
    frames, verts, tex, seq, seed = WORKLOADS[name]
    t0 = time.time()
    drc, ktx, info = synth.make_sequence(frames, verts, tex, sequence_size=seq, seed=seed + 1000 * rank)
    info["gen_s"] = round(time.time() - t0, 2)
    return drc, ktx, info


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
    """Bounded sample of the same workload: nseg segments and the geometry frames they cover."""
    nseg = min(nseg, len(ktx))
    return drc[: nseg * seq], ktx[:nseg]


def stage_bytes(info, P_total, frames, bytes_in_geo, bytes_in_tex):
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
        "blocks": frames * nblk * (4 + 64),                                    # 2x u16 indices in, 64 B RGBA out
    }
    return g, t


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU work of the cpu_baseline sample")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    frames, verts, tex, seq, seed = WORKLOADS[args.workload]
    ncores = os.cpu_count() or 1
    workload_name = f"configs[1]: {frames}-frame V2 seq, {verts // 1000}k verts/frame Draco, {tex}^2 ETC1S KTX2 batch={seq} (synthetic, tools/synth seed {seed})"

    # ------------------------------------------------------------------ reference arm (CPU oracle)
    if args.impl == "reference":
        if rank != 0:
            return 0
        drc, ktx, info = make_workload(args.workload, 0)
        sd, sk = cpu_sample(drc, ktx, seq, 6)                       # 42 frames + 6 segments per step
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
    import torch
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    uv = importlib.import_module("universal-volumetric_b200")
    drc, ktx, info = make_workload(args.workload, rank)
    ctx = uv.Context(local, profiling=True)
    player = uv.V2Player(ctx)
    n_d, n_k = len(drc), len(ktx)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def e2e_step():
        g, t = player.decode_step_raw(drc, ktx, uv.MEM_HOST)
        return g, t, ctx.stats(0, combined=True), ctx.stats(1, combined=True)

    def resident_step():
        g, t = player.replay_step_raw(n_d, n_k, uv.MEM_DEVICE)
        return g, t, ctx.stats(0, combined=True), ctx.stats(1, combined=True)

    # warm-up (also uploads the batch that the resident steps replay)
    for _ in range(max(args.warmup, 1)):
        g, t, sg, st = e2e_step()
    assert all(x.status == 0 for x in g) and all(x.status == 0 for x in t), "decode failed on the bench workload"
    P_total = sum(x.num_points for x in g); F_total = sum(x.num_faces for x in g)
    texels = sum(x.width * x.height * x.layers for x in t)
    for _ in range(max(args.warmup, 1)):
        resident_step()

    clocks = ClockSampler(local); clocks.start()
    # ---- timed: resident inputs (value)
    dev_ms = 0.0; launches = 0; stage_acc = {}
    barrier(); t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.flush_l2()
        g, t, sg, st = resident_step()
        dev_ms += max(sg["device_ms"], st["device_ms"]); launches += sg["kernel_launches"] + st["kernel_launches"]
        for k, v in list(sg["stages"].items()) + [("tex_" + k, v) for k, v in st["stages"].items()]:
            stage_acc[k] = stage_acc.get(k, 0.0) + v
    barrier(); wall_resident = time.perf_counter() - t0
    # ---- timed: end to end through the C ABI with host buffers (e2e)
    e2e_s = 0.0
    barrier()
    for _ in range(args.steps):
        ctx.flush_l2()
        t1 = time.perf_counter(); g, t, sg, st = e2e_step(); e2e_s += time.perf_counter() - t1
        launches_e2e = sg["kernel_launches"] + st["kernel_launches"]
    barrier()
    clk = clocks.stop()
    h2d = sg["bytes_in"] + st["bytes_in"]; d2h = sg["bytes_out"] + st["bytes_out"]
    # max over ranks
    if world > 1:
        v = torch.tensor([dev_ms, e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(v, op=dist.ReduceOp.MAX); dev_ms, e2e_s = float(v[0]), float(v[1])
    total_frames = frames * world * args.steps
    value = total_frames / (dev_ms / 1e3); e2e = total_frames / e2e_s

    # ---- roofline per stage
    peak, peak_src = measured_peak()
    gb, tb = stage_bytes(info, P_total, frames, sg["bytes_in"], st["bytes_in"])
    stages = {}
    for k, ms in stage_acc.items():
        nbytes = gb.get(k.replace("(s1)", "")) if not k.startswith("tex_") else tb.get(k[4:])
        per = ms / args.steps
        stages[k] = {"ms": round(per, 4), "share": round(ms / (dev_ms if world == 1 else sum(stage_acc.values())), 4)}
        if nbytes and per > 0:
            stages[k]["gbs"] = round(nbytes / (per * 1e-3) / 1e9, 2); stages[k]["frac"] = round(stages[k]["gbs"] / peak, 5)
    kernel_stages = {k: v for k, v in stages.items() if k not in ("h2d", "d2h", "tex_h2d", "tex_d2h", "counts_readback")}
    ksum = sum(v["ms"] for v in kernel_stages.values()) or 1.0
    for v in kernel_stages.values():
        v["share_of_kernel_time"] = round(v["ms"] / ksum, 4)          # comparable with the ncu launch-list shares
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        if tr.get("workload") == args.workload:
            traffic = tr
    except Exception:
        pass
    dom = max(kernel_stages, key=lambda k: kernel_stages[k]["ms"])
    dom_bytes = gb.get(dom.replace("(s1)", "")) if not dom.startswith("tex_") else tb.get(dom[4:])
    roof = {"bound": "hbm", "kernel": dom, "achieved": kernel_stages[dom].get("gbs"), "peak": peak, "unit": "GB/s", "frac": kernel_stages[dom].get("frac"),
            "traffic": (traffic or {}).get(dom.replace("(s1)", "")), "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes, "ms_per_launch": kernel_stages[dom]["ms"],
            "note": "dominant stage is a latency-bound serial walk (one warp per frame); HBM-bound stages are listed in `stages`"}
    step_bytes = sg["bytes_in"] + st["bytes_in"] + sg["bytes_out"] + st["bytes_out"]
    pipeline = {"bytes_per_frame": step_bytes / frames, "achieved_gbs": step_bytes * world * args.steps / (dev_ms / 1e3) / 1e9}
    pipeline["frac"] = pipeline["achieved_gbs"] / (peak * world)

    line = None
    if rank == 0:
        # ---- cpu baseline on a bounded sample (rank 0, N=1 only)
        cpu = None
        if world == 1:
            nseg = 2
            sd, sk = cpu_sample(drc, ktx, seq, nseg)
            tg, tt, p, x = cpu_oracle_run(sd, sk, ncores)          # probe
            per_seg = max(tg + tt, 1e-3) / nseg
            nseg = int(max(2, min(len(ktx), args.cpu_seconds / per_seg)))
            sd, sk = cpu_sample(drc, ktx, seq, nseg)
            tg, tt, p, x = cpu_oracle_run(sd, sk, ncores)
            cpu = {"value": len(sd) / (tg + tt), "unit": "frames/s", "cores": ncores, "kind": "port",
                   "sample": f"first {len(sd)} frames + {len(sk)} segments of the workload, {ncores} threads (thread pool over frames/segments), oracle/liboracle.so",
                   "geometry_ms_per_frame_per_core": tg / len(sd) * 1e3 * min(ncores, len(sd)), "texture_ms_per_frame_per_core": tt / len(sd) * 1e3 * min(ncores, len(sk))}
        line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32+f32",
                "data": "synthetic",
                "config": {"workload": workload_name, "frames_per_gpu": frames, "segments_per_gpu": n_k, "verts": info["verts"], "faces": info["faces"],
                           "points_per_frame": P_total / frames, "distinct_geometry_frames": info["distinct_geometry"], "l2": "flushed between timed iterations (256 MiB memset)",
                           "parallelism": f"frames sharded, {world} rank(s), no data-path collective"},
                "mverts_per_s": P_total * world * args.steps / (dev_ms / 1e3) / 1e6, "mtexels_per_s": texels * world * args.steps / (dev_ms / 1e3) / 1e6,
                "roofline": roof, "pipeline_roofline": pipeline, "stages": stages,
                "cpu_baseline": cpu,
                "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s / args.steps * 1e3,
                        "path": "uvol_decode_v2_batch (geometry and texture streams concurrent), UVOL_MEM_HOST"},
                "gpu_launches": launches + launches_e2e * args.steps, "clocks": clk,
                "wall_ms_per_step_resident": wall_resident / args.steps * 1e3, "workload_gen_s": info["gen_s"]}
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

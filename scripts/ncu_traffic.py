"""ncu report (`ncu --set full --clock-control none -o X python scripts/prof_c3.py N`) -> profiles/<tag>_traffic.json.

Per launch: duration, DRAM bytes read / written, achieved occupancy, registers, grid.  Per bench stage (the names bench.py prints in
`stages`): DRAM bytes of the launches that make up the stage, summed over ONE decode pass (the last one in the report), plus the
frame count of the capture so that bench.py can scale the figure to its own batch (all stages are linear in the frame count).

usage: python scripts/ncu_traffic.py report.ncu-rep workload frames out.json
"""
import csv
import io
import json
import subprocess
import sys

rep, workload, frames, out = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]


def col(name):
    return hdr.index(name)


def num(r, name, scale_bytes=False):
    if name not in hdr:
        return 0.0
    i = col(name)
    try:
        v = float(r[i].replace(",", ""))
    except ValueError:
        v = 0.0
    if v != v:          # "nan": the metric was not collected for this launch
        v = 0.0
    if scale_bytes:
        u = units[i].lower()
        v *= {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1.0)
    return v


def tscale(name):
    u = units[col(name)].lower()
    return {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3, "s": 1e3}.get(u, 1e-6)


launches = []
for r in rows[2:]:
    name = r[col("Kernel Name")]
    short = name.split("(")[0].split("::")[-1].split("<")[0]
    launches.append({"kernel": short, "full_name": name[:120], "ms": num(r, "gpu__time_duration.sum") * tscale("gpu__time_duration.sum"),
                     "dram_read": num(r, "dram__bytes_read.sum", True), "dram_write": num(r, "dram__bytes_write.sum", True),
                     "warps_active_pct": num(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                     "sm_throughput_pct": num(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
                     "regs": num(r, "launch__registers_per_thread"), "grid": r[col("Grid Size")] if "Grid Size" in hdr else "", "block": r[col("Block Size")] if "Block Size" in hdr else ""})

# the last decode pass: the driver decodes the same batch twice, so every kernel name appears an even number of times -- the second
# half of each name's launches, in launch order (the two streams interleave differently from pass to pass)
seen, total = {}, {}
for l in launches:
    total[l["kernel"]] = total.get(l["kernel"], 0) + 1
last = []
for l in launches:
    i = seen.get(l["kernel"], 0); seen[l["kernel"]] = i + 1
    if total[l["kernel"]] % 2 or i >= total[l["kernel"]] // 2:
        last.append(l)
order = {}
stages = {}


def add(stage, l):
    s = stages.setdefault(stage, {"dram_bytes": 0.0, "ms": 0.0, "launches": 0})
    s["dram_bytes"] += l["dram_read"] + l["dram_write"]; s["ms"] += l["ms"]; s["launches"] += 1


for l in last:
    k = l["kernel"]; n = order.get(k, 0); order[k] = n + 1
    full = l["full_name"]
    if k == "k_rans": add(["rans_attr", "rans_ctx", "rans_recheck"][min(n, 2)], l)
    elif k == "k_rabs_lanes": add(["rabs_seams", "rabs_aux"][min(n, 1)], l)
    elif k.startswith("k_edgebreaker"): add("edgebreaker", l)
    elif k in ("k_seam_count", "k_seams"): add("seams", l)
    elif k == "k_attr_fan": add("attr_tables", l)
    elif k == "k_basis_globals": add("tex_globals", l)
    elif k == "k_scan": add("attr_tables" if n == 0 else "point_count", l)
    elif k == "k_point_fan": add("point_count" if n == 0 else "point_assign", l)
    elif k == "k_plan2": add("plan2", l)
    elif k == "k_face_records": add("face_records", l)
    elif k == "k_traverse": add("traverse", l)
    elif k == "k_parents": add("parents", l)
    elif k == "k_predict_wrap": add("predict_wrap", l)
    elif k == "k_normals": add("normals", l)
    elif k == "k_expand": add("expand_pnc" if n == 0 else "expand", l)
    elif k == "k_uv_prepare": add("uv_prepare", l)
    elif k == "k_predict_uv": add("predict_uv", l)
    elif k.startswith("k_uastc_blocks") or k.startswith("k_etc1s_blocks"): add("tex_blocks", l)
    elif k.startswith("k_etc1s_slices") or k.startswith("k_slice"): add("tex_slices", l)
    elif k.startswith("k_corto_faces"): add("faces", l)
    elif k.startswith("k_corto_delta"): add("delta", l)
    elif k.startswith("k_corto_dequant"): add("dequant", l)
    elif k.startswith("k_corto_values"): add("values", l)
    elif k.startswith("k_tunstall"): add("tunstall", l)
    else: add("other:" + k, l)

doc = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per stage of ONE decode pass, from `ncu --set full --clock-control none` "
                   "(per-launch list below; cold-cache, serialised launches: use the SHARES, not the absolute times)",
       "workload": workload, "frames": frames}
for s, v in stages.items():
    doc[s] = v["dram_bytes"]
doc["stage_ms_under_ncu"] = {s: round(v["ms"], 3) for s, v in stages.items()}
doc["launches"] = [{k: (round(v, 3) if isinstance(v, float) else v) for k, v in l.items() if k != "full_name"} for l in last]
json.dump(doc, open(out, "w"), indent=1)
tot = sum(v["ms"] for v in stages.values()) or 1.0
for s, v in sorted(stages.items(), key=lambda t: -t[1]["ms"]):
    print("%-16s %8.3f ms %5.1f%%  dram %8.3f GB  (%d launches)" % (s, v["ms"], 100 * v["ms"] / tot, v["dram_bytes"] / 1e9, v["launches"]))

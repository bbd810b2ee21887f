"""Summarise an ncu report: per kernel duration, instruction count, stall ratios, and the hottest SASS lines."""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 14
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr = rows[0]
keys = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__average_warp_latency_per_inst_issued.ratio"] + [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:70], "id", r[0])
    print("  ", {k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): round(float(r[hdr.index(k)]), 2) for k in keys if r[hdr.index(k)] not in ("", "0")})
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
cur = None; data = []
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        if cur: data.append(cur)
        cur = [r[1][:60], None, []]
    elif r and r[0] == "Address": cur[1] = r
    elif cur and cur[1] and len(r) > 5:
        h = cur[1]
        try: cur[2].append((r[h.index("Source")].strip(), int(r[h.index("Instructions Executed")]), int(r[h.index("# Samples")])))
        except Exception: pass
if cur: data.append(cur)
for name, h, lines in data:
    tot = sum(l[2] for l in lines) or 1
    print("==", name, "samples", tot)
    for i, (s, e, sm) in sorted(enumerate(lines), key=lambda t: -t[1][2])[:top]:
        print("   %5d %5.1f%%  exec %9d  %s" % (i, 100.0 * sm / tot, e, s[:80]))

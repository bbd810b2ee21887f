import json, sys
for p in sys.argv[1:]:
    try:
        d = json.loads(open(p).read().strip().splitlines()[-1])
    except Exception as e:
        print(p, "unreadable", e); continue
    print(p, "value %.0f fps  %.1f ms/step  e2e %.0f fps (%.0f ms)  cpu %s  windows %s scratch %s GB" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"],
          d["cpu_baseline"] and round(d["cpu_baseline"]["value"]), d["config"].get("windows"), d["config"].get("scratch_gb_largest_window")))
    print("   ", {k: v["ms"] for k, v in d["stages"].items() if v["ms"] > d["ms_per_step"] * 0.01})

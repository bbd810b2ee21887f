#!/bin/bash
# reduced ncu evidence session (one gpurun call, ONE GPU): launch list of the default bench command, every kernel of a 56-frame C3
# window with the full section set (-> per-stage DRAM traffic), and the three UASTC block kernels (RGBA32 / BC7 / ASTC targets).
TAG=${1:-r02z}
O=gpurun_out
T=/tmp/ncu_$TAG; mkdir -p $T
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches_bench_c3.csv python bench.py --steps 1 --warmup 1 --cpu-seconds 0.5 --no-extra-targets > $O/${TAG}_ncu_bench.log 2>&1; tail -c 150 $O/${TAG}_ncu_bench.log; echo
timeout 1200 ncu --set full --clock-control none -k "regex:^k_" -o $T/c3_window56 -f python scripts/prof_c3.py 2>&1 | tail -1
ncu -i $T/c3_window56.ncu-rep --page raw --csv > $O/${TAG}_c3_window56_raw.csv; ls -la $T/c3_window56.ncu-rep $O/${TAG}_c3_window56_raw.csv
python scripts/ncu_traffic.py $O/${TAG}_c3_window56_raw.csv c3 56 $O/${TAG}_traffic_c3.json
timeout 600 ncu --set full --clock-control none -k "regex:^k_uastc" -o $T/tex_targets -f python scripts/prof_tex_targets.py 2>&1 | tail -1
ncu -i $T/tex_targets.ncu-rep --page raw --csv > $O/${TAG}_tex_targets_raw.csv
python scripts/ncu_traffic.py $O/${TAG}_tex_targets_raw.csv tex_targets 28 $O/${TAG}_traffic_tex_targets.json
du -sh $O

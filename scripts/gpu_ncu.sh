#!/bin/bash
# ncu evidence session (one gpurun call, ONE GPU): the launch list of the default bench command, every kernel of a 56-frame C3 window
# and of a 64-frame V1 batch with the full section set (-> per-stage DRAM traffic; raw pages kept as CSV, the reports themselves are
# too large to travel), a source-level capture of the three longest serial kernels, and the TMA on/off A/B of the block kernels.
TAG=${1:-r02}
O=gpurun_out
T=/tmp/ncu_$TAG; mkdir -p $T
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches_bench_c3.csv python bench.py --steps 1 --warmup 1 --cpu-seconds 0.5 --no-extra-targets > $O/${TAG}_ncu_bench.log 2>&1; tail -c 150 $O/${TAG}_ncu_bench.log; echo
timeout 1500 ncu --set full --clock-control none -k "regex:^k_" -o $T/c3_window56 -f python scripts/prof_c3.py 2>&1 | tail -1
ncu -i $T/c3_window56.ncu-rep --page raw --csv > $O/${TAG}_c3_window56_raw.csv; ls -la $T/c3_window56.ncu-rep $O/${TAG}_c3_window56_raw.csv
python scripts/ncu_traffic.py $O/${TAG}_c3_window56_raw.csv c3 56 $O/${TAG}_traffic_c3.json
timeout 900 ncu --set full --clock-control none -k "regex:^k_corto|k_tunstall" -o $T/c5_64 -f python scripts/prof_corto.py 2>&1 | tail -1
ncu -i $T/c5_64.ncu-rep --page raw --csv > $O/${TAG}_c5_64_raw.csv
python scripts/ncu_traffic.py $O/${TAG}_c5_64_raw.csv c5 64 $O/${TAG}_traffic_c5.json
# source-level view of the longest serial kernels (second pass of the 56-frame window)
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_traverse|k_edgebreaker_valence2|k_predict_uv" -s 3 -c 3 -o $O/${TAG}_c3_serial -f python scripts/prof_c3.py 2>&1 | tail -1
python scripts/ncu_hot.py $O/${TAG}_c3_serial.ncu-rep 12 > $O/${TAG}_ncu_hot_serial.txt 2>&1; head -12 $O/${TAG}_ncu_hot_serial.txt
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_corto_faces" -s 1 -c 1 -o $O/${TAG}_c5_faces -f python scripts/prof_corto.py 2>&1 | tail -1
python scripts/ncu_hot.py $O/${TAG}_c5_faces.ncu-rep 14 > $O/${TAG}_ncu_hot_corto.txt 2>&1; head -8 $O/${TAG}_ncu_hot_corto.txt
{ echo "== TMA on"; python scripts/exp_uastc.py 2>&1 | tail -6; python scripts/exp_etc1s.py 2>&1 | tail -2; echo "== TMA off (UVOL_NO_TMA=1)"; UVOL_NO_TMA=1 python scripts/exp_uastc.py 2>&1 | tail -6; UVOL_NO_TMA=1 python scripts/exp_etc1s.py 2>&1 | tail -2; } > $O/${TAG}_tma_ab.txt; cat $O/${TAG}_tma_ab.txt
du -sh $O

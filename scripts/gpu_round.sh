#!/bin/bash
# Round measurement session (one gpurun call on ONE GPU): parity tests, the bench workloads with their reference arms, the ncu
# launch list of the default bench command, one `ncu --set full` capture of every kernel of a 56-frame C3 window (-> per-stage DRAM
# traffic) and of the V1 kernels, and the TMA on/off A/B of the texture block kernels.  Outputs -> gpurun_out/$TAG_*; the summaries
# worth keeping are copied into profiles/ by hand.
TAG=${1:-r02}
O=gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 > $O/${TAG}_bench_c3.json 2> $O/${TAG}_bench_c3.err || tail -3 $O/${TAG}_bench_c3.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_reference_c3.json 2>/dev/null
for w in c2 c5 liam; do
  timeout 400 python bench.py --workload $w --steps 5 --warmup 3 > $O/${TAG}_bench_$w.json 2>/dev/null
  timeout 400 python bench.py --workload $w --impl reference --steps 3 --warmup 1 > $O/${TAG}_reference_$w.json 2>/dev/null
done
for w in c3 c2 c5 liam; do python - <<P
import json
b=json.loads(open('$O/${TAG}_bench_$w.json').read().strip().splitlines()[-1]); r=json.loads(open('$O/${TAG}_reference_$w.json').read().strip().splitlines()[-1])
print('$w value %.0f e2e %.0f %s | reference %.0f (%s, %d cores) -> e2e ratio %.2f | roofline %s %.4f' % (b['value'], b['e2e']['value'], b['unit'], r['value'], r['cpu_baseline']['kind'], r['cpu_baseline']['cores'], b['e2e']['value'] / r['value'], b['roofline'].get('kernel'), b['roofline'].get('frac') or 0))
P
done
# launch list of the bench command (cold-cache, serialised: shares only)
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches_bench_c3.csv python bench.py --steps 1 --warmup 1 --cpu-seconds 0.5 --no-extra-targets > $O/${TAG}_ncu_bench.log 2>&1; tail -c 200 $O/${TAG}_ncu_bench.log
# every kernel of a 56-frame C3 window, full sections (two passes; the second is summarised)
timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:^k_" -o $O/${TAG}_c3_window56 -f python scripts/prof_c3.py 2>&1 | tail -2
python scripts/ncu_traffic.py $O/${TAG}_c3_window56.ncu-rep c3 56 $O/${TAG}_traffic_c3.json
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:^k_corto|k_tunstall" -o $O/${TAG}_c5_64 -f python scripts/prof_corto.py 2>&1 | tail -2
python scripts/ncu_traffic.py $O/${TAG}_c5_64.ncu-rep c5 64 $O/${TAG}_traffic_c5.json
# TMA staging A/B (CUDA-event times of the block kernels, tables / codebooks staged by cp.async.bulk vs by a copy loop)
{ echo "== TMA on"; python scripts/exp_uastc.py 2>&1 | tail -6; python scripts/exp_etc1s.py 2>&1 | tail -2; echo "== TMA off (UVOL_NO_TMA=1)"; UVOL_NO_TMA=1 python scripts/exp_uastc.py 2>&1 | tail -6; UVOL_NO_TMA=1 python scripts/exp_etc1s.py 2>&1 | tail -2; } > $O/${TAG}_tma_ab.txt; cat $O/${TAG}_tma_ab.txt

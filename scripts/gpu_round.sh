#!/bin/bash
# Round measurement session (one gpurun call): parity tests, the three bench workloads + their reference arms, the ncu launch list of
# the default bench command and one `ncu --set full` capture of the hot kernels on a 56-frame C3 window.  Outputs -> gpurun_out/$TAG_*.
TAG=${1:-r01f}
O=gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 600 python bench.py --steps 5 --warmup 3 > $O/${TAG}_bench_c3.json 2> $O/${TAG}_bench_c3.err || tail -3 $O/${TAG}_bench_c3.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_reference_c3.json 2>/dev/null
timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 > $O/${TAG}_bench_c2.json 2>/dev/null
timeout 300 python bench.py --workload c2 --impl reference --steps 3 --warmup 1 > $O/${TAG}_reference_c2.json 2>/dev/null
timeout 300 python bench.py --workload c5 --steps 5 --warmup 3 > $O/${TAG}_bench_c5.json 2>/dev/null
timeout 300 python bench.py --workload c5 --impl reference --steps 3 --warmup 1 > $O/${TAG}_reference_c5.json 2>/dev/null
python scripts/sum_bench.py $O/${TAG}_bench_c3.json $O/${TAG}_bench_c2.json
for w in c3 c2 c5; do python -c "
import json,sys
d=json.loads(open('$O/${TAG}_reference_$w.json').read().strip().splitlines()[-1]); print('reference $w', round(d['value'],1), d['cpu_baseline']['kind'], d['cpu_baseline']['cores'])"; done
python -c "
import json
d=json.loads(open('$O/${TAG}_bench_c5.json').read().strip().splitlines()[-1]); print('c5 value', round(d['value']), 'e2e', round(d['e2e']['value']), 'cpu', round(d['cpu_baseline']['value']), d['stages'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches_bench_c3.csv python bench.py --steps 1 --warmup 1 --cpu-seconds 0.5 > $O/${TAG}_ncu_bench.log 2>&1; tail -c 200 $O/${TAG}_ncu_bench.log
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_uastc_blocks|k_traverse|k_edgebreaker_valence2|k_corner_records|k_predict_uv|k_expand|k_rans" -s 8 -c 8 -o $O/${TAG}_c3_hot python scripts/prof_c3.py 2>&1 | tail -2

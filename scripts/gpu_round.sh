#!/bin/bash
# Round measurement session (one gpurun call on ONE GPU): parity tests, then every bench workload with its reference arm.
# Outputs -> gpurun_out/$TAG_*; what is worth keeping is copied into profiles/ by hand.  The ncu evidence is scripts/gpu_ncu.sh.
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

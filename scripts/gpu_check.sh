#!/bin/bash
# quick GPU loop: parity tests + short bench
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py --steps 2 --warmup 1 --cpu-seconds 1 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.0f fps  e2e %.0f fps  ms/step %.1f  cpu %.0f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['cpu_baseline']['value']))
print({k:v['ms'] for k,v in d['stages'].items() if v['ms']>0.5})
"

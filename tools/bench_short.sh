timeout 200 python bench.py --steps 5 --warmup 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step']); [print(k, round(v['ms_per_step'],3)) for k,v in d['kernels'].items()]
"

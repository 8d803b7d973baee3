python -m pytest tests -x -q -m gpu 2>&1 | tail -4
run() { python bench.py --steps 20 --warmup 3 --skip-cpu 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1 value %.3e ms/step %.3f e2e %.3f (%.3e) k2 %.4f frac %.3f stream %.0f'%(d['value'],d['ms_per_step'],d['e2e']['ms_per_step'],d['e2e']['value'],d['roofline']['ms'],d['roofline']['frac'],d['roofline_stream']['achieved']), d['e2e'])
    elif 'rror' in l or 'Trace' in l: print(l[:300])
"; }
run a; run b

run() { python bench.py --steps 10 --warmup 3 --skip-cpu --skip-stream 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1 value %.3e ms/step %.3f e2e %.3f k2 %.4f frac %.3f'%(d['value'],d['ms_per_step'],d['e2e']['ms_per_step'],d['roofline']['ms'],d['roofline']['frac']), d['plan'])
    elif 'rror' in l or 'Trace' in l: print(l[:300])
"; }
run balanced; BI_MMA_FULL_UNITS=1 run full; run balanced; BI_MMA_FULL_UNITS=1 run full
BI_MMA_FULL_UNITS=1 BI_MMA_TARGET_UNITS=7104 run full7104

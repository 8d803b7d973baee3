python -m pytest tests -x -q -m gpu 2>&1 | tail -6
run() { python bench.py --steps 20 --warmup 3 --skip-cpu 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1 value %.3e ms/step %.3f e2e %.3f k2 %.4f frac %.3f stream %.0f'%(d['value'],d['ms_per_step'],d['e2e']['ms_per_step'],d['roofline']['ms'],d['roofline']['frac'],d['roofline_stream']['achieved']))
    elif 'rror' in l or 'Trace' in l: print(l[:300])
"; }
run tmap; BI_MMA_NO_TENSORMAP=1 run rows; run tmap; BI_MMA_NO_TENSORMAP=1 run rows

run() { python bench.py --steps 20 --warmup 3 --skip-cpu 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1 value %.3e ms/step %.3f e2e %.3f k2 %.4f frac %.3f stream %.0f'%(d['value'],d['ms_per_step'],d['e2e']['ms_per_step'],d['roofline']['ms'],d['roofline']['frac'],d['roofline_stream']['achieved']), d['plan']['work_units'])
    elif 'rror' in l or 'Trace' in l: print(l[:300])
"; }
run t28416; BI_MMA_TARGET_UNITS=14208 run t14208;  BI_MMA_TARGET_UNITS=56832 run t56832
python -m pytest tests -x -q -m gpu 2>&1 | tail -2

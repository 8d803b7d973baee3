python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "unbinned or grouped or batch_shape or anchor_hit or nan_inf or zero_rates" 2>&1 | tail -3
for tu in 3552 7104; do
BI_MMA_TARGET_UNITS=$tu python bench.py --steps 10 --warmup 3 --skip-cpu 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('TU $tu value %.3e ms/step %.3f e2e %.3f k2 %.4f frac %.3f stream %.0f GB/s units %d'%(d['value'],d['ms_per_step'],d['e2e']['ms_per_step'],d['roofline']['ms'],d['roofline']['frac'],d['roofline_stream']['achieved'],d['plan']['work_items']))
    elif 'rror' in l: print(l[:300])
"
done

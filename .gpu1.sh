set -x
python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "unbinned or grouped or batch_shape or anchor_hit or nan_inf or zero_rates" > gpurun_out/pytest5.log 2>&1; tail -15 gpurun_out/pytest5.log
python bench.py --steps 10 --warmup 3 --skip-cpu > gpurun_out/bench5.log 2>&1; echo bench exit $?
python - <<'PY'
import json
for l in open('gpurun_out/bench5.log'):
    if l.startswith('{'):
        d=json.loads(l)
        print('value %.3e e2e %.3e ms/step %.3f e2e ms %.3f'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['e2e']['ms_per_step']))
        print('roofline',{k:v for k,v in d['roofline'].items() if k in('achieved','peak','frac','ms','share_of_step')})
        print('stream',{k:v for k,v in (d['roofline_stream'] or {}).items() if k in('kernel','achieved','frac','ms','plain_read_gbs_this_run')})
        print('plan',d['plan'],'peak',d['fp64_fma_peak_tflops'])
PY
tail -3 gpurun_out/bench5.log | cut -c1-600
python bench.py --steps 10 --warmup 3 --skip-cpu --kernel grouped --stream-kernel stream 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('LEGACY value %.3e ms/step %.3f'%(d['value'],d['ms_per_step']), d['roofline']['ms'], d['roofline_stream']['achieved'])
"

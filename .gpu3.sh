python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 > gpurun_out/bench6.log 2>&1; echo bench rc $?
python - <<'PY'
import json
for l in open('gpurun_out/bench6.log'):
    if l.startswith('{'):
        d=json.loads(l)
        print('value %.3e e2e %.3e ms/step %.3f e2e ms %.3f launches %s'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['e2e']['ms_per_step'],d['gpu_launches']))
        print('roofline',{k:v for k,v in d['roofline'].items() if k in('achieved','peak','frac','ms','share_of_step')})
        print('stream',{k:v for k,v in (d['roofline_stream'] or {}).items() if k in('kernel','achieved','frac','ms')})
        print('cpu',d['cpu_baseline']['value'], d['plan'], d['clocks'])
PY
tail -2 gpurun_out/bench6.log | cut -c1-300

python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.log 2>&1; echo rc $?
grep "^{" gpurun_out/bench_2gpu.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('n_gpus',d['n_gpus'],'value %.3e ms/step %.3f e2e %.3e'%(d['value'],d['ms_per_step'],d['e2e']['value']), d['roofline']['frac'])
"
tail -5 gpurun_out/bench_2gpu.log | cut -c1-300
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>&1 | tail -2 | cut -c1-400

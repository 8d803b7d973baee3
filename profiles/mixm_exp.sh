timeout 250 python -m pytest tests/test_gpu_template.py -x -q 2>&1 | tail -3
for N in 100000000 20000000 2000000; do
TPL_POINTS=7,11,64 timeout 90 python profiles/template_bench.py 0 $N 2>&1 | grep C5 | sed "s/C5 mixture bin_major=0: //; s/K=96, //; s/; prepared.*//"
done

timeout 300 python -m pytest tests/test_gpu_philox.py tests/test_gpu_ensemble.py -q -x 2>&1 | tail -2
run() { MCL_PHILOX_SHARE_BM=$1 timeout 90 python bench.py --workload $2 --replicas $3 --steps 3 --warmup 1 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('share_bm=$1 $2', round(d['value']/1e9,1), round(d['roofline']['frac'],3))"; }
run 0 c2 2960; run 1 c2 2960; run 0 c2 2960; run 1 c2 2960; run 0 c5 5920; run 1 c5 5920

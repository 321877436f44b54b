timeout 200 python -m pytest tests/test_gpu_ensemble.py -q -x -k "histograms or conserve" 2>&1 | tail -2
run() { MCL_B200_LIB=$PWD/scripts/ab_libs/$1.so timeout 90 python bench.py --workload $2 --replicas $3 --steps 3 --warmup 1 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 $2', round(d['value']/1e9,1), round(d['roofline']['frac'],3))"; }
run libbase c2 2960; run libvar c2 2960; run libbase c5 5920; run libvar c5 5920

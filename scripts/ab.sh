run() { MCL_PHILOX_NT=$2 timeout 60 python bench.py --workload $1 --replicas $3 --steps 1 --warmup 1 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 NT=$2', round(d['value']/1e9,1), round(d['roofline']['frac'],3))"; }
run c2 256 1480; run c2 256 1480

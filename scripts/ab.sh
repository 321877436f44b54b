# quick sanity + throughput of the tree's own build
timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
run() { timeout 90 python bench.py --workload $1 --replicas $2 --steps 3 --warmup 1 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e9,1), round(d['roofline']['frac'],3))"; }
run c2 2960; run c5 5920

# scratch A/B runner: bash scripts/ab.sh <libA.so> <libB.so> ...   (paths relative to the repo root)
run() { MCL_B200_LIB=$PWD/$1 timeout 120 python bench.py --workload $2 --replicas $3 --steps 3 --warmup 1 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 $2', round(d['value']/1e9,1), round(d['roofline']['frac'],3))"; }
for lib in "$@"; do run $lib c2 2960; run $lib c5 5920; done

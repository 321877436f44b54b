# usage: bash scripts/ab_one.sh <lib.so> <workload> <replicas>   (extra env passes through) -> value/1e9 and roofline fraction
MCL_B200_LIB=$PWD/$1 timeout 120 python bench.py --workload $2 --replicas $3 --steps 3 --warmup 1 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 $2', round(d['value']/1e9,1), round(d['roofline']['frac'],3))"

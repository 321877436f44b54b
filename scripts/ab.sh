run() { MCL_B200_LIB=$PWD/scripts/ab_libs/$1.so timeout 90 python bench.py --workload $2 --replicas $3 --steps 3 --warmup 1 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 $2', round(d['value']/1e9,1), round(d['roofline']['frac'],3))"; }
run libbase c5 5920; run libvar2 c5 5920; run libbase c2 2960; run libvar2 c2 2960
for lib in libbase libvar2; do echo $lib; MCL_B200_LIB=$PWD/scripts/ab_libs/$lib.so timeout 100 python scripts/c4_probe.py 2>&1 | grep -E "S= 16384|S=  4096"; done

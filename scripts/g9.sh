mkdir -p gpurun_out
python scripts/c4_timing.py 2>&1 | grep -E "mcl_objective|python wall" > gpurun_out/g9_c4.log; cat gpurun_out/g9_c4.log
ab() { timeout 200 python bench.py --workload $1 --replicas $2 --steps 3 --warmup 2 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 $2 $3', round(d['value']/1e9,1), round(d['roofline']['frac'],3), round(d['ms_per_step'],2), round(d['roofline']['achieved']/1e9,1), 'e2e', round(d['e2e']['value']/1e9,1))"; }
{
ab c5 5920 nt64
MCL_PHILOX_NT=128 ab c5 5920 nt128
MCL_PHILOX_NT=256 ab c5 5920 nt256
MCL_PHILOX_NT=32 ab c5 5920 nt32
ab c5 50000 full
ab c4 4096 smallbox
ab c3 2560 relist
ab c1 8 default
} > gpurun_out/g9_ab.log 2>&1
cat gpurun_out/g9_ab.log

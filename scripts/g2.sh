mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q -x -s 2>&1 | tail -60) > gpurun_out/g2_pytest.log
tail -5 gpurun_out/g2_pytest.log
ab() { MCL_B200_LIB=$1 timeout 150 python bench.py --workload $2 --replicas $3 --steps 2 --warmup 1 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 $2 $4', round(d['value']/1e9,1), round(d['roofline']['frac'],3), round(d['ms_per_step'],2), round(d['roofline']['achieved']/1e9,1))"; }
NEW=$PWD/mcluminescence_b200/_lib/libmcl_b200.so
{
ab $PWD/scripts/ab_libs/base.so c2 2960 one
ab $NEW c2 2960 one
MCL_BENCH_TWO_CHANNEL=1 ab $NEW c2 2960 two
ab $NEW c5 5920 one
ab $NEW c3 2560 relist
MCL_PHILOX_RELIST=0 ab $NEW c3 2560 norelist
ab $NEW c4 4096 smem
MCL_PHILOX_SMEM_SLAB=0 ab $NEW c4 4096 global
} > gpurun_out/g2_ab.log 2>&1
cat gpurun_out/g2_ab.log

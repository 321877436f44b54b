mkdir -p gpurun_out
ab() { MCL_B200_LIB=$1 timeout 150 python bench.py --workload $2 --replicas $3 --steps 2 --warmup 2 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$(basename $1) $2 $4', round(d['value']/1e9,1), round(d['roofline']['frac'],3), round(d['ms_per_step'],2), round(d['roofline']['achieved']/1e9,1))"; }
NEW=$PWD/mcluminescence_b200/_lib/libmcl_b200.so
{
for v in $NEW $PWD/scripts/ab_libs/r01.so $PWD/scripts/ab_libs/oldtie.so; do
ab $v c2 2960 one
ab $v c5 5920 one
done
MCL_BENCH_TWO_CHANNEL=1 ab $NEW c2 2960 two
ab $NEW c3 2560 relist
ab $NEW c4 4096 smallbox
ab $NEW c4 16384 smallbox
} > gpurun_out/g6_ab.log 2>&1
cat gpurun_out/g6_ab.log
# ncu: C2 block kernel (one wave of 3 CTAs per SM) and the one-warp kernel on a C4 population
timeout 400 ncu --set full --clock-control none --import-source on -k regex:philox_kernel -s 1 -c 1 -o gpurun_out/prof_r02a_c2 python bench.py --workload c2 --steps 1 --warmup 1 --replicas 444 --no-cpu > gpurun_out/ncu_r02a_c2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:smallbox_kernel -s 5 -c 1 -o gpurun_out/prof_r02a_c4 python bench.py --workload c4 --steps 1 --warmup 1 --replicas 4096 --no-cpu > gpurun_out/ncu_r02a_c4.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3

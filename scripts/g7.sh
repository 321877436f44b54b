mkdir -p gpurun_out
ab() { MCL_B200_LIB=$1 timeout 200 python bench.py --workload $2 --replicas $3 --steps 2 --warmup 2 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$(basename $1) $2 $3 $4', round(d['value']/1e9,1), round(d['roofline']['frac'],3), round(d['ms_per_step'],2), round(d['roofline']['achieved']/1e9,1))"; }
NEW=$PWD/mcluminescence_b200/_lib/libmcl_b200.so
{
for v in $NEW $PWD/scripts/ab_libs/nodefer.so $PWD/scripts/ab_libs/r01.so; do
ab $v c2 2960 one
ab $v c2 10000 one
ab $v c5 5920 one
done
} > gpurun_out/g7_ab.log 2>&1
cat gpurun_out/g7_ab.log
(timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15) > gpurun_out/g7_pytest.log
tail -4 gpurun_out/g7_pytest.log

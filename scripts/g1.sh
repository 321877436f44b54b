# GPU call: full GPU test suite, A/B of the previous kernel vs this one (C2 one-/two-channel, C5), other configs
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x -s 2>&1 | tail -60) > gpurun_out/g1_pytest.log
tail -5 gpurun_out/g1_pytest.log
ab() { MCL_B200_LIB=$1 timeout 150 python bench.py --workload $2 --replicas $3 --steps 2 --warmup 1 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 $2 $4', round(d['value']/1e9,1), round(d['roofline']['frac'],3))"; }
{
ab $PWD/scripts/ab_libs/base.so c2 2960 one
ab $PWD/mcluminescence_b200/_lib/libmcl_b200.so c2 2960 one
MCL_BENCH_TWO_CHANNEL=1 ab $PWD/scripts/ab_libs/base.so c2 2960 two
MCL_BENCH_TWO_CHANNEL=1 ab $PWD/mcluminescence_b200/_lib/libmcl_b200.so c2 2960 two
ab $PWD/scripts/ab_libs/base.so c5 5920 one
ab $PWD/mcluminescence_b200/_lib/libmcl_b200.so c5 5920 one
} > gpurun_out/g1_ab.log 2>&1
cat gpurun_out/g1_ab.log
(timeout 200 python scripts/bench_configs.py 2>&1 | tail -60) > gpurun_out/g1_configs.log
grep -E "esteps_per_s|seconds" gpurun_out/g1_configs.log

# Full round-end measurement set on one GPU: bench lines (default C2, reference arm, the other configs), ncu launch list,
# ncu --set full captures of the two native kernels, sanitizer passes, GPU tests.   usage: bash scripts/round_bench.sh <tag>
set -x
tag=${1:-r02}
mkdir -p gpurun_out
timeout 900 python bench.py 2> gpurun_out/bench_${tag}_n1.err | tail -1 > gpurun_out/bench_${tag}_n1.json
timeout 600 python bench.py --impl reference 2> gpurun_out/bench_${tag}_ref.err | tail -1 > gpurun_out/bench_${tag}_ref.json
for wl in c4 c5 c3 c1; do
  timeout 600 python bench.py --workload $wl 2> gpurun_out/bench_${tag}_$wl.err | tail -1 > gpurun_out/bench_${tag}_$wl.json
done
timeout 600 python bench.py --workload c5 --scaling strong --no-cpu 2>/dev/null | tail -1 > gpurun_out/bench_${tag}_c5_strong_n1.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${tag}.csv python bench.py --steps 2 --warmup 1 --replicas 1480 --no-cpu > gpurun_out/ncu_launches_${tag}.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${tag}_c4.csv python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launches_${tag}_c4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:philox_kernel -s 1 -c 1 -o gpurun_out/prof_${tag}_c2 python bench.py --steps 1 --warmup 1 --replicas 444 --no-cpu > gpurun_out/ncu_full_${tag}_c2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:smallbox_kernel -s 8 -c 1 -o gpurun_out/prof_${tag}_c4 python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_full_${tag}_c4.log 2>&1
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool python scripts/sanitize_small.py > gpurun_out/sanitize_${tag}_$tool.log 2>&1; tail -3 gpurun_out/sanitize_${tag}_$tool.log
done
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8) > gpurun_out/pytest_${tag}.log; tail -3 gpurun_out/pytest_${tag}.log
for f in n1 ref c4 c5 c3 c1; do cut -c1-260 gpurun_out/bench_${tag}_$f.json; done

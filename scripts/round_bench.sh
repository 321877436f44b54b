# Full round-end measurement set on one GPU: bench line, reference arm, ncu launch list, one ncu --set full capture.
set -x
tag=${1:-r01}
timeout 600 python bench.py 2> gpurun_out/bench_${tag}_n1.err | tail -1 > gpurun_out/bench_${tag}_n1.json
timeout 600 python bench.py --impl reference 2> gpurun_out/bench_${tag}_ref.err | tail -1 > gpurun_out/bench_${tag}_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${tag}.csv python bench.py --steps 2 --warmup 1 --replicas 1480 --no-cpu > gpurun_out/ncu_launches_${tag}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:philox_kernel -s 1 -c 1 -o gpurun_out/prof_${tag}_c2 python bench.py --steps 1 --warmup 1 --replicas 444 --no-cpu > gpurun_out/ncu_full_${tag}.log 2>&1
cut -c1-300 gpurun_out/bench_${tag}_n1.json; cut -c1-300 gpurun_out/bench_${tag}_ref.json

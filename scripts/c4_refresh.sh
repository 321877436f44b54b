tag=r02d
timeout 600 python bench.py --workload c4 2> gpurun_out/bench_${tag}_c4.err | tail -1 > gpurun_out/bench_${tag}_c4.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${tag}_c4.csv python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launches_${tag}_c4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:smallbox_kernel -s 8 -c 1 -o gpurun_out/prof_${tag}_c4 python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_full_${tag}_c4.log 2>&1
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4) > gpurun_out/pytest_${tag}.log; tail -2 gpurun_out/pytest_${tag}.log
cut -c1-300 gpurun_out/bench_${tag}_c4.json

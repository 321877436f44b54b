# usage: bash scripts/prof.sh <tag> <workload> <replicas> [NT]
tag=$1; wl=$2; rep=$3; nt=$4
if [ -n "$nt" ]; then export MCL_PHILOX_NT=$nt; fi
timeout 600 ncu --set full --clock-control none --import-source on -k regex:philox_kernel -s 1 -c 1 -o gpurun_out/prof_$tag python bench.py --workload $wl --steps 1 --warmup 1 --replicas $rep --no-cpu > gpurun_out/ncu_$tag.log 2>&1
tail -2 gpurun_out/ncu_$tag.log

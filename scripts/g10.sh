# multi-GPU check: N ranks (default 2) -- weak C2 (short), strong C5 and C4, plus the N=1 counterparts of the strong runs
N=${1:-2}
mkdir -p gpurun_out
run() { # workload scaling replicas
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $1 --scaling $2 --replicas $3 --steps 3 --warmup 2 --no-cpu 2>gpurun_out/g10_n${N}_$1.err | tail -1 > gpurun_out/g10_n${N}_$1_$2.json
python -c "import json; d=json.load(open('gpurun_out/g10_n${N}_$1_$2.json')); print('N=$N $1 $2', round(d['value']/1e9,1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']/1e9,1))"
}
run c2 weak 1480
run c5 strong 50000
run c4 strong 4096
run c4 strong 32768
if [ "$N" = "2" ]; then
for wl in "c2 weak 1480" "c5 strong 50000" "c4 strong 4096" "c4 strong 32768"; do set -- $wl
timeout 400 python bench.py --gpus 1 --workload $1 --scaling $2 --replicas $3 --steps 3 --warmup 2 --no-cpu 2>/dev/null | tail -1 > gpurun_out/g10_n1_$1_$2_$3.json
python -c "import json; d=json.load(open('gpurun_out/g10_n1_$1_$2_$3.json')); print('N=1 $1 $2 $3', round(d['value']/1e9,1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']/1e9,1))"
done
fi

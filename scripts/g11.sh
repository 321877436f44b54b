# 8-GPU check (charged 8x): strong C5 (50 000 replicas) and weak C2 (1480 replicas per GPU), device-timed max over ranks
N=${1:-8}
mkdir -p gpurun_out
run() { # workload scaling replicas
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $1 --scaling $2 --replicas $3 --steps 3 --warmup 3 --no-cpu 2>gpurun_out/g11_n${N}_$1.err | tail -1 > gpurun_out/g11_n${N}_$1_$2.json
python -c "import json; d=json.load(open('gpurun_out/g11_n${N}_$1_$2.json')); print('N=$N $1 $2', round(d['value']/1e9,1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']/1e9,1))"
}
run c5 strong 50000
run c2 weak 1480

mkdir -p gpurun_out
ab() { timeout 200 python bench.py --workload $1 --replicas $2 --steps 3 --warmup 2 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 $2 $3', round(d['value']/1e9,1), round(d['roofline']['frac'],3), round(d['ms_per_step'],2), round(d['roofline']['achieved']/1e9,1), 'e2e', round(d['e2e']['value']/1e9,1))"; }
{
ab c4 4096 smallbox
ab c4 16384 smallbox
ab c3 2560 relist
} > gpurun_out/g8_ab.log 2>&1
cat gpurun_out/g8_ab.log
# multi-GPU: weak C2 (short), strong C5 and C4
for wl in "c2 weak 1480" "c5 strong 50000" "c4 strong 4096"; do set -- $wl
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload $1 --scaling $2 --replicas $3 --steps 2 --warmup 2 --no-cpu 2>gpurun_out/g8_n2_$1.err | tail -1 > gpurun_out/g8_n2_$1.json
python -c "import json; d=json.load(open('gpurun_out/g8_n2_$1.json')); print('N=2 $1 $2', round(d['value']/1e9,1), d['ms_per_step'], d['config'])"
timeout 300 python bench.py --gpus 1 --workload $1 --scaling $2 --replicas $3 --steps 2 --warmup 2 --no-cpu 2>/dev/null | tail -1 > gpurun_out/g8_n1_$1.json
python -c "import json; d=json.load(open('gpurun_out/g8_n1_$1.json')); print('N=1 $1 $2', round(d['value']/1e9,1), d['ms_per_step'])"
done

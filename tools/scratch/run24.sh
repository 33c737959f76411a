mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2k_bench_n8.json 2> gpurun_out/r2k_bench_n8.err; echo rc8=$?
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2k_bench_n4.json 2> gpurun_out/r2k_bench_n4.err; echo rc4=$?
python -c "
import json
for n in (8,4):
    d=json.loads(open('gpurun_out/r2k_bench_n%d.json'%n).read().strip().splitlines()[-1]); print(n, d['value'], d['e2e']['value'], {k:v.get('ms') for k,v in d['configs'].items()})"

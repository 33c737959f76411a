mkdir -p gpurun_out
python -m pytest tests/test_gpu_distributed.py tests/test_gpu_kzg.py -x -q -m gpu 2>&1 | tail -4 > gpurun_out/r2k_tests_2gpu.log; cat gpurun_out/r2k_tests_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2k_bench_n2.json 2> gpurun_out/r2k_bench_n2.err; echo rc2=$?
python -c "
import json; d=json.loads(open('gpurun_out/r2k_bench_n2.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], {k:v.get('ms') for k,v in d['configs'].items()})"

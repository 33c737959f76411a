python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2h_bench_n2.json 2> gpurun_out/r2h_bench_n2.err; echo rc2=$?
python bench.py --steps 20 --warmup 5 > gpurun_out/r2h_bench_n1.json 2> gpurun_out/r2h_bench_n1.err; echo rc1=$?

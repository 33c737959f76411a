mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2j_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2j_bench_n1.json 2> gpurun_out/r2j_bench_n1.err; echo rc=$?
tail -3 gpurun_out/r2j_tests.log

python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py --no-cpu-baseline > gpurun_out/r2g_bench_n1.json 2> gpurun_out/r2g_bench_n1.err; echo bench rc=$?

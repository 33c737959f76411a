mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2l_tests.log; tail -2 gpurun_out/r2l_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/r2l_bench_n1.json 2> gpurun_out/r2l_bench_n1.err; echo rc=$?
python -c "
import json; d=json.loads(open('gpurun_out/r2l_bench_n1.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['gpu_launches'], d['breakdown_ms_per_step'])"

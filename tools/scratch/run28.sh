python -m pytest tests/test_gpu_kzg.py tests/test_gpu_protocols.py tests/test_gpu_vector_commitment.py -x -q -m gpu 2>&1 | tail -2
python tools/msm_bench.py 20 22 2>/dev/null | grep '^{' | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['log_n'], d['scalars'], round(d['wall_ms'],3), round(d['points_per_s']/1e6,1), round(d['accumulate_ms'],3), round(d['sort_ms'],3), round(d['merge_ms'],3), round(d['reduce_ms'],3))"
python tools/bench_configs.py c5 2>&1 | tail -2 | cut -c1-600
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-fold --no-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('device', d['value'], 'e2e', d['e2e']['value'], d['breakdown_ms_per_step'])"

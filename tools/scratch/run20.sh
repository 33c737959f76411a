python -m pytest tests/test_gpu_kzg.py tests/test_gpu_protocols.py -x -q -m gpu 2>&1 | tail -3
for s in 17 18 19; do python tools/shape_n8.py $s 10 2>&1 | tail -1; done
python tools/shape_n8.py 20 5 2>&1 | tail -1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-fold --no-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('device', d['value'], 'e2e', d['e2e']['value'], d['breakdown_ms_per_step'])"

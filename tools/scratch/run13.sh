python -m pytest tests/test_gpu_kzg.py tests/test_gpu_protocols.py tests/test_gpu_lagrange.py tests/test_gpu_full_size.py -x -q -m gpu 2>&1 | tail -3
python tools/shape_shout.py 1 3 2>&1 | tail -1
python tools/shape_n8.py 17 10 2>&1 | tail -1
python tools/shape_n8.py 18 5 2>&1 | tail -1
python tools/shape_n8.py 19 5 2>&1 | tail -1
python bench.py --no-cpu-baseline --no-fold --no-configs --steps 20 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('device', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), d['breakdown_ms_per_step'])
"

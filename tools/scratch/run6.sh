python -m pytest tests/test_gpu_full_size.py tests/test_gpu_protocols.py -x -q -m gpu 2>&1 | tail -3
for t in h2d_overlap=1 h2d_overlap=0 h2d_overlap=1 h2d_overlap=0; do
TSGPU_TUNING=$t python bench.py --no-cpu-baseline --no-fold --no-configs --steps 20 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$t', 'device', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'A e2e', round(d['distribution_A']['e2e'],3), 'fw e2e', round(d['full_width_values']['e2e'],3), d['e2e']['gpu_launches'])
"
done

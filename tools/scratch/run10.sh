python -m pytest tests/test_gpu_kzg.py -x -q -m gpu 2>&1 | tail -3
for t in h2d_overlap=1 h2d_overlap=0; do
TSGPU_TUNING=$t python bench.py --no-cpu-baseline --no-fold --steps 5 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$t', 'device', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'C3', d['configs']['C3'].get('ms'), 'C5', d['configs']['C5'].get('ms'))
"
done
for lg in 18 19; do for c in 17 $lg 20; do echo "== 2^$lg ops/rank, c=$c"; TSGPU_TABLE_WINDOW_BITS=$c python tools/shape_n8.py $lg 5 2>&1 | tail -1 | cut -c1-420; done; done

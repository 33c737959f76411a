mkdir -p gpurun_out
python tools/msm_bench.py 16 17 18 19 20 21 22 2>/dev/null | grep '^{' > gpurun_out/r02_msm_bench_sizes.jsonl
python -c "
import json
for l in open('gpurun_out/r02_msm_bench_sizes.jsonl'):
    d=json.loads(l); print(d['log_n'], d['scalars'], round(d['wall_ms'],3), round(d['points_per_s']/1e6,1), round(d['accumulate_ms'],3), round(d['sort_ms'],3), round(d['merge_ms'],3), round(d['reduce_ms'],3))"

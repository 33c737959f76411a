mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fold --no-configs"
for G in 32 128; do
  TSGPU_L2_FETCH=$G python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-fold --no-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('L2_FETCH $G device', d['value'], 'e2e', d['e2e']['value'], d['breakdown_ms_per_step']['msm_accumulate_4x'])"
  TSGPU_L2_FETCH=$G timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_msm_accumulate -c 2 --csv --log-file gpurun_out/l2fetch_$G.csv $B > /dev/null 2>&1
  tail -2 gpurun_out/l2fetch_$G.csv | awk -F, '{print $(NF-2), $(NF-1), $NF}'
done
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_msm_accumulate -c 2 --csv --log-file gpurun_out/l2fetch_default.csv $B > /dev/null 2>&1
tail -6 gpurun_out/l2fetch_default.csv | awk -F, '{print $(NF-2), $(NF-1), $NF}'

mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2k_tests.log; tail -2 gpurun_out/r2k_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench_n1.json 2> gpurun_out/r2k_bench_n1.err; echo rc=$?
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fold --no-configs"
$B > gpurun_out/r02_bench_short.json 2> gpurun_out/r02_bench_short.err || exit 1
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/r02_ncu_list.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_msm_accumulate" -c 6 -f -o gpurun_out/r02_acc_in_bench $B > gpurun_out/r02_ncu_acc.log 2>&1
python tools/shape_n8.py 17 3 > gpurun_out/r02_shape_n8.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file gpurun_out/r02_launches_shape_n8.csv python tools/shape_n8.py 17 1 > gpurun_out/r02_ncu_shape_n8.log 2>&1
ls -la gpurun_out | tail -8

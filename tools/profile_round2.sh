# Round-2 profiling recipe (one GPU).  Every ncu pass runs only after the same command has exited 0 without ncu.
#   1. launch list of the bench command (per-launch durations: shares of the step)
#   2. `--set full` of k_msm_accumulate inside the bench command: the first launches belong to the guard proof, launches 3-6 are open passes (2 x 2^20 full-width scalars)
#   3. `--set full` of the sum-check / reduction kernels on the small targets workload (tools/ncu_targets.py)
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fold --no-configs"
$B > gpurun_out/r02_bench_short.json 2> gpurun_out/r02_bench_short.err || exit 1
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/r02_ncu_list.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_msm_accumulate" -c 6 -f -o gpurun_out/r02_acc_in_bench $B > gpurun_out/r02_ncu_acc.log 2>&1
python tools/ncu_targets.py > gpurun_out/r02_targets.log 2>&1 && timeout 500 ncu --set full --clock-control none --import-source on \
    -k regex:"k_msm_accumulate|k_bind$|k_bind_eval2_claim|k_round_eval|k_msm_span_sums|k_msm_bit_sums|k_msm_scatter|k_msm_digits" -c 10 -f -o gpurun_out/r02_targets \
    python tools/ncu_targets.py > gpurun_out/r02_ncu_targets.log 2>&1
ls -la gpurun_out | tail -12
#   4. launch list of the per-rank shape of an 8-way sharded proof (2^17 operations per rank)
python tools/shape_n8.py 17 3 > gpurun_out/r02_shape_n8.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file gpurun_out/r02_launches_shape_n8.csv python tools/shape_n8.py 17 1 > gpurun_out/r02_ncu_shape_n8.log 2>&1

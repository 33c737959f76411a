mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fold > gpurun_out/s4_bench_short.log 2>&1 || exit 1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/launches_s4.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fold > gpurun_out/s4_ncu_list.log 2>&1
python tools/ncu_targets.py > gpurun_out/s4_targets.log 2>&1 && timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_msm_accumulate|k_bind$|k_bind_eval2_claim|k_round_eval|k_msm_span_sums|k_table_sum|k_table_add" -c 8 -f -o gpurun_out/prof_s4 python tools/ncu_targets.py > gpurun_out/s4_ncu_full.log 2>&1
# compute-sanitizer is closed on this pool (runs under it left GPUs needing a reset): bounds are covered by the parity tests instead
ls -la gpurun_out

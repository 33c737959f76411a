"""The work ONE rank does in an 8-way sharded Twist proof of 2^20 operations, on one GPU: a 2^17-op trace through tsgpu_twist_prove_sharded with a one-rank
communicator (same slice length, window width c = 17 and launch chain; no exchange).  For `ncu --metrics gpu__time_duration.sum` launch lists of that shape.
usage: python tools/shape_n8.py [log_ops_per_rank] [reps]"""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
ts = importlib.import_module("multilinear-map-cryptography_b200")
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 17
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = ts.Context(0)
ctx.comm_init(1, 0)
pp, vp = ts.setup_params(ctx, log_n - 2)
n = 1 << log_n
addr, vals_u64, isw = bench.trace_random(log_n, 16, ts.chacha20_u64(bytes([2]) * 32, 3 * n))
vals = ts.fe_vec(vals_u64)
tw = ts.Twist.new(pp)
p = tw.prove_sharded(addr, vals, n)
assert tw.verify(p, vp)
ctx.set_tuning("kernel_timing", 1)
for _ in range(2):
    tw.prove_sharded(addr, vals, n)
ctx.timer_reset()
ctx.synchronize(); t0 = time.perf_counter()
for _ in range(reps):
    tw.prove_sharded(addr, vals, n)
ctx.synchronize(); dt = (time.perf_counter() - t0) / reps
print({"log_ops": log_n, "ms_per_proof_wall": dt * 1e3, **{k: ctx.timer_read(k)[0] / reps for k in ("msm_total", "msm_sort", "msm_accumulate", "msm_merge", "msm_reduce", "open_bary", "wall_commit", "wall_exchange", "wall_transcript", "wall_open_partial", "wall_open_finish")}})

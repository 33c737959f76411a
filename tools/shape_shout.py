"""The work ONE rank does in an N-way sharded Shout proof of C3 (2^20-entry table, 2^22 lookups), on one GPU: a 2^(20-s)-entry table and 2^(22-s) lookups through
tsgpu_shout_prove_sharded with a one-rank communicator (same slice lengths, window widths and launch chain; no exchange).
usage: python tools/shape_shout.py [s = log2 N] [reps]      (TSGPU_TUNING=key=value,... as in bench.py)"""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
ts = importlib.import_module("multilinear-map-cryptography_b200")
s = int(sys.argv[1]) if len(sys.argv) > 1 else 1
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
logT, logL = 20 - s, 22 - s
ctx = ts.Context(0)
for kv in filter(None, os.environ.get("TSGPU_TUNING", "").split(",")):
    k, v = kv.split("="); ctx.set_tuning(k, int(v))
ctx.comm_init(1, 0)
pp, vp = ts.setup_params(ctx, logL - 2)
T, L = 1 << logT, 1 << logL
i = np.arange(T, dtype=np.uint64)
idx = ts.chacha20_u64(bytes([3]) * 32, L) % np.uint64(T)
ent = ts.fe_vec(i * i)
sh = ts.Shout.new(pp)
p = sh.prove_sharded(ent, T, idx, L)
assert sh.verify(p, vp)
ctx.set_tuning("kernel_timing", 1)
for _ in range(2):
    sh.prove_sharded(ent, T, idx, L)
ctx.timer_reset()
ctx.synchronize(); t0 = time.perf_counter()
for _ in range(reps):
    sh.prove_sharded(ent, T, idx, L)
ctx.synchronize(); dt = (time.perf_counter() - t0) / reps
names = ("msm_total", "msm_sort", "msm_accumulate", "msm_merge", "msm_reduce", "open_bary", "wall_commit", "wall_transcript", "wall_open_partial", "wall_open_finish")
print({"logT": logT, "logL": logL, "ms_per_proof_wall": round(dt * 1e3, 3), **{k: round(ctx.timer_read(k)[0] / reps, 3) for k in names}})

"""Phase timings of the two batched MSM passes of one Twist::prove (commit pass over the raw values, open pass over the two
quotient vectors) at 2^20 operations.  usage: python tools/pass_bench.py"""
import ctypes as C, importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
ts = importlib.import_module("multilinear-map-cryptography_b200")
from importlib import import_module
B = import_module("multilinear-map-cryptography_b200.binding")
ctx = ts.Context(0)
pp, vp = ts.setup_params(ctx, 18)
n = 1 << 20
addr, vals_u64, isw = bench.trace_random(20, 16, ts.chacha20_u64(bytes([2]) * 32, 3 << 20))
pa = ctx.poly_from_u64(addr, n); pv = ctx.poly_upload_padded(ts.fe_vec(vals_u64), n)
lib = B.lib()
arr = (C.c_void_p * 2)(pa._h, pv._h)
outs = np.empty((2, 12), dtype=np.uint64); vals = np.empty((2, 4), dtype=np.uint64)
z = ts.fe_vec(np.array([0x123456789abcdef], dtype=np.uint64))
def run(name, fn, reps=5):
    for _ in range(2): fn()
    ctx.set_tuning("kernel_timing", 1); ctx.timer_reset()
    for _ in range(reps): fn()
    r = {k: round(ctx.timer_read(k)[0] / reps, 3) for k in ("msm_total", "msm_sort", "msm_accumulate", "msm_merge", "msm_reduce", "open_bary")}
    ctx.set_tuning("kernel_timing", 0)
    print(name, json.dumps(r), flush=True)
import hashlib
for quad in ([int(a) for a in sys.argv[1:]] or [1]):
  ctx.set_tuning("msm_quad_tree", quad)
  print("msm_quad_tree =", quad)
  run("commit pass", lambda: ctx.check(lib.tsgpu_kzg_commit_values_batch_dev(ctx._h, pp.srs._h, arr, C.c_size_t(2), B._p(outs))))
  h1 = hashlib.sha256(outs.tobytes()).hexdigest()[:16]
  run("open pass  ", lambda: ctx.check(lib.tsgpu_kzg_open_values_batch_dev(ctx._h, pp.srs._h, arr, C.c_size_t(2), B._p(z), B._p(vals), B._p(outs))))
  print("result hashes", h1, hashlib.sha256(outs.tobytes() + vals.tobytes()).hexdigest()[:16])

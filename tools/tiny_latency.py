import importlib, sys, time, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ts = importlib.import_module("multilinear-map-cryptography_b200")
ctx = ts.Context(0)
for log_size, n in ((4, 8), (6, 32), (10, 128), (10, 4096)):
    t0 = time.perf_counter(); pp, vp = ts.setup_params(ctx, log_size); t1 = time.perf_counter()
    tw = ts.Twist.new(pp)
    addr = (np.arange(n) % (1 << log_size)).astype(np.uint64); vals = ts.fe_vec(np.arange(n, dtype=np.uint64) * 42)
    ts_ = []
    for it in range(4):
        a = time.perf_counter(); l0 = ctx.launch_count; p = tw.prove_arrays(addr, vals); ts_.append((round((time.perf_counter() - a) * 1e3, 3), ctx.launch_count - l0))
    ctx.set_tuning("eval_basis", 0)
    tc = []
    for it in range(3):
        a = time.perf_counter(); l0 = ctx.launch_count; p2 = tw.prove_arrays(addr, vals); tc.append((round((time.perf_counter() - a) * 1e3, 3), ctx.launch_count - l0))
    ctx.set_tuning("eval_basis", 1)
    assert p.to_bytes() == p2.to_bytes()
    a = time.perf_counter(); ok = tw.verify(p, vp); tv = time.perf_counter() - a
    print(f"log_size {log_size} n {n}: setup {1e3*(t1-t0):.1f} ms; prove (ms, launches) eval-basis {ts_}; coefficient {tc}; verify {1e3*tv:.1f} ms", flush=True)

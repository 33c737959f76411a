"""Per-call wall times of Twist::prove on both paths, to spot host-side stalls (allocator, event creation, ...)."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
ts = importlib.import_module("multilinear-map-cryptography_b200")
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = ts.Context(0, stream.cuda_stream)
pp, vp = ts.setup_params(ctx, 18)
n = 1 << 20
addr, vals_u64, isw = bench.trace_random(20, 16, ts.chacha20_u64(bytes([2]) * 32, 3 << 20))
vals = ts.fe_vec(vals_u64)
addr_pin = torch.empty(n, dtype=torch.int64, pin_memory=True); addr_pin.numpy().view(np.uint64)[:] = addr
vals_pin = torch.empty((n, 4), dtype=torch.int64, pin_memory=True); vals_pin.numpy().view(np.uint64)[:] = vals
addr_h = addr_pin.numpy().view(np.uint64); vals_h = vals_pin.numpy().view(np.uint64)
tw = ts.Twist.new(pp)
def series(name, f, reps):
    out = []
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize(); out.append((time.perf_counter() - t0) * 1e3); del r
    print(name, " ".join("%.1f" % x for x in out), flush=True)
a = ctx.poly_from_u64(addr_h); v = ctx.poly_upload_padded(vals_h, n)
for eb in (0, 1):
    for timing in (0, 1):
        ctx.set_tuning("eval_basis", eb); ctx.set_tuning("kernel_timing", timing)
        series(f"eval_basis={eb} timing={timing} prove_device", lambda: tw.prove_device(a.clone(), v.clone()), 30)
        cl = [(a.clone(), v.clone()) for _ in range(30)]
        it = iter(cl)
        series(f"eval_basis={eb} timing={timing} preclone    ", lambda: tw.prove_device(*next(it)), 30)
        del cl, it
        ctx.timer_reset()

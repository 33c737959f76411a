"""Wall-clock breakdown of the Twist::prove host path at 2^20 ops (scratch tool)."""
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
def T(f, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best * 1e3, r
print("prove_arrays pinned   %.1f ms" % T(lambda: tw.prove_arrays(addr_h, vals_h, isw))[0])
print("prove_arrays pageable %.1f ms" % T(lambda: tw.prove_arrays(addr, vals, isw))[0])
print("poly_from_u64         %.2f ms" % T(lambda: ctx.poly_from_u64(addr_h))[0])
print("poly_upload_padded    %.2f ms" % T(lambda: ctx.poly_upload_padded(vals_h, n))[0])
a = ctx.poly_from_u64(addr_h); v = ctx.poly_upload_padded(vals_h, n)
print("clone x2              %.2f ms" % T(lambda: (a.clone(), v.clone()))[0])
print("prove_device          %.1f ms" % T(lambda: tw.prove_device(a.clone(), v.clone()))[0])
print("interpolate           %.2f ms" % T(lambda: a.clone().interpolate_iota())[0])
c = a.clone().interpolate_iota()
print("commit_dev            %.2f ms" % T(lambda: ts.KZGCommitment.commit(pp.srs, c))[0])
z = vals_h[5]
print("open_dev              %.2f ms" % T(lambda: ts.KZGCommitment.open(pp.srs, c, z))[0])
cv = v.clone().interpolate_iota()
print("commit_dev(values)    %.2f ms" % T(lambda: ts.KZGCommitment.commit(pp.srs, cv))[0])
ctx.set_tuning("kernel_timing", 1); ctx.timer_reset()
ts.KZGCommitment.commit(pp.srs, c); ts.KZGCommitment.commit(pp.srs, cv)
print("timers: acc", ctx.timer_read("msm_accumulate"), "total", ctx.timer_read("msm_total"))

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
addr, vals_u64, isw = bench.synthetic_trace(20, 16, 2)
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
series("prove_arrays", lambda: tw.prove_arrays(addr_h, vals_h, isw), 25)
series("upload_only ", lambda: (ctx.poly_from_u64(addr_h), ctx.poly_upload_padded(vals_h, n)), 25)
a = ctx.poly_from_u64(addr_h); v = ctx.poly_upload_padded(vals_h, n)
series("prove_device", lambda: tw.prove_device(a.clone(), v.clone()), 25)
series("prove_arrays", lambda: tw.prove_arrays(addr_h, vals_h, isw), 25)

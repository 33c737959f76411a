"""Quick device timing of the sum-check kernels (bind / round_eval / fused) at one size.  Scratch tool;
bench.py is the contract benchmark."""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch
ts = importlib.import_module("multilinear-map-cryptography_b200")
import oracle as O

nv = int(sys.argv[1]) if len(sys.argv) > 1 else 26
tma_log = int(sys.argv[2]) if len(sys.argv) > 2 else -1      # >= 0: tables with at least 2^tma_log positions per stream use the TMA-ring kernels
torch.cuda.init()
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = ts.Context(0, stream.cuda_stream)
pf_log = int(sys.argv[3]) if len(sys.argv) > 3 else -1       # >= 0: d = 2 rounds with at least 2^pf_log positions use the warp-private prefetch kernels
ctx.set_tuning("prefetch_min_log2", pf_log)
w = O.chacha_fr_rand(bytes([4]) * 32, nv)
r = O.chacha_fr_rand(bytes([6]) * 32, 1)
t0 = time.time()
A = ctx.table_eq(w)
B = ctx.table_eq(w[::-1].copy())
ctx.synchronize()
print("generated 2 tables of 2^%d in %.2fs" % (nv, time.time() - t0), flush=True)
N = 1 << nv
res = {"nv": nv, "prefetch_min_log2": pf_log}

def timeit(fn, setup, reps=5):
    best = 1e9
    for _ in range(reps + 2):
        obj = setup()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream); fn(obj); e1.record(stream); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

# bind, d = 1
ms = timeit(lambda t: t.bind(r), lambda: A.clone())
res["bind_ms"] = ms; res["bind_gbs"] = 48.0 * N / (ms * 1e-3) / 1e9
# round eval d = 2 (includes the 128-byte D2H of the result)
def mk():
    a, b = A.clone(), B.clone()
    return ctx.sumcheck([a, b])
ms = timeit(lambda sc: sc.round_eval(), mk)
res["eval2_ms"] = ms; res["eval2_gbs"] = 64.0 * N / (ms * 1e-3) / 1e9
ms = timeit(lambda sc: sc.bind_eval(r), mk)
res["bind_eval2_ms"] = ms; res["bind_eval2_gbs"] = 96.0 * N / (ms * 1e-3) / 1e9
ms = timeit(lambda sc: sc.bind_eval(r, claim=r), mk)     # timing only: any claim value costs the same
res["bind_eval2_claim_ms"] = ms; res["bind_eval2_claim_gbs"] = 96.0 * N / (ms * 1e-3) / 1e9
# full sum-check d=2 with random challenges (no transcript) - device time only
def full(sc):
    sc.round_eval()
    while sc.vars_left > 1:
        sc.bind_eval(r, claim=r)
    sc.bind(r)
ms = timeit(full, mk, reps=3)
res["sumcheck2_ms"] = ms; res["sumcheck2_gbs"] = 256.0 * N / (ms * 1e-3) / 1e9
# evaluate
pt = O.chacha_fr_rand(bytes([9]) * 32, nv)
ms = timeit(lambda t: t.evaluate(pt), lambda: A)
res["evaluate_ms"] = ms; res["evaluate_gbs"] = 32.0 * N / (ms * 1e-3) / 1e9
print(json.dumps(res))

"""Tiny workload for `ncu --set full`: one G1 MSM of 2^18 points (precomputed window tables, one shared bucket set), one fold
(bind), one round evaluation and one fused bind+eval (claim form) of 2^24-entry tables, one barycentric opening pass.
Keep it short - ncu replays every profiled kernel ~40 times."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
ts = importlib.import_module("multilinear-map-cryptography_b200")
ctx = ts.Context(0)
tau = ts.fe(123456789)
srs = ctx.srs_generate(tau, 1 << 18)
rng = np.random.default_rng(1)
sc = rng.integers(0, 1 << 62, size=(1 << 18, 4), dtype=np.uint64)      # any limbs < r are valid Montgomery scalars
sc[:, 3] &= (1 << 60) - 1
c = ts.KZGCommitment.commit(srs, sc)
w = sc[:24].copy(); r = sc[30:31].copy()
A = ctx.table_eq(w); B = ctx.table_eq(w[::-1].copy())
A.clone().bind(r)
s = ctx.sumcheck([A, B])
s.round_eval(); s.bind_eval(r, claim=r)
pv = ctx.poly_upload(sc)
ts.KZGCommitment.open_values(srs, pv, r)
ctx.synchronize()
print("ok", ts.g1_compress(c).hex()[:16])

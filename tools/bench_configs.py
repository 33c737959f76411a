"""Single-GPU measurements of BASELINE.json configs C3 (Shout, 2^20-entry table / 2^22 lookups), C4 (stand-alone sum-check over
a 2^26-entry eq x one-hot product, C++ host loop + transcript) and C5 (G1 MSM of 2^24 points), with the inputs SURVEY 8(d)
defines.  One JSON line per config.  CUDA events on the library stream; best of a few runs after warm-up.
usage: python tools/bench_configs.py [c3] [c4] [c5] [--c4-log N] [--c5-log N]"""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch
ts = importlib.import_module("multilinear-map-cryptography_b200")
import oracle as O

args = sys.argv[1:]
def opt(name, default):
    return int(args[args.index(name) + 1]) if name in args else default
which = [a for a in args if a in ("c3", "c3check", "c4", "c4check", "c5")] or ["c3", "c3check", "c4", "c4check", "c5"]
C3_T, C3_L = opt("--c3-table-log", 20), opt("--c3-lookups-log", 22)
C4_LOG, C5_LOG = opt("--c4-log", 26), opt("--c5-log", 24)

torch.cuda.init()
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = ts.Context(0, stream.cuda_stream)
HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream); fn(); e1.record(stream); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def c3check():
    """the real lookup-correctness sum-check (non-parity mode, host/read_check.cpp) on the C3 statement: prove + verify"""
    T, L = 1 << C3_T, 1 << C3_L
    entries = ts.fe_vec(np.arange(T, dtype=np.uint64) ** 2)
    idx = O.chacha_u64(bytes([3]) * 32, L) % np.uint64(T)
    vals = np.ascontiguousarray(entries[idx.astype(np.int64)])
    rc = ts.ShoutReadCheck(ctx)
    claim, proof, _ = rc.prove_arrays(entries, idx, vals, ts.Transcript())
    assert rc.verify_arrays(entries, idx, vals, proof, ts.Transcript())
    out = {"config": "C3-read-check", "workload": f"Shout read-checking sum-check, 2^{C3_T}-entry table, 2^{C3_L} lookups (host buffers in, proof out)", "n_gpus": 1,
           "rounds": int(proof.round_polynomials.shape[0])}
    out["prove_ms"] = timed(lambda: rc.prove_arrays(entries, idx, vals, ts.Transcript()), reps=3, warm=1)
    out["verify_ms"] = timed(lambda: rc.verify_arrays(entries, idx, vals, proof, ts.Transcript()), reps=3, warm=1)
    out["h2d_bytes"] = int(entries.nbytes + idx.nbytes + vals.nbytes)
    out["lookups_per_s"] = L / (out["prove_ms"] * 1e-3)
    print(json.dumps(out), flush=True)


def c4check():
    """the real memory-consistency sum-checks (non-parity mode, host/memory_check.cpp) on the shape of config 4: 2^10 cells x 2^16 cycles =
    2^26-entry (weighted one-hot) x Val tables, 26 + 16 rounds; trace from the reference generator (src/benchmarks.rs:88-99)"""
    logK, logT = opt("--c4check-cells-log", 10), opt("--c4check-cycles-log", 16)
    K, n = 1 << logK, 1 << logT
    i = np.arange(n, dtype=np.uint64)
    isw = (i % 3 == 0).astype(np.uint8)
    addr = np.where(isw == 1, i % K, (i // 2) % K).astype(np.uint64)
    vals_u = np.zeros(n, dtype=np.uint64); mem = {}
    for j in range(n):
        a = int(addr[j])
        if isw[j]:
            mem[a] = 42 * j
        vals_u[j] = mem.get(a, 0)
    vals = ts.fe_vec(vals_u)
    mc = ts.TwistMemoryCheck(ctx)
    proof = mc.prove_arrays(addr, vals, isw, K, ts.Transcript())
    assert mc.verify_arrays(addr, vals, isw, K, proof, ts.Transcript())
    out = {"config": "C4-memory-check", "workload": f"Twist read-checking + Val-evaluation sum-checks, 2^{logK} cells x 2^{logT} cycles (2^{logK + logT}-entry tables built on the device)",
           "n_gpus": 1, "rounds": [int(proof.read_check.round_polynomials.shape[0]), int(proof.val_evaluation.round_polynomials.shape[0])]}
    out["prove_ms"] = timed(lambda: mc.prove_arrays(addr, vals, isw, K, ts.Transcript()), reps=3, warm=1)
    out["verify_ms"] = timed(lambda: mc.verify_arrays(addr, vals, isw, K, proof, ts.Transcript()), reps=3, warm=1)
    print(json.dumps(out), flush=True)


def c3():
    """Shout::prove: table i^2 (src/benchmarks.rs:167-169), 2^22 lookups at ChaCha20 seed [3; 32] indices"""
    T, L = 1 << C3_T, 1 << C3_L
    t0 = time.perf_counter()
    pp, vp = ts.setup_params(ctx, C3_L - 2)
    setup_s = time.perf_counter() - t0
    entries = ts.fe_vec(np.arange(T, dtype=np.uint64) ** 2)
    idx = O.chacha_u64(bytes([3]) * 32, L) % np.uint64(T)
    shout = ts.Shout.new(pp)
    proof = shout.prove_arrays(entries, idx)
    assert shout.verify(proof, vp)
    out = {"config": "C3", "workload": f"Shout::prove, 2^{C3_T}-entry table, 2^{C3_L} lookups, setup_params({C3_L - 2})", "n_gpus": 1, "setup_s": setup_s,
           "proof_bytes": len(proof.to_bytes())}
    out["prove_ms_e2e_host_buffers"] = timed(lambda: shout.prove_arrays(entries, idx), reps=3, warm=1)
    ctx.set_tuning("eval_basis", 0)
    pc = shout.prove_arrays(entries, idx)
    assert pc.to_bytes() == proof.to_bytes()
    out["coefficient_path_ms"] = timed(lambda: shout.prove_arrays(entries, idx), reps=2, warm=1)
    ctx.set_tuning("eval_basis", 1)
    out["lookups_per_s"] = L / (out["prove_ms_e2e_host_buffers"] * 1e-3)
    print(json.dumps(out), flush=True)


def c4():
    """SumCheck over A = eq(w, .), B = one-hot K = 2^10 x T = 2^(n-10): w and the addresses from ChaCha20 seed [4; 32]"""
    nv = C4_LOG; logK = 10 if nv > 10 else 1
    rows = 1 << (nv - logK)
    w, addr = O.chacha_fr_then_u64(bytes([4]) * 32, nv, rows)
    addr = addr % np.uint64(1 << logK)
    A = ctx.table_eq(w.reshape(nv, 4)); B = ctx.table_one_hot_rows(addr, logK, nv)
    sc = ctx.sumcheck([A.clone(), B.clone()]); ev = sc.round_eval(); sc.end()
    dd = importlib.import_module("multilinear-map-cryptography_b200.distributed")
    claimed = dd.fr_add(ev[0], ev[1])
    proof = ts.SumCheck(nv, claimed).prove_product(ctx, [A.clone(), B.clone()], ts.Transcript())
    ok, _ = ts.SumCheck(nv, claimed).verify(proof, ts.Transcript())
    assert ok
    clones = [[A.clone(), B.clone()] for _ in range(6)]
    it = iter(clones)
    ms = timed(lambda: ts.SumCheck(nv, claimed).prove_product(ctx, next(it), ts.Transcript()), reps=4, warm=2)
    N = 1 << nv
    gbs = 128.0 * 2 * N / (ms * 1e-3) / 1e9
    print(json.dumps({"config": "C4", "workload": f"sum-check, 2 tables of 2^{nv} entries (eq x one-hot), {nv} rounds, host transcript", "n_gpus": 1, "ms": ms,
                      "algorithmic_GBps": gbs, "frac_of_measured_hbm": gbs / HBM, "hbm_peak_gbs": HBM, "algorithmic_bytes": 256.0 * N}), flush=True)


def c5():
    """G1 MSM of 2^24 points: bases = g1_powers of setup_params (tau from seed [42; 32]), scalars Fr::rand from seed [5; 32]"""
    n = 1 << C5_LOG
    tau, _ = O.setup_scalars()
    t0 = time.perf_counter()
    srs = ctx.srs_generate(tau, n)
    setup_s = time.perf_counter() - t0
    sc = O.chacha_fr_rand(bytes([5]) * 32, n).reshape(n, 4)
    poly = ctx.poly_upload(sc)
    c = ts.KZGCommitment.commit(srs, poly)
    ctx.set_tuning("kernel_timing", 1)
    e0 = ctx.counter("msm_entries")
    ms = timed(lambda: ts.KZGCommitment.commit(srs, poly), reps=3, warm=1)
    calls = 4
    acc, cnt = ctx.timer_read("msm_accumulate")
    ph = {k: ctx.timer_read("msm_" + k)[0] / max(cnt, 1) for k in ("sort", "merge", "reduce")}
    ctx.set_tuning("kernel_timing", 0)
    entries = (ctx.counter("msm_entries") - e0) / calls
    acc_ms = acc / max(cnt, 1)
    print(json.dumps({"config": "C5", "workload": f"G1 MSM, 2^{C5_LOG} points, full-width scalars", "n_gpus": 1, "ms": ms, "points_per_s": n / (ms * 1e-3), "srs_setup_s": setup_s,
                      "accumulate_ms": acc_ms, "sort_ms": ph["sort"], "merge_ms": ph["merge"], "reduce_ms": ph["reduce"], "bucket_entries": entries,
                      "accumulate_TIMAD_s": entries * 1360 / (acc_ms * 1e-3) / 1e12 if acc_ms else None, "commitment": ts.g1_compress(c).hex()[:16]}), flush=True)


for name in which:
    {"c3": c3, "c3check": c3check, "c4": c4, "c4check": c4check, "c5": c5}[name]()
ctx.close()

"""MSM micro-benchmark on one GPU: full-width and small-scalar MSMs over a generated SRS, per-phase device times.
usage: python tools/msm_bench.py [log_n ...]"""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
ts = importlib.import_module("multilinear-map-cryptography_b200")

def main():
    logs = [int(a) for a in sys.argv[1:]] or [20]
    ctx = ts.Context(0)
    import oracle as O
    tau, _ = O.setup_scalars()
    out = []
    for lg in logs:
        n = 1 << lg
        srs = ctx.srs_generate(tau, n)
        rng = np.random.default_rng(lg)
        cases = {
            "full_width": ctx.poly_upload(O.chacha_fr_rand(bytes([5]) * 32, n).reshape(n, 4)) if lg <= 22 else None,
            "u16": ctx.poly_from_u64(rng.integers(0, 1 << 16, size=n, dtype=np.uint64), n),
            "u63": ctx.poly_from_u64(rng.integers(0, 1 << 63, size=n, dtype=np.uint64), n),
            "ones": ctx.poly_from_u64(np.ones(n, dtype=np.uint64), n),
        }
        if cases["full_width"] is None:
            # large sizes: full-width scalars generated on the device path as products (u63 * u63 stays random enough for timing)
            cases["full_width"] = ctx.poly_upload(np.ascontiguousarray(rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)))
        for name, poly in cases.items():
            for _ in range(2):
                ts.KZGCommitment.commit(srs, poly)
            ctx.set_tuning("kernel_timing", 1); ctx.timer_reset()
            c0 = ctx.counter("msm_entries")
            t0 = time.perf_counter()
            reps = 5
            for _ in range(reps):
                ts.KZGCommitment.commit(srs, poly)
            ctx.synchronize()
            wall = (time.perf_counter() - t0) / reps * 1e3
            tot, _ = ctx.timer_read("msm_total"); acc, _ = ctx.timer_read("msm_accumulate")
            phases = {k: ctx.timer_read("msm_" + k)[0] / reps for k in ("sort", "merge", "reduce")}
            ctx.set_tuning("kernel_timing", 0)
            ent = (ctx.counter("msm_entries") - c0) / reps
            out.append({"log_n": lg, "scalars": name, "wall_ms": wall, "msm_total_ms": tot / reps, "accumulate_ms": acc / reps, "sort_ms": phases["sort"], "merge_ms": phases["merge"], "reduce_ms": phases["reduce"], "entries": ent,
                        "points_per_s": n / (wall * 1e-3), "acc_TIMAD_s": ent * 1360 / (acc / reps * 1e-3) / 1e12 if acc else None})
            print(json.dumps(out[-1]), flush=True)
        del cases, srs
    ctx.close()

if __name__ == "__main__":
    main()

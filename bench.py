#!/usr/bin/env python
"""bench.py - headline benchmark of the B200 prover backend.

Workload (BASELINE.json configs[1], "C2"): Twist::prove over a synthetic MemoryTrace of 2^20 random read/write operations on 2^16
memory cells, setup_params(18) (prove rejects more than 4 * 2^log_size operations, src/twist.rs:108).  The trace is SURVEY 8(d)'s
distribution B: ChaCha20Rng::from_seed([2; 32]), per operation three next_u64 draws a, b, c -> address a mod 2^16, write iff b & 1,
written value Fr::from(c); reads return the simulated memory content.  A "step" is one Twist::prove.

    python bench.py --gpus N --steps K --warmup W            our arm (torchrun launches N ranks for N > 1)
    python bench.py --impl reference ...                     CPU arm: the oracle port of the reference prover on the SAME trace

One JSON line on stdout (rank 0).
 N = 1: `value` = ms per proof with the padded vectors resident in HBM; `e2e.value` = ms per proof through the public API from pinned
        host memory (H2D of the trace + D2H of the proof inside the timed region).  Beside the headline the line carries distribution A
        (the generator of src/benchmarks.rs:88-99), full-width values, the coefficient path, and BASELINE configs C3 / C4 / C5 (`configs`),
        each with its own roofline block, and the CPU baseline: the oracle's prover on the same 2^20-op trace, all host cores.
 N > 1: `value` = ms for ONE proof of the same trace sharded over the N ranks ("scaling": "strong"; byte-identical to the one-GPU proof,
        asserted on rank 0 before timing); `replicas` = one independent proof per GPU; `configs` = C3 sharded, C4 hypercube-sharded,
        C5 point-sharded.
"""
import argparse
import importlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "multilinear-map-cryptography_b200"

LOG_OPS = int(os.environ.get("TSGPU_BENCH_LOG_OPS", "20"))      # the env override exists for the CPU test of the contract line only
LOG_CELLS = 16
LOG_SIZE = max(LOG_OPS - 2, 0)  # setup_params(18): max_operations = 2^20
FOLD_LOG = 26                   # BASELINE metric "sumcheck fold GB/s": bind of one 2^26-entry table
JSON_OUT = sys.stdout
IMAD_PEAK_TOPS = 18.5           # measured on this pool with tools/ubench.cu (plain IMAD.WIDE issue rate, profiles/r01_ubench_imad_forms.json)
METRIC = "twist_prove_ms_at_2^20_ops"
R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617
# identical in both arms (the driver compares `config` of the two lines)
CONFIG = {
    "workload": f"Twist::prove, 2^{LOG_CELLS} cells, 2^{LOG_OPS} random read/write ops, setup_params({LOG_SIZE})",
    "distribution": "B (SURVEY 8d): ChaCha20Rng::from_seed([2; 32]); per op next_u64 a, b, c: address a mod 2^16, write iff b & 1, value Fr::from(c); reads return the simulated memory",
    "l2": "per-step working set (2 x 32 MiB vectors, 2 x 64 MiB basis points, 0.8 GiB window tables, ~0.2 GiB MSM scratch) exceeds the 126 MB L2; no explicit flush",
}


# ------------------------------------------------------------------------------------------------ synthetic traces
def simulate_memory(addr, is_write, fresh):
    """values[j] = the value operation j carries: the written value for a write, the last value written to the address (0 if none) for a
    read - MemoryTrace::read / write (src/twist.rs:41-70), vectorised"""
    n = addr.shape[0]
    order = np.lexsort((np.arange(n), addr))
    a_s, w_s, f_s = addr[order], is_write[order], fresh[order]
    idx = np.where(w_s == 1, np.arange(n), -1)
    seg_start = np.r_[True, a_s[1:] != a_s[:-1]]
    first_of_seg = np.flatnonzero(seg_start)[np.cumsum(seg_start) - 1]
    last_write = np.maximum.accumulate(idx)
    valid = last_write >= first_of_seg
    vals_sorted = np.where(valid, f_s[np.clip(last_write, 0, n - 1)], 0).astype(np.uint64)
    values = np.empty(n, dtype=np.uint64)
    values[order] = vals_sorted
    return values


def trace_random(log_ops, log_cells, stream):
    """distribution B from 3 n u64 draws of ChaCha20Rng::from_seed([2; 32]) (`stream`): (addresses, u64 values, is_write)"""
    n = 1 << log_ops
    s = np.asarray(stream, dtype=np.uint64).reshape(n, 3)
    addr = s[:, 0] % np.uint64(1 << log_cells)
    is_write = (s[:, 1] & np.uint64(1)).astype(np.uint8)
    return addr, simulate_memory(addr, is_write, s[:, 2].copy()), is_write


def trace_ref_pattern(log_ops, log_cells):
    """distribution A, the reference's own benchmark generator (src/benchmarks.rs:88-99): op i writes 42 i to cell i mod M when i % 3 == 0,
    else reads cell (i / 2) mod M"""
    n = 1 << log_ops
    i = np.arange(n, dtype=np.uint64)
    is_write = (i % np.uint64(3) == 0).astype(np.uint8)
    M = np.uint64(1 << log_cells)
    addr = np.where(is_write == 1, i % M, (i // np.uint64(2)) % M).astype(np.uint64)
    return addr, simulate_memory(addr, is_write, i * np.uint64(42)), is_write


class ClockSampler:
    """SM clock and throttle reasons while the timed regions run, sampled in-process through NVML (nvidia_ml_py);
    spawning `nvidia-smi -lms` instead stalled the device for hundreds of ms at unpredictable times."""

    def __init__(self, index: int):
        self.index = index; self.samples = []; self.reasons = set(); self.max_mhz = None
        self._stop = threading.Event(); self.t = None; self.err = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            import torch
            props = torch.cuda.get_device_properties(self.index)
            bus = props.pci_bus_id if hasattr(props, "pci_bus_id") else None
            self.h = None
            if bus is not None:                      # honour CUDA_VISIBLE_DEVICES remapping through the PCI bus id of the torch device
                for i in range(pynvml.nvmlDeviceGetCount()):
                    h = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if pynvml.nvmlDeviceGetPciInfo(h).bus == bus:
                        self.h = h
            if self.h is None:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._run, daemon=True); self.t.start()
        except Exception as e:   # noqa: BLE001
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception as e:   # noqa: BLE001
                self.err = repr(e)
            self._stop.wait(0.02)

    def mark(self):
        """forget samples taken so far (called when the timed region starts)"""
        self.samples = []; self.reasons = set()

    def stop(self):
        self._stop.set()
        if self.t:
            self.t.join(timeout=1)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable: " + str(self.err)]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "samples": len(self.samples), "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_prover(threads: int, steps: int, warmup: int, budget_s: float):
    """The oracle's arkworks-class CPU prover (oracle/oracle.cpp, fast tier: NTT interpolation, Pippenger MSM with ark-ec's window rule,
    all host threads) on the SAME 2^LOG_OPS-op trace as the GPU arm - no sampling, no extrapolation.  Runs `warmup` untimed and up to `steps`
    timed proofs, stopping early once `budget_s` of timed work is spent (a proof takes ~20-30 s).  -> (mean ms, steps run, warmup run, note)"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    O.build()
    n = 1 << LOG_OPS
    addr, vals_u64, isw = trace_random(LOG_OPS, LOG_CELLS, O.chacha_u64(bytes([2]) * 32, 3 * n))
    vals = O.fr_from_u64(vals_u64)
    t0 = time.perf_counter()
    powers = O.setup_g1_powers(n + 1, fast=True, threads=threads)
    setup_s = time.perf_counter() - t0
    times = []
    spent = 0.0
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.twist_prove(powers, n, addr, vals, isw, fast=True, threads=threads)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt); spent += dt
            if spent + dt > budget_s:
                break
    ms = 1e3 * float(np.mean(times))
    return ms, len(times), warmup, (f"oracle fast tier (NTT interpolation + Pippenger, {threads} threads): {len(times)} full Twist::prove of the 2^{LOG_OPS}-op trace "
                                    f"({ms:.0f} ms each; SRS setup {setup_s:.1f} s outside the timed region)")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    cores = O.ncpu()
    value, steps, warmup, sample = cpu_prover(cores, max(1, args.steps), min(args.warmup, 1), budget_s=float(os.environ.get("TSGPU_REF_BUDGET_S", "120")))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "ms", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": value, "higher_is_better": False, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None,
        "dtype": "u64x4 (BN254 Fr/Fq, 4x64-bit Montgomery)", "data": "synthetic", "config": CONFIG,
        "note": "the Rust reference cannot be built here (no toolchain) and its O(n^3) interpolation cannot reach 2^20 operations; this arm is the oracle's quasi-linear CPU "
                "prover (kind \"port\") producing the same proof bytes; steps are capped by a time budget, `steps` / `warmup` are the numbers actually run",
        "cpu_baseline": {"value": value, "unit": "ms", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=JSON_OUT, flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the C3 / C4 / C5 side configs")
    ap.add_argument("--no-fold", action="store_true")
    args = ap.parse_args()
    # stdout carries exactly ONE line, the JSON: keep a private handle to it and point fd 1 at stderr for everything else that writes there
    # (NCCL prints its version line to stdout from inside the library communicator at N > 1, whatever NCCL_DEBUG_FILE says)
    global JSON_OUT
    sys.stdout.flush()
    JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ts = importlib.import_module(PKG)
    dd = importlib.import_module(PKG + ".distributed")
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = ts.Context(local, stream.cuda_stream)
    W = max(args.warmup, 3); K = max(args.steps, 1)
    KS = min(K, 5)                                               # side measurements: at most 5 timed steps each
    for kv in filter(None, os.environ.get("TSGPU_TUNING", "").split(",")):      # A/B switches for experiments, e.g. TSGPU_TUNING=eval_basis=0
        key, val = kv.split("=")
        ctx.set_tuning(key, int(val))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:   # noqa: BLE001
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0)); hbm_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(*xs):
        if world == 1:
            return [float(x) for x in xs]
        t = torch.tensor(list(xs), device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)

    def timed(fn, steps, warm):
        """CUDA events on the library stream around `steps` calls after `warm` untimed ones, barrier + synchronize on both sides: ms per call (this rank)"""
        for _ in range(warm):
            fn()
        barrier()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        return e0.elapsed_time(e1) / steps

    def pinned(arr):
        a = np.ascontiguousarray(arr, dtype=np.uint64)
        t = torch.empty(a.shape, dtype=torch.int64, pin_memory=True)
        t.numpy().view(np.uint64)[...] = a
        return t, t.numpy().view(np.uint64)

    sampler = ClockSampler(local); sampler.start()
    t0 = time.perf_counter()
    pp, vp = ts.setup_params(ctx, LOG_SIZE)                     # SRS of 2^20 + 1 points generated on the device
    setup_s = time.perf_counter() - t0
    twist = ts.Twist.new(pp)
    n = 1 << LOG_OPS
    addr, vals_u64, isw = trace_random(LOG_OPS, LOG_CELLS, ts.chacha20_u64(bytes([2]) * 32, 3 * n))   # the same trace on every rank
    keep = []                                                   # pinned tensors stay alive while their numpy views are in use
    t_, addr_h = pinned(addr); keep.append(t_)
    t_, vals_h = pinned(ts.fe_vec(vals_u64)); keep.append(t_)

    # ---- correctness guard: the proof must verify (transcript + sum-check + openings) before anything is timed
    proof = twist.prove_arrays(addr_h, vals_h, isw)
    assert twist.verify(proof, vp), "benchmark proof does not verify"
    proof_bytes = proof.to_bytes()

    def time_twist(a_h, v_h, w_flags, steps):
        """-> (e2e ms, e2e launches, device-resident result dict): one Twist::prove per step"""
        l0 = [0]

        def step_e2e():
            p = twist.prove_arrays(a_h, v_h, w_flags)
            _ = p.final_evaluation                               # the D2H'd proof contents are read
        e2e = timed(step_e2e, steps, W)
        l0[0] = ctx.launch_count
        step_e2e()
        e2e_launches = ctx.launch_count - l0[0]
        base_a = ctx.poly_from_u64(a_h); base_v = ctx.poly_upload_padded(v_h, a_h.shape[0])
        # every input copy is made BEFORE the warm-up, so that the stream-ordered pool has its final size when the timed region starts
        clones = [(base_a.clone(), base_v.clone()) for _ in range(W + steps)]
        ctx.set_tuning("kernel_timing", 1)
        for a, v in clones[:W]:
            twist.prove_device(a, v)
        ctx.timer_reset()
        barrier()
        c0 = {k: ctx.counter(k) for k in ("launches", "msm_calls", "msm_points", "msm_entries")}
        e0.record(stream)
        for a, v in clones[W:]:
            twist.prove_device(a, v)
        e1.record(stream)
        barrier()
        r = {"ms": e0.elapsed_time(e1) / steps, "steps": steps}
        r.update({k: ctx.counter(k) - c0[k] for k in c0})
        for name in ("msm_accumulate", "msm_total", "msm_sort", "msm_merge", "msm_reduce", "interpolate", "open_bary", "open_scan"):
            r[name + "_ms"], r[name + "_cnt"] = ctx.timer_read(name)
        ctx.set_tuning("kernel_timing", 0)
        return e2e, int(e2e_launches), r

    # ---- roofline of k_msm_accumulate (integer-pipe bound), the dominant kernel of the step.  Two accountings, both reported:
    #  * `achieved` (the contract's): SURVEY 8(d)'s per-unit figure x the units of a launch / the launch duration.  The figure is
    #    16 windows x 11 Fq products x 136 IMAD = 23 936 IMAD per point of a full-width MSM; full-width scalars occur in the open pass (two
    #    quotient vectors of 2^20 points each per launch), so achieved = 2 n x 23 936 / t(open-pass share of the accumulate time), the share
    #    taken by bucket entries (13 per full-width point).
    #  * `executed`: what the kernel really does - one mixed XYZZ + affine addition (8M + 2S; the Y coordinate's two products share one
    #    reduction: 1288 multiply-adds) per bucket entry, entries counted by the library.
    imad_per_add = 8 * (2 * 8 * 8 + 8) + (2 * 64 + 72)
    survey_imad_per_point = 16 * 11 * 136

    def acc_roofline(r, full_width_points_per_step):
        steps = r["steps"]
        launches = max(r["msm_accumulate_cnt"], 1)
        launch_ms = r["msm_accumulate_ms"] / launches
        entries = r["msm_entries"] / max(r["msm_calls"], 1)
        executed = entries * imad_per_add / (launch_ms * 1e-3) / 1e12 if launch_ms > 0 else 0.0
        fw_entries = 13.0 * full_width_points_per_step * steps
        fw_ms = r["msm_accumulate_ms"] * min(1.0, fw_entries / max(r["msm_entries"], 1))
        ach = full_width_points_per_step * steps * survey_imad_per_point / (fw_ms * 1e-3) / 1e12 if fw_ms > 0 else 0.0
        return {"kernel": "k_msm_accumulate", "bound": "int32-pipe", "achieved": ach, "peak": IMAD_PEAK_TOPS, "unit": "TIMAD/s",
                "frac": ach / IMAD_PEAK_TOPS, "traffic": None, "launch_ms": launch_ms, "launches": r["msm_accumulate_cnt"],
                "executed": executed, "executed_frac": executed / IMAD_PEAK_TOPS, "entries_per_launch": entries,
                "share_of_step": r["msm_accumulate_ms"] / steps / r["ms"] if r["ms"] > 0 else None}

    def breakdown(r):
        s = r["steps"]
        return {"msm_4x": r["msm_total_ms"] / s, "msm_accumulate_4x": r["msm_accumulate_ms"] / s, "msm_sort_4x": r["msm_sort_ms"] / s,
                "msm_chunk_merge_4x": r["msm_merge_ms"] / s, "msm_window_reduce_4x": r["msm_reduce_ms"] / s, "interpolate_2x": r["interpolate_ms"] / s,
                "open_barycentric_2x": r["open_bary_ms"] / s, "open_scan_2x": r["open_scan_ms"] / s, "bucket_entries": r["msm_entries"] // s,
                "gpu_launches": r["launches"] // s}

    line = {"metric": METRIC, "unit": "ms", "n_gpus": world, "steps": K, "warmup": W, "higher_is_better": False, "vs_baseline": None,
            "dtype": "u32x8 (BN254 Fr/Fq, 256-bit Montgomery on the integer pipe)", "data": "synthetic", "config": CONFIG,
            "path": "default: evaluation-basis SRS (commit / open from the values, no interpolation); byte-identical to the coefficient path (asserted in the run)",
            "proof_bytes": len(proof_bytes), "setup_s": setup_s}

    if world == 1:
        # ---- coefficient path (tuning eval_basis = 0): interpolate -> commit -> open on coefficients, the reference's own sequence of steps
        ctx.set_tuning("eval_basis", 0)
        proof_c = twist.prove_arrays(addr_h, vals_h, isw)
        coef_e2e, _, coef = time_twist(addr_h, vals_h, isw, KS)
        ctx.set_tuning("eval_basis", 1)
        assert proof_bytes == proof_c.to_bytes(), "the two paths must give identical proof bytes"
        # ---- headline
        sampler.mark()
        e2e_ms, e2e_launches, dflt = time_twist(addr_h, vals_h, isw, K)
        clocks = sampler.stop()
        dev_ms = dflt["ms"]
        roofline = acc_roofline(dflt, 2 * n)           # the two opening quotients are the full-width scalars of the default path
        roofline["peak_source"] = "tools/ubench.cu on this pool: 18.5 T IMAD.WIDE/s without carry-in (the carry-chained form the multiplier needs issues at half that rate); MEASURED_PEAKS.json has no integer figure"
        roofline["algorithmic_unit"] = "SURVEY 8(d): 16 windows x 11 Fq products x 136 IMAD = 23 936 IMAD per full-width point; `executed` = bucket entries x 1288 multiply-adds"
        # DRAM bytes per launch from `ncu --set full` of this command (profiles/r02_ncu_full_summary.md, tools/profile_round2.sh): commit pass 0.493 GB read +
        # 0.017 GB written, open pass 3.625 + 0.146 GB; mean over the two launches of a proof, like `launch_ms`.  Algorithmic bytes: entries x (64 B point + 4 B entry).
        roofline["traffic"] = 0.5 * ((0.493 + 0.017) + (3.625 + 0.146)) * 1e9
        roofline["algorithmic_bytes"] = roofline["entries_per_launch"] * 68.0
        line.update({"value": dev_ms, "ms_per_step": dev_ms, "scaling": "weak", "clocks": clocks,
                     "e2e": {"value": e2e_ms, "unit": "ms", "h2d_bytes_per_step": int(addr_h.nbytes + vals_h.nbytes), "d2h_bytes_per_step": int(4 * 16 * 96 + 2 * 32),
                             "gpu_launches": e2e_launches},
                     "gpu_launches": int(dflt["launches"] // K), "roofline": roofline, "breakdown_ms_per_step": breakdown(dflt),
                     "coefficient_path": {"value": coef["ms"], "unit": "ms", "e2e": coef_e2e, "breakdown_ms_per_step": breakdown(coef), "roofline": acc_roofline(coef, 4 * n),
                                          "note": "tsgpu_set_tuning(eval_basis, 0): 2 interpolations + 4 full-width MSMs + 2 Horner/quotient scans (what an SRS without its trapdoor runs)"},
                     "msm_points_per_s": dflt["msm_points"] / (dflt["msm_total_ms"] * 1e-3) if dflt["msm_total_ms"] > 0 else None,
                     "msm_full_width_points_per_s": coef["msm_points"] / (coef["msm_total_ms"] * 1e-3) if coef["msm_total_ms"] > 0 else None,
                     "ops_per_s": n / (dev_ms * 1e-3)})
        # ---- the same step on distribution A (the reference's own generator) and with full-width written values (no short-scalar commit pass)
        a_addr, a_vals, a_isw = trace_ref_pattern(LOG_OPS, LOG_CELLS)
        t_, a_addr_h = pinned(a_addr); keep.append(t_)
        t_, a_vals_h = pinned(ts.fe_vec(a_vals)); keep.append(t_)
        assert twist.verify(twist.prove_arrays(a_addr_h, a_vals_h, a_isw), vp)
        a_e2e, _, a_dev = time_twist(a_addr_h, a_vals_h, a_isw, KS)
        line["distribution_A"] = {"value": a_dev["ms"], "e2e": a_e2e, "unit": "ms", "what": "src/benchmarks.rs:88-99: write 42 i to cell i mod 2^16 when i % 3 == 0, else read cell (i / 2) mod 2^16",
                                  "breakdown_ms_per_step": breakdown(a_dev)}
        wide, _ = ts.chacha20_fr_then_u64(bytes([6]) * 32, n)                     # uniform Fr values: every scalar of the commit pass is full width
        src = simulate_memory(addr, isw, np.arange(1, n + 1, dtype=np.uint64))   # 1-based index of the write each operation's value comes from (0: never written)
        w_vals = np.where((src > 0)[:, None], wide[np.maximum(src, 1).astype(np.int64) - 1], np.uint64(0))
        t_, w_vals_h = pinned(w_vals); keep.append(t_)
        assert twist.verify(twist.prove_arrays(addr_h, w_vals_h, isw), vp)
        w_e2e, _, w_dev = time_twist(addr_h, w_vals_h, isw, KS)
        line["full_width_values"] = {"value": w_dev["ms"], "e2e": w_e2e, "unit": "ms", "what": "same addresses, written values uniform in Fr (Fr::rand, seed [6; 32])",
                                     "breakdown_ms_per_step": breakdown(w_dev)}
        del a_dev, w_dev
    else:
        # ---- N > 1: ONE proof of the same trace sharded over the ranks (evaluation-basis slices per rank, three small all-gathers over the library's
        # NCCL communicator).  Byte identity with the one-GPU proof is asserted on every rank before timing.
        ctx.comm_init_torch()
        lo, hi = twist.shard_range(n, rank, world)
        t_, a_s = pinned(addr[lo:hi]); keep.append(t_)
        t_, v_s = pinned(ts.fe_vec(vals_u64[lo:hi])); keep.append(t_)
        ps = twist.prove_sharded(a_s, v_s, n)
        assert ps.to_bytes() == proof_bytes, "the sharded proof must equal the one-GPU proof byte for byte"
        m_loc = n // world
        base_a = ctx.poly_from_u64(a_s, m_loc); base_v = ctx.poly_upload_padded(v_s, m_loc)
        assert twist.prove_sharded_device(base_a, base_v, n).to_bytes() == proof_bytes
        sampler.mark()

        def step_sharded_e2e():
            p = twist.prove_sharded(a_s, v_s, n)
            _ = p.final_evaluation
        e2e_local = timed(step_sharded_e2e, K, W)
        l0 = ctx.launch_count
        dev_local = timed(lambda: twist.prove_sharded_device(base_a, base_v, n), K, W)
        launches = (ctx.launch_count - l0) // (K + W)
        clocks = sampler.stop()
        # one independent proof per GPU (replicas: weak scaling, no data-path collective)
        rep_local = timed(lambda: twist.prove_arrays(addr_h, vals_h, isw), KS, W)
        dev_ms, e2e_ms, rep_ms = max_over_ranks(dev_local, e2e_local, rep_local)
        line.update({"value": dev_ms, "ms_per_step": dev_ms, "scaling": "strong", "clocks": clocks,
                     "what": f"ONE Twist::prove of the 2^{LOG_OPS}-op trace sharded over {world} ranks (tsgpu_twist_prove_sharded[_dev]); proof bytes == the one-GPU proof (asserted on every rank)",
                     "e2e": {"value": e2e_ms, "unit": "ms", "h2d_bytes_per_step": int(a_s.nbytes + v_s.nbytes), "d2h_bytes_per_step": int(4 * 16 * 96 + 2 * 32),
                             "note": "bytes per rank; every rank uploads its slice of the trace"},
                     "gpu_launches": int(launches), "ops_per_s": n / (dev_ms * 1e-3),
                     "exchange": "peer mailboxes over NVLink (single-kernel all-gathers)" if ctx.comm_peer_exchange else "ncclAllGather",
                     "replicas": {"value": rep_ms, "unit": "ms", "scaling": "weak", "what": "one independent proof per GPU from host buffers", "ops_per_s_all_gpus": world * n / (rep_ms * 1e-3)},
                     "roofline": {"kernel": "k_msm_accumulate", "bound": "int32-pipe", "achieved": None, "peak": IMAD_PEAK_TOPS, "unit": "TIMAD/s", "frac": None, "traffic": None,
                                  "note": "per-kernel accounting is reported by the N = 1 line; at N > 1 the step is latency-bound (collectives + fixed per-pass launch chains)"}})

    # ---- side measurement (BASELINE metric 'sumcheck fold GB/s'): bind one 2^26-entry table, HBM-bound
    if not args.no_fold and rank == 0 and LOG_OPS >= 20:
        nv = FOLD_LOG
        w, _ = ts.chacha20_fr_then_u64(bytes([7]) * 32, nv + 1)
        T = ctx.table_eq(w[:nv])
        r = np.ascontiguousarray(w[nv:nv + 1])
        ctx.set_tuning("kernel_timing", 1)
        for _ in range(3):
            T.clone().bind(r)
        cl = [T.clone() for _ in range(5)]
        ctx.timer_reset()
        for c in cl:
            c.bind(r)
        ctx.synchronize()
        b_ms, b_cnt = ctx.timer_read("bind")
        ctx.set_tuning("kernel_timing", 0)
        gbs = 48.0 * (1 << nv) / (b_ms / b_cnt * 1e-3) / 1e9
        line["roofline_fold"] = {"kernel": "k_bind", "bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                                 # DRAM bytes per launch from `ncu --set full` (profiles/r02_ncu_full_summary.md: 0.537 GB read + 0.229 GB written at 2^24 entries), scaled to 2^nv
                                 "traffic": (0.537 + 0.229) * 1e9 * (1 << nv) / (1 << 24), "algorithmic_bytes": 48.0 * (1 << nv),
                                 "launch_ms": b_ms / b_cnt, "workload": f"one 2^{nv}-entry Fr table (2 GiB), 48 B per output entry", "peak_source": hbm_src}
        del cl, T

    # ---- BASELINE configs C3 / C4 / C5 with SURVEY 8(d)'s inputs
    if not args.no_configs and LOG_OPS >= 20:
        cfg = {}
        try:
            cfg["C4"] = bench_c4(ts, dd, ctx, rank, world, timed, max_over_ranks, hbm_peak, hbm_src, KS)
            cfg["C5"] = bench_c5(ts, dd, ctx, rank, world, timed, max_over_ranks, KS)
            cfg["C3"] = bench_c3(ts, ctx, rank, world, timed, max_over_ranks, pinned, KS)
        except Exception as e:   # noqa: BLE001 - a side config must not take the headline down with it
            import traceback
            cfg["error"] = repr(e) + " | " + traceback.format_exc().splitlines()[-3].strip()
        line["configs"] = cfg

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle as O
        cores = O.ncpu()
        v, _, _, sample = cpu_prover(cores, 1, 0, budget_s=30.0)
        line["cpu_baseline"] = {"value": v, "unit": "ms", "cores": cores, "kind": "port", "sample": sample}
    elif rank == 0:
        line["cpu_baseline"] = None

    if rank == 0:
        print(json.dumps(line), file=JSON_OUT, flush=True)
    if world > 1:
        ctx.close()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ BASELINE configs 3, 4, 5
def bench_c4(ts, dd, ctx, rank, world, timed, max_over_ranks, hbm_peak, hbm_src, steps):
    """C4: SumCheck::prove over A = eq(w, .), B = one-hot 2^10 addresses x 2^16 cycles (2^26 entries each); w = 26 Fr::rand and the addresses from
    ChaCha20Rng::from_seed([4; 32]).  N > 1: the hypercube sliced by the high index bits, one exchange of the round values per round."""
    nv, logK = 26, 10
    logG = world.bit_length() - 1
    nloc = nv - logG
    rows = 1 << (nv - logK)
    w, addr = ts.chacha20_fr_then_u64(bytes([4]) * 32, nv, rows)
    addr = addr % np.uint64(1 << logK)
    rows_loc = rows >> logG
    A = ctx.table_eq(w[:nloc])
    if logG:                                                 # eq over the high variables at this rank's bits, as one scalar
        s = 1
        for k in range(logG):
            wk = ts.fe_to_int(w[nloc + k])
            s = s * (wk if (rank >> k) & 1 else 1 - wk) % R_MOD
        A = A.scalar_mul(ts.fe_from_int(s))
    B = ctx.table_one_hot_rows(np.ascontiguousarray(addr[rank * rows_loc:(rank + 1) * rows_loc]), logK, nloc)
    # claimed sum = sum_i A_i B_i over all ranks (one round evaluation: g(0) + g(1))
    sc = ctx.sumcheck([A.clone(), B.clone()]); ev = sc.round_eval(); sc.end()
    part = dd.fr_add(ev[0], ev[1])
    if world > 1:
        allp = ctx.comm_allgather(part.reshape(1, 4)).reshape(-1, 4)
        claimed = allp[0]
        for g in range(1, world):
            claimed = dd.fr_add(claimed, allp[g])
    else:
        claimed = part
    prove = (lambda tt: ts.SumCheck(nv, claimed).prove_product_sharded(ctx, tt, ts.Transcript())) if world > 1 else \
            (lambda tt: ts.SumCheck(nv, claimed).prove_product(ctx, tt, ts.Transcript()))
    proof = prove([A.clone(), B.clone()])
    ok, _ = ts.SumCheck(nv, claimed).verify(proof, ts.Transcript())
    assert ok, "C4 proof does not verify"
    import hashlib
    out = {"workload": f"SumCheck::prove, 2 tables of 2^{nv} entries (eq x one-hot 2^{logK} x 2^{nv - logK}), {nv} rounds, host transcript" + (f", hypercube sharded over {world} ranks" if world > 1 else ""),
           "n_gpus": world, "scaling": "strong", "proof_sha256": hashlib.sha256(proof.round_polynomials.tobytes() + proof.final_evaluation.tobytes()).hexdigest()[:16]}
    N = float(1 << nv)
    for name, flag in (("ms", 0), ("ms_deferred_claim_check", 1)):
        ctx.set_tuning("deferred_claim_check", flag)
        clones = [[A.clone(), B.clone()] for _ in range(steps + 2)]
        it = iter(clones)
        (ms,) = max_over_ranks(timed(lambda: prove(next(it)), steps, 2))
        out[name] = ms
        del clones, it
    ctx.set_tuning("deferred_claim_check", 0)
    ctx.set_tuning("sc_tail", 0)                                  # the same proof with one launch + stream synchronisation per round all the way down
    clones = [[A.clone(), B.clone()] for _ in range(steps + 2)]
    it = iter(clones)
    (out["ms_without_persistent_tail"],) = max_over_ranks(timed(lambda: prove(next(it)), steps, 2))
    ctx.set_tuning("sc_tail", 1)
    del clones, it
    if world > 1:
        out["exchange"] = "peer mailboxes over NVLink: round sums inside the round kernel, no NCCL call per round" if ctx.comm_peer_exchange else "ncclAllReduce per round"
        if ctx.comm_peer_exchange:                            # the same proof over the NCCL collectives, for comparison
            ctx.set_tuning("peer_exchange", 0)
            clones = [[A.clone(), B.clone()] for _ in range(steps + 2)]
            it = iter(clones)
            (out["ms_nccl_allreduce_per_round"],) = max_over_ranks(timed(lambda: prove(next(it)), steps, 2))
            ctx.set_tuning("peer_exchange", 1)
            del clones, it
        out["link_bytes_per_round_per_rank"] = 2 * 32 * (world - 1)    # two field elements (g(0), g(2)) to each peer
    gbs = 128.0 * 2 * N / (out["ms"] * 1e-3) / 1e9
    out["roofline"] = {"kernel": "k_round_eval<2> + k_bind_eval2_claim (whole protocol)", "bound": "hbm", "achieved": gbs, "peak": hbm_peak * world, "unit": "GB/s",
                       "frac": gbs / (hbm_peak * world), "traffic": None, "algorithmic_bytes": 256.0 * N, "peak_source": hbm_src + (f" x {world} GPUs" if world > 1 else ""),
                       "achieved_deferred": 128.0 * 2 * N / (out["ms_deferred_claim_check"] * 1e-3) / 1e9,
                       "note": "128 d N bytes (SURVEY 8d) over the whole prove call incl. the 26 host transcript round trips; default = the reference's deterministic round-0 check, "
                               "`deferred` = the opt-in claim-form round 0"}
    return out


def bench_c5(ts, dd, ctx, rank, world, timed, max_over_ranks, steps):
    """C5: G1 MSM of 2^24 points: bases = g1_powers of setup_params (tau = first Fr::rand of seed [42; 32]), scalars 2^24 x Fr::rand from seed [5; 32]
    (uniform, full width).  N > 1: sliced by points, partial results all-gathered and added."""
    logn = 24
    n = 1 << logn
    tau = ts.chacha20_fr_then_u64(bytes([42]) * 32, 1)[0][0]
    lo, hi = dd.slice_bounds(n, rank, world)
    t0 = time.perf_counter()
    srs = ctx.srs_generate_range(tau, lo, hi - lo) if world > 1 else ctx.srs_generate(tau, n)
    sc, _ = ts.chacha20_fr_then_u64(bytes([5]) * 32, n)
    poly = ctx.poly_upload(np.ascontiguousarray(sc[lo:hi])); del sc
    ctx.synchronize()
    setup_s = time.perf_counter() - t0
    coll = dd.Collective() if world > 1 else None

    def step():
        part = ts.KZGCommitment.commit(srs, poly)
        return dd.sharded_commit(part, coll) if world > 1 else part
    c = step()
    ctx.set_tuning("kernel_timing", 1)
    step(); ctx.timer_reset()
    e_0 = ctx.counter("msm_entries"); c_0 = ctx.counter("msm_calls")
    (ms,) = max_over_ranks(timed(step, steps, 1))
    ctx.synchronize()
    acc, cnt = ctx.timer_read("msm_accumulate")
    ph = {k: ctx.timer_read("msm_" + k)[0] / max(cnt, 1) for k in ("sort", "merge", "reduce")}
    ctx.set_tuning("kernel_timing", 0)
    calls = max(ctx.counter("msm_calls") - c_0, 1)
    entries = (ctx.counter("msm_entries") - e_0) / calls
    acc_ms = acc / max(cnt, 1)
    pts = n / (ms * 1e-3)
    model = (hi - lo) * 16 * 11 * 136 / (acc_ms * 1e-3) / 1e12 if acc_ms else None
    out = {"workload": f"KZGCommitment::commit = G1 MSM of 2^{logn} points, uniform full-width scalars" + (f", point-sharded over {world} ranks" if world > 1 else ""),
           "n_gpus": world, "scaling": "strong", "ms": ms, "points_per_s": pts, "setup_s": setup_s, "commitment": ts.g1_compress(c).hex()[:16],
           "accumulate_ms": acc_ms, "sort_ms": ph["sort"], "merge_ms": ph["merge"], "reduce_ms": ph["reduce"], "bucket_entries": entries,
           "roofline": {"kernel": "k_msm_accumulate", "bound": "int32-pipe", "achieved": model, "peak": IMAD_PEAK_TOPS, "unit": "TIMAD/s", "frac": model / IMAD_PEAK_TOPS if model else None,
                        "traffic": None, "executed": entries * 1288 / (acc_ms * 1e-3) / 1e12 if acc_ms else None,
                        "note": "per rank: local points x 23 936 model IMAD (SURVEY 8d) / accumulate launch time; `executed` = bucket entries x 1288 multiply-adds"}}
    del srs, poly
    return out


def bench_c3(ts, ctx, rank, world, timed, max_over_ranks, pinned, steps):
    """C3: Shout::prove, table entries[i] = i^2 (src/benchmarks.rs:167-169), T = 2^20; 2^22 lookups at ChaCha20Rng::from_seed([3; 32]) indices
    (next_u64 mod T).  N > 1: ONE proof sharded over the ranks (table and lookup vector sliced by position)."""
    logT, logL = 20, 22
    T, L = 1 << logT, 1 << logL
    t0 = time.perf_counter()
    pp, vp = ts.setup_params(ctx, logL - 2)
    setup_s = time.perf_counter() - t0
    shout = ts.Shout.new(pp)
    i = np.arange(T, dtype=np.uint64)
    idx = ts.chacha20_u64(bytes([3]) * 32, L) % np.uint64(T)
    keep = []
    if world > 1:
        elo, ehi = shout.shard_range(T, rank, world); llo, lhi = shout.shard_range(L, rank, world)
        t_, e_h = pinned(ts.fe_vec(i[elo:ehi] * i[elo:ehi])); keep.append(t_)
        t_, l_h = pinned(idx[llo:lhi]); keep.append(t_)
        step = lambda: shout.prove_sharded(e_h, T, l_h, L)      # noqa: E731
    else:
        t_, e_h = pinned(ts.fe_vec(i * i)); keep.append(t_)
        t_, l_h = pinned(idx); keep.append(t_)
        step = lambda: shout.prove_arrays(e_h, l_h)             # noqa: E731
    proof = step()
    assert shout.verify(proof, vp), "C3 proof does not verify"
    import hashlib
    ctx.set_tuning("kernel_timing", 1)
    step(); ctx.timer_reset()
    (ms,) = max_over_ranks(timed(step, steps, 2))
    ctx.synchronize()
    acc, cnt = ctx.timer_read("msm_accumulate")
    ctx.set_tuning("kernel_timing", 0)
    out = {"workload": f"Shout::prove, 2^{logT}-entry table (i^2), 2^{logL} lookups (ChaCha20 seed [3; 32]), setup_params({logL - 2}), host buffers in, proof out"
                       + (f", ONE proof sharded over {world} ranks" if world > 1 else ""),
           "n_gpus": world, "scaling": "strong", "ms": ms, "lookups_per_s": L / (ms * 1e-3), "setup_s": setup_s, "proof_sha256": hashlib.sha256(proof.to_bytes()).hexdigest()[:16],
           "h2d_bytes_per_step": int(e_h.nbytes + l_h.nbytes), "accumulate_ms_per_proof": acc / (steps + 2) if cnt else None,
           "roofline": {"kernel": "k_msm_accumulate", "bound": "int32-pipe", "achieved": None, "peak": IMAD_PEAK_TOPS, "unit": "TIMAD/s", "frac": None, "traffic": None,
                        "share_of_step": (acc / (steps + 2)) / ms if cnt and ms else None,
                        "note": "same kernel as the headline (rate reported there); here its share of the proof"}}
    del pp, vp, shout
    return out


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py - headline benchmark of the B200 prover backend.

Workload (BASELINE.json configs[1]): Twist::prove over a synthetic MemoryTrace of 2^20 read/write operations on
2^16 memory cells, setup_params(18) (prove rejects more than 4 * 2^log_size operations, src/twist.rs:108).
A "step" is one Twist::prove.  Default path: 4 G1 MSMs of 2^20 points over the evaluation-basis SRS (2 commitments over the
raw addresses / values, 2 opening quotients), 2 barycentric evaluation + quotient passes, the 20-round (all-zero) sum-check
transcript on the host.  `coefficient_path` reports the reference's own sequence (2 interpolations of 2^20 points, 4 full-width
MSMs, 2 Horner/quotient scans) - identical proof bytes, checked in the run.

    python bench.py --gpus N --steps K --warmup W            our arm (torchrun launches N ranks for N > 1)
    python bench.py --impl reference ...                     CPU arm: the oracle port of the reference prover

One JSON line on stdout (rank 0).  `value` = ms per proof with the padded vectors resident in HBM;
`e2e.value` = ms per proof through the public API from pinned host memory (H2D of the trace + D2H of the proof
inside the timed region).  Multi-GPU (N > 1): each rank proves an independent trace of the same shape (traces are
independent objects; no data-path collective), so `value` stays "ms per proof" and throughput scales with N.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "multilinear-map-cryptography_b200"

LOG_OPS = 20
LOG_CELLS = 16
LOG_SIZE = 18            # setup_params(18): max_operations = 2^20
CPU_SAMPLE_LOG = 16      # CPU arms time a 2^16-op sample and scale linearly (optimistic for the CPU: the work is n log^2 n)
FOLD_LOG = 26            # side measurement: sum-check fold of one 2^26-entry table (BASELINE metric "sumcheck fold GB/s")
JSON_OUT = sys.stdout
IMAD_PEAK_TOPS = 18.5    # measured on this pool with tools/ubench.cu (plain IMAD.WIDE issue rate, profiles/r01_ubench.json)


def synthetic_trace(log_ops: int, log_cells: int, seed: int):
    """2^log_ops operations: address uniform in [0, 2^log_cells), write with probability 1/2 of a fresh 63-bit value,
    reads return the last value written to the address (0 if none) - the semantics of MemoryTrace::read/write."""
    n = 1 << log_ops
    rng = np.random.default_rng(seed)
    addr = rng.integers(0, 1 << log_cells, size=n, dtype=np.uint64)
    is_write = rng.integers(0, 2, size=n, dtype=np.uint8)
    fresh = rng.integers(0, 1 << 63, size=n, dtype=np.uint64)
    # last-write-wins simulation, vectorised: for every op, index of the latest write to the same address at or before it
    order = np.lexsort((np.arange(n), addr))
    a_s, w_s, f_s = addr[order], is_write[order], fresh[order]
    idx = np.where(w_s == 1, np.arange(n), -1)
    seg_start = np.r_[True, a_s[1:] != a_s[:-1]]
    seg_id = np.cumsum(seg_start) - 1
    first_of_seg = np.flatnonzero(seg_start)[seg_id]
    last_write = np.maximum.accumulate(np.where(idx >= 0, idx, -1))
    valid = last_write >= first_of_seg
    vals_sorted = np.where(valid, f_s[np.clip(last_write, 0, n - 1)], 0).astype(np.uint64)
    values = np.empty(n, dtype=np.uint64)
    values[order] = vals_sorted
    return addr, values, is_write


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
            # honour CUDA_VISIBLE_DEVICES remapping through the PCI bus id of the torch device
            import torch
            bus = torch.cuda.get_device_properties(self.index).pci_bus_id if hasattr(torch.cuda.get_device_properties(self.index), "pci_bus_id") else None
            self.h = None
            if bus is not None:
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


def cpu_prover_sample(threads: int, steps: int, warmup: int):
    """The oracle's arkworks-class CPU prover (oracle/oracle.cpp, fast tier: NTT interpolation, Pippenger MSM with
    ark-ec's window rule, threaded) on a 2^CPU_SAMPLE_LOG-op trace of the same distribution; ms scaled to 2^20 ops."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    O.build()
    n = 1 << CPU_SAMPLE_LOG
    addr, vals_u64, isw = synthetic_trace(CPU_SAMPLE_LOG, LOG_CELLS, seed=2)
    vals = O.fr_from_u64(vals_u64)
    powers = O.setup_g1_powers(n + 1, fast=True, threads=threads)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.twist_prove(powers, n, addr, vals, isw, fast=True, threads=threads)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    sample_ms = 1e3 * float(np.mean(times))
    scale = (1 << LOG_OPS) / n
    return sample_ms * scale, sample_ms, f"oracle fast tier, Twist::prove on 2^{CPU_SAMPLE_LOG} ops ({sample_ms:.0f} ms measured), scaled x{int(scale)} linearly to 2^{LOG_OPS} ops"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    cores = O.ncpu()
    steps = max(1, min(args.steps, 3)); warmup = min(args.warmup, 1)
    value, sample_ms, sample = cpu_prover_sample(cores, steps, warmup)
    line = {
        "impl": "reference", "metric": "twist_prove_ms_at_2^20_ops", "value": value, "unit": "ms", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": value, "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64x4 (BN254 Fr/Fq, 4x64-bit Montgomery)", "data": "synthetic",
        "config": {"workload": f"Twist::prove, 2^{LOG_CELLS} cells, 2^{LOG_OPS} ops, setup_params({LOG_SIZE}) - CPU sample scaled", "note":
                   "the Rust reference cannot be built here (no toolchain) and its O(n^3) interpolation cannot reach 2^20; this arm is the oracle's quasi-linear CPU prover"},
        "cpu_baseline": {"value": value, "unit": "ms", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=JSON_OUT, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line (NCCL_DEBUG=VERSION prints to stdout otherwise)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ts = importlib.import_module(PKG)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = ts.Context(local, stream.cuda_stream)
    W = max(args.warmup, 3); K = max(args.steps, 1)
    for kv in filter(None, os.environ.get("TSGPU_TUNING", "").split(",")):      # A/B switches for experiments, e.g. TSGPU_TUNING=msm_acc_waves=0,msm_two_level=0
        key, val = kv.split("=")
        ctx.set_tuning(key, int(val))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local); sampler.start()
    t0 = time.perf_counter()
    pp, vp = ts.setup_params(ctx, LOG_SIZE)                     # SRS of 2^20 + 1 points generated on the device
    setup_s = time.perf_counter() - t0
    twist = ts.Twist.new(pp)
    n = 1 << LOG_OPS
    addr, vals_u64, isw = synthetic_trace(LOG_OPS, LOG_CELLS, seed=2 + rank)
    vals = ts.fe_vec(vals_u64)
    # pinned host buffers: what the public API call reads from
    addr_pin = torch.empty(n, dtype=torch.int64, pin_memory=True); addr_pin.numpy().view(np.uint64)[:] = addr
    vals_pin = torch.empty((n, 4), dtype=torch.int64, pin_memory=True); vals_pin.numpy().view(np.uint64)[:] = vals
    addr_h = addr_pin.numpy().view(np.uint64); vals_h = vals_pin.numpy().view(np.uint64)

    # ---- correctness guard: the proof must verify (transcript + sum-check + openings) before anything is timed
    proof = twist.prove_arrays(addr_h, vals_h, isw)
    assert twist.verify(proof, vp), "benchmark proof does not verify"
    proof_len = len(proof.to_bytes())

    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    base_a = ctx.poly_from_u64(addr_h); base_v = ctx.poly_upload_padded(vals_h, n)

    def time_e2e():
        """public API, pinned host buffers in, proof read back: ms per proof, launches per proof"""
        for _ in range(W):
            twist.prove_arrays(addr_h, vals_h, isw)
        barrier()
        l0 = ctx.launch_count
        e0.record(stream)
        for _ in range(K):
            p = twist.prove_arrays(addr_h, vals_h, isw)
            _ = p.final_evaluation                               # D2H'd proof contents are read
        e1.record(stream)
        barrier()
        return e0.elapsed_time(e1) / K, (ctx.launch_count - l0) // K

    def time_device():
        """padded vectors already in HBM: per-proof ms, launches, MSM work counters and kernel timers of the timed region"""
        # every input copy is made BEFORE the warm-up, so that the stream-ordered pool has its final size when the timed region starts
        # (growing the pool inside a step stalls it for ~100 ms: tools/diag_stall.py)
        warm = [(base_a.clone(), base_v.clone()) for _ in range(W)]
        clones = [(base_a.clone(), base_v.clone()) for _ in range(K)]
        ctx.set_tuning("kernel_timing", 1)
        for a, v in warm:
            twist.prove_device(a, v)
        del warm
        ctx.timer_reset()
        barrier()
        c0 = {k: ctx.counter(k) for k in ("launches", "msm_calls", "msm_points", "msm_entries")}
        e0.record(stream)
        for a, v in clones:
            twist.prove_device(a, v)
        e1.record(stream)
        barrier()
        r = {"ms": e0.elapsed_time(e1) / K}
        r.update({k: ctx.counter(k) - c0[k] for k in c0})
        for name in ("msm_accumulate", "msm_total", "msm_sort", "msm_merge", "msm_reduce", "interpolate", "open_bary", "open_scan"):
            r[name + "_ms"], r[name + "_cnt"] = ctx.timer_read(name)
        ctx.set_tuning("kernel_timing", 0)
        return r

    # ---- coefficient path (tuning eval_basis = 0): interpolate -> commit -> open on coefficients, the reference's own sequence of steps
    ctx.set_tuning("eval_basis", 0)
    proof_c = twist.prove_arrays(addr_h, vals_h, isw)
    coef_e2e_ms, coef_e2e_launches = time_e2e()
    coef = time_device()
    # ---- default path (evaluation-basis SRS prepared by setup_params): commitments and openings straight from the values
    ctx.set_tuning("eval_basis", 1)
    assert twist.prove_arrays(addr_h, vals_h, isw).to_bytes() == proof_c.to_bytes(), "the two paths must give identical proof bytes"
    sampler.mark()
    e2e_ms, e2e_launches = time_e2e()
    dflt = time_device()
    clocks = sampler.stop()
    dev_ms = dflt["ms"]

    if world > 1:
        t = torch.tensor([dev_ms, e2e_ms, coef["ms"], coef_e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms, coef["ms"], coef_e2e_ms = (float(x) for x in t)

    # ---- N > 1, strong-scaling companion: ONE proof of the same trace shape sharded over the ranks (evaluation-basis slices per rank, three small
    # all-gathers over the library's NCCL communicator; byte-identical to the one-GPU proof, tests/test_gpu_distributed.py).  Reported beside the
    # weak-scaling headline (one independent proof per GPU); host buffers in, proof out, CUDA events, max over ranks.
    sharded = None
    if world > 1:
        try:
            ctx.comm_init_torch()
            addr0, vals0_u64, _ = synthetic_trace(LOG_OPS, LOG_CELLS, seed=2)                    # the same trace on every rank; each passes its slice
            lo, hi = twist.shard_range(n, rank, world)
            a_pin = torch.empty(hi - lo, dtype=torch.int64, pin_memory=True); a_pin.numpy().view(np.uint64)[:] = addr0[lo:hi]
            v_pin = torch.empty((hi - lo, 4), dtype=torch.int64, pin_memory=True); v_pin.numpy().view(np.uint64)[:] = ts.fe_vec(vals0_u64[lo:hi])
            a_s = a_pin.numpy().view(np.uint64); v_s = v_pin.numpy().view(np.uint64)
            ps = twist.prove_sharded(a_s, v_s, n)
            assert twist.verify(ps, vp), "sharded proof does not verify"
            for _ in range(W):
                twist.prove_sharded(a_s, v_s, n)
            barrier()
            e0.record(stream)
            for _ in range(K):
                twist.prove_sharded(a_s, v_s, n)
            e1.record(stream)
            barrier()
            t = torch.tensor([e0.elapsed_time(e1) / K], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sharded = {"value": float(t[0]), "unit": "ms", "scaling": "strong", "ops_per_s": n / (float(t[0]) * 1e-3),
                       "what": "ONE Twist::prove of 2^20 ops sharded over the ranks (tsgpu_twist_prove_sharded), host buffers"}
        except Exception as e:   # noqa: BLE001
            sharded = {"error": repr(e)}

    # ---- roofline of the dominant kernel of the step: MSM bucket accumulation (integer-pipe bound).
    # Two accountings, both reported:
    #  * `achieved` (the contract's): SURVEY 8(d)'s per-unit figure x the units of a launch / the launch duration.  The figure is
    #    n x 16 windows x 11 Fq products x 136 IMAD = 23 936 IMAD per point of a full-width MSM; the launches that process full-width
    #    scalars are the open passes (two quotient vectors of 2^20 points each per launch), so achieved = 2 n x 23 936 / t(open-pass launch);
    #    t(open pass) = (accumulate time of the step - commit-pass share), the commit pass being timed by its own entry count at the same rate.
    #  * `executed`: what the kernel really does - one mixed XYZZ + affine addition (8M + 2S: 8 Fq products x 136 IMAD + 200 for the fused Y coordinate = 1288) per bucket entry
    #    (non-zero signed digit, counted by the library): fewer than the model because the window tables need 13 additions per point, not 16.
    imad_per_add = 8 * (2 * 8 * 8 + 8) + (2 * 64 + 72)     # 8 Montgomery products + the fused a b - c d of the Y coordinate (one reduction)
    survey_imad_per_point = 16 * 11 * 136

    def acc_roofline(r, full_width_points_per_step):
        launches = max(r["msm_accumulate_cnt"], 1)
        launch_ms = r["msm_accumulate_ms"] / launches
        entries = r["msm_entries"] / max(r["msm_calls"], 1)
        executed = entries * imad_per_add / (launch_ms * 1e-3) / 1e12 if launch_ms > 0 else 0.0
        # time spent on full-width scalars: total accumulate time x their share of the bucket entries (13 per point)
        fw_entries = 13.0 * full_width_points_per_step * K
        fw_ms = r["msm_accumulate_ms"] * min(1.0, fw_entries / max(r["msm_entries"], 1))
        ach = full_width_points_per_step * K * survey_imad_per_point / (fw_ms * 1e-3) / 1e12 if fw_ms > 0 else 0.0
        return {"kernel": "k_msm_accumulate", "bound": "int32-pipe", "achieved": ach, "peak": IMAD_PEAK_TOPS, "unit": "TIMAD/s",
                "frac": ach / IMAD_PEAK_TOPS, "traffic": None, "launch_ms": launch_ms, "launches": r["msm_accumulate_cnt"],
                "executed": executed, "executed_frac": executed / IMAD_PEAK_TOPS, "entries_per_launch": entries,
                "share_of_step": r["msm_accumulate_ms"] / K / r["ms"] if r["ms"] > 0 else None}

    roofline = acc_roofline(dflt, 2 * n)           # default path: the two opening quotients are the full-width scalars
    roofline["peak_source"] = "tools/ubench.cu on this pool: 18.5 T IMAD.WIDE/s without carry predicate (the carry-chained form the multiplier needs issues at half that rate)"
    roofline["algorithmic_unit"] = "SURVEY 8(d): 16 windows x 11 Fq products x 136 IMAD = 23 936 IMAD per full-width point; `executed` = bucket entries x 1288 IMAD (8M + 2S with the Y coordinate's two products sharing one reduction)"
    # DRAM bytes per launch from `ncu --set full` of this same command (profiles/r01_ncu_accumulate_in_bench.md): commit pass 0.587 + 0.068 GB,
    # open pass 3.622 + 0.146 GB; mean over the two launches of a proof, like `launch_ms`.  Algorithmic bytes: entries x (64 B point + 4 B entry).
    roofline["traffic"] = 0.5 * ((0.586715 + 0.067997) + (3.621562 + 0.145975)) * 1e9
    roofline["algorithmic_bytes"] = roofline["entries_per_launch"] * 68.0
    roofline["ncu"] = "FMA-heavy pipe 88-89% busy, thread efficiency 31.8/32, DRAM 0.87 TB/s (profiles/r01_ncu_accumulate_in_bench.md)"

    # ---- side measurement (BASELINE metric 'sumcheck fold GB/s'): bind one 2^26-entry table, HBM-bound
    fold = None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0)); hbm_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback 6650 GB/s"
    if not args.no_fold and rank == 0:
        nv = FOLD_LOG
        w = np.ascontiguousarray(vals_h[:nv])
        r = np.ascontiguousarray(vals_h[nv:nv + 1])
        T = ctx.table_eq(w)
        ctx.set_tuning("kernel_timing", 1)
        for _ in range(3):
            T.clone().bind(r)
        cl = [T.clone() for _ in range(5)]
        ctx.timer_reset()
        for c in cl:
            c.bind(r)
        b_ms, b_cnt = ctx.timer_read("bind")
        ctx.set_tuning("kernel_timing", 0)
        gbs = 48.0 * (1 << nv) / (b_ms / b_cnt * 1e-3) / 1e9
        fold = {"kernel": "k_bind", "bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                # DRAM bytes per launch from `ncu --set full` (profiles/r01_ncu_full_summary.md: 0.5369 GB read + 0.2257 GB written at 2^24 entries), scaled to 2^nv
                "traffic": (0.536879 + 0.225681) * 1e9 * (1 << nv) / (1 << 24), "algorithmic_bytes": 48.0 * (1 << nv),
                "launch_ms": b_ms / b_cnt, "workload": f"one 2^{nv}-entry Fr table (2 GiB), 48 B per output entry", "peak_source": hbm_src}
        del cl, T

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle as O
        cores = O.ncpu()
        v, sample_ms, sample = cpu_prover_sample(cores, 1, 0)
        cpu = {"value": v, "unit": "ms", "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        def breakdown(r):
            return {"msm_4x": r["msm_total_ms"] / K, "msm_accumulate_4x": r["msm_accumulate_ms"] / K, "msm_sort_4x": r["msm_sort_ms"] / K,
                    "msm_chunk_merge_4x": r["msm_merge_ms"] / K, "msm_window_reduce_4x": r["msm_reduce_ms"] / K, "interpolate_2x": r["interpolate_ms"] / K,
                    "open_barycentric_2x": r["open_bary_ms"] / K, "open_scan_2x": r["open_scan_ms"] / K, "bucket_entries": r["msm_entries"] // K,
                    "gpu_launches": r["launches"] // K}
        line = {
            "metric": "twist_prove_ms_at_2^20_ops", "value": dev_ms, "unit": "ms", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms, "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32x8 (BN254 Fr/Fq, 256-bit Montgomery on the integer pipe)", "data": "synthetic",
            "config": {"workload": f"Twist::prove, 2^{LOG_CELLS} cells, 2^{LOG_OPS} random read/write ops, setup_params({LOG_SIZE})",
                       "path": "default: evaluation-basis SRS (commit / open from the values, no interpolation); byte-identical to the coefficient path, which is timed in `coefficient_path`",
                       "per_gpu": "one independent proof per GPU" if world > 1 else "single GPU",
                       "l2": "per-step working set (2 x 32 MiB vectors, 2 x 64 MiB SRS, ~200 MiB MSM scratch per MSM) exceeds the 126 MB L2; no explicit flush",
                       "proof_bytes": proof_len, "setup_s": setup_s},
            "clocks": clocks,
            "e2e": {"value": e2e_ms, "unit": "ms", "h2d_bytes_per_step": int(addr_h.nbytes + vals_h.nbytes), "d2h_bytes_per_step": int(4 * 16 * 96 + 2 * 32),
                    "gpu_launches": int(e2e_launches)},
            "gpu_launches": int(dflt["launches"]),
            "roofline": roofline,
            "roofline_fold": fold,
            "cpu_baseline": cpu,
            "breakdown_ms_per_step": breakdown(dflt),
            "coefficient_path": {"value": coef["ms"], "unit": "ms", "e2e": coef_e2e_ms, "breakdown_ms_per_step": breakdown(coef), "roofline": acc_roofline(coef, 4 * n),
                                 "note": "tsgpu_set_tuning(eval_basis, 0): 2 interpolations + 4 full-width MSMs + 2 Horner/quotient scans"},
            "msm_points_per_s": dflt["msm_points"] / (dflt["msm_total_ms"] * 1e-3) if dflt["msm_total_ms"] > 0 else None,
            "msm_full_width_points_per_s": coef["msm_points"] / (coef["msm_total_ms"] * 1e-3) if coef["msm_total_ms"] > 0 else None,
            "ops_per_s_all_gpus": world * n / (dev_ms * 1e-3),
            "one_proof_sharded": sharded,
        }
        print(json.dumps(line), file=JSON_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

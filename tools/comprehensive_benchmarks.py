#!/usr/bin/env python
"""The reference's benchmark CLI (examples/comprehensive_benchmarks.rs + src/benchmarks.rs `ProtocolBenchmarks`) re-pointed at the GPU backend:
same modes, same synthetic generators (src/benchmarks.rs:88-99 for Twist, :167-177 for Shout), same operation scaling per size
(:54-66, :137-149), same table columns.  Timing is wall clock around setup / prove / verify as in the reference (`Instant`), on cuda:0.

    python tools/comprehensive_benchmarks.py                 log sizes 4-8, 256 operations in the comparison (the reference's default)
    python tools/comprehensive_benchmarks.py quick|full|dev  4-6 / 4-10 / 4-5
    python tools/comprehensive_benchmarks.py custom --min-log-size A --max-log-size B --operations N
    python tools/comprehensive_benchmarks.py twist-only|shout-only [--min-log-size A --max-log-size B]
    python tools/comprehensive_benchmarks.py b200            log sizes 10-18 with FULL traces (4 * 2^log_size operations = max_operations),
                                                             the sizes the CPU reference cannot reach (its interpolation is O(n^3))
Every proof is verified (pairing check on the CPU), as the reference harness asserts."""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ts = importlib.import_module("multilinear-map-cryptography_b200")


class BenchmarkResults:                                   # src/benchmarks.rs:9-38
    def __init__(self, setup_time, prove_time, verify_time, proof_size, num_operations, memory_usage):
        self.setup_time, self.prove_time, self.verify_time = setup_time, prove_time, verify_time
        self.proof_size, self.num_operations, self.memory_usage = proof_size, num_operations, memory_usage

    def prove_ops_per_second(self):
        return self.num_operations / self.prove_time if self.prove_time > 0 else 0.0

    def verify_ops_per_second(self):
        return self.num_operations / self.verify_time if self.verify_time > 0 else 0.0

    def total_time(self):
        return self.setup_time + self.prove_time + self.verify_time


def scaled_operations(size: int) -> int:                  # src/benchmarks.rs:54-66 / :137-149
    return size // 2 if size <= 64 else (size // 4 if size <= 512 else size // 8)


class ProtocolBenchmarks:
    """src/benchmarks.rs:42-365; one CUDA context shared by all runs, the params (SRS on the device) rebuilt per size like the reference"""

    def __init__(self, device: int = 0):
        self.ctx = ts.Context(device)

    def benchmark_twist_single(self, log_size: int, num_operations: int) -> BenchmarkResults:      # :76-125
        t0 = time.perf_counter()
        pp, vp = ts.setup_params(self.ctx, log_size)
        twist = ts.Twist.new(pp)
        setup = time.perf_counter() - t0
        memory_size = 1 << log_size
        i = np.arange(num_operations, dtype=np.uint64)
        is_write = (i % 3 == 0)
        addr = np.where(is_write, i % memory_size, (i // 2) % memory_size).astype(np.uint64)
        # reads return the simulated memory content (MemoryTrace::read): last value written to the address, 0 before any write
        values = np.zeros(num_operations, dtype=np.uint64)
        mem = {}
        for k in range(num_operations):
            a = int(addr[k])
            if is_write[k]:
                mem[a] = (k * 42) & 0xFFFFFFFFFFFFFFFF
            values[k] = mem.get(a, 0)
        vals = ts.fe_vec(values)
        t0 = time.perf_counter()
        proof = twist.prove_arrays(addr, vals, is_write.astype(np.uint8))
        prove = time.perf_counter() - t0
        t0 = time.perf_counter()
        ok = twist.verify(proof, vp)
        verify = time.perf_counter() - t0
        assert ok, "Proof verification failed"
        res = BenchmarkResults(setup, prove, verify, len(proof.to_bytes()), num_operations, 32 * (memory_size + 2 * num_operations))
        pp.free()
        return res

    def benchmark_shout_single(self, log_size: int, num_lookups: int) -> BenchmarkResults:          # :158-203
        t0 = time.perf_counter()
        pp, vp = ts.setup_params(self.ctx, log_size)
        shout = ts.Shout.new(pp)
        setup = time.perf_counter() - t0
        table_size = 1 << log_size
        i = np.arange(table_size, dtype=np.uint64)
        entries = ts.fe_vec(i * i)
        idx = (np.arange(num_lookups, dtype=np.uint64) % table_size).astype(np.uint64)
        t0 = time.perf_counter()
        proof = shout.prove_arrays(entries, idx)
        prove = time.perf_counter() - t0
        t0 = time.perf_counter()
        ok = shout.verify(proof, vp)
        verify = time.perf_counter() - t0
        assert ok, "Proof verification failed"
        res = BenchmarkResults(setup, prove, verify, len(proof.to_bytes()), num_lookups, 32 * (table_size + 2 * num_lookups))
        pp.free()
        return res

    def benchmark_twist_scaling_range(self, lo: int, hi: int, full_traces: bool = False):
        out = []
        for log_size in range(lo, hi + 1):
            size = 1 << log_size
            n = 4 * size if full_traces else scaled_operations(size)
            print(f"  Testing Twist with memory size: {size} (2^{log_size}), operations: {n}")
            out.append((size, self.benchmark_twist_single(log_size, n)))
        return out

    def benchmark_shout_scaling_range(self, lo: int, hi: int, full_traces: bool = False):
        out = []
        for log_size in range(lo, hi + 1):
            size = 1 << log_size
            n = 4 * size if full_traces else scaled_operations(size)
            print(f"  Testing Shout with table size: {size} (2^{log_size}), lookups: {n}")
            out.append((size, self.benchmark_shout_single(log_size, n)))
        return out

    def comparative_benchmark(self, log_size: int, num_operations: int):                            # :206-211
        return self.benchmark_twist_single(log_size, num_operations), self.benchmark_shout_single(log_size, num_operations)

    @staticmethod
    def print_scaling_results(protocol: str, results):                                              # :289-305
        print("Size\t| Setup(ms)\t| Prove(ms)\t| Verify(ms)\t| Proof(KB)\t| Ops/sec")
        print("--------|---------------|---------------|---------------|---------------|--------")
        for size, r in results:
            print(f"{size}\t| {r.setup_time * 1e3:.2f}\t\t| {r.prove_time * 1e3:.2f}\t\t| {r.verify_time * 1e3:.2f}\t\t| {r.proof_size / 1024:.2f}\t\t| {r.prove_ops_per_second():.0f}")

    @staticmethod
    def print_comparative_results(t, s):
        print("Protocol | Prove(ms) | Verify(ms) | Proof(KB) | Ops/sec | Memory(KB)")
        print("---------|-----------|------------|-----------|---------|----------")
        for name, r in (("Twist", t), ("Shout", s)):
            print(f"{name}    | {r.prove_time * 1e3:.2f}      | {r.verify_time * 1e3:.2f}       | {r.proof_size / 1024:.2f}      | {r.prove_ops_per_second():.0f}     | {r.memory_usage / 1024:.1f}")

    def run_comprehensive_benchmark_with_params(self, lo: int, hi: int, num_ops: int, full_traces: bool = False):   # :219-238
        print("Twist and Shout Protocol Benchmark Suite (B200 backend)")
        print("========================================================\n")
        print("Twist Protocol Scaling Analysis:")
        tw = self.benchmark_twist_scaling_range(lo, hi, full_traces)
        self.print_scaling_results("Twist", tw)
        print("\nShout Protocol Scaling Analysis:")
        sh = self.benchmark_shout_scaling_range(lo, hi, full_traces)
        self.print_scaling_results("Shout", sh)
        mid = (lo + hi) // 2
        print(f"\nComparative Analysis (Memory/Table Size: {1 << mid}):")
        t, s = self.comparative_benchmark(mid, min(num_ops, 4 << mid))
        self.print_comparative_results(t, s)
        return tw, sh


def parse_flags(argv, lo, hi, ops):
    i = 0
    while i < len(argv):
        if argv[i] in ("--min-log-size", "--max-log-size", "--operations") and i + 1 < len(argv):
            v = int(argv[i + 1])
            if argv[i] == "--min-log-size":
                lo = v
            elif argv[i] == "--max-log-size":
                hi = v
            else:
                ops = v
            i += 2
        else:
            raise SystemExit(f"Unknown argument: {argv[i]}")
    if lo > hi:
        raise SystemExit("min-log-size must not exceed max-log-size")
    return lo, hi, ops


def main():
    argv = sys.argv[1:]
    mode = argv[0] if argv else "default"
    if mode in ("help", "--help", "-h"):
        print(__doc__)
        return
    pb = ProtocolBenchmarks()
    if mode == "default":
        pb.run_comprehensive_benchmark_with_params(4, 8, 256)
    elif mode == "quick":
        pb.run_comprehensive_benchmark_with_params(4, 6, 64)
    elif mode == "full":
        pb.run_comprehensive_benchmark_with_params(4, 10, 256)
    elif mode == "dev":
        pb.run_comprehensive_benchmark_with_params(4, 5, 32)
    elif mode == "custom":
        lo, hi, ops = parse_flags(argv[1:], 4, 8, 256)
        pb.run_comprehensive_benchmark_with_params(lo, hi, ops)
    elif mode == "twist-only":
        lo, hi, _ = parse_flags(argv[1:], 4, 8, 256)
        pb.print_scaling_results("Twist", pb.benchmark_twist_scaling_range(lo, hi))
    elif mode == "shout-only":
        lo, hi, _ = parse_flags(argv[1:], 4, 8, 256)
        pb.print_scaling_results("Shout", pb.benchmark_shout_scaling_range(lo, hi))
    elif mode == "b200":
        lo, hi, ops = parse_flags(argv[1:], 10, 18, 1 << 16)
        tw, sh = pb.run_comprehensive_benchmark_with_params(lo, hi, ops, full_traces=True)
        print(json.dumps({"mode": "b200", "twist": [{"log_size": s.bit_length() - 1, "ops": r.num_operations, "setup_ms": r.setup_time * 1e3, "prove_ms": r.prove_time * 1e3,
                                                      "verify_ms": r.verify_time * 1e3, "proof_bytes": r.proof_size} for s, r in tw],
                          "shout": [{"log_size": s.bit_length() - 1, "lookups": r.num_operations, "setup_ms": r.setup_time * 1e3, "prove_ms": r.prove_time * 1e3,
                                     "verify_ms": r.verify_time * 1e3, "proof_bytes": r.proof_size} for s, r in sh]}))
    else:
        print(f"Unknown mode: {mode}\nUse 'help' for usage information.")
        raise SystemExit(1)
    print(f"\nBenchmarks completed ({pb.ctx.launch_count} kernel launches on the device).")


if __name__ == "__main__":
    main()

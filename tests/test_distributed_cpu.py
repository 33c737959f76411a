"""CPU, world_size 2 over gloo: the host-side logic of the sharded provers (distributed.py) - hypercube slicing,
limb all-reduce + carry/mod-r, transcript lock-step, tail hand-over, point-sharded commitment combine - with the
oracle as the per-rank round engine (the product engine needs a GPU; the GPU twin is tests/test_gpu_distributed.py)."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "multilinear-map-cryptography_b200"


class OracleRoundEngine:
    """per-rank rounds computed with the CPU oracle's conversions and Python integers (test infrastructure)"""

    def __init__(self):
        import oracle as O
        self.O = O

    def fresh(self):
        return OracleRoundEngine()

    def begin(self, tables):
        self.t = [self.O.fr_to_ints(np.ascontiguousarray(t, dtype=np.uint64).reshape(-1, 4)) for t in tables]

    @property
    def vars_left(self):
        return len(self.t[0]).bit_length() - 1

    def round_eval(self):
        R = self.O.R_MOD
        half = len(self.t[0]) // 2
        ev = []
        for x in range(4):
            s = 0
            for i in range(half):
                p = 1
                for t in self.t:
                    p = p * (t[2 * i] + x * (t[2 * i + 1] - t[2 * i])) % R
                s = (s + p) % R
            ev.append(s)
        return self.O.fr_from_ints(ev)

    def bind(self, r):
        R = self.O.R_MOD
        rr = self.O.fr_to_ints(r)[0]
        self.t = [[(t[2 * i] + rr * (t[2 * i + 1] - t[2 * i])) % R for i in range(len(t) // 2)] for t in self.t]

    def bind_eval(self, r):
        self.bind(r)
        return self.round_eval()

    def final(self):
        return self.O.fr_from_ints([t[0] for t in self.t])


def _worker(rank, world, port, nv, d, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import importlib
    ts = importlib.import_module(PKG)
    dd = importlib.import_module(PKG + ".distributed")
    import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 1 << nv
        tables = [O.chacha_fr_rand(bytes([40 + t]) * 32, n) for t in range(d)]
        ints = [O.fr_to_ints(t) for t in tables]
        claimed = 0
        for i in range(n):
            p = 1
            for t in range(d):
                p = p * ints[t][i] % O.R_MOD
            claimed = (claimed + p) % O.R_MOD
        claimed_m = O.fr_from_ints([claimed])[0]
        coll = dd.Collective()
        lo, hi = dd.slice_bounds(n, rank, world)
        local = [t[lo:hi] for t in tables]
        proof, chals, finals = dd.ShardedSumCheck(nv, claimed_m, coll).prove_product(OracleRoundEngine(), local, ts.Transcript())
        ref = O.sumcheck_prove_product(tables, claimed_m, mode="closure" if nv <= 5 else "tables")
        ok = (proof.round_polynomials == ref["round_polynomials"]).all() and (proof.final_evaluation == ref["final_evaluation"]).all() \
            and (chals == ref["challenges"]).all()
        try:      # a wrong claim must fail in round 0 on every rank (sumcheck.rs:77-84)
            dd.ShardedSumCheck(nv, O.fr_from_ints([7])[0], coll).prove_product(OracleRoundEngine(), local, ts.Transcript())
            ok = False
        except ts.TwistAndShoutError as e:
            ok = ok and e.variant == "SumCheck" and "Round 0" in str(e)
        big = O.fr_from_ints([O.R_MOD - 1, 5, 0, (1 << 200) + 3])          # limb all-reduce of extreme values
        tot = coll.all_reduce_fr(big)
        ok = ok and O.limbs_to_ints(tot) == [(world * x) % O.R_MOD for x in O.limbs_to_ints(big)]
        m = 64                                                             # point-sharded commitment
        pw = O.setup_g1_powers(m, fast=True)
        poly = O.chacha_fr_rand(bytes([9]) * 32, m)
        a, b = dd.slice_bounds(m, rank, world)
        part = O.msm_pippenger(O.g1_batch_to_affine(pw[a:b]), poly[a:b])
        total = dd.sharded_commit(part, coll)
        ok = ok and O.g1_compress(total) == O.g1_compress(O.kzg_commit(pw, poly))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nv,d", [(4, 2), (6, 3), (1, 1)])
def test_sharded_sumcheck_and_commit_world2_gloo(nv, d):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + nv * 7 + d
    procs = [ctx.Process(target=_worker, args=(r, 2, port, nv, d, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    results = sorted(q.get(timeout=5) for _ in range(2))
    assert results == [(0, True), (1, True)]
    assert all(p.exitcode == 0 for p in procs)


def test_limb_sum_conversion_single_process(tsgpu, oracle):
    import importlib
    dd = importlib.import_module(PKG + ".distributed")
    xs = [oracle.R_MOD - 1, 0, 1, (1 << 253) + 12345]
    a = oracle.ints_to_limbs(xs)
    s = dd.to_limb_sums(a) * 8            # as if 8 ranks had contributed the same value
    assert oracle.limbs_to_ints(dd.from_limb_sums(s)) == [8 * x % oracle.R_MOD for x in xs]
    ev = oracle.fr_from_ints([54, 54 + 51 + 3, 54 + 102 + 12, 54 + 153 + 27])      # 54 + 51x + 3x^2 at 0..3 (Appendix C.3)
    assert oracle.fr_to_ints(dd.round_coeffs(ev)) == [54, 51, 3, 0]


def test_shard_ranges_partition_the_padded_vectors():
    """Twist.shard_range / Shout.shard_range (what tsgpu_twist_prove_sharded / tsgpu_shout_prove_sharded expect from rank r): contiguous,
    disjoint, covering exactly the real operations, each inside the rank's slice [r m / G, (r + 1) m / G) of the padded length m"""
    import importlib
    ts = importlib.import_module("multilinear-map-cryptography_b200")
    for total in (0, 1, 2, 3, 5, 8, 37, 1000, 4096, (1 << 16) - 77, 1 << 16):
        m = 1
        while m < total:
            m <<= 1
        for world in (1, 2, 4, 8):
            if m < world:
                continue
            prev = 0
            for rank in range(world):
                lo, hi = ts.Twist.shard_range(total, rank, world)
                assert (lo, hi) == ts.Shout.shard_range(total, rank, world)
                assert lo == prev and lo <= hi <= total
                count = m // world
                assert hi - lo <= count and (hi == lo or (lo >= rank * count and hi <= (rank + 1) * count))
                prev = hi
            assert prev == total

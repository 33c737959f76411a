"""One-process-per-GPU sharding of the two pieces of the path that shard naturally:

* sum-check over a product of MLE tables: rank g owns the slice of every table whose reference index has high bits g
  (a contiguous range); rounds over the local variables are purely local plus ONE small all-reduce of the round
  evaluations g(0..3) per round - sent as 32 zero-extended u64 limbs so that an integer SUM all-reduce (NCCL over
  NVLink, gloo in the CPU tests) is exact, then carried and reduced mod r on the host; every rank feeds the same
  Fiat-Shamir transcript and derives the same challenge.  When one entry per table is left per rank the G x d values
  are all-gathered and each rank finishes the last log2(G) rounds on its own (they are tiny).
* KZG commitment: the MSM is sliced by points; each rank runs the full Pippenger on its slice and the G partial
  results (one G1 point each) are all-gathered and added on the host.

torch.distributed is only the plumbing; the arithmetic is libtsgpu's.  The round engine is pluggable so the host-side
logic (slicing, limb all-reduce, transcript lock-step, tail hand-over) can be tested on CPU with world_size 2 over gloo
(tests/test_distributed_cpu.py uses the oracle as the per-rank engine)."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from .binding import Context, SumCheckProof, Transcript, TwistAndShoutError, _fr, _p, lib


# ------------------------------------------------------------------------------------------------ host helpers
def round_coeffs(evals: np.ndarray) -> np.ndarray:
    evals = _fr(evals, 4)
    out = np.empty((4, 4), dtype=np.uint64)
    lib().tsgpu_sumcheck_round_coeffs(_p(evals), _p(out))
    return out


def horner(coeffs: np.ndarray, x: np.ndarray) -> np.ndarray:
    coeffs = _fr(coeffs); x = _fr(x, 1)
    out = np.empty(4, dtype=np.uint64)
    lib().tsgpu_horner_eval(_p(coeffs), C.c_size_t(coeffs.shape[0]), _p(x), _p(out))
    return out


def fr_add(a, b) -> np.ndarray:
    out = np.empty(4, dtype=np.uint64)
    lib().tsgpu_fr_add(_p(_fr(a, 1)), _p(_fr(b, 1)), _p(out))
    return out


def fr_mul(a, b) -> np.ndarray:
    out = np.empty(4, dtype=np.uint64)
    lib().tsgpu_fr_mul(_p(_fr(a, 1)), _p(_fr(b, 1)), _p(out))
    return out


def to_limb_sums(x: np.ndarray) -> np.ndarray:
    """(k, 4) uint64 field elements -> (k, 8) int64 zero-extended 32-bit limbs (the all-reduce payload)"""
    x = _fr(x)
    return np.ascontiguousarray(x.view(np.uint32).reshape(-1, 8).astype(np.int64))


def from_limb_sums(s: np.ndarray) -> np.ndarray:
    s = np.ascontiguousarray(s, dtype=np.int64).reshape(-1, 8)
    out = np.empty((s.shape[0], 4), dtype=np.uint64)
    lib().tsgpu_fr_from_limb_sums(_p(s.view(np.uint64)), C.c_size_t(s.shape[0]), _p(out))
    return out


def slice_bounds(n: int, rank: int, world: int):
    """contiguous slice of a length-n vector owned by `rank` (n and world powers of two for tables)"""
    per = n // world
    return rank * per, (rank + 1) * per


MONT_ONE = np.array([0xac96341c4ffffffb, 0x36fc76959f60cd29, 0x666ea36f7879462e, 0x0e0a77c19a07df2f], dtype=np.uint64)


# ------------------------------------------------------------------------------------------------ collectives
class Collective:
    """Thin wrapper over torch.distributed (NCCL on GPUs, gloo on CPU); world size 1 needs no process group."""

    def __init__(self, group=None, device: Optional[str] = None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.enabled = dist.is_available() and dist.is_initialized()
        self.group = group
        self.world = dist.get_world_size(group) if self.enabled else 1
        self.rank = dist.get_rank(group) if self.enabled else 0
        if device is None:
            device = "cuda" if self.enabled and dist.get_backend(group) == "nccl" else "cpu"
        self.device = device

    def all_reduce_fr(self, x: np.ndarray) -> np.ndarray:
        """exact field sum over ranks of (k, 4) elements through an integer SUM all-reduce of 32-bit limbs"""
        if self.world == 1:
            return _fr(x).copy()
        t = self.torch.from_numpy(to_limb_sums(x)).to(self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return from_limb_sums(t.cpu().numpy())

    def all_gather_u64(self, x: np.ndarray) -> np.ndarray:
        """(m,) uint64 per rank -> (world, m)"""
        x = np.ascontiguousarray(x, dtype=np.uint64).reshape(-1)
        if self.world == 1:
            return x.reshape(1, -1).copy()
        t = self.torch.from_numpy(x.view(np.int64).copy()).to(self.device)
        out = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t, group=self.group)
        return np.stack([o.cpu().numpy().view(np.uint64) for o in out])


# ------------------------------------------------------------------------------------------------ engines
class DeviceRoundEngine:
    """the per-rank round engine of the product: tables resident on this rank's GPU, tsgpu_sc_* rounds"""

    def __init__(self, ctx: Context):
        self.ctx = ctx
        self.sc = None

    def fresh(self) -> "DeviceRoundEngine":
        return DeviceRoundEngine(self.ctx)

    def begin(self, tables: Sequence[np.ndarray]):
        self.tabs = [self.ctx.table_upload(t) for t in tables]
        self.sc = self.ctx.sumcheck(self.tabs)

    def begin_device(self, tables):
        self.tabs = list(tables)
        self.sc = self.ctx.sumcheck(self.tabs)

    @property
    def vars_left(self) -> int:
        return self.sc.vars_left

    def round_eval(self) -> np.ndarray:
        return self.sc.round_eval()

    def bind_eval(self, r) -> np.ndarray:
        return self.sc.bind_eval(r)

    def bind(self, r):
        self.sc.bind(r)

    def final(self) -> np.ndarray:
        f = self.sc.final()
        self.sc.end()
        return f


# ------------------------------------------------------------------------------------------------ sharded sum-check
class ShardedSumCheck:
    """SumCheck::new(num_vars, claimed_sum).prove(|v| prod_t mle_t.evaluate(v), transcript) (src/sumcheck.rs:56-110)
    with the hypercube sliced over the ranks of `coll`.  Every rank returns the same proof."""

    def __init__(self, num_vars: int, claimed_sum, coll: Optional[Collective] = None):
        self.num_vars = num_vars
        self.claimed_sum = _fr(claimed_sum, 1).reshape(4)
        self.coll = coll or Collective()

    def prove_product(self, engine, local_tables, transcript: Transcript, device_tables: bool = False):
        """local_tables: this rank's contiguous slice of each table (reference index order), 2^(num_vars - log2 G) entries
        (host arrays, or Table handles with device_tables=True).  engine: rounds over the local slice; `engine.fresh()`
        gives the engine for the gathered G-entry tail tables."""
        coll = self.coll
        G = coll.world
        logG = G.bit_length() - 1
        if (1 << logG) != G:
            raise TwistAndShoutError(1, "number of ranks must be a power of two")
        n_local = self.num_vars - logG
        if n_local < 0:
            raise TwistAndShoutError(1, "more ranks than table entries")
        d = len(local_tables)
        if device_tables:
            engine.begin_device(local_tables)
        else:
            engine.begin(local_tables)
        if engine.vars_left != n_local:
            raise TwistAndShoutError(1, "Number of variables must match")
        current = self.claimed_sum
        zero = np.zeros(4, dtype=np.uint64)
        round_polys: List[np.ndarray] = []
        challenges: List[np.ndarray] = []

        def absorb(rnd: int, evals_total: np.ndarray):
            nonlocal current
            coeffs = round_coeffs(evals_total)
            g0, g1 = horner(coeffs, zero), horner(coeffs, MONT_ONE)
            if not (fr_add(g0, g1) == current).all():                              # sumcheck.rs:77-84
                raise TwistAndShoutError(6, f"Round {rnd} consistency check failed")
            round_polys.append(coeffs)
            transcript.append_field_elements(f"sumcheck_round_{rnd}".encode(), coeffs)
            r = transcript.challenge_field_element(f"sumcheck_challenge_{rnd}".encode())
            challenges.append(r)
            current = horner(coeffs, r)
            return r

        # ---- rounds over the local variables: partial evaluations + one all-reduce per round
        ev = engine.round_eval() if n_local else None
        for rnd in range(n_local):
            total = coll.all_reduce_fr(ev)
            r = absorb(rnd, total)
            if rnd + 1 < n_local:
                ev = engine.bind_eval(r)
            else:
                engine.bind(r)
        local_finals = engine.final()
        # ---- tail: gather one entry per table per rank (rank = high index bits), finish on every rank
        if logG:
            gathered = coll.all_gather_u64(local_finals.reshape(-1)).reshape(G, d, 4)
            tail_tables = [np.ascontiguousarray(gathered[:, t, :]) for t in range(d)]
            tail = engine.fresh()
            tail.begin(tail_tables)
            ev = tail.round_eval()
            for k in range(logG):
                r = absorb(n_local + k, ev)
                if k + 1 < logG:
                    ev = tail.bind_eval(r)
                else:
                    tail.bind(r)
            finals = tail.final()
        else:
            finals = local_finals
        fe = finals[0]
        for t in range(1, d):
            fe = fr_mul(fe, finals[t])
        proof = SumCheckProof(np.stack(round_polys) if round_polys else np.empty((0, 4, 4), dtype=np.uint64), fe)
        return proof, (np.stack(challenges) if challenges else np.empty((0, 4), dtype=np.uint64)), finals


# ------------------------------------------------------------------------------------------------ sharded commitment
def sharded_commit(local_commit, coll: Optional[Collective] = None) -> np.ndarray:
    """`local_commit`: this rank's MSM over its point slice (G1 Jacobian, uint64[12]).  Returns the sum over ranks."""
    coll = coll or Collective()
    pts = coll.all_gather_u64(np.ascontiguousarray(local_commit, dtype=np.uint64).reshape(12))
    acc = pts[0].copy()
    for g in range(1, pts.shape[0]):
        out = np.empty(12, dtype=np.uint64)
        lib().tsgpu_g1_add(_p(acc), _p(np.ascontiguousarray(pts[g])), _p(out))
        acc = out
    return acc

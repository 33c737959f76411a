"""GPU: the sharded provers of distributed.py with the product's device round engine.  World size 1 always runs;
the 2-rank NCCL case runs when the box has two GPUs (gpurun --gpus 2)."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import seed_bytes

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "multilinear-map-cryptography_b200"


def test_sharded_driver_world1_matches_oracle(ctx, tsgpu, oracle):
    import importlib
    dd = importlib.import_module(PKG + ".distributed")
    nv, d = 12, 2
    tables = [oracle.chacha_fr_rand(seed_bytes(80 + t), 1 << nv) for t in range(d)]
    # claimed sum from the oracle's table prover (round 0 evaluations)
    ev = ctx.sumcheck([ctx.table_upload(t) for t in tables]).round_eval()
    claimed = dd.fr_add(ev[0], ev[1])
    ref = oracle.sumcheck_prove_product(tables, claimed, mode="tables")
    proof, chals, finals = dd.ShardedSumCheck(nv, claimed).prove_product(dd.DeviceRoundEngine(ctx), tables, tsgpu.Transcript())
    assert (proof.round_polynomials == ref["round_polynomials"]).all() and (proof.final_evaluation == ref["final_evaluation"]).all()
    assert (finals == ref["finals"]).all()


def _nccl_worker(rank, world, port, nv, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import importlib
    import torch.distributed as dist
    ts = importlib.import_module(PKG)
    dd = importlib.import_module(PKG + ".distributed")
    import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        ctx = ts.Context(rank)
        n = 1 << nv
        tables = [O.chacha_fr_rand(bytes([60 + t]) * 32, n) for t in range(2)]
        lo, hi = dd.slice_bounds(n, rank, world)
        coll = dd.Collective()
        # claimed sum = all-reduced round-0 g(0) + g(1)
        sc = ctx.sumcheck([ctx.table_upload(t[lo:hi]) for t in tables])
        tot = coll.all_reduce_fr(sc.round_eval()); sc.end()
        claimed = dd.fr_add(tot[0], tot[1])
        proof, chals, finals = dd.ShardedSumCheck(nv, claimed, coll).prove_product(dd.DeviceRoundEngine(ctx), [t[lo:hi] for t in tables], ts.Transcript())
        ref = O.sumcheck_prove_product(tables, claimed, mode="tables")
        ok = (proof.round_polynomials == ref["round_polynomials"]).all() and (proof.final_evaluation == ref["final_evaluation"]).all()
        # point-sharded KZG commitment with per-rank SRS slices generated on the device
        tau, _ = O.setup_scalars()
        m = 1 << 12
        a, b = dd.slice_bounds(m, rank, world)
        srs = ctx.srs_generate_range(tau, a, b - a)
        poly = O.chacha_fr_rand(bytes([11]) * 32, m)
        part = ts.KZGCommitment.commit(srs, poly[a:b])
        total = dd.sharded_commit(part, coll)
        pw = O.setup_g1_powers(m, fast=True)
        ok = ok and O.g1_compress(total) == O.g1_compress(O.msm_pippenger(O.g1_batch_to_affine(pw), poly))
        q.put((rank, bool(ok)))
        ctx.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_sharded_sumcheck_and_commit_two_gpus_nccl():
    import torch.multiprocessing as mp
    mpctx = mp.get_context("spawn")
    q = mpctx.Queue()
    procs = [mpctx.Process(target=_nccl_worker, args=(r, 2, 29871, 14, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
    assert sorted(q.get(timeout=5) for _ in range(2)) == [(0, True), (1, True)]

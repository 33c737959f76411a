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


@pytest.mark.parametrize("nv,d", [(1, 2), (5, 1), (12, 2), (9, 3)])
def test_native_sharded_prove_one_rank_equals_unsharded(tsgpu, oracle, nv, d):
    """tsgpu_sumcheck_prove_product_sharded with a one-rank communicator (no NCCL needed): same proof as the oracle"""
    c = tsgpu.Context(0)
    try:
        c.comm_init(1, 0)
        assert c.comm_size == 1 and c.comm_rank == 0
        tables = [oracle.chacha_fr_rand(seed_bytes(30 + 4 * d + t + nv), 1 << nv) for t in range(d)]
        ev = c.sumcheck([c.table_upload(t) for t in tables]).round_eval()
        import importlib
        dd = importlib.import_module(PKG + ".distributed")
        claimed = dd.fr_add(ev[0], ev[1])
        ref = oracle.sumcheck_prove_product(tables, claimed, mode="tables")
        proof, chals, finals = tsgpu.SumCheck(nv, claimed).prove_product_sharded(c, [c.table_upload(t) for t in tables], tsgpu.Transcript(), return_aux=True)
        assert (proof.round_polynomials == ref["round_polynomials"]).all() and (proof.final_evaluation == ref["final_evaluation"]).all()
        assert (finals == ref["finals"]).all()
        pt = oracle.chacha_fr_rand(seed_bytes(nv + 3), nv).reshape(nv, 4)
        assert (c.table_upload(tables[0]).evaluate_sharded(nv, pt) == oracle.mle_evaluate(tables[0], pt, fold=True)).all()
        with pytest.raises(tsgpu.TwistAndShoutError) as e:       # wrong claim: the reference's round-0 error
            tsgpu.SumCheck(nv, tsgpu.fe(12345)).prove_product_sharded(c, [c.table_upload(t) for t in tables], tsgpu.Transcript())
        assert e.value.variant == "SumCheck" and "Round 0 consistency check failed" in str(e.value)
    finally:
        c.close()


def _native_worker(rank, world, port, nv, d, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import importlib
    import torch.distributed as dist
    ts = importlib.import_module(PKG)
    dd = importlib.import_module(PKG + ".distributed")
    import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)          # only carries the NCCL id; the data path is the library's NCCL
    try:
        ctx = ts.Context(rank)
        ctx.comm_init_torch()
        n = 1 << nv
        tables = [O.chacha_fr_rand(bytes([70 + t]) * 32, n) for t in range(d)]
        lo, hi = dd.slice_bounds(n, rank, world)
        full = [np.asarray(t).reshape(n, 4) for t in tables]
        ints = [O.fr_to_ints(t) for t in full]
        tot = 0
        for i in range(n):
            p = 1
            for t in range(d):
                p = p * ints[t][i] % O.R_MOD
            tot = (tot + p) % O.R_MOD
        claimed = O.fr_from_ints([tot])[0]
        ref = O.sumcheck_prove_product(tables, claimed, mode="tables")
        ok = True
        # both exchange paths (peer mailboxes: the round sums travel inside the round kernel; NCCL: all-reduce per round), both round-0 forms, and a
        # wrong claim, which every rank must reject with the reference's error
        for peer in ((1, 0) if ctx.comm_peer_exchange else (0,)):
            ctx.set_tuning("peer_exchange", peer)
            for deferred in (0, 1):
                ctx.set_tuning("deferred_claim_check", deferred)
                proof, chals, finals = ts.SumCheck(nv, claimed).prove_product_sharded(ctx, [ctx.table_upload(t[lo:hi]) for t in full], ts.Transcript(), return_aux=True)
                ok = ok and (proof.round_polynomials == ref["round_polynomials"]).all() and (proof.final_evaluation == ref["final_evaluation"]).all()
                ok = ok and (finals == ref["finals"]).all() and (chals == ref["challenges"]).all()
                try:
                    ts.SumCheck(nv, ts.fe(12345)).prove_product_sharded(ctx, [ctx.table_upload(t[lo:hi]) for t in full], ts.Transcript())
                    ok = False
                except ts.TwistAndShoutError as e:
                    ok = ok and e.variant == "SumCheck" and "Round 0 consistency check failed" in str(e)
            ctx.set_tuning("deferred_claim_check", 0)
        ctx.set_tuning("peer_exchange", 1)
        q.put(("peer_exchange", rank, ctx.comm_peer_exchange))
        # sharded MultilinearExtension::evaluate of table 0
        pt = O.chacha_fr_rand(bytes([90]) * 32, nv).reshape(nv, 4)
        got = ctx.table_upload(full[0][lo:hi]).evaluate_sharded(nv, pt)
        ok = ok and (got == O.mle_evaluate(tables[0], pt, fold=True)).all()
        # host all-gather through the library communicator
        g = ctx.comm_allgather(np.arange(12, dtype=np.uint64) + 100 * rank)
        ok = ok and all((g[r] == np.arange(12, dtype=np.uint64) + 100 * r).all() for r in range(world))
        q.put((rank, bool(ok)))
        ctx.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("nv,d", [(13, 2), (10, 3), (2, 2)])
def test_native_sharded_sumcheck_two_gpus_library_nccl(nv, d):
    import torch.multiprocessing as mp
    mpctx = mp.get_context("spawn")
    q = mpctx.Queue()
    procs = [mpctx.Process(target=_native_worker, args=(r, 2, 29873 + nv, nv, d, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
    got = [q.get(timeout=5) for _ in range(4)]
    assert sorted(x for x in got if len(x) == 2) == [(0, True), (1, True)]
    peers = sorted(x[1:] for x in got if len(x) == 3)
    assert peers[0][1] == peers[1][1]                       # both ranks agree on the exchange path (NVLink boxes: peer mailboxes)


def test_sharded_twist_one_rank_equals_plain_prove(tsgpu):
    """tsgpu_twist_prove_sharded with a one-rank communicator: the slice code path (basis slice = whole basis) gives the same bytes"""
    c = tsgpu.Context(0)
    try:
        c.comm_init(1, 0)
        pp, vp = tsgpu.setup_params(c, 10)
        for nops in (2, 37, 1000, 4096):
            rng = np.random.default_rng(nops)
            addr = rng.integers(0, 1 << 10, size=nops).astype(np.uint64)
            vals = tsgpu.fe_vec(rng.integers(0, 1 << 63, size=nops, dtype=np.uint64))
            tw = tsgpu.Twist.new(pp)
            a = tw.prove_arrays(addr, vals)
            b = tw.prove_sharded(addr, vals, nops)
            assert a.to_bytes() == b.to_bytes() and tw.verify(b, vp)
    finally:
        c.close()


def test_sharded_shout_one_rank_equals_plain_prove(tsgpu):
    """tsgpu_shout_prove_sharded with a one-rank communicator, table and lookup vectors of equal and of different padded lengths"""
    c = tsgpu.Context(0)
    try:
        c.comm_init(1, 0)
        pp, vp = tsgpu.setup_params(c, 10)
        sh = tsgpu.Shout.new(pp)
        for nent, nlook in ((3, 4), (8, 5), (1000, 37), (64, 4096), (1024, 1024), (4096, 1)):
            rng = np.random.default_rng(nent * 7 + nlook)
            entries = tsgpu.fe_vec(rng.integers(0, 1 << 63, size=nent, dtype=np.uint64))
            idx = rng.integers(0, nent, size=nlook).astype(np.uint64)
            a = sh.prove_arrays(entries, idx)
            b = sh.prove_sharded(entries, nent, idx, nlook)
            assert a.to_bytes() == b.to_bytes() and sh.verify(b, vp)
        with pytest.raises(tsgpu.TwistAndShoutError) as e:
            sh.prove_sharded(entries, 4096, np.zeros(4097, dtype=np.uint64), 4097)
        assert e.value.message == "Too many lookup operations"
        with pytest.raises(tsgpu.TwistAndShoutError) as e:                       # a rank must pass exactly its range
            sh.prove_sharded(entries[:10], 4096, idx, 1)
        assert e.value.variant == "InvalidParameters"
    finally:
        c.close()


def _twist_worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import importlib
    import torch.distributed as dist
    ts = importlib.import_module(PKG)
    import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ctx = ts.Context(rank)
        ctx.comm_init_torch()
        pp, vp = ts.setup_params(ctx, 14)
        ok = True
        for nops, wide in ((2, False), (5, False), (1000, False), (4096, True), ((1 << 16) - 77, False), (1 << 16, True)):
            rng = np.random.default_rng(nops)                       # same trace on every rank
            addr = rng.integers(0, 1 << 14, size=nops).astype(np.uint64)
            vals = O.chacha_fr_rand(bytes([nops % 251]) * 32, nops).reshape(nops, 4) if wide else ts.fe_vec(rng.integers(0, 1 << 63, size=nops, dtype=np.uint64))
            tw = ts.Twist.new(pp)
            lo, hi = tw.shard_range(nops, rank, world)
            ctx.set_tuning("peer_exchange", 0)                      # NCCL all-gathers
            via_nccl = tw.prove_sharded(addr[lo:hi], vals[lo:hi], nops)
            ctx.set_tuning("peer_exchange", 1)                      # single-kernel all-gathers over the peer mailboxes (when mapped)
            sharded = tw.prove_sharded(addr[lo:hi], vals[lo:hi], nops)
            ok = ok and tw.verify(sharded, vp) and sharded.to_bytes() == via_nccl.to_bytes()
            if rank == 0:
                ok = ok and sharded.to_bytes() == tw.prove_arrays(addr, vals).to_bytes()      # one-GPU proof of the whole trace
            gathered = ctx.comm_allgather(np.frombuffer(sharded.to_bytes()[:64], dtype=np.uint64))
            ok = ok and (gathered[0] == gathered[1]).all()                                    # both ranks hold the same commitments
        sh = ts.Shout.new(pp)
        for nent, nlook in ((2, 2), (3, 7), (1000, 4096), (1 << 14, 1 << 16), (1 << 12, (1 << 12) - 5), ((1 << 14) - 3, 100)):
            rng = np.random.default_rng(nent + 3 * nlook)
            entries = O.chacha_fr_rand(bytes([nent % 251]) * 32, nent).reshape(nent, 4) if nent == 1000 else ts.fe_vec(rng.integers(0, 1 << 63, size=nent, dtype=np.uint64))
            idx = rng.integers(0, nent, size=nlook).astype(np.uint64)
            elo, ehi = sh.shard_range(nent, rank, world); llo, lhi = sh.shard_range(nlook, rank, world)
            sharded = sh.prove_sharded(entries[elo:ehi], nent, idx[llo:lhi], nlook)
            ok = ok and sh.verify(sharded, vp)
            if rank == 0:
                ok = ok and sharded.to_bytes() == sh.prove_arrays(entries, idx).to_bytes()
            gathered = ctx.comm_allgather(np.frombuffer(sharded.to_bytes()[:64], dtype=np.uint64))
            ok = ok and (gathered[0] == gathered[1]).all()
        q.put((rank, bool(ok)))
        ctx.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_sharded_twist_and_shout_two_gpus_same_bytes_as_one_gpu():
    import torch.multiprocessing as mp
    mpctx = mp.get_context("spawn")
    q = mpctx.Queue()
    procs = [mpctx.Process(target=_twist_worker, args=(r, 2, 29891, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
    assert sorted(q.get(timeout=5) for _ in range(2)) == [(0, True), (1, True)]


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

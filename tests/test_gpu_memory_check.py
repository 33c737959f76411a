"""GPU: the memory-consistency sum-checks of Twist (read-checking over (cell, cycle) + Val-evaluation) that the reference leaves as a stub
(src/twist.rs:181-214), and their table builders.  Both parts are the reference's own SumCheck::prove (src/sumcheck.rs:56-110) applied to
product closures: the CPU oracle runs exactly that on the same transcript (closure form at tiny sizes, table form above) over tables built
here in plain Python, and the device result must match bit for bit.  Non-parity mode: Twist::prove itself is unchanged."""
import numpy as np
import pytest

from conftest import seed_bytes

pytestmark = pytest.mark.gpu


def _trace(oracle, tsgpu, K, n, seed, wide=False):
    """random trace on K cells: (addresses, values Fr, is_write) with reads returning the simulated content (MemoryTrace semantics)"""
    rng = np.random.default_rng(seed)
    addr = rng.integers(0, K, size=n).astype(np.uint64)
    isw = rng.integers(0, 2, size=n).astype(np.uint8)
    fresh = oracle.chacha_fr_rand(seed_bytes(seed), max(n, 1)).reshape(-1, 4) if wide else tsgpu.fe_vec(rng.integers(1, 1 << 63, size=max(n, 1), dtype=np.uint64))
    vals = np.zeros((n, 4), dtype=np.uint64)
    mem = {}
    for j in range(n):
        a = int(addr[j])
        if isw[j]:
            mem[a] = fresh[j]
        vals[j] = mem.get(a, np.zeros(4, dtype=np.uint64))
    return addr, vals, isw


def _eq_ints(oracle, pt):
    return oracle._eq_ints(pt)


def _lt_point_ints(oracle, b_ints, t, p):
    return oracle.lt_point_ints(b_ints, t)


def _oracle_memory_check(oracle, addr, vals, isw, K, mode):
    """host/memory_check.cpp restated with oracle primitives and Python integers (oracle/oracle.py: twist_memory_check_prove)"""
    return oracle.twist_memory_check_prove(addr, vals, isw, K, mode, with_write_check=True)


@pytest.mark.parametrize("K,n,mode", [(1, 1, "closure"), (2, 2, "closure"), (4, 3, "closure"), (2, 8, "closure"), (8, 6, "closure"),
                                      (4, 16, "tables"), (16, 50, "tables"), (64, 64, "tables"), (8, 500, "tables"), (256, 30, "tables")])
def test_memory_check_matches_reference_sumcheck_on_the_real_closures(ctx, tsgpu, oracle, K, n, mode):
    addr, vals, isw = _trace(oracle, tsgpu, K, n, seed=K * 1000 + n, wide=(n % 2 == 0))
    mc = tsgpu.TwistMemoryCheck(ctx)
    proof = mc.prove_arrays(addr, vals, isw, K, tsgpu.Transcript())
    c1, c2, ref1, ref2, c3, c4, ref3, ref4 = _oracle_memory_check(oracle, addr, vals, isw, K, mode)
    assert (proof.claims[0] == c1).all() and (proof.claims[1] == c2).all()
    assert (proof.read_check.round_polynomials == ref1["round_polynomials"]).all() and (proof.read_check.final_evaluation == ref1["final_evaluation"]).all()
    assert (proof.val_evaluation.round_polynomials == ref2["round_polynomials"]).all() and (proof.val_evaluation.final_evaluation == ref2["final_evaluation"]).all()
    # write-checking (third sum-check) and the Val-evaluation of its closing claim
    assert (proof.write_claims[0] == c3).all() and (proof.write_claims[1] == c4).all()
    assert (proof.write_check.round_polynomials == ref3["round_polynomials"]).all() and (proof.write_check.final_evaluation == ref3["final_evaluation"]).all()
    assert (proof.write_val_evaluation.round_polynomials == ref4["round_polynomials"]).all() and (proof.write_val_evaluation.final_evaluation == ref4["final_evaluation"]).all()
    assert mc.verify_arrays(addr, vals, isw, K, proof, tsgpu.Transcript())
    if n >= 2 and isw.any():
        # a tampered write-checking part must not verify: claim, a round coefficient, the final evaluation
        import copy
        for field in ("write_claims", "write_check", "write_val_evaluation"):
            bad = copy.deepcopy(proof)
            if field == "write_claims":
                bad.write_claims[1] = tsgpu.fe_add(bad.write_claims[1], tsgpu.fe(1))
            else:
                getattr(bad, field).final_evaluation[:] = tsgpu.fe_add(getattr(bad, field).final_evaluation, tsgpu.fe(1))
            assert not mc.verify_arrays(addr, vals, isw, K, bad, tsgpu.Transcript()), field


def test_memory_check_reference_demo_trace(ctx, tsgpu, oracle):
    """examples/demo.rs:33-46: 8 cells, W(0,42) W(1,100) R(0) R(1) W(0,43) R(0); README: 256 cells, W(0,42) W(1,100) R(0)"""
    for cells, ops in ((8, [("W", 0, 42), ("W", 1, 100), ("R", 0, 0), ("R", 1, 0), ("W", 0, 43), ("R", 0, 0)]), (256, [("W", 0, 42), ("W", 1, 100), ("R", 0, 0)])):
        trace = tsgpu.MemoryTrace.new(cells)
        for kind, a, v in ops:
            trace.write(a, tsgpu.fe(v)) if kind == "W" else trace.read(a)
        mc = tsgpu.TwistMemoryCheck(ctx)
        proof = mc.prove(trace, tsgpu.Transcript())
        assert proof.read_check.round_polynomials.shape[0] == (cells.bit_length() - 1) + (len(ops) - 1).bit_length()
        assert proof.read_check.round_polynomials.any()               # a real constraint: not the all-zero cubics of the stub
        assert mc.verify(trace, proof, tsgpu.Transcript())


def test_inconsistent_read_is_rejected(ctx, tsgpu, oracle):
    K, n = 16, 40
    addr, vals, isw = _trace(oracle, tsgpu, K, n, seed=77)
    mc = tsgpu.TwistMemoryCheck(ctx)
    j = int(np.flatnonzero(isw == 0)[3])
    bad = vals.copy(); bad[j] = tsgpu.fe(123456789)                                   # this read returns something that was never written there
    with pytest.raises(tsgpu.TwistAndShoutError) as e:                                 # the prover's own check (sumcheck.rs:77-84)
        mc.prove_arrays(addr, bad, isw, K, tsgpu.Transcript())
    assert e.value.variant == "SumCheck" and e.value.message == "Round 0 consistency check failed"
    proof = mc.prove_arrays(addr, vals, isw, K, tsgpu.Transcript())
    assert mc.verify_arrays(addr, vals, isw, K, proof, tsgpu.Transcript())
    assert not mc.verify_arrays(addr, bad, isw, K, proof, tsgpu.Transcript())         # the honest proof does not fit the false statement
    addr2 = addr.copy(); addr2[j] = (addr2[j] + 1) % K
    assert not mc.verify_arrays(addr2, vals, isw, K, proof, tsgpu.Transcript())
    for part in ("read_check", "val_evaluation"):                                      # tampering with either part
        t = tsgpu.TwistMemoryCheck.Proof(proof.claims.copy(), tsgpu.SumCheckProof(proof.read_check.round_polynomials.copy(), proof.read_check.final_evaluation.copy()),
                                         tsgpu.SumCheckProof(proof.val_evaluation.round_polynomials.copy(), proof.val_evaluation.final_evaluation.copy()))
        getattr(t, part).final_evaluation[0] ^= 1
        assert not mc.verify_arrays(addr, vals, isw, K, t, tsgpu.Transcript())
    t = tsgpu.TwistMemoryCheck.Proof(proof.claims.copy(), proof.read_check, proof.val_evaluation)
    t.claims[1][0] ^= 1                                                                # a wrong Val~(x*, j*)
    assert not mc.verify_arrays(addr, vals, isw, K, t, tsgpu.Transcript())
    with pytest.raises(tsgpu.TwistAndShoutError) as e:                                 # twist.rs:49-53
        mc.prove_arrays(np.array([K], dtype=np.uint64), vals[:1], isw[:1], K, tsgpu.Transcript())
    assert e.value.message == "Address out of bounds"
    with pytest.raises(tsgpu.TwistAndShoutError) as e:                                 # twist.rs:38
        mc.prove_arrays(addr, vals, isw, 12, tsgpu.Transcript())
    assert e.value.message == "Memory size must be power of 2"


@pytest.mark.parametrize("k,t", [(0, 0), (3, 2), (2, 6), (6, 7)])
def test_memory_table_builders(ctx, tsgpu, oracle, k, t):
    """Val(x, j), the weighted one-hot matrix, the elementwise product and LT~(., b) against plain Python"""
    import ctypes as C
    import importlib
    bd = importlib.import_module(tsgpu.__name__ + ".binding")
    p = oracle.R_MOD
    K, T = 1 << k, 1 << t
    n = T - T // 4
    addr, vals, isw = _trace(oracle, tsgpu, K, n, seed=5 * k + t, wide=True)
    vi = oracle.fr_to_ints(vals) if n else []
    h = C.c_void_p()
    ctx.check(tsgpu.lib().tsgpu_table_memory_values(ctx._h, bd._p(addr), bd._p(isw), bd._p(vals), C.c_size_t(n), C.c_uint(k), C.c_uint(t), C.byref(h)))
    got = oracle.fr_to_ints(bd.Table(ctx, h).download())
    mem = [0] * K
    for j in range(T):
        assert [got[x + K * j] for x in range(K)] == mem, f"Val(., {j})"
        if j < n and isw[j]:
            mem[int(addr[j])] = vi[j]
    w = oracle.chacha_fr_rand(seed_bytes(31 + t), T).reshape(T, 4)
    W = ctx.table_upload(w, t)
    wi = oracle.fr_to_ints(w)
    for flag in (0, 1):
        h = C.c_void_p()
        ctx.check(tsgpu.lib().tsgpu_table_one_hot_weighted(ctx._h, W._h, bd._p(addr), bd._p(isw), C.c_int(flag), C.c_size_t(n), C.c_uint(k), C.byref(h)))
        got = oracle.fr_to_ints(bd.Table(ctx, h).download())
        want = [0] * (K * T)
        for j in range(n):
            if isw[j] == flag:
                want[int(addr[j]) + K * j] = wi[j]
        assert got == want
    h = C.c_void_p()
    ctx.check(tsgpu.lib().tsgpu_table_mul(ctx._h, W._h, W._h, C.byref(h)))
    assert oracle.fr_to_ints(bd.Table(ctx, h).download()) == [x * x % p for x in wi]
    b = oracle.chacha_fr_rand(seed_bytes(41 + t), max(t, 1)).reshape(-1, 4)[:t]
    h = C.c_void_p()
    ctx.check(tsgpu.lib().tsgpu_table_lt_point(ctx._h, bd._p(np.ascontiguousarray(b)) if t else None, C.c_uint(t), C.byref(h)))
    lt = bd.Table(ctx, h)
    assert oracle.fr_to_ints(lt.download()) == _lt_point_ints(oracle, oracle.fr_to_ints(b) if t else [], t, p)
    # at a boolean point c the table is the indicator [a < c]
    if t:
        c = 5 % T
        bits = oracle.fr_from_ints([(c >> i) & 1 for i in range(t)])
        h = C.c_void_p()
        ctx.check(tsgpu.lib().tsgpu_table_lt_point(ctx._h, bd._p(bits), C.c_uint(t), C.byref(h)))
        assert oracle.fr_to_ints(bd.Table(ctx, h).download()) == [1 if a < c else 0 for a in range(T)]


def test_memory_check_config4_shape(ctx, tsgpu, oracle):
    """2^8 cells x 2^14 cycles = 2^22-entry tables (BASELINE config 4 scaled down 16x): the reference generator pattern of src/benchmarks.rs:88-99
    proves and verifies; 22 + 14 rounds"""
    K, n = 1 << 8, 1 << 14
    i = np.arange(n, dtype=np.uint64)
    isw = (i % 3 == 0).astype(np.uint8)
    addr = np.where(isw == 1, i % K, (i // 2) % K).astype(np.uint64)
    vals_u = np.zeros(n, dtype=np.uint64); mem = {}
    for j in range(n):
        a = int(addr[j])
        if isw[j]:
            mem[a] = 42 * j
        vals_u[j] = mem.get(a, 0)
    vals = tsgpu.fe_vec(vals_u)
    mc = tsgpu.TwistMemoryCheck(ctx)
    proof = mc.prove_arrays(addr, vals, isw, K, tsgpu.Transcript())
    assert proof.read_check.round_polynomials.shape == (22, 4, 4) and proof.val_evaluation.round_polynomials.shape == (14, 4, 4)
    assert mc.verify_arrays(addr, vals, isw, K, proof, tsgpu.Transcript())


def test_memory_check_challenges_depend_on_the_trace(ctx, tsgpu, oracle):
    """Fiat-Shamir binding (see tests/test_gpu_read_check.py): two read values shifted so that the read claim at a statement-independent
    point r0 is unchanged - a false trace that a transcript ignoring the statement would accept - are rejected by prover and verifier"""
    p = oracle.R_MOD
    K, n = 8, 16
    addr, vals, isw = _trace(oracle, tsgpu, K, n, seed=77)
    reads = [j for j in range(n) if not isw[j]]
    assert len(reads) >= 2
    a, b = reads[0], reads[-1]
    r0 = tsgpu.Transcript().challenge_field_elements(b"memory_check_point", 4)
    eq0 = oracle.fr_to_ints(oracle.eq_table(r0))
    v = oracle.fr_to_ints(vals)
    delta = 99
    v[a] = (v[a] + delta) % p
    v[b] = (v[b] - delta * eq0[a] * pow(eq0[b], -1, p)) % p
    forged = oracle.fr_from_ints(v)
    mc = tsgpu.TwistMemoryCheck(ctx)
    with pytest.raises(tsgpu.TwistAndShoutError) as e:
        mc.prove_arrays(addr, forged, isw, K, tsgpu.Transcript())
    assert e.value.variant == "SumCheck" and e.value.message == "Round 0 consistency check failed"
    proof = mc.prove_arrays(addr, vals, isw, K, tsgpu.Transcript())
    assert mc.verify_arrays(addr, vals, isw, K, proof, tsgpu.Transcript())
    assert not mc.verify_arrays(addr, forged, isw, K, proof, tsgpu.Transcript())


def test_memory_check_bound_to_the_commitments_of_a_twist_proof(ctx, tsgpu, oracle):
    """src/twist.rs:181-214 ("a production implementation ... tie to the commitments"): the constraint sum-checks run on a transcript that first absorbed the
    address / value commitment hashes of the byte-identical Twist proof (labels of twist.rs:157-160), and the verifier recomputes both commitments from the
    clear statement.  The oracle runs the same sum-checks on ITS transcript after appending ITS hashes of ITS commitments."""
    pp, vp = tsgpu.setup_params(ctx, 6)
    trace = tsgpu.MemoryTrace.new(16)
    rng = np.random.default_rng(11)
    for j in range(50):
        a = int(rng.integers(0, 16))
        trace.write(a, tsgpu.fe(int(rng.integers(1, 1 << 62)))) if rng.integers(0, 2) else trace.read(a)
    tw = tsgpu.Twist.new(pp)
    proof = tw.prove(trace)
    assert tw.verify(proof, vp)
    mc = tsgpu.TwistMemoryCheck(ctx)
    cproof = mc.prove(trace, mc.bind(tsgpu.Transcript(), proof))
    assert mc.verify(trace, cproof, mc.bind(tsgpu.Transcript(), proof))
    assert mc.commitments_match(pp, proof, trace)
    # oracle: same protocol on the oracle's transcript, hashes of the oracle's own commitments (Jacobian points out of the oracle's proof parser are not needed:
    # the proof bytes already equal the oracle's - test_gpu_protocols - so the device commitments are the oracle's; the hash is the oracle's g1_hash)
    addr, vals, isw = trace.arrays()
    otr = oracle.Transcript()
    otr.append_field_element(b"address_commitment", oracle.g1_hash(proof.commitments[0]))
    otr.append_field_element(b"value_commitment", oracle.g1_hash(proof.commitments[1]))
    c1, c2, ref1, ref2, c3, c4, ref3, ref4 = oracle.twist_memory_check_prove(addr, vals, isw, 16, "tables", with_write_check=True, transcript=otr)
    assert (cproof.claims[0] == c1).all() and (cproof.read_check.round_polynomials == ref1["round_polynomials"]).all()
    assert (cproof.write_claims[0] == c3).all() and (cproof.write_val_evaluation.round_polynomials == ref4["round_polynomials"]).all()
    # unbound, or bound to the proof of ANOTHER trace: every challenge differs, nothing verifies
    assert not mc.verify(trace, cproof, tsgpu.Transcript())
    other = tsgpu.MemoryTrace.new(16)
    for op in trace.operations[:-1]:
        other.write(op.address, op.value) if op.kind == "W" else other.read(op.address)
    other.write(3, tsgpu.fe(999))
    proof2 = tw.prove(other)
    assert not mc.verify(trace, cproof, mc.bind(tsgpu.Transcript(), proof2))
    assert not mc.commitments_match(pp, proof2, trace) and mc.commitments_match(pp, proof2, other)
    # same trace under the coefficient path commits to the same points
    ctx.set_tuning("eval_basis", 0)
    try:
        assert mc.commitments_match(pp, proof, trace) and not mc.commitments_match(pp, proof, other)
    finally:
        ctx.set_tuning("eval_basis", 1)

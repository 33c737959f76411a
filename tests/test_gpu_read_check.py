"""GPU: the lookup-correctness sum-check (core Shout read-checking) that the reference leaves as a stub (src/shout.rs:157-184), and its
building blocks (scatter_add / gather / inner_product).  The proof is the reference's own SumCheck::prove (src/sumcheck.rs:56-110) applied
to the closure |x| ra~(x, r) * Val~(x): the CPU oracle runs exactly that - closure form at small sizes, table form above - on the same
transcript, and the device result must match it bit for bit.  Non-parity mode: Shout::prove itself is unchanged (tests/test_gpu_protocols.py)."""
import numpy as np
import pytest

from conftest import seed_bytes

pytestmark = pytest.mark.gpu


def _statement(oracle, tsgpu, nent, nlook, seed, wide=True):
    rng = np.random.default_rng(seed)
    entries = oracle.chacha_fr_rand(seed_bytes(seed), nent).reshape(nent, 4) if wide else tsgpu.fe_vec(rng.integers(0, 1 << 63, size=nent, dtype=np.uint64))
    idx = rng.integers(0, nent, size=nlook).astype(np.uint64)
    return entries, idx, entries[idx.astype(np.int64)] if nlook else np.empty((0, 4), dtype=np.uint64)


def _oracle_read_check(oracle, entries, idx, vals, mode):
    """the protocol of host/read_check.cpp restated with oracle primitives (oracle/oracle.py: shout_read_check_prove)"""
    return oracle.shout_read_check_prove(entries, idx, vals, mode)


@pytest.mark.parametrize("nent,nlook,mode", [(1, 1, "closure"), (2, 1, "closure"), (3, 4, "closure"), (8, 5, "closure"), (16, 37, "closure"),
                                             (50, 8, "closure"), (1000, 300, "tables"), (4096, 1 << 14, "tables"), ((1 << 14) - 3, 1000, "tables")])
def test_read_check_matches_reference_sumcheck_on_the_real_closure(ctx, tsgpu, oracle, nent, nlook, mode):
    entries, idx, vals = _statement(oracle, tsgpu, nent, nlook, seed=nent + nlook)
    claim, proof, ch = tsgpu.ShoutReadCheck(ctx).prove_arrays(entries, idx, vals, tsgpu.Transcript())
    want_claim, ref = _oracle_read_check(oracle, entries, idx, vals, mode)
    assert (claim == want_claim).all()
    assert (proof.round_polynomials == ref["round_polynomials"]).all()
    assert (proof.final_evaluation == ref["final_evaluation"]).all()
    assert (ch == ref["challenges"]).all()
    assert tsgpu.ShoutReadCheck(ctx).verify_arrays(entries, idx, vals, proof, tsgpu.Transcript())
    if proof.round_polynomials.shape[0]:
        assert proof.round_polynomials.any()                 # a real constraint: not the all-zero cubics of the stub


def test_read_check_reference_demo_table(ctx, tsgpu, oracle):
    """examples/demo.rs:66-79: squares table 0..7, lookups [3, 5, 0, 7]; README: [1, 4, 9], lookup(1)"""
    for entries, looks in (([i * i for i in range(8)], [3, 5, 0, 7]), ([1, 4, 9], [1])):
        table = tsgpu.LookupTable.new(tsgpu.fe_vec(entries))
        for i in looks:
            assert tsgpu.fe_to_int(table.lookup(i)) == entries[i]
        rc = tsgpu.ShoutReadCheck(ctx)
        claim, proof, _ = rc.prove(table, tsgpu.Transcript())
        assert rc.verify(table, proof, tsgpu.Transcript())


def test_wrong_lookup_value_is_rejected(ctx, tsgpu, oracle):
    entries, idx, vals = _statement(oracle, tsgpu, 64, 100, seed=5)
    rc = tsgpu.ShoutReadCheck(ctx)
    bad = vals.copy(); bad[17] = entries[(int(idx[17]) + 1) % 64]            # lookup 17 claims the neighbouring entry
    with pytest.raises(tsgpu.TwistAndShoutError) as e:                       # the prover's own check (sumcheck.rs:77-84)
        rc.prove_arrays(entries, idx, bad, tsgpu.Transcript())
    assert e.value.variant == "SumCheck" and e.value.message == "Round 0 consistency check failed"
    # an honest proof of the true statement does not verify against the false one, nor against other indices, nor when tampered
    claim, proof, _ = rc.prove_arrays(entries, idx, vals, tsgpu.Transcript())
    assert rc.verify_arrays(entries, idx, vals, proof, tsgpu.Transcript())
    assert not rc.verify_arrays(entries, idx, bad, proof, tsgpu.Transcript())
    idx2 = idx.copy(); idx2[3] = (idx2[3] + 1) % 64
    assert not rc.verify_arrays(entries, idx2, vals, proof, tsgpu.Transcript())
    t = tsgpu.SumCheckProof(proof.round_polynomials.copy(), proof.final_evaluation.copy())
    t.final_evaluation[0] ^= 1
    assert not rc.verify_arrays(entries, idx, vals, t, tsgpu.Transcript())
    with pytest.raises(tsgpu.TwistAndShoutError) as e:                       # sumcheck.rs:118-122
        rc.verify_arrays(entries, idx, vals, tsgpu.SumCheckProof(proof.round_polynomials[:-1], proof.final_evaluation), tsgpu.Transcript())
    assert e.value.message == "Proof has wrong number of rounds"
    with pytest.raises(tsgpu.TwistAndShoutError) as e:                       # shout.rs:44-48
        rc.prove_arrays(entries, np.array([64], dtype=np.uint64), vals[:1], tsgpu.Transcript())
    assert e.value.message == "Lookup index out of bounds"


@pytest.mark.parametrize("l,k,skew", [(0, 0, False), (5, 3, False), (12, 12, False), (16, 10, False), (16, 10, True), (18, 20, False)])
def test_scatter_add_gather_inner_product(ctx, tsgpu, oracle, l, k, skew):
    L, K = 1 << l, 1 << k
    p = oracle.R_MOD
    rng = np.random.default_rng(l * 31 + k)
    n = L - (L // 5)                                                          # fewer indices than weights: the tail does not contribute
    w = oracle.chacha_fr_rand(seed_bytes(l + 2 * k + 1), L).reshape(L, 4)
    idx = (np.full(n, 7 % K) if skew else rng.integers(0, K, size=n)).astype(np.uint64)   # skew: every atomic hits one bucket
    W = ctx.table_upload(w)
    got = W.scatter_add(idx, k)
    assert got.num_vars == k
    wi = oracle.fr_to_ints(w)
    want = [0] * K
    for j in range(n):
        want[int(idx[j])] += wi[j]
    assert (got.download() == oracle.fr_from_ints([x % p for x in want])).all()
    src = oracle.chacha_fr_rand(seed_bytes(99 + k), K).reshape(K, 4)
    G = ctx.table_upload(src).gather(idx, l)
    back = G.download()
    assert (back[:n] == src[idx.astype(np.int64)]).all() and (back[n:] == 0).all()
    # <W, gather(S, idx)> == <scatter_add(W, idx), S>: the adjoint identity the verifier's closing check relies on
    lhs = oracle.fr_to_ints(W.inner_product(G))[0]
    rhs = oracle.fr_to_ints(got.inner_product(ctx.table_upload(src)))[0]
    assert lhs == rhs == sum(wi[j] * oracle.fr_to_ints(src[int(idx[j])])[0] for j in range(min(n, 64))) % p if n <= 64 else lhs == rhs


def test_read_check_benchmark_shape(ctx, tsgpu, oracle):
    """2^16-entry table of squares, 2^18 lookups i mod T (the generator of src/benchmarks.rs:167-177): proves, verifies, and the claim
    equals the multilinear extension of the returned values at the transcript's point"""
    T, L = 1 << 16, 1 << 18
    i = np.arange(T, dtype=np.uint64)
    entries = tsgpu.fe_vec(i * i)
    idx = (np.arange(L, dtype=np.uint64) % T).astype(np.uint64)
    vals = entries[idx.astype(np.int64)]
    rc = tsgpu.ShoutReadCheck(ctx)
    claim, proof, ch = rc.prove_arrays(entries, idx, vals, tsgpu.Transcript())
    assert proof.round_polynomials.shape == (16, 4, 4)
    tr = tsgpu.Transcript()
    tr.append_field_elements(b"read_check_statement", oracle.statement_digest_elements(
        b"shout_read_check", [T, L], [entries.tobytes(), idx.tobytes(), np.ascontiguousarray(vals).tobytes()]))
    r = tr.challenge_field_elements(b"read_check_point", 18)
    assert (ctx.mle_evaluate(vals, r) == claim).all()
    assert rc.verify_arrays(entries, idx, vals, proof, tsgpu.Transcript())


def test_challenges_depend_on_the_statement(ctx, tsgpu, oracle):
    """Fiat-Shamir binding: the point r is drawn AFTER the statement digest is absorbed.  Against a statement-independent r0 (what a fresh
    transcript would hand out) one can shift two returned values so that sum_j eq(r0, j)(v'_j - v_j) = 0: the false statement would share the
    claim rv~(r0) with the true one and prove / verify.  With the digest in the transcript the false statement is rejected on both sides."""
    p = oracle.R_MOD
    entries, idx, vals = _statement(oracle, tsgpu, 32, 16, seed=9)
    r0 = tsgpu.Transcript().challenge_field_elements(b"read_check_point", 4)            # the point of a transcript that ignores the statement
    eq0 = oracle.fr_to_ints(oracle.eq_table(r0))
    v = oracle.fr_to_ints(vals)
    delta = 12345
    v[2] = (v[2] + delta) % p
    v[7] = (v[7] - delta * eq0[2] * pow(eq0[7], -1, p)) % p                             # keeps sum_j eq(r0, j) v_j unchanged
    forged = oracle.fr_from_ints(v)
    assert sum(e * a for e, a in zip(eq0, v)) % p == sum(e * a for e, a in zip(eq0, oracle.fr_to_ints(vals))) % p
    rc = tsgpu.ShoutReadCheck(ctx)
    with pytest.raises(tsgpu.TwistAndShoutError) as e:
        rc.prove_arrays(entries, idx, forged, tsgpu.Transcript())
    assert e.value.variant == "SumCheck" and e.value.message == "Round 0 consistency check failed"
    claim, proof, _ = rc.prove_arrays(entries, idx, vals, tsgpu.Transcript())
    assert rc.verify_arrays(entries, idx, vals, proof, tsgpu.Transcript())
    assert not rc.verify_arrays(entries, idx, forged, proof, tsgpu.Transcript())
    # the drawn point moves with every part of the statement
    pts = set()
    for ent, ix, vv in ((entries, idx, vals), (entries, idx, forged), (entries[::-1].copy(), idx, vals), (entries, idx[::-1].copy(), vals)):
        tr = tsgpu.Transcript()
        tr.append_field_elements(b"read_check_statement", oracle.statement_digest_elements(
            b"shout_read_check", [32, 16], [ent.tobytes(), ix.tobytes(), np.ascontiguousarray(vv).tobytes()]))
        pts.add(tr.challenge_field_elements(b"read_check_point", 4).tobytes())
    assert len(pts) == 4


def test_read_check_bound_to_the_commitments_of_a_shout_proof(ctx, tsgpu, oracle):
    """src/shout.rs:157-184: the read-checking sum-check on a transcript that first absorbed the table / index commitment hashes of the byte-identical
    Shout proof (labels of shout.rs:129-133); the verifier recomputes both commitments from the clear statement."""
    pp, vp = tsgpu.setup_params(ctx, 6)
    table = tsgpu.LookupTable.new(tsgpu.fe_vec([i * i + 7 for i in range(40)]))
    rng = np.random.default_rng(3)
    for i in rng.integers(0, 40, size=33):
        table.lookup(int(i))
    sh = tsgpu.Shout.new(pp)
    proof = sh.prove(table)
    assert sh.verify(proof, vp)
    rc = tsgpu.ShoutReadCheck(ctx)
    claim, cproof, ch = rc.prove(table, rc.bind(tsgpu.Transcript(), proof))
    assert rc.verify(table, cproof, rc.bind(tsgpu.Transcript(), proof))
    assert rc.commitments_match(pp, proof, table)
    idx, vals = rc._statement(table)
    otr = oracle.Transcript()
    otr.append_field_element(b"table_commitment", oracle.g1_hash(proof.commitments[0]))
    otr.append_field_element(b"index_commitment", oracle.g1_hash(proof.commitments[1]))
    want_claim, ref = oracle.shout_read_check_prove(table.entries, idx, vals, "tables", transcript=otr)
    assert (claim == want_claim).all() and (cproof.round_polynomials == ref["round_polynomials"]).all() and (ch == ref["challenges"]).all()
    assert not rc.verify(table, cproof, tsgpu.Transcript())                       # unbound transcript: other challenges
    other = tsgpu.LookupTable.new(tsgpu.fe_vec([i * i + 7 for i in range(39)] + [5]))
    for l in table.lookups:
        other.lookup(l.index)
    proof2 = sh.prove(other)
    assert not rc.verify(table, cproof, rc.bind(tsgpu.Transcript(), proof2))
    assert not rc.commitments_match(pp, proof2, table) and rc.commitments_match(pp, proof2, other)

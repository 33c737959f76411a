"""GPU parity: MultilinearExtension::evaluate / partial_evaluate (src/polynomials.rs:85-161) and the table
generators, through the C ABI, vs the CPU oracle - bit-exact."""
import numpy as np
import pytest

from conftest import seed_bytes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nv", [0, 1, 2, 4, 7, 10, 13])
def test_evaluate_matches_oracle(ctx, oracle, nv):
    n = 1 << nv
    evals = oracle.chacha_fr_rand(seed_bytes(nv + 1), n)
    if n > 4:
        evals[3] = 0   # the reference skips zero entries (polynomials.rs:95-97); value unchanged
    point = oracle.chacha_fr_rand(seed_bytes(100 + nv), nv).reshape(nv, 4)
    ref = oracle.mle_evaluate(evals, point, fold=(nv > 10))
    got = ctx.mle_evaluate(evals, point)
    assert (got == ref).all()


def test_evaluate_large_vs_fold(ctx, oracle):
    nv = 20
    evals = oracle.chacha_fr_rand(seed_bytes(7), 1 << nv)
    point = oracle.chacha_fr_rand(seed_bytes(8), nv)
    ref = oracle.mle_evaluate(evals, point, fold=True)
    t = ctx.table_upload(evals)
    assert (t.evaluate(point) == ref).all()
    # boolean point returns the table entry (polynomials_tests: evaluate at hypercube vertices)
    idx = 0b1011_0011_1010_0101_1100
    bits = oracle.fr_from_ints([(idx >> j) & 1 for j in range(nv)])
    assert (t.evaluate(bits) == evals[idx]).all()


@pytest.mark.parametrize("nv,k", [(2, 1), (4, 0), (4, 4), (6, 3), (9, 2), (12, 5), (12, 11)])
def test_partial_evaluate_matches_oracle(ctx, oracle, nv, k):
    evals = oracle.chacha_fr_rand(seed_bytes(nv * 16 + k), 1 << nv)
    fixed = oracle.chacha_fr_rand(seed_bytes(200 + k), k).reshape(k, 4)
    ref = oracle.mle_partial_evaluate(evals, fixed, fold=(nv > 9))
    got = ctx.mle_partial_evaluate(evals, fixed)
    assert got.shape == ref.shape and (got == ref).all()


def test_reference_anchor_values(ctx, oracle):
    """tests/polynomial_tests.rs:93-131 anchors: [1,2,3,4] at (1/2,1/2) = 10/4; partial_evaluate([1]) -> [2,4]."""
    evals = oracle.fr_from_ints([1, 2, 3, 4])
    half = pow(2, -1, oracle.R_MOD)
    got = ctx.mle_evaluate(evals, oracle.fr_from_ints([half, half]))
    assert oracle.fr_to_ints(got)[0] == 10 * pow(4, -1, oracle.R_MOD) % oracle.R_MOD
    pe = ctx.mle_partial_evaluate(evals, oracle.fr_from_ints([1]))
    assert oracle.fr_to_ints(pe) == [2, 4]


def test_upload_download_roundtrip_and_padding(ctx, oracle):
    evals = oracle.chacha_fr_rand(seed_bytes(5), 37)
    t = ctx.table_upload(evals, num_vars=6)          # from_evaluations_vec pads with zeros (polynomials.rs:40-50)
    back = t.download()
    assert (back[:37] == evals).all() and (back[37:] == 0).all()
    t2 = ctx.table_upload(evals, num_vars=5)         # ... or truncates
    assert (t2.download() == evals[:32]).all()


@pytest.mark.parametrize("nv", [0, 1, 3, 9, 14])
def test_eq_table_generator(ctx, oracle, nv):
    w = oracle.chacha_fr_rand(seed_bytes(33 + nv), nv).reshape(nv, 4)
    got = ctx.table_eq(w).download()
    assert (got == oracle.eq_table(w)).all()


def test_one_hot_and_u64_generators(ctx, oracle):
    rows, log_k = 64, 4
    idx = oracle.chacha_u64(seed_bytes(3), rows) % (1 << log_k)
    t = ctx.table_one_hot_rows(idx, log_k, 10).download()
    ints = oracle.fr_to_ints(t)
    expect = [0] * 1024
    for r in range(rows):
        expect[r * 16 + int(idx[r])] = 1
    assert ints == expect
    v = oracle.chacha_u64(seed_bytes(4), 100)
    t = ctx.table_from_u64(v, 7).download()
    assert (t[:100] == oracle.fr_from_u64(v)).all() and (t[100:] == 0).all()


def test_bind_is_partial_evaluate_of_one_variable(ctx, oracle):
    evals = oracle.chacha_fr_rand(seed_bytes(77), 1 << 10)
    r = oracle.chacha_fr_rand(seed_bytes(78), 1)
    t = ctx.table_upload(evals)
    t.bind(r)
    assert (t.download() == oracle.mle_partial_evaluate(evals, r, fold=True)).all()

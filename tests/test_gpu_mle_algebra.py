"""GPU parity: the MultilinearExtension constructors and algebra (src/polynomials.rs:28-82,164-195: from_sparse, one_hot, add,
scalar_mul, sum_evaluations) and LessThanPolynomial (src/polynomials.rs:201-293), through the C ABI vs the CPU oracle - bit-exact.
The small cases restate the reference's own tests (src/polynomials.rs:415-477, tests/polynomial_tests.rs)."""
import numpy as np
import pytest

from conftest import seed_bytes

pytestmark = pytest.mark.gpu


def test_one_hot_polynomial(ctx, tsgpu, oracle):
    """src/polynomials.rs:415-428: one_hot(3, 5) evaluates to 1 at the bits of 5 and to 0 at every other vertex"""
    mle = tsgpu.MultilinearExtension.one_hot(ctx, 3, 5)
    assert mle.num_vars == 3
    for i in range(8):
        bits = oracle.fr_from_ints([(i >> j) & 1 for j in range(3)])
        assert oracle.fr_to_ints(mle.evaluate(bits)) == [1 if i == 5 else 0]
    ev = oracle.fr_to_ints(mle.evaluations)
    assert ev == [0, 0, 0, 0, 0, 1, 0, 0]
    with pytest.raises(tsgpu.TwistAndShoutError) as e:                       # assert!(index < size), polynomials.rs:73
        tsgpu.MultilinearExtension.one_hot(ctx, 3, 8)
    assert e.value.variant == "Polynomial" and e.value.message == "Index 8 out of bounds for size 8"


@pytest.mark.parametrize("nv,count", [(0, 1), (3, 0), (3, 4), (10, 300), (16, 5000)])
def test_from_sparse_matches_sequential_assignment(ctx, tsgpu, oracle, nv, count):
    rng = np.random.default_rng(nv * 100 + count)
    size = 1 << nv
    idx = rng.integers(0, size, size=count)                                   # repeated indices: the last value wins (polynomials.rs:58-61)
    vals = oracle.chacha_fr_rand(seed_bytes(nv + count), max(count, 1)).reshape(-1, 4)[:count]
    want = np.zeros((size, 4), dtype=np.uint64)
    for i, v in zip(idx, vals):
        want[i] = v
    mle = tsgpu.MultilinearExtension.from_sparse(ctx, nv, [(int(i), v) for i, v in zip(idx, vals)])
    assert mle.num_vars == nv and (mle.evaluations == want).all()
    if count:
        with pytest.raises(tsgpu.TwistAndShoutError) as e:
            tsgpu.MultilinearExtension.from_sparse(ctx, nv, [(size, vals[0])])
        assert e.value.message == f"Index {size} out of bounds for size {size}"


def test_polynomial_operations_reference_case(ctx, tsgpu, oracle):
    """src/polynomials.rs:462-477: [1,2] + [3,4] = [4,6]; 3 * [1,2] = [3,6]"""
    m1 = tsgpu.MultilinearExtension.from_evaluations(ctx, oracle.fr_from_ints([1, 2]))
    m2 = tsgpu.MultilinearExtension.from_evaluations(ctx, oracle.fr_from_ints([3, 4]))
    assert oracle.fr_to_ints(m1.add(m2).evaluations) == [4, 6]
    assert oracle.fr_to_ints(m1.scalar_mul(tsgpu.fe(3)).evaluations) == [3, 6]
    assert oracle.fr_to_ints(m1.sum_evaluations()) == [3]
    m3 = tsgpu.MultilinearExtension.from_evaluations(ctx, oracle.fr_from_ints([1, 2, 3, 4]))
    with pytest.raises(tsgpu.TwistAndShoutError) as e:                        # assert_eq!(num_vars), polynomials.rs:165
        m1.add(m3)
    assert e.value.message == "Number of variables must match"
    with pytest.raises(tsgpu.TwistAndShoutError) as e:                        # polynomials.rs:86-87
        m3.evaluate(oracle.fr_from_ints([1]))
    assert e.value.message == "Point dimension must match number of variables"
    with pytest.raises(ValueError):                                           # polynomials.rs:30-31
        tsgpu.MultilinearExtension.from_evaluations(ctx, oracle.fr_from_ints([1, 2, 3]))


@pytest.mark.parametrize("nv", [0, 1, 5, 9, 14, 20])
def test_add_scalar_mul_sum_match_bigint_arithmetic(ctx, tsgpu, oracle, nv):
    n = 1 << nv
    a = oracle.chacha_fr_rand(seed_bytes(50 + nv), n).reshape(n, 4)
    b = oracle.chacha_fr_rand(seed_bytes(90 + nv), n).reshape(n, 4)
    s = oracle.chacha_fr_rand(seed_bytes(130 + nv), 1).reshape(4)
    A = tsgpu.MultilinearExtension.from_evaluations(ctx, a); B = tsgpu.MultilinearExtension.from_evaluations(ctx, b)
    ai, bi, si = oracle.fr_to_ints(a), oracle.fr_to_ints(b), oracle.fr_to_ints(s)[0]
    p = oracle.R_MOD
    assert (A.add(B).evaluations == oracle.fr_from_ints([(x + y) % p for x, y in zip(ai, bi)])).all()
    assert (A.scalar_mul(s).evaluations == oracle.fr_from_ints([x * si % p for x in ai])).all()
    assert (A.sum_evaluations() == oracle.fr_from_ints([sum(ai) % p])[0]).all()
    # sum_evaluations is what a d = 1 sum-check claims: g(0) + g(1) of round 0
    if nv:
        sc = ctx.sumcheck([A.table.clone()]); ev = sc.round_eval(); sc.end()
        assert oracle.fr_to_ints(A.sum_evaluations())[0] == sum(oracle.fr_to_ints(ev[:2])) % p
    # linearity of the extension: (a + s b)(r) = a(r) + s b(r)
    r = oracle.chacha_fr_rand(seed_bytes(170 + nv), nv).reshape(nv, 4)
    lhs = oracle.fr_to_ints(A.add(B.scalar_mul(s)).evaluate(r))[0]
    assert lhs == (oracle.fr_to_ints(A.evaluate(r))[0] + si * oracle.fr_to_ints(B.evaluate(r))[0]) % p


def test_less_than_polynomial_reference_cases(ctx, tsgpu, oracle):
    """src/polynomials.rs:430-443 and tests/polynomial_tests.rs (lt rows)"""
    lt = tsgpu.LessThanPolynomial.new(3)
    F, T = False, True
    assert oracle.fr_to_ints(lt.evaluate_at_bits([F, F, F], [T, F, F])) == [1]     # 0 < 1
    assert oracle.fr_to_ints(lt.evaluate_at_bits([T, F, F], [F, F, F])) == [0]     # 1 > 0
    assert oracle.fr_to_ints(lt.evaluate_at_bits([T, F, F], [T, F, F])) == [0]     # 1 == 1
    assert oracle.fr_to_ints(lt.evaluate_at_bits([F, T, F], [T, F, F])) == [1]     # first differing bit decides
    assert oracle.fr_to_ints(lt.evaluate_at_field_elements(tsgpu.fe(2), tsgpu.fe(1))) == [1]
    assert oracle.fr_to_ints(lt.evaluate_at_field_elements(tsgpu.fe(1), tsgpu.fe(2))) == [0]
    assert oracle.fr_to_ints(lt.evaluate_at_field_elements(tsgpu.fe(8 + 5), tsgpu.fe(5))) == [0]   # only the low num_vars bits count


@pytest.mark.parametrize("nv", [0, 1, 2, 3, 5, 8])
def test_less_than_table_matches_oracle(ctx, tsgpu, oracle, nv):
    mle = tsgpu.LessThanPolynomial.new(nv).to_multilinear_extension(ctx)
    assert mle.num_vars == 2 * nv
    got = mle.evaluations
    assert (got == oracle.lt_table(nv)).all()
    # every entry agrees with evaluate_at_field_elements on (a, b) = (index & mask, index >> nv)   (polynomials.rs:249-257)
    lt = tsgpu.LessThanPolynomial.new(nv)
    rng = np.random.default_rng(nv)
    for i in rng.integers(0, 1 << (2 * nv), size=min(16, 1 << (2 * nv))):
        a, b = int(i) & ((1 << nv) - 1), int(i) >> nv
        assert (got[int(i)] == lt.evaluate_at_field_elements(tsgpu.fe(a), tsgpu.fe(b))).all()


def test_less_than_table_large_count(ctx, tsgpu, oracle):
    """nv = 11: 2^22 entries; the number of pairs with lt = 1 is (4^nv - 2^nv) / 2 (lt is a strict total order on the bit-reversed values)"""
    nv = 11
    mle = tsgpu.LessThanPolynomial.new(nv).to_multilinear_extension(ctx)
    assert oracle.fr_to_ints(mle.sum_evaluations())[0] == ((1 << (2 * nv)) - (1 << nv)) // 2

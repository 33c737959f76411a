"""The reference's own 72 tests (tests/*.rs and the in-file #[cfg(test)] modules of src/*.rs), restated one for one against this backend's mirror
of the reference API - same names (prefixed with their file), same inputs, same assertions.  What the reference checks with `assert!(x.is_err())`
is checked here as a raised TwistAndShoutError with the reference's variant / message; `panic!` (assert) cases as the mirrored exception.
The closure-typed SumCheck::prove cannot cross an FFI; its two tests use the structured sibling on the same polynomial (x1 * x2 as the product of
the two coordinate MLEs), which tests/test_gpu_sumcheck.py shows to be byte-identical to the closure form."""
import numpy as np
import pytest

# tests that take the `ctx` fixture need the device and carry the gpu marker; the rest (input containers, host helpers) also run in the CPU suite


@pytest.fixture(scope="module")
def ts(tsgpu):
    return tsgpu


def F(ts, x):
    return ts.fe(x)


def eq(a, b):
    return (np.asarray(a, dtype=np.uint64).reshape(-1) == np.asarray(b, dtype=np.uint64).reshape(-1)).all()


def prove_verify_twist(ts, ctx, log_size, trace):
    pp, vp = ts.setup_params(ctx, log_size)
    twist = ts.Twist.new(pp)
    proof = twist.prove(trace)
    return proof, twist.verify(proof, vp), twist, vp


def prove_verify_shout(ts, ctx, log_size, table):
    pp, vp = ts.setup_params(ctx, log_size)
    shout = ts.Shout.new(pp)
    proof = shout.prove(table)
    return proof, shout.verify(proof, vp)


# ================================================================================================ tests/integration_tests.rs
@pytest.mark.gpu
def test_integration__full_memory_consistency_workflow(ts, ctx):
    trace = ts.MemoryTrace.new(8)
    trace.write(0, F(ts, 42)); trace.write(1, F(ts, 100)); trace.write(2, F(ts, 200))
    assert eq(trace.read(0), F(ts, 42)) and eq(trace.read(1), F(ts, 100))
    trace.write(0, F(ts, 43)); trace.write(3, F(ts, 300))
    assert eq(trace.read(0), F(ts, 43)) and eq(trace.read(3), F(ts, 300))
    _, ok, _, _ = prove_verify_twist(ts, ctx, 3, trace)
    assert ok


@pytest.mark.gpu
def test_integration__full_lookup_workflow(ts, ctx):
    table = ts.LookupTable.new(ts.fe_vec([0, 1, 4, 9, 16, 25, 36, 49]))
    for i, sq in ((3, 9), (5, 25), (0, 0), (7, 49)):
        assert eq(table.lookup(i), F(ts, sq))
    _, ok = prove_verify_shout(ts, ctx, 3, table)
    assert ok


@pytest.mark.gpu
def test_integration__commitment_scheme_integration(ts, ctx):
    pp, vp = ts.setup_params(ctx, 3)
    poly = ts.fe_vec([1, 2, 3])
    commitment = ts.KZGCommitment.commit(pp.srs, poly)
    for x in (0, 1, 2, 5):
        value, proof = ts.KZGCommitment.open(pp.srs, poly, F(ts, x))
        assert ts.kzg_verify(vp, commitment, F(ts, x), value, proof)
        assert ts.fe_to_int(value) == 1 + 2 * x + 3 * x * x


@pytest.mark.gpu
def test_integration__combined_twist_and_shout(ts, ctx):
    pp, vp = ts.setup_params(ctx, 3)
    opcodes = ts.LookupTable.new(ts.fe_vec(list(range(8))))
    memory = ts.MemoryTrace.new(8)
    opcodes.lookup(1); memory.write(0, F(ts, 42))
    opcodes.lookup(1); memory.write(1, F(ts, 58))
    opcodes.lookup(3)
    a, b = memory.read(0), memory.read(1)
    memory.write(2, ts.fe_add(a, b))
    assert eq(memory.read(2), F(ts, 100))
    opcodes.lookup(7)
    twist, shout = ts.Twist.new(pp), ts.Shout.new(pp)
    assert twist.verify(twist.prove(memory), vp)
    assert shout.verify(shout.prove(opcodes), vp)


@pytest.mark.gpu
def test_integration__polynomial_commitment_consistency(ts, ctx):
    pp, vp = ts.setup_params(ctx, 4)
    mle = ts.MultilinearExtension.from_evaluations(ctx, ts.fe_vec([10, 20, 30, 40, 50, 60, 70, 80]))
    fixed = ts.fe_vec([2, 3])
    partial = mle.partial_evaluate(fixed)
    ev = partial.evaluations
    coeffs = ts.poly_utils.lagrange_interpolate(ctx, [(F(ts, i), ev[i]) for i in range(ev.shape[0])])
    commitment = ts.KZGCommitment.commit(pp.srs, coeffs)
    value, proof = ts.KZGCommitment.open(pp.srs, coeffs, F(ts, 10))
    assert ts.kzg_verify(vp, commitment, F(ts, 10), value, proof)
    mle.evaluate(np.stack([fixed[0], fixed[1], F(ts, 10)]))


@pytest.mark.gpu
def test_integration__parameter_compatibility(ts, ctx):
    pp, vp = ts.setup_params(ctx, 4)
    assert pp.log_size == vp.log_size and pp.max_operations == vp.max_operations and pp.fiat_shamir_seed == vp.fiat_shamir_seed
    poly = ts.fe_vec([1, 2])
    commitment = ts.KZGCommitment.commit(pp.srs, poly)
    value, proof = ts.KZGCommitment.open(pp.srs, poly, F(ts, 5))
    assert ts.kzg_verify(vp, commitment, F(ts, 5), value, proof)


def _sumcheck_x1_times_x2(ts, ctx):
    # f(x1, x2) = x1 * x2 = (MLE of [0, 1, 0, 1]) * (MLE of [0, 0, 1, 1]); sum over {0,1}^2 = 1
    sumcheck = ts.SumCheck(2, F(ts, 1))
    tables = [ctx.table_upload(ts.fe_vec([0, 1, 0, 1])), ctx.table_upload(ts.fe_vec([0, 0, 1, 1]))]
    proof = sumcheck.prove_product(ctx, tables, ts.Transcript(bytes([42]) * 32))
    ok, _ = sumcheck.verify(proof, ts.Transcript(bytes([42]) * 32))
    return ok


@pytest.mark.gpu
def test_integration__sumcheck_protocol_basic(ts, ctx):
    assert _sumcheck_x1_times_x2(ts, ctx)


@pytest.mark.gpu
def test_integration__error_handling(ts, ctx):
    pp, _ = ts.setup_params(ctx, 2)
    large = ts.MemoryTrace.new(4)
    for i in range(100):
        large.write(i % 4, F(ts, i))
    with pytest.raises(ts.TwistAndShoutError) as e:
        ts.Twist.new(pp).prove(large)
    assert e.value.variant == "InvalidParameters" and e.value.message == "Too many operations"
    table = ts.LookupTable.new(ts.fe_vec([1] * 4))
    for _ in range(100):
        table.lookup(0)
    with pytest.raises(ts.TwistAndShoutError) as e:
        ts.Shout.new(pp).prove(table)
    assert e.value.message == "Too many lookup operations"
    trace = ts.MemoryTrace.new(4)
    for call in (lambda: trace.write(4, F(ts, 1)), lambda: trace.read(10), lambda: ts.LookupTable.new(ts.fe_vec([1, 1])).lookup(2)):
        with pytest.raises(ts.TwistAndShoutError):
            call()


# ================================================================================================ tests/polynomial_tests.rs
@pytest.mark.gpu
def test_polynomial__multilinear_extension_creation(ts, ctx):
    ev = ts.fe_vec([1, 2, 3, 4])
    mle = ts.MultilinearExtension.from_evaluations(ctx, ev)
    assert mle.num_vars == 2 and eq(mle.evaluations, ev)


@pytest.mark.gpu
def test_polynomial__multilinear_extension_power_of_two_requirement(ts, ctx):
    assert ts.MultilinearExtension.from_evaluations(ctx, ts.fe_vec([1] * 8)).num_vars == 3
    with pytest.raises(ValueError):                                                # panics in the reference (assert_eq!)
        ts.MultilinearExtension.from_evaluations(ctx, ts.fe_vec([1] * 7))


@pytest.mark.gpu
def test_polynomial__multilinear_extension_from_sparse(ts, ctx):
    mle = ts.MultilinearExtension.from_sparse(ctx, 3, [(0, F(ts, 10)), (2, F(ts, 30)), (5, F(ts, 60))])
    assert mle.num_vars == 3
    assert [ts.fe_to_int(v) for v in mle.evaluations] == [10, 0, 30, 0, 0, 60, 0, 0]


@pytest.mark.gpu
def test_polynomial__one_hot_polynomial(ts, ctx):
    mle = ts.MultilinearExtension.one_hot(ctx, 3, 5)
    assert mle.num_vars == 3 and [ts.fe_to_int(v) for v in mle.evaluations] == [1 if i == 5 else 0 for i in range(8)]


@pytest.mark.gpu
def test_polynomial__multilinear_extension_evaluation_at_boolean_points(ts, ctx):
    mle = ts.MultilinearExtension.from_evaluations(ctx, ts.fe_vec([1, 2, 3, 4]))
    for pt, want in (((0, 0), 1), ((1, 0), 2), ((0, 1), 3), ((1, 1), 4)):
        assert eq(mle.evaluate(ts.fe_vec(list(pt))), F(ts, want))


@pytest.mark.gpu
def test_polynomial__multilinear_extension_evaluation_at_random_points(ts, ctx):
    mle = ts.MultilinearExtension.from_evaluations(ctx, ts.fe_vec([1, 2, 3, 4]))
    half = ts.fe_inverse(F(ts, 2))
    expected = ts.fe_mul(F(ts, 10), ts.fe_inverse(F(ts, 4)))
    assert eq(mle.evaluate(np.stack([half, half])), expected)


@pytest.mark.gpu
def test_polynomial__partial_evaluation(ts, ctx):
    mle = ts.MultilinearExtension.from_evaluations(ctx, ts.fe_vec([1, 2, 3, 4]))
    partial = mle.partial_evaluate(ts.fe_vec([1]))
    assert partial.num_vars == 1
    assert eq(partial.evaluate(ts.fe_vec([0])), F(ts, 2)) and eq(partial.evaluate(ts.fe_vec([1])), F(ts, 4))


@pytest.mark.gpu
def test_polynomial__polynomial_arithmetic(ts, ctx):
    m1 = ts.MultilinearExtension.from_evaluations(ctx, ts.fe_vec([1, 2]))
    m2 = ts.MultilinearExtension.from_evaluations(ctx, ts.fe_vec([3, 4]))
    assert eq(m1.add(m2).evaluations, ts.fe_vec([4, 6]))
    assert eq(m1.scalar_mul(F(ts, 3)).evaluations, ts.fe_vec([3, 6]))
    assert eq(m1.sum_evaluations(), F(ts, 3))


def test_polynomial__less_than_polynomial(ts):
    lt = ts.LessThanPolynomial.new(3)
    f, t = False, True
    assert eq(lt.evaluate_at_bits([f, f, f], [t, f, f]), F(ts, 1))
    assert eq(lt.evaluate_at_bits([t, f, f], [f, f, f]), F(ts, 0))
    assert eq(lt.evaluate_at_bits([t, f, f], [t, f, f]), F(ts, 0))
    assert eq(lt.evaluate_at_bits([f, t, f], [t, t, f]), F(ts, 1))
    assert eq(lt.evaluate_at_bits([t, t, f], [f, t, f]), F(ts, 0))


@pytest.mark.gpu
def test_polynomial__less_than_polynomial_multilinear_extension(ts, ctx):
    mle = ts.LessThanPolynomial.new(2).to_multilinear_extension(ctx)
    assert mle.num_vars == 4
    assert eq(mle.evaluate(ts.fe_vec([0, 0, 0, 1])), F(ts, 1))
    assert eq(mle.evaluate(ts.fe_vec([0, 1, 0, 0])), F(ts, 0))


@pytest.mark.gpu
def test_polynomial__lagrange_interpolation(ts, ctx):
    coeffs = ts.poly_utils.lagrange_interpolate(ctx, [(F(ts, 0), F(ts, 0)), (F(ts, 1), F(ts, 1)), (F(ts, 2), F(ts, 4))])
    assert [ts.fe_to_int(c) for c in coeffs] == [0, 0, 1]


def test_polynomial__polynomial_evaluation(ts):
    assert eq(ts.poly_utils.evaluate_polynomial(ts.fe_vec([1, 2, 3]), F(ts, 5)), F(ts, 86))


def test_polynomial__polynomial_derivative(ts):
    assert [ts.fe_to_int(c) for c in ts.poly_utils.derivative(ts.fe_vec([5, 1, 2, 3]))] == [1, 4, 9]


@pytest.mark.gpu
def test_polynomial__sparse_multilinear_extension(ts, ctx):
    mle = ts.MultilinearExtension.from_sparse(ctx, 3, [(0, F(ts, 100)), (7, F(ts, 700))])
    assert eq(mle.evaluate(ts.fe_vec([0, 0, 0])), F(ts, 100))
    assert eq(mle.evaluate(ts.fe_vec([1, 1, 1])), F(ts, 700))
    assert eq(mle.evaluate(ts.fe_vec([1, 1, 0])), F(ts, 0))


@pytest.mark.gpu
def test_polynomial__multilinear_extension_random_evaluation(ts, ctx, oracle):
    ev = oracle.chacha_fr_rand(bytes(32), 16).reshape(16, 4)          # 16 x FieldElement::rand of a fixed-seed ChaCha20 (the reference uses ark_std::test_rng)
    mle = ts.MultilinearExtension.from_evaluations(ctx, ev)
    for i in range(16):
        assert eq(mle.evaluate(ts.fe_vec([(i >> j) & 1 for j in range(4)])), ev[i])


@pytest.mark.gpu
def test_polynomial__multilinear_extension_properties(ts, ctx):
    m1 = ts.MultilinearExtension.from_evaluations(ctx, ts.fe_vec([1, 2, 3, 4]))
    m2 = ts.MultilinearExtension.from_evaluations(ctx, ts.fe_vec([5, 6, 7, 8]))
    pt = ts.fe_vec([3, 7])
    assert eq(m1.add(m2).evaluate(pt), ts.fe_add(m1.evaluate(pt), m2.evaluate(pt)))
    assert eq(m1.scalar_mul(F(ts, 5)).evaluate(pt), ts.fe_mul(m1.evaluate(pt), F(ts, 5)))


# ================================================================================================ tests/production_tests.rs
@pytest.mark.gpu
def test_production__twist_with_opening_proofs(ts, ctx):
    trace = ts.MemoryTrace.new(16)
    trace.write(0, F(ts, 42)); trace.write(1, F(ts, 73)); trace.write(2, F(ts, 100))
    assert eq(trace.read(0), F(ts, 42)) and eq(trace.read(1), F(ts, 73))
    trace.write(0, F(ts, 999)); trace.write(1, F(ts, 888))
    assert eq(trace.read(0), F(ts, 999)) and eq(trace.read(1), F(ts, 888))
    proof, ok, _, _ = prove_verify_twist(ts, ctx, 4, trace)
    assert ok
    assert proof.round_polynomials.shape[0] > 0


@pytest.mark.gpu
def test_production__shout_with_opening_proofs(ts, ctx):
    table = ts.LookupTable.new(ts.fe_vec([10, 20, 30, 40, 50]))
    for i in (0, 2, 4, 1, 3):
        table.lookup(i)
    proof, ok = prove_verify_shout(ts, ctx, 4, table)
    assert ok
    assert proof.round_polynomials.shape[0] > 0


@pytest.mark.gpu
def test_production__twist_with_multilinear_extensions(ts, ctx):
    trace = ts.MemoryTrace.new(8)
    for i in range(8):
        trace.write(i, F(ts, i * i + 1))
    for i in reversed(range(8)):
        trace.read(i)
    proof, ok, _, _ = prove_verify_twist(ts, ctx, 3, trace)
    assert ok
    assert proof.round_polynomials.shape[0] == 4                                   # 16 operations -> log_ops = 4


@pytest.mark.gpu
def test_production__shout_edge_cases(ts, ctx):
    pp, vp = ts.setup_params(ctx, 2)
    shout = ts.Shout.new(pp)
    small = ts.LookupTable.new(ts.fe_vec([123])); small.lookup(0)
    assert shout.verify(shout.prove(small), vp)
    rep = ts.LookupTable.new(ts.fe_vec([456, 789]))
    for i in (0, 0, 1, 0):
        rep.lookup(i)
    assert shout.verify(shout.prove(rep), vp)


@pytest.mark.gpu
def test_production__proof_non_malleability(ts, ctx):
    trace = ts.MemoryTrace.new(8)
    trace.write(0, F(ts, 42)); trace.write(1, F(ts, 73))
    proof, ok, twist, vp = prove_verify_twist(ts, ctx, 3, trace)
    assert ok
    assert proof.final_evaluations.shape[0] == 2
    proof.tamper_final_evaluation(0, F(ts, 999))
    # the reference leaves the outcome open ("our simplified version may pass"); with the pairing check in place the tampered proof is rejected
    assert twist.verify(proof, vp) is False


# ================================================================================================ tests/shout_tests.rs
def test_shout__lookup_table_basic_operations(ts):
    table = ts.LookupTable.new(ts.fe_vec([10, 20, 30, 40, 50]))
    assert eq(table.lookup(0), F(ts, 10)) and eq(table.lookup(2), F(ts, 30)) and eq(table.lookup(4), F(ts, 50))
    assert len(table.lookups) == 3 and table.size() == 5


def test_shout__lookup_table_bounds_checking(ts):
    table = ts.LookupTable.new(ts.fe_vec([100, 200, 300]))
    for i in (0, 1, 2):
        table.lookup(i)
    for i in (3, 100):
        with pytest.raises(ts.TwistAndShoutError) as e:
            table.lookup(i)
        assert e.value.variant == "InvalidParameters" and e.value.message == "Lookup index out of bounds"


def test_shout__lookup_table_empty(ts):
    table = ts.LookupTable.new([])
    assert table.size() == 0
    with pytest.raises(ts.TwistAndShoutError):
        table.lookup(0)


def test_shout__lookup_table_single_entry(ts):
    table = ts.LookupTable.new(ts.fe_vec([42]))
    assert table.size() == 1 and eq(table.lookup(0), F(ts, 42))
    with pytest.raises(ts.TwistAndShoutError):
        table.lookup(1)


def _shout_case(ts, ctx, log_size, entries, lookups):
    table = ts.LookupTable.new(ts.fe_vec(entries))
    for i in lookups:
        table.lookup(i)
    return prove_verify_shout(ts, ctx, log_size, table)[1]


@pytest.mark.gpu
def test_shout__protocol_basic_lookup(ts, ctx):
    assert _shout_case(ts, ctx, 3, [100, 200, 300, 400], [0, 2, 3, 1])


@pytest.mark.gpu
def test_shout__protocol_no_lookups(ts, ctx):
    assert _shout_case(ts, ctx, 2, [10, 20, 30, 40], [])


@pytest.mark.gpu
def test_shout__protocol_single_lookup(ts, ctx):
    assert _shout_case(ts, ctx, 2, [1000, 2000], [1])


@pytest.mark.gpu
def test_shout__protocol_repeated_lookups(ts, ctx):
    assert _shout_case(ts, ctx, 2, [111, 222, 333], [0, 0, 1, 0, 2, 1])


@pytest.mark.gpu
def test_shout__protocol_all_indices(ts, ctx):
    assert _shout_case(ts, ctx, 2, [10, 20, 30, 40], range(4))


@pytest.mark.gpu
def test_shout__protocol_reverse_order(ts, ctx):
    assert _shout_case(ts, ctx, 2, [100, 200, 300, 400], reversed(range(4)))


@pytest.mark.gpu
def test_shout__protocol_large_table(ts, ctx):
    assert _shout_case(ts, ctx, 4, [i * 10 for i in range(16)], [0, 5, 10, 15, 2, 8, 1, 14])


@pytest.mark.gpu
def test_shout__protocol_exceeds_operations_limit(ts, ctx):
    with pytest.raises(ts.TwistAndShoutError) as e:
        _shout_case(ts, ctx, 1, [1, 2], [0] * 20)
    assert e.value.message == "Too many lookup operations"


def test_shout__lookup_op_structure(ts):
    op = ts.LookupOp(5, F(ts, 42))
    assert op.index == 5 and eq(op.value, F(ts, 42))
    op2 = op
    assert op.index == op2.index and eq(op.value, op2.value)


@pytest.mark.gpu
def test_shout__protocol_zero_values(ts, ctx):
    assert _shout_case(ts, ctx, 2, [0, 100, 0, 200], [0, 1, 2, 3])


@pytest.mark.gpu
def test_shout__protocol_duplicate_values(ts, ctx):
    assert _shout_case(ts, ctx, 2, [100, 200, 100, 300], [0, 2, 1])
# ================================================================================================ tests/twist_tests.rs
def test_twist__memory_trace_basic_operations(ts):
    trace = ts.MemoryTrace.new(16)
    trace.write(0, F(ts, 42)); trace.write(5, F(ts, 100)); trace.write(15, F(ts, 255))
    assert eq(trace.read(0), F(ts, 42)) and eq(trace.read(5), F(ts, 100)) and eq(trace.read(15), F(ts, 255))
    assert eq(trace.read(10), F(ts, 0))
    assert len(trace.operations) == 7


def test_twist__memory_trace_write_then_read(ts):
    trace = ts.MemoryTrace.new(8)
    trace.write(3, F(ts, 123)); assert eq(trace.read(3), F(ts, 123))
    trace.write(3, F(ts, 456)); assert eq(trace.read(3), F(ts, 456))


def test_twist__memory_trace_bounds_checking(ts):
    trace = ts.MemoryTrace.new(4)
    trace.write(0, F(ts, 1)); trace.write(3, F(ts, 2)); trace.read(0); trace.read(3)
    for call in (lambda: trace.write(4, F(ts, 1)), lambda: trace.write(100, F(ts, 1)), lambda: trace.read(4), lambda: trace.read(100)):
        with pytest.raises(ts.TwistAndShoutError) as e:
            call()
        assert e.value.variant == "InvalidParameters" and e.value.message == "Address out of bounds"


def _twist_case(ts, ctx, log_size, cells, ops):
    trace = ts.MemoryTrace.new(cells)
    for op in ops:
        trace.write(op[1], F(ts, op[2])) if op[0] == "W" else trace.read(op[1])
    return prove_verify_twist(ts, ctx, log_size, trace)[1]


@pytest.mark.gpu
def test_twist__protocol_small_trace(ts, ctx):
    assert _twist_case(ts, ctx, 3, 8, [("W", 0, 10), ("W", 1, 20), ("R", 0), ("W", 2, 30), ("R", 1), ("R", 2)])


@pytest.mark.gpu
def test_twist__protocol_empty_trace(ts, ctx):
    assert _twist_case(ts, ctx, 2, 4, [])


@pytest.mark.gpu
def test_twist__protocol_only_reads(ts, ctx):
    assert _twist_case(ts, ctx, 2, 4, [("R", i) for i in range(4)])


@pytest.mark.gpu
def test_twist__protocol_only_writes(ts, ctx):
    assert _twist_case(ts, ctx, 2, 4, [("W", i, i + 1) for i in range(4)])


@pytest.mark.gpu
def test_twist__protocol_repeated_operations(ts, ctx):
    assert _twist_case(ts, ctx, 2, 4, [("W", 0, 100), ("R", 0), ("W", 0, 200), ("R", 0), ("W", 0, 300), ("R", 0)])


@pytest.mark.gpu
def test_twist__protocol_max_operations(ts, ctx):
    assert _twist_case(ts, ctx, 2, 4, [("W", i % 4, i + 1) for i in range(15)])


@pytest.mark.gpu
def test_twist__protocol_exceeds_operations_limit(ts, ctx):
    with pytest.raises(ts.TwistAndShoutError) as e:
        _twist_case(ts, ctx, 1, 2, [("W", i % 2, i + 1) for i in range(10)])
    assert e.value.message == "Too many operations"


def test_twist__memory_operation_types(ts):
    read_op = ts.MemoryOp("R", 5, F(ts, 42)); write_op = ts.MemoryOp("W", 10, F(ts, 100))
    assert read_op.kind == "R" and read_op.address == 5 and eq(read_op.value, F(ts, 42))
    assert write_op.kind == "W" and write_op.address == 10 and eq(write_op.value, F(ts, 100))
    assert read_op.kind != write_op.kind


# ================================================================================================ src/commitments.rs (in-file tests)
@pytest.mark.gpu
def test_commitments__kzg_commitment(ts, ctx):
    pp, vp = ts.setup_params(ctx, 4)
    poly = ts.fe_vec([1, 2, 3])
    commitment = ts.KZGCommitment.commit(pp.srs, poly)
    value, proof = ts.KZGCommitment.open(pp.srs, poly, F(ts, 5))
    assert eq(value, F(ts, 86))
    assert ts.kzg_verify(vp, commitment, F(ts, 5), value, proof)
    assert not ts.kzg_verify(vp, commitment, F(ts, 5), F(ts, 87), proof)


@pytest.mark.gpu
def test_commitments__kzg_vector_commitment(ts, ctx):
    pp, vp = ts.setup_params(ctx, 4)
    vector = ts.fe_vec([10, 20, 30, 40])
    commitment = ts.KZGVectorCommitment.commit(pp.srs, vector)
    value, proof = ts.KZGVectorCommitment.open(pp.srs, vector, 2)
    assert eq(value, F(ts, 30))
    assert ts.KZGVectorCommitment.verify(vp, commitment, 2, value, proof)


def test_commitments__polynomial_division(ts):
    q = ts.polynomial_division([ts.fe_from_int(-1), F(ts, 0), F(ts, 1)], [ts.fe_from_int(-1), F(ts, 1)])
    assert [ts.fe_to_int(c) for c in q] == [1, 1]


# ================================================================================================ src/lib.rs, src/polynomials.rs, src/shout.rs, src/sumcheck.rs, src/twist.rs, src/utils.rs
@pytest.mark.gpu
def test_lib__library_imports(ts, ctx):
    ts.setup_params(ctx, 4)


@pytest.mark.gpu
def test_polynomials_rs__multilinear_extension_evaluation(ts, ctx, oracle):
    mle = ts.MultilinearExtension.from_evaluations(ctx, ts.fe_vec([1, 2, 3, 4]))
    for pt, want in (((0, 0), 1), ((1, 0), 2), ((0, 1), 3), ((1, 1), 4)):
        assert eq(mle.evaluate(ts.fe_vec(list(pt))), F(ts, want))
    r = oracle.chacha_fr_rand(bytes(32), 2).reshape(2, 4)
    r1, r2 = ts.fe_to_int(r[0]), ts.fe_to_int(r[1])
    expected = (1 * (1 - r1) * (1 - r2) + 2 * r1 * (1 - r2) + 3 * (1 - r1) * r2 + 4 * r1 * r2) % oracle.R_MOD
    assert ts.fe_to_int(mle.evaluate(r)) == expected


@pytest.mark.gpu
def test_polynomials_rs__one_hot_polynomial(ts, ctx):
    mle = ts.MultilinearExtension.one_hot(ctx, 3, 5)
    for i in range(8):
        assert eq(mle.evaluate(ts.fe_vec([(i >> j) & 1 for j in range(3)])), F(ts, 1 if i == 5 else 0))


def test_polynomials_rs__less_than_polynomial(ts):
    lt = ts.LessThanPolynomial.new(3)
    f, t = False, True
    assert eq(lt.evaluate_at_bits([f, f, f], [t, f, f]), F(ts, 1))
    assert eq(lt.evaluate_at_bits([t, f, f], [f, f, f]), F(ts, 0))
    assert eq(lt.evaluate_at_bits([t, f, f], [t, f, f]), F(ts, 0))
    assert eq(lt.evaluate_at_bits([f, t, f], [t, f, f]), F(ts, 1))       # the first differing bit (bit 0) decides


@pytest.mark.gpu
def test_polynomials_rs__partial_evaluation(ts, ctx):
    partial = ts.MultilinearExtension.from_evaluations(ctx, ts.fe_vec([1, 2, 3, 4])).partial_evaluate(ts.fe_vec([1]))
    assert partial.num_vars == 1
    assert eq(partial.evaluate(ts.fe_vec([0])), F(ts, 2)) and eq(partial.evaluate(ts.fe_vec([1])), F(ts, 4))


@pytest.mark.gpu
def test_polynomials_rs__polynomial_operations(ts, ctx):
    m1 = ts.MultilinearExtension.from_evaluations(ctx, ts.fe_vec([1, 2]))
    m2 = ts.MultilinearExtension.from_evaluations(ctx, ts.fe_vec([3, 4]))
    assert eq(m1.add(m2).evaluations, ts.fe_vec([4, 6])) and eq(m1.scalar_mul(F(ts, 3)).evaluations, ts.fe_vec([3, 6]))


def test_shout_rs__lookup_table(ts):
    table = ts.LookupTable.new(ts.fe_vec([10, 20, 30, 40]))
    assert eq(table.lookup(0), F(ts, 10)) and eq(table.lookup(2), F(ts, 30)) and len(table.lookups) == 2


@pytest.mark.gpu
def test_shout_rs__shout_prove_verify(ts, ctx):
    assert _shout_case(ts, ctx, 4, [100, 200, 300, 400], [0, 2, 1])


@pytest.mark.gpu
def test_sumcheck_rs__sumcheck_simple(ts, ctx):
    ts.setup_params(ctx, 2)
    assert _sumcheck_x1_times_x2(ts, ctx)


def test_twist_rs__memory_trace(ts):
    trace = ts.MemoryTrace.new(8)
    trace.write(0, F(ts, 42)); trace.write(1, F(ts, 73))
    assert eq(trace.read(0), F(ts, 42)) and eq(trace.read(1), F(ts, 73)) and len(trace.operations) == 4


@pytest.mark.gpu
def test_twist_rs__twist_prove_verify(ts, ctx):
    assert _twist_case(ts, ctx, 4, 16, [("W", 0, 42), ("W", 1, 73), ("R", 0)])


@pytest.mark.gpu
def test_utils_rs__setup_params(ts, ctx):
    pp, vp = ts.setup_params(ctx, 4)
    assert pp.log_size == 4 and vp.log_size == 4 and pp.max_operations == 64
    assert len(pp.srs) > 0


def test_utils_rs__transcript(ts):
    transcript = ts.Transcript(bytes([42]) * 32)
    transcript.append_field_element(b"test", F(ts, 123))
    assert ts.fe_to_int(transcript.challenge_field_element(b"challenge")) != 0


def test_utils_rs__field_utils(ts):
    assert eq(ts.field_utils.inner_product(ts.fe_vec([1, 2]), ts.fe_vec([3, 4])), F(ts, 11))
    assert eq(ts.field_utils.powers(F(ts, 2), 4), ts.fe_vec([1, 2, 4, 8]))
    els = ts.fe_vec([2, 3, 5])
    for e, inv in zip(els, ts.field_utils.batch_inverse(els)):
        assert eq(ts.fe_mul(e, inv), F(ts, 1))

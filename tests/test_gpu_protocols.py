"""GPU parity for the whole path: setup_params, Twist::prove/verify, Shout::prove/verify through the C ABI.
Mirrors the reference's tests/twist_tests.rs, tests/shout_tests.rs, tests/integration_tests.rs (round trips,
limits, empty inputs) and adds what the reference never checks: canonical proof bytes identical to the golden
vectors and to the CPU oracle."""
import json
import os

import numpy as np
import pytest

from conftest import seed_bytes

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "appendix_c.json")))


@pytest.fixture(scope="module")
def params3(ctx, tsgpu):
    return tsgpu.setup_params(ctx, 3)


@pytest.fixture(scope="module")
def params10(ctx, tsgpu):
    return tsgpu.setup_params(ctx, 10)


def test_setup_params_matches_reference(ctx, tsgpu, oracle, params3):
    """src/utils.rs:277-284 (max_operations == 4 * 2^log) + Appendix C.1 values"""
    pp, vp = params3
    assert pp.log_size == 3 and vp.log_size == 3 and pp.max_operations == 32
    assert str(oracle.fr_to_ints(pp.tau)[0]) == GOLD["tau"] and pp.fiat_shamir_seed.hex() == GOLD["fiat_shamir_seed"]
    assert len(pp.srs) == 33
    pw = pp.srs.download()
    assert oracle.g1_compress(pw[1]).hex() == GOLD["g1_powers_1"] and oracle.g1_compress(pw[32]).hex() == GOLD["g1_powers_32"]
    pp4, _ = tsgpu.setup_params(ctx, 4)
    assert pp4.max_operations == 64


def test_demo_twist_golden_bytes(ctx, tsgpu, params3):
    """examples/demo.rs:33-61 trace; bytes pinned in tests/golden/appendix_c.json"""
    pp, vp = params3
    trace = tsgpu.MemoryTrace.new(8)
    trace.write(0, tsgpu.fe(42)); trace.write(1, tsgpu.fe(100))
    assert tsgpu.fe_to_int(trace.read(0)) == 42 and tsgpu.fe_to_int(trace.read(1)) == 100
    trace.write(0, tsgpu.fe(43))
    assert tsgpu.fe_to_int(trace.read(0)) == 43
    twist = tsgpu.Twist.new(pp)
    proof = twist.prove(trace)
    assert proof.to_bytes().hex() == GOLD["twist_demo"]["proof_hex"]
    assert twist.verify(proof, vp)
    assert len(proof.round_polynomials) == 3 and (proof.round_polynomials == 0).all() and (proof.final_evaluation == 0).all()


def test_demo_shout_golden_bytes(ctx, tsgpu, params3):
    """examples/demo.rs:66-93"""
    pp, vp = params3
    table = tsgpu.LookupTable.new(tsgpu.fe_vec([i * i for i in range(8)]))
    for i in (3, 5, 0, 7):
        assert tsgpu.fe_to_int(table.lookup(i)) == i * i
    shout = tsgpu.Shout.new(pp)
    proof = shout.prove(table)
    assert proof.to_bytes().hex() == GOLD["shout_demo"]["proof_hex"]
    assert shout.verify(proof, vp)


def test_readme_quick_start_golden_bytes(ctx, tsgpu):
    """README.md:40-60 with setup_params(8)"""
    pp, vp = tsgpu.setup_params(ctx, 8)
    trace = tsgpu.MemoryTrace.new(256)
    trace.write(0, tsgpu.fe(42)); trace.write(1, tsgpu.fe(100)); trace.read(0)
    proof = tsgpu.Twist.new(pp).prove(trace)
    assert proof.to_bytes().hex() == GOLD["twist_readme"]["proof_hex"]
    assert tsgpu.Twist.new(pp).verify(proof, vp)
    table = tsgpu.LookupTable.new(tsgpu.fe_vec([1, 4, 9]))
    table.lookup(1)
    sproof = tsgpu.Shout.new(pp).prove(table)
    assert sproof.to_bytes().hex() == GOLD["shout_readme"]["proof_hex"]
    assert len(sproof.opening_proofs) == 0 and len(sproof.round_polynomials) == 0      # one lookup: no rounds, no openings
    assert tsgpu.Shout.new(pp).verify(sproof, vp)


def test_empty_trace_and_no_lookups(ctx, tsgpu, params3):
    """tests/twist_tests.rs:88-99, tests/shout_tests.rs:100-118"""
    pp, vp = params3
    proof = tsgpu.Twist.new(pp).prove(tsgpu.MemoryTrace.new(8))
    assert proof.to_bytes().hex() == GOLD["twist_empty"]["proof_hex"]
    assert tsgpu.Twist.new(pp).verify(proof, vp) and len(proof.final_evaluations) == 0
    sproof = tsgpu.Shout.new(pp).prove(tsgpu.LookupTable.new(tsgpu.fe_vec([1, 2, 3, 4])))
    assert sproof.to_bytes().hex() == GOLD["shout_no_lookups"]["proof_hex"]
    assert tsgpu.Shout.new(pp).verify(sproof, vp)


def test_operation_limits(ctx, tsgpu):
    """tests/twist_tests.rs:180-196 / tests/shout_tests.rs:243-263: more than max_operations is InvalidParameters"""
    pp, _ = tsgpu.setup_params(ctx, 1)            # max_operations = 8
    trace = tsgpu.MemoryTrace.new(2)
    for i in range(9):
        trace.write(i % 2, tsgpu.fe(i))
    with pytest.raises(tsgpu.TwistAndShoutError) as e:
        tsgpu.Twist.new(pp).prove(trace)
    assert e.value.variant == "InvalidParameters" and "Too many operations" in str(e.value)
    table = tsgpu.LookupTable.new(tsgpu.fe_vec([1, 2]))
    for i in range(9):
        table.lookup(i % 2)
    with pytest.raises(tsgpu.TwistAndShoutError) as e:
        tsgpu.Shout.new(pp).prove(table)
    assert e.value.variant == "InvalidParameters" and "Too many lookup operations" in str(e.value)
    trace8 = tsgpu.MemoryTrace.new(2)
    for i in range(8):
        trace8.write(i % 2, tsgpu.fe(i))
    assert tsgpu.Twist.new(pp).verify(tsgpu.Twist.new(pp).prove(trace8), pp)      # exactly at the limit is fine


def test_bounds_errors(tsgpu):
    """src/twist.rs:49-53,62-66; src/shout.rs:44-48"""
    trace = tsgpu.MemoryTrace.new(4)
    with pytest.raises(tsgpu.TwistAndShoutError) as e:
        trace.write(4, tsgpu.fe(1))
    assert "Address out of bounds" in str(e.value)
    with pytest.raises(tsgpu.TwistAndShoutError):
        trace.read(100)
    with pytest.raises(AssertionError):
        tsgpu.MemoryTrace.new(6)
    table = tsgpu.LookupTable.new(tsgpu.fe_vec([1, 2, 3]))
    with pytest.raises(tsgpu.TwistAndShoutError) as e:
        table.lookup(3)
    assert "Lookup index out of bounds" in str(e.value)


@pytest.mark.parametrize("nops", [1, 2, 5, 16, 33, 200, 1000, 4096])
def test_twist_random_traces_match_oracle_bytes(ctx, tsgpu, oracle, params10, nops):
    pp, vp = params10
    pw = pp.srs.download()
    rng = np.random.default_rng(nops)
    addr = rng.integers(0, 1 << 10, size=nops).astype(np.uint64)
    vals = oracle.chacha_fr_rand(seed_bytes(nops), nops).reshape(nops, 4)
    isw = rng.integers(0, 2, size=nops).astype(np.uint8)
    proof = tsgpu.Twist.new(pp).prove_arrays(addr, vals, isw)
    want, z = oracle.twist_prove(pw, pp.max_operations, addr, vals, isw, fast=True)
    assert proof.to_bytes() == want
    if nops > 1:
        assert (proof.opening_point == z).all()
    assert tsgpu.Twist.new(pp).verify(proof, vp)
    if nops <= 33:      # the reference's own O(n^3) algorithms, restated verbatim
        want_verbatim, _ = oracle.twist_prove(pw, pp.max_operations, addr, vals, isw, fast=False)
        assert proof.to_bytes() == want_verbatim


@pytest.mark.parametrize("nent,nlook", [(1, 1), (3, 2), (8, 4), (100, 37), (1000, 4096), (4096, 700)])
def test_shout_random_tables_match_oracle_bytes(ctx, tsgpu, oracle, params10, nent, nlook):
    pp, vp = params10
    pw = pp.srs.download()
    rng = np.random.default_rng(nent * 7 + nlook)
    entries = oracle.chacha_fr_rand(seed_bytes(nent), nent).reshape(nent, 4)
    idx = rng.integers(0, nent, size=nlook).astype(np.uint64)
    proof = tsgpu.Shout.new(pp).prove_arrays(entries, idx)
    want, _ = oracle.shout_prove(pw, pp.max_operations, entries, idx, fast=True)
    assert proof.to_bytes() == want
    assert tsgpu.Shout.new(pp).verify(proof, vp)


def test_tampered_opening_is_rejected(ctx, tsgpu, params3):
    """verify returns Ok(false) for a wrong evaluation (src/commitments.rs:523-541 wrong-value rejection)"""
    pp, vp = params3
    trace = tsgpu.MemoryTrace.new(8)
    for i in range(6):
        trace.write(i, tsgpu.fe(10 + i))
    twist = tsgpu.Twist.new(pp)
    proof = twist.prove(trace)
    assert twist.verify(proof, vp)
    proof.tamper_final_evaluation(0, tsgpu.fe(999))
    assert not twist.verify(proof, vp)


def test_reference_benchmark_pattern_2p16(ctx, tsgpu, oracle):
    """src/benchmarks.rs:88-99 generator at 2^16 ops (setup_params(14)): bytes equal to the CPU oracle's fast tier"""
    log_size = 14
    pp, vp = tsgpu.setup_params(ctx, log_size)
    n = 1 << 16
    msize = 1 << log_size
    mem = {}
    addr = np.empty(n, dtype=np.uint64); ints = [0] * n; isw = np.zeros(n, dtype=np.uint8)
    for i in range(n):
        if i % 3 == 0:
            a = i % msize; v = i * 42; mem[a] = v; isw[i] = 1
        else:
            a = (i // 2) % msize; v = mem.get(a, 0)
        addr[i] = a; ints[i] = v
    vals = tsgpu.fe_vec(ints)
    proof = tsgpu.Twist.new(pp).prove_arrays(addr, vals, isw)
    assert tsgpu.Twist.new(pp).verify(proof, vp)
    want, _ = oracle.twist_prove(pp.srs.download(), pp.max_operations, addr, vals, isw, fast=True)
    assert proof.to_bytes() == want

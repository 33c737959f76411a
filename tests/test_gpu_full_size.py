"""GPU parity at the FULL sizes of BASELINE.json, where the CPU oracle is too slow to recompute the answer, through
size-independent properties of the domain: the pairing equation of every KZG opening (Twist::verify), byte equality of
the two proving paths, linearity of the commitment, sum-check prove -> verify, the fold / evaluate identity
f(r, x) = fold(f, r)(x), and consistency of the final evaluation with MultilinearExtension::evaluate.  Bit-exact."""
import importlib

import numpy as np
import pytest

from conftest import seed_bytes

pytestmark = pytest.mark.gpu
PKG = "multilinear-map-cryptography_b200"


def _trace(n, cells, seed):
    rng = np.random.default_rng(seed)
    addr = rng.integers(0, cells, size=n, dtype=np.uint64)
    vals = rng.integers(0, 1 << 63, size=n, dtype=np.uint64)
    isw = rng.integers(0, 2, size=n, dtype=np.uint8)
    return addr, vals, isw


@pytest.fixture(scope="module")
def params18(ctx, tsgpu):
    return tsgpu.setup_params(ctx, 18)          # max_operations = 2^20 (configs[1])


def test_twist_2p20_ops_verifies_and_both_paths_agree(ctx, tsgpu, params18):
    """configs[1]: 2^16 cells, 2^20 operations.  verify() replays the transcript and checks both openings with the pairing."""
    pp, vp = params18
    n = 1 << 20
    addr, vals_u64, isw = _trace(n, 1 << 16, 7)
    vals = tsgpu.fe_vec(vals_u64)
    twist = tsgpu.Twist.new(pp)
    proof = twist.prove_arrays(addr, vals, isw)
    assert len(proof.round_polynomials) == 20 and len(proof.opening_proofs) == 2
    assert twist.verify(proof, vp)
    try:
        ctx.set_tuning("eval_basis", 0)
        coeff = twist.prove_arrays(addr, vals, isw)
    finally:
        ctx.set_tuning("eval_basis", 1)
    assert coeff.to_bytes() == proof.to_bytes()
    proof.tamper_final_evaluation(1, tsgpu.fe(1))
    assert not twist.verify(proof, vp)


def test_twist_2p20_overlapped_upload_gives_the_same_bytes(ctx, tsgpu, params18):
    """Twist::prove from host buffers commits the address vector while the values travel on a side stream (h2d_overlap, traces of >= 24 MiB of values):
    same bytes as with the plain upload + one batched commit pass; a vector may be freed while its copy is in flight."""
    import ctypes as C
    bd = importlib.import_module(PKG + ".binding")
    pp, vp = params18
    for n in (1 << 20, (1 << 20) - 12345):
        addr, vals_u64, isw = _trace(n, 1 << 16, 21)
        vals = tsgpu.fe_vec(vals_u64)
        twist = tsgpu.Twist.new(pp)
        a = twist.prove_arrays(addr, vals, isw).to_bytes()
        try:
            ctx.set_tuning("h2d_overlap", 0)
            b = twist.prove_arrays(addr, vals, isw).to_bytes()
        finally:
            ctx.set_tuning("h2d_overlap", 1)
        assert a == b
    h = C.c_void_p()
    ctx.check(tsgpu.lib().tsgpu_poly_upload_padded_async(ctx._h, bd._p(vals), C.c_size_t(n), C.c_size_t(1 << 20), C.byref(h)))
    assert tsgpu.lib().tsgpu_poly_in_flight(h) == 1
    tsgpu.lib().tsgpu_poly_free(ctx._h, h)                      # freed while the copy may still run: the free is ordered behind it
    h = C.c_void_p()
    ctx.check(tsgpu.lib().tsgpu_poly_upload_padded_async(ctx._h, bd._p(vals), C.c_size_t(n), C.c_size_t(1 << 20), C.byref(h)))
    ctx.check(tsgpu.lib().tsgpu_poly_wait(ctx._h, h))
    assert tsgpu.lib().tsgpu_poly_in_flight(h) == 0
    out = np.empty((1 << 20, 4), dtype=np.uint64)
    ctx.check(tsgpu.lib().tsgpu_poly_download(ctx._h, h, bd._p(out)))
    assert (out[:n] == vals).all() and not out[n:].any()
    tsgpu.lib().tsgpu_poly_free(ctx._h, h)


def test_twist_2p20_full_width_values_and_ragged_length(ctx, tsgpu, oracle, params18):
    """field-sized memory values (the short-scalar tables must be bypassed) and a length that is not a power of two (zero padding)"""
    pp, vp = params18
    n = (1 << 20) - 12345
    addr, _, isw = _trace(n, 1 << 16, 9)
    vals = oracle.chacha_fr_rand(seed_bytes(77), n).reshape(n, 4)
    twist = tsgpu.Twist.new(pp)
    proof = twist.prove_arrays(addr, vals, isw)
    assert twist.verify(proof, vp)
    try:
        ctx.set_tuning("eval_basis", 0)
        assert twist.prove_arrays(addr, vals, isw).to_bytes() == proof.to_bytes()
    finally:
        ctx.set_tuning("eval_basis", 1)


def test_commitment_is_linear_at_2p20(ctx, tsgpu, oracle, params18):
    """commit(a) + commit(b) == commit(a + b) for 2^20 full-width coefficients (window-table MSM), and for value vectors"""
    pp, _ = params18
    n = 1 << 20
    a = oracle.chacha_fr_rand(seed_bytes(1), n).reshape(n, 4)
    b = oracle.chacha_fr_rand(seed_bytes(2), n).reshape(n, 4)
    s = oracle.field_binop("fr", "add", a, b)
    K = tsgpu.KZGCommitment
    pa, pb, ps = ctx.poly_upload(a), ctx.poly_upload(b), ctx.poly_upload(s)
    lib = tsgpu.lib()
    for commit in (K.commit, K.commit_values):
        ca, cb, cs = commit(pp.srs, pa), commit(pp.srs, pb), commit(pp.srs, ps)
        out = np.empty(12, dtype=np.uint64)
        import ctypes as C
        lib.tsgpu_g1_add(ca.ctypes.data_as(C.c_void_p), cb.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        assert tsgpu.g1_compress(out) == tsgpu.g1_compress(cs)


def test_shout_2p18_table_2p20_lookups_verifies(ctx, tsgpu, params18):
    """the Shout shape of configs[2] scaled to the 2^20-operation parameters: two vectors of different lengths in one batched pass"""
    pp, vp = params18
    T, L = 1 << 18, 1 << 20
    entries = tsgpu.fe_vec(np.arange(T, dtype=np.uint64) ** 2)
    idx = np.random.default_rng(3).integers(0, T, size=L, dtype=np.uint64)
    shout = tsgpu.Shout.new(pp)
    proof = shout.prove_arrays(entries, idx)
    assert shout.verify(proof, vp)
    try:
        ctx.set_tuning("eval_basis", 0)
        assert shout.prove_arrays(entries, idx).to_bytes() == proof.to_bytes()
    finally:
        ctx.set_tuning("eval_basis", 1)


@pytest.mark.parametrize("nv", [22, 24])
def test_sumcheck_full_size_prove_verify_and_final_evaluation(ctx, tsgpu, oracle, nv):
    """configs[3] shape (eq x one-hot) at 2^22 / 2^24 entries: the proof verifies, and final_evaluation equals the product of
    MultilinearExtension::evaluate of the two tables at the challenge point (sumcheck.rs:104)"""
    dd = importlib.import_module(PKG + ".distributed")
    logK = 10
    w, addr = oracle.chacha_fr_then_u64(bytes([4]) * 32, nv, 1 << (nv - logK))
    A = ctx.table_eq(w.reshape(nv, 4)); B = ctx.table_one_hot_rows(addr % np.uint64(1 << logK), logK, nv)
    sc = ctx.sumcheck([A.clone(), B.clone()]); ev = sc.round_eval(); sc.end()
    claimed = dd.fr_add(ev[0], ev[1])
    proof, chals, finals = tsgpu.SumCheck(nv, claimed).prove_product(ctx, [A.clone(), B.clone()], tsgpu.Transcript(), return_aux=True)
    ok, vch = tsgpu.SumCheck(nv, claimed).verify(proof, tsgpu.Transcript())
    assert ok and (vch == chals).all()
    ea, eb = A.evaluate(chals), B.evaluate(chals)
    assert (finals[0] == ea).all() and (finals[1] == eb).all()
    assert (proof.final_evaluation == dd.fr_mul(ea, eb)).all()


def test_fold_identity_at_2p24(ctx, tsgpu, oracle):
    """f(r_0, x_1..) == bind(f, r_0)(x_1..): the constant-table fold against the evaluate kernels on a 2^24-entry table"""
    nv = 24
    w = oracle.chacha_fr_rand(seed_bytes(5), nv).reshape(nv, 4)
    pt = oracle.chacha_fr_rand(seed_bytes(6), nv).reshape(nv, 4)
    T = ctx.table_eq(w)
    full = T.evaluate(pt)
    F = T.clone(); F.bind(pt[0:1])
    assert F.num_vars == nv - 1
    assert (F.evaluate(pt[1:]) == full).all()
    # eq(w, pt) has a closed form: prod_j (w_j pt_j + (1 - w_j)(1 - pt_j))
    P = oracle.R_MOD
    wi, pi = oracle.fr_to_ints(w), oracle.fr_to_ints(pt)
    want = 1
    for a, b in zip(wi, pi):
        want = want * ((a * b + (1 - a) * (1 - b)) % P) % P
    assert oracle.fr_to_ints(full.reshape(1, 4))[0] == want

//! cuda-sys - B200 (sm_100a) prover backend for the `twist-and-shout` crate.
//!
//! Thin FFI to `libtsgpu.so` (`include/tsgpu.h`) plus safe wrappers that keep the reference's API surface:
//! `setup_params`, `Twist::prove/verify`, `Shout::prove/verify`, the `CommitmentScheme` trait
//! (src/commitments.rs:15-59) and `MultilinearExtension::evaluate / partial_evaluate` (src/polynomials.rs:85,126).
//! The Fiat-Shamir transcript stays on the host; every heavy step runs in hand-written CUDA kernels.
//! There is no CPU fallback: every call fails with the library's error when no CUDA device is present.
//!
//! STATUS: NOT compiled in the repository's build image (no Rust toolchain there).  The C ABI bound here is the
//! one the repository builds and tests (`tests/`, ~230 GPU parity tests through ctypes).
pub mod ffi;

use ark_bn254::{Fr, G1Projective};
use ark_ff::{One, Zero};
use std::ffi::CStr;
use std::os::raw::c_int;
use std::ptr;
use twist_and_shout::commitments::{CommitmentScheme, KZGCommitment, KZGCommitmentValue, KZGProof};
use twist_and_shout::polynomials::{poly_utils, MultilinearExtension};
use twist_and_shout::sumcheck::{SumCheck, SumCheckProof};
use twist_and_shout::utils::{field_utils, CommitmentParams, CommitmentVerificationKey, ProverParams, Transcript, VerifierParams};   // the two commitment structs live in utils.rs:52-76
use twist_and_shout::{LookupTable, MemoryOp, MemoryTrace, Result, ShoutProof, TwistAndShoutError, TwistProof};

// layout contract of the boundary (ark-ff / ark-ec 0.4.2): no conversion on either side
const _: () = assert!(core::mem::size_of::<Fr>() == 32 && core::mem::size_of::<G1Projective>() == 96);

/// One GPU + one stream.  Not re-entrant: one proof at a time per context; use one context per thread / per GPU.
pub struct Context { raw: *mut ffi::Ctx }
unsafe impl Send for Context {}

impl Context {
    pub fn new(device: i32) -> Result<Self> {
        let mut raw = ptr::null_mut();
        let rc = unsafe { ffi::tsgpu_init(device as c_int, ptr::null_mut(), &mut raw) };
        if rc != 0 { return Err(TwistAndShoutError::ProofGeneration(format!("tsgpu_init failed on device {device} (no CUDA device?)"))); }
        assert_eq!(unsafe { ffi::tsgpu_abi_version() }, 1);
        Ok(Self { raw })
    }
    fn check(&self, rc: c_int) -> Result<()> {
        use TwistAndShoutError as E;
        if rc == 0 { return Ok(()); }
        let msg = unsafe { CStr::from_ptr(ffi::tsgpu_last_error(self.raw)) }.to_string_lossy().into_owned();
        Err(match rc { 1 => E::InvalidParameters(msg), 2 => E::ProofGeneration(msg), 3 => E::ProofVerification(msg),
                       4 => E::Commitment(msg), 5 => E::Polynomial(msg), _ => E::SumCheck(msg) })
    }
    /// kernels launched so far (the driver's evidence that the CUDA path ran)
    pub fn launch_count(&self) -> u64 { unsafe { ffi::tsgpu_launch_count(self.raw) } }
    /// one process per GPU: rank 0 creates the id (`comm_unique_id`), the host program carries it to every rank
    pub fn comm_unique_id() -> [u8; 128] { let mut id = [0u8; 128]; unsafe { ffi::tsgpu_comm_unique_id(id.as_mut_ptr()) }; id }
    pub fn comm_init(&self, nranks: i32, rank: i32, id: &[u8; 128]) -> Result<()> { self.check(unsafe { ffi::tsgpu_comm_init(self.raw, nranks, rank, id.as_ptr()) }) }
}
impl Drop for Context { fn drop(&mut self) { unsafe { ffi::tsgpu_destroy(self.raw) } } }

// ------------------------------------------------------------------------------------------------ SRS handle
/// `CommitmentParams.g1_powers` resident on the device as affine points (+ window tables, built on first use).
pub struct DeviceSrs<'c> { ctx: &'c Context, raw: *mut ffi::Srs }
impl<'c> DeviceSrs<'c> {
    /// upload `params.g1_powers` (batch-normalised to affine on the device)
    pub fn upload(ctx: &'c Context, params: &CommitmentParams) -> Result<Self> {
        let mut raw = ptr::null_mut();
        ctx.check(unsafe { ffi::tsgpu_srs_upload(ctx.raw, params.g1_powers.as_ptr(), params.g1_powers.len(), &mut raw) })?;
        Ok(Self { ctx, raw })
    }
    /// `g1_powers[i] = G * tau^i`, i < n, generated on the device (src/utils.rs:89-96) - same group elements, z = 1
    pub fn generate(ctx: &'c Context, tau: Fr, n: usize) -> Result<Self> {
        let mut raw = ptr::null_mut();
        ctx.check(unsafe { ffi::tsgpu_srs_generate(ctx.raw, &tau, n, &mut raw) })?;
        Ok(Self { ctx, raw })
    }
    pub fn len(&self) -> usize { unsafe { ffi::tsgpu_srs_len(self.raw) } }
    pub fn download(&self) -> Result<Vec<G1Projective>> {
        let mut v = vec![G1Projective::default(); self.len()];
        self.ctx.check(unsafe { ffi::tsgpu_srs_download(self.ctx.raw, self.raw, 0, v.len(), v.as_mut_ptr()) })?;
        Ok(v)
    }
    /// `KZGCommitment::commit` (src/commitments.rs:162-180): Pippenger MSM on the device
    pub fn commit(&self, polynomial: &[Fr]) -> Result<KZGCommitmentValue> {
        let mut out = G1Projective::default();
        self.ctx.check(unsafe { ffi::tsgpu_kzg_commit(self.ctx.raw, self.raw, polynomial.as_ptr(), polynomial.len(), &mut out) })?;
        Ok(KZGCommitmentValue { commitment: out })
    }
    /// `KZGCommitment::open` (src/commitments.rs:182-199): Horner value + quotient scan + MSM
    pub fn open(&self, polynomial: &[Fr], point: Fr) -> Result<(Fr, KZGProof)> {
        let (mut v, mut pi) = (Fr::zero(), G1Projective::default());
        self.ctx.check(unsafe { ffi::tsgpu_kzg_open(self.ctx.raw, self.raw, polynomial.as_ptr(), polynomial.len(), &point, &mut v, &mut pi) })?;
        Ok((v, KZGProof { proof: pi }))
    }
}
impl Drop for DeviceSrs<'_> { fn drop(&mut self) { unsafe { ffi::tsgpu_srs_free(self.ctx.raw, self.raw) } } }

/// `impl CommitmentScheme` with the reference's associated-function shape (no `self`): the context and the uploaded SRS
/// live in a thread-local cache keyed on the identity of `params.g1_powers` (pointer, length, last element), because
/// `Twist::new` / `Shout::new` deep-copy the params (src/twist.rs:100-104).
pub struct GpuKzg;
thread_local! {
    static CACHE: std::cell::RefCell<Option<(usize, usize, G1Projective, &'static Context, DeviceSrs<'static>)>> = std::cell::RefCell::new(None);
}
fn with_cached_srs<T>(params: &CommitmentParams, f: impl FnOnce(&DeviceSrs<'static>) -> Result<T>) -> Result<T> {
    CACHE.with(|c| {
        let mut c = c.borrow_mut();
        let key = (params.g1_powers.as_ptr() as usize, params.g1_powers.len(), *params.g1_powers.last().unwrap_or(&G1Projective::default()));
        let hit = matches!(&*c, Some((p, n, last, _, _)) if *p == key.0 && *n == key.1 && *last == key.2);
        if !hit {
            *c = None;                                                   // drops the previous SRS before its context
            let ctx: &'static Context = Box::leak(Box::new(Context::new(0)?));
            let srs = DeviceSrs::upload(ctx, params)?;
            *c = Some((key.0, key.1, key.2, ctx, srs));
        }
        f(&c.as_ref().unwrap().4)
    })
}
impl CommitmentScheme for GpuKzg {
    type Commitment = KZGCommitmentValue;
    type Proof = KZGProof;
    type Params = CommitmentParams;
    type VerifyKey = CommitmentVerificationKey;
    fn commit(params: &Self::Params, polynomial: &[Fr]) -> Result<Self::Commitment> { with_cached_srs(params, |s| s.commit(polynomial)) }
    fn open(params: &Self::Params, polynomial: &[Fr], point: Fr) -> Result<(Fr, Self::Proof)> { with_cached_srs(params, |s| s.open(polynomial, point)) }
    fn verify(vk: &Self::VerifyKey, c: &Self::Commitment, point: Fr, value: Fr, proof: &Self::Proof) -> Result<bool> {
        KZGCommitment::verify(vk, c, point, value, proof)               // two pairings: stays on the CPU (arkworks)
    }
}

// ------------------------------------------------------------------------------------------------ MLE
/// `MultilinearExtension::evaluate` (src/polynomials.rs:85-122): one streaming pass over the table on the device
pub fn mle_evaluate(ctx: &Context, mle: &MultilinearExtension, point: &[Fr]) -> Result<Fr> {
    assert_eq!(point.len(), mle.num_vars, "Point dimension must match number of variables");
    let mut out = Fr::zero();
    ctx.check(unsafe { ffi::tsgpu_mle_evaluate(ctx.raw, mle.evaluations.as_ptr(), mle.num_vars as u32, point.as_ptr(), &mut out) })?;
    Ok(out)
}
/// `MultilinearExtension::partial_evaluate` (src/polynomials.rs:126-161): fixes the first k variables
pub fn mle_partial_evaluate(ctx: &Context, mle: &MultilinearExtension, fixed: &[Fr]) -> Result<MultilinearExtension> {
    assert!(fixed.len() <= mle.num_vars, "Cannot fix more variables than available");
    let mut out = vec![Fr::zero(); 1usize << (mle.num_vars - fixed.len())];
    ctx.check(unsafe { ffi::tsgpu_mle_partial_evaluate(ctx.raw, mle.evaluations.as_ptr(), mle.num_vars as u32, fixed.as_ptr(), fixed.len() as u32, out.as_mut_ptr()) })?;
    Ok(MultilinearExtension::from_evaluations_vec(mle.num_vars - fixed.len(), out))
}

// ------------------------------------------------------------------------------------------------ sum-check
/// Structured sibling of `SumCheck::prove` (the closure-typed original cannot be offloaded): byte-identical to
/// `sc.prove(|v| tables.iter().map(|t| t.evaluate(v)).product(), transcript)` for 1..=3 tables.
/// Round evaluation and table folding run on the device; interpolation of the four evaluations, the
/// g(0) + g(1) check and the transcript are the reference's own code (src/sumcheck.rs:77-100).
pub fn sumcheck_prove_product(ctx: &Context, sc: &SumCheck, tables: &[MultilinearExtension], transcript: &mut Transcript) -> Result<SumCheckProof> {
    assert!((1..=3).contains(&tables.len()) && tables.iter().all(|t| t.num_vars == sc.num_vars));
    let mut handles: Vec<*mut ffi::Table> = Vec::new();
    let free_all = |h: &Vec<*mut ffi::Table>| for &t in h { unsafe { ffi::tsgpu_table_free(ctx.raw, t) } };
    for t in tables {
        let mut h = ptr::null_mut();
        if let Err(e) = ctx.check(unsafe { ffi::tsgpu_table_upload(ctx.raw, t.evaluations.as_ptr(), t.evaluations.len(), t.num_vars as u32, &mut h) }) { free_all(&handles); return Err(e); }
        handles.push(h);
    }
    let mut raw_sc = ptr::null_mut();
    if let Err(e) = ctx.check(unsafe { ffi::tsgpu_sc_begin(ctx.raw, handles.as_ptr(), handles.len() as c_int, &mut raw_sc) }) { free_all(&handles); return Err(e); }
    // this loop enqueues nothing else on the context until tsgpu_sc_end: the small d = 2 rounds may run in the resident tail kernel
    unsafe { ffi::tsgpu_sc_exclusive(raw_sc, 1); }
    let result = (|| -> Result<SumCheckProof> {
        let xs: Vec<Fr> = (0..4u64).map(Fr::from).collect();
        let mut round_polynomials = Vec::with_capacity(sc.num_vars);
        let mut current_sum = sc.claimed_sum;
        let mut evals = [Fr::zero(); 4];
        let mut finals = vec![Fr::zero(); tables.len()];
        if sc.num_vars > 0 { ctx.check(unsafe { ffi::tsgpu_sc_round_eval(raw_sc, evals.as_mut_ptr()) })?; }
        for round in 0..sc.num_vars {
            let pts: Vec<(Fr, Fr)> = xs.iter().cloned().zip(evals.iter().cloned()).collect();
            let round_poly = poly_utils::lagrange_interpolate(&pts);                          // sumcheck.rs:201-205
            let g0 = field_utils::horner_eval(&round_poly, Fr::zero());
            let g1 = field_utils::horner_eval(&round_poly, Fr::one());
            if g0 + g1 != current_sum { return Err(TwistAndShoutError::SumCheck(format!("Round {} consistency check failed", round))); }
            transcript.append_field_elements(format!("sumcheck_round_{}", round).as_bytes(), &round_poly);
            let challenge = transcript.challenge_field_element(format!("sumcheck_challenge_{}", round).as_bytes());
            current_sum = field_utils::horner_eval(&round_poly, challenge);
            round_polynomials.push(round_poly);
            if round + 1 < sc.num_vars {
                // fused fold + next round; with two tables g(1) is derived from the claim (= current_sum) on the device
                let rc = if tables.len() == 2 { unsafe { ffi::tsgpu_sc_bind_eval_claim(raw_sc, &challenge, &current_sum, evals.as_mut_ptr()) } }
                         else { unsafe { ffi::tsgpu_sc_bind_eval(raw_sc, &challenge, evals.as_mut_ptr()) } };
                ctx.check(rc)?;
            } else {
                ctx.check(unsafe { ffi::tsgpu_sc_bind(raw_sc, &challenge) })?;
            }
        }
        ctx.check(unsafe { ffi::tsgpu_sc_final(raw_sc, finals.as_mut_ptr()) })?;
        let final_evaluation = finals.iter().fold(Fr::one(), |a, b| a * b);                 // polynomial(&fixed_variables), sumcheck.rs:104
        Ok(SumCheckProof { round_polynomials, final_evaluation })
    })();
    unsafe { ffi::tsgpu_sc_end(raw_sc) };
    free_all(&handles);
    result
}

// ------------------------------------------------------------------------------------------------ setup_params, Twist, Shout
/// `setup_params(log_size)` (src/utils.rs:79-131) with the SRS generated on the device.  The returned reference structs are
/// complete (`g1_powers` downloaded), so every reference function keeps working on them; `GpuParams` keeps the device copy.
pub struct GpuParams<'c> { ctx: &'c Context, raw: *mut ffi::Params, pub prover: ProverParams, pub verifier: VerifierParams }
pub fn setup_params<'c>(ctx: &'c Context, log_size: usize) -> Result<GpuParams<'c>> {
    let mut raw = ptr::null_mut();
    ctx.check(unsafe { ffi::tsgpu_setup_params(ctx.raw, log_size, &mut raw) })?;
    // tau, g2_tau and the seed come from the same ChaCha20 stream as the reference; the G2 side is three group elements: computed by
    // the reference itself on a tiny instance would cost a full CPU setup, so they are rebuilt here from tau (utils.rs:84,98,101-102)
    let (mut tau, mut seed) = (Fr::zero(), [0u8; 32]);
    unsafe { ffi::tsgpu_params_tau(raw, &mut tau); ffi::tsgpu_params_fiat_shamir_seed(raw, seed.as_mut_ptr()); }
    let srs = unsafe { ffi::tsgpu_params_srs(raw) };
    let n = unsafe { ffi::tsgpu_srs_len(srs) };
    let mut g1_powers = vec![G1Projective::default(); n];
    ctx.check(unsafe { ffi::tsgpu_srs_download(ctx.raw, srs, 0, n, g1_powers.as_mut_ptr()) })?;
    use ark_ec::Group;
    let (g1_gen, g2_gen) = (G1Projective::generator(), ark_bn254::G2Projective::generator());
    let max_operations = unsafe { ffi::tsgpu_params_max_operations(raw) };
    let prover = ProverParams { log_size, max_operations, fiat_shamir_seed: seed,
                                commitment_params: CommitmentParams { g1_powers, g2_generator: g2_gen, tau: Some(tau) } };
    let verifier = VerifierParams { log_size, max_operations, fiat_shamir_seed: seed,
                                    commitment_vk: CommitmentVerificationKey { g1_generator: g1_gen, g2_generator: g2_gen, g2_tau: g2_gen * tau } };
    Ok(GpuParams { ctx, raw, prover, verifier })
}
impl Drop for GpuParams<'_> { fn drop(&mut self) { unsafe { ffi::tsgpu_params_free(self.ctx.raw, self.raw) } } }

struct ProofHandle(*mut ffi::Proof);
impl Drop for ProofHandle { fn drop(&mut self) { unsafe { ffi::tsgpu_proof_free(self.0) } } }
impl ProofHandle {
    fn parts(&self) -> (KZGCommitmentValue, KZGCommitmentValue, SumCheckProof, Vec<KZGProof>, Vec<Fr>) {
        unsafe {
            let (mut c0, mut c1) = (G1Projective::default(), G1Projective::default());
            ffi::tsgpu_proof_commitment(self.0, 0, &mut c0); ffi::tsgpu_proof_commitment(self.0, 1, &mut c1);
            let rounds = ffi::tsgpu_proof_num_rounds(self.0);
            let mut flat = vec![Fr::zero(); 4 * rounds];
            ffi::tsgpu_proof_round_polynomials(self.0, flat.as_mut_ptr());
            let mut fe = Fr::zero();
            ffi::tsgpu_proof_final_evaluation(self.0, &mut fe);
            let (mut ops, mut evs) = (Vec::new(), Vec::new());
            for i in 0..ffi::tsgpu_proof_num_openings(self.0) {
                let (mut p, mut v) = (G1Projective::default(), Fr::zero());
                ffi::tsgpu_proof_opening(self.0, i, &mut p, &mut v);
                ops.push(KZGProof { proof: p }); evs.push(v);
            }
            (KZGCommitmentValue { commitment: c0 }, KZGCommitmentValue { commitment: c1 },
             SumCheckProof { round_polynomials: flat.chunks(4).map(|c| c.to_vec()).collect(), final_evaluation: fe }, ops, evs)
        }
    }
    /// canonical proof bytes (ark-serialize encodings)
    fn bytes(&self) -> Vec<u8> {
        let n = unsafe { ffi::tsgpu_proof_bytes(self.0, ptr::null_mut(), 0) };
        let mut b = vec![0u8; n];
        unsafe { ffi::tsgpu_proof_bytes(self.0, b.as_mut_ptr(), n) };
        b
    }
}

/// `Twist::prove` (src/twist.rs:107-252) on the device; the result is the reference's own `TwistProof`, so
/// `Twist::verify` of the unchanged crate accepts it.
pub fn twist_prove(params: &GpuParams, trace: &MemoryTrace) -> Result<TwistProof> {
    let (mut addr, mut vals, mut isw) = (Vec::new(), Vec::new(), Vec::new());
    for op in &trace.operations {
        match op { MemoryOp::Read { address, value } => { addr.push(*address as u64); vals.push(*value); isw.push(0u8); }
                   MemoryOp::Write { address, value } => { addr.push(*address as u64); vals.push(*value); isw.push(1u8); } }
    }
    let mut raw = ptr::null_mut();
    params.ctx.check(unsafe { ffi::tsgpu_twist_prove(params.ctx.raw, params.raw, addr.as_ptr(), vals.as_ptr(), isw.as_ptr(), addr.len(), &mut raw) })?;
    let (address_commitment, value_commitment, consistency_proof, opening_proofs, final_evaluations) = ProofHandle(raw).parts();
    Ok(TwistProof { address_commitment, value_commitment, consistency_proof, opening_proofs, final_evaluations })
}
/// `Shout::prove` (src/shout.rs:97-222) on the device
pub fn shout_prove(params: &GpuParams, table: &LookupTable) -> Result<ShoutProof> {
    let idx: Vec<u64> = table.lookups.iter().map(|l| l.index as u64).collect();
    let mut raw = ptr::null_mut();
    params.ctx.check(unsafe { ffi::tsgpu_shout_prove(params.ctx.raw, params.raw, table.entries.as_ptr(), table.entries.len(), idx.as_ptr(), idx.len(), &mut raw) })?;
    let (table_commitment, index_commitment, lookup_proof, opening_proofs, final_evaluations) = ProofHandle(raw).parts();
    Ok(ShoutProof { table_commitment, index_commitment, lookup_proof, opening_proofs, final_evaluations })
}
/// canonical bytes of a device proof of the same trace (for comparisons with `tests/golden/appendix_c.json` of the repository)
pub fn twist_prove_bytes(params: &GpuParams, addresses: &[u64], values: &[Fr]) -> Result<Vec<u8>> {
    assert_eq!(addresses.len(), values.len());
    let mut raw = ptr::null_mut();
    params.ctx.check(unsafe { ffi::tsgpu_twist_prove(params.ctx.raw, params.raw, addresses.as_ptr(), values.as_ptr(), ptr::null(), addresses.len(), &mut raw) })?;
    Ok(ProofHandle(raw).bytes())
}

#[cfg(test)]
mod tests {
    //! The parity checks a maintainer runs against REAL arkworks (this is what finally pins the oracle):
    //! every device result must equal the unchanged reference on the same inputs.
    use super::*;
    use ark_std::UniformRand;
    use twist_and_shout::{Shout, Twist};

    #[test]
    fn demo_proofs_equal_the_reference() {
        let ctx = Context::new(0).unwrap();
        let gp = setup_params(&ctx, 3).unwrap();
        let (pp, vp) = twist_and_shout::setup_params(3);
        assert_eq!(gp.prover.commitment_params.g1_powers, pp.commitment_params.g1_powers);
        assert_eq!(gp.prover.fiat_shamir_seed, pp.fiat_shamir_seed);
        let mut trace = MemoryTrace::new(8);                                   // examples/demo.rs:33-46
        trace.write(0, Fr::from(42u64)).unwrap(); trace.write(1, Fr::from(100u64)).unwrap();
        trace.read(0).unwrap(); trace.read(1).unwrap(); trace.write(0, Fr::from(43u64)).unwrap(); trace.read(0).unwrap();
        let want = Twist::new(&pp).prove(&trace).unwrap();
        let got = twist_prove(&gp, &trace).unwrap();
        assert_eq!(got.address_commitment, want.address_commitment);
        assert_eq!(got.value_commitment, want.value_commitment);
        assert_eq!(got.consistency_proof.round_polynomials, want.consistency_proof.round_polynomials);
        assert_eq!(got.final_evaluations, want.final_evaluations);
        for (a, b) in got.opening_proofs.iter().zip(&want.opening_proofs) { assert_eq!(a.proof, b.proof); }
        assert!(Twist::new(&pp).verify(&got, &vp).unwrap());
        let mut table = LookupTable::new((0..8u64).map(|i| Fr::from(i * i)).collect());   // examples/demo.rs:66-79
        for i in [3usize, 5, 0, 7] { table.lookup(i).unwrap(); }
        let want = Shout::new(&pp).prove(&table).unwrap();
        let got = shout_prove(&gp, &table).unwrap();
        assert_eq!(got.table_commitment, want.table_commitment);
        assert_eq!(got.index_commitment, want.index_commitment);
        assert_eq!(got.final_evaluations, want.final_evaluations);
        assert!(Shout::new(&pp).verify(&got, &vp).unwrap());
    }

    #[test]
    fn product_sumcheck_equals_the_closure_sumcheck() {
        let ctx = Context::new(0).unwrap();
        let mut rng = ark_std::test_rng();
        let nv = 6;
        let a = MultilinearExtension::from_evaluations((0..1 << nv).map(|_| Fr::rand(&mut rng)).collect());
        let b = MultilinearExtension::from_evaluations((0..1 << nv).map(|_| Fr::rand(&mut rng)).collect());
        let claimed: Fr = a.evaluations.iter().zip(&b.evaluations).map(|(x, y)| *x * y).sum();
        let sc = SumCheck::new(nv, claimed);
        let (a2, b2) = (a.clone(), b.clone());
        let want = sc.prove(move |v| a2.evaluate(v) * b2.evaluate(v), &mut Transcript::new(&[0u8; 32])).unwrap();
        let got = sumcheck_prove_product(&ctx, &sc, &[a.clone(), b.clone()], &mut Transcript::new(&[0u8; 32])).unwrap();
        assert_eq!(got.round_polynomials, want.round_polynomials);
        assert_eq!(got.final_evaluation, want.final_evaluation);
        let r: Vec<Fr> = (0..nv).map(|_| Fr::rand(&mut rng)).collect();
        assert_eq!(mle_evaluate(&ctx, &a, &r).unwrap(), a.evaluate(&r));
        assert_eq!(mle_partial_evaluate(&ctx, &a, &r[..2]).unwrap().evaluations, a.partial_evaluate(&r[..2]).evaluations);
    }

    #[test]
    fn kzg_commit_open_equal_the_reference() {
        let (pp, vp) = twist_and_shout::setup_params(2);
        let mut rng = ark_std::test_rng();
        let poly: Vec<Fr> = (0..13).map(|_| Fr::rand(&mut rng)).collect();
        let z = Fr::rand(&mut rng);
        let c = GpuKzg::commit(&pp.commitment_params, &poly).unwrap();
        assert_eq!(c, KZGCommitment::commit(&pp.commitment_params, &poly).unwrap());
        let (v, pi) = GpuKzg::open(&pp.commitment_params, &poly, z).unwrap();
        let (v2, pi2) = KZGCommitment::open(&pp.commitment_params, &poly, z).unwrap();
        assert_eq!(v, v2); assert_eq!(pi.proof, pi2.proof);
        assert!(GpuKzg::verify(&vp.commitment_vk, &c, z, v, &pi).unwrap());
    }
}

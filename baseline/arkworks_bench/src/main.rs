//! CPU baseline on real arkworks (baseline "B-rs").  Two measurements per log2 size given on the command line:
//!   * the UNMODIFIED reference `Twist::prove` on the generator pattern of src/benchmarks.rs:88-99 - only for sizes <= 2^10
//!     (its interpolation is O(n^3), src/polynomials.rs:301-352);
//!   * `ark_ec::VariableBaseMSM::msm` over the reference's own SRS points with uniform scalars - the arkworks-class
//!     commitment (the reference itself commits with n serial scalar multiplications, src/commitments.rs:173-177).
//! Prints one JSON line per measurement with the rayon thread count.
//!
//! `cargo run --release -- golden > ../../tests/golden/from_arkworks.json` instead writes the golden vectors of
//! tests/golden/appendix_c.json as REAL arkworks / rand_chacha / std produce them (tau, the Fiat-Shamir seed, two SRS points, the
//! transcript test challenge, the demo / README / empty proofs in the canonical byte layout of SURVEY Appendix D, the C.3 product
//! sum-check).  tests/test_oracle_golden.py::test_vectors_from_real_arkworks compares that file with the oracle's vectors key by key when
//! it exists: this is the one command that PINS the oracle to the reference (the repository's build image has no Rust toolchain).
use ark_bn254::{Fr, G1Affine, G1Projective};
use ark_ec::{CurveGroup, VariableBaseMSM};
use ark_std::UniformRand;
use std::time::Instant;
use twist_and_shout::{setup_params, MemoryTrace, Twist};

fn best_of<T>(k: usize, mut f: impl FnMut() -> T) -> f64 {
    (0..k).map(|_| { let t = Instant::now(); let _ = std::hint::black_box(f()); t.elapsed().as_secs_f64() }).fold(f64::MAX, f64::min)
}

// ------------------------------------------------------------------------------------------------ golden vectors
mod golden {
    use ark_bn254::Fr;
    use ark_ec::CurveGroup;
    use ark_ff::{BigInteger, PrimeField};
    use ark_serialize::CanonicalSerialize;
    use twist_and_shout::commitments::{KZGCommitmentValue, KZGProof};
    use twist_and_shout::polynomials::MultilinearExtension;
    use twist_and_shout::sumcheck::{SumCheck, SumCheckProof};
    use twist_and_shout::utils::Transcript;
    use twist_and_shout::{setup_params, LookupTable, MemoryTrace, Shout, Twist};

    fn hex(b: &[u8]) -> String { b.iter().map(|x| format!("{x:02x}")).collect() }
    fn g1(p: &ark_bn254::G1Projective) -> Vec<u8> { let mut v = Vec::new(); p.into_affine().serialize_compressed(&mut v).unwrap(); v }
    fn fr(x: &Fr) -> Vec<u8> { let mut v = Vec::new(); x.serialize_compressed(&mut v).unwrap(); v }
    // compressed(C0) | compressed(C1) | u64 rounds | per round (u64 4 | 4 x Fr) | Fr final | u64 #openings | G1.. | u64 #evals | Fr..
    fn proof_bytes(c0: &KZGCommitmentValue, c1: &KZGCommitmentValue, sc: &SumCheckProof, op: &[KZGProof], ev: &[Fr]) -> Vec<u8> {
        let mut b = Vec::new();
        b.extend(g1(&c0.commitment)); b.extend(g1(&c1.commitment));
        b.extend((sc.round_polynomials.len() as u64).to_le_bytes());
        for r in &sc.round_polynomials { b.extend((r.len() as u64).to_le_bytes()); for c in r { b.extend(fr(c)); } }
        b.extend(fr(&sc.final_evaluation));
        b.extend((op.len() as u64).to_le_bytes()); for p in op { b.extend(g1(&p.proof)); }
        b.extend((ev.len() as u64).to_le_bytes()); for v in ev { b.extend(fr(v)); }
        b
    }
    fn twist(log_size: usize, cells: usize, ops: &[(char, usize, u64)]) -> String {
        let (pp, vp) = setup_params(log_size);
        let mut t = MemoryTrace::new(cells);
        for &(k, a, v) in ops { if k == 'W' { t.write(a, Fr::from(v)).unwrap(); } else { assert_eq!(t.read(a).unwrap(), Fr::from(v)); } }
        let tw = Twist::new(&pp);
        let p = tw.prove(&t).unwrap();
        assert!(tw.verify(&p, &vp).unwrap());
        let ops_json: Vec<String> = ops.iter().map(|(k, a, v)| format!("[\"{k}\", {a}, {v}]")).collect();
        format!("{{\"log_size\": {log_size}, \"ops\": [{}], \"proof_hex\": \"{}\"}}", ops_json.join(", "),
                hex(&proof_bytes(&p.address_commitment, &p.value_commitment, &p.consistency_proof, &p.opening_proofs, &p.final_evaluations)))
    }
    fn shout(log_size: usize, entries: &[u64], lookups: &[usize]) -> String {
        let (pp, vp) = setup_params(log_size);
        let mut t = LookupTable::new(entries.iter().map(|&e| Fr::from(e)).collect());
        for &i in lookups { t.lookup(i).unwrap(); }
        let sh = Shout::new(&pp);
        let p = sh.prove(&t).unwrap();
        assert!(sh.verify(&p, &vp).unwrap());
        format!("{{\"log_size\": {log_size}, \"entries\": {:?}, \"lookups\": {:?}, \"proof_hex\": \"{}\"}}", entries, lookups,
                hex(&proof_bytes(&p.table_commitment, &p.index_commitment, &p.lookup_proof, &p.opening_proofs, &p.final_evaluations)))
    }
    pub fn emit() {
        let (pp, _) = setup_params(3);
        let tau = pp.commitment_params.tau.unwrap();
        let mont: Vec<u8> = tau.0 .0.iter().rev().flat_map(|l| l.to_be_bytes()).collect();          // the in-memory Montgomery limbs, most significant first
        let (pp5, _) = setup_params(5);                                                             // 129 powers: indices 1 and 32 exist
        let mut tr = Transcript::new(&[0u8; 32]);
        tr.append_field_element(b"test", &Fr::from(123u64));
        let ch = tr.challenge_field_element(b"challenge");
        // C.3: product sum-check of two 8-entry tables on a fresh transcript
        let a = MultilinearExtension::from_evaluations(&(1u64..=8).map(Fr::from).collect::<Vec<_>>());
        let b = MultilinearExtension::from_evaluations(&[3u64, 1, 4, 1, 5, 9, 2, 6].iter().map(|&x| Fr::from(x)).collect::<Vec<_>>());
        let mut tr3 = Transcript::new(&[0u8; 32]);
        let sc = SumCheck::new(3, Fr::from(162u64)).prove(|v| a.evaluate(v) * b.evaluate(v), &mut tr3).unwrap();
        let rounds: Vec<String> = sc.round_polynomials.iter().map(|r| format!("[{}]", r.iter().map(|c| format!("\"{}\"", c.into_bigint())).collect::<Vec<_>>().join(", "))).collect();
        println!("{{");
        println!("  \"provenance\": \"baseline/arkworks_bench golden: real arkworks 0.4 / rand_chacha 0.3 / Rust std\",");
        println!("  \"tau\": \"{}\",", tau.into_bigint());
        println!("  \"tau_montgomery_limbs_hex\": \"0x{}\",", hex(&mont).trim_start_matches('0'));
        println!("  \"fiat_shamir_seed\": \"{}\",", hex(&pp.fiat_shamir_seed));
        println!("  \"g1_powers_1\": \"{}\",", hex(&g1(&pp5.commitment_params.g1_powers[1])));
        println!("  \"g1_powers_32\": \"{}\",", hex(&g1(&pp5.commitment_params.g1_powers[32])));
        println!("  \"transcript_test_challenge\": \"{}\",", ch.into_bigint());
        println!("  \"twist_demo\": {},", twist(3, 8, &[('W', 0, 42), ('W', 1, 100), ('R', 0, 42), ('R', 1, 100), ('W', 0, 43), ('R', 0, 43)]));
        println!("  \"shout_demo\": {},", shout(3, &[0, 1, 4, 9, 16, 25, 36, 49], &[3, 5, 0, 7]));
        println!("  \"twist_readme\": {},", twist(8, 256, &[('W', 0, 42), ('W', 1, 100), ('R', 0, 42)]));
        println!("  \"shout_readme\": {},", shout(8, &[1, 4, 9], &[1]));
        println!("  \"twist_empty\": {},", twist(3, 8, &[]));
        println!("  \"shout_no_lookups\": {},", shout(3, &[1, 2, 3, 4], &[]));
        println!("  \"sumcheck_c3\": {{\"A\": [1, 2, 3, 4, 5, 6, 7, 8], \"B\": [3, 1, 4, 1, 5, 9, 2, 6], \"claimed_sum\": 162, \"round_polynomials\": [{}], \"final_evaluation\": \"{}\"}}",
                 rounds.join(", "), sc.final_evaluation.into_bigint());
        println!("}}");
        let _ = BigInteger::to_bytes_le(&tau.into_bigint());
    }
}

fn main() {
    let cores = rayon::current_num_threads();
    if std::env::args().nth(1).as_deref() == Some("golden") {
        golden::emit();
        return;
    }
    for arg in std::env::args().skip(1) {
        let log_n: usize = arg.parse().expect("log2 size");
        let n = 1usize << log_n;
        if log_n <= 10 {
            let log_size = log_n.saturating_sub(2);
            let (pp, _vp) = setup_params(log_size);
            let cells = (n / 16).max(1).next_power_of_two();
            let mut trace = MemoryTrace::new(cells);
            for i in 0..n {                                              // src/benchmarks.rs:88-99
                if i % 3 == 0 { trace.write(i % cells, Fr::from((42 * i) as u64)).unwrap(); } else { trace.read((i / 2) % cells).unwrap(); }
            }
            let twist = Twist::new(&pp);
            let s = best_of(3, || twist.prove(&trace).unwrap());
            println!("{{\"what\": \"reference Twist::prove\", \"log_ops\": {log_n}, \"ms\": {:.3}, \"cores\": {cores}}}", s * 1e3);
        }
        let mut rng = ark_std::test_rng();
        let g = G1Projective::rand(&mut rng);
        let tau = Fr::rand(&mut rng);
        let mut p = g;
        let proj: Vec<G1Projective> = (0..n).map(|_| { let q = p; p *= tau; q }).collect();
        let bases: Vec<G1Affine> = G1Projective::normalize_batch(&proj);
        let scalars: Vec<Fr> = (0..n).map(|_| Fr::rand(&mut rng)).collect();
        let s = best_of(3, || G1Projective::msm(&bases, &scalars).unwrap());
        println!("{{\"what\": \"ark_ec VariableBaseMSM\", \"log_points\": {log_n}, \"ms\": {:.3}, \"points_per_s\": {:.0}, \"cores\": {cores}}}", s * 1e3, n as f64 / s);
    }
}

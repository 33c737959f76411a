//! CPU baseline on real arkworks (baseline "B-rs").  Two measurements per log2 size given on the command line:
//!   * the UNMODIFIED reference `Twist::prove` on the generator pattern of src/benchmarks.rs:88-99 - only for sizes <= 2^10
//!     (its interpolation is O(n^3), src/polynomials.rs:301-352);
//!   * `ark_ec::VariableBaseMSM::msm` over the reference's own SRS points with uniform scalars - the arkworks-class
//!     commitment (the reference itself commits with n serial scalar multiplications, src/commitments.rs:173-177).
//! Prints one JSON line per measurement with the rayon thread count.
use ark_bn254::{Fr, G1Affine, G1Projective};
use ark_ec::{CurveGroup, VariableBaseMSM};
use ark_std::UniformRand;
use std::time::Instant;
use twist_and_shout::{setup_params, MemoryTrace, Twist};

fn best_of<T>(k: usize, mut f: impl FnMut() -> T) -> f64 {
    (0..k).map(|_| { let t = Instant::now(); let _ = std::hint::black_box(f()); t.elapsed().as_secs_f64() }).fold(f64::MAX, f64::min)
}

fn main() {
    let cores = rayon::current_num_threads();
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

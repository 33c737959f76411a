//! Raw `extern "C"` declarations of include/tsgpu.h (abi version 1).  Layout contract: `ark_bn254::Fr` is one
//! `BigInt<4>` = `[u64; 4]` (little-endian limbs of a * 2^256 mod r) = `tsgpu_fr`; `G1Projective` is Jacobian
//! `{x, y, z: Fq}` = 96 bytes = `tsgpu_g1` (ark-ff / ark-ec 0.4.2; asserted in lib.rs).
#![allow(non_camel_case_types, dead_code)]
use ark_bn254::{Fr, G1Projective};
use std::os::raw::{c_char, c_int, c_uint, c_void};

macro_rules! opaque { ($($n:ident),*) => { $( #[repr(C)] pub struct $n { _p: [u8; 0] } )* } }
opaque!(Ctx, Table, Sc, Srs, Poly, Params, Proof, TranscriptH);

extern "C" {
    pub fn tsgpu_abi_version() -> c_int;
    pub fn tsgpu_init(device: c_int, stream: *mut c_void, out: *mut *mut Ctx) -> c_int;
    pub fn tsgpu_destroy(ctx: *mut Ctx);
    pub fn tsgpu_last_error(ctx: *const Ctx) -> *const c_char;
    pub fn tsgpu_launch_count(ctx: *const Ctx) -> u64;

    // MultilinearExtension (src/polynomials.rs:18-196)
    pub fn tsgpu_table_upload(ctx: *mut Ctx, evals: *const Fr, n: usize, num_vars: c_uint, out: *mut *mut Table) -> c_int;
    pub fn tsgpu_table_download(ctx: *mut Ctx, t: *const Table, out: *mut Fr) -> c_int;
    pub fn tsgpu_table_num_vars(t: *const Table) -> c_uint;
    pub fn tsgpu_table_free(ctx: *mut Ctx, t: *mut Table);
    pub fn tsgpu_table_one_hot(ctx: *mut Ctx, num_vars: c_uint, index: usize, out: *mut *mut Table) -> c_int;
    pub fn tsgpu_table_from_sparse(ctx: *mut Ctx, num_vars: c_uint, indices: *const u64, values: *const Fr, count: usize, out: *mut *mut Table) -> c_int;
    pub fn tsgpu_table_less_than(ctx: *mut Ctx, num_vars: c_uint, out: *mut *mut Table) -> c_int;
    pub fn tsgpu_table_add(ctx: *mut Ctx, a: *const Table, b: *const Table, out: *mut *mut Table) -> c_int;
    pub fn tsgpu_table_scalar_mul(ctx: *mut Ctx, a: *const Table, scalar: *const Fr, out: *mut *mut Table) -> c_int;
    pub fn tsgpu_table_sum_evaluations(ctx: *mut Ctx, t: *const Table, out: *mut Fr) -> c_int;
    pub fn tsgpu_mle_evaluate(ctx: *mut Ctx, evals: *const Fr, num_vars: c_uint, point: *const Fr, out: *mut Fr) -> c_int;
    pub fn tsgpu_mle_partial_evaluate(ctx: *mut Ctx, evals: *const Fr, num_vars: c_uint, fixed: *const Fr, k: c_uint, out: *mut Fr) -> c_int;

    // lookup / memory-checking building blocks (csrc/lookup.cu)
    pub fn tsgpu_table_scatter_add(ctx: *mut Ctx, weights: *const Table, idx: *const u64, n: usize, log_k: c_uint, out: *mut *mut Table) -> c_int;
    pub fn tsgpu_table_gather(ctx: *mut Ctx, src: *const Table, idx: *const u64, n: usize, num_vars: c_uint, out: *mut *mut Table) -> c_int;
    pub fn tsgpu_table_inner_product(ctx: *mut Ctx, a: *const Table, b: *const Table, out: *mut Fr) -> c_int;
    pub fn tsgpu_table_mul(ctx: *mut Ctx, a: *const Table, b: *const Table, out: *mut *mut Table) -> c_int;
    pub fn tsgpu_table_memory_values(ctx: *mut Ctx, addresses: *const u64, is_write: *const u8, values: *const Fr, n: usize, log_cells: c_uint,
                                     log_cycles: c_uint, out: *mut *mut Table) -> c_int;
    pub fn tsgpu_table_one_hot_weighted(ctx: *mut Ctx, weights: *const Table, addresses: *const u64, select: *const u8, flag: c_int, n: usize,
                                        log_cells: c_uint, out: *mut *mut Table) -> c_int;
    pub fn tsgpu_table_lt_point(ctx: *mut Ctx, point: *const Fr, num_vars: c_uint, out: *mut *mut Table) -> c_int;

    // sum-check rounds, the transcript stays with the caller (src/sumcheck.rs:56-110,156-207)
    pub fn tsgpu_sc_begin(ctx: *mut Ctx, tables: *const *mut Table, d: c_int, out: *mut *mut Sc) -> c_int;
    pub fn tsgpu_sc_num_vars(sc: *const Sc) -> c_uint;
    pub fn tsgpu_sc_exclusive(sc: *mut Sc, on: c_int) -> c_int;      // 1: nothing else is enqueued on the context during the rounds (enables the resident tail kernel)
    pub fn tsgpu_sc_round_eval(sc: *mut Sc, evals: *mut Fr) -> c_int;
    pub fn tsgpu_sc_bind(sc: *mut Sc, r: *const Fr) -> c_int;
    pub fn tsgpu_sc_bind_eval(sc: *mut Sc, r: *const Fr, evals: *mut Fr) -> c_int;
    pub fn tsgpu_sc_bind_eval_claim(sc: *mut Sc, r: *const Fr, claim: *const Fr, evals: *mut Fr) -> c_int;
    pub fn tsgpu_sc_final(sc: *mut Sc, finals: *mut Fr) -> c_int;
    pub fn tsgpu_sc_end(sc: *mut Sc);

    // SRS + KZG (src/utils.rs:89-96, src/commitments.rs:162-199)
    pub fn tsgpu_srs_generate(ctx: *mut Ctx, tau: *const Fr, n: usize, out: *mut *mut Srs) -> c_int;
    pub fn tsgpu_srs_upload(ctx: *mut Ctx, powers: *const G1Projective, n: usize, out: *mut *mut Srs) -> c_int;
    pub fn tsgpu_srs_download(ctx: *mut Ctx, srs: *const Srs, first: usize, count: usize, out: *mut G1Projective) -> c_int;
    pub fn tsgpu_srs_len(srs: *const Srs) -> usize;
    pub fn tsgpu_srs_free(ctx: *mut Ctx, srs: *mut Srs);
    pub fn tsgpu_kzg_commit(ctx: *mut Ctx, srs: *const Srs, poly: *const Fr, n: usize, out: *mut G1Projective) -> c_int;
    pub fn tsgpu_kzg_open(ctx: *mut Ctx, srs: *const Srs, poly: *const Fr, n: usize, z: *const Fr, value: *mut Fr, proof: *mut G1Projective) -> c_int;
    pub fn tsgpu_interpolate_iota(ctx: *mut Ctx, values: *const Fr, n: usize, coeffs: *mut Fr) -> c_int;
    pub fn tsgpu_msm_g1(ctx: *mut Ctx, bases_affine_xy: *const c_void, scalars: *const Fr, n: usize, out: *mut G1Projective) -> c_int;

    // whole-protocol entry points (host/protocols.cpp): setup_params, Twist / Shout prove + verify
    pub fn tsgpu_setup_params(ctx: *mut Ctx, log_size: usize, out: *mut *mut Params) -> c_int;
    pub fn tsgpu_params_free(ctx: *mut Ctx, p: *mut Params);
    pub fn tsgpu_params_max_operations(p: *const Params) -> usize;
    pub fn tsgpu_params_tau(p: *const Params, out: *mut Fr);
    pub fn tsgpu_params_fiat_shamir_seed(p: *const Params, out: *mut u8);
    pub fn tsgpu_params_srs(p: *const Params) -> *const Srs;
    pub fn tsgpu_twist_prove(ctx: *mut Ctx, params: *const Params, addresses: *const u64, values: *const Fr, is_write: *const u8,
                             num_operations: usize, out: *mut *mut Proof) -> c_int;
    pub fn tsgpu_shout_prove(ctx: *mut Ctx, params: *const Params, entries: *const Fr, num_entries: usize, lookup_indices: *const u64,
                             num_lookups: usize, out: *mut *mut Proof) -> c_int;
    pub fn tsgpu_twist_prove_sharded(ctx: *mut Ctx, params: *const Params, addresses: *const u64, values: *const Fr, num_local: usize,
                                     total_operations: usize, out: *mut *mut Proof) -> c_int;
    pub fn tsgpu_twist_prove_sharded_dev(ctx: *mut Ctx, params: *const Params, local_addresses: *mut Poly, local_values: *mut Poly,
                                         padded_operations: usize, out: *mut *mut Proof) -> c_int;
    pub fn tsgpu_comm_peer_exchange(ctx: *const Ctx) -> c_int;     // 1: peer mailboxes over NVLink (round sums inside the round kernel), 0: NCCL collectives
    pub fn tsgpu_shout_prove_sharded(ctx: *mut Ctx, params: *const Params, entries: *const Fr, num_local_entries: usize, total_entries: usize,
                                     lookup_indices: *const u64, num_local_lookups: usize, total_lookups: usize, out: *mut *mut Proof) -> c_int;
    pub fn tsgpu_twist_verify(ctx: *mut Ctx, params: *const Params, proof: *const Proof, valid: *mut c_int) -> c_int;
    pub fn tsgpu_shout_verify(ctx: *mut Ctx, params: *const Params, proof: *const Proof, valid: *mut c_int) -> c_int;
    pub fn tsgpu_proof_num_rounds(p: *const Proof) -> usize;
    pub fn tsgpu_proof_num_openings(p: *const Proof) -> usize;
    pub fn tsgpu_proof_commitment(p: *const Proof, which: c_int, out: *mut G1Projective);
    pub fn tsgpu_proof_round_polynomials(p: *const Proof, out: *mut Fr);
    pub fn tsgpu_proof_final_evaluation(p: *const Proof, out: *mut Fr);
    pub fn tsgpu_proof_opening(p: *const Proof, i: usize, proof: *mut G1Projective, value: *mut Fr);
    pub fn tsgpu_proof_bytes(p: *const Proof, out: *mut u8, capacity: usize) -> usize;
    pub fn tsgpu_proof_free(p: *mut Proof);

    // the real constraint sum-checks the reference leaves as stubs (non-parity mode; host/read_check.cpp, host/memory_check.cpp).
    // `transcript` is the library's own Transcript handle (tsgpu_transcript_new): these protocols draw their challenges inside the call.
    pub fn tsgpu_transcript_new(seed32: *const u8) -> *mut TranscriptH;
    pub fn tsgpu_transcript_free(t: *mut TranscriptH);
    pub fn tsgpu_shout_read_check_prove(ctx: *mut Ctx, entries: *const Fr, num_entries: usize, lookup_indices: *const u64, lookup_values: *const Fr,
                                        num_lookups: usize, transcript: *mut TranscriptH, claimed_sum: *mut Fr, round_polys: *mut Fr,
                                        final_evaluation: *mut Fr, challenges: *mut Fr) -> c_int;
    pub fn tsgpu_shout_read_check_verify(ctx: *mut Ctx, entries: *const Fr, num_entries: usize, lookup_indices: *const u64, lookup_values: *const Fr,
                                         num_lookups: usize, transcript: *mut TranscriptH, round_polys: *const Fr, num_rounds: usize,
                                         final_evaluation: *const Fr, valid: *mut c_int) -> c_int;
    pub fn tsgpu_twist_memory_check_prove(ctx: *mut Ctx, addresses: *const u64, values: *const Fr, is_write: *const u8, num_operations: usize,
                                          memory_size: usize, transcript: *mut TranscriptH, claims: *mut Fr, rounds1: *mut Fr, final1: *mut Fr,
                                          rounds2: *mut Fr, final2: *mut Fr) -> c_int;
    pub fn tsgpu_twist_memory_check_verify(ctx: *mut Ctx, addresses: *const u64, values: *const Fr, is_write: *const u8, num_operations: usize,
                                           memory_size: usize, transcript: *mut TranscriptH, claims: *const Fr, rounds1: *const Fr, num_rounds1: usize,
                                           final1: *const Fr, rounds2: *const Fr, num_rounds2: usize, final2: *const Fr, valid: *mut c_int) -> c_int;
    // write-checking (third sum-check of Twist) + the Val-evaluation of the claim it ends in; same transcript, after the memory check
    pub fn tsgpu_twist_write_check_prove(ctx: *mut Ctx, addresses: *const u64, values: *const Fr, is_write: *const u8, num_operations: usize,
                                         memory_size: usize, transcript: *mut TranscriptH, claims: *mut Fr, rounds3: *mut Fr, final3: *mut Fr,
                                         rounds4: *mut Fr, final4: *mut Fr) -> c_int;
    pub fn tsgpu_twist_write_check_verify(ctx: *mut Ctx, addresses: *const u64, values: *const Fr, is_write: *const u8, num_operations: usize,
                                          memory_size: usize, transcript: *mut TranscriptH, claims: *const Fr, rounds3: *const Fr, num_rounds3: usize,
                                          final3: *const Fr, rounds4: *const Fr, num_rounds4: usize, final4: *const Fr, valid: *mut c_int) -> c_int;

    // binding the constraint sum-checks to the KZG commitments of a Twist (is_shout = 0) / Shout (1) proof
    pub fn tsgpu_transcript_bind_proof(transcript: *mut TranscriptH, proof: *const Proof, is_shout: c_int) -> c_int;
    pub fn tsgpu_twist_commitments_match(ctx: *mut Ctx, params: *const Params, proof: *const Proof, addresses: *const u64, values: *const Fr,
                                         num_operations: usize, matches: *mut c_int) -> c_int;
    pub fn tsgpu_shout_commitments_match(ctx: *mut Ctx, params: *const Params, proof: *const Proof, entries: *const Fr, num_entries: usize,
                                         lookup_indices: *const u64, num_lookups: usize, matches: *mut c_int) -> c_int;

    // multi-GPU (one process per GPU; the host program carries the 128-byte NCCL id)
    pub fn tsgpu_comm_unique_id(out: *mut u8) -> c_int;
    pub fn tsgpu_comm_init(ctx: *mut Ctx, nranks: c_int, rank: c_int, id: *const u8) -> c_int;
    pub fn tsgpu_comm_allgather(ctx: *mut Ctx, input: *const c_void, bytes: usize, out: *mut c_void) -> c_int;
    pub fn tsgpu_sumcheck_prove_product_sharded(ctx: *mut Ctx, tables: *const *mut Table, d: c_int, num_vars: c_uint, claimed_sum: *const Fr,
                                                transcript: *mut TranscriptH, round_polys: *mut Fr, final_evaluation: *mut Fr,
                                                challenges: *mut Fr, table_finals: *mut Fr) -> c_int;
}

/* tsgpu.h - C ABI of the B200 (sm_100a) prover backend for the Twist & Shout hot path.
 *
 * This is the drop-in boundary: the entry points are what a Rust `cuda-sys` crate placed next to the
 * reference crate (`twist-and-shout`, /root/reference) would bind with `extern "C"` (see INTEGRATION.md
 * for the binding and for which reference function each call replaces).  Plain pointers and sizes
 * only; no C++ or torch types.
 *
 * Data layout (reference types, unchanged - no conversion at the boundary):
 *   tsgpu_fr  = ark_bn254::Fr  = [u64; 4] little-endian limbs holding a * 2^256 mod r   (src/utils.rs:14)
 *   tsgpu_g1  = G1Projective   = Jacobian {x, y, z} over Fq, Montgomery limbs, identity z = 0 (src/utils.rs:17)
 *   tsgpu_g1a = affine (x, y), Montgomery limbs, identity = all-zero
 * Host buffers passed in are caller-owned and only read for the duration of the call; results are
 * written before the call returns (calls are synchronous with respect to the host).
 *
 * Errors: every call returns 0 on success or one of the TSGPU_E_* codes, which map 1:1 onto the
 * reference's TwistAndShoutError variants (src/lib.rs:59-78); tsgpu_last_error() returns the message
 * (the reference's own strings where one exists).  Nothing unwinds across the boundary.
 * There is no CPU fallback: without a CUDA device tsgpu_init fails with TSGPU_E_PROOF_GENERATION.
 */
#ifndef TSGPU_H
#define TSGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TSGPU_ABI_VERSION 1

/* TwistAndShoutError variants, src/lib.rs:59-78 */
#define TSGPU_OK 0
#define TSGPU_E_INVALID_PARAMETERS 1
#define TSGPU_E_PROOF_GENERATION 2
#define TSGPU_E_PROOF_VERIFICATION 3
#define TSGPU_E_COMMITMENT 4
#define TSGPU_E_POLYNOMIAL 5
#define TSGPU_E_SUMCHECK 6

typedef struct { uint64_t l[4]; } tsgpu_fr;
typedef struct { uint64_t x[4], y[4], z[4]; } tsgpu_g1;
typedef struct { uint64_t x[4], y[4]; } tsgpu_g1a;

typedef struct tsgpu_ctx tsgpu_ctx;       /* one GPU, one stream, scratch memory; not re-entrant */
typedef struct tsgpu_table tsgpu_table;   /* an MLE evaluation table resident in HBM */
typedef struct tsgpu_sc tsgpu_sc;         /* a sum-check prover instance (tables being folded) */
typedef struct tsgpu_srs tsgpu_srs;       /* CommitmentParams.g1_powers resident in HBM (affine) */
typedef struct tsgpu_poly tsgpu_poly;     /* a coefficient vector resident in HBM */

int tsgpu_abi_version(void);

/* ---- context ------------------------------------------------------------------------------------ */
/* `stream` is a cudaStream_t the caller owns (e.g. torch's current stream) or NULL to create one. */
int tsgpu_init(int device, void* stream, tsgpu_ctx** out);
void tsgpu_destroy(tsgpu_ctx* ctx);
const char* tsgpu_last_error(const tsgpu_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's `gpu_launches`) */
uint64_t tsgpu_launch_count(const tsgpu_ctx* ctx);
int tsgpu_synchronize(tsgpu_ctx* ctx);
int tsgpu_sm_count(const tsgpu_ctx* ctx);
/* work counters since tsgpu_init: "launches", "msm_calls", "msm_points", "msm_entries" (non-zero digits = bucket additions) */
uint64_t tsgpu_counter_read(const tsgpu_ctx* ctx, const char* name);
/* tuning knobs: "prefetch_min_log2" (log2 of the positions per launch from which the d = 2 sum-check rounds use the warp-private TMA prefetch
 * kernels; default 21), "deferred_claim_check" (see tsgpu_sumcheck_prove_product), "peer_exchange" (sharded paths: peer mailboxes over NVLink
 * or NCCL collectives), "eval_basis" (0: Twist / Shout::prove interpolate and commit coefficients), "msm_tables" (0: per-window bucket sets),
 * "sc_tail" (d = 2 claim-form rounds on tables of at most 2^11 entries run in ONE persistent kernel that exchanges round values and challenges with the
 * host through a mapped mailbox; default 1), "kernel_timing" (per-phase CUDA-event timers behind tsgpu_timer_read) */
int tsgpu_set_tuning(tsgpu_ctx* ctx, const char* key, long value);
/* with tuning key "kernel_timing" = 1 the library brackets its main kernels with CUDA events on the context
 * stream; names: "msm_accumulate", "msm_total", "interpolate", "open_scan", "sc_round_eval", "sc_bind_eval", "bind" */
int tsgpu_timer_read(tsgpu_ctx* ctx, const char* name, double* total_ms, uint64_t* count);
void tsgpu_timer_reset(tsgpu_ctx* ctx);

/* ---- MLE tables: MultilinearExtension { num_vars, evaluations }  (src/polynomials.rs:18-82) ------ */
/* from_evaluations / from_evaluations_vec: `n` host entries, zero-padded or truncated to 2^num_vars
 * (polynomials.rs:40-50).  Uploaded once; lives in HBM until freed. */
int tsgpu_table_upload(tsgpu_ctx* ctx, const tsgpu_fr* evals, size_t n, unsigned num_vars, tsgpu_table** out);
/* evaluations back to the host in reference index order (2^num_vars entries) */
int tsgpu_table_download(tsgpu_ctx* ctx, const tsgpu_table* t, tsgpu_fr* out);
int tsgpu_table_clone(tsgpu_ctx* ctx, const tsgpu_table* t, tsgpu_table** out);
unsigned tsgpu_table_num_vars(const tsgpu_table* t);
void tsgpu_table_free(tsgpu_ctx* ctx, tsgpu_table* t);
/* device-side generators for the tables the protocols are built from:
 *   eq(w, .)[i] = prod_j (bit_j(i) ? w_j : 1 - w_j)      - the basis polynomial of polynomials.rs:108-122
 *   one-hot matrix: entry [row * 2^log_k + idx[row]] = 1  - MultilinearExtension::one_hot (polynomials.rs:71-82) per row
 *   from_u64: entry i = Fr::from(v[i])                    - twist.rs:119 / shout.rs:112 */
int tsgpu_table_eq(tsgpu_ctx* ctx, const tsgpu_fr* w, unsigned num_vars, tsgpu_table** out);
int tsgpu_table_one_hot_rows(tsgpu_ctx* ctx, const uint64_t* idx, size_t rows, unsigned log_k, unsigned num_vars, tsgpu_table** out);
int tsgpu_table_from_u64(tsgpu_ctx* ctx, const uint64_t* v, size_t n, unsigned num_vars, tsgpu_table** out);
/* MultilinearExtension::one_hot(num_vars, index) (src/polynomials.rs:71-82).  index >= 2^num_vars panics in the reference
 * ("Index {i} out of bounds for size {n}"): TSGPU_E_POLYNOMIAL with that message here. */
int tsgpu_table_one_hot(tsgpu_ctx* ctx, unsigned num_vars, size_t index, tsgpu_table** out);
/* MultilinearExtension::from_sparse(num_vars, &[(index, value)]) (src/polynomials.rs:52-67): a repeated index keeps its last value;
 * the same out-of-bounds behaviour as one_hot. */
int tsgpu_table_from_sparse(tsgpu_ctx* ctx, unsigned num_vars, const uint64_t* indices, const tsgpu_fr* values, size_t count, tsgpu_table** out);
/* LessThanPolynomial::new(num_vars).to_multilinear_extension() (src/polynomials.rs:243-263): 2 * num_vars variables, entry at
 * reference index a | (b << num_vars) = lt(a, b), decided by the first differing bit counted from bit 0 (src/polynomials.rs:222-239). */
int tsgpu_table_less_than(tsgpu_ctx* ctx, unsigned num_vars, tsgpu_table** out);
/* LessThanPolynomial::evaluate_at_field_elements (src/polynomials.rs:213-220): the low num_vars bits of the canonical integers.  Host. */
void tsgpu_lt_evaluate_at_field_elements(unsigned num_vars, const tsgpu_fr* a, const tsgpu_fr* b, tsgpu_fr* out);
/* MultilinearExtension::add / scalar_mul / sum_evaluations (src/polynomials.rs:164-195).  add: TSGPU_E_POLYNOMIAL
 * "Number of variables must match" where the reference asserts. */
int tsgpu_table_add(tsgpu_ctx* ctx, const tsgpu_table* a, const tsgpu_table* b, tsgpu_table** out);
int tsgpu_table_scalar_mul(tsgpu_ctx* ctx, const tsgpu_table* a, const tsgpu_fr* scalar, tsgpu_table** out);
int tsgpu_table_sum_evaluations(tsgpu_ctx* ctx, const tsgpu_table* t, tsgpu_fr* out);
/* Lookup-argument building blocks: the one-hot matrix ra(x, j) = [idx[j] == x] of Shout (src/polynomials.rs:71-82 row by row) applied
 * to a vector without materialising its 2^log_k x n entries.
 *   scatter_add: out[x] = sum over j < n with idx[j] == x of weights[j]   (ra~(., r) for weights = eq(r, .))
 *   gather:      out[j] = src[idx[j]] for j < n, zero up to 2^num_vars
 *   inner_product: field_utils::inner_product (src/utils.rs:210-213) of two tables of equal size
 * An index >= the table size is TSGPU_E_INVALID_PARAMETERS "Lookup index out of bounds" (src/shout.rs:44-48). */
int tsgpu_table_scatter_add(tsgpu_ctx* ctx, const tsgpu_table* weights, const uint64_t* idx, size_t n, unsigned log_k, tsgpu_table** out);
int tsgpu_table_gather(tsgpu_ctx* ctx, const tsgpu_table* src, const uint64_t* idx, size_t n, unsigned num_vars, tsgpu_table** out);
int tsgpu_table_inner_product(tsgpu_ctx* ctx, const tsgpu_table* a, const tsgpu_table* b, tsgpu_fr* out);

/* ---- MultilinearExtension::evaluate / partial_evaluate  (src/polynomials.rs:85-161) -------------- */
/* host-buffer forms (copy in, compute, copy out) */
int tsgpu_mle_evaluate(tsgpu_ctx* ctx, const tsgpu_fr* evals, unsigned num_vars, const tsgpu_fr* point, tsgpu_fr* out);
int tsgpu_mle_partial_evaluate(tsgpu_ctx* ctx, const tsgpu_fr* evals, unsigned num_vars, const tsgpu_fr* fixed, unsigned k, tsgpu_fr* out);
/* HBM-resident forms */
int tsgpu_table_evaluate(tsgpu_ctx* ctx, const tsgpu_table* t, const tsgpu_fr* point, tsgpu_fr* out);
int tsgpu_table_partial_evaluate(tsgpu_ctx* ctx, const tsgpu_table* t, const tsgpu_fr* fixed, unsigned k, tsgpu_table** out);
/* one fold: T'[i] = T[2i] + r (T[2i+1] - T[2i]) in place (num_vars decreases by one) */
int tsgpu_table_bind(tsgpu_ctx* ctx, tsgpu_table* t, const tsgpu_fr* r);

/* ---- sum-check prover round  (SumCheck::prove / compute_round_polynomial, src/sumcheck.rs:56-110,156-207)
 * Structured sibling of the closure-typed SumCheck::prove for f(v) = prod_{t<d} mle_t.evaluate(v), d in 1..3.
 * Round-stepped so the Fiat-Shamir transcript stays with the caller:
 *     begin -> round_eval -> [transcript] -> bind_eval(r_0) -> [transcript] -> ... -> bind_eval(r_{n-2})
 *           -> [transcript] -> final(r_{n-1})
 * The tables are consumed (folded in place). */
int tsgpu_sc_begin(tsgpu_ctx* ctx, tsgpu_table* const* tables, int d, tsgpu_sc** out);
unsigned tsgpu_sc_num_vars(const tsgpu_sc* sc);   /* variables still unbound */
/* on = 1: the caller promises to enqueue nothing else on this context until the rounds of `sc` end (tsgpu_sc_final / tsgpu_sc_end).  The d = 2 claim-form
 * rounds on tables of at most 2^11 entries then run inside ONE persistent kernel that stays resident between the calls and trades round values for challenges
 * through a mapped mailbox (a round costs a PCIe round trip instead of a launch + stream synchronisation); any other tsgpu_sc_* call makes it hand the tables
 * back first.  The library's own SumCheck::prove loops set it; default off: interleaving several handles on one context stays legal. */
int tsgpu_sc_exclusive(tsgpu_sc* sc, int on);
/* g(0), g(1), g(2), g(3) of the current round (sumcheck.rs:175-198) */
int tsgpu_sc_round_eval(tsgpu_sc* sc, tsgpu_fr evals[4]);
/* the same four values for d = 2 when g(0) + g(1) = claim is vouched for by the caller, who then owes the check of the claim itself (the library's
 * SumCheck::prove loop pays it at the end, see tsgpu_sumcheck_prove_product): two products per pair instead of three.  d != 2: plain evaluation. */
int tsgpu_sc_round_eval_claim(tsgpu_sc* sc, const tsgpu_fr* claim, tsgpu_fr evals[4]);
/* bind the current variable to r (sumcheck.rs:99) */
int tsgpu_sc_bind(tsgpu_sc* sc, const tsgpu_fr* r);
/* fused: bind to r, then evaluate the next round's g(0..3) in the same pass */
int tsgpu_sc_bind_eval(tsgpu_sc* sc, const tsgpu_fr* r, tsgpu_fr evals[4]);
/* same with the running claim g_k(r) of SumCheck::prove (src/sumcheck.rs:86-100) supplied: g(1) = claim - g(0) is derived, not summed */
int tsgpu_sc_bind_eval_claim(tsgpu_sc* sc, const tsgpu_fr* r, const tsgpu_fr* claim, tsgpu_fr evals[4]);
/* after all variables are bound: the d table values mle_t(r_0..r_{n-1}); their product is
 * SumCheckProof.final_evaluation (sumcheck.rs:104) */
int tsgpu_sc_final(tsgpu_sc* sc, tsgpu_fr* finals);
void tsgpu_sc_end(tsgpu_sc* sc);

/* ---- SRS: CommitmentParams.g1_powers resident in HBM  (src/utils.rs:53-63, 89-96) ---------------------------
 * generate: g1_powers[i] = G * tau^i for i < n, what the setup_params loop computes with n = max_degree + 1
 *           (fixed-base byte windows on the device instead of n serial double-and-add multiplications).
 * upload:   an existing Vec<G1Projective> (Jacobian) from the host; normalised to affine once on the device.
 * download: back to the host as Jacobian points with z = 1 (the same group elements). */
int tsgpu_srs_generate(tsgpu_ctx* ctx, const tsgpu_fr* tau, size_t n, tsgpu_srs** out);
int tsgpu_srs_upload(tsgpu_ctx* ctx, const tsgpu_g1* powers, size_t n, tsgpu_srs** out);
/* powers first .. first + n - 1 only: the slice a point-sharded MSM rank holds */
int tsgpu_srs_generate_range(tsgpu_ctx* ctx, const tsgpu_fr* tau, size_t first, size_t n, tsgpu_srs** out);
int tsgpu_srs_download(tsgpu_ctx* ctx, const tsgpu_srs* srs, size_t first, size_t count, tsgpu_g1* out);
size_t tsgpu_srs_len(const tsgpu_srs* srs);
void tsgpu_srs_free(tsgpu_ctx* ctx, tsgpu_srs* srs);

/* ---- coefficient vectors resident in HBM (low -> high, reference order) ---------------------------------- */
int tsgpu_poly_upload(tsgpu_ctx* ctx, const tsgpu_fr* coeffs, size_t n, tsgpu_poly** out);
int tsgpu_poly_download(tsgpu_ctx* ctx, const tsgpu_poly* p, tsgpu_fr* out);
size_t tsgpu_poly_len(const tsgpu_poly* p);
int tsgpu_poly_clone(tsgpu_ctx* ctx, const tsgpu_poly* p, tsgpu_poly** out);
/* padded vectors as Twist/Shout::prove build them: Fr::from(v[i]) / given values, zero-filled up to `padded`
 * (src/twist.rs:115-148, src/shout.rs:105-118) */
int tsgpu_poly_from_u64(tsgpu_ctx* ctx, const uint64_t* v, size_t n, size_t padded, tsgpu_poly** out);
int tsgpu_poly_upload_padded(tsgpu_ctx* ctx, const tsgpu_fr* vals, size_t n, size_t padded, tsgpu_poly** out);
/* same vector, but the host -> device copy runs on a side stream of the context and the call returns without waiting: work enqueued on the context
 * afterwards overlaps the transfer (tsgpu_twist_prove commits the address vector while the 32 n bytes of values travel).  No other entry point waits for
 * the copy by itself: tsgpu_poly_wait orders the context's stream behind it (a no-op for vectors that are not in flight); tsgpu_poly_free may be called at
 * any time.  `vals` must stay valid until then; pinned memory is what makes the copy asynchronous. */
int tsgpu_poly_upload_padded_async(tsgpu_ctx* ctx, const tsgpu_fr* vals, size_t n, size_t padded, tsgpu_poly** out);
int tsgpu_poly_wait(tsgpu_ctx* ctx, tsgpu_poly* p);
int tsgpu_poly_in_flight(const tsgpu_poly* p);

/* ---- vector_to_polynomial: poly_utils::lagrange_interpolate over x_i = i  (src/twist.rs:307-315,
 * src/shout.rs:277-285, src/polynomials.rs:301-352) --------------------------------------------------------
 * Same (unique) coefficients in O(n log^2 n): Newton differences as a convolution + falling-factorial ->
 * monomial conversion, NTTs over Fr.  n must be a power of two <= 2^27 (Twist/Shout pad to one);
 * otherwise TSGPU_E_POLYNOMIAL.  prepare() builds the size-dependent tables once (done implicitly on first use). */
int tsgpu_interpolate_prepare(tsgpu_ctx* ctx, unsigned log_n);
int tsgpu_interpolate_iota(tsgpu_ctx* ctx, const tsgpu_fr* values, size_t n, tsgpu_fr* coeffs);   /* any n (host buffers) */
int tsgpu_poly_interpolate_iota(tsgpu_ctx* ctx, tsgpu_poly* values_to_coeffs_in_place);
void tsgpu_poly_free(tsgpu_ctx* ctx, tsgpu_poly* p);

/* ---- CommitmentScheme for KZGCommitment  (src/commitments.rs:156-199) -------------------------------------
 * commit: C = sum_i polynomial[i] * g1_powers[i]  (Pippenger MSM); TSGPU_E_COMMITMENT
 *         "Polynomial degree exceeds setup size" when n > srs length (commitments.rs:166-170).
 * open:   value = P(z) (Horner, :305-313), proof = commit((P - value) / (x - z)) (:317-375, :194).
 * Results are G1Projective values: any Jacobian representative of the group element. */
int tsgpu_kzg_commit(tsgpu_ctx* ctx, const tsgpu_srs* srs, const tsgpu_fr* polynomial, size_t n, tsgpu_g1* out);
int tsgpu_kzg_open(tsgpu_ctx* ctx, const tsgpu_srs* srs, const tsgpu_fr* polynomial, size_t n, const tsgpu_fr* z,
                   tsgpu_fr* value, tsgpu_g1* proof);
int tsgpu_kzg_commit_dev(tsgpu_ctx* ctx, const tsgpu_srs* srs, const tsgpu_poly* polynomial, tsgpu_g1* out);
int tsgpu_kzg_open_dev(tsgpu_ctx* ctx, const tsgpu_srs* srs, const tsgpu_poly* polynomial, const tsgpu_fr* z,
                       tsgpu_fr* value, tsgpu_g1* proof);
/* ---- evaluation-basis form of vector_to_polynomial + commit / open  (src/twist.rs:151-160,226-243, 307-315) ----
 * Twist/Shout::prove only ever commit to interpolants of a vector on the nodes 0..m-1.  The commitment is linear in
 * the VALUES: commit(P) = sum_j v_j * [L_j(tau)]_1.  lagrange_prepare builds those basis points for a power-of-two m
 * from the trapdoor that setup_params holds (CommitmentParams.tau, src/utils.rs:84,107); an SRS made by
 * tsgpu_srs_upload has no trapdoor (can_lagrange == 0) and callers stay on interpolate + commit.
 * commit_values: one MSM over the raw values (same group element as interpolate + commit).
 * open_values:   value = P(z) by the barycentric formula, proof = sum_j Q(j) [L_j(tau)]_1 with
 *                Q(j) = (v_j - value) / (j - z); TSGPU_E_POLYNOMIAL if z is one of the nodes. */
int tsgpu_srs_can_lagrange(const tsgpu_srs* srs);
int tsgpu_srs_has_lagrange(const tsgpu_srs* srs, size_t m);
int tsgpu_srs_lagrange_prepare(tsgpu_ctx* ctx, const tsgpu_srs* srs, size_t m);
int tsgpu_kzg_commit_values_dev(tsgpu_ctx* ctx, const tsgpu_srs* srs, const tsgpu_poly* values, tsgpu_g1* out);
int tsgpu_kzg_open_values_dev(tsgpu_ctx* ctx, const tsgpu_srs* srs, const tsgpu_poly* values, const tsgpu_fr* z,
                              tsgpu_fr* value, tsgpu_g1* proof);
/* ---- batched forms: `count` (<= 4) commitments / openings in ONE MSM pass.  Twist::prove and Shout::prove commit to two
 * vectors and open both at the same point (src/twist.rs:151-160,226-243): the bucket sets of the MSMs are laid side by
 * side so that sorting, accumulation, chunk merge and window reduction run once over the union.  Results are the same
 * group / field elements as `count` separate calls. */
int tsgpu_kzg_commit_batch_dev(tsgpu_ctx* ctx, const tsgpu_srs* srs, const tsgpu_poly* const* polynomials, size_t count, tsgpu_g1* outs);
int tsgpu_kzg_open_batch_dev(tsgpu_ctx* ctx, const tsgpu_srs* srs, const tsgpu_poly* const* polynomials, size_t count, const tsgpu_fr* z,
                             tsgpu_fr* values, tsgpu_g1* proofs);
int tsgpu_kzg_commit_values_batch_dev(tsgpu_ctx* ctx, const tsgpu_srs* srs, const tsgpu_poly* const* values, size_t count, tsgpu_g1* outs);
int tsgpu_kzg_open_values_batch_dev(tsgpu_ctx* ctx, const tsgpu_srs* srs, const tsgpu_poly* const* values, size_t count, const tsgpu_fr* z,
                                    tsgpu_fr* out_values, tsgpu_g1* proofs);
/* ---- VectorCommitmentScheme for KZGVectorCommitment  (src/commitments.rs:378-483) ------------------------------
 * commit: KZG commitment to lagrange_interpolate((i, vector[i])) - any length, degree < len.
 * open:   (vector[index], opening proof at the point Fr::from(index)); TSGPU_E_COMMITMENT "Index out of bounds".
 * verify: the KZG pairing check at that point (CPU). */
int tsgpu_vector_commit(tsgpu_ctx* ctx, const tsgpu_srs* srs, const tsgpu_fr* vector, size_t n, tsgpu_g1* out);
int tsgpu_vector_open(tsgpu_ctx* ctx, const tsgpu_srs* srs, const tsgpu_fr* vector, size_t n, size_t index, tsgpu_fr* value, tsgpu_g1* proof);
/* ---- sharded evaluation-basis commit / open: this rank holds the nodes [first, first + count) of the m-node domain --------
 * (count = length of the value slices).  commit: partial commitments, to be added over the ranks.  open, phase 1: out[0] = product
 * of this rank's (z - j), out[1 + i] = this rank's partial barycentric sum of slice i; the caller multiplies the products of all ranks
 * (N(z)), adds the partial sums and forms value_i = N(z) * sum_i.  open, phase 2: partial quotient commitments for those values. */
int tsgpu_srs_lagrange_prepare_range(tsgpu_ctx* ctx, const tsgpu_srs* srs, size_t m, size_t first, size_t count);
int tsgpu_kzg_commit_values_slice_batch_dev(tsgpu_ctx* ctx, const tsgpu_srs* srs, size_t m, size_t first, const tsgpu_poly* const* slices, size_t k,
                                            tsgpu_g1* partials);
int tsgpu_kzg_open_values_slice_partial(tsgpu_ctx* ctx, size_t m, size_t first, const tsgpu_poly* const* slices, size_t k, const tsgpu_fr* z, tsgpu_fr* out);
int tsgpu_kzg_open_values_slice_finish(tsgpu_ctx* ctx, const tsgpu_srs* srs, size_t m, size_t first, const tsgpu_poly* const* slices, size_t k,
                                       const tsgpu_fr* values, tsgpu_g1* partial_proofs);
/* plain G1 MSM over caller-supplied affine bases: sum_i scalars[i] * bases[i] */
int tsgpu_msm_g1(tsgpu_ctx* ctx, const tsgpu_g1a* bases, const tsgpu_fr* scalars, size_t n, tsgpu_g1* out);
/* CPU helpers on single points: KZGCommitmentValue::hash (commitments.rs:73-84), ark-serialize compressed bytes
 * (commitments.rs:106-118), group equality (derive(PartialEq) on KZGCommitmentValue, commitments.rs:66) */
void tsgpu_g1_hash(const tsgpu_g1* p, tsgpu_fr* out);
void tsgpu_g1_compress(const tsgpu_g1* p, uint8_t out[32]);
int tsgpu_g1_equal(const tsgpu_g1* a, const tsgpu_g1* b);
void tsgpu_g1_add(const tsgpu_g1* a, const tsgpu_g1* b, tsgpu_g1* out);

/* ---- host side of the path: Transcript and the SumCheck::prove / verify loops --------------------------
 * These run on the CPU (the Fiat-Shamir transcript stays on the host) and drive the round kernels above.
 * A Rust host would keep its own Transcript (src/utils.rs:134-204) and call the tsgpu_sc_* entry points;
 * these C entry points exist so that non-Rust callers (the parity tests, bench.py) use the same logic. */
typedef struct tsgpu_transcript tsgpu_transcript;
tsgpu_transcript* tsgpu_transcript_new(const uint8_t* seed32);   /* Transcript::new - the seed has no effect (utils.rs:190) */
void tsgpu_transcript_free(tsgpu_transcript* t);
void tsgpu_transcript_append(tsgpu_transcript* t, const char* label, size_t label_len, const tsgpu_fr* elems, size_t n);
void tsgpu_transcript_challenge(tsgpu_transcript* t, const char* label, size_t label_len, tsgpu_fr* out);
size_t tsgpu_transcript_state_len(const tsgpu_transcript* t);

/* SumCheck::new(num_vars, claimed_sum).prove(|v| prod_t mle_t.evaluate(v), transcript)   (sumcheck.rs:56-110)
 * round_polys: num_vars x 4 coefficients (low -> high); challenges / table_finals may be NULL.
 * Returns TSGPU_E_SUMCHECK with "Round {k} consistency check failed" when claimed_sum is wrong.
 * Default: the reference's deterministic check - round 0 is summed in full and g(0) + g(1) == claimed_sum is tested before anything is
 * appended to the transcript (sumcheck.rs:77-84); on that error the tables are unmodified.
 * Opt-in, tsgpu_set_tuning(ctx, "deferred_claim_check", 1), d = 2 only: round 0 also runs in the claim form (one product per pair less in
 * the largest round) and the claimed sum is checked against the product of the bound tables at the end.  Same proofs; a wrong claim gives
 * the same error and the transcript is rolled back to its state at entry, but it is detected only after all rounds ran (except with
 * probability ~ 2 num_vars / |Fr|) and the tables ARE consumed on that error path. */
int tsgpu_sumcheck_prove_product(tsgpu_ctx* ctx, tsgpu_table* const* tables, int d, const tsgpu_fr* claimed_sum,
                                 tsgpu_transcript* transcript, tsgpu_fr* round_polys, tsgpu_fr* final_evaluation,
                                 tsgpu_fr* challenges, tsgpu_fr* table_finals);
/* SumCheck::verify (sumcheck.rs:113-153): *valid = 1/0; TSGPU_E_SUMCHECK for a wrong number of rounds */
int tsgpu_sumcheck_verify(unsigned num_vars, const tsgpu_fr* claimed_sum, const tsgpu_fr* round_polys, size_t num_rounds,
                          const tsgpu_fr* final_evaluation, tsgpu_transcript* transcript, int* valid, tsgpu_fr* challenges);

/* CPU helpers for sharded (one process per GPU) provers, see multilinear-map-cryptography_b200/distributed.py:
 * round coefficients from g(0..3) (sumcheck.rs:201-205), Horner (utils.rs:217-221), field add / mul, and the
 * conversion of an integer all-reduce result (8 sums of zero-extended 32-bit limbs per element) back to Fr */
void tsgpu_sumcheck_round_coeffs(const tsgpu_fr evals[4], tsgpu_fr coeffs[4]);
void tsgpu_horner_eval(const tsgpu_fr* coeffs, size_t n, const tsgpu_fr* x, tsgpu_fr* out);
void tsgpu_fr_add(const tsgpu_fr* a, const tsgpu_fr* b, tsgpu_fr* out);
void tsgpu_fr_mul(const tsgpu_fr* a, const tsgpu_fr* b, tsgpu_fr* out);
void tsgpu_fr_from_limb_sums(const uint64_t* sums, size_t n, tsgpu_fr* out);

/* ---- multi-GPU: one process per GPU, one NCCL communicator per context ---------------------------------
 * The host program moves the 128-byte unique id from rank 0 to every rank (torch.distributed, MPI, a file ...).  NCCL is
 * loaded at run time; a single rank needs neither NCCL nor an id.  The number of ranks must be a power of two. */
int tsgpu_comm_unique_id(uint8_t out[128]);
int tsgpu_comm_init(tsgpu_ctx* ctx, int nranks, int rank, const uint8_t id[128]);
int tsgpu_comm_size(const tsgpu_ctx* ctx);
/* 1 when the ranks exchange over peer-mapped mailboxes (CUDA IPC over NVLink / NVSwitch): the sharded sum-check sums its round values inside the
 * round kernel and the small all-gathers are one single-block kernel; 0: NCCL collectives (no peer access, or tsgpu_set_tuning("peer_exchange", 0)) */
int tsgpu_comm_peer_exchange(const tsgpu_ctx* ctx);
int tsgpu_comm_rank(const tsgpu_ctx* ctx);
void tsgpu_comm_destroy(tsgpu_ctx* ctx);
/* all-gather of `bytes` (multiple of 8) per rank, host to host: e.g. per-rank partial MSM results (point-sliced commitment),
 * which the caller adds with tsgpu_g1_add */
int tsgpu_comm_allgather(tsgpu_ctx* ctx, const void* in, size_t bytes, void* out);
/* MultilinearExtension::evaluate of a table sliced over the ranks (`local`: this rank's slice, num_vars - log2(ranks) variables);
 * one local pass + an all-reduce of one field element; every rank gets the value */
int tsgpu_table_evaluate_sharded(tsgpu_ctx* ctx, const tsgpu_table* local, unsigned num_vars, const tsgpu_fr* point, tsgpu_fr* out);
/* SumCheck::prove for a product of tables (src/sumcheck.rs:56-110) with the hypercube sliced over the ranks: `tables` are this
 * rank's slices (reference index high bits = rank; num_vars - log2(ranks) variables; consumed).  One 256-byte integer
 * all-reduce of the round evaluations per round; every rank returns the same proof.  Outputs as tsgpu_sumcheck_prove_product. */
int tsgpu_sumcheck_prove_product_sharded(tsgpu_ctx* ctx, tsgpu_table* const* tables, int d, unsigned num_vars, const tsgpu_fr* claimed_sum,
                                         tsgpu_transcript* transcript, tsgpu_fr* round_polys, tsgpu_fr* final_evaluation,
                                         tsgpu_fr* challenges, tsgpu_fr* table_finals);

/* ---- setup_params, Twist::prove / verify, Shout::prove / verify  (host orchestration over the calls above) --
 * setup_params(log_size) (src/utils.rs:79-131): max_operations = 4 * 2^log_size, tau = first Fr::rand of
 * ChaCha20Rng::from_seed([42; 32]), g1_powers[0 ..= max_operations] generated on the device, 32-byte
 * Fiat-Shamir seed from the same rng. */
typedef struct tsgpu_params tsgpu_params;
typedef struct tsgpu_proof tsgpu_proof;    /* TwistProof / ShoutProof: 2 commitments, SumCheckProof, 0|2 openings, 0|2 evaluations */
int tsgpu_setup_params(tsgpu_ctx* ctx, size_t log_size, tsgpu_params** out);
/* VerifierParams only (log_size, max_operations, commitment_vk, fiat_shamir_seed): CPU, no context needed */
int tsgpu_setup_verifier_params(size_t log_size, tsgpu_params** out);
void tsgpu_params_free(tsgpu_ctx* ctx, tsgpu_params* p);
size_t tsgpu_params_log_size(const tsgpu_params* p);
size_t tsgpu_params_max_operations(const tsgpu_params* p);
void tsgpu_params_tau(const tsgpu_params* p, tsgpu_fr* out);
void tsgpu_params_fiat_shamir_seed(const tsgpu_params* p, uint8_t out[32]);
const tsgpu_srs* tsgpu_params_srs(const tsgpu_params* p);

/* Twist::prove(&MemoryTrace) (src/twist.rs:107-252).  operations[i] = {address: addresses[i], value: values[i]},
 * is_write[i] != 0 for MemoryOp::Write.  TSGPU_E_INVALID_PARAMETERS "Too many operations" beyond max_operations. */
int tsgpu_twist_prove(tsgpu_ctx* ctx, const tsgpu_params* params, const uint64_t* addresses, const tsgpu_fr* values,
                      const uint8_t* is_write, size_t num_operations, tsgpu_proof** out);
/* same with the two zero-padded vectors already in HBM (consumed) */
int tsgpu_twist_prove_dev(tsgpu_ctx* ctx, const tsgpu_params* params, tsgpu_poly* padded_addresses, tsgpu_poly* padded_values, tsgpu_proof** out);
/* ONE proof sharded over the ranks of the context's communicator (tsgpu_comm_init): rank r passes the operations of the padded
 * positions [r m / G, (r + 1) m / G), m = next_power_of_two(total_operations) - num_local of them exist.  Three small all-gathers
 * (partial commitments, opening sums, partial opening proofs); every rank returns the same proof, byte-identical to tsgpu_twist_prove. */
int tsgpu_twist_prove_sharded(tsgpu_ctx* ctx, const tsgpu_params* params, const uint64_t* addresses, const tsgpu_fr* values, size_t num_local,
                              size_t total_operations, tsgpu_proof** out);
/* same with this rank's zero-padded slices (padded_operations / ranks entries each) already resident in HBM (not consumed) */
int tsgpu_twist_prove_sharded_dev(tsgpu_ctx* ctx, const tsgpu_params* params, tsgpu_poly* local_addresses, tsgpu_poly* local_values,
                                  size_t padded_operations, tsgpu_proof** out);
/* Shout::prove(&LookupTable) (src/shout.rs:97-222): entries = table.entries, lookup_indices[i] = lookups[i].index.
 * TSGPU_E_INVALID_PARAMETERS "Too many lookup operations" beyond max_operations. */
int tsgpu_shout_prove(tsgpu_ctx* ctx, const tsgpu_params* params, const tsgpu_fr* entries, size_t num_entries,
                      const uint64_t* lookup_indices, size_t num_lookups, tsgpu_proof** out);
/* ONE Shout proof sharded over the ranks (BASELINE config 3 at 2/4/8 GPUs): rank r passes the table entries of the padded positions
 * [r mt / G, (r + 1) mt / G), mt = next_power_of_two(total_entries), and the lookup indices of [r ml / G, (r + 1) ml / G),
 * ml = next_power_of_two(total_lookups).  Same all-gathers as the Twist form (two more when mt != ml: the two openings then run one
 * after the other); every rank returns the same proof, byte-identical to tsgpu_shout_prove. */
int tsgpu_shout_prove_sharded(tsgpu_ctx* ctx, const tsgpu_params* params, const tsgpu_fr* entries, size_t num_local_entries, size_t total_entries,
                              const uint64_t* lookup_indices, size_t num_local_lookups, size_t total_lookups, tsgpu_proof** out);
/* Read/write memory (Twist) tables over (cell x, cycle j), reference index x + 2^log_cells * j - the layout of BASELINE config 4.
 *   memory_values:   Val(x, j) = content of cell x just before operation j (MemoryTrace semantics, src/twist.rs:48-70; zero-initialised memory)
 *   one_hot_weighted: out[addresses[j] + 2^log_cells j] = weights[j] for the operations with select[j] == flag (weights: 2^log_cycles entries)
 *   mul:             elementwise product;   lt_point: out[a] = LT~(a, point), [a < c] in the natural integer order, multilinear in c
 * An address >= 2^log_cells is TSGPU_E_INVALID_PARAMETERS "Address out of bounds" (src/twist.rs:49-53). */
int tsgpu_table_memory_values(tsgpu_ctx* ctx, const uint64_t* addresses, const uint8_t* is_write, const tsgpu_fr* values, size_t n, unsigned log_cells,
                              unsigned log_cycles, tsgpu_table** out);
int tsgpu_table_one_hot_weighted(tsgpu_ctx* ctx, const tsgpu_table* weights, const uint64_t* addresses, const uint8_t* select, int flag, size_t n,
                                 unsigned log_cells, tsgpu_table** out);
int tsgpu_table_mul(tsgpu_ctx* ctx, const tsgpu_table* a, const tsgpu_table* b, tsgpu_table** out);
/* table(x, j) <- rows(j) - table(x, j) in place (rows: the cycle variables only; table: cells x cycles) - wv(j) - Val(x, j) of Twist write-checking */
int tsgpu_table_broadcast_rows_minus(tsgpu_ctx* ctx, const tsgpu_table* rows, tsgpu_table* table);
int tsgpu_table_lt_point(tsgpu_ctx* ctx, const tsgpu_fr* point, unsigned num_vars, tsgpu_table** out);

/* ---- The lookup-correctness sum-check the reference leaves as a stub (src/shout.rs:157-184: the closure returns zero on every branch
 * and says "In a production implementation, this would involve more complex constraints").  NOT part of the reference's proofs - a
 * separate, explicitly non-parity mode; Shout::prove above stays byte-identical to the reference.
 * Statement: lookup j < num_lookups reads entries[lookup_indices[j]] and returns lookup_values[j] (LookupOp { index, value }).
 * Protocol (core Shout read-checking): the statement is bound first - a BLAKE2b-256 tree digest of (sizes, entries, indices, values)
 * (host/statement_digest.hpp) is appended as two field elements under "read_check_statement", so that every challenge depends on the
 * whole statement; then r = transcript.challenge_field_elements("read_check_point", log2 L) with L the padded number of lookups; claim = rv~(r), the multilinear extension of the returned values, appended as "read_check_claim"; then
 * SumCheck::new(log2 K, claim).prove(|x| ra~(x, r) * Val~(x)) (src/sumcheck.rs:56-110, same labels) over the K padded table entries,
 * with ra~(x, r) = sum_j eq(r, j) [idx_j == x] built by scatter_add and never as a K x L matrix.  A wrong returned value makes the
 * prover fail with the reference's own error, TSGPU_E_SUMCHECK "Round 0 consistency check failed".
 * round_polys: log2 K x 4 coefficients; challenges (optional): log2 K elements.
 * verify: replays the transcript, SumCheck::verify, and closes with final_evaluation == ra~(x*, r) * Val~(x*) computed from the
 * statement (gather + inner product + MultilinearExtension::evaluate on the device). */
int tsgpu_shout_read_check_prove(tsgpu_ctx* ctx, const tsgpu_fr* entries, size_t num_entries, const uint64_t* lookup_indices,
                                 const tsgpu_fr* lookup_values, size_t num_lookups, tsgpu_transcript* transcript,
                                 tsgpu_fr* claimed_sum, tsgpu_fr* round_polys, tsgpu_fr* final_evaluation, tsgpu_fr* challenges);
int tsgpu_shout_read_check_verify(tsgpu_ctx* ctx, const tsgpu_fr* entries, size_t num_entries, const uint64_t* lookup_indices,
                                  const tsgpu_fr* lookup_values, size_t num_lookups, tsgpu_transcript* transcript,
                                  const tsgpu_fr* round_polys, size_t num_rounds, const tsgpu_fr* final_evaluation, int* valid);
/* ---- The memory-consistency sum-checks the reference leaves as a stub (src/twist.rs:181-214: the closure returns zero on every branch;
 * "In a production implementation, this would involve a more complex constraint").  NOT part of the reference's proofs - the second half of the
 * non-parity mode; Twist::prove above stays byte-identical to the reference.
 * Statement: operation j < n is Read/Write { address, value } on a zero-initialised memory of memory_size = 2^k cells; every Read must return the
 * value last written to its address (0 if none).  T = 2^t = the padded number of operations.  Two sum-checks on one transcript:
 *   1. read-checking over (x, j), k + t rounds:   sum_j eq(r, j) [read_j] value_j  =  sum_{x, j} ( eq(r, j) [read_j] ra(x, j) ) * Val(x, j)
 *      the digest of (n, memory_size, addresses, values, is_write) is appended first ("memory_check_statement", two field elements), then
 *      r = challenge_field_elements("memory_check_point", t); claim appended as "memory_read_claim"; ra(x, j) = [address_j == x];
 *      Val(x, j) = content of cell x before operation j.  Ends at a point (x*, j*); the prover sends Val~(x*, j*) ("memory_val_claim").
 *   2. Val-evaluation over j', t rounds:   Val~(x*, j*)  =  sum_j' ( Inc_j' eq(x*, address_j') ) * LT~(j', j*)
 *      Inc_j' = value written minus the previous content (0 for reads); LT = [j' < j] in the natural order.
 * Both are the reference's own SumCheck::prove (src/sumcheck.rs:56-110) on those product closures.  A read that returns a wrong value makes the
 * prover fail with TSGPU_E_SUMCHECK "Round 0 consistency check failed".
 * Outputs: claims[2] = {read claim, Val~(x*, j*)}; rounds1 (k + t) x 4, final1; rounds2 t x 4, final2.
 * verify recomputes both claims' closing values from the statement (device gathers / inner products) after SumCheck::verify of each part. */
int tsgpu_twist_memory_check_prove(tsgpu_ctx* ctx, const uint64_t* addresses, const tsgpu_fr* values, const uint8_t* is_write, size_t num_operations,
                                   size_t memory_size, tsgpu_transcript* transcript, tsgpu_fr claims[2], tsgpu_fr* rounds1, tsgpu_fr* final1,
                                   tsgpu_fr* rounds2, tsgpu_fr* final2);
int tsgpu_twist_memory_check_verify(tsgpu_ctx* ctx, const uint64_t* addresses, const tsgpu_fr* values, const uint8_t* is_write, size_t num_operations,
                                    size_t memory_size, tsgpu_transcript* transcript, const tsgpu_fr claims[2], const tsgpu_fr* rounds1, size_t num_rounds1,
                                    const tsgpu_fr* final1, const tsgpu_fr* rounds2, size_t num_rounds2, const tsgpu_fr* final2, int* valid);
/* ---- write-checking, the third sum-check of Twist (with its own Val-evaluation): Inc_j = [write_j] (value_j - Val(address_j, j)) is consistent with
 * the written values and Val,   sum_j eq(r', j) Inc_j  =  sum_{x, j} ( eq(r', j) [write_j] ra(x, j) ) * ( value_j - Val(x, j) )   (k + t rounds),
 * ending in a claim Val~(x**, j**) that a second Val-evaluation sum-check (t rounds) proves.  Labels: "memory_write_statement" (digest),
 * "memory_write_point", "memory_write_claim", "memory_val_claim_2".  Runs on the caller's transcript - normally right after
 * tsgpu_twist_memory_check_prove on the same one.  claims[2] = {write claim, Val~(x**, j**)}; rounds3 (k + t) x 4, final3; rounds4 t x 4, final4.
 * A write whose recorded value differs from what the later reads return is caught by read-checking; a wrong Inc / Val pairing by this one. */
int tsgpu_twist_write_check_prove(tsgpu_ctx* ctx, const uint64_t* addresses, const tsgpu_fr* values, const uint8_t* is_write, size_t num_operations,
                                  size_t memory_size, tsgpu_transcript* transcript, tsgpu_fr claims[2], tsgpu_fr* rounds3, tsgpu_fr* final3,
                                  tsgpu_fr* rounds4, tsgpu_fr* final4);
int tsgpu_twist_write_check_verify(tsgpu_ctx* ctx, const uint64_t* addresses, const tsgpu_fr* values, const uint8_t* is_write, size_t num_operations,
                                   size_t memory_size, tsgpu_transcript* transcript, const tsgpu_fr claims[2], const tsgpu_fr* rounds3, size_t num_rounds3,
                                   const tsgpu_fr* final3, const tsgpu_fr* rounds4, size_t num_rounds4, const tsgpu_fr* final4, int* valid);
/* ---- binding the constraint sum-checks above to the KZG commitments of a Twist / Shout proof.
 * tsgpu_transcript_bind_proof: appends the two commitment hashes of `proof` under the labels Twist::prove (is_shout = 0: "address_commitment",
 * "value_commitment", src/twist.rs:157-160) or Shout::prove (1: "table_commitment", "index_commitment", src/shout.rs:129-133) use; called by prover and verifier
 * on the fresh transcript they then hand to tsgpu_*_check_prove / _verify, it makes every challenge of the sum-checks depend on the commitments.
 * tsgpu_*_commitments_match: recomputes the two commitments from the clear statement on the device and compares them with the proof's (*match = 1 / 0): what the
 * verifier runs to know that the commitments it verified openings of commit to the statement the sum-checks are about.  (Non-succinct: the verifier reads the statement.) */
int tsgpu_transcript_bind_proof(tsgpu_transcript* transcript, const tsgpu_proof* proof, int is_shout);
int tsgpu_twist_commitments_match(tsgpu_ctx* ctx, const tsgpu_params* params, const tsgpu_proof* proof, const uint64_t* addresses, const tsgpu_fr* values,
                                  size_t num_operations, int* match);
int tsgpu_shout_commitments_match(tsgpu_ctx* ctx, const tsgpu_params* params, const tsgpu_proof* proof, const tsgpu_fr* entries, size_t num_entries,
                                  const uint64_t* lookup_indices, size_t num_lookups, int* match);
/* Twist::verify / Shout::verify (src/twist.rs:255-304, src/shout.rs:225-274): transcript replay, SumCheck::verify and
 * the two KZGCommitment::verify pairing checks (src/commitments.rs:201-228) - all on the CPU, as in the reference. */
int tsgpu_twist_verify(tsgpu_ctx* ctx, const tsgpu_params* params, const tsgpu_proof* proof, int* valid);
int tsgpu_shout_verify(tsgpu_ctx* ctx, const tsgpu_params* params, const tsgpu_proof* proof, int* valid);

/* KZGCommitment::verify / batch_verify with the verification key held in the params (CPU, BN254 optimal-ate pairing) */
int tsgpu_kzg_verify(const tsgpu_params* params, const tsgpu_g1* commitment, const tsgpu_fr* point, const tsgpu_fr* value,
                     const tsgpu_g1* proof, int* valid);
/* KZGVectorCommitment::verify (src/commitments.rs:471-481) */
int tsgpu_vector_verify(const tsgpu_params* params, const tsgpu_g1* commitment, size_t index, const tsgpu_fr* value, const tsgpu_g1* proof, int* valid);
int tsgpu_kzg_batch_verify(const tsgpu_params* params, const tsgpu_g1* commitments, const tsgpu_fr* points, const tsgpu_fr* values,
                           const tsgpu_g1* proofs, size_t n, int* valid);
/* pairing self-test hooks: prod_i e(a_i G1, b_i G2) == 1 ?;  G2 generator on the twist and of order r */
int tsgpu_pairing_product_of_generators_is_one(const tsgpu_fr* a, const tsgpu_fr* b, size_t n);
int tsgpu_g2_generator_checks(void);
/* prod_i e(P_i, Q_i) == 1 for arbitrary points in canonical coordinates (4 little-endian 64-bit limbs per Fq element): g1 = n x (x, y), g2 = n x (x.c0, x.c1, y.c0, y.c1), all-zero = identity;
 * -1 when a coordinate is not reduced or a point is off its curve.  The pairing is the one behind KZGCommitment::verify (src/commitments.rs:211-226: Bn254::pairing); this hook lets the
 * published alt_bn128 pairing-check vectors (EIP-197) run against it. */
int tsgpu_pairing_check_points(const uint64_t* g1, const uint64_t* g2, size_t n);
int tsgpu_pairing_self_check(void);   /* split final exponentiation == plain power, Fq12 inverse / square / Frobenius, Jacobian G2 multiplication: 1 when all agree */

size_t tsgpu_proof_num_rounds(const tsgpu_proof* p);
size_t tsgpu_proof_num_openings(const tsgpu_proof* p);
void tsgpu_proof_commitment(const tsgpu_proof* p, int which, tsgpu_g1* out);
void tsgpu_proof_round_polynomials(const tsgpu_proof* p, tsgpu_fr* out);           /* num_rounds x 4 */
void tsgpu_proof_final_evaluation(const tsgpu_proof* p, tsgpu_fr* out);
void tsgpu_proof_opening(const tsgpu_proof* p, size_t i, tsgpu_g1* proof, tsgpu_fr* value);
void tsgpu_proof_opening_point(const tsgpu_proof* p, tsgpu_fr* out);
void tsgpu_proof_set_final_evaluation(tsgpu_proof* p, size_t i, const tsgpu_fr* v);   /* tamper helper for verify tests */
/* canonical proof bytes (the reference proofs have no serialisation; layout built from ark-serialize encodings,
 * DESIGN.md "proof bytes"); returns the length, writes when capacity suffices */
size_t tsgpu_proof_bytes(const tsgpu_proof* p, uint8_t* out, size_t capacity);
/* CPU conversions: FieldElement::from(u64) and into_bigint() */
void tsgpu_fr_from_u64(const uint64_t* in, size_t n, tsgpu_fr* out);
void tsgpu_fr_to_canonical(const tsgpu_fr* in, size_t n, tsgpu_fr* out);
void tsgpu_proof_free(tsgpu_proof* p);

/* ---- Host helpers around the path (no device work).
 * tsgpu_chacha20_u64: n outputs of ChaCha20Rng::from_seed(seed32).next_u64() (rand_chacha 0.3.1 / rand_core BlockRng word order) - the
 * generator the reference draws tau, the Fiat-Shamir challenges (src/utils.rs:81,172-192) and its benchmark traces from; exported so that
 * callers can build SURVEY 8(d)'s seeded synthetic traces (distribution B) without a second implementation.
 * tsgpu_statement_digest: the 32-byte binding digest the non-parity constraint sum-checks absorb before their first challenge
 * (host/statement_digest.hpp: two-level BLAKE2b-256 tree; domain = at most 16 bytes). */
void tsgpu_chacha20_u64(const uint8_t* seed32, size_t n, uint64_t* out);
/* from one ChaCha20Rng::from_seed(seed32): num_fr draws of Fr::rand (ark-ff 0.4.2: four next_u64, top two bits masked, rejection) and then
 * num_u64 draws of next_u64 - e.g. tau (seed [42; 32], src/utils.rs:81-84) or SURVEY 8(d)'s config-4 inputs (seed [4; 32]) */
void tsgpu_chacha20_fr_then_u64(const uint8_t* seed32, size_t num_fr, tsgpu_fr* out_fr, size_t num_u64, uint64_t* out_u64);
void tsgpu_statement_digest(const char* domain, const uint64_t* header, size_t num_header, const void* const* segments,
                            const size_t* segment_bytes, size_t num_segments, uint8_t out32[32]);

#ifdef __cplusplus
}
#endif
#endif /* TSGPU_H */

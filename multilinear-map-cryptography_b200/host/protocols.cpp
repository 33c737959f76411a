// host/protocols.cpp - host-side mirror of setup_params, Twist::prove and Shout::prove
// (src/utils.rs:79-131, src/twist.rs:107-252, src/shout.rs:97-222): same inputs, same transcript order, same
// error behaviour, every heavy step on the device through the C ABI of include/tsgpu.h.
//
// verify() replays the transcript and the sum-check exactly as the reference does and checks each KZG opening with
// the pairing equation of src/commitments.rs:201-228 (host/pairing.hpp, CPU).
#include <cstring>
#include <new>
#include <string>
#include <vector>
#include "../csrc/context.cuh"
#include "field64.hpp"
#include "pairing.hpp"
#include "sumcheck_host.hpp"
#include "transcript.hpp"

using namespace tsg;
using namespace tsg::host;

struct tsgpu_params {                 // ProverParams (+ the verifier's copy of the seed), src/utils.rs:21-50
    size_t log_size = 0;
    size_t max_operations = 0;
    tsgpu_fr tau;                     // CommitmentParams.tau (always Some)
    uint8_t fiat_shamir_seed[32];
    tsgpu_srs* srs = nullptr;         // g1_powers[0 ..= max_degree] on the device
    VerifyKey vk;                     // CommitmentVerificationKey: g1 generator, g2 generator, g2_tau (utils.rs:110-114)
};

struct tsgpu_proof {                  // TwistProof / ShoutProof (src/twist.rs:76-89, src/shout.rs:64-79)
    tsgpu_g1 commitments[2];
    std::vector<tsgpu_fr> round_polynomials;   // rounds x 4
    tsgpu_fr final_evaluation;
    std::vector<tsgpu_g1> opening_proofs;      // 0 or 2
    std::vector<tsgpu_fr> final_evaluations;   // 0 or 2
    tsgpu_fr opening_point;                    // challenges[0] (not part of the proof; kept for inspection)
};

namespace {

constexpr size_t H2D_OVERLAP_MIN_BYTES = (size_t)24 << 20;   // below this the second MSM pass costs more than the transfer it hides

size_t next_pow2(size_t n) { size_t p = 1; while (p < n) p <<= 1; return p; }   // 0usize.next_power_of_two() == 1
unsigned log2_of(size_t p) { unsigned l = 0; while (((size_t)1 << l) < p) ++l; return l; }

fr_t fr_of(const tsgpu_fr& x) { fr_t r; memcpy(r.l, x.l, 32); return r; }
tsgpu_fr abi_of(const fr_t& x) { tsgpu_fr r; memcpy(r.l, x.l, 32); return r; }

// Shared tail of Twist::prove (twist.rs:151-251) and Shout::prove (shout.rs:121-221): the two padded vectors are
// already on the device as `pa`, `pb` (values, natural order); they are interpolated in place.
int prove_two_vectors(tsgpu_ctx* ctx, const tsgpu_params* params, tsgpu_poly* pa, tsgpu_poly* pb, const char* label_a, const char* label_b,
                      unsigned rounds, tsgpu_proof** out) {
    int rc;
    // Evaluation-basis path (csrc/lagrange.cu): with [L_j(tau)]_1 available for both vector lengths the commitment is one MSM
    // over the raw values and vector_to_polynomial is never materialised; same group elements, same bytes.  Otherwise (SRS
    // uploaded without its trapdoor, or the path switched off) interpolate and commit coefficients as the reference does.
    bool eval_basis = ctx->eval_basis && tsgpu_srs_can_lagrange(params->srs) &&
                      tsgpu_srs_lagrange_prepare(ctx, params->srs, tsgpu_poly_len(pa)) == TSGPU_OK &&
                      tsgpu_srs_lagrange_prepare(ctx, params->srs, tsgpu_poly_len(pb)) == TSGPU_OK;
    if (!eval_basis) {
        if ((rc = tsgpu_poly_wait(ctx, pa)) || (rc = tsgpu_poly_wait(ctx, pb))) return rc;
        if ((rc = tsgpu_poly_interpolate_iota(ctx, pa))) return rc;            // vector_to_polynomial
        if ((rc = tsgpu_poly_interpolate_iota(ctx, pb))) return rc;
    }
    tsgpu_proof* pr = new (std::nothrow) tsgpu_proof;
    if (!pr) return fail(ctx, TSGPU_E_PROOF_GENERATION, "out of host memory");
    memset(&pr->opening_point, 0, 32);
    const tsgpu_poly* both[2] = {pa, pb};
    {
    WallTimer wt(ctx, "wall_commit");
    if (eval_basis && (tsgpu_poly_in_flight(pa) || tsgpu_poly_in_flight(pb))) {
        // one vector is still travelling on the side stream (tsgpu_twist_prove / tsgpu_shout_prove from host buffers): commit the resident one alone while it
        // does, then the other.  Two passes instead of one batched pass cost ~0.25 ms of device time at 2^20 operations and hide most of the 0.6 ms transfer.
        const int first = tsgpu_poly_in_flight(pa) ? 1 : 0, second = 1 - first;
        tsgpu_poly* late = second ? pb : pa;
        rc = tsgpu_kzg_commit_values_batch_dev(ctx, params->srs, &both[first], 1, &pr->commitments[first]);
        if (!rc) rc = tsgpu_poly_wait(ctx, late);
        if (!rc) rc = tsgpu_kzg_commit_values_batch_dev(ctx, params->srs, &both[second], 1, &pr->commitments[second]);
    } else {
        rc = tsgpu_poly_wait(ctx, pa);
        if (!rc) rc = tsgpu_poly_wait(ctx, pb);
        if (!rc) rc = eval_basis ? tsgpu_kzg_commit_values_batch_dev(ctx, params->srs, both, 2, pr->commitments)     // both commitments in one MSM pass
                                 : tsgpu_kzg_commit_batch_dev(ctx, params->srs, both, 2, pr->commitments);
    }
    }
    if (rc) { delete pr; return rc; }
    WallTimer* wtr = new WallTimer(ctx, "wall_transcript");
    Transcript tr(params->fiat_shamir_seed);
    tsgpu_fr h;
    tsgpu_g1_hash(&pr->commitments[0], &h); tr.append_field_element(label_a, fr_of(h));
    tsgpu_g1_hash(&pr->commitments[1], &h); tr.append_field_element(label_b, fr_of(h));
    // SumCheck::new(rounds, 0).prove(closure): the reference closure returns zero on every branch
    // (twist.rs:191-213, shout.rs:166-183), so every round polynomial is the zero cubic, the running sum stays
    // zero, and only the transcript advances (sumcheck.rs:86-100).  final_evaluation = closure(challenges) = 0.
    const fr_t zero4[4] = {fr_t::zero(), fr_t::zero(), fr_t::zero(), fr_t::zero()};
    for (unsigned round = 0; round < rounds; ++round) {
        for (int k = 0; k < 4; ++k) pr->round_polynomials.push_back(abi_of(zero4[k]));
        tr.append_field_elements("sumcheck_round_" + std::to_string(round), zero4, 4);
        (void)tr.challenge_field_element("sumcheck_challenge_" + std::to_string(round));
    }
    pr->final_evaluation = abi_of(fr_t::zero());
    std::vector<fr_t> ch = tr.challenge_field_elements("opening_challenges", rounds);   // twist.rs:219
    delete wtr;
    if (!ch.empty()) {                                                                  // twist.rs:226-243
        WallTimer wt(ctx, "wall_open");
        tsgpu_fr z = abi_of(ch[0]);
        pr->opening_point = z;
        if (eval_basis) {
            // the barycentric opening needs z outside the nodes 0..n-1 (a 2^-230 event): otherwise finish on coefficients
            Fr64 zc = Fr64::from_raw(z.l).from_mont();
            size_t nmax = tsgpu_poly_len(pa) > tsgpu_poly_len(pb) ? tsgpu_poly_len(pa) : tsgpu_poly_len(pb);
            if (!zc.l[1] && !zc.l[2] && !zc.l[3] && zc.l[0] < nmax) {
                if ((rc = tsgpu_poly_interpolate_iota(ctx, pa)) || (rc = tsgpu_poly_interpolate_iota(ctx, pb))) { delete pr; return rc; }
                eval_basis = false;
            }
        }
        tsgpu_fr vs[2]; tsgpu_g1 pis[2];
        rc = eval_basis ? tsgpu_kzg_open_values_batch_dev(ctx, params->srs, both, 2, &z, vs, pis)       // both openings: one MSM pass over the two quotients
                        : tsgpu_kzg_open_batch_dev(ctx, params->srs, both, 2, &z, vs, pis);
        if (rc) { delete pr; return rc; }
        for (int i = 0; i < 2; ++i) { pr->opening_proofs.push_back(pis[i]); pr->final_evaluations.push_back(vs[i]); }
    }
    *out = pr;
    return TSGPU_OK;
}

}  // namespace

extern "C" {

// ------------------------------------------------------------------------------------ setup_params
int tsgpu_setup_params(tsgpu_ctx* ctx, size_t log_size, tsgpu_params** out) {
    if (!ctx || !out) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    if (log_size > 25) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "log_size too large for the Fr two-adicity");
    tsgpu_params* p = new (std::nothrow) tsgpu_params;
    if (!p) return fail(ctx, TSGPU_E_PROOF_GENERATION, "out of host memory");
    p->log_size = log_size;
    p->max_operations = (size_t)1 << (log_size + 2);                       // utils.rs:80
    uint8_t seed[32]; memset(seed, 42, 32);
    ChaCha20Rng rng(seed);                                                 // utils.rs:81
    fr_t tau = rng.rand_field<fr_t>();                                     // utils.rs:84
    p->tau = abi_of(tau);
    const size_t max_degree = next_pow2(p->max_operations);                // utils.rs:89
    int rc = tsgpu_srs_generate(ctx, &p->tau, max_degree + 1, &p->srs);    // utils.rs:93-96
    if (rc) { delete p; return rc; }
    p->vk.g1_generator = G1J::generator();
    p->vk.g2_generator = G2A::generator();
    p->vk.g2_tau = p->vk.g2_generator.mul(Fr64::from_raw(p->tau.l));              // utils.rs:98
    rng.fill_bytes(p->fiat_shamir_seed, 32);                               // utils.rs:101-102
    rc = tsgpu_interpolate_prepare(ctx, (unsigned)(log_size + 2));
    if (!rc && ctx->eval_basis) rc = tsgpu_srs_lagrange_prepare(ctx, p->srs, p->max_operations);   // other lengths are built on first use
    if (rc) { tsgpu_srs_free(ctx, p->srs); delete p; return rc; }
    *out = p;
    return TSGPU_OK;
}
// VerifierParams only (src/utils.rs:36-50,110-128): no SRS, no GPU - what a verifier process needs
int tsgpu_setup_verifier_params(size_t log_size, tsgpu_params** out) {
    if (!out) return TSGPU_E_INVALID_PARAMETERS;
    tsgpu_params* p = new (std::nothrow) tsgpu_params;
    if (!p) return TSGPU_E_PROOF_GENERATION;
    p->log_size = log_size;
    p->max_operations = (size_t)1 << (log_size + 2);
    uint8_t seed[32]; memset(seed, 42, 32);
    ChaCha20Rng rng(seed);
    fr_t tau = rng.rand_field<fr_t>();
    p->tau = abi_of(tau);
    p->vk.g1_generator = G1J::generator();
    p->vk.g2_generator = G2A::generator();
    p->vk.g2_tau = p->vk.g2_generator.mul(Fr64::from_raw(p->tau.l));
    rng.fill_bytes(p->fiat_shamir_seed, 32);
    p->srs = nullptr;
    *out = p;
    return TSGPU_OK;
}
void tsgpu_params_free(tsgpu_ctx* ctx, tsgpu_params* p) {
    if (!p) return;
    if (p->srs) tsgpu_srs_free(ctx, p->srs);
    delete p;
}
size_t tsgpu_params_log_size(const tsgpu_params* p) { return p->log_size; }
size_t tsgpu_params_max_operations(const tsgpu_params* p) { return p->max_operations; }
void tsgpu_params_tau(const tsgpu_params* p, tsgpu_fr* out) { *out = p->tau; }
void tsgpu_params_fiat_shamir_seed(const tsgpu_params* p, uint8_t out[32]) { memcpy(out, p->fiat_shamir_seed, 32); }
const tsgpu_srs* tsgpu_params_srs(const tsgpu_params* p) { return p->srs; }

// ------------------------------------------------------------------------------------ Twist::prove
// operations[i] = Read/Write { address: addresses[i], value: values[i] } (twist.rs:16-20); is_write is the
// op_type vector the reference also builds (twist.rs:132-138) - it never influences the proof.
int tsgpu_twist_prove(tsgpu_ctx* ctx, const tsgpu_params* params, const uint64_t* addresses, const tsgpu_fr* values,
                      const uint8_t* is_write, size_t num_operations, tsgpu_proof** out) {
    (void)is_write;
    if (!ctx || !params || !out || ((!addresses || !values) && num_operations)) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    if (num_operations > params->max_operations) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "Too many operations");   // twist.rs:108-112
    const size_t padded = next_pow2(num_operations);                        // .next_power_of_two().max(1), twist.rs:141
    tsgpu_poly *pa = nullptr, *pv = nullptr;
    int rc = tsgpu_poly_from_u64(ctx, addresses, num_operations, padded, &pa);
    // large traces: the values (32 bytes each) travel on the side stream while the address vector is committed
    const bool overlap = ctx->h2d_overlap && num_operations * sizeof(tsgpu_fr) >= H2D_OVERLAP_MIN_BYTES;
    if (!rc) rc = overlap ? tsgpu_poly_upload_padded_async(ctx, values, num_operations, padded, &pv) : tsgpu_poly_upload_padded(ctx, values, num_operations, padded, &pv);
    if (!rc) rc = prove_two_vectors(ctx, params, pa, pv, "address_commitment", "value_commitment", log2_of(padded), out);
    tsgpu_poly_free(ctx, pa); tsgpu_poly_free(ctx, pv);
    return rc;
}
// same, with the two padded vectors already resident in HBM (consumed: interpolated in place)
int tsgpu_twist_prove_dev(tsgpu_ctx* ctx, const tsgpu_params* params, tsgpu_poly* padded_addresses, tsgpu_poly* padded_values, tsgpu_proof** out) {
    if (!ctx || !params || !padded_addresses || !padded_values || !out) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    size_t n = tsgpu_poly_len(padded_addresses);
    if (n != tsgpu_poly_len(padded_values) || next_pow2(n) != n) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "padded vectors must have equal power-of-two length");
    if (n > params->max_operations) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "Too many operations");
    return prove_two_vectors(ctx, params, padded_addresses, padded_values, "address_commitment", "value_commitment", log2_of(n), out);
}

}  // extern "C"

// ------------------------------------------------------------------------------------ Twist::prove / Shout::prove, one proof sharded over the ranks
// The two padded vectors (lengths ma, mb: powers of two, each >= G) are sliced by position over the G ranks of the context's
// communicator: rank r holds the entries [r m / G, (r + 1) m / G) of each (`pa`, `pb`: this rank's slices, zero padded).  Each rank
// commits its slices over ITS slice of the evaluation basis; the partial commitments (one G1 point per vector and rank) are
// all-gathered and added - the "MSM sliced by points with a final cross-GPU sum".  All ranks run the same transcript.  The opening
// value P(z) = N(z) sum_j w_j v_j / (z - j) needs one more all-gather (per rank: the product of its (z - j) and its partial sums),
// the quotient commitment a third.  Vectors of equal length (always so for Twist) share every pass and every all-gather; vectors of
// different lengths (Shout: table and lookups) are opened one after the other (two more small all-gathers).  Every rank returns the
// same proof, byte-identical to the one-GPU proof.
namespace {
int prove_two_vectors_sharded(tsgpu_ctx* ctx, const tsgpu_params* params, tsgpu_poly* pa, size_t ma, tsgpu_poly* pb, size_t mb,
                              const char* label_a, const char* label_b, unsigned rounds, tsgpu_proof** out) {
    const size_t G = (size_t)tsgpu_comm_size(ctx), rank = (size_t)tsgpu_comm_rank(ctx);
    const size_t m[2] = {ma, mb};
    const size_t first[2] = {rank * (ma / G), rank * (mb / G)};
    const tsgpu_poly* both[2] = {pa, pb};
    const bool same = ma == mb;
    tsgpu_proof* pr = new (std::nothrow) tsgpu_proof;
    if (!pr) return fail(ctx, TSGPU_E_PROOF_GENERATION, "out of host memory");
    memset(&pr->opening_point, 0, 32);
    auto sum_over_ranks = [&](const tsgpu_g1* mine, tsgpu_g1* total) -> int {      // mine[2] -> all-gather -> total[2]
        std::vector<tsgpu_g1> all(2 * G);
        int r = tsgpu_comm_allgather(ctx, mine, 2 * sizeof(tsgpu_g1), all.data());
        if (r) return r;
        for (int i = 0; i < 2; ++i) {
            G1J acc = G1J::identity();
            for (size_t g = 0; g < G; ++g) { G1J p; memcpy(&p, &all[2 * g + i], 96); acc = acc.add(p); }
            memcpy(&total[i], &acc, 96);
        }
        return TSGPU_OK;
    };
    // value_i = (prod over ranks of slice products) * (sum over ranks of partial sums), from the gathered (product, sums...) records
    auto combine = [&](const std::vector<tsgpu_fr>& all, size_t rec, size_t k, tsgpu_fr* vs) {
        Fr64 nz = Fr64::one();
        for (size_t g = 0; g < G; ++g) nz = nz * Fr64::from_raw(all[rec * g].l);
        for (size_t i = 0; i < k; ++i) {
            Fr64 s = Fr64::zero();
            for (size_t g = 0; g < G; ++g) s = s + Fr64::from_raw(all[rec * g + 1 + i].l);
            Fr64 v = nz * s;
            memcpy(vs[i].l, v.l, 32);
        }
    };
    int rc;
    tsgpu_g1 part[2];
    {
        WallTimer wt(ctx, "wall_commit");
        if (same) rc = tsgpu_kzg_commit_values_slice_batch_dev(ctx, params->srs, ma, first[0], both, 2, part);
        else {
            rc = tsgpu_kzg_commit_values_slice_batch_dev(ctx, params->srs, m[0], first[0], &both[0], 1, &part[0]);
            if (!rc) rc = tsgpu_kzg_commit_values_slice_batch_dev(ctx, params->srs, m[1], first[1], &both[1], 1, &part[1]);
        }
    }
    if (!rc) { WallTimer wt(ctx, "wall_exchange"); rc = sum_over_ranks(part, pr->commitments); }
    if (!rc) {
        WallTimer* wtr = new WallTimer(ctx, "wall_transcript");
        Transcript tr(params->fiat_shamir_seed);
        tsgpu_fr h;
        tsgpu_g1_hash(&pr->commitments[0], &h); tr.append_field_element(label_a, fr_of(h));
        tsgpu_g1_hash(&pr->commitments[1], &h); tr.append_field_element(label_b, fr_of(h));
        const fr_t zero4[4] = {fr_t::zero(), fr_t::zero(), fr_t::zero(), fr_t::zero()};
        for (unsigned round = 0; round < rounds; ++round) {
            for (int k = 0; k < 4; ++k) pr->round_polynomials.push_back(abi_of(zero4[k]));
            tr.append_field_elements("sumcheck_round_" + std::to_string(round), zero4, 4);
            (void)tr.challenge_field_element("sumcheck_challenge_" + std::to_string(round));
        }
        pr->final_evaluation = abi_of(fr_t::zero());
        std::vector<fr_t> ch = tr.challenge_field_elements("opening_challenges", rounds);
        delete wtr;
        if (!ch.empty()) {
            tsgpu_fr z = abi_of(ch[0]);
            pr->opening_point = z;
            tsgpu_fr vs[2]; tsgpu_g1 pis[2];
            if (same) {
                tsgpu_fr mine[3];                               // product of this rank's (z - j), partial sums of the two vectors
                { WallTimer wt(ctx, "wall_open_partial"); rc = tsgpu_kzg_open_values_slice_partial(ctx, ma, first[0], both, 2, &z, mine); }
                std::vector<tsgpu_fr> all(3 * G);
                if (!rc) { WallTimer wt(ctx, "wall_exchange"); rc = tsgpu_comm_allgather(ctx, mine, sizeof(mine), all.data()); }
                if (!rc) {
                    combine(all, 3, 2, vs);
                    WallTimer wt(ctx, "wall_open_finish");
                    rc = tsgpu_kzg_open_values_slice_finish(ctx, params->srs, ma, first[0], both, 2, vs, part);
                }
            } else {
                for (int i = 0; i < 2 && !rc; ++i) {            // the node inverses of a slice live in the context between the two phases
                    tsgpu_fr mine[2];
                    rc = tsgpu_kzg_open_values_slice_partial(ctx, m[i], first[i], &both[i], 1, &z, mine);
                    std::vector<tsgpu_fr> all(2 * G);
                    if (!rc) rc = tsgpu_comm_allgather(ctx, mine, sizeof(mine), all.data());
                    if (!rc) {
                        combine(all, 2, 1, &vs[i]);
                        rc = tsgpu_kzg_open_values_slice_finish(ctx, params->srs, m[i], first[i], &both[i], 1, &vs[i], &part[i]);
                    }
                }
            }
            if (!rc) { WallTimer wt(ctx, "wall_exchange"); rc = sum_over_ranks(part, pis); }
            if (!rc) for (int i = 0; i < 2; ++i) { pr->opening_proofs.push_back(pis[i]); pr->final_evaluations.push_back(vs[i]); }
        }
    }
    if (rc) { delete pr; return rc; }
    *out = pr;
    return TSGPU_OK;
}
// how many of `total` real entries fall into this rank's range [rank m / G, (rank + 1) m / G) of a vector padded to m
size_t shard_expect(size_t total, size_t m, size_t G, size_t rank) {
    const size_t count = m / G, first = rank * count;
    return total > first ? (total - first < count ? total - first : count) : 0;
}
}  // namespace

extern "C" {
// rank r passes the operations [r m / G, (r + 1) m / G) that exist, m = next_power_of_two(total_operations) (num_local of them; the rest
// of its range is the zero padding)
int tsgpu_twist_prove_sharded(tsgpu_ctx* ctx, const tsgpu_params* params, const uint64_t* addresses, const tsgpu_fr* values, size_t num_local,
                              size_t total_operations, tsgpu_proof** out) {
    if (!ctx || !params || !out || ((!addresses || !values) && num_local)) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    if (total_operations > params->max_operations) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "Too many operations");   // twist.rs:108-112
    const size_t G = (size_t)tsgpu_comm_size(ctx), rank = (size_t)tsgpu_comm_rank(ctx);
    const size_t m = next_pow2(total_operations);
    if (m < G || !tsgpu_srs_can_lagrange(params->srs)) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "sharded proving needs at least one padded operation per rank and an SRS with its trapdoor");
    if (num_local != shard_expect(total_operations, m, G, rank)) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "this rank must pass exactly the operations of its range");
    tsgpu_poly *pa = nullptr, *pv = nullptr;
    int rc = tsgpu_poly_from_u64(ctx, addresses, num_local, m / G, &pa);
    if (!rc) rc = tsgpu_poly_upload_padded(ctx, values, num_local, m / G, &pv);
    if (!rc) rc = prove_two_vectors_sharded(ctx, params, pa, m, pv, m, "address_commitment", "value_commitment", log2_of(m), out);
    tsgpu_poly_free(ctx, pa); tsgpu_poly_free(ctx, pv);
    return rc;
}
// same, with this rank's zero-padded slices (m / G entries each) already resident in HBM (not consumed)
int tsgpu_twist_prove_sharded_dev(tsgpu_ctx* ctx, const tsgpu_params* params, tsgpu_poly* local_addresses, tsgpu_poly* local_values,
                                  size_t padded_operations, tsgpu_proof** out) {
    if (!ctx || !params || !out || !local_addresses || !local_values) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    const size_t G = (size_t)tsgpu_comm_size(ctx), m = padded_operations;
    if (m > params->max_operations) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "Too many operations");
    if (next_pow2(m) != m || m < G || !tsgpu_srs_can_lagrange(params->srs)) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "sharded proving needs a power-of-two padded length, at least one padded operation per rank and an SRS with its trapdoor");
    if (tsgpu_poly_len(local_addresses) != m / G || tsgpu_poly_len(local_values) != m / G) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "this rank must pass slices of padded_operations / ranks entries");
    return prove_two_vectors_sharded(ctx, params, local_addresses, m, local_values, m, "address_commitment", "value_commitment", log2_of(m), out);
}
// Shout::prove (shout.rs:97-222) sharded the same way: rank r passes the table entries and the lookup indices that fall into its range of the
// padded table (length next_power_of_two(total_entries)) resp. of the padded lookup vector (length next_power_of_two(total_lookups)).
int tsgpu_shout_prove_sharded(tsgpu_ctx* ctx, const tsgpu_params* params, const tsgpu_fr* entries, size_t num_local_entries, size_t total_entries,
                              const uint64_t* lookup_indices, size_t num_local_lookups, size_t total_lookups, tsgpu_proof** out) {
    if (!ctx || !params || !out || (!entries && num_local_entries) || (!lookup_indices && num_local_lookups)) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    if (total_lookups > params->max_operations) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "Too many lookup operations");   // shout.rs:98-102
    const size_t G = (size_t)tsgpu_comm_size(ctx), rank = (size_t)tsgpu_comm_rank(ctx);
    const size_t mt = next_pow2(total_entries), ml = next_pow2(total_lookups);                // shout.rs:105,116
    if (mt < G || ml < G || !tsgpu_srs_can_lagrange(params->srs)) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "sharded proving needs at least one padded entry and one padded lookup per rank and an SRS with its trapdoor");
    if (mt > tsgpu_srs_len(params->srs)) return fail(ctx, TSGPU_E_COMMITMENT, "Polynomial degree exceeds setup size");               // commitments.rs:166-170
    if (num_local_entries != shard_expect(total_entries, mt, G, rank) || num_local_lookups != shard_expect(total_lookups, ml, G, rank))
        return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "this rank must pass exactly the entries and lookups of its ranges");
    tsgpu_poly *pt = nullptr, *pi = nullptr;
    int rc = tsgpu_poly_upload_padded(ctx, entries, num_local_entries, mt / G, &pt);
    if (!rc) rc = tsgpu_poly_from_u64(ctx, lookup_indices, num_local_lookups, ml / G, &pi);
    if (!rc) rc = prove_two_vectors_sharded(ctx, params, pt, mt, pi, ml, "table_commitment", "index_commitment", log2_of(ml), out);
    tsgpu_poly_free(ctx, pt); tsgpu_poly_free(ctx, pi);
    return rc;
}

// ------------------------------------------------------------------------------------ Shout::prove
int tsgpu_shout_prove(tsgpu_ctx* ctx, const tsgpu_params* params, const tsgpu_fr* entries, size_t num_entries,
                      const uint64_t* lookup_indices, size_t num_lookups, tsgpu_proof** out) {
    if (!ctx || !params || !out || (!entries && num_entries) || (!lookup_indices && num_lookups)) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    if (num_lookups > params->max_operations) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "Too many lookup operations");   // shout.rs:98-102
    const size_t table_size = next_pow2(num_entries);                       // shout.rs:105
    const size_t lookups_size = next_pow2(num_lookups);                     // shout.rs:116
    tsgpu_poly *pt = nullptr, *pi = nullptr;
    // (committing the index vector while the table travels on the side stream, as tsgpu_twist_prove does with its values, was measured on C3 - 2^20 entries, 2^22
    //  lookups - and did not pay: 19.14 against 18.95 ms; the two commit passes of different lengths batch too well)
    int rc = tsgpu_poly_from_u64(ctx, lookup_indices, num_lookups, lookups_size, &pi);
    if (!rc) rc = tsgpu_poly_upload_padded(ctx, entries, num_entries, table_size, &pt);
    if (!rc) rc = prove_two_vectors(ctx, params, pt, pi, "table_commitment", "index_commitment", log2_of(lookups_size), out);
    tsgpu_poly_free(ctx, pt); tsgpu_poly_free(ctx, pi);
    return rc;
}

// ------------------------------------------------------------------------------------ binding the constraint sum-checks to the commitments of a proof
// The non-parity sum-checks (host/read_check.cpp, host/memory_check.cpp) run on a transcript of the caller.  To tie them to the KZG commitments of the
// byte-identical Twist / Shout proof: (1) the transcript first absorbs the two commitment hashes exactly as Twist::prove / Shout::prove absorb them
// (tsgpu_transcript_bind_proof), so every challenge of the sum-checks depends on the commitments; (2) the verifier, who holds the statement in the clear,
// recomputes the two commitments from it on the device (same MSM as the prover) and compares (tsgpu_*_commitments_match): the proof's commitments then
// provably commit to THIS statement.  The verifier is not succinct - that would take a multilinear opening argument the reference does not have.
namespace {
int commitments_of_two_vectors(tsgpu_ctx* ctx, const tsgpu_params* params, tsgpu_poly* pa, tsgpu_poly* pb, tsgpu_g1 out[2]) {
    const bool eval_basis = ctx->eval_basis && tsgpu_srs_can_lagrange(params->srs) &&
                            tsgpu_srs_lagrange_prepare(ctx, params->srs, tsgpu_poly_len(pa)) == TSGPU_OK &&
                            tsgpu_srs_lagrange_prepare(ctx, params->srs, tsgpu_poly_len(pb)) == TSGPU_OK;
    int rc;
    if (!eval_basis) {
        if ((rc = tsgpu_poly_interpolate_iota(ctx, pa))) return rc;
        if ((rc = tsgpu_poly_interpolate_iota(ctx, pb))) return rc;
    }
    const tsgpu_poly* both[2] = {pa, pb};
    return eval_basis ? tsgpu_kzg_commit_values_batch_dev(ctx, params->srs, both, 2, out) : tsgpu_kzg_commit_batch_dev(ctx, params->srs, both, 2, out);
}
int match_result(const tsgpu_proof* proof, const tsgpu_g1 c[2], int* match) {
    G1J a0, a1, b0, b1;
    memcpy(&a0, &proof->commitments[0], 96); memcpy(&a1, &proof->commitments[1], 96); memcpy(&b0, &c[0], 96); memcpy(&b1, &c[1], 96);
    *match = (a0.equals(b0) && a1.equals(b1)) ? 1 : 0;
    return TSGPU_OK;
}
}  // namespace
int tsgpu_transcript_bind_proof(tsgpu_transcript* transcript, const tsgpu_proof* proof, int is_shout) {
    if (!transcript || !proof) return TSGPU_E_INVALID_PARAMETERS;
    Transcript& tr = *tsgpu_transcript_inner(transcript);
    tsgpu_fr h;
    tsgpu_g1_hash(&proof->commitments[0], &h); tr.append_field_element(is_shout ? "table_commitment" : "address_commitment", fr_of(h));   // shout.rs:129, twist.rs:157
    tsgpu_g1_hash(&proof->commitments[1], &h); tr.append_field_element(is_shout ? "index_commitment" : "value_commitment", fr_of(h));
    return TSGPU_OK;
}
int tsgpu_twist_commitments_match(tsgpu_ctx* ctx, const tsgpu_params* params, const tsgpu_proof* proof, const uint64_t* addresses, const tsgpu_fr* values,
                                  size_t num_operations, int* match) {
    if (!ctx || !params || !proof || !match || ((!addresses || !values) && num_operations)) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    if (num_operations > params->max_operations) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "Too many operations");
    const size_t padded = next_pow2(num_operations);
    tsgpu_poly *pa = nullptr, *pv = nullptr;
    tsgpu_g1 c[2];
    int rc = tsgpu_poly_from_u64(ctx, addresses, num_operations, padded, &pa);
    if (!rc) rc = tsgpu_poly_upload_padded(ctx, values, num_operations, padded, &pv);
    if (!rc) rc = commitments_of_two_vectors(ctx, params, pa, pv, c);
    tsgpu_poly_free(ctx, pa); tsgpu_poly_free(ctx, pv);
    return rc ? rc : match_result(proof, c, match);
}
int tsgpu_shout_commitments_match(tsgpu_ctx* ctx, const tsgpu_params* params, const tsgpu_proof* proof, const tsgpu_fr* entries, size_t num_entries,
                                  const uint64_t* lookup_indices, size_t num_lookups, int* match) {
    if (!ctx || !params || !proof || !match || (!entries && num_entries) || (!lookup_indices && num_lookups)) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    if (num_lookups > params->max_operations) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "Too many lookup operations");
    tsgpu_poly *pt = nullptr, *pi = nullptr;
    tsgpu_g1 c[2];
    int rc = tsgpu_poly_upload_padded(ctx, entries, num_entries, next_pow2(num_entries), &pt);
    if (!rc) rc = tsgpu_poly_from_u64(ctx, lookup_indices, num_lookups, next_pow2(num_lookups), &pi);
    if (!rc) rc = commitments_of_two_vectors(ctx, params, pt, pi, c);
    tsgpu_poly_free(ctx, pt); tsgpu_poly_free(ctx, pi);
    return rc ? rc : match_result(proof, c, match);
}

// ------------------------------------------------------------------------------------ proof accessors
size_t tsgpu_proof_num_rounds(const tsgpu_proof* p) { return p->round_polynomials.size() / 4; }
size_t tsgpu_proof_num_openings(const tsgpu_proof* p) { return p->opening_proofs.size(); }
void tsgpu_proof_commitment(const tsgpu_proof* p, int which, tsgpu_g1* out) { *out = p->commitments[which ? 1 : 0]; }
void tsgpu_proof_round_polynomials(const tsgpu_proof* p, tsgpu_fr* out) { if (!p->round_polynomials.empty()) memcpy(out, p->round_polynomials.data(), p->round_polynomials.size() * 32); }
void tsgpu_proof_final_evaluation(const tsgpu_proof* p, tsgpu_fr* out) { *out = p->final_evaluation; }
void tsgpu_proof_opening(const tsgpu_proof* p, size_t i, tsgpu_g1* proof, tsgpu_fr* value) { *proof = p->opening_proofs[i]; *value = p->final_evaluations[i]; }
void tsgpu_proof_opening_point(const tsgpu_proof* p, tsgpu_fr* out) { *out = p->opening_point; }
void tsgpu_proof_free(tsgpu_proof* p) { delete p; }

// canonical bytes: compressed(C0) | compressed(C1) | u64 rounds | per round (u64 4 | 4 x Fr) |
// Fr final_evaluation | u64 #openings | compressed... | u64 #evals | Fr...
size_t tsgpu_proof_bytes(const tsgpu_proof* p, uint8_t* out, size_t capacity) {
    std::vector<uint8_t> b;
    auto put_u64 = [&](uint64_t v) { for (int i = 0; i < 8; ++i) b.push_back((uint8_t)(v >> (8 * i))); };
    auto put_fr = [&](const tsgpu_fr& x) { uint8_t t[32]; fr_to_bytes(fr_of(x), t); b.insert(b.end(), t, t + 32); };
    auto put_g1 = [&](const tsgpu_g1& g) { uint8_t t[32]; tsgpu_g1_compress(&g, t); b.insert(b.end(), t, t + 32); };
    put_g1(p->commitments[0]); put_g1(p->commitments[1]);
    size_t rounds = p->round_polynomials.size() / 4;
    put_u64(rounds);
    for (size_t r = 0; r < rounds; ++r) { put_u64(4); for (int k = 0; k < 4; ++k) put_fr(p->round_polynomials[4 * r + k]); }
    put_fr(p->final_evaluation);
    put_u64(p->opening_proofs.size()); for (auto& g : p->opening_proofs) put_g1(g);
    put_u64(p->final_evaluations.size()); for (auto& v : p->final_evaluations) put_fr(v);
    if (out && capacity >= b.size()) memcpy(out, b.data(), b.size());
    return b.size();
}

// ------------------------------------------------------------------------------------ verify (transcript + sum-check + pairing KZG)
// Twist::verify / Shout::verify control flow (twist.rs:255-304, shout.rs:225-274); *valid = 1/0.
static int verify_common(tsgpu_ctx* ctx, const tsgpu_params* params, const tsgpu_proof* proof, const char* label_a, const char* label_b, int* valid) {
    if (!params || !proof || !valid) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    Transcript tr(params->fiat_shamir_seed);
    tsgpu_fr h;
    tsgpu_g1_hash(&proof->commitments[0], &h); tr.append_field_element(label_a, fr_of(h));
    tsgpu_g1_hash(&proof->commitments[1], &h); tr.append_field_element(label_b, fr_of(h));
    const unsigned num_vars = (unsigned)(proof->round_polynomials.size() / 4);          // twist.rs:263
    SumCheckProof sc;
    for (unsigned r = 0; r < num_vars; ++r) {
        std::vector<fr_t> c(4);
        for (int k = 0; k < 4; ++k) c[k] = fr_of(proof->round_polynomials[4 * r + k]);
        sc.round_polynomials.push_back(c);
    }
    sc.final_evaluation = fr_of(proof->final_evaluation);
    int ok = sumcheck_verify(num_vars, fr_t::zero(), sc, tr, nullptr);
    if (ok < 0) return fail(ctx, TSGPU_E_SUMCHECK, "Proof has wrong number of rounds");
    if (!ok) { *valid = 0; return TSGPU_OK; }
    std::vector<fr_t> ch = tr.challenge_field_elements("opening_challenges", num_vars);
    if (!ch.empty() && proof->opening_proofs.size() >= 2 && proof->final_evaluations.size() >= 2) {   // twist.rs:275
        Fr64 z = Fr64::from_raw(ch[0].l);
        for (int i = 0; i < 2; ++i) {                                                                 // KZGCommitment::verify x 2
            G1J C, pi; memcpy(&C, &proof->commitments[i], 96); memcpy(&pi, &proof->opening_proofs[i], 96);
            Fr64 v = Fr64::from_raw(proof->final_evaluations[i].l);
            if (!kzg_verify(params->vk, C, z, v, pi)) { *valid = 0; return TSGPU_OK; }
        }
    }
    *valid = 1;
    return TSGPU_OK;
}
int tsgpu_twist_verify(tsgpu_ctx* ctx, const tsgpu_params* params, const tsgpu_proof* proof, int* valid) {
    return verify_common(ctx, params, proof, "address_commitment", "value_commitment", valid);
}
int tsgpu_shout_verify(tsgpu_ctx* ctx, const tsgpu_params* params, const tsgpu_proof* proof, int* valid) {
    return verify_common(ctx, params, proof, "table_commitment", "index_commitment", valid);
}
// tamper helper for tests of the verify path: overwrite one final evaluation
void tsgpu_proof_set_final_evaluation(tsgpu_proof* p, size_t i, const tsgpu_fr* v) { if (i < p->final_evaluations.size()) p->final_evaluations[i] = *v; }

}  // extern "C"

// ------------------------------------------------------------------------------------ small host conversions
extern "C" {
// FieldElement::from(u64) (twist.rs:119): canonical small integer -> Montgomery limbs.  CPU.
void tsgpu_fr_from_u64(const uint64_t* in, size_t n, tsgpu_fr* out) {
    for (size_t i = 0; i < n; ++i) { Fr64 f = Fr64::from_u64(in[i]); memcpy(out[i].l, f.l, 32); }
}
// LessThanPolynomial::evaluate_at_field_elements (polynomials.rs:213-220): field_to_bits takes the low num_vars bits of into_bigint()
// (bits beyond 256 read as false), evaluate_at_bits returns on the first differing bit counted from bit 0 (polynomials.rs:222-239).  CPU.
void tsgpu_lt_evaluate_at_field_elements(unsigned num_vars, const tsgpu_fr* a, const tsgpu_fr* b, tsgpu_fr* out) {
    Fr64 ca = Fr64::from_raw(a->l).from_mont(), cb = Fr64::from_raw(b->l).from_mont();
    Fr64 res = Fr64::zero();
    for (unsigned i = 0; i < num_vars && i < 256; ++i) {
        const bool ba = (ca.l[i / 64] >> (i % 64)) & 1, bb = (cb.l[i / 64] >> (i % 64)) & 1;
        if (ba && !bb) break;
        if (!ba && bb) { res = Fr64::one(); break; }
    }
    memcpy(out->l, res.l, 32);
}
// into_bigint(): Montgomery limbs -> canonical integer limbs.  CPU.
void tsgpu_fr_to_canonical(const tsgpu_fr* in, size_t n, tsgpu_fr* out) {
    for (size_t i = 0; i < n; ++i) { Fr64 f = Fr64::from_raw(in[i].l).from_mont(); memcpy(out[i].l, f.l, 32); }
}
}

// ------------------------------------------------------------------------------------ KZG verify / pairing (CPU)
extern "C" {
// KZGCommitment::verify(vk, commitment, point, value, proof) (src/commitments.rs:201-228); vk taken from params
int tsgpu_kzg_verify(const tsgpu_params* params, const tsgpu_g1* commitment, const tsgpu_fr* point, const tsgpu_fr* value,
                     const tsgpu_g1* proof, int* valid) {
    if (!params || !commitment || !point || !value || !proof || !valid) return TSGPU_E_INVALID_PARAMETERS;
    G1J C, pi; memcpy(&C, commitment, 96); memcpy(&pi, proof, 96);
    if (!C.is_valid() || !pi.is_valid()) { *valid = 0; return TSGPU_OK; }      // not points of G1: nothing the reference's types could even hold
    *valid = kzg_verify(params->vk, C, Fr64::from_raw(point->l), Fr64::from_raw(value->l), pi) ? 1 : 0;
    return TSGPU_OK;
}
// KZGVectorCommitment::verify (src/commitments.rs:471-481): the KZG check at point Fr::from(index)
int tsgpu_vector_verify(const tsgpu_params* params, const tsgpu_g1* commitment, size_t index, const tsgpu_fr* value, const tsgpu_g1* proof, int* valid) {
    Fr64 pt = Fr64::from_u64((uint64_t)index);
    tsgpu_fr z; memcpy(z.l, pt.l, 32);
    return tsgpu_kzg_verify(params, commitment, &z, value, proof, valid);
}
// KZGCommitment::batch_verify (src/commitments.rs:230-301): random linear combination with gamma_i = Fr::rand of
// ChaCha20Rng::from_seed([42; 32]); TSGPU_E_COMMITMENT "Batch verify input lengths must match" is the caller's
// concern here (one length parameter).  Empty batch verifies.  NOTE: the reference formula applies gamma_i to the proof AND to
// the G2 side, so it rejects every non-empty batch (it is never called in the reference); mirrored as written.
int tsgpu_kzg_batch_verify(const tsgpu_params* params, const tsgpu_g1* commitments, const tsgpu_fr* points, const tsgpu_fr* values,
                           const tsgpu_g1* proofs, size_t n, int* valid) {
    if (!params || !valid || (n && (!commitments || !points || !values || !proofs))) return TSGPU_E_INVALID_PARAMETERS;
    if (n == 0) { *valid = 1; return TSGPU_OK; }
    uint8_t seed[32]; memset(seed, 42, 32);
    ChaCha20Rng rng(seed);
    G1J bc = G1J::identity(), bp = G1J::identity();
    Fr64 bv = Fr64::zero();
    G2A bg2 = G2A::infinity();
    for (size_t i = 0; i < n; ++i) {
        fr_t g32 = rng.rand_field<fr_t>();
        Fr64 gamma = Fr64::from_raw(g32.l);
        G1J C, pi; memcpy(&C, &commitments[i], 96); memcpy(&pi, &proofs[i], 96);
        if (!C.is_valid() || !pi.is_valid()) { *valid = 0; return TSGPU_OK; }
        bc = bc.add(C.mul(gamma));
        bv = bv + Fr64::from_raw(values[i].l) * gamma;
        bp = bp.add(pi.mul(gamma));
        G2A t = params->vk.g2_tau.add(params->vk.g2_generator.mul(Fr64::from_raw(points[i].l)).neg());
        bg2 = bg2.add(t.mul(gamma));
    }
    G1J left = bc.add(params->vk.g1_generator.mul(bv).neg());
    *valid = pairing_product_is_one({left, bp.neg()}, {params->vk.g2_generator, bg2}) ? 1 : 0;
    return TSGPU_OK;
}
// prod_i e(a_i * G1, b_i * G2) == 1 ?  (self-test hook for the pairing: bilinearity / non-degeneracy checks)
int tsgpu_pairing_product_of_generators_is_one(const tsgpu_fr* a, const tsgpu_fr* b, size_t n) {
    std::vector<G1J> P; std::vector<G2A> Q;
    for (size_t i = 0; i < n; ++i) {
        P.push_back(G1J::generator().mul(Fr64::from_raw(a[i].l)));
        Q.push_back(G2A::generator().mul(Fr64::from_raw(b[i].l)));
    }
    return pairing_product_is_one(P, Q) ? 1 : 0;
}
// prod_i e(P_i, Q_i) == 1 ? for ARBITRARY points given by their canonical coordinates (four little-endian 64-bit limbs each): g1 = n x (x, y), (0, 0) = the identity;
// g2 = n x (x.c0, x.c1, y.c0, y.c1) over Fq2 = Fq[u] / (u^2 + 1), all zero = the identity.  -1: a coordinate is not reduced or a point is not on its curve; else 1 / 0.
// Lets published pairing-check vectors (EIP-197 / alt_bn128 precompile tests) run against this pairing (tests/test_published_kats.py).
int tsgpu_pairing_check_points(const uint64_t* g1, const uint64_t* g2, size_t n) {
    if ((!g1 || !g2) && n) return -1;
    std::vector<G1J> P; std::vector<G2A> Q;
    for (size_t i = 0; i < n; ++i) {
        const uint64_t* a = g1 + 8 * i; const uint64_t* b = g2 + 16 * i;
        for (int k = 0; k < 2; ++k) if (Fq64::geq_mod(a + 4 * k)) return -1;
        for (int k = 0; k < 4; ++k) if (Fq64::geq_mod(b + 4 * k)) return -1;
        uint64_t any1 = 0, any2 = 0;
        for (int k = 0; k < 8; ++k) any1 |= a[k];
        for (int k = 0; k < 16; ++k) any2 |= b[k];
        G1J p = G1J::identity();
        if (any1) { p.x = Fq64::from_raw(a) * Fq64::r2(); p.y = Fq64::from_raw(a + 4) * Fq64::r2(); p.z = Fq64::one(); }
        if (!p.is_valid()) return -1;
        G2A q = G2A::infinity();
        if (any2) {
            q.x = Fq2{Fq64::from_raw(b) * Fq64::r2(), Fq64::from_raw(b + 4) * Fq64::r2()};
            q.y = Fq2{Fq64::from_raw(b + 8) * Fq64::r2(), Fq64::from_raw(b + 12) * Fq64::r2()};
            q.inf = false;
        }
        if (!q.on_curve()) return -1;
        P.push_back(p); Q.push_back(q);
    }
    return pairing_product_is_one(P, Q) ? 1 : 0;
}
// self-check of the faster pairing pieces: the split final exponentiation equals the plain power by (p^12 - 1) / r, the Fq12 inverse and
// the symmetric square agree with the product, the Jacobian G2 scalar multiplication equals repeated affine additions
int tsgpu_pairing_self_check(void) {
    Fq64 ax, ay;
    G1J P = G1J::generator().mul(Fr64::from_u64(12345));
    P.to_affine(ax, ay);
    G2A Q = G2A::generator();
    Fq12 f = miller_loop(ax, ay, false, Q);
    if (!(final_exponentiation(f) == final_exponentiation_plain(f))) return 0;
    {   // the Euclidean field inverse against the Fermat power, both fields: small, large, structured and pseudo-random elements
        Fq64 xq = Fq64::from_u64(3); Fr64 xr = Fr64::from_u64(5);
        for (int i = 0; i < 200; ++i) {
            if (!(xq.inverse() == xq.inverse_fermat()) || !(xq * xq.inverse() == Fq64::one())) return 0;
            if (!(xr.inverse() == xr.inverse_fermat()) || !(xr * xr.inverse() == Fr64::one())) return 0;
            xq = xq * xq + Fq64::from_u64(0x9e3779b97f4a7c15ull + i); xr = xr * xr + Fr64::from_u64(0xc2b2ae3d27d4eb4full + i);
        }
        const Fq64 eq[] = {Fq64::one(), Fq64::zero() - Fq64::one(), Fq64::from_u64(2), Fq64::from_raw(Fq64::one().l).dbl(), Fq64::r2(), Fq64::from_u64(1).inverse_fermat()};
        for (const Fq64& e : eq) if (!(e.inverse() == e.inverse_fermat())) return 0;
        const Fr64 er[] = {Fr64::one(), Fr64::zero() - Fr64::one(), Fr64::from_u64(2), Fr64::r2()};
        for (const Fr64& e : er) if (!(e.inverse() == e.inverse_fermat())) return 0;
        if (!Fq64::zero().inverse().is_zero() || !Fr64::zero().inverse().is_zero()) return 0;
        {   // unreduced limbs (possible only for bytes handed in over the C ABI): p itself -> 0, p + 2 -> the inverse of 2
            Fq64 m; for (int i = 0; i < 4; ++i) m.l[i] = Fq64::modl(i);
            if (!m.inverse().is_zero()) return 0;
            Fq64 m2 = m; m2.l[0] += 2;
            Fq64 two; two.l[0] = 2; two.l[1] = two.l[2] = two.l[3] = 0;
            if (!(m2.inverse() == two.inverse())) return 0;
        }
    }
    if (!(f * f.inverse()).is_one()) return 0;
    if (!(f.sqr() == f * f)) return 0;
    if (!(f.frobenius().frobenius().frobenius().frobenius().frobenius().frobenius() == f.conj6())) return 0;   // (f^p)^..6 times = f^(p^6)
    G2A acc = G2A::infinity();
    for (int i = 0; i < 11; ++i) acc = acc.add(Q);
    G2A m = Q.mul(Fr64::from_u64(11));
    if (m.inf || !(m.x == acc.x) || !(m.y == acc.y)) return 0;
    if (!Q.mul(Fr64::from_u64(1)).on_curve() || !(Q.mul(Fr64::from_u64(2)).x == Q.add(Q).x)) return 0;
    return 1;
}
int tsgpu_g2_generator_checks(void) {   // on the twist, and of order r: (r - 1) Q + Q = infinity
    G2A g = G2A::generator();
    if (!g.on_curve()) return 0;
    Fr64 m1 = Fr64::zero() - Fr64::one();
    G2A t = g.mul(m1).add(g);
    return t.inf ? 1 : 0;
}
}

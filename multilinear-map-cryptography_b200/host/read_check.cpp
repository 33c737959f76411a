// host/read_check.cpp - the lookup-correctness sum-check of Shout ("core Shout" read-checking), the constraint the reference
// leaves as a stub: its closure returns zero on every branch (src/shout.rs:157-184), so Shout::prove proves nothing about the
// lookups.  This is a separate, explicitly NON-PARITY mode; tsgpu_shout_prove stays byte-identical to the reference.
//
//   statement   lookup j reads entries[idx_j] and returns v_j                       (LookupOp { index, value }, shout.rs:17-22)
//   identity    for every r:  rv~(r) = sum_{x in {0,1}^k} ra~(x, r) * Val~(x)
//               rv~ = MLE of the returned values, Val~ = MLE of the padded table, ra(x, j) = [idx_j == x]
//   protocol    transcript <- digest(statement) (host/statement_digest.hpp: every challenge depends on the whole statement);
//               r <- transcript; claim = rv~(r); SumCheck::new(k, claim).prove(|x| ra~(x, r) * Val~(x))   (sumcheck.rs:56-110)
//   verifier    SumCheck::verify (sumcheck.rs:113-153) + final_evaluation == ra~(x*, r) * Val~(x*) from the statement
//
// All field work runs on the device through the table calls of include/tsgpu.h; the transcript and the round checks are host code.
#include <cstring>
#include <string>
#include <vector>
#include "../csrc/context.cuh"
#include "field64.hpp"
#include "statement_digest.hpp"
#include "sumcheck_host.hpp"
#include "transcript.hpp"

using namespace tsg;
using namespace tsg::host;

namespace {

size_t next_pow2(size_t n) { size_t p = 1; while (p < n) p <<= 1; return p; }
unsigned log2_of(size_t p) { unsigned l = 0; while (((size_t)1 << l) < p) ++l; return l; }
tsgpu_fr abi_of(const fr_t& x) { tsgpu_fr r; memcpy(r.l, x.l, 32); return r; }
fr_t fr_of(const tsgpu_fr& x) { fr_t r; memcpy(r.l, x.l, 32); return r; }

struct Tables {                       // frees whatever was created, on every path
    tsgpu_ctx* ctx;
    std::vector<tsgpu_table*> t;
    explicit Tables(tsgpu_ctx* c) : ctx(c) { t.reserve(16); }   // slots stay put: fewer than 16 are ever taken
    ~Tables() { for (tsgpu_table* x : t) tsgpu_table_free(ctx, x); }
    tsgpu_table** slot() { t.push_back(nullptr); return &t.back(); }
};

// the statement enters the transcript as two field elements (low / high 128 bits of its digest) before any challenge is drawn
void absorb_statement(Transcript& tr, const tsgpu_fr* entries, size_t num_entries, const uint64_t* lookup_indices, const tsgpu_fr* lookup_values,
                      size_t num_lookups) {
    const uint64_t header[2] = {(uint64_t)num_entries, (uint64_t)num_lookups};
    const StatementSegment segs[3] = {{entries, 32 * num_entries}, {lookup_indices, 8 * num_lookups}, {lookup_values, 32 * num_lookups}};
    uint8_t d[32];
    statement_digest("shout_read_check", header, 2, segs, 3, d);
    fr_t fe[2];
    for (int h = 0; h < 2; ++h) {
        Fr64 x = Fr64::zero();
        memcpy(x.l, d + 16 * h, 16);
        x = x * Fr64::r2();
        memcpy(fe[h].l, x.l, 32);
    }
    tr.append_field_elements("read_check_statement", fe, 2);
}

// the common opening of prover and verifier: the statement digest, the point r, the eq(r, .) table and the claim rv~(r)
int open_statement(tsgpu_ctx* ctx, const tsgpu_fr* entries, size_t num_entries, const uint64_t* lookup_indices, const tsgpu_fr* lookup_values,
                   size_t num_lookups, unsigned l, Transcript& tr, Tables& tabs, tsgpu_table** eq_r, tsgpu_fr* claim) {
    absorb_statement(tr, entries, num_entries, lookup_indices, lookup_values, num_lookups);
    std::vector<fr_t> r = tr.challenge_field_elements("read_check_point", l);
    std::vector<tsgpu_fr> r_abi(l ? l : 1);
    for (unsigned i = 0; i < l; ++i) r_abi[i] = abi_of(r[i]);
    tsgpu_table** e = tabs.slot();
    int rc = tsgpu_table_eq(ctx, r_abi.data(), l, e);
    if (rc) return rc;
    tsgpu_table** rv = tabs.slot();
    rc = tsgpu_table_upload(ctx, lookup_values, num_lookups, l, rv);           // from_evaluations_vec: zero padded to 2^l
    if (rc) return rc;
    rc = tsgpu_table_inner_product(ctx, *e, *rv, claim);                       // rv~(r) = sum_j eq(r, j) v_j
    if (rc) return rc;
    tr.append_field_element("read_check_claim", fr_of(*claim));
    *eq_r = *e;
    return TSGPU_OK;
}

}  // namespace

extern "C" {

int tsgpu_shout_read_check_prove(tsgpu_ctx* ctx, const tsgpu_fr* entries, size_t num_entries, const uint64_t* lookup_indices,
                                 const tsgpu_fr* lookup_values, size_t num_lookups, tsgpu_transcript* transcript,
                                 tsgpu_fr* claimed_sum, tsgpu_fr* round_polys, tsgpu_fr* final_evaluation, tsgpu_fr* challenges) {
    if (!ctx || !transcript || !claimed_sum || !final_evaluation || (!entries && num_entries) ||
        ((!lookup_indices || !lookup_values) && num_lookups)) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    const size_t K = next_pow2(num_entries), L = next_pow2(num_lookups);       // shout.rs:105,116
    const unsigned k = log2_of(K), l = log2_of(L);
    if (k && !round_polys) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    for (size_t j = 0; j < num_lookups; ++j)
        if (lookup_indices[j] >= num_entries) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "Lookup index out of bounds");      // shout.rs:44-48
    Transcript& tr = *tsgpu_transcript_inner(transcript);
    Tables tabs(ctx);
    tsgpu_table* eq_r = nullptr;
    int rc = open_statement(ctx, entries, num_entries, lookup_indices, lookup_values, num_lookups, l, tr, tabs, &eq_r, claimed_sum);
    if (rc) return rc;
    // the two sum-check tables over the table index x: ra~(x, r) and Val(x); consumed (folded in place) by the rounds
    tsgpu_table* pair[2] = {nullptr, nullptr};
    rc = tsgpu_table_scatter_add(ctx, eq_r, lookup_indices, num_lookups, k, &pair[0]);
    if (!rc) rc = tsgpu_table_upload(ctx, entries, num_entries, k, &pair[1]);
    SumCheckProof proof;
    std::vector<fr_t> ch;
    std::string err;
    if (!rc) {
        rc = sumcheck_prove_product(ctx, pair, 2, fr_of(*claimed_sum), tr, proof, &ch, nullptr, err);
        if (rc) fail(ctx, rc, err.c_str());
    }
    tsgpu_table_free(ctx, pair[0]); tsgpu_table_free(ctx, pair[1]);
    if (rc) return rc;
    for (unsigned round = 0; round < k; ++round) {
        for (int c = 0; c < 4; ++c) round_polys[4 * round + c] = abi_of(proof.round_polynomials[round][c]);
        if (challenges) challenges[round] = abi_of(ch[round]);
    }
    *final_evaluation = abi_of(proof.final_evaluation);
    return TSGPU_OK;
}

int tsgpu_shout_read_check_verify(tsgpu_ctx* ctx, const tsgpu_fr* entries, size_t num_entries, const uint64_t* lookup_indices,
                                  const tsgpu_fr* lookup_values, size_t num_lookups, tsgpu_transcript* transcript,
                                  const tsgpu_fr* round_polys, size_t num_rounds, const tsgpu_fr* final_evaluation, int* valid) {
    if (!ctx || !transcript || !final_evaluation || !valid || (!entries && num_entries) || (!round_polys && num_rounds) ||
        ((!lookup_indices || !lookup_values) && num_lookups)) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    const size_t K = next_pow2(num_entries), L = next_pow2(num_lookups);
    const unsigned k = log2_of(K), l = log2_of(L);
    for (size_t j = 0; j < num_lookups; ++j)
        if (lookup_indices[j] >= num_entries) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "Lookup index out of bounds");
    Transcript& tr = *tsgpu_transcript_inner(transcript);
    Tables tabs(ctx);
    tsgpu_table* eq_r = nullptr;
    tsgpu_fr claim;
    int rc = open_statement(ctx, entries, num_entries, lookup_indices, lookup_values, num_lookups, l, tr, tabs, &eq_r, &claim);
    if (rc) return rc;
    SumCheckProof proof;
    for (size_t round = 0; round < num_rounds; ++round) {
        std::vector<fr_t> c(4);
        for (int i = 0; i < 4; ++i) c[i] = fr_of(round_polys[4 * round + i]);
        proof.round_polynomials.push_back(c);
    }
    proof.final_evaluation = fr_of(*final_evaluation);
    std::vector<fr_t> ch;
    int ok = sumcheck_verify(k, fr_of(claim), proof, tr, &ch);
    if (ok < 0) return fail(ctx, TSGPU_E_SUMCHECK, "Proof has wrong number of rounds");                                       // sumcheck.rs:118-122
    if (!ok) { *valid = 0; return TSGPU_OK; }
    // closing check: final_evaluation == ra~(x*, r) * Val~(x*), x* = the sum-check challenges
    std::vector<tsgpu_fr> x_abi(k ? k : 1);
    for (unsigned i = 0; i < k; ++i) x_abi[i] = abi_of(ch[i]);
    tsgpu_table** eq_x = tabs.slot();
    rc = tsgpu_table_eq(ctx, x_abi.data(), k, eq_x);
    if (rc) return rc;
    tsgpu_table** g = tabs.slot();
    rc = tsgpu_table_gather(ctx, *eq_x, lookup_indices, num_lookups, l, g);    // eq(x*, idx_j)
    if (rc) return rc;
    tsgpu_fr ra, val;
    rc = tsgpu_table_inner_product(ctx, eq_r, *g, &ra);                        // ra~(x*, r) = sum_j eq(r, j) eq(x*, idx_j)
    if (rc) return rc;
    tsgpu_table** v = tabs.slot();
    rc = tsgpu_table_upload(ctx, entries, num_entries, k, v);
    if (!rc) rc = tsgpu_table_evaluate(ctx, *v, x_abi.data(), &val);           // Val~(x*)
    if (rc) return rc;
    *valid = (fr_of(ra) * fr_of(val) == proof.final_evaluation) ? 1 : 0;
    return TSGPU_OK;
}

}  // extern "C"

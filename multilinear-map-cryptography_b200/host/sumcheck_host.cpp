// host/sumcheck_host.cpp - SumCheck::prove / verify host loop (transcript on the CPU, rounds on the GPU).
#include "sumcheck_host.hpp"
#include <cstring>
#include "../csrc/context.cuh"
#include "field64.hpp"
#include "statement_digest.hpp"

namespace tsg {
namespace host {

// native 64-bit limbs for the per-round host arithmetic (fr_t on the host runs the device limb code with an emulated carry flag, ~10x slower)
static inline Fr64 w64(const fr_t& x) { return Fr64::from_raw(x.l); }
static inline fr_t n32(const Fr64& x) { fr_t r; memcpy(r.l, x.l, 32); return r; }

void interpolate4(const fr_t e32[4], fr_t c[4]) {
    // forward differences on x = 0,1,2,3, then Newton -> monomial:
    //   P = e0 + D1 x + D2 x(x-1)/2 + D3 x(x-1)(x-2)/6
    static const Fr64 inv2 = Fr64::from_u64(2).inverse(), inv3 = Fr64::from_u64(3).inverse(), inv6 = inv2 * inv3;
    const Fr64 e[4] = {w64(e32[0]), w64(e32[1]), w64(e32[2]), w64(e32[3])};
    Fr64 d1 = e[1] - e[0];
    Fr64 d2 = e[2] - e[1] - e[1] + e[0];
    Fr64 d3 = e[3] - e[2] - e[2] - e[2] + e[1] + e[1] + e[1] - e[0];
    Fr64 h2 = d2 * inv2;
    c[0] = e32[0];
    c[1] = n32(d1 - h2 + d3 * inv3);
    c[2] = n32(h2 - d3 * inv2);
    c[3] = n32(d3 * inv6);
}

static inline void to_abi(const fr_t& x, tsgpu_fr* o) { memcpy(o->l, x.l, 32); }
static inline fr_t from_abi(const tsgpu_fr& x) { fr_t r; memcpy(r.l, x.l, 32); return r; }

int sumcheck_prove_product(tsgpu_ctx* ctx, tsgpu_table* const* tables, int d, const fr_t& claimed_sum, Transcript& tr,
                           SumCheckProof& proof, std::vector<fr_t>* challenges, std::vector<fr_t>* table_finals, std::string& err) {
    tsgpu_sc* sc = nullptr;
    int rc = tsgpu_sc_begin(ctx, tables, d, &sc);
    if (rc) { err = tsgpu_last_error(ctx); return rc; }
    const unsigned num_vars = tsgpu_sc_num_vars(sc);
    tsgpu_sc_exclusive(sc, 1);             // this loop owns the context until tsgpu_sc_end: the small rounds may run in the resident tail kernel
    proof.round_polynomials.clear();
    if (challenges) challenges->clear();
    fr_t current = claimed_sum;
    tsgpu_fr ev[4];
    // DEFAULT: round 0 is summed in full and g(0) + g(1) == claimed_sum is checked before anything is appended - the reference's deterministic
    // check (sumcheck.rs:77-84): a wrong claim fails at once, with the transcript untouched (the tables are bound in place only on success paths
    // after round 0; on this error they are still the caller's unmodified tables).
    // OPT-IN (tsgpu_set_tuning("deferred_claim_check", 1)), d = 2: round 0 also runs in the claim form (g(1) = claimed_sum - g(0): a third of its products saved).  The reference's round-0 check
    // g(0) + g(1) == claimed_sum (sumcheck.rs:77-84) - the only one that can fail for honest tables - is then paid at the END: a wrong claim off by
    // delta shifts every later running sum by delta * prod_i L_1(r_i) (L_1 = the Lagrange basis polynomial of node 1), so the product of the
    // bound tables differs from the last running sum unless some challenge hits a root of L_1 (probability ~ 2 n / |Fr|, ~2^-247).  On a mismatch
    // the transcript is rolled back to its state at entry (the reference fails before touching it) and the same error is returned.
    const bool deferred = d == 2 && num_vars > 0 && ctx->deferred_claim_check;
    const size_t transcript_mark = tr.state_len();
    if (num_vars) {
        tsgpu_fr cl; to_abi(claimed_sum, &cl);
        rc = deferred ? tsgpu_sc_round_eval_claim(sc, &cl, ev) : tsgpu_sc_round_eval(sc, ev);
        if (rc) goto cuda_fail;
    }
    for (unsigned round = 0; round < num_vars; ++round) {
        fr_t e[4], coeffs[4];
        for (int i = 0; i < 4; ++i) e[i] = from_abi(ev[i]);
        interpolate4(e, coeffs);
        // g(0) + g(1) must equal the running sum (sumcheck.rs:77-84)
        fr_t g0 = horner_eval(coeffs, 4, fr_t::zero());
        fr_t g1 = horner_eval(coeffs, 4, fr_t::one());
        if (g0 + g1 != current) {
            tsgpu_sc_end(sc);
            err = "Round " + std::to_string(round) + " consistency check failed";
            return TSGPU_E_SUMCHECK;
        }
        proof.round_polynomials.emplace_back(coeffs, coeffs + 4);
        tr.append_field_elements("sumcheck_round_" + std::to_string(round), coeffs, 4);
        fr_t r = tr.challenge_field_element("sumcheck_challenge_" + std::to_string(round));
        if (challenges) challenges->push_back(r);
        current = horner_eval(coeffs, 4, r);
        tsgpu_fr rr; to_abi(r, &rr);
        tsgpu_fr cur; to_abi(current, &cur);
        if (round + 1 < num_vars) rc = tsgpu_sc_bind_eval_claim(sc, &rr, &cur, ev);   // fused bind + next round; its g(0) + g(1) is `current`
        else rc = tsgpu_sc_bind(sc, &rr);
        if (rc) goto cuda_fail;
    }
    {
        // final_evaluation = polynomial(&fixed_variables) (sumcheck.rs:104) = product of the bound tables
        std::vector<tsgpu_fr> fin((size_t)d);
        rc = tsgpu_sc_final(sc, fin.data());
        if (rc) goto cuda_fail;
        fr_t fe = fr_t::one();
        if (table_finals) table_finals->clear();
        for (int t = 0; t < d; ++t) { fr_t v = from_abi(fin[t]); fe = fe * v; if (table_finals) table_finals->push_back(v); }
        proof.final_evaluation = fe;
        if (deferred && fe != current) {       // the deferred round-0 check
            tsgpu_sc_end(sc);
            tr.truncate(transcript_mark);
            proof.round_polynomials.clear();
            if (challenges) challenges->clear();
            err = "Round 0 consistency check failed";
            return TSGPU_E_SUMCHECK;
        }
    }
    tsgpu_sc_end(sc);
    return TSGPU_OK;
cuda_fail:
    err = tsgpu_last_error(ctx);
    tsgpu_sc_end(sc);
    return rc;
}

int sumcheck_verify(unsigned num_vars, const fr_t& claimed_sum, const SumCheckProof& proof, Transcript& tr, std::vector<fr_t>* challenges) {
    if (proof.round_polynomials.size() != num_vars) return -1;
    fr_t current = claimed_sum;
    if (challenges) challenges->clear();
    for (unsigned round = 0; round < num_vars; ++round) {
        const std::vector<fr_t>& c = proof.round_polynomials[round];
        fr_t g0 = horner_eval(c.data(), c.size(), fr_t::zero());
        fr_t g1 = horner_eval(c.data(), c.size(), fr_t::one());
        if (g0 + g1 != current) return 0;
        tr.append_field_elements("sumcheck_round_" + std::to_string(round), c.data(), c.size());
        fr_t r = tr.challenge_field_element("sumcheck_challenge_" + std::to_string(round));
        if (challenges) challenges->push_back(r);
        current = horner_eval(c.data(), c.size(), r);
    }
    return current == proof.final_evaluation ? 1 : 0;
}

}  // namespace host
}  // namespace tsg

// ------------------------------------------------------------------------------------------ C ABI
using namespace tsg;
using namespace tsg::host;

struct tsgpu_transcript { Transcript tr; };
tsg::host::Transcript* tsgpu_transcript_inner(tsgpu_transcript* t) { return &t->tr; }

extern "C" {

tsgpu_transcript* tsgpu_transcript_new(const uint8_t* seed32) { return new tsgpu_transcript{Transcript(seed32)}; }
void tsgpu_transcript_free(tsgpu_transcript* t) { delete t; }
void tsgpu_transcript_append(tsgpu_transcript* t, const char* label, size_t label_len, const tsgpu_fr* elems, size_t n) {
    std::vector<fr_t> v(n);
    for (size_t i = 0; i < n; ++i) memcpy(v[i].l, elems[i].l, 32);
    t->tr.append_field_elements(std::string(label, label_len), v.data(), n);
}
void tsgpu_transcript_challenge(tsgpu_transcript* t, const char* label, size_t label_len, tsgpu_fr* out) {
    fr_t c = t->tr.challenge_field_element(std::string(label, label_len));
    memcpy(out->l, c.l, 32);
}
size_t tsgpu_transcript_state_len(const tsgpu_transcript* t) { return t->tr.state_len(); }

void tsgpu_chacha20_u64(const uint8_t* seed32, size_t n, uint64_t* out) {
    ChaCha20Rng rng(seed32);
    for (size_t i = 0; i < n; ++i) out[i] = rng.next_u64();
}
void tsgpu_chacha20_fr_then_u64(const uint8_t* seed32, size_t num_fr, tsgpu_fr* out_fr, size_t num_u64, uint64_t* out_u64) {
    ChaCha20Rng rng(seed32);
    for (size_t i = 0; i < num_fr; ++i) { fr_t f = rng.rand_field<fr_t>(); memcpy(out_fr[i].l, f.l, 32); }
    for (size_t i = 0; i < num_u64; ++i) out_u64[i] = rng.next_u64();
}
void tsgpu_statement_digest(const char* domain, const uint64_t* header, size_t num_header, const void* const* segments,
                            const size_t* segment_bytes, size_t num_segments, uint8_t out32[32]) {
    std::vector<StatementSegment> segs(num_segments);
    for (size_t i = 0; i < num_segments; ++i) segs[i] = {segments[i], segment_bytes[i]};
    statement_digest(domain, header, num_header, segs.data(), num_segments, out32);
}

int tsgpu_sumcheck_prove_product(tsgpu_ctx* ctx, tsgpu_table* const* tables, int d, const tsgpu_fr* claimed_sum,
                                 tsgpu_transcript* transcript, tsgpu_fr* round_polys, tsgpu_fr* final_evaluation,
                                 tsgpu_fr* challenges, tsgpu_fr* table_finals) {
    if (!ctx || !tables || !claimed_sum || !transcript || !final_evaluation) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    SumCheckProof proof; std::vector<fr_t> ch, fin; std::string err;
    fr_t cs; memcpy(cs.l, claimed_sum->l, 32);
    int rc = sumcheck_prove_product(ctx, tables, d, cs, transcript->tr, proof, &ch, &fin, err);
    if (rc) return fail(ctx, rc, err);
    for (size_t r = 0; r < proof.round_polynomials.size(); ++r)
        for (int k = 0; k < 4; ++k) memcpy(round_polys[4 * r + k].l, proof.round_polynomials[r][k].l, 32);
    memcpy(final_evaluation->l, proof.final_evaluation.l, 32);
    if (challenges) for (size_t i = 0; i < ch.size(); ++i) memcpy(challenges[i].l, ch[i].l, 32);
    if (table_finals) for (size_t i = 0; i < fin.size(); ++i) memcpy(table_finals[i].l, fin[i].l, 32);
    return TSGPU_OK;
}

int tsgpu_sumcheck_verify(unsigned num_vars, const tsgpu_fr* claimed_sum, const tsgpu_fr* round_polys, size_t num_rounds,
                          const tsgpu_fr* final_evaluation, tsgpu_transcript* transcript, int* valid, tsgpu_fr* challenges) {
    if (!claimed_sum || !final_evaluation || !transcript || !valid || (!round_polys && num_rounds)) return TSGPU_E_INVALID_PARAMETERS;
    SumCheckProof proof;
    for (size_t r = 0; r < num_rounds; ++r) {
        std::vector<fr_t> c(4);
        for (int k = 0; k < 4; ++k) memcpy(c[k].l, round_polys[4 * r + k].l, 32);
        proof.round_polynomials.push_back(c);
    }
    memcpy(proof.final_evaluation.l, final_evaluation->l, 32);
    fr_t cs; memcpy(cs.l, claimed_sum->l, 32);
    std::vector<fr_t> ch;
    int rc = sumcheck_verify(num_vars, cs, proof, transcript->tr, &ch);
    if (rc < 0) return TSGPU_E_SUMCHECK;   // "Proof has wrong number of rounds" (sumcheck.rs:118-122)
    *valid = rc;
    if (challenges) for (size_t i = 0; i < ch.size(); ++i) memcpy(challenges[i].l, ch[i].l, 32);
    return TSGPU_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------ helpers for sharded provers
#include "field64.hpp"
extern "C" {

// the 4 coefficients SumCheck::compute_round_polynomial returns for evaluations g(0..3) (sumcheck.rs:201-205)
void tsgpu_sumcheck_round_coeffs(const tsgpu_fr evals[4], tsgpu_fr coeffs[4]) {
    fr_t e[4], c[4];
    for (int i = 0; i < 4; ++i) memcpy(e[i].l, evals[i].l, 32);
    interpolate4(e, c);
    for (int i = 0; i < 4; ++i) memcpy(coeffs[i].l, c[i].l, 32);
}
// field_utils::horner_eval (utils.rs:217-221)
void tsgpu_horner_eval(const tsgpu_fr* coeffs, size_t n, const tsgpu_fr* x, tsgpu_fr* out) {
    std::vector<fr_t> c(n);
    for (size_t i = 0; i < n; ++i) memcpy(c[i].l, coeffs[i].l, 32);
    fr_t xx; memcpy(xx.l, x->l, 32);
    fr_t r = horner_eval(c.data(), n, xx);
    memcpy(out->l, r.l, 32);
}
void tsgpu_fr_add(const tsgpu_fr* a, const tsgpu_fr* b, tsgpu_fr* out) {
    fr_t x, y; memcpy(x.l, a->l, 32); memcpy(y.l, b->l, 32);
    fr_t r = x + y; memcpy(out->l, r.l, 32);
}
void tsgpu_fr_mul(const tsgpu_fr* a, const tsgpu_fr* b, tsgpu_fr* out) {
    fr_t x, y; memcpy(x.l, a->l, 32); memcpy(y.l, b->l, 32);
    fr_t r = x * y; memcpy(out->l, r.l, 32);
}
// Collective result -> field element: each element arrives as 8 sums of zero-extended 32-bit limbs (what an
// integer all-reduce over the ranks produces); value = sum_i limb_sum[i] * 2^(32 i) mod r.  Montgomery form is
// linear, so summing the limbs of Montgomery representations and reducing gives the Montgomery form of the sum.
void tsgpu_fr_from_limb_sums(const uint64_t* sums, size_t n, tsgpu_fr* out) {
    using tsg::host::Fr64;
    for (size_t e = 0; e < n; ++e) {
        // carry-propagate into 9 x 32-bit limbs (sums < 2^32 * ranks)
        uint64_t carry = 0; uint32_t limbs[10];
        for (int i = 0; i < 8; ++i) { uint64_t v = sums[8 * e + i] + carry; limbs[i] = (uint32_t)v; carry = v >> 32; }
        limbs[8] = (uint32_t)carry; limbs[9] = (uint32_t)(carry >> 32);
        // low 256 bits + high * 2^256: 2^256 mod r is the Montgomery one
        uint64_t lo[4];
        for (int i = 0; i < 4; ++i) lo[i] = (uint64_t)limbs[2 * i] | ((uint64_t)limbs[2 * i + 1] << 32);
        while (Fr64::geq_mod(lo)) Fr64::sub_mod(lo);
        Fr64 acc; memcpy(acc.l, lo, 32);
        uint64_t hi = (uint64_t)limbs[8] | ((uint64_t)limbs[9] << 32);
        Fr64 one = Fr64::one();   // raw limbs of 2^256 mod r
        // hi * (2^256 mod r) by double-and-add on raw residues (hi < number of ranks)
        Fr64 add = Fr64::zero();
        for (int b = 63; b >= 0; --b) { add = add + add; if ((hi >> b) & 1) add = add + one; }
        acc = acc + add;
        memcpy(out[e].l, acc.l, 32);
    }
}

}  // extern "C"

// host/sumcheck_host.cpp - SumCheck::prove / verify host loop (transcript on the CPU, rounds on the GPU).
#include "sumcheck_host.hpp"
#include <cstring>
#include "../csrc/context.cuh"

namespace tsg {
namespace host {

void interpolate4(const fr_t e[4], fr_t c[4]) {
    // forward differences on x = 0,1,2,3, then Newton -> monomial:
    //   P = e0 + D1 x + D2 x(x-1)/2 + D3 x(x-1)(x-2)/6
    const fr_t inv2 = fr_t::from_u64(2).inverse(), inv3 = fr_t::from_u64(3).inverse(), inv6 = inv2 * inv3;
    fr_t d1 = e[1] - e[0];
    fr_t d2 = e[2] - e[1] - e[1] + e[0];
    fr_t d3 = e[3] - e[2] - e[2] - e[2] + e[1] + e[1] + e[1] - e[0];
    fr_t h2 = d2 * inv2;
    c[0] = e[0];
    c[1] = d1 - h2 + d3 * inv3;
    c[2] = h2 - d3 * inv2;
    c[3] = d3 * inv6;
}

static inline void to_abi(const fr_t& x, tsgpu_fr* o) { memcpy(o->l, x.l, 32); }
static inline fr_t from_abi(const tsgpu_fr& x) { fr_t r; memcpy(r.l, x.l, 32); return r; }

int sumcheck_prove_product(tsgpu_ctx* ctx, tsgpu_table* const* tables, int d, const fr_t& claimed_sum, Transcript& tr,
                           SumCheckProof& proof, std::vector<fr_t>* challenges, std::vector<fr_t>* table_finals, std::string& err) {
    tsgpu_sc* sc = nullptr;
    int rc = tsgpu_sc_begin(ctx, tables, d, &sc);
    if (rc) { err = tsgpu_last_error(ctx); return rc; }
    const unsigned num_vars = tsgpu_sc_num_vars(sc);
    proof.round_polynomials.clear();
    if (challenges) challenges->clear();
    fr_t current = claimed_sum;
    tsgpu_fr ev[4];
    if (num_vars) { rc = tsgpu_sc_round_eval(sc, ev); if (rc) goto cuda_fail; }
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
        if (round + 1 < num_vars) rc = tsgpu_sc_bind_eval(sc, &rr, ev);   // fused bind + next round
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

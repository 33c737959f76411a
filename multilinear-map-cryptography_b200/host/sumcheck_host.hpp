// host/sumcheck_host.hpp - host-side mirror of SumCheck (src/sumcheck.rs) driving the device round kernels.
#pragma once
#include <string>
#include <vector>
#include "../../include/tsgpu.h"
#include "field64.hpp"
#include "transcript.hpp"

namespace tsg {
namespace host {

// field_utils::horner_eval (src/utils.rs:217-221)
inline fr_t horner_eval(const fr_t* coeffs, size_t n, const fr_t& x) {
    const Fr64 x64 = Fr64::from_raw(x.l);          // native 64-bit limbs (see field64.hpp)
    Fr64 acc = Fr64::zero();
    for (size_t i = n; i-- > 0;) acc = acc * x64 + Fr64::from_raw(coeffs[i].l);
    fr_t r; memcpy(r.l, acc.l, 32);
    return r;
}

// The 4 monomial coefficients of the cubic through (0,e0),(1,e1),(2,e2),(3,e3): what
// poly_utils::lagrange_interpolate returns at src/sumcheck.rs:201-205 (the interpolant is unique).
void interpolate4(const fr_t evals[4], fr_t coeffs[4]);

struct SumCheckProof {                       // src/sumcheck.rs:25-31
    std::vector<std::vector<fr_t>> round_polynomials;
    fr_t final_evaluation;
};

// SumCheck::prove (src/sumcheck.rs:56-110) for f(v) = prod_t table_t.evaluate(v), on the device.
// Returns TSGPU_OK or TSGPU_E_SUMCHECK ("Round {k} consistency check failed"); `tables` are consumed.
int sumcheck_prove_product(tsgpu_ctx* ctx, tsgpu_table* const* tables, int d, const fr_t& claimed_sum, Transcript& tr,
                           SumCheckProof& proof, std::vector<fr_t>* challenges, std::vector<fr_t>* table_finals, std::string& err);

// SumCheck::verify (src/sumcheck.rs:113-153): 1 valid, 0 invalid, -1 "Proof has wrong number of rounds"
int sumcheck_verify(unsigned num_vars, const fr_t& claimed_sum, const SumCheckProof& proof, Transcript& tr, std::vector<fr_t>* challenges);

}  // namespace host
}  // namespace tsg

// the Transcript behind an opaque tsgpu_transcript handle (for library-internal host loops)
tsg::host::Transcript* tsgpu_transcript_inner(tsgpu_transcript* t);


// host/memory_check.cpp - the memory-consistency sum-checks of Twist (read-checking + Val-evaluation), the constraint the reference
// leaves as a stub: its closure returns zero on every branch (src/twist.rs:181-214), so Twist::prove proves nothing about the trace.
// A separate, explicitly NON-PARITY mode (second half); tsgpu_twist_prove stays byte-identical to the reference.
//
//   statement   operation j < n: Read/Write { address_j, value_j } on 2^k zero-initialised cells (MemoryTrace, twist.rs:16-70);
//               every Read returns the value last written to its address
//   tables      over (cell x, cycle j), index x + 2^k j - the shape of BASELINE config 4 (2^10 x 2^16 = 2^26 entries):
//               ra(x, j) = [address_j == x],  Val(x, j) = content of x before operation j,  Inc_j = written value - previous content
//   0. transcript <- digest(statement) (host/statement_digest.hpp), so that every challenge depends on the whole trace
//   1. read-checking   sum_j eq(r, j) [read_j] value_j = sum_{x, j} ( eq(r, j) [read_j] ra(x, j) ) * Val(x, j)          k + t rounds
//   2. Val-evaluation  Val~(x*, j*) = sum_j' ( Inc_j' eq(x*, address_j') ) * LT~(j', j*)                                  t rounds
//   3. write-checking  sum_j eq(r', j) Inc_j = sum_{x, j} ( eq(r', j) [write_j] ra(x, j) ) * ( value_j - Val(x, j) )              k + t rounds
//   4. Val-evaluation of the Val~ claim that 3 ends in                                                                    t rounds
//      (tsgpu_twist_write_check_prove / _verify, at the end of this file)
//   All are SumCheck::prove (src/sumcheck.rs:56-110) on the product closures, driven by host/sumcheck_host.cpp on one transcript.
//   The verifier has the statement in the clear and recomputes the two closing values from it (device gathers and inner products).
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
    explicit Tables(tsgpu_ctx* c) : ctx(c) { t.reserve(32); }   // slots stay put: fewer than 32 are ever taken
    ~Tables() { for (tsgpu_table* x : t) tsgpu_table_free(ctx, x); }
    tsgpu_table** slot() { t.push_back(nullptr); return &t.back(); }
    void release(tsgpu_table* x) { for (auto& y : t) if (y == x) y = nullptr; }   // ownership handed to a sum-check (freed by the caller)
};

struct Statement {
    const uint64_t* addr; const tsgpu_fr* values; const uint8_t* is_write; size_t n;
    unsigned k, t;
};

int check_statement(tsgpu_ctx* ctx, const uint64_t* addresses, const tsgpu_fr* values, const uint8_t* is_write, size_t n, size_t memory_size, Statement* st) {
    if ((!addresses || !values || !is_write) && n) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    if (memory_size == 0 || (memory_size & (memory_size - 1))) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "Memory size must be power of 2");   // twist.rs:38
    st->addr = addresses; st->values = values; st->is_write = is_write; st->n = n;
    st->k = log2_of(memory_size); st->t = log2_of(next_pow2(n));
    if (st->k + st->t > 28) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "memory_size x operations exceeds 2^28 table entries");
    for (size_t j = 0; j < n; ++j) if (addresses[j] >= memory_size) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "Address out of bounds");     // twist.rs:49-53
    return TSGPU_OK;
}

// the statement enters the transcript (two field elements: low / high 128 bits of its digest) before any challenge is drawn
void absorb_statement(const Statement& st, Transcript& tr, const char* label) {
    const uint64_t header[2] = {(uint64_t)st.n, (uint64_t)1 << st.k};
    const StatementSegment segs[3] = {{st.addr, 8 * st.n}, {st.values, 32 * st.n}, {st.is_write, st.n}};
    uint8_t d[32];
    statement_digest("twist_memory_chk", header, 2, segs, 3, d);
    fr_t fe[2];
    for (int h = 0; h < 2; ++h) {
        Fr64 x = Fr64::zero();
        memcpy(x.l, d + 16 * h, 16);
        x = x * Fr64::r2();
        memcpy(fe[h].l, x.l, 32);
    }
    tr.append_field_elements(label, fe, 2);
}

// r, eq(r, .) and the read claim sum_j eq(r, j) [read_j] value_j - the opening both sides share
int open_statement(tsgpu_ctx* ctx, const Statement& st, Transcript& tr, Tables& tabs, tsgpu_table** eq_r, tsgpu_fr* claim) {
    absorb_statement(st, tr, "memory_check_statement");
    std::vector<fr_t> r = tr.challenge_field_elements("memory_check_point", st.t);
    std::vector<tsgpu_fr> r_abi(st.t ? st.t : 1);
    for (unsigned i = 0; i < st.t; ++i) r_abi[i] = abi_of(r[i]);
    tsgpu_table** e = tabs.slot();
    int rc = tsgpu_table_eq(ctx, r_abi.data(), st.t, e);
    if (rc) return rc;
    std::vector<tsgpu_fr> rd(st.n ? st.n : 1);
    for (size_t j = 0; j < st.n; ++j) { if (st.is_write[j]) memset(&rd[j], 0, 32); else rd[j] = st.values[j]; }
    tsgpu_table** rdt = tabs.slot();
    rc = tsgpu_table_upload(ctx, rd.data(), st.n, st.t, rdt);
    if (!rc) rc = tsgpu_table_inner_product(ctx, *e, *rdt, claim);
    if (rc) return rc;
    tr.append_field_element("memory_read_claim", fr_of(*claim));
    *eq_r = *e;
    return TSGPU_OK;
}

// U[j] = Inc_j * eq(x*, address_j): the first table of the Val-evaluation sum-check, from the statement
int increments_table(tsgpu_ctx* ctx, const Statement& st, const std::vector<fr_t>& x_star, Tables& tabs, tsgpu_table** out) {
    std::vector<Fr64> mem((size_t)1 << st.k, Fr64::zero());
    std::vector<tsgpu_fr> inc(st.n ? st.n : 1);
    for (size_t j = 0; j < st.n; ++j) {
        if (st.is_write[j]) {
            Fr64 v = Fr64::from_raw(st.values[j].l), d = v - mem[st.addr[j]];
            memcpy(inc[j].l, d.l, 32);
            mem[st.addr[j]] = v;
        } else memset(&inc[j], 0, 32);
    }
    tsgpu_table** inct = tabs.slot();
    int rc = tsgpu_table_upload(ctx, inc.data(), st.n, st.t, inct);
    if (rc) return rc;
    std::vector<tsgpu_fr> x_abi(st.k ? st.k : 1);
    for (unsigned i = 0; i < st.k; ++i) x_abi[i] = abi_of(x_star[i]);
    tsgpu_table** eqx = tabs.slot();
    rc = tsgpu_table_eq(ctx, x_abi.data(), st.k, eqx);
    if (rc) return rc;
    tsgpu_table** g = tabs.slot();
    rc = tsgpu_table_gather(ctx, *eqx, st.addr, st.n, st.t, g);
    if (rc) return rc;
    tsgpu_table** u = tabs.slot();
    rc = tsgpu_table_mul(ctx, *inct, *g, u);
    if (rc) return rc;
    *out = *u;
    return TSGPU_OK;
}

// LT~(a, b) for two field points: sum_i (1 - a_i) b_i prod_{l > i} (a_l b_l + (1 - a_l)(1 - b_l)), bit t - 1 most significant
Fr64 lt_eval(const std::vector<fr_t>& a, const std::vector<fr_t>& b) {
    Fr64 prefix = Fr64::one(), acc = Fr64::zero(), one = Fr64::one();
    for (size_t i = a.size(); i-- > 0;) {
        Fr64 ai = Fr64::from_raw(a[i].l), bi = Fr64::from_raw(b[i].l);
        acc = acc + prefix * (one - ai) * bi;
        prefix = prefix * (ai * bi + (one - ai) * (one - bi));
    }
    return acc;
}

void export_rounds(const SumCheckProof& p, tsgpu_fr* rounds, tsgpu_fr* fin) {
    for (size_t r = 0; r < p.round_polynomials.size(); ++r)
        for (int c = 0; c < 4; ++c) rounds[4 * r + c] = abi_of(p.round_polynomials[r][c]);
    *fin = abi_of(p.final_evaluation);
}
void import_rounds(const tsgpu_fr* rounds, size_t n, const tsgpu_fr* fin, SumCheckProof* p) {
    for (size_t r = 0; r < n; ++r) {
        std::vector<fr_t> c(4);
        for (int i = 0; i < 4; ++i) c[i] = fr_of(rounds[4 * r + i]);
        p->round_polynomials.push_back(c);
    }
    p->final_evaluation = fr_of(*fin);
}

}  // namespace

extern "C" {

int tsgpu_twist_memory_check_prove(tsgpu_ctx* ctx, const uint64_t* addresses, const tsgpu_fr* values, const uint8_t* is_write, size_t num_operations,
                                   size_t memory_size, tsgpu_transcript* transcript, tsgpu_fr claims[2], tsgpu_fr* rounds1, tsgpu_fr* final1,
                                   tsgpu_fr* rounds2, tsgpu_fr* final2) {
    if (!ctx || !transcript || !claims || !final1 || !final2) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    Statement st;
    int rc = check_statement(ctx, addresses, values, is_write, num_operations, memory_size, &st);
    if (rc) return rc;
    if ((st.k + st.t && !rounds1) || (st.t && !rounds2)) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    Transcript& tr = *tsgpu_transcript_inner(transcript);
    Tables tabs(ctx);
    tsgpu_table* eq_r = nullptr;
    if ((rc = open_statement(ctx, st, tr, tabs, &eq_r, &claims[0]))) return rc;
    std::string err;
    // ---- 1. read-checking over (x, j)
    tsgpu_table* pair1[2] = {nullptr, nullptr};
    rc = tsgpu_table_one_hot_weighted(ctx, eq_r, st.addr, st.is_write, /*flag: reads*/ 0, st.n, st.k, &pair1[0]);
    if (!rc) rc = tsgpu_table_memory_values(ctx, st.addr, st.is_write, st.values, st.n, st.k, st.t, &pair1[1]);
    SumCheckProof p1; std::vector<fr_t> ch1, fin1;
    if (!rc) {
        rc = sumcheck_prove_product(ctx, pair1, 2, fr_of(claims[0]), tr, p1, &ch1, &fin1, err);
        if (rc) fail(ctx, rc, err.c_str());
    }
    tsgpu_table_free(ctx, pair1[0]); tsgpu_table_free(ctx, pair1[1]);
    if (rc) return rc;
    export_rounds(p1, rounds1, final1);
    claims[1] = abi_of(fin1[1]);                                               // Val~(x*, j*)
    tr.append_field_element("memory_val_claim", fin1[1]);
    // ---- 2. Val-evaluation over j'
    std::vector<fr_t> x_star(ch1.begin(), ch1.begin() + st.k), j_star(ch1.begin() + st.k, ch1.end());
    tsgpu_table* u = nullptr;
    if ((rc = increments_table(ctx, st, x_star, tabs, &u))) return rc;
    std::vector<tsgpu_fr> j_abi(st.t ? st.t : 1);
    for (unsigned i = 0; i < st.t; ++i) j_abi[i] = abi_of(j_star[i]);
    tsgpu_table* pair2[2] = {u, nullptr};
    tabs.release(u);
    rc = tsgpu_table_lt_point(ctx, j_abi.data(), st.t, &pair2[1]);
    SumCheckProof p2;
    if (!rc) {
        rc = sumcheck_prove_product(ctx, pair2, 2, fin1[1], tr, p2, nullptr, nullptr, err);
        if (rc) fail(ctx, rc, err.c_str());
    }
    tsgpu_table_free(ctx, pair2[0]); tsgpu_table_free(ctx, pair2[1]);
    if (rc) return rc;
    export_rounds(p2, rounds2, final2);
    return TSGPU_OK;
}

int tsgpu_twist_memory_check_verify(tsgpu_ctx* ctx, const uint64_t* addresses, const tsgpu_fr* values, const uint8_t* is_write, size_t num_operations,
                                    size_t memory_size, tsgpu_transcript* transcript, const tsgpu_fr claims[2], const tsgpu_fr* rounds1, size_t num_rounds1,
                                    const tsgpu_fr* final1, const tsgpu_fr* rounds2, size_t num_rounds2, const tsgpu_fr* final2, int* valid) {
    if (!ctx || !transcript || !claims || !final1 || !final2 || !valid || (!rounds1 && num_rounds1) || (!rounds2 && num_rounds2))
        return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    Statement st;
    int rc = check_statement(ctx, addresses, values, is_write, num_operations, memory_size, &st);
    if (rc) return rc;
    Transcript& tr = *tsgpu_transcript_inner(transcript);
    Tables tabs(ctx);
    tsgpu_table* eq_r = nullptr;
    tsgpu_fr claim1;
    if ((rc = open_statement(ctx, st, tr, tabs, &eq_r, &claim1))) return rc;
    *valid = 0;
    if (fr_of(claim1) != fr_of(claims[0])) return TSGPU_OK;
    SumCheckProof p1, p2;
    import_rounds(rounds1, num_rounds1, final1, &p1);
    import_rounds(rounds2, num_rounds2, final2, &p2);
    std::vector<fr_t> ch1, ch2;
    int ok = sumcheck_verify(st.k + st.t, fr_of(claim1), p1, tr, &ch1);
    if (ok < 0) return fail(ctx, TSGPU_E_SUMCHECK, "Proof has wrong number of rounds");
    if (!ok) return TSGPU_OK;
    std::vector<fr_t> x_star(ch1.begin(), ch1.begin() + st.k), j_star(ch1.begin() + st.k, ch1.end());
    // closing 1: final1 == ( sum_{read j} eq(r, j) eq(j*, j) eq(x*, address_j) ) * Val~(x*, j*)
    std::vector<tsgpu_fr> x_abi(st.k ? st.k : 1), j_abi(st.t ? st.t : 1);
    for (unsigned i = 0; i < st.k; ++i) x_abi[i] = abi_of(x_star[i]);
    for (unsigned i = 0; i < st.t; ++i) j_abi[i] = abi_of(j_star[i]);
    tsgpu_table** eqj = tabs.slot();
    if ((rc = tsgpu_table_eq(ctx, j_abi.data(), st.t, eqj))) return rc;
    tsgpu_table** eqx = tabs.slot();
    if ((rc = tsgpu_table_eq(ctx, x_abi.data(), st.k, eqx))) return rc;
    tsgpu_table** g = tabs.slot();
    if ((rc = tsgpu_table_gather(ctx, *eqx, st.addr, st.n, st.t, g))) return rc;
    std::vector<tsgpu_fr> sel(st.n ? st.n : 1);
    const fr_t one = fr_t::one();
    for (size_t j = 0; j < st.n; ++j) { if (st.is_write[j]) memset(&sel[j], 0, 32); else sel[j] = abi_of(one); }
    tsgpu_table** selt = tabs.slot();
    if ((rc = tsgpu_table_upload(ctx, sel.data(), st.n, st.t, selt))) return rc;
    tsgpu_table** m1 = tabs.slot();
    if ((rc = tsgpu_table_mul(ctx, eq_r, *eqj, m1))) return rc;
    tsgpu_table** m2 = tabs.slot();
    if ((rc = tsgpu_table_mul(ctx, *g, *selt, m2))) return rc;
    tsgpu_fr ra;
    if ((rc = tsgpu_table_inner_product(ctx, *m1, *m2, &ra))) return rc;
    if (fr_of(ra) * fr_of(claims[1]) != p1.final_evaluation) return TSGPU_OK;
    tr.append_field_element("memory_val_claim", fr_of(claims[1]));
    ok = sumcheck_verify(st.t, fr_of(claims[1]), p2, tr, &ch2);
    if (ok < 0) return fail(ctx, TSGPU_E_SUMCHECK, "Proof has wrong number of rounds");
    if (!ok) return TSGPU_OK;
    // closing 2: final2 == U~(j**) * LT~(j**, j*)
    tsgpu_table* u = nullptr;
    if ((rc = increments_table(ctx, st, x_star, tabs, &u))) return rc;
    std::vector<tsgpu_fr> jj_abi(st.t ? st.t : 1);
    for (unsigned i = 0; i < st.t; ++i) jj_abi[i] = abi_of(ch2[i]);
    tsgpu_fr u_val;
    if ((rc = tsgpu_table_evaluate(ctx, u, jj_abi.data(), &u_val))) return rc;
    Fr64 lt = lt_eval(ch2, j_star);
    fr_t lt32; memcpy(lt32.l, lt.l, 32);
    *valid = (fr_of(u_val) * lt32 == p2.final_evaluation) ? 1 : 0;
    return TSGPU_OK;
}

// ---- 3. write-checking + 4. its Val-evaluation: the third sum-check of Twist.  Inc_j = [write_j] (value_j - Val(address_j, j)) is what a write adds to
// its cell; the prover of the paper commits to Inc and proves it consistent with the written values and Val:
//     sum_j eq(r', j) Inc_j  =  sum_{x, j} ( eq(r', j) [write_j] wa(x, j) ) * ( wv(j) - Val(x, j) )                                   k + t rounds
// with wa(x, j) = [address_j == x] and wv(j) = value_j broadcast over the cells.  It ends at a new point (x**, j**): the second factor there is
// value~(j**) - Val~(x**, j**), so the prover sends Val~(x**, j**) and proves it by a second Val-evaluation sum-check (t rounds), exactly as part 2.
// Runs on the caller's transcript (after tsgpu_twist_memory_check_prove on the same transcript, or alone: the statement digest is absorbed here too).
// claims[2] = {sum_j eq(r', j) Inc_j, Val~(x**, j**)}; rounds3 (k + t) x 4, final3; rounds4 t x 4, final4.
namespace {
// Inc as a table over the cycles (zero for reads and for the padding)
int inc_table(tsgpu_ctx* ctx, const Statement& st, Tables& tabs, tsgpu_table** out) {
    std::vector<Fr64> mem((size_t)1 << st.k, Fr64::zero());
    std::vector<tsgpu_fr> inc(st.n ? st.n : 1);
    for (size_t j = 0; j < st.n; ++j) {
        if (st.is_write[j]) {
            Fr64 v = Fr64::from_raw(st.values[j].l), d = v - mem[st.addr[j]];
            memcpy(inc[j].l, d.l, 32);
            mem[st.addr[j]] = v;
        } else memset(&inc[j], 0, 32);
    }
    tsgpu_table** t = tabs.slot();
    int rc = tsgpu_table_upload(ctx, inc.data(), st.n, st.t, t);
    if (!rc) *out = *t;
    return rc;
}
int open_write_statement(tsgpu_ctx* ctx, const Statement& st, Transcript& tr, Tables& tabs, tsgpu_table** eq_r, tsgpu_fr* claim) {
    absorb_statement(st, tr, "memory_write_statement");
    std::vector<fr_t> r = tr.challenge_field_elements("memory_write_point", st.t);
    std::vector<tsgpu_fr> r_abi(st.t ? st.t : 1);
    for (unsigned i = 0; i < st.t; ++i) r_abi[i] = abi_of(r[i]);
    tsgpu_table** e = tabs.slot();
    int rc = tsgpu_table_eq(ctx, r_abi.data(), st.t, e);
    if (rc) return rc;
    tsgpu_table* inc = nullptr;
    if ((rc = inc_table(ctx, st, tabs, &inc))) return rc;
    if ((rc = tsgpu_table_inner_product(ctx, *e, inc, claim))) return rc;
    tr.append_field_element("memory_write_claim", fr_of(*claim));
    *eq_r = *e;
    return TSGPU_OK;
}
}  // namespace

int tsgpu_twist_write_check_prove(tsgpu_ctx* ctx, const uint64_t* addresses, const tsgpu_fr* values, const uint8_t* is_write, size_t num_operations,
                                  size_t memory_size, tsgpu_transcript* transcript, tsgpu_fr claims[2], tsgpu_fr* rounds3, tsgpu_fr* final3,
                                  tsgpu_fr* rounds4, tsgpu_fr* final4) {
    if (!ctx || !transcript || !claims || !final3 || !final4) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    Statement st;
    int rc = check_statement(ctx, addresses, values, is_write, num_operations, memory_size, &st);
    if (rc) return rc;
    if ((st.k + st.t && !rounds3) || (st.t && !rounds4)) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    Transcript& tr = *tsgpu_transcript_inner(transcript);
    Tables tabs(ctx);
    tsgpu_table* eq_r = nullptr;
    if ((rc = open_write_statement(ctx, st, tr, tabs, &eq_r, &claims[0]))) return rc;
    std::string err;
    // ---- 3. write-checking over (x, j): WA = eq(r', j) [write_j] wa(x, j);  D = wv(j) - Val(x, j)
    tsgpu_table* pair3[2] = {nullptr, nullptr};
    tsgpu_table** wv = tabs.slot();
    rc = tsgpu_table_one_hot_weighted(ctx, eq_r, st.addr, st.is_write, /*flag: writes*/ 1, st.n, st.k, &pair3[0]);
    if (!rc) rc = tsgpu_table_memory_values(ctx, st.addr, st.is_write, st.values, st.n, st.k, st.t, &pair3[1]);
    if (!rc) rc = tsgpu_table_upload(ctx, st.values, st.n, st.t, wv);
    if (!rc) rc = tsgpu_table_broadcast_rows_minus(ctx, *wv, pair3[1]);
    SumCheckProof p3; std::vector<fr_t> ch3, fin3;
    if (!rc) {
        rc = sumcheck_prove_product(ctx, pair3, 2, fr_of(claims[0]), tr, p3, &ch3, &fin3, err);
        if (rc) fail(ctx, rc, err.c_str());
    }
    tsgpu_table_free(ctx, pair3[0]); tsgpu_table_free(ctx, pair3[1]);
    if (rc) return rc;
    export_rounds(p3, rounds3, final3);
    // Val~(x**, j**) = value~(j**) - D~(x**, j**)
    std::vector<fr_t> x_star(ch3.begin(), ch3.begin() + st.k), j_star(ch3.begin() + st.k, ch3.end());
    std::vector<tsgpu_fr> j_abi(st.t ? st.t : 1);
    for (unsigned i = 0; i < st.t; ++i) j_abi[i] = abi_of(j_star[i]);
    tsgpu_fr wv_at;
    if ((rc = tsgpu_table_evaluate(ctx, *wv, j_abi.data(), &wv_at))) return rc;
    const fr_t val_claim = fr_of(wv_at) - fin3[1];
    claims[1] = abi_of(val_claim);
    tr.append_field_element("memory_val_claim_2", val_claim);
    // ---- 4. Val-evaluation at (x**, j**)
    tsgpu_table* u = nullptr;
    if ((rc = increments_table(ctx, st, x_star, tabs, &u))) return rc;
    tsgpu_table* pair4[2] = {u, nullptr};
    tabs.release(u);
    rc = tsgpu_table_lt_point(ctx, j_abi.data(), st.t, &pair4[1]);
    SumCheckProof p4;
    if (!rc) {
        rc = sumcheck_prove_product(ctx, pair4, 2, val_claim, tr, p4, nullptr, nullptr, err);
        if (rc) fail(ctx, rc, err.c_str());
    }
    tsgpu_table_free(ctx, pair4[0]); tsgpu_table_free(ctx, pair4[1]);
    if (rc) return rc;
    export_rounds(p4, rounds4, final4);
    return TSGPU_OK;
}

int tsgpu_twist_write_check_verify(tsgpu_ctx* ctx, const uint64_t* addresses, const tsgpu_fr* values, const uint8_t* is_write, size_t num_operations,
                                   size_t memory_size, tsgpu_transcript* transcript, const tsgpu_fr claims[2], const tsgpu_fr* rounds3, size_t num_rounds3,
                                   const tsgpu_fr* final3, const tsgpu_fr* rounds4, size_t num_rounds4, const tsgpu_fr* final4, int* valid) {
    if (!ctx || !transcript || !claims || !final3 || !final4 || !valid || (!rounds3 && num_rounds3) || (!rounds4 && num_rounds4))
        return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    Statement st;
    int rc = check_statement(ctx, addresses, values, is_write, num_operations, memory_size, &st);
    if (rc) return rc;
    Transcript& tr = *tsgpu_transcript_inner(transcript);
    Tables tabs(ctx);
    tsgpu_table* eq_r = nullptr;
    tsgpu_fr claim3;
    if ((rc = open_write_statement(ctx, st, tr, tabs, &eq_r, &claim3))) return rc;
    *valid = 0;
    if (fr_of(claim3) != fr_of(claims[0])) return TSGPU_OK;
    SumCheckProof p3, p4;
    import_rounds(rounds3, num_rounds3, final3, &p3);
    import_rounds(rounds4, num_rounds4, final4, &p4);
    std::vector<fr_t> ch3, ch4;
    int ok = sumcheck_verify(st.k + st.t, fr_of(claim3), p3, tr, &ch3);
    if (ok < 0) return fail(ctx, TSGPU_E_SUMCHECK, "Proof has wrong number of rounds");
    if (!ok) return TSGPU_OK;
    std::vector<fr_t> x_star(ch3.begin(), ch3.begin() + st.k), j_star(ch3.begin() + st.k, ch3.end());
    // closing 3: final3 == ( sum_{write j} eq(r', j) eq(j**, j) eq(x**, address_j) ) * ( value~(j**) - Val~(x**, j**) )
    std::vector<tsgpu_fr> x_abi(st.k ? st.k : 1), j_abi(st.t ? st.t : 1);
    for (unsigned i = 0; i < st.k; ++i) x_abi[i] = abi_of(x_star[i]);
    for (unsigned i = 0; i < st.t; ++i) j_abi[i] = abi_of(j_star[i]);
    tsgpu_table** eqj = tabs.slot();
    if ((rc = tsgpu_table_eq(ctx, j_abi.data(), st.t, eqj))) return rc;
    tsgpu_table** eqx = tabs.slot();
    if ((rc = tsgpu_table_eq(ctx, x_abi.data(), st.k, eqx))) return rc;
    tsgpu_table** g = tabs.slot();
    if ((rc = tsgpu_table_gather(ctx, *eqx, st.addr, st.n, st.t, g))) return rc;
    std::vector<tsgpu_fr> sel(st.n ? st.n : 1);
    const fr_t one = fr_t::one();
    for (size_t j = 0; j < st.n; ++j) { if (st.is_write[j]) sel[j] = abi_of(one); else memset(&sel[j], 0, 32); }
    tsgpu_table** selt = tabs.slot();
    if ((rc = tsgpu_table_upload(ctx, sel.data(), st.n, st.t, selt))) return rc;
    tsgpu_table** m1 = tabs.slot();
    if ((rc = tsgpu_table_mul(ctx, eq_r, *eqj, m1))) return rc;
    tsgpu_table** m2 = tabs.slot();
    if ((rc = tsgpu_table_mul(ctx, *g, *selt, m2))) return rc;
    tsgpu_fr wa;
    if ((rc = tsgpu_table_inner_product(ctx, *m1, *m2, &wa))) return rc;
    tsgpu_table** wv = tabs.slot();
    if ((rc = tsgpu_table_upload(ctx, st.values, st.n, st.t, wv))) return rc;
    tsgpu_fr wv_at;
    if ((rc = tsgpu_table_evaluate(ctx, *wv, j_abi.data(), &wv_at))) return rc;
    if (fr_of(wa) * (fr_of(wv_at) - fr_of(claims[1])) != p3.final_evaluation) return TSGPU_OK;
    tr.append_field_element("memory_val_claim_2", fr_of(claims[1]));
    ok = sumcheck_verify(st.t, fr_of(claims[1]), p4, tr, &ch4);
    if (ok < 0) return fail(ctx, TSGPU_E_SUMCHECK, "Proof has wrong number of rounds");
    if (!ok) return TSGPU_OK;
    // closing 4: final4 == U~(j***) * LT~(j***, j**)
    tsgpu_table* u = nullptr;
    if ((rc = increments_table(ctx, st, x_star, tabs, &u))) return rc;
    std::vector<tsgpu_fr> jj_abi(st.t ? st.t : 1);
    for (unsigned i = 0; i < st.t; ++i) jj_abi[i] = abi_of(ch4[i]);
    tsgpu_fr u_val;
    if ((rc = tsgpu_table_evaluate(ctx, u, jj_abi.data(), &u_val))) return rc;
    Fr64 lt = lt_eval(ch4, j_star);
    fr_t lt32; memcpy(lt32.l, lt.l, 32);
    *valid = (fr_of(u_val) * lt32 == p4.final_evaluation) ? 1 : 0;
    return TSGPU_OK;
}

}  // extern "C"

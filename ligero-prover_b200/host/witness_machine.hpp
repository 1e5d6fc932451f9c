// The witness machine behind the .wat / .wasm front end (wat_emitter.hpp): what the reference spreads over witness_manager
// (pools, constraints, release into rows: include/zkp/backend/witness_manager.hpp), ligetron_backend (expression evaluation, bit
// (de)composition, the bitwise / comparison / division gadgets: include/zkp/backend/core.hpp, ligetron_backend.hpp) and the C++
// object lifetimes that decide WHEN a witness is released.  Test infrastructure pins it to runs of the reference itself
// (tests/test_refctx_cpu.py); see the parity statement in wat_emitter.hpp.
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <utility>
#include <vector>

#include "../csrc/host_fr.h"
#include "fiat_shamir.hpp"
#include "row_packer.hpp"

namespace ligero::cuda::host {

using lgr::host::Fr;

// ---- the witness machine ------------------------------------------------------------------------------------
struct wat_stats {
    uint64_t private_consts = 0, asserts = 0, arithmetic_ops = 0;
    uint64_t linear_witnesses = 0, quadratic_slots = 0, linear_constraints = 0;   // witnesses committed to linear rows; slots; draws from the linear stream
    uint64_t violated_constraints = 0;           // > 0: the program's assertions do not hold (the proof will not validate)
};

// What the reference spreads over witness_manager (pools, constraints, release into rows), ligetron_backend (expression
// evaluation, bit (de)composition, the bitwise / comparison / division gadgets) and the C++ object lifetimes that decide WHEN
// a witness is released (shared_ptr<lazy_witness> with a committing deleter, core.hpp:60-100,277-300).  Here: witnesses are
// indices into one table with a reference count; `wref` is the counted handle (the last one to die releases the witness
// into the packer), `bitvec` a vector of handles that dies most-significant-bit first and -- like the reference's
// decomposed_bits, which has a destructor and hence no move constructor -- can only be copied, `expr` a small run-time tree
// in place of the reference's expression templates.  The gadgets below are written so that handles are created, copied and
// dropped where the reference's are; the C++ rules (temporaries die at the end of the full expression, locals in reverse
// order of declaration) then produce the reference's release order without it being spelled out.
class witness_machine {
public:
    using wid = uint32_t;
    static constexpr wid none = 0xFFFFFFFFu;

    class wref {
    public:
        wref() = default;
        wref(witness_machine *m, wid w) : m_(m), w_(w) { m_->w_[w_].refs++; }
        wref(const wref &o) : m_(o.m_), w_(o.w_) { if (m_) m_->w_[w_].refs++; }
        wref(wref &&o) noexcept : m_(o.m_), w_(o.w_) { o.m_ = nullptr; }
        wref &operator=(const wref &o) { wref t(o); swap(t); return *this; }           // the old witness goes before the
        wref &operator=(wref &&o) noexcept { wref t(std::move(o)); swap(t); return *this; }   // assignment returns (shared_ptr)
        ~wref() { reset(); }
        void reset() {
            if (!m_) return;
            witness_machine *m = m_;
            m_ = nullptr;
            if (--m->w_[w_].refs == 0) m->release(w_);
        }
        explicit operator bool() const { return m_ != nullptr; }
        wid id() const { return w_; }
        const Fr &val() const { return m_->w_[w_].val; }

    private:
        void swap(wref &o) { std::swap(m_, o.m_); std::swap(w_, o.w_); }
        witness_machine *m_ = nullptr;
        wid w_ = 0;
    };

    // decomposed_bits (core.hpp:93-150): least significant bit first
    class bitvec {
    public:
        bitvec() = default;
        bitvec(const bitvec &) = default;                    // copy only: "moving" a bitvec shares its witnesses with the source
        bitvec &operator=(const bitvec &) = default;
        ~bitvec() { while (!b_.empty()) b_.pop_back(); }
        size_t size() const { return b_.size(); }
        wref &operator[](size_t i) { return b_[i]; }
        const wref &operator[](size_t i) const { return b_[i]; }
        void push_back(wref w) { b_.push_back(std::move(w)); }
        void push_lsb(wref w, size_t n) { b_.insert(b_.begin(), n, w); }
        void push_msb(wref w, size_t n) { b_.insert(b_.end(), n, w); }
        void drop_lsb(size_t n) {
            for (size_t i = n; i-- > 0;) b_[i].reset();
            b_.erase(b_.begin(), b_.begin() + (ptrdiff_t)n);
        }
        void drop_msb(size_t n) { for (size_t i = 0; i < n; i++) b_.pop_back(); }
        void clear() { while (!b_.empty()) b_.pop_back(); }

    private:
        std::vector<wref> b_;
    };

    // an expression over witnesses and constants (core.hpp:152-270); children die first operand first, as a std::tuple does
    struct expr {
        enum kind_t : uint8_t { WIT, CONST, ADD, SUB, MUL, NOT, AND };
        kind_t kind;
        wref w;
        Fr k{};
        std::unique_ptr<expr> a, b;
        expr(const wref &x) : kind(WIT), w(x) {}             // NOLINT: implicit on purpose (x & ~y reads like the reference)
        expr(wref &&x) : kind(WIT), w(std::move(x)) {}       // NOLINT
        explicit expr(const Fr &c) : kind(CONST), k(c) {}
        expr(kind_t kd, expr &&x) : kind(kd), a(new expr(std::move(x))) {}
        expr(kind_t kd, expr &&x, expr &&y) : kind(kd), a(new expr(std::move(x))), b(new expr(std::move(y))) {}
        expr(expr &&) = default;
        expr &operator=(expr &&) = delete;
        ~expr() { w.reset(); a.reset(); b.reset(); }
    };
    static expr K(uint64_t v) { return expr(lgr::host::from_u64(v)); }
    static expr K(const Fr &v) { return expr(v); }

    // rows go to `pk` as witnesses are released; with a stage-1 seed the linear-test coefficients are drawn as the reference draws them
    witness_machine(row_packer &pk, const uint8_t *stage1_seed) : pk_(pk), seeded_(stage1_seed != nullptr) {
        static const uint8_t any_iv[16] = {0};
        if (seeded_) rng_.init(stage1_seed, any_iv);
    }
    witness_machine(const witness_machine &) = delete;
    witness_machine &operator=(const witness_machine &) = delete;

    // ---- witness_manager --------------------------------------------------------------------------------------
    wid acquire_raw(const Fr &v) {
        wid w;
        if (!free_w_.empty()) { w = free_w_.back(); free_w_.pop_back(); w_[w] = wit{v, zero(), none, 0, 0}; }
        else { w_.push_back(wit{v, zero(), none, 0, 0}); w = (wid)(w_.size() - 1); }
        return w;
    }
    wref acquire(const Fr &v) { return wref(this, acquire_raw(v)); }
    const Fr &value(wid w) const { return w_[w].val; }

    Fr draw() {                                               // generate_linear_random (witness_manager.hpp:344-348)
        draws_++;
        if (!seeded_) return zero();
        uint32_t limbs[8];
        rng_.next(limbs);
        return lgr::host::from_u32(limbs);
    }
    // (without a seed every rho is zero: the coefficient rows and const_sum stay zero and nothing needs adding)
    void coef_add(wid w, const Fr &r) { if (seeded_) w_[w].coef = lgr::host::add(w_[w].coef, r); }
    void coef_sub(wid w, const Fr &r) { if (seeded_) w_[w].coef = sub(w_[w].coef, r); }
    void const_add(const Fr &r) { if (seeded_) const_sum_ = lgr::host::add(const_sum_, r); }
    void const_sub(const Fr &r) { if (seeded_) const_sum_ = sub(const_sum_, r); }

    // constrain_equal (witness_manager.hpp:421-429): one draw, +r on a, -r on b
    void constrain_equal(wid a, wid b) {
        if (!(w_[a].val == w_[b].val)) violated_++;
        const Fr r = draw();
        coef_add(a, r);
        coef_sub(b, r);
    }
    // constrain_constant (:399-419): one draw, +r on the witness, -v*r on the constant
    void constrain_constant(wid w, const Fr &v) {
        if (!(w_[w].val == v)) violated_++;
        const Fr r = draw();
        coef_add(w, r);
        const_sub(fmul(v, r));
    }
    wid clone_raw(wid w) { const wid c = acquire_raw(w_[w].val); constrain_equal(w, c); return c; }       // clone_witness (:393-397)

    // constrain_quadratic (:474-492): slot positions (a, b, c); a witness that already sits in a slot is replaced by a clone
    void constrain_quadratic(wid c, wid a, wid b) {
        if (!(fmul(w_[a].val, w_[b].val) == w_[c].val)) violated_++;
        uint32_t s;
        if (!free_s_.empty()) { s = free_s_.back(); free_s_.pop_back(); slots_[s] = slot{}; }
        else { slots_.push_back(slot{}); s = (uint32_t)slots_.size() - 1; }
        nslots_++;
        const wid arr[3] = {a, b, c};
        for (int i = 0; i < 3; i++) {
            wid w = arr[i];
            const bool taken = w_[w].slot != none;
            if (taken) w = clone_raw(arr[i]);
            w_[w].slot = s; w_[w].pos = i;
            slots_[s].w[i] = w;
            if (taken) release(w);
        }
    }
    // constrain_bit (:431-441)
    void constrain_bit(wid b) {
        const wid b1 = clone_raw(b), b2 = clone_raw(b);
        constrain_quadratic(b, b1, b2);
        release(b1);
        release(b2);
    }

    // commit_release_witness (:117-186)
    void release(wid w) {
        // (an Fr is four little-endian 64-bit limbs: the same 32 bytes as the eight 32-bit limbs of a row element)
        const auto limbs = [](const Fr &x) { return reinterpret_cast<const uint32_t *>(x.v); };
        if (w_[w].slot == none) {
            pk_.push_linear(limbs(w_[w].val), limbs(w_[w].coef));
            nlinear_++;
            free_w_.push_back(w);
            return;
        }
        const uint32_t si = w_[w].slot;
        slot &s = slots_[si];
        s.ready[w_[w].pos] = true;
        if (!(s.ready[0] && s.ready[1] && s.ready[2])) return;
        const wit &x = w_[s.w[0]], &y = w_[s.w[1]], &z = w_[s.w[2]];
        pk_.push_quadratic(limbs(x.val), limbs(y.val), limbs(z.val), limbs(x.coef), limbs(y.coef), limbs(z.coef));
        for (int j = 0; j < 3; j++) free_w_.push_back(s.w[j]);
        free_s_.push_back(si);
    }

    // ---- ligetron_backend ---------------------------------------------------------------------------------------
    // eval(expr) -> a witness holding its value (core.hpp:303-311 and the eval_impl overloads :319-690)
    wref eval(const expr &e) {
        switch (e.kind) {
        case expr::WIT: return e.w;
        case expr::CONST: {                                   // zkexpr<constant>::eval (:179-184)
            const wid w = acquire_raw(e.k);
            constrain_constant(w, e.k);
            return wref(this, w);
        }
        case expr::MUL:
            if (e.b->kind == expr::CONST) break;
            [[fallthrough]];
        case expr::AND: {                                     // :537-550, :637-652: operands materialised, slot (x, y, z)
            wref x = eval(*e.a);
            wref y = eval(*e.b);
            const Fr zv = e.kind == expr::AND ? lgr::host::from_u64(x.val().v[0] & y.val().v[0]) : fmul(x.val(), y.val());
            const wid z = acquire_raw(zv);
            constrain_quadratic(z, x.id(), y.id());
            return wref(this, z);
        }
        default: break;
        }
        // linear forms: a fresh witness takes -r, the leaves take +-r (scaled), constants go to const_sum
        const wid w = acquire_raw(zero());
        const Fr r = draw();
        coef_sub(w, r);
        const Fr out = spread(e, r);
        w_[w].val = out;
        return wref(this, w);
    }
    // eval(expr, result, rand): value of the expression, `r` handed down to its leaves
    Fr spread(const expr &e, const Fr &r) {
        using lgr::host::add; using lgr::host::mul;
        switch (e.kind) {
        case expr::WIT: coef_add(e.w.id(), r); return e.w.val();
        case expr::ADD: {
            const Fr x = spread(*e.a, r);
            if (e.b->kind == expr::CONST) { const_add(fmul(e.b->k, r)); return add(x, e.b->k); }
            const Fr y = spread(*e.b, r);
            return add(x, y);
        }
        case expr::SUB: {
            if (e.a->kind == expr::CONST) {                   // K - x
                const Fr x = spread(*e.b, neg(r));
                const_add(fmul(e.a->k, r));
                return sub(e.a->k, x);
            }
            const Fr x = spread(*e.a, r);
            if (e.b->kind == expr::CONST) { const_sub(fmul(e.b->k, r)); return sub(x, e.b->k); }
            const Fr y = spread(*e.b, neg(r));
            return sub(x, y);
        }
        case expr::NOT: {                                     // 1 - x
            const Fr x = spread(*e.a, neg(r));
            const_add(r);
            return lgr::host::from_u64(1 - x.v[0]);
        }
        case expr::MUL:
            if (e.b->kind == expr::CONST) {                   // x * K: the leaf takes K*r
                const Fr x = spread(*e.a, fmul(e.b->k, r));
                return fmul(x, e.b->k);
            }
            [[fallthrough]];
        case expr::AND: {                                     // a product inside a linear form: its own witness takes r
            wref z = eval(e);
            const Fr out = z.val();
            coef_add(z.id(), r);
            return out;
        }
        default: throw std::logic_error("wat: a bare constant inside an expression");
        }
    }

    wref duplicate(const wref &w) { return wref(this, clone_raw(w.id())); }
    void assert_const(const wref &w, uint64_t v) { constrain_constant(w.id(), lgr::host::from_u64(v)); }
    void assert_equal(const wref &x, const wref &y) { constrain_equal(x.id(), y.id()); }

    // bit_decompose (core.hpp:714-741) with constrain_bit per bit
    bitvec bit_decompose(const wref &x, size_t nbits) {
        bitvec bits;
        const Fr rho = draw();
        coef_sub(x.id(), rho);
        for (size_t i = 0; i < nbits; i++) {
            const wid b = acquire_raw(lgr::host::from_u64(i < 256 ? (w_[x.id()].val.v[i >> 6] >> (i & 63)) & 1 : 0));
            constrain_bit(b);
            coef_add(b, shl(rho, (int)i));
            bits.push_back(wref(this, b));
        }
        return bits;
    }
    // bit_decompose_constant (:743-759)
    bitvec bit_decompose_constant(uint64_t k, size_t nbits) {
        bitvec bits;
        for (size_t i = 0; i < nbits; i++) {
            const Fr bit = lgr::host::from_u64(i < 64 ? (k >> i) & 1 : 0);
            const wid b = acquire_raw(bit);
            constrain_constant(b, bit);
            bits.push_back(wref(this, b));
        }
        return bits;
    }
    // bit_compose (:761-781); the bits stay with the caller
    wref bit_compose(const bitvec &bits) {
        const wid sum = acquire_raw(zero());
        const Fr rho = draw();
        coef_sub(sum, rho);
        Fr acc = zero();
        for (size_t i = 0; i < bits.size(); i++) {
            acc = lgr::host::add(acc, shl(bits[i].val(), (int)i));
            coef_add(bits[i].id(), shl(rho, (int)i));
        }
        w_[sum].val = acc;
        return wref(this, sum);
    }
    static uint64_t bit_compose_constant(const bitvec &bits) {           // :783-787, low 64 bits
        uint64_t v = 0;
        for (size_t i = 0; i < bits.size() && i < 64; i++) v |= (bits[i].val().v[0] & 1) << i;
        return v;
    }

    wref bitwise_xor(const wref &x, const wref &y);
    wref bitwise_xnor(const wref &x, const wref &y);
    wref bitwise_eqz(const bitvec &x);
    wref bitwise_eq(const bitvec &x, const bitvec &y);
    std::pair<wref, wref> bitwise_gt(const bitvec &x, const bitvec &y, bool is_signed);
    std::pair<wref, wref> idivide_qr(const wref &x, const wref &y);

    // witness_manager::finalize (the mask rows are the prover's business)
    void finish(uint32_t const_sum[8]) {
        pk_.finalize();
        if (const_sum) lgr::host::to_u32(const_sum, const_sum_);
    }
    uint64_t draws() const { return draws_; }
    uint64_t violated() const { return violated_; }
    uint64_t slots_made() const { return nslots_; }
    uint64_t linear_released() const { return nlinear_; }

    static Fr zero() { return Fr{{0, 0, 0, 0}}; }
    // a * b mod p.  Most values of an integer program are below 2^64 (bits, bytes, machine words, small constants): their
    // product is below 2^128 < p and needs no reduction; in the stage-1 run every rho is zero as well
    static Fr fmul(const Fr &a, const Fr &b) {
        if (!(a.v[1] | a.v[2] | a.v[3] | b.v[1] | b.v[2] | b.v[3])) {
            const unsigned __int128 c = (unsigned __int128)a.v[0] * b.v[0];
            return Fr{{(uint64_t)c, (uint64_t)(c >> 64), 0, 0}};
        }
        if (!(a.v[0] | a.v[1] | a.v[2] | a.v[3]) || !(b.v[0] | b.v[1] | b.v[2] | b.v[3])) return zero();
        // 2^i * rho (bit composition spreads one rho over 32 / 64 bits): the Montgomery form of 2^i comes from a table, one multiplication left
        const int ia = pow2_index(a), ib = ia < 0 ? pow2_index(b) : -1;
        if (ia >= 0) return lgr::host::montmul(pow2_mont()[(size_t)ia], b);
        if (ib >= 0) return lgr::host::montmul(pow2_mont()[(size_t)ib], a);
        return lgr::host::mul(a, b);
    }
    static int pow2_index(const Fr &a) {                      // i if a = 2^i (i < 254), else -1
        int at = -1;
        for (int j = 0; j < 4; j++) {
            if (!a.v[j]) continue;
            if (at >= 0 || (a.v[j] & (a.v[j] - 1))) return -1;
            at = 64 * j + __builtin_ctzll(a.v[j]);
        }
        return at < 254 ? at : -1;
    }
    static const std::vector<Fr> &pow2_mont() {
        static const std::vector<Fr> table = [] {
            std::vector<Fr> t;
            Fr x = lgr::host::from_u64(1);
            for (int i = 0; i < 254; i++) { t.push_back(lgr::host::to_mont(x)); x = lgr::host::add(x, x); }
            return t;
        }();
        return table;
    }
    static Fr sub(const Fr &a, const Fr &b) {
        Fr r;
        if (lgr::host::sub4(r.v, a.v, b.v)) lgr::host::add4(r.v, r.v, lgr::host::kP);
        return r;
    }
    static Fr neg(const Fr &a) { return sub(zero(), a); }
    static Fr shl(const Fr &a, int i) {                       // a * 2^i mod p, i < 254
        Fr p2 = zero();
        p2.v[i >> 6] = 1ULL << (i & 63);
        return fmul(a, p2);
    }

private:
    struct wit { Fr val, coef; uint32_t slot; int pos; uint32_t refs; };
    struct slot { wid w[3] = {0, 0, 0}; bool ready[3] = {false, false, false}; };
    row_packer &pk_;
    bool seeded_;
    fr_random_stream rng_;
    std::vector<wit> w_;
    std::vector<slot> slots_;
    std::vector<wid> free_w_;
    std::vector<uint32_t> free_s_;
    Fr const_sum_ = zero();
    uint64_t draws_ = 0, violated_ = 0, nslots_ = 0, nlinear_ = 0;
};

using wexpr = witness_machine::expr;
inline wexpr operator+(wexpr x, wexpr y) { return wexpr(wexpr::ADD, std::move(x), std::move(y)); }
inline wexpr operator-(wexpr x, wexpr y) { return wexpr(wexpr::SUB, std::move(x), std::move(y)); }
inline wexpr operator*(wexpr x, wexpr y) { return wexpr(wexpr::MUL, std::move(x), std::move(y)); }
inline wexpr operator&(wexpr x, wexpr y) { return wexpr(wexpr::AND, std::move(x), std::move(y)); }
inline wexpr operator~(wexpr x) { return wexpr(wexpr::NOT, std::move(x)); }

// the bit gadgets (core.hpp:789-852)
inline witness_machine::wref witness_machine::bitwise_xor(const wref &x, const wref &y) { return eval(x + y - (x & y) * K(2)); }
inline witness_machine::wref witness_machine::bitwise_xnor(const wref &x, const wref &y) { return eval(~(x + y - (x & y) * K(2))); }
inline witness_machine::wref witness_machine::bitwise_eqz(const bitvec &x) {
    wref eqz = eval(~x[0]);
    for (size_t i = 1; i < x.size(); i++) eqz = eval(eqz & ~x[i]);
    return eqz;
}
inline witness_machine::wref witness_machine::bitwise_eq(const bitvec &x, const bitvec &y) {
    wref eq = bitwise_xnor(x[0], y[0]);
    for (size_t i = 1; i < x.size(); i++) eq = eval(eq & bitwise_xnor(x[i], y[i]));
    return eq;
}
inline std::pair<witness_machine::wref, witness_machine::wref> witness_machine::bitwise_gt(const bitvec &x, const bitvec &y, bool is_signed) {
    const size_t msb = x.size() - 1;
    wref gt, eq;
    if (is_signed) gt = eval(~x[msb] & y[msb]);
    else gt = eval(x[msb] & ~y[msb]);
    eq = bitwise_xnor(x[msb], y[msb]);
    for (size_t i = msb; i-- > 0;) {
        wref same = bitwise_xnor(x[i], y[i]);
        gt = eval(gt + (eq & x[i] & ~y[i]));
        eq = eval(eq & same);
    }
    return std::make_pair(std::move(gt), std::move(eq));
}
// idivide_qr (core.hpp:692-712): q, r with q*y + r tied to x (the caller range-checks q and r)
inline std::pair<witness_machine::wref, witness_machine::wref> witness_machine::idivide_qr(const wref &x, const wref &y) {
    const Fr xv = x.val(), yv = y.val();
    if (xv.v[2] | xv.v[3] | yv.v[2] | yv.v[3]) throw std::invalid_argument("wat: division operands beyond 128 bits");
    const unsigned __int128 xn = ((unsigned __int128)xv.v[1] << 64) | xv.v[0], yn = ((unsigned __int128)yv.v[1] << 64) | yv.v[0];
    if (!yn) throw std::invalid_argument("wat: integer divide by zero");
    const unsigned __int128 qn = xn / yn, rn = xn % yn;
    wref q = acquire(Fr{{(uint64_t)qn, (uint64_t)(qn >> 64), 0, 0}});
    wref r = acquire(Fr{{(uint64_t)rn, (uint64_t)(rn >> 64), 0, 0}});
    wref tmp = eval(q * y + r);
    constrain_equal(tmp.id(), x.id());
    return std::make_pair(std::move(q), std::move(r));
}

}  // namespace ligero::cuda::host

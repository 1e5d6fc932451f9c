// Bounded interpreter boundary (BASELINE config 4): a front end for the folded-WAT subset the reference's
// arithmetic tests use (tests/i64_mul.wat, i64_add.wat, i64_sub.wat: imports env.i64_private_const / env.assert_equal,
// one exported function of folded i64.const / i64.mul / i64.add / i64.sub / call forms) and the witness emitter behind it.
// It is NOT the reference's interpreter (include/interpreter_impl.hpp + include/zkp/backend/*.hpp, out of scope): it produces
// a constraint system with the same meaning and feeds it through row_packer -> matrix_prover, so that a .wat goes from text
// to a verifying proof through the product's own entry point.
//
// What each form contributes (cf. include/host_modules/env.hpp:64-77,178-188 and the backend's bit_decompose,
// include/zkp/backend/ligero.hpp:862-890):
//   (call $i64_private_const (i64.const v))   witness x = v, range-checked by a 64-bit decomposition: 64 bit slots b_i with
//                                             b_i * b_i = b_i and the linear constraint x - sum 2^i b_i = 0
//   (i64.mul a b)                             slot (a', b', p) with a' = a, b' = b, p = a*b in the field (< 2^128), p decomposed
//                                             into 128 bits, result = the low 64 bits recomposed (wrap-around of i64.mul)
//   (i64.add a b) / (i64.sub a b)             s = a + b resp. a - b + 2^64 (a constant term: it lands in const_sum), 65-bit
//                                             decomposition, result = low 64 bits
//   (call $assert_equal a b)                  linear constraint a - b = 0
// A quadratic slot holds three field elements (x, y, z) with x*y = z enforced by the quadratic test; "b*b = b" therefore
// also needs x = y and x = z as linear constraints.  Every linear constraint c gets its own random rho_c from the LINEAR
// stream (AES-CTR keyed with the stage-1 seed, nonbatch_context.hpp:105-112); the coefficient of a slot is
// sum_c rho_c * (multiplier of the slot in c), and const_sum = sum_c rho_c * constant_c (zkp/common.hpp:68-79).
//
// Parity statement: the ORDER in which the reference releases witnesses (and so the row order) follows the
// lifetime of C++ temporaries inside its interpreter and cannot be pinned without running it; this emitter releases them in
// creation order.  The proof it leads to is a valid Ligero proof of the same statement, not a byte-identical one.
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../csrc/host_fr.h"
#include "fiat_shamir.hpp"
#include "row_packer.hpp"

namespace ligero::cuda::host {

using lgr::host::Fr;

// ---- S-expressions ------------------------------------------------------------------------------------------
struct sexpr {
    bool is_list = false;
    std::string atom;                 // atoms; string literals keep their quotes
    std::vector<sexpr> list;
    const std::string &head() const { static const std::string none; return (is_list && !list.empty() && !list[0].is_list) ? list[0].atom : none; }
};

class sexpr_parser {
public:
    explicit sexpr_parser(const std::string &text) : s_(text) {}
    sexpr parse_top() {
        skip();
        sexpr e = parse();
        skip();
        if (pos_ != s_.size()) throw std::invalid_argument("wat: trailing text after the module");
        return e;
    }

private:
    void skip() {
        for (;;) {
            while (pos_ < s_.size() && (s_[pos_] == ' ' || s_[pos_] == '\n' || s_[pos_] == '\t' || s_[pos_] == '\r')) pos_++;
            if (pos_ + 1 < s_.size() && s_[pos_] == ';' && s_[pos_ + 1] == ';') { while (pos_ < s_.size() && s_[pos_] != '\n') pos_++; continue; }
            if (pos_ + 1 < s_.size() && s_[pos_] == '(' && s_[pos_ + 1] == ';') {
                int depth = 0;
                while (pos_ + 1 < s_.size()) {
                    if (s_[pos_] == '(' && s_[pos_ + 1] == ';') { depth++; pos_ += 2; }
                    else if (s_[pos_] == ';' && s_[pos_ + 1] == ')') { depth--; pos_ += 2; if (!depth) break; }
                    else pos_++;
                }
                continue;
            }
            return;
        }
    }
    sexpr parse() {
        if (pos_ >= s_.size()) throw std::invalid_argument("wat: unexpected end of text");
        sexpr e;
        if (s_[pos_] == '(') {
            pos_++;
            e.is_list = true;
            for (;;) {
                skip();
                if (pos_ >= s_.size()) throw std::invalid_argument("wat: unbalanced parenthesis");
                if (s_[pos_] == ')') { pos_++; return e; }
                e.list.push_back(parse());
            }
        }
        if (s_[pos_] == ')') throw std::invalid_argument("wat: unexpected ')'");
        const size_t b = pos_;
        if (s_[pos_] == '"') {
            pos_++;
            while (pos_ < s_.size() && s_[pos_] != '"') pos_ += (s_[pos_] == '\\') ? 2 : 1;
            if (pos_ >= s_.size()) throw std::invalid_argument("wat: unterminated string");
            pos_++;
        } else {
            while (pos_ < s_.size() && !strchr(" \n\t\r()", s_[pos_])) pos_++;
        }
        e.atom = s_.substr(b, pos_ - b);
        return e;
    }
    const std::string &s_;
    size_t pos_ = 0;
};

// ---- the constraint system ----------------------------------------------------------------------------------
struct wat_stats {
    uint64_t private_consts = 0, asserts = 0, arithmetic_ops = 0;
    uint64_t linear_witnesses = 0, quadratic_slots = 0, linear_constraints = 0;
    uint64_t violated_constraints = 0;           // > 0: the program's assertions do not hold (the proof will not validate)
};

class constraint_system {
public:
    struct ref { bool quad = false; uint32_t pos = 0; size_t index = 0; };           // a slot: linear[index] or quad[index].{x,y,z}
    struct term { ref slot; Fr mult; };
    struct constraint { std::vector<term> terms; Fr constant{{0, 0, 0, 0}}; };       // sum mult * value(slot) + constant == 0

    ref new_linear(const Fr &v) { lin_.push_back(v); return ref{false, 0, lin_.size() - 1}; }
    size_t new_slot(const Fr &x, const Fr &y, const Fr &z) { quad_.push_back({x, y, z}); return quad_.size() - 1; }
    static ref at(size_t slot, uint32_t pos) { return ref{true, pos, slot}; }
    const Fr &value(const ref &r) const { return r.quad ? quad_[r.index][r.pos] : lin_[r.index]; }
    void equal(const ref &a, const ref &b) { constraint c; c.terms = {{a, one()}, {b, minus_one()}}; cons_.push_back(std::move(c)); }
    void add(constraint c) { cons_.push_back(std::move(c)); }

    // bit slot: (b, b, b) with x = y and x = z; returns the x position
    ref new_bit(uint64_t bit) {
        const Fr b = lgr::host::from_u64(bit);
        const size_t s = new_slot(b, b, b);
        equal(at(s, 0), at(s, 1));
        equal(at(s, 0), at(s, 2));
        return at(s, 0);
    }
    // v (an integer < 2^nbits held in `holder`) = sum 2^i bit_i; returns the bit slots
    std::vector<ref> decompose(const ref &holder, unsigned __int128 v, int nbits) {
        std::vector<ref> bits;
        constraint c;
        c.terms.push_back({holder, one()});
        Fr w = minus_one();                                                         // -(2^i)
        for (int i = 0; i < nbits; i++) {
            bits.push_back(new_bit((uint64_t)((v >> i) & 1)));
            c.terms.push_back({bits.back(), w});
            w = lgr::host::add(w, w);
        }
        cons_.push_back(std::move(c));
        return bits;
    }
    // new linear witness = sum_{i < n} 2^i bits[i]
    ref compose(const std::vector<ref> &bits, int n, uint64_t v) {
        const ref z = new_linear(lgr::host::from_u64(v));
        constraint c;
        c.terms.push_back({z, minus_one()});
        Fr w = one();
        for (int i = 0; i < n; i++) { c.terms.push_back({bits[(size_t)i], w}); w = lgr::host::add(w, w); }
        cons_.push_back(std::move(c));
        return z;
    }

    size_t num_linear() const { return lin_.size(); }
    size_t num_slots() const { return quad_.size(); }
    size_t num_constraints() const { return cons_.size(); }
    uint64_t violated() const {
        uint64_t bad = 0;
        for (const constraint &c : cons_) {
            Fr acc = c.constant;
            for (const term &t : c.terms) acc = lgr::host::add(acc, lgr::host::mul(t.mult, value(t.slot)));
            if (acc.v[0] | acc.v[1] | acc.v[2] | acc.v[3]) bad++;
        }
        for (const auto &q : quad_) if (!(lgr::host::mul(q[0], q[1]) == q[2])) bad++;
        return bad;
    }

    // rows in creation order through the reference's packing rule; with a stage-1 seed also the linear-test coefficients
    void pack(row_packer &pk, const uint8_t *stage1_seed, uint32_t const_sum[8]) const {
        std::vector<Fr> cl(lin_.size(), zero());
        std::vector<std::array<Fr, 3>> cq(quad_.size(), std::array<Fr, 3>{zero(), zero(), zero()});
        Fr cs = zero();
        if (stage1_seed) {
            static const uint8_t any_iv[16] = {0};
            fr_random_stream rng(stage1_seed, any_iv);                                // the linear engine of nonbatch_context.hpp:105-112
            for (const constraint &c : cons_) {
                uint32_t limbs[8];
                rng.next(limbs);
                const Fr rho = lgr::host::to_mont(lgr::host::from_u32(limbs));
                for (const term &t : c.terms) {
                    Fr &dst = t.slot.quad ? cq[t.slot.index][t.slot.pos] : cl[t.slot.index];
                    dst = lgr::host::add(dst, lgr::host::montmul(rho, t.mult));
                }
                cs = lgr::host::add(cs, lgr::host::montmul(rho, c.constant));
            }
        }
        uint32_t v[3][8], c[3][8];
        for (size_t i = 0; i < lin_.size(); i++) {
            lgr::host::to_u32(v[0], lin_[i]); lgr::host::to_u32(c[0], cl[i]);
            pk.push_linear(v[0], c[0]);
        }
        for (size_t i = 0; i < quad_.size(); i++) {
            for (int j = 0; j < 3; j++) { lgr::host::to_u32(v[j], quad_[i][j]); lgr::host::to_u32(c[j], cq[i][j]); }
            pk.push_quadratic(v[0], v[1], v[2], c[0], c[1], c[2]);
        }
        pk.finalize();
        if (const_sum) lgr::host::to_u32(const_sum, cs);
    }

private:
    static Fr zero() { return Fr{{0, 0, 0, 0}}; }
    static Fr one() { return Fr{{1, 0, 0, 0}}; }
    static Fr minus_one() { Fr r; const uint64_t o[4] = {1, 0, 0, 0}; lgr::host::sub4(r.v, lgr::host::kP, o); return r; }
    std::vector<Fr> lin_;
    std::vector<std::array<Fr, 3>> quad_;
    std::vector<constraint> cons_;
};

// ---- front end ------------------------------------------------------------------------------------------------
class wat_program {
public:
    explicit wat_program(const std::string &text) {
        sexpr_parser p(text);
        const sexpr top = p.parse_top();
        if (top.head() != "module") throw std::invalid_argument("wat: expected (module ...)");
        std::string start;
        for (size_t i = 1; i < top.list.size(); i++) {
            const sexpr &f = top.list[i];
            if (f.head() == "import") {
                // (import "env" "name" (func $id ...))
                if (f.list.size() < 4 || f.list[3].head() != "func" || f.list[3].list.size() < 2) throw std::invalid_argument("wat: unsupported import");
                if (f.list[1].atom != "\"env\"") throw std::invalid_argument("wat: only the env host module is supported (no WASI, no bn254fr / vbn254fr imports)");
                imports_[f.list[3].list[1].atom] = unquote(f.list[2].atom);
            } else if (f.head() == "func") {
                if (f.list.size() < 2 || f.list[1].is_list) throw std::invalid_argument("wat: functions must be named");
                funcs_[f.list[1].atom] = &f;
            } else if (f.head() == "export") {
                if (f.list.size() >= 3 && unquote(f.list[1].atom) == "_start" && f.list[2].head() == "func" && f.list[2].list.size() == 2) start = f.list[2].list[1].atom;
            } else {
                throw std::invalid_argument("wat: unsupported module field (" + f.head() + ")");
            }
        }
        if (start.empty() || !funcs_.count(start)) throw std::invalid_argument("wat: no exported _start function");
        module_ = top;                      // keep the tree alive; re-point the function table into the copy
        funcs_.clear();
        for (size_t i = 1; i < module_.list.size(); i++) if (module_.list[i].head() == "func") funcs_[module_.list[i].list[1].atom] = &module_.list[i];
        start_ = start;
    }

    void run(constraint_system &cs, wat_stats &st) const {
        const sexpr &f = *funcs_.at(start_);
        for (size_t i = 2; i < f.list.size(); i++) {
            const std::string &h = f.list[i].head();
            if (h == "param" || h == "result" || h == "local" || h == "type") {
                if (h != "type" && f.list[i].list.size() > 1) throw std::invalid_argument("wat: _start with parameters / locals is not supported");
                continue;
            }
            eval(f.list[i], cs, st);
        }
        st.linear_witnesses = cs.num_linear();
        st.quadratic_slots = cs.num_slots();
        st.linear_constraints = cs.num_constraints();
        st.violated_constraints = cs.violated();
    }

private:
    struct value { bool present = false, witness = false; uint64_t v = 0; constraint_system::ref slot; };
    static std::string unquote(const std::string &s) { return (s.size() >= 2 && s.front() == '"') ? s.substr(1, s.size() - 2) : s; }
    static uint64_t parse_i64(const std::string &t) {
        std::string s;
        for (char ch : t) if (ch != '_') s.push_back(ch);
        bool neg = false;
        size_t i = 0;
        if (i < s.size() && (s[i] == '-' || s[i] == '+')) { neg = s[i] == '-'; i++; }
        if (i >= s.size()) throw std::invalid_argument("wat: bad integer literal " + t);
        unsigned __int128 acc = 0;
        if (s.compare(i, 2, "0x") == 0 || s.compare(i, 2, "0X") == 0) {
            for (i += 2; i < s.size(); i++) {
                const char ch = s[i];
                const int d = (ch >= '0' && ch <= '9') ? ch - '0' : ((ch >= 'a' && ch <= 'f') ? ch - 'a' + 10 : ((ch >= 'A' && ch <= 'F') ? ch - 'A' + 10 : -1));
                if (d < 0) throw std::invalid_argument("wat: bad integer literal " + t);
                acc = acc * 16 + (unsigned)d;
                if (acc >> 64) throw std::invalid_argument("wat: integer literal out of range " + t);
            }
        } else {
            for (; i < s.size(); i++) {
                if (s[i] < '0' || s[i] > '9') throw std::invalid_argument("wat: bad integer literal " + t);
                acc = acc * 10 + (unsigned)(s[i] - '0');
                if (acc >> 64) throw std::invalid_argument("wat: integer literal out of range " + t);
            }
        }
        const uint64_t u = (uint64_t)acc;
        return neg ? (uint64_t)(0 - u) : u;
    }
    static value concrete(uint64_t v) { value r; r.present = true; r.v = v; return r; }
    static value witness(uint64_t v, const constraint_system::ref &slot) { value r; r.present = r.witness = true; r.v = v; r.slot = slot; return r; }

    // i64_private_const (env.hpp:178-188): a fresh witness with a 64-bit range check
    static value private_const(uint64_t v, constraint_system &cs, wat_stats &st) {
        st.private_consts++;
        const constraint_system::ref x = cs.new_linear(lgr::host::from_u64(v));
        cs.decompose(x, v, 64);
        return witness(v, x);
    }
    static value promote(const value &a, constraint_system &cs) {        // a compile-time constant entering a constraint
        if (a.witness) return a;
        const constraint_system::ref x = cs.new_linear(lgr::host::from_u64(a.v));
        constraint_system::constraint c;                                  // x - v = 0 pins it to the public constant
        c.terms.push_back({x, Fr{{1, 0, 0, 0}}});
        const Fr fv = lgr::host::from_u64(a.v);
        const uint64_t z[4] = {0, 0, 0, 0};
        Fr neg; if (a.v) lgr::host::sub4(neg.v, lgr::host::kP, fv.v); else memcpy(neg.v, z, 32);
        c.constant = neg;
        cs.add(std::move(c));
        return witness(a.v, x);
    }

    value binop(const std::string &op, value a, value b, constraint_system &cs, wat_stats &st) const {
        if (!a.present || !b.present) throw std::invalid_argument("wat: " + op + " needs two operands");
        if (!a.witness && !b.witness) {
            return concrete(op == "i64.mul" ? a.v * b.v : (op == "i64.add" ? a.v + b.v : a.v - b.v));
        }
        a = promote(a, cs); b = promote(b, cs);
        st.arithmetic_ops++;
        if (op == "i64.mul") {
            const unsigned __int128 p = (unsigned __int128)a.v * b.v;
            Fr fp{{(uint64_t)p, (uint64_t)(p >> 64), 0, 0}};
            const size_t s = cs.new_slot(lgr::host::from_u64(a.v), lgr::host::from_u64(b.v), fp);
            cs.equal(constraint_system::at(s, 0), a.slot);
            cs.equal(constraint_system::at(s, 1), b.slot);
            const std::vector<constraint_system::ref> bits = cs.decompose(constraint_system::at(s, 2), p, 128);
            return witness((uint64_t)p, cs.compose(bits, 64, (uint64_t)p));
        }
        // add: s = a + b;  sub: s = a - b + 2^64 (never negative)
        const bool sub = op == "i64.sub";
        const unsigned __int128 sv = sub ? ((unsigned __int128)a.v + (((unsigned __int128)1) << 64) - b.v) : ((unsigned __int128)a.v + b.v);
        const constraint_system::ref sslot = cs.new_linear(Fr{{(uint64_t)sv, (uint64_t)(sv >> 64), 0, 0}});
        constraint_system::constraint c;
        Fr m1; const uint64_t o[4] = {1, 0, 0, 0}; lgr::host::sub4(m1.v, lgr::host::kP, o);
        c.terms.push_back({sslot, Fr{{1, 0, 0, 0}}});
        c.terms.push_back({a.slot, m1});
        c.terms.push_back({b.slot, sub ? Fr{{1, 0, 0, 0}} : m1});
        if (sub) { const uint64_t two64[4] = {0, 1, 0, 0}; lgr::host::sub4(c.constant.v, lgr::host::kP, two64); }   // - 2^64
        cs.add(std::move(c));
        const std::vector<constraint_system::ref> bits = cs.decompose(sslot, sv, 65);
        return witness((uint64_t)sv, cs.compose(bits, 64, (uint64_t)sv));
    }

    value eval(const sexpr &e, constraint_system &cs, wat_stats &st) const {
        if (!e.is_list) throw std::invalid_argument("wat: only folded instructions are supported (" + e.atom + ")");
        const std::string &h = e.head();
        if (h == "i64.const") {
            if (e.list.size() != 2) throw std::invalid_argument("wat: i64.const takes one literal");
            return concrete(parse_i64(e.list[1].atom));
        }
        if (h == "i64.mul" || h == "i64.add" || h == "i64.sub") {
            if (e.list.size() != 3) throw std::invalid_argument("wat: " + h + " takes two folded operands");
            const value a = eval(e.list[1], cs, st);
            const value b = eval(e.list[2], cs, st);
            return binop(h, a, b, cs, st);
        }
        if (h == "call") {
            if (e.list.size() < 2) throw std::invalid_argument("wat: call without a target");
            const auto it = imports_.find(e.list[1].atom);
            if (it == imports_.end()) throw std::invalid_argument("wat: call of a non-imported function is not supported (" + e.list[1].atom + ")");
            std::vector<value> args;
            for (size_t i = 2; i < e.list.size(); i++) args.push_back(eval(e.list[i], cs, st));
            if (it->second == "i64_private_const") {
                if (args.size() != 1 || args[0].witness) throw std::invalid_argument("wat: i64_private_const takes one constant");
                return private_const(args[0].v, cs, st);
            }
            if (it->second == "assert_equal") {
                if (args.size() != 2) throw std::invalid_argument("wat: assert_equal takes two operands");
                st.asserts++;
                const value a = promote(args[0], cs), b = promote(args[1], cs);   // make_witness on both sides (env.hpp:68-69)
                cs.equal(a.slot, b.slot);
                return value{};
            }
            throw std::invalid_argument("wat: env." + it->second + " is not supported by the bounded front end");
        }
        throw std::invalid_argument("wat: unsupported instruction " + h);
    }

    sexpr module_;
    std::map<std::string, std::string> imports_;
    std::map<std::string, const sexpr *> funcs_;
    std::string start_;
};

}  // namespace ligero::cuda::host

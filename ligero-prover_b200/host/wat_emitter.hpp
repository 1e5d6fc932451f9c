// Bounded interpreter boundary (BASELINE config 4): a front end for the folded-WAT subset the reference's
// arithmetic tests use (tests/i64_mul.wat, i64_add.wat, i64_sub.wat and their i32 twins: imports env.i64_private_const /
// env.i32_private_const / env.assert_equal, one exported function of folded iNN.const / iNN.mul / iNN.add / iNN.sub / call
// forms) and the witness emitter behind it.
// It is NOT the reference's interpreter (include/interpreter_impl.hpp + include/zkp/backend/*.hpp: a general WASM machine over
// an expression-template backend, out of scope); it is a small witness machine that gives every form of the subset the
// meaning the reference gives it -- which witnesses exist, which linear-test randomness lands on them, and WHEN each one is
// released into a row:
//   (call $i64_private_const (i64.const v))   env.hpp:178-188: witness x = v, 64-bit decomposition (core.hpp:714-741: one
//                                             draw rho, x gets -rho, bit i gets +rho*2^i; every bit b is a slot (b', b'', b)
//                                             with two clones tied to it by one draw each, witness_manager.hpp:431-441);
//                                             x is released when the call returns, the bits travel on the stack
//   (i64.mul a b)                             interpreter_impl.hpp:351-391: operands recomposed (core.hpp:761-781: sum witness,
//                                             one draw, bits get +rho*2^i), slot (x, y, x*y), 128-bit decomposition of the
//                                             product, top 64 bits released at once, then product, y, x -- so the slot lands
//                                             after them -- and only then the operands' bits (a's, then b's, each most
//                                             significant first: the handler's popped stack values die last)
//   (i64.add a b) / (i64.sub a b)             :262-349: s = x + y resp. (2^64 - y) + x with one draw (the constant lands in
//                                             const_sum), 65-bit decomposition, top bit released, then s, y, x
//   (call $assert_equal a b)                  env.hpp:64-77: both sides recomposed, one draw ties them, b's witness is
//                                             released before a's
//   a literal operand                         a bare witness (nonbatch_context.hpp:275-299): no constraint of its own
// Linear-test randomness comes from the LINEAR stream (AES-CTR keyed with the stage-1 seed, nonbatch_context.hpp:105-112),
// drawn in execution order, so the program is run again once the seed exists (as the reference re-runs it in stage 2).
//
// Parity statement: pinned to a run of the reference itself.  tests/refctx/ref_contexts.cpp compiles the reference's own
// interpreter, env module, backend and witness manager and runs programs of this subset through them; on tests/i64_mul.wat
// (tests/golden/refctx_i64_mul_k8192.json), on the repo's mul64.wat and on random expression trees this emitter produces the
// same rows, the same coefficient rows and the same const_sum, element for element (tests/test_refctx_cpu.py).  The
// release order follows the lifetimes of the reference's C++ objects as GCC orders them; another compiler's unspecified
// evaluation order could move witnesses inside a row.
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

// ---- the witness machine ------------------------------------------------------------------------------------
struct wat_stats {
    uint64_t private_consts = 0, asserts = 0, arithmetic_ops = 0;
    uint64_t linear_witnesses = 0, quadratic_slots = 0, linear_constraints = 0;   // linear_constraints = draws from the linear stream
    uint64_t violated_constraints = 0;           // > 0: the program's assertions do not hold (the proof will not validate)
};

class witness_machine {
public:
    using wid = uint32_t;

    // rows go to `pk` as witnesses are released; with a stage-1 seed the linear-test coefficients are drawn as the reference draws them
    witness_machine(row_packer &pk, const uint8_t *stage1_seed) : pk_(pk), seeded_(stage1_seed != nullptr) {
        static const uint8_t any_iv[16] = {0};
        if (seeded_) rng_.init(stage1_seed, any_iv);
    }

    wid acquire(const Fr &v) { w_.push_back(wit{v, zero(), -1, 0}); return (wid)(w_.size() - 1); }
    const Fr &value(wid w) const { return w_[w].val; }

    Fr draw() {                                               // generate_linear_random (witness_manager.hpp:344-348)
        draws_++;
        if (!seeded_) return zero();
        uint32_t limbs[8];
        rng_.next(limbs);
        return lgr::host::from_u32(limbs);
    }
    void coef_add(wid w, const Fr &r) { w_[w].coef = lgr::host::add(w_[w].coef, r); }
    void coef_sub(wid w, const Fr &r) { w_[w].coef = sub(w_[w].coef, r); }
    void const_add(const Fr &r) { const_sum_ = lgr::host::add(const_sum_, r); }

    // constrain_equal (witness_manager.hpp:421-429): one draw, +r on a, -r on b
    void equal(wid a, wid b) {
        if (!(w_[a].val == w_[b].val)) violated_++;
        const Fr r = draw();
        coef_add(a, r);
        coef_sub(b, r);
    }
    wid clone(wid w) { const wid c = acquire(w_[w].val); equal(w, c); return c; }       // clone_witness (:393-397)

    // constrain_quadratic (:474-492): slot positions (a, b, c); a witness that already sits in a slot is replaced by a clone
    void quadratic(wid c, wid a, wid b) {
        if (!(lgr::host::mul(w_[a].val, w_[b].val) == w_[c].val)) violated_++;
        slots_.push_back(slot{});
        const int s = (int)slots_.size() - 1;
        const wid arr[3] = {a, b, c};
        for (int i = 0; i < 3; i++) {
            wid w = arr[i];
            const bool taken = w_[w].slot >= 0;
            if (taken) w = clone(arr[i]);
            w_[w].slot = s; w_[w].pos = i;
            slots_[(size_t)s].w[i] = w;
            if (taken) release(w);
        }
    }

    // commit_release_witness (:117-186)
    void release(wid w) {
        wit &x = w_[w];
        uint32_t v[3][8], c[3][8];
        if (x.slot < 0) {
            lgr::host::to_u32(v[0], x.val); lgr::host::to_u32(c[0], x.coef);
            pk_.push_linear(v[0], c[0]);
            return;
        }
        slot &s = slots_[(size_t)x.slot];
        s.ready[x.pos] = true;
        if (!(s.ready[0] && s.ready[1] && s.ready[2])) return;
        for (int j = 0; j < 3; j++) { lgr::host::to_u32(v[j], w_[s.w[j]].val); lgr::host::to_u32(c[j], w_[s.w[j]].coef); }
        pk_.push_quadratic(v[0], v[1], v[2], c[0], c[1], c[2]);
    }

    // bit_decompose (core.hpp:714-741) with constrain_bit (witness_manager.hpp:431-441) per bit
    std::vector<wid> decompose(wid x, int nbits) {
        const Fr rho = draw();
        coef_sub(x, rho);
        const Fr v = w_[x].val;
        std::vector<wid> bits;
        for (int i = 0; i < nbits; i++) {
            const wid b = acquire(lgr::host::from_u64((v.v[i >> 6] >> (i & 63)) & 1));
            const wid b1 = clone(b), b2 = clone(b);
            quadratic(b, b1, b2);
            release(b1);
            release(b2);
            coef_add(b, shl(rho, i));
            bits.push_back(b);
        }
        return bits;
    }
    // bit_compose (core.hpp:761-781); the caller releases the bits
    wid compose(const std::vector<wid> &bits) {
        const wid sum = acquire(zero());
        const Fr rho = draw();
        coef_sub(sum, rho);
        Fr acc = zero();
        for (size_t i = 0; i < bits.size(); i++) {
            acc = lgr::host::add(acc, shl(w_[bits[i]].val, (int)i));
            coef_add(bits[i], shl(rho, (int)i));
        }
        w_[sum].val = acc;
        return sum;
    }

    // witness_manager::finalize (the mask rows are the prover's business)
    void finish(uint32_t const_sum[8]) {
        pk_.finalize();
        if (const_sum) lgr::host::to_u32(const_sum, const_sum_);
    }
    uint64_t draws() const { return draws_; }
    uint64_t violated() const { return violated_; }

    static Fr zero() { return Fr{{0, 0, 0, 0}}; }
    static Fr sub(const Fr &a, const Fr &b) {
        Fr r;
        if (lgr::host::sub4(r.v, a.v, b.v)) lgr::host::add4(r.v, r.v, lgr::host::kP);
        return r;
    }
    static Fr shl(const Fr &a, int i) {                       // a * 2^i mod p, i < 254
        Fr p2 = zero();
        p2.v[i >> 6] = 1ULL << (i & 63);
        return lgr::host::mul(a, p2);
    }

private:
    struct wit { Fr val, coef; int slot; int pos; };
    struct slot { wid w[3] = {0, 0, 0}; bool ready[3] = {false, false, false}; };
    row_packer &pk_;
    bool seeded_;
    fr_random_stream rng_;
    std::vector<wit> w_;
    std::vector<slot> slots_;
    Fr const_sum_ = zero();
    uint64_t draws_ = 0, violated_ = 0;
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

    // one execution of _start on the machine (rows leave through the machine's packer as witnesses are released)
    void run(witness_machine &m, wat_stats &st) const {
        const sexpr &f = *funcs_.at(start_);
        for (size_t i = 2; i < f.list.size(); i++) {
            const std::string &h = f.list[i].head();
            if (h == "param" || h == "result" || h == "local" || h == "type") {
                if (h != "type" && f.list[i].list.size() > 1) throw std::invalid_argument("wat: _start with parameters / locals is not supported");
                continue;
            }
            value leftover = eval(f.list[i], m, st);
            drop(leftover, m);                                // a value nobody consumed dies with the frame
        }
        st.linear_constraints = m.draws();
        st.violated_constraints = m.violated();
    }

private:
    using wid = witness_machine::wid;
    // a stack value: nothing, a literal, or the bit witnesses of a 64-bit result (decomposed_bits, least significant first)
    struct value { bool present = false, bits = false; uint64_t v = 0; std::vector<wid> b; };   // v is kept reduced to the value's width
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
    static value decomposed(uint64_t v, std::vector<wid> b) { value r; r.present = r.bits = true; r.v = v; r.b = std::move(b); return r; }
    // ~decomposed_bits (core.hpp:100-105): most significant bit first
    static void drop(value &a, witness_machine &m) {
        for (size_t i = a.b.size(); i-- > 0;) m.release(a.b[i]);
        a.b.clear();
    }
    // make_witness (nonbatch_context.hpp:275-299): a literal becomes a bare witness, bits are recomposed.  The bits do
    // NOT die here: decomposed_bits has a destructor and therefore no move constructor, so the popped stack value the
    // opcode handler holds keeps a copy until the handler returns -- after its witnesses, first operand first (sx is
    // declared after sy).
    static wid make_witness(const value &a, witness_machine &m, wat_stats &st) {
        st.linear_witnesses++;
        return a.bits ? m.compose(a.b) : m.acquire(lgr::host::from_u64(a.v));
    }
    // i32_private_const / i64_private_const (env.hpp:166-188): a fresh witness with a range check of its width
    static value private_const(uint64_t v, int width, witness_machine &m, wat_stats &st) {
        st.private_consts++;
        st.linear_witnesses++;
        st.quadratic_slots += (uint64_t)width;
        const wid x = m.acquire(lgr::host::from_u64(v));
        std::vector<wid> bits = m.decompose(x, width);
        m.release(x);
        return decomposed(v, std::move(bits));
    }

    // exec_inn_mul / exec_inn_add / exec_inn_sub on operands of `width` = 32 or 64 bits (interpreter_impl.hpp:262-391)
    value binop(const std::string &op, int width, value a, value b, witness_machine &m, wat_stats &st) const {
        if (!a.present || !b.present) throw std::invalid_argument("wat: " + op + " needs two operands");
        const uint64_t mask = width == 64 ? ~0ULL : 0xFFFFFFFFULL;
        const std::string f = op.substr(4);
        if (!a.bits && !b.bits) {
            return concrete((f == "mul" ? a.v * b.v : (f == "add" ? a.v + b.v : a.v - b.v)) & mask);
        }
        st.arithmetic_ops++;
        const uint64_t av = a.v, bv = b.v;
        const wid x = make_witness(a, m, st), y = make_witness(b, m, st);
        if (f == "mul") {
            const unsigned __int128 p = (unsigned __int128)av * bv;
            const wid z = m.acquire(Fr{{(uint64_t)p, (uint64_t)(p >> 64), 0, 0}});
            m.quadratic(z, x, y);
            st.quadratic_slots += 2 * (uint64_t)width + 1;
            std::vector<wid> bits = m.decompose(z, 2 * width);
            for (int i = 2 * width - 1; i >= width; i--) m.release(bits[(size_t)i]);      // drop_msb(width)
            bits.resize((size_t)width);
            m.release(z); m.release(y); m.release(x);                         // big_result, y, x leave scope in that order,
            drop(a, m); drop(b, m);                                           // then the popped operands: sx (declared last), sy
            return decomposed((uint64_t)p & mask, std::move(bits));
        }
        // add: s = x + y;  sub: s = (2^width - y) + x (never negative)
        const bool sub = f == "sub";
        const unsigned __int128 sv = sub ? ((unsigned __int128)av + (((unsigned __int128)1) << width) - bv) : ((unsigned __int128)av + bv);
        const wid sw = m.acquire(Fr{{(uint64_t)sv, (uint64_t)(sv >> 64), 0, 0}});
        st.linear_witnesses++;
        const Fr r = m.draw();
        m.coef_sub(sw, r);
        if (sub) {                                                            // y takes -r, the constant adds 2^width * r, x takes +r
            m.coef_sub(y, r);
            m.const_add(witness_machine::shl(r, width));
            m.coef_add(x, r);
        } else {
            m.coef_add(x, r);
            m.coef_add(y, r);
        }
        st.quadratic_slots += (uint64_t)width + 1;
        std::vector<wid> bits = m.decompose(sw, width + 1);
        m.release(bits[(size_t)width]);                                       // drop_msb(1)
        bits.resize((size_t)width);
        m.release(sw); m.release(y); m.release(x);
        drop(a, m); drop(b, m);
        return decomposed((uint64_t)sv & mask, std::move(bits));
    }

    value eval(const sexpr &e, witness_machine &m, wat_stats &st) const {
        if (!e.is_list) throw std::invalid_argument("wat: only folded instructions are supported (" + e.atom + ")");
        const std::string &h = e.head();
        if (h == "i64.const" || h == "i32.const") {
            if (e.list.size() != 2) throw std::invalid_argument("wat: " + h + " takes one literal");
            const uint64_t v = parse_i64(e.list[1].atom);
            if (h == "i32.const" && v > 0xFFFFFFFFULL && v < 0xFFFFFFFF80000000ULL) throw std::invalid_argument("wat: integer literal out of range " + e.list[1].atom);
            return concrete(h == "i32.const" ? (v & 0xFFFFFFFFULL) : v);
        }
        if (h == "i64.mul" || h == "i64.add" || h == "i64.sub" || h == "i32.mul" || h == "i32.add" || h == "i32.sub") {
            if (e.list.size() != 3) throw std::invalid_argument("wat: " + h + " takes two folded operands");
            value a = eval(e.list[1], m, st);
            value b = eval(e.list[2], m, st);
            return binop(h, h[1] == '3' ? 32 : 64, std::move(a), std::move(b), m, st);
        }
        if (h == "call") {
            if (e.list.size() < 2) throw std::invalid_argument("wat: call without a target");
            const auto it = imports_.find(e.list[1].atom);
            if (it == imports_.end()) throw std::invalid_argument("wat: call of a non-imported function is not supported (" + e.list[1].atom + ")");
            std::vector<value> args;
            for (size_t i = 2; i < e.list.size(); i++) args.push_back(eval(e.list[i], m, st));
            if (it->second == "i64_private_const" || it->second == "i32_private_const") {
                if (args.size() != 1 || args[0].bits) throw std::invalid_argument("wat: " + it->second + " takes one constant");
                const int width = it->second[1] == '3' ? 32 : 64;
                return private_const(width == 32 ? (args[0].v & 0xFFFFFFFFULL) : args[0].v, width, m, st);
            }
            if (it->second == "assert_equal") {                                // env.hpp:64-77
                if (args.size() != 2) throw std::invalid_argument("wat: assert_equal takes two operands");
                st.asserts++;
                const wid wx = make_witness(args[0], m, st), wy = make_witness(args[1], m, st);
                m.equal(wx, wy);
                m.release(wy); m.release(wx);
                drop(args[0], m); drop(args[1], m);
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

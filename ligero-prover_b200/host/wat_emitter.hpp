// Interpreter boundary (BASELINE config 4): a front end for WebAssembly programs over the env and
// wasi_snapshot_preview1 host modules -- WebAssembly text (folded, as the reference's tests/*.wat are written, or plain) and
// WebAssembly binaries, both of which the reference's prover takes (src/webgpu_prover.cpp:189-207) -- and the witness machine
// behind it.  It stands where the reference's interpreter stands (include/interpreter.hpp, interpreter_impl.hpp over the
// expression-template backend of include/zkp/backend/*.hpp) without being a translation of it.  What it takes: every integer
// instruction the reference implements (interpreter_impl.hpp:155-1309), floating point on numbers (:1314-1853), select, drop,
// nop, locals, i32 / i64 globals, structured control flow, calls of the module's own functions and call_indirect through a
// function table, references and the table instructions (:1926-2106), linear memory (loads, stores, memory.size / grow / fill / copy / init, data segments), the env functions
// iNN_private_const / assert_equal / assert_zero / assert_one / assert_constant / witness_cast / assert_is_concrete
// (host_modules/env.hpp) and wasi args_sizes_get / args_get / fd_write / proc_exit / random_get (host_modules/wasi_preview1.hpp;
// args_get marks the bytes of the private arguments, which is how secret inputs reach a guest).  The other host modules
// (bn254fr, vbn254fr, uint256, ecc) and passive element segments are refused.
// It gives each instruction the meaning the reference gives it: which witnesses exist, which draws of the
// linear stream land on them, and WHEN each one is released into a row.  That order is decided in the reference by C++
// object lifetimes (a witness is committed when the last shared_ptr to it dies), so the machine below is built from counted
// handles with the same copy / move / destruction behaviour and its gadgets mirror where the reference creates, copies and
// drops them; see the comments at witness_machine.
// Linear-test randomness comes from the LINEAR stream (AES-CTR keyed with the stage-1 seed, nonbatch_context.hpp:105-112),
// drawn in execution order, so the program is run again once the seed exists (as the reference re-runs it in stage 2).
//
// Parity statement: pinned to runs of the reference itself.  tests/refctx/ref_contexts.cpp compiles the reference's own
// interpreter, env and WASI modules, backend and witness manager and runs programs through them as token streams; on all 69
// programs of the reference's tests/ (tests/i64_mul.wat = BASELINE config 4 among them), on the repo's mul64.wat / arith32.wat /
// intops.wat / wasi_args.wat (tests/golden/refctx_*.json), indirect.wat and on random programs over the whole instruction set
// (expression trees, locals and functions, memory, control flow, floating point and globals) this emitter produces the same
// rows, the same coefficient rows and the same const_sum, element for element (tests/test_refctx_cpu.py).
// The release order follows the lifetimes of C++ objects as GCC orders them (parameters, temporaries, structured
// bindings); this file leans on the same rules and is built with the same compiler.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <random>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

#include "../csrc/host_fr.h"
#include "fiat_shamir.hpp"
#include "row_packer.hpp"
#include "witness_machine.hpp"

namespace ligero::cuda::host {

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
            if (++depth_ > 2000) throw std::invalid_argument("wat: forms nested too deeply");
            pos_++;
            e.is_list = true;
            for (;;) {
                skip();
                if (pos_ >= s_.size()) throw std::invalid_argument("wat: unbalanced parenthesis");
                if (s_[pos_] == ')') { pos_++; depth_--; return e; }
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
    int depth_ = 0;
};

// ---- front end ------------------------------------------------------------------------------------------------
// A program is a module whose functions are flat instruction lists, read either from WebAssembly text (the folded style of
// the reference's tests/*.wat, plain instruction sequences too) or from a WebAssembly binary (the other input the reference's
// prover takes, src/webgpu_prover.cpp:189-207): type, import, function, table, memory, global, export, element, code and
// data sections.  Execution starts at the exported `_start`.  No wabt on either path.
class wat_program {
public:
    explicit wat_program(const std::string &data) {
        if (data.size() >= 4 && data.compare(0, 4, std::string("\0asm", 4)) == 0) parse_binary(data);
        else parse_text(data);
        if (const char *limit = getenv("LGRP_WAT_STEP_LIMIT")) {   // executed instructions per run (default 2 x 10^8); 0 = no bound
            char *end = nullptr;
            const unsigned long long n = strtoull(limit, &end, 10);
            if (end && !*end && *limit) step_limit_ = n ? n : ~0ULL;
        }
    }

    // one execution of _start on the machine (rows leave through the machine's packer as witnesses are released)
    void run(witness_machine &m, wat_stats &st) const {
        run_state rs{m, st};
        rs.step_limit = step_limit_;
        for (const global_t &g : globals_) rs.globals.push_back(g.init);
        rs.table = table_;
        rs.elems = elem_segs_;
        rs.memory.assign((size_t)mem_pages_ * 65536, 0);
        rs.max_pages = mem_max_;
        for (const data_t &d : datas_) {                      // instantiate (runtime.hpp:537-556): active segments are copied in and dropped
            rs.datas.emplace_back(d.bytes.begin(), d.bytes.end());
            if (!d.active) continue;
            if ((uint64_t)d.offset + d.bytes.size() > rs.memory.size()) throw std::invalid_argument("wat: data segment does not fit the memory");
            std::copy(d.bytes.begin(), d.bytes.end(), rs.memory.begin() + d.offset);
            rs.datas.back().clear();
        }
        exit_code_ = -1;
        const flow done = call(start_, rs, 0);
        if (done.kind == flow::exit) exit_code_ = (int)done.label;
        while (!rs.stack.empty()) rs.stack.pop_back();
        st.linear_constraints = m.draws();
        st.violated_constraints = m.violated();
        st.quadratic_slots = m.slots_made();
        st.linear_witnesses = m.linear_released();
    }
    size_t instructions() const { size_t n = 0; for (const func_t &f : funcs_) n += f.code.size(); return n; }
    // the program's arguments (argv[0] included) and which of them are private: what wasi args_get hands the guest; the bytes of
    // a private argument are marked in memory, so loading them commits witnesses (host_modules/wasi_preview1.hpp:71-100)
    void set_args(std::vector<std::vector<uint8_t>> args, std::set<int> private_indices) { args_ = std::move(args); private_ = std::move(private_indices); }
    void set_echo(bool on) { echo_ = on; }                   // forward what the guest writes to fd 1 / 2 (fd_write) to stdout / stderr
    int exit_code() const { return exit_code_; }              // of the last run: -1 unless the guest called proc_exit
    void set_step_limit(uint64_t n) { step_limit_ = n; }      // loops make running time a property of the program: executed instructions are bounded

private:
    using wref = witness_machine::wref;
    using bitvec = witness_machine::bitvec;
    // stack_value (stack_value.hpp:84-110) restricted to what the integer subset puts on the stack: a native number
    // (tagged i32 / i64 like native_numeric), one witness, or the bit witnesses of a value.  Moving a value moves the
    // witness handle and COPIES the bits (see bitvec), which is what keeps popped operand bits alive until a handler returns.
    struct value;
    // wasm_frame (stack_value.hpp:72-80,261-266): the locals of one activation; the last local dies first.  The frame lives ON
    // the operand stack (as in the reference), below the activation's values, and dies when that slot is dropped
    struct frame_t {
        std::vector<value> locals;
        uint32_t arity = 0;
        ~frame_t() { while (!locals.empty()) locals.pop_back(); }
    };
    struct value {
        enum kind_t : uint8_t { NUM, WIT, BITS, LABEL, FRAME, REF } kind = NUM;
        bool is64 = false;
        bool isf = false;                                     // NUM: a floating-point number (native_numeric's f32 / f64 tags); `num` holds its bits
        uint64_t num = 0;                                     // NUM: the number; LABEL: the arity of the block it closes; REF: a function index (imports first), ~0 = null
        wref wit;
        bitvec bits;
        std::unique_ptr<frame_t> frame;
        value() = default;
        value(value &&) = default;
        value(const value &) = delete;
        // std::variant's move assignment: the same alternative is assigned member-wise (bits element by element), another one
        // destroys what is held first (a witness is dropped, bits die most significant first, a frame takes its locals with
        // it) and then takes the new value
        value &operator=(value &&o) {
            if (kind != o.kind) {
                if (kind == WIT) wit.reset();
                else if (kind == BITS) bits.clear();
                else if (kind == FRAME) frame.reset();
                kind = o.kind;
            }
            is64 = o.is64; isf = o.isf; num = o.num;
            if (kind == WIT) wit = std::move(o.wit);
            else if (kind == BITS) bits = o.bits;
            else if (kind == FRAME) frame = std::move(o.frame);
            return *this;
        }
        static value ref(int64_t f) { value r; r.kind = REF; r.num = f < 0 ? ~0ULL : (uint64_t)f; return r; }     // reference_t (types.hpp:46)
        int64_t as_ref() const { return num == ~0ULL ? -1 : (int64_t)num; }
        static value label(uint32_t arity) { value r; r.kind = LABEL; r.num = arity; return r; }
        static value of(std::unique_ptr<frame_t> f) { value r; r.kind = FRAME; r.frame = std::move(f); return r; }
        // local.get / local.tee (interpreter_impl.hpp:1855-1900): another handle on the same witnesses
        value share() const { value r; r.kind = kind; r.is64 = is64; r.isf = isf; r.num = num; r.wit = wit; r.bits = bits; return r; }
        static value u32(uint32_t v) { value r; r.num = v; return r; }
        static value u64(uint64_t v) { value r; r.is64 = true; r.num = v; return r; }
        static value f32(float v) { value r; r.isf = true; uint32_t b; memcpy(&b, &v, 4); r.num = b; return r; }
        static value f64(double v) { value r; r.isf = r.is64 = true; memcpy(&r.num, &v, 8); return r; }
        float as_f32() const { const uint32_t b = (uint32_t)num; float v; memcpy(&v, &b, 4); return v; }
        double as_f64() const { double v; memcpy(&v, &num, 8); return v; }
        static value of(wref w) { value r; r.kind = WIT; r.wit = std::move(w); return r; }
        static value of(bitvec b) { value r; r.kind = BITS; r.bits = b; return r; }
        uint32_t as_u32() const { return (uint32_t)num; }
        uint64_t as_u64() const { return num; }
    };
    // memory_instance (runtime.hpp:106-176): concrete bytes plus the set of byte ranges a witness was stored to.  A load that
    // touches such a range yields a FRESH witness holding the concrete value (nothing ties it to the stored one -- the
    // reference's model, kept); stores of numbers, memory.fill and memory.init clear the mark, memory.copy carries it along.
    struct secret_ranges {
        std::map<uint32_t, uint32_t> iv;                      // begin -> end, right-open, disjoint, touching ranges joined
        void add(uint32_t b, uint32_t e) {
            if (b >= e) return;
            auto it = iv.lower_bound(b);
            if (it != iv.begin() && std::prev(it)->second >= b) { --it; b = it->first; }
            while (it != iv.end() && it->first <= e) { e = std::max(e, it->second); it = iv.erase(it); }
            iv[b] = e;
        }
        void subtract(uint32_t b, uint32_t e) {
            if (b >= e) return;
            auto it = iv.lower_bound(b);
            if (it != iv.begin() && std::prev(it)->second > b) --it;
            while (it != iv.end() && it->first < e) {
                const uint32_t ib = it->first, ie = it->second;
                it = iv.erase(it);
                if (ib < b) iv[ib] = b;
                if (ie > e) { iv[e] = ie; break; }
            }
        }
        bool intersects(uint32_t b, uint32_t e) const {
            if (b >= e) return false;
            auto it = iv.lower_bound(b);
            if (it != iv.begin() && std::prev(it)->second > b) return true;
            return it != iv.end() && it->first < e;
        }
    };
    struct run_state {
        witness_machine &m;
        wat_stats &st;
        std::vector<value> stack;
        std::vector<uint8_t> memory;
        secret_ranges secrets;
        std::vector<std::vector<uint8_t>> datas;
        uint32_t max_pages = 0;
        std::vector<frame_t *> frames;                        // current_frame() = frames.back()
        std::vector<uint64_t> globals;
        std::vector<int64_t> table;                           // table 0 (table.set / grow / fill / copy / init change it)
        std::vector<std::vector<int64_t>> elems;              // the element segments (elem.drop empties one)
        std::mt19937 rand{1145141919};                        // wasi random_get (wasi_preview1.hpp:47,203-215)
        uint64_t steps = 0, step_limit = 0;
        void push(value v) { stack.push_back(std::move(v)); }
        // drop_n_below (nonbatch_context.hpp:128-138): the `n` values under the top `pos` leave the stack.  First every one of
        // them is handed to destroy_value BY VALUE (:238-247) -- a witness or a frame moves into that parameter and dies
        // there, bits are only copied -- then the range is erased: what is left of it dies as std::vector::erase moves the
        // upper values down (assignment rules of `value`) and destroys the tail
        void drop_n_below(size_t n, size_t pos) {
            if (n + pos > stack.size()) throw std::logic_error("wat: operand stack underflow while unwinding");
            const size_t end = stack.size() - pos, begin = end - n;
            for (size_t i = begin; i < end; i++) {
                value gone(std::move(stack[i]));
                if (gone.kind == value::FRAME) frames.pop_back();
            }
            stack.erase(stack.begin() + (ptrdiff_t)begin, stack.begin() + (ptrdiff_t)end);
        }
        // block_entry (:153-155): the label goes under the block's parameters
        void block_entry(size_t params, size_t arity) {
            if (params > stack.size()) throw std::logic_error("wat: operand stack underflow at a block");
            stack.insert(stack.end() - (ptrdiff_t)params, value::label((uint32_t)arity));
        }
        value pop() {
            if (stack.empty()) throw std::invalid_argument("wat: operand stack underflow");
            value top = std::move(stack.back());
            stack.pop_back();
            return top;
        }
        // nonbatch_context.hpp:249-316
        uint64_t make_numeric(value s) {
            if (s.kind == value::REF) throw std::invalid_argument("wat: a reference where a number is needed (the reference traps: Unexpected stack value)");
            if (s.kind == value::NUM) return s.num;
            if (s.kind == value::WIT) return s.wit.val().v[0];
            return witness_machine::bit_compose_constant(s.bits);
        }
        wref make_witness(value s) {
            if (s.kind == value::REF) throw std::invalid_argument("wat: a reference where a witness is needed (the reference traps: Unexpected stack value)");
            if (s.kind == value::NUM && s.isf) throw std::invalid_argument("wat: a floating-point value where a witness is needed (the reference traps: Unexpected numeric)");
            if (s.kind == value::NUM) return m.acquire(lgr::host::from_u64(s.is64 ? s.as_u64() : s.as_u32()));
            if (s.kind == value::WIT) return std::move(s.wit);
            return m.bit_compose(s.bits);
        }
        bitvec make_decomposed(value s, size_t nbits) {
            if (s.kind == value::REF) throw std::invalid_argument("wat: a reference where a number is needed (the reference traps: Unexpected stack value)");
            if (s.kind == value::NUM) return m.bit_decompose_constant(s.as_u64(), nbits);
            if (s.kind == value::WIT) return m.bit_decompose(s.wit, nbits);
            return s.bits;
        }
    };

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
    static value numeric(bool is64, uint64_t v) { return is64 ? value::u64(v) : value::u32((uint32_t)v); }
    // value types: the integer widths 32 / 64 and, next to them, the two floating-point types
    static constexpr uint8_t F32 = 33, F64 = 65, FUNCREF = 0x70, EXTERNREF = 0x6F;
    static bool is_float(uint8_t t) { return t == F32 || t == F64; }
    static bool is_ref(uint8_t t) { return t == FUNCREF || t == EXTERNREF; }
    static std::string type_name(uint8_t t) { return t == 32 ? "i32" : (t == 64 ? "i64" : (t == F32 ? "f32" : (t == F64 ? "f64" : (t == FUNCREF ? "funcref" : (t == EXTERNREF ? "externref" : "?"))))); }
    static size_t bytes_of(uint8_t t) { return (t == 32 || t == F32) ? 4 : 8; }
    static value of_bits(uint8_t t, uint64_t bits) {         // a number of type `t` from its bit pattern
        value r = numeric(t == 64 || t == F64, bits);
        r.isf = is_float(t);
        return r;
    }
    static int64_t sext(uint64_t v, int w) { return w == 64 ? (int64_t)v : (int64_t)(int32_t)(uint32_t)v; }

    // ---- the integer instructions (interpreter_impl.hpp:155-1309).  Locals are declared in the reference's order: that
    // order is the release order of whatever they still hold when the handler returns.
    enum class op { clz, ctz, popcnt, add, sub, mul, div, rem, and_, or_, xor_, shl, shr, rotl, rotr, eqz, eq, ne, lt, gt, le, ge,
                    extend8, extend16, extend32, extend_i32, wrap };

    // a number in a stack slot, written in place: instructions whose operands are all numbers touch no witness, so they replace the top
    // of the stack instead of popping and pushing values (which carry a witness handle, a bit vector and a frame pointer along)
    static void set_number(value &slot, bool is64, uint64_t v) { slot.num = is64 ? v : (uint64_t)(uint32_t)v; slot.is64 = is64; slot.isf = false; }
    static void unary_number(op o, int w, bool sgn, uint64_t num, uint64_t &r, bool &r64) {
        const uint64_t mask = w == 64 ? ~0ULL : 0xFFFFFFFFULL, x = num & mask;
        r64 = false;
        switch (o) {
        case op::clz: r = x ? (uint32_t)(__builtin_clzll(x) - (64 - w)) : (uint32_t)w; break;
        case op::ctz: r = x ? (uint32_t)__builtin_ctzll(x) : (uint32_t)w; break;
        case op::popcnt: r = (uint32_t)__builtin_popcountll(x); break;
        case op::eqz: r = x == 0; r64 = w == 64; break;
        case op::extend8: r = (uint64_t)(int64_t)(int8_t)x; r64 = w == 64; break;
        case op::extend16: r = w == 64 ? (uint64_t)(uint16_t)x : (uint64_t)(int64_t)(int16_t)x; r64 = w == 64; break;   // (sic: the i64 form zero-extends, :1208)
        case op::extend32: r = (uint64_t)(int64_t)(int32_t)x; r64 = true; break;
        case op::extend_i32: r = sgn ? (uint64_t)(int64_t)(int32_t)(uint32_t)num : (uint64_t)(uint32_t)num; r64 = true; break;
        case op::wrap: r = (uint32_t)num; break;
        default: throw std::logic_error("wat: not a unary instruction");
        }
    }
    static void unary(op o, int w, bool sgn, run_state &rs) {
        witness_machine &m = rs.m;
        if (!rs.stack.empty() && rs.stack.back().kind == value::NUM) {
            uint64_t r; bool r64;
            unary_number(o, w, sgn, rs.stack.back().num, r, r64);
            set_number(rs.stack.back(), r64, r);
            return;
        }
        value sx = rs.pop();
        rs.st.arithmetic_ops++;
        const size_t nb = (size_t)w, msb = nb - 1;
        switch (o) {
        case op::clz: case op::ctz: {                          // :155-227: running "all zero so far" flag, summed
            bitvec bits = rs.make_decomposed(std::move(sx), nb);
            const bool up = o == op::ctz;
            wref acc = m.eval(~bits[up ? 0 : msb]);
            wref cont = m.duplicate(acc);
            for (size_t j = 1; j < nb; j++) {
                const size_t i = up ? j : msb - j;
                cont = m.eval(cont & ~bits[i]);
                acc = m.eval(acc + cont);
            }
            rs.push(value::of(std::move(acc)));
            break;
        }
        case op::popcnt: {                                     // :230-262
            bitvec bits = rs.make_decomposed(std::move(sx), nb);
            wref acc = m.eval(witness_machine::K(0));
            for (size_t i = 0; i < nb; i++) acc = m.eval(acc + bits[i]);
            rs.push(value::of(std::move(acc)));
            break;
        }
        case op::eqz: {                                        // :889-916
            bitvec x = rs.make_decomposed(std::move(sx), nb);
            wref acc = m.eval(~x[0]);
            for (size_t i = 1; i < nb; i++) acc = m.eval(acc & ~x[i]);
            rs.push(value::of(std::move(acc)));
            break;
        }
        case op::extend8: case op::extend16: case op::extend32: {   // :1164-1255: the sign bit cloned upwards
            const size_t from = o == op::extend8 ? 8 : (o == op::extend16 ? 16 : 32);
            bitvec bits = rs.make_decomposed(std::move(sx), nb);
            bits.drop_msb(nb - from);
            for (size_t i = from; i < nb; i++) bits.push_back(m.duplicate(bits[from - 1]));
            rs.push(value::of(bits));
            break;
        }
        case op::extend_i32: {                                 // :1258-1289
            bitvec bits = rs.make_decomposed(std::move(sx), 32);
            if (sgn) {
                for (size_t i = 32; i < 64; i++) bits.push_back(m.duplicate(bits[31]));
            } else {
                wref zero = m.eval(witness_machine::K(0));
                bits.push_msb(zero, 32);
            }
            rs.push(value::of(bits));
            break;
        }
        case op::wrap: {                                       // :1292-1309
            bitvec bits = rs.make_decomposed(std::move(sx), 64);
            bits.drop_msb(32);
            rs.push(value::of(bits));
            break;
        }
        default: throw std::logic_error("wat: not a unary instruction");
        }
    }

    // shl / shr_s / shr_u / rotl / rotr (:706-886): the count is read off its witnesses (no constraint), the bits are re-wired
    static void shift(op o, int w, bool sgn, run_state &rs) {
        witness_machine &m = rs.m;
        if (rs.stack.size() >= 2 && rs.stack.back().kind == value::NUM && rs.stack[rs.stack.size() - 2].kind == value::NUM) {
            const uint32_t n = (uint32_t)rs.stack.back().num % (uint32_t)w;
            rs.stack.pop_back();
            const uint64_t mask = w == 64 ? ~0ULL : 0xFFFFFFFFULL, x = rs.stack.back().num & mask;
            uint64_t r = 0;
            switch (o) {
            case op::shl: r = x << n; break;
            case op::shr: r = sgn ? (uint64_t)(sext(x, w) >> n) : x >> n; break;
            case op::rotl: r = n ? (x << n) | (x >> ((uint32_t)w - n)) : x; break;
            case op::rotr: r = n ? (x >> n) | (x << ((uint32_t)w - n)) : x; break;
            default: throw std::logic_error("wat: not a shift");
            }
            set_number(rs.stack.back(), w == 64, r & mask);
            return;
        }
        value cnt = rs.pop();
        value sx = rs.pop();
        const size_t nb = (size_t)w, msb = nb - 1;
        const uint32_t n = (uint32_t)rs.make_numeric(std::move(cnt)) % (uint32_t)nb;
        if (sx.kind == value::NUM) {
            const uint64_t mask = w == 64 ? ~0ULL : 0xFFFFFFFFULL, x = sx.num & mask;
            uint64_t r = 0;
            switch (o) {
            case op::shl: r = x << n; break;
            case op::shr: r = sgn ? (uint64_t)(sext(x, w) >> n) : x >> n; break;
            case op::rotl: r = n ? (x << n) | (x >> (nb - n)) : x; break;
            case op::rotr: r = n ? (x >> n) | (x << (nb - n)) : x; break;
            default: throw std::logic_error("wat: not a shift");
            }
            rs.push(numeric(w == 64, r & mask));
            return;
        }
        rs.st.arithmetic_ops++;
        bitvec x = rs.make_decomposed(std::move(sx), nb);
        if (o == op::shl) {
            wref zero = m.eval(witness_machine::K(0));
            x.push_lsb(zero, n);
            x.drop_msb(n);
            rs.push(value::of(x));
        } else if (o == op::shr) {
            bitvec unused;
            if (sgn) {
                wref pad = m.duplicate(x[msb]);
                x.drop_lsb(n);
                x.push_msb(pad, n);
            } else {
                wref zero = m.eval(witness_machine::K(0));
                x.drop_lsb(n);
                x.push_msb(zero, n);
            }
            rs.push(value::of(x));
        } else {
            bitvec moved;
            if (o == op::rotl) {
                for (size_t i = 0; i < n; i++) moved.push_back(std::move(x[nb - n + i]));
                for (size_t i = n; i < nb; i++) moved.push_back(std::move(x[i - n]));
            } else {
                for (size_t i = n; i < nb; i++) moved.push_back(std::move(x[i]));
                for (size_t i = 0; i < n; i++) moved.push_back(std::move(x[i]));
            }
            rs.push(value::of(moved));
        }
    }

    static void binary_numbers(op o, int w, bool sgn, uint64_t xin, uint64_t yin, uint64_t &r, bool &r64) {
        const uint64_t mask = w == 64 ? ~0ULL : 0xFFFFFFFFULL, x = xin & mask, y = yin & mask;
        const int64_t xs = sext(x, w), ys = sext(y, w);
        r64 = w == 64;
        switch (o) {
        case op::add: r = (x + y) & mask; break;
        case op::sub: r = (x - y) & mask; break;
        case op::mul: r = (x * y) & mask; break;
        case op::div: case op::rem:
            if (!y) throw std::invalid_argument("wat: integer divide by zero");
            if (sgn) {
                if (ys == -1) r = o == op::div ? (uint64_t)(0 - (uint64_t)xs) : 0;    // (INT_MIN / -1 traps in WASM; wraps here)
                else r = (uint64_t)(o == op::div ? xs / ys : xs % ys);
            } else r = o == op::div ? x / y : x % y;
            r &= mask;
            break;
        case op::and_: r = x & y; break;
        case op::or_: r = x | y; break;
        case op::xor_: r = x ^ y; break;
        case op::eq: r = x == y; r64 = false; break;
        case op::ne: r = x != y; r64 = false; break;
        case op::lt: r = sgn ? xs < ys : x < y; break;
        case op::gt: r = sgn ? xs > ys : x > y; break;
        case op::le: r = sgn ? xs <= ys : x <= y; break;
        case op::ge: r = sgn ? xs >= ys : x >= y; break;
        default: throw std::logic_error("wat: not a binary instruction");
        }
    }
    static void binary(op o, int w, bool sgn, run_state &rs) {
        witness_machine &m = rs.m;
        if (rs.stack.size() >= 2 && rs.stack.back().kind == value::NUM && rs.stack[rs.stack.size() - 2].kind == value::NUM) {
            uint64_t r; bool r64;
            binary_numbers(o, w, sgn, rs.stack[rs.stack.size() - 2].num, rs.stack.back().num, r, r64);
            rs.stack.pop_back();
            set_number(rs.stack.back(), r64, r);
            return;
        }
        value sy = rs.pop();
        value sx = rs.pop();
        const size_t nb = (size_t)w, msb = nb - 1;
        rs.st.arithmetic_ops++;
        switch (o) {
        case op::add: case op::sub: case op::mul: {            // :265-392: the overflowing result, decomposed, top bits dropped
            wref x = rs.make_witness(std::move(sx));
            wref y = rs.make_witness(std::move(sy));
            Fr pow2 = witness_machine::zero();
            pow2.v[nb >> 6] = 1ULL << (nb & 63);               // 2^nb
            wref wide = o == op::add ? m.eval(x + y) : (o == op::sub ? m.eval(witness_machine::K(pow2) - y + x) : m.eval(x * y));
            const size_t extra = o == op::mul ? nb : 1;
            bitvec bits = m.bit_decompose(wide, nb + extra);
            bits.drop_msb(extra);
            rs.push(value::of(bits));
            break;
        }
        case op::div: case op::rem: divide(o, nb, sgn, sx, sy, rs); break;
        case op::and_: case op::or_: case op::xor_: {          // :597-703: bit by bit
            bitvec x = rs.make_decomposed(std::move(sx), nb);
            bitvec y = rs.make_decomposed(std::move(sy), nb);
            bitvec out;
            for (size_t i = 0; i < nb; i++) {
                if (o == op::and_) out.push_back(m.eval(x[i] & y[i]));
                else if (o == op::or_) out.push_back(m.eval(x[i] + y[i] - (x[i] & y[i])));
                else { wref b = m.bitwise_xor(x[i], y[i]); out.push_back(std::move(b)); }
            }
            rs.push(value::of(out));
            break;
        }
        case op::eq: case op::ne: {                            // :919-978
            bitvec x = rs.make_decomposed(std::move(sx), nb);
            bitvec y = rs.make_decomposed(std::move(sy), nb);
            wref r = o == op::eq ? m.bitwise_eq(x, y) : m.eval(~m.bitwise_eq(x, y));
            rs.push(value::of(std::move(r)));
            break;
        }
        case op::lt: case op::gt: case op::le: case op::ge: {  // :981-1161: (gt, eq) scanned from the top bit down
            bitvec x = rs.make_decomposed(std::move(sx), nb);
            bitvec y = rs.make_decomposed(std::move(sy), nb);
            std::pair<wref, wref> ge = m.bitwise_gt(x, y, sgn);
            if (o == op::gt) { rs.push(value::of(std::move(ge.first))); break; }
            wref r = o == op::lt ? m.eval(~(ge.first + ge.second)) : (o == op::le ? m.eval(~ge.first) : m.eval(ge.first + ge.second));
            rs.push(value::of(std::move(r)));
            break;
        }
        default: throw std::logic_error("wat: not a binary instruction");
        }
        (void)msb;
    }

    // div_s / div_u / rem_s / rem_u (:395-594): quotient and remainder as fresh witnesses with q*y + r = x, both range-checked,
    // r < y by the comparison gadget; the signed forms divide absolute values and put the sign back
    static void divide(op o, size_t nb, bool sgn, value &sx, value &sy, run_state &rs) {
        witness_machine &m = rs.m;
        const size_t msb = nb - 1;
        wref x = rs.make_witness(std::move(sx));
        wref y = rs.make_witness(std::move(sy));
        if (!sgn) {
            std::pair<wref, wref> qr = m.idivide_qr(x, y);
            { bitvec range_q = m.bit_decompose(qr.first, nb); }
            bitvec by = m.bit_decompose(y, nb);
            bitvec br = m.bit_decompose(qr.second, nb);
            std::pair<wref, wref> ge = m.bitwise_gt(by, br, sgn);
            m.assert_const(ge.first, 1);
            m.assert_const(ge.second, 0);
            rs.push(value::of(std::move(o == op::div ? qr.first : qr.second)));
            return;
        }
        bitvec bx = m.bit_decompose(x, nb);
        bitvec by = m.bit_decompose(y, nb);
        Fr pow2 = witness_machine::zero();
        pow2.v[nb >> 6] = 1ULL << (nb & 63);
        const auto K = [](const Fr &v) { return witness_machine::K(v); };
        wref abs_x = m.eval(bx[msb] * (K(pow2) - x) + ~bx[msb] * x);
        wref abs_y = m.eval(by[msb] * (K(pow2) - y) + ~by[msb] * y);
        std::pair<wref, wref> qr = m.idivide_qr(abs_x, abs_y);
        { bitvec range_q = m.bit_decompose(qr.first, nb); }
        bitvec abs_y_bits = m.bit_decompose(abs_y, nb);
        bitvec br = m.bit_decompose(qr.second, nb);
        std::pair<wref, wref> ge = m.bitwise_gt(abs_y_bits, br, sgn);
        m.assert_const(ge.first, 1);
        m.assert_const(ge.second, 0);
        if (o == op::div) {
            wref negative = m.bitwise_xor(bx[msb], by[msb]);
            wref ovf_q = m.eval(K(pow2) - qr.first);
            bitvec bneg_q = m.bit_decompose(ovf_q, nb + 1);
            bneg_q.drop_msb(1);
            wref neg_q = m.bit_compose(bneg_q);
            wref res_q = m.eval(negative * neg_q + ~negative * qr.first);
            rs.push(value::of(std::move(res_q)));
        } else {
            wref ovf_r = m.eval(K(pow2) - qr.second);
            bitvec bneg_r = m.bit_decompose(ovf_r, nb + 1);
            bneg_r.drop_msb(1);
            wref neg_r = m.bit_compose(bneg_r);
            wref res_r = m.eval(bx[msb] * neg_r + ~bx[msb] * qr.second);
            rs.push(value::of(std::move(res_r)));
        }
    }

    // ---- floating point (interpreter_impl.hpp:1314-1853): numbers only.  The reference computes on native float / double
    // with the <cmath> functions named below and reads every operand with std::get<native_numeric> -- a witness there ends
    // its run (bad_variant_access), here it is reported.  Nothing of this reaches a row unless a result is converted back
    // to an integer and committed (iNN.trunc_* / reinterpret -> iNN_private_const); that is how the tests observe it.
    enum class fop : uint8_t { abs, neg, ceil, floor, trunc, nearest, sqrt, add, sub, mul, div, min, max, copysign, eq, ne, lt, gt, le, ge,
                               convert, demote, promote, reinterpret, trunc_to_int, trunc_sat };
    static value pop_number(run_state &rs, const char *what) {
        value v = rs.pop();
        if (v.kind != value::NUM) throw std::invalid_argument(std::string("wat: ") + what + " takes concrete operands (a witness here ends the reference's run)");
        return v;
    }
    template <typename F> static F get_float(const value &v) { if constexpr (sizeof(F) == 4) return v.as_f32(); else return v.as_f64(); }
    template <typename F> static value put_float(F v) { if constexpr (sizeof(F) == 4) return value::f32(v); else return value::f64(v); }
    template <typename F> static void float_arith(fop o, run_state &rs) {
        if (o <= fop::sqrt) {
            const F x = get_float<F>(pop_number(rs, "a floating-point instruction"));
            switch (o) {
            case fop::abs: rs.push(put_float<F>(std::fabs(x))); break;
            case fop::neg: rs.push(put_float<F>(-x)); break;
            case fop::ceil: rs.push(put_float<F>(std::ceil(x))); break;
            case fop::floor: rs.push(put_float<F>(std::floor(x))); break;
            case fop::trunc: rs.push(put_float<F>(std::trunc(x))); break;
            case fop::nearest: rs.push(put_float<F>(std::nearbyint(x))); break;
            default: rs.push(put_float<F>(std::sqrt(x))); break;
            }
            return;
        }
        const F y = get_float<F>(pop_number(rs, "a floating-point instruction"));
        const F x = get_float<F>(pop_number(rs, "a floating-point instruction"));
        // two NaN operands: SSE returns the FIRST SOURCE of the instruction, quieted -- and which C++ operand that is, is the
        // compiler's choice for the commutative operations.  The reference's handlers (interpreter_impl.hpp:1478-1522), compiled
        // with GCC 13 at -O1 and at its release -O3 alike, return the FIRST operand's NaN for add, sub, mul and div; this compiler,
        // inlining the same expressions here, commuted add and mul (found by the differential test).  Spelled out, so that it
        // does not depend on how this file is compiled
        if (o <= fop::div && std::isnan(x) && std::isnan(y)) {
            value r = put_float<F>(x);
            r.num |= sizeof(F) == 4 ? 0x00400000ULL : 0x0008000000000000ULL;
            rs.push(std::move(r));
            return;
        }
        switch (o) {
        case fop::add: rs.push(put_float<F>(x + y)); break;
        case fop::sub: rs.push(put_float<F>(x - y)); break;
        case fop::mul: rs.push(put_float<F>(x * y)); break;
        case fop::div: rs.push(put_float<F>(x / y)); break;
        // min / max (:1524-1562): a NaN operand gives the default NaN, otherwise std::fmin / std::fmax -- which in glibc on x86-64
        // are MINSS / MAXSS with the second operand as the source: on a tie the SECOND operand is returned, so max(-0, +0) = +0 and
        // max(+0, -0) = -0 (WebAssembly says +0 for both).  Spelled out, because a compiler that knows no NaN is left may expand the
        // call itself with the operands the other way round
        case fop::min: rs.push(put_float<F>((std::isnan(x) || std::isnan(y)) ? std::numeric_limits<F>::quiet_NaN() : (x < y ? x : y))); break;
        case fop::max: rs.push(put_float<F>((std::isnan(x) || std::isnan(y)) ? std::numeric_limits<F>::quiet_NaN() : (x > y ? x : y))); break;
        case fop::copysign: rs.push(put_float<F>(std::copysign(x, y))); break;
        case fop::eq: rs.push(value::u32(x == y)); break;
        case fop::ne: rs.push(value::u32(x != y)); break;
        case fop::lt: rs.push(value::u32(x < y)); break;
        case fop::gt: rs.push(value::u32(x > y)); break;
        case fop::le: rs.push(value::u32(x <= y)); break;
        case fop::ge: rs.push(value::u32(x >= y)); break;
        default: throw std::logic_error("wat: unknown floating-point instruction");
        }
    }
    // iNN.trunc_fMM_s/u and iNN.trunc_sat_fMM_s/u (:1661-1853): the range tests are the reference's (in double; a float widens exactly)
    static uint64_t float_to_int(double v, int width, bool sgn, bool sat) {
        const double range = sgn ? (width == 32 ? 2147483648.0 : 9223372036854775808.0) : (width == 32 ? 4294967296.0 : 18446744073709551616.0);
        const uint64_t top = width == 32 ? (sgn ? 0x7FFFFFFFULL : 0xFFFFFFFFULL) : (sgn ? 0x7FFFFFFFFFFFFFFFULL : ~0ULL);
        const uint64_t bottom = sgn ? (width == 32 ? 0x80000000ULL : 0x8000000000000000ULL) : 0;
        const bool low = sgn ? v < -range : v <= -1.0;
        if (!sat && (std::isnan(v) || v >= range || low)) throw std::invalid_argument("wat: integer overflow in a float-to-integer conversion");
        if (std::isnan(v)) return 0;
        if (v >= range) return top;
        if (low) return bottom;
        const double t = std::trunc(v);
        if (!sgn) return width == 32 ? (uint64_t)(uint32_t)t : (uint64_t)t;
        return width == 32 ? (uint64_t)(uint32_t)(int32_t)t : (uint64_t)(int64_t)t;
    }
    struct opinfo { op o; int arity; bool sgn; };           // arity 3 marks the shifts / rotates (two operands, the count is read as a number)
    static bool lookup(const std::string &name, opinfo &out) {
        static const std::map<std::string, opinfo> table = {
            {"clz", {op::clz, 1, false}}, {"ctz", {op::ctz, 1, false}}, {"popcnt", {op::popcnt, 1, false}}, {"eqz", {op::eqz, 1, false}},
            {"extend8_s", {op::extend8, 1, true}}, {"extend16_s", {op::extend16, 1, true}}, {"extend32_s", {op::extend32, 1, true}},
            {"extend_i32_s", {op::extend_i32, 1, true}}, {"extend_i32_u", {op::extend_i32, 1, false}}, {"wrap_i64", {op::wrap, 1, false}},
            {"add", {op::add, 2, false}}, {"sub", {op::sub, 2, false}}, {"mul", {op::mul, 2, false}},
            {"div_s", {op::div, 2, true}}, {"div_u", {op::div, 2, false}}, {"rem_s", {op::rem, 2, true}}, {"rem_u", {op::rem, 2, false}},
            {"and", {op::and_, 2, false}}, {"or", {op::or_, 2, false}}, {"xor", {op::xor_, 2, false}},
            {"shl", {op::shl, 3, false}}, {"shr_s", {op::shr, 3, true}}, {"shr_u", {op::shr, 3, false}}, {"rotl", {op::rotl, 3, false}}, {"rotr", {op::rotr, 3, false}},
            {"eq", {op::eq, 2, false}}, {"ne", {op::ne, 2, false}},
            {"lt_s", {op::lt, 2, true}}, {"lt_u", {op::lt, 2, false}}, {"gt_s", {op::gt, 2, true}}, {"gt_u", {op::gt, 2, false}},
            {"le_s", {op::le, 2, true}}, {"le_u", {op::le, 2, false}}, {"ge_s", {op::ge, 2, true}}, {"ge_u", {op::ge, 2, false}},
        };
        const auto it = table.find(name);
        if (it == table.end()) return false;
        out = it->second;
        return true;
    }

    // env host functions (include/host_modules/env.hpp:40-110,160-190)
    enum class host_fn : uint8_t { i32_private_const, i64_private_const, assert_equal, assert_zero, assert_one, assert_constant, witness_cast, assert_is_concrete,
                                   wasi_args_sizes_get, wasi_args_get, wasi_fd_write, wasi_proc_exit, wasi_random_get, print_str, dump_memory, unsupported };
    static bool host_lookup(const std::string &name, host_fn &out) {
        static const std::map<std::string, host_fn> table = {
            {"i32_private_const", host_fn::i32_private_const}, {"i64_private_const", host_fn::i64_private_const}, {"assert_equal", host_fn::assert_equal},
            {"assert_zero", host_fn::assert_zero}, {"assert_one", host_fn::assert_one}, {"assert_constant", host_fn::assert_constant},
            {"witness_cast_u32", host_fn::witness_cast}, {"witness_cast_u64", host_fn::witness_cast}, {"assert_is_concrete", host_fn::assert_is_concrete},
            {"print_str", host_fn::print_str}, {"dump_memory", host_fn::dump_memory},
        };
        const auto it = table.find(name);
        if (it == table.end()) return false;
        out = it->second;
        return true;
    }
    static void host(host_fn f, run_state &rs) {
        witness_machine &m = rs.m;
        switch (f) {
        case host_fn::i32_private_const: case host_fn::i64_private_const: {   // a fresh witness, range-checked by its decomposition
            const int width = f == host_fn::i32_private_const ? 32 : 64;
            const value lit = rs.pop();
            if (lit.kind != value::NUM) throw std::invalid_argument(std::string("wat: ") + (width == 32 ? "i32" : "i64") + "_private_const takes one constant");
            rs.st.private_consts++;
            wref x = m.acquire(lgr::host::from_u64(width == 32 ? lit.as_u32() : lit.as_u64()));
            bitvec checked = m.bit_decompose(x, (size_t)width);
            rs.push(value::of(checked));
            break;
        }
        case host_fn::assert_equal: {
            rs.st.asserts++;
            value sy = rs.pop();
            value sx = rs.pop();
            wref wx = rs.make_witness(std::move(sx));
            wref wy = rs.make_witness(std::move(sy));
            m.assert_equal(wx, wy);
            break;
        }
        case host_fn::assert_zero: case host_fn::assert_one: {
            rs.st.asserts++;
            value s = rs.pop();
            wref w = rs.make_witness(std::move(s));
            m.assert_const(w, f == host_fn::assert_one ? 1 : 0);
            break;
        }
        case host_fn::assert_constant: {                       // ties the witness to the value it holds
            rs.st.asserts++;
            value s = rs.pop();
            wref w = rs.make_witness(std::move(s));
            m.constrain_constant(w.id(), w.val());
            break;
        }
        case host_fn::witness_cast: {                          // the value as ONE witness (the stack keeps a second handle)
            value sx = rs.pop();
            wref wx = rs.make_witness(std::move(sx));
            rs.push(value::of(wx));
            break;
        }
        case host_fn::assert_is_concrete: {
            value s = rs.pop();
            if (s.kind != value::NUM) throw std::invalid_argument("wat: assert_is_concrete: value is a witness");
            break;
        }
        default: throw std::logic_error("wat: not an env function");
        }
    }

    struct ins {
        enum kind_t : uint8_t { konst, unary_op, shift_op, binary_op, host_call, func_call, local_get, local_set, local_tee, select, drop, nop,
                                load, store, memory_size, memory_grow, memory_fill, memory_copy, memory_init, data_drop,
                                block, loop, if_, else_, end, br, br_if, br_table, return_, unreachable, float_op, global_get, global_set, call_indirect,
                                ref_null, ref_is_null, ref_func, table_get, table_set, table_size, table_grow, table_fill, table_copy, table_init, elem_drop } kind;
        uint8_t o = 0;                                        // op, fop or host_fn; bytes moved by a load / store
        uint8_t width = 0;                                    // the value type the instruction computes in (32, 64, F32, F64); conversions: of the result
        bool sgn = false;
        uint64_t imm = 0;                                     // literal, local / function / data index (module functions from 0), memory offset,
                                                              // branch depth / table; block, loop, if: index of the matching `end`, o = parameters, width = results
        uint32_t aux = 0;                                     // if: index of its `else` (of its `end` when there is none); conversions: the source type
    };
    // a module function: signature, locals (parameters first), flat body
    struct func_t {
        std::vector<uint8_t> params, results, locals;         // widths (32 / 64); locals = params + declared locals
        std::vector<ins> code;
        std::vector<std::vector<uint32_t>> tables;            // br_table targets, the default last
    };

    static void floating(const ins &i, run_state &rs) {
        const fop o = (fop)i.o;
        if (o <= fop::ge) {
            if (i.width == F32) float_arith<float>(o, rs); else float_arith<double>(o, rs);
            return;
        }
        const value x = pop_number(rs, "a conversion to or from floating point");
        switch (o) {
        case fop::convert:                                    // fNN.convert_iMM_s/u (:1577-1629)
            if (i.width == F32) rs.push(value::f32(i.aux == 32 ? (i.sgn ? static_cast<float>(static_cast<int32_t>(x.as_u32())) : static_cast<float>(x.as_u32()))
                                                               : (i.sgn ? static_cast<float>(static_cast<int64_t>(x.as_u64())) : static_cast<float>(x.as_u64()))));
            else rs.push(value::f64(i.aux == 32 ? (i.sgn ? static_cast<double>(static_cast<int32_t>(x.as_u32())) : static_cast<double>(x.as_u32()))
                                                : (i.sgn ? static_cast<double>(static_cast<int64_t>(x.as_u64())) : static_cast<double>(x.as_u64()))));
            break;
        case fop::demote: rs.push(value::f32(static_cast<float>(x.as_f64()))); break;
        case fop::promote: rs.push(value::f64(static_cast<double>(x.as_f32()))); break;
        case fop::reinterpret: rs.push(of_bits(i.width, bytes_of(i.width) == 4 ? (x.num & 0xFFFFFFFFULL) : x.num)); break;
        case fop::trunc_to_int: case fop::trunc_sat:
            rs.push(numeric(i.width == 64, float_to_int(i.aux == F32 ? (double)x.as_f32() : x.as_f64(), i.width, i.sgn, o == fop::trunc_sat)));
            break;
        default: throw std::logic_error("wat: unknown floating-point instruction");
        }
    }

    // exec_result (types.hpp:53-86): how an instruction ended -- fell through, jumps `label` more blocks out, or returns
    struct flow {
        enum { ok, jump, ret, exit } kind = ok;               // exit: the guest called proc_exit (exec_exit); `label` carries the code
        uint32_t label = 0;
        bool unwind() {                                       // one block left behind; true when the jump ends here
            if (kind != jump) return false;
            if (label == 0) { kind = ok; return true; }
            --label;
            return false;
        }
    };
    // run_call (interpreter.hpp:274-345): arguments become the first locals, declared locals start at zero, the frame goes on
    // the stack; when the body is through it is dropped from under the results -- and with it whatever the locals still
    // hold, last local first.  `return` has dropped it already
    flow call(size_t fi, run_state &rs, int depth) const {
        if (depth > 200) throw std::invalid_argument("wat: call depth exceeded");
        const func_t &f = funcs_[fi];
        std::vector<value> arguments;
        for (size_t i = 0; i < f.params.size(); i++) arguments.emplace_back(rs.pop());
        std::reverse(arguments.begin(), arguments.end());
        auto frame = std::make_unique<frame_t>();
        frame->arity = (uint32_t)f.results.size();
        frame->locals = std::move(arguments);
        for (size_t i = f.params.size(); i < f.locals.size(); i++) frame->locals.emplace_back(of_bits(f.locals[i], 0));
        rs.frames.push_back(frame.get());
        rs.push(value::of(std::move(frame)));
        const size_t base = rs.stack.size();
        for (size_t pc = 0; pc < f.code.size(); pc++) {
            const flow r = step(f, pc, rs, depth, base);
            if (r.kind == flow::ret) return flow{};           // (a jump that leaves the body is not caught by the reference either; the validator rejects it)
            if (r.kind == flow::exit) {                       // (:338-349) everything down to and including this activation's frame leaves, then the caller unwinds
                size_t distance = 0;
                while (distance < rs.stack.size() && rs.stack[rs.stack.size() - 1 - distance].kind != value::FRAME) distance++;
                if (distance < rs.stack.size()) rs.drop_n_below(distance + 1, 0);
                return r;
            }
        }
        rs.drop_n_below(1, f.results.size());
        return flow{};
    }
    // wasi_snapshot_preview1 (include/host_modules/wasi_preview1.hpp): the functions a guest needs to receive its arguments, print
    // and stop.  Pointers and lengths are read with as_u32() in the reference, i.e. they must be numbers
    static uint32_t wasi_u32(run_state &rs) {
        value v = rs.pop();
        if (v.kind != value::NUM) throw std::invalid_argument("wat: a WASI call takes concrete operands (a witness here ends the reference's run)");
        return v.as_u32();
    }
    static uint8_t *wasi_mem(run_state &rs, uint64_t at, uint64_t n) {
        if (at + n > rs.memory.size()) throw std::invalid_argument("wat: a WASI call reaches outside the memory");
        return rs.memory.data() + at;
    }
    static void wasi_put32(run_state &rs, uint32_t at, uint32_t v) { memcpy(wasi_mem(rs, at, 4), &v, 4); }
    flow host_call(host_fn f, run_state &rs) const {
        switch (f) {
        case host_fn::wasi_args_sizes_get: {                  // (:52-69) argc and the total size of the argument bytes
            const uint32_t size_ptr = wasi_u32(rs), count_ptr = wasi_u32(rs);
            uint32_t total = 0;
            for (const auto &a : args_) total += (uint32_t)a.size();
            wasi_put32(rs, count_ptr, (uint32_t)args_.size());
            wasi_put32(rs, size_ptr, total);
            rs.push(value::u32(0));
            break;
        }
        case host_fn::wasi_args_get: {                        // (:71-100) the pointers, the bytes, and the mark on every private argument
            uint32_t buffer = wasi_u32(rs), argv = wasi_u32(rs);
            for (size_t i = 0; i < args_.size(); i++) {
                wasi_put32(rs, argv, buffer);
                argv += 4;
                if (!args_[i].empty()) memcpy(wasi_mem(rs, buffer, args_[i].size()), args_[i].data(), args_[i].size());
                if (private_.count((int)i)) rs.secrets.add(buffer, buffer + (uint32_t)args_[i].size());
                buffer += (uint32_t)args_[i].size();
            }
            rs.push(value::u32(0));
            break;
        }
        case host_fn::wasi_fd_write: {                        // (:167-193) writev; only stdout / stderr exist here
            const uint32_t nwrite_ptr = wasi_u32(rs), iovec_len = wasi_u32(rs), iovec_ptr = wasi_u32(rs), fd = wasi_u32(rs);
            wasi_mem(rs, nwrite_ptr, 4);
            if (iovec_len > 1024) throw std::invalid_argument("wat: fd_write with too many buffers");
            uint64_t written = 0;
            std::string text;
            for (uint32_t i = 0; i < iovec_len; i++) {
                uint32_t ptr, len;
                memcpy(&ptr, wasi_mem(rs, (uint64_t)iovec_ptr + 8 * (uint64_t)i, 8), 4);
                memcpy(&len, wasi_mem(rs, (uint64_t)iovec_ptr + 8 * (uint64_t)i + 4, 4), 4);
                text.append((const char *)wasi_mem(rs, ptr, len), len);
                written += len;
            }
            const bool known = fd == 1 || fd == 2;
            if (known && echo_) { fwrite(text.data(), 1, text.size(), fd == 1 ? stdout : stderr); fflush(fd == 1 ? stdout : stderr); }
            rs.push(value::u32(known ? 0 : 9));               // errno of a failed writev (EBADF), as the reference reports it
            wasi_put32(rs, nwrite_ptr, known ? (uint32_t)written : 0xFFFFFFFFu);
            break;
        }
        case host_fn::wasi_proc_exit: {                       // (:195-200) ends the run: every activation is unwound (call)
            value code = rs.pop();
            flow r; r.kind = flow::exit; r.label = (uint32_t)rs.make_numeric(std::move(code));
            return r;
        }
        case host_fn::wasi_random_get: {                      // (:202-215) NOT random: mt19937 seeded with a constant, one draw per byte
            const uint32_t len = wasi_u32(rs), ptr = wasi_u32(rs);
            uint8_t *buf = wasi_mem(rs, ptr, len);
            std::uniform_int_distribution<> dist(0, 255);
            for (uint32_t i = 0; i < len; i++) buf[i] = (uint8_t)dist(rs.rand);
            rs.push(value::u32(0));
            break;
        }
        case host_fn::print_str: case host_fn::dump_memory: {  // env.print_str / dump_memory (host_modules/env.hpp:92-126): (ptr, len) of linear memory to stdout, raw or as hex
            const uint32_t len = wasi_u32(rs), ptr = wasi_u32(rs);
            const uint8_t *bytes = wasi_mem(rs, ptr, len);
            if (echo_) {
                if (f == host_fn::print_str) fwrite(bytes, 1, len, stdout);
                else { fputs("@dump: ", stdout); for (uint32_t i = 0; i < len; i++) printf("%02X", bytes[i]); fputc('\n', stdout); }
                fflush(stdout);
            }
            break;
        }
        case host_fn::unsupported: throw std::logic_error("wat: unresolved import");   // (step() reports these by name)
        default: host(f, rs); break;
        }
        return flow{};
    }
    // the body of a block / of one arm of an if (run_scoped_block, run_if_then_else: interpreter.hpp:91-131,175-214)
    flow run_body(const func_t &f, size_t lo, size_t hi, size_t results, run_state &rs, int depth, size_t base) const {
        for (size_t pc = lo; pc < hi; pc++) {
            flow r = step(f, pc, rs, depth, base);
            if (r.kind != flow::ok) { r.unwind(); return r; }
        }
        rs.drop_n_below(1, results);                          // the label leaves from under the results
        return flow{};
    }
    // one instruction; `pc` is left on its last index (the `end` of a block)
    flow step(const func_t &f, size_t &pc, run_state &rs, int depth, size_t base) const {
        const ins &i = f.code[pc];
        if (++rs.steps > rs.step_limit) throw std::invalid_argument("wat: step budget exceeded (" + std::to_string(rs.step_limit) + " instructions)");
        switch (i.kind) {
        case ins::konst: {
            rs.stack.emplace_back();
            value &v = rs.stack.back();
            v.num = i.imm; v.is64 = i.width == 64 || i.width == F64; v.isf = is_float(i.width);
            break;
        }
        case ins::float_op: floating(i, rs); break;
        case ins::global_get: rs.push(of_bits(globals_[(size_t)i.imm].type, rs.globals[(size_t)i.imm])); break;   // exec_global_get / set (interpreter_impl.hpp:1902-1924):
        case ins::global_set: rs.globals[(size_t)i.imm] = pop_number(rs, "global.set").num; break;                // a global holds a native number
        case ins::unary_op: unary((op)i.o, i.width, i.sgn, rs); break;
        case ins::shift_op: shift((op)i.o, i.width, i.sgn, rs); break;
        case ins::binary_op: binary((op)i.o, i.width, i.sgn, rs); break;
        case ins::host_call:
            // imports resolve when they are CALLED in the reference (nonbatch_context.hpp:201-208): a module may import what the front end
            // does not provide as long as it does not call it
            if ((host_fn)i.o == host_fn::unsupported) throw std::invalid_argument("wat: call of " + unprovided_[(size_t)i.imm] + ", which the front end does not provide");
            return host_call((host_fn)i.o, rs);
        case ins::func_call: return call((size_t)i.imm, rs, depth + 1);
        case ins::call_indirect: {                            // run_call_indirect (interpreter.hpp:372-398): the index is read with as_u32(), i.e. it must be a number
            value v = rs.pop();
            if (v.kind != value::NUM) throw std::invalid_argument("wat: call_indirect takes a concrete index (a witness here ends the reference's run)");
            const uint32_t at = v.as_u32();
            if (at >= rs.table.size()) throw std::invalid_argument("wat: call_indirect: index out of bound");
            if (rs.table[at] < 0) throw std::invalid_argument("wat: call_indirect: null pointer");
            if ((uint64_t)rs.table[at] < nimports_) throw std::invalid_argument("wat: an indirect call of an imported function is not supported");
            const size_t target = (size_t)rs.table[at] - nimports_;
            const func_t &callee = funcs_[target];
            const sig_t &want_sig = sigs_[(size_t)i.imm];       // (the reference leaves this check as a TODO; a mismatch would misread its stack.  WebAssembly traps)
            if (callee.params != want_sig.params || callee.results != want_sig.results) throw std::invalid_argument("wat: call_indirect: indirect call type mismatch");
            return call(target, rs, depth + 1);
        }
        case ins::ref_null: case ins::ref_is_null: case ins::ref_func: case ins::table_get: case ins::table_set: case ins::table_size: case ins::table_grow:
        case ins::table_fill: case ins::table_copy: case ins::table_init: case ins::elem_drop:
            reference_op(i, rs);
            break;
        case ins::local_get: {
            const value &l = rs.frames.back()->locals[(size_t)i.imm];
            if (l.kind != value::NUM) { rs.push(l.share()); break; }
            rs.stack.emplace_back();
            value &v = rs.stack.back();
            v.num = l.num; v.is64 = l.is64; v.isf = l.isf;
            break;
        }
        case ins::local_set: case ins::local_tee: {
            value &l = rs.frames.back()->locals[(size_t)i.imm];
            if (rs.stack.empty()) throw std::invalid_argument("wat: operand stack underflow");
            value &top = rs.stack.back();
            if (l.kind == value::NUM && top.kind == value::NUM) {              // number over number: member-wise, nothing is released
                l.num = top.num; l.is64 = top.is64; l.isf = top.isf;
                if (i.kind == ins::local_set) rs.stack.pop_back();
            } else if (i.kind == ins::local_set) l = rs.pop();
            else l = top.share();
            break;
        }
        case ins::select: select(rs); break;
        case ins::drop: rs.pop(); break;                      // exec_drop (interpreter_impl.hpp:112-116)
        case ins::nop: break;
        case ins::load: load(i, rs); break;
        case ins::store: store(i, rs); break;
        case ins::block: {
            const size_t end = (size_t)i.imm;
            rs.block_entry(i.o, i.width);
            const flow r = run_body(f, pc + 1, end, i.width, rs, depth, base);
            pc = end;
            return r;
        }
        case ins::loop: {                                     // run_loop (:133-173): the label is re-entered on every round, with the parameters as its arity
            const size_t end = (size_t)i.imm;
            for (;;) {
                rs.block_entry(i.o, i.o);
                bool again = false;
                for (size_t q = pc + 1; q < end; q++) {
                    flow r = step(f, q, rs, depth, base);
                    if (r.kind == flow::ok) continue;
                    if (r.unwind()) { again = true; break; }
                    pc = end;
                    return r;
                }
                if (again) continue;
                rs.drop_n_below(1, i.width);
                pc = end;
                return flow{};
            }
        }
        case ins::if_: {
            const size_t end = (size_t)i.imm, els = (size_t)i.aux;
            value tmp = rs.pop();
            const uint32_t c = (uint32_t)rs.make_numeric(std::move(tmp));
            rs.block_entry(i.o, i.width);
            const flow r = c ? run_body(f, pc + 1, els, i.width, rs, depth, base) : run_body(f, std::min(els + 1, end), end, i.width, rs, depth, base);
            pc = end;
            return r;
        }
        case ins::br: return branch((uint32_t)i.imm, rs);
        case ins::br_if: {                                    // run_br_if (:232-240)
            value v = rs.pop();
            const uint32_t cond = (uint32_t)rs.make_numeric(std::move(v));
            if (cond) return branch((uint32_t)i.imm, rs);
            break;
        }
        case ins::br_table: {                                 // run_br_table (:242-253)
            value v = rs.pop();
            const uint32_t k = (uint32_t)rs.make_numeric(std::move(v));
            const std::vector<uint32_t> &t = f.tables[(size_t)i.imm];
            return branch(k + 1 < t.size() ? t[k] : t.back(), rs);
        }
        case ins::return_: {                                  // run_return (:255-272): everything down to and including the frame leaves, the results stay
            const size_t arity = rs.frames.back()->arity;
            size_t distance = 0;
            while (distance < rs.stack.size() && rs.stack[rs.stack.size() - 1 - distance].kind != value::FRAME) distance++;
            if (distance >= rs.stack.size() || distance < arity) throw std::logic_error("wat: return without a frame");
            rs.drop_n_below(distance - arity + 1, arity);
            flow r; r.kind = flow::ret;
            return r;
        }
        case ins::unreachable: throw std::invalid_argument("wat: unreachable executed");
        case ins::else_: case ins::end: throw std::logic_error("wat: stray block delimiter");
        default: bulk_memory(i, rs); break;
        }
        return flow{};
    }
    // run_br (interpreter.hpp:216-230): find the label `l` blocks out, drop it and everything above it except its arity
    static flow branch(uint32_t l, run_state &rs) {
        size_t distance = 0, seen = 0;
        for (;; distance++) {
            if (distance >= rs.stack.size()) throw std::logic_error("wat: branch target not on the stack");
            if (rs.stack[rs.stack.size() - 1 - distance].kind == value::LABEL && seen++ == l) break;
        }
        const size_t n = (size_t)rs.stack[rs.stack.size() - 1 - distance].num;
        if (distance < n) throw std::logic_error("wat: branch without its operands");
        rs.drop_n_below(distance - n + 1, n);
        flow r; r.kind = flow::jump; r.label = l;
        return r;
    }
    // references and the table (interpreter_impl.hpp:1926-2106): a reference is a function index or null; indices and counts are
    // read with as_u32(), references with as_ref() -- anything else on the stack ends the reference's run
    static int64_t pop_ref(run_state &rs, const char *what) {
        value v = rs.pop();
        if (v.kind != value::REF) throw std::invalid_argument(std::string("wat: ") + what + " takes a reference");
        return v.as_ref();
    }
    static void reference_op(const ins &i, run_state &rs) {
        const auto index = [&](const char *what) { return pop_number(rs, what).as_u32(); };
        std::vector<int64_t> &tab = rs.table;
        switch (i.kind) {
        case ins::ref_null: rs.push(value::ref(-1)); break;
        case ins::ref_is_null: rs.push(value::u32(pop_ref(rs, "ref.is_null") < 0)); break;
        case ins::ref_func: rs.push(value::ref((int64_t)i.imm)); break;
        case ins::table_get: {
            const uint32_t at = index("table.get");
            if (at >= tab.size()) throw std::invalid_argument("wat: table_get: index out of range");
            rs.push(value::ref(tab[at]));
            break;
        }
        case ins::table_set: {
            const int64_t r = pop_ref(rs, "table.set");
            const uint32_t at = index("table.set");
            if (at >= tab.size()) throw std::invalid_argument("wat: table_set: index out of range");
            tab[at] = r;
            break;
        }
        case ins::table_size: rs.push(value::u32((uint32_t)tab.size())); break;
        case ins::table_grow: {                               // (:1997-2019) no maximum is consulted: it grows until memory runs out (capped here)
            const uint32_t n = index("table.grow");
            const int64_t r = pop_ref(rs, "table.grow");
            if ((uint64_t)tab.size() + n > 1000000) { rs.push(value::u32(0xFFFFFFFFu)); break; }
            rs.push(value::u32((uint32_t)tab.size()));
            tab.insert(tab.end(), n, r);
            break;
        }
        case ins::table_fill: {
            const uint32_t n = index("table.fill");
            const int64_t r = pop_ref(rs, "table.fill");
            const uint32_t at = index("table.fill");
            if ((uint64_t)at + n > tab.size()) throw std::invalid_argument("wat: table_fill: index out of bound");
            std::fill_n(tab.begin() + at, n, r);
            break;
        }
        case ins::table_copy: {
            const uint32_t n = index("table.copy"), src = index("table.copy"), dst = index("table.copy");
            if ((uint64_t)src + n > tab.size() || (uint64_t)dst + n > tab.size()) throw std::invalid_argument("wat: table_copy: index out of bound");
            if (dst <= src) std::copy(tab.begin() + src, tab.begin() + src + n, tab.begin() + dst);
            else std::copy_backward(tab.begin() + src, tab.begin() + src + n, tab.begin() + dst + n);
            break;
        }
        case ins::table_init: {
            const uint32_t n = index("table.init"), src = index("table.init"), dst = index("table.init");
            const std::vector<int64_t> &seg = rs.elems[(size_t)i.imm];
            if ((uint64_t)src + n > seg.size() || (uint64_t)dst + n > tab.size()) throw std::invalid_argument("wat: table_init: index out of bound");
            std::copy(seg.begin() + src, seg.begin() + src + n, tab.begin() + dst);
            break;
        }
        case ins::elem_drop: rs.elems[(size_t)i.imm].clear(); break;
        default: throw std::logic_error("wat: unknown instruction kind");
        }
    }
    // do_load (interpreter_impl.hpp:2206-2228): the address is read as a number; a range that holds a stored witness gives a new witness
    static void load(const ins &i, run_state &rs) {
        const bool plain = !rs.stack.empty() && rs.stack.back().kind == value::NUM;      // a number as the address: the slot is reused for the result
        uint64_t ea;
        if (plain) ea = (uint64_t)(uint32_t)rs.stack.back().num + i.imm;
        else { value tmp = rs.pop(); ea = (uint64_t)(uint32_t)rs.make_numeric(std::move(tmp)) + i.imm; }
        const uint64_t n = i.o;
        if (ea + n > rs.memory.size()) throw std::invalid_argument("wat: invalid memory address");
        uint64_t c = 0;
        memcpy(&c, rs.memory.data() + ea, (size_t)n);
        if (i.sgn && n < 8 && (c >> (8 * n - 1)) & 1) c |= ~0ULL << (8 * n);
        if (bytes_of(i.width) == 4) c &= 0xFFFFFFFFULL;
        if (plain) {
            if (!rs.secrets.iv.empty() && rs.secrets.intersects((uint32_t)ea, (uint32_t)(ea + n))) rs.stack.pop_back();
            else { value &v = rs.stack.back(); v.num = c; v.is64 = i.width == 64 || i.width == F64; v.isf = is_float(i.width); return; }
        }
        if (rs.secrets.intersects((uint32_t)ea, (uint32_t)(ea + n))) rs.push(value::of(rs.make_witness(of_bits(i.width, c))));   // (a float here traps, as in the reference)
        else rs.push(of_bits(i.width, c));
    }
    // do_store (:2310-2344): the value is read as a number and its witnesses are let go; the range is marked iff it was not a number
    static void store(const ins &i, run_state &rs) {
        const size_t depth = rs.stack.size();
        if (depth >= 2 && rs.stack[depth - 1].kind == value::NUM && rs.stack[depth - 2].kind == value::NUM) {      // a number to an address that is a number
            const uint64_t ea = (uint64_t)(uint32_t)rs.stack[depth - 2].num + i.imm, n = i.o;
            if (ea + n > rs.memory.size()) throw std::invalid_argument("wat: invalid memory address");
            if (!rs.secrets.iv.empty()) rs.secrets.subtract((uint32_t)ea, (uint32_t)(ea + n));
            uint64_t c = rs.stack[depth - 1].num;
            if (bytes_of(i.width) == 4) c &= 0xFFFFFFFFULL;
            memcpy(rs.memory.data() + ea, &c, (size_t)n);
            rs.stack.pop_back();
            rs.stack.pop_back();
            return;
        }
        value tmp = rs.pop();
        value addr = rs.pop();
        const uint64_t ea = (uint64_t)(uint32_t)rs.make_numeric(std::move(addr)) + i.imm, n = i.o;
        if (ea + n > rs.memory.size()) throw std::invalid_argument("wat: invalid memory address");
        if (tmp.kind == value::NUM) rs.secrets.subtract((uint32_t)ea, (uint32_t)(ea + n));
        else rs.secrets.add((uint32_t)ea, (uint32_t)(ea + n));
        uint64_t c = rs.make_numeric(std::move(tmp));
        if (bytes_of(i.width) == 4) c &= 0xFFFFFFFFULL;
        memcpy(rs.memory.data() + ea, &c, (size_t)n);
    }
    // memory.size / grow / fill / copy / init, data.drop (:2108-2204): operands must be numbers (the reference reads them with std::get)
    static uint32_t concrete(run_state &rs, const char *what) {
        value v = rs.pop();
        if (v.kind != value::NUM) throw std::invalid_argument(std::string("wat: ") + what + " takes concrete operands");
        return v.as_u32();
    }
    static void bulk_memory(const ins &i, run_state &rs) {
        const uint32_t page = 65536;
        switch (i.kind) {
        case ins::memory_size: rs.push(value::u32((uint32_t)(rs.memory.size() / page))); break;
        case ins::memory_grow: {
            const uint32_t sz = (uint32_t)(rs.memory.size() / page), n = concrete(rs, "memory.grow");
            const uint64_t len = (uint64_t)sz + n;
            if (len > 4096 || (rs.max_pages && len > rs.max_pages)) rs.push(value::u32(0xFFFFFFFFu));   // (the front end caps memory at 256 MiB; the reference at 4 GiB)
            else { rs.memory.resize((size_t)len * page); rs.push(value::u32(sz)); }
            break;
        }
        case ins::memory_fill: {
            const uint32_t n = concrete(rs, "memory.fill"), val = concrete(rs, "memory.fill"), d = concrete(rs, "memory.fill");
            if ((uint64_t)d + n > rs.memory.size()) throw std::invalid_argument("wat: memory.fill: invalid address");
            std::fill_n(rs.memory.begin() + d, n, (uint8_t)val);
            rs.secrets.subtract(d, d + n);
            break;
        }
        case ins::memory_copy: {                              // memcpy_secrets (runtime.hpp:136-172)
            const uint32_t count = concrete(rs, "memory.copy"), src = concrete(rs, "memory.copy"), dst = concrete(rs, "memory.copy");
            if ((uint64_t)src + count > rs.memory.size() || (uint64_t)dst + count > rs.memory.size()) throw std::invalid_argument("wat: memory.copy: out of range");
            secret_ranges moved;
            const uint32_t offset = dst - src;                // wraps, as in the reference: a marked range that starts more than `dst` bytes before `src` wraps to an empty range and loses its mark
            for (const auto &r : rs.secrets.iv) if (r.first < src + count && r.second > src) moved.add(r.first + offset, r.second + offset);
            rs.secrets.subtract(dst, dst + count);
            for (const auto &r : moved.iv) rs.secrets.add(std::max(r.first, dst), std::min(r.second, dst + count));
            memmove(rs.memory.data() + dst, rs.memory.data() + src, count);
            break;
        }
        case ins::memory_init: {
            const uint32_t n = concrete(rs, "memory.init"), sidx = concrete(rs, "memory.init"), d = concrete(rs, "memory.init");
            const std::vector<uint8_t> &data = rs.datas[(size_t)i.imm];
            if ((uint64_t)sidx + n > data.size() || (uint64_t)d + n > rs.memory.size()) throw std::invalid_argument("wat: memory.init: invalid address");
            std::copy(data.begin() + sidx, data.begin() + sidx + n, rs.memory.begin() + d);
            rs.secrets.subtract(d, d + n);
            break;
        }
        case ins::data_drop: rs.datas[(size_t)i.imm].clear(); break;
        default: throw std::logic_error("wat: unknown instruction kind");
        }
    }

    // exec_select (interpreter_impl.hpp:118-140): a concrete condition removes one of the two values from the stack; a witness
    // condition is compared with zero bit by bit and the result is is_zero * second + ~is_zero * first, one new witness
    static void select(run_state &rs) {
        witness_machine &m = rs.m;
        value sc = rs.pop();
        if (sc.kind == value::NUM) {
            rs.drop_n_below(1, sc.as_u32() ? 0 : 1);
            return;
        }
        rs.st.arithmetic_ops++;
        bitvec c = rs.make_decomposed(std::move(sc), 32);
        wref f = rs.make_witness(rs.pop());
        wref t = rs.make_witness(rs.pop());
        wref is_zero = m.bitwise_eqz(c);
        wref v = m.eval(is_zero * f + ~is_zero * t);
        rs.push(value::of(std::move(v)));
    }

    // ---- building the instruction lists.  A light validator rides along: the static width (32 / 64) of every stack slot.
    // wabt would have parsed (not validated) the module for the reference; here an instruction applied to a value of the
    // other width is rejected (the handlers index operand bits by the instruction's width).  Host calls are checked for
    // operand COUNT only: the reference's own tests call assert_equal (param i64 i64) with i32 operands.
    struct import_t {
        std::string module, field;
        bool typed = false;                                   // its signature is known (a function the front end does not provide can then be
        std::vector<uint8_t> params, results;                 // called in the text: the call fails when it is executed, as in the reference)
    };
    func_t *cur_ = nullptr;                                   // the function being built
    std::vector<uint8_t> types_;                              // 32 / 64 / F32 / F64; 0 = any (a value popped in unreachable code)
    // the enclosing blocks of the instruction being added (the function body is the outermost), as WebAssembly validation keeps them
    struct ctrl_t {
        ins::kind_t kind;                                     // block, loop, if_, else_ (the second arm of an if); nop for the function body
        std::vector<uint8_t> start, end;                      // parameter / result widths
        size_t height;                                        // of types_ when the block was entered (below its parameters)
        bool unreachable;
        size_t pc;                                            // where its opening instruction sits
        std::string name;                                     // $label in text
    };
    std::vector<ctrl_t> ctrl_;

    uint8_t pop_type(const std::string &shown) {
        if (types_.size() <= ctrl_.back().height) {
            if (ctrl_.back().unreachable) return 0;
            throw std::invalid_argument("wat: operand stack underflow at " + shown);
        }
        const uint8_t t = types_.back();
        types_.pop_back();
        return t;
    }
    void want(int width, const std::string &shown) {
        const uint8_t t = pop_type(shown);
        if (t && t != width) throw std::invalid_argument("wat: type mismatch: " + shown + " applied to an " + type_name(t) + " value");
    }
    void want_all(const std::vector<uint8_t> &ts, const std::string &shown) { for (size_t i = ts.size(); i-- > 0;) want(ts[i], shown); }
    void dead_code() { types_.resize(ctrl_.back().height); ctrl_.back().unreachable = true; }
    // block / loop / if with its block type; `name` is the $label of the text format
    void emit_block(ins::kind_t k, const std::vector<uint8_t> &params, const std::vector<uint8_t> &results, const std::string &name) {
        const char *shown = k == ins::block ? "block" : (k == ins::loop ? "loop" : "if");
        if (params.size() > 255 || results.size() > 255 || ctrl_.size() > 500) throw std::invalid_argument("wat: block too large or nested too deeply");
        if (k == ins::if_) want(32, shown);
        want_all(params, shown);
        ctrl_.push_back(ctrl_t{k, params, results, types_.size(), false, cur_->code.size(), name});
        types_.insert(types_.end(), params.begin(), params.end());
        ins i; i.kind = k; i.o = (uint8_t)params.size(); i.width = (uint8_t)results.size();
        cur_->code.push_back(i);
    }
    void close_arm(const char *shown) {                       // the values a block leaves must be exactly its results
        ctrl_t &c = ctrl_.back();
        want_all(c.end, shown);
        if (types_.size() != c.height) throw std::invalid_argument(std::string("wat: values left on the stack at ") + shown);
    }
    void emit_else() {
        if (ctrl_.size() < 2 || ctrl_.back().kind != ins::if_) throw std::invalid_argument("wat: else without an if");
        close_arm("else");
        ctrl_t &c = ctrl_.back();
        c.kind = ins::else_; c.unreachable = false;
        types_.insert(types_.end(), c.start.begin(), c.start.end());
        cur_->code[c.pc].aux = (uint32_t)cur_->code.size();
        put(ins::else_);
    }
    void emit_end() {
        if (ctrl_.size() < 2) throw std::invalid_argument("wat: end without a block");
        if (ctrl_.back().kind == ins::if_ && ctrl_.back().start != ctrl_.back().end) throw std::invalid_argument("wat: an if without an else must leave what it takes");
        close_arm("end");
        const ctrl_t c = ctrl_.back();
        ctrl_.pop_back();
        types_.insert(types_.end(), c.end.begin(), c.end.end());
        ins &open = cur_->code[c.pc];
        open.imm = cur_->code.size();
        if (c.kind != ins::else_) open.aux = (uint32_t)cur_->code.size();
        put(ins::end);
    }
    // the values a branch to the block `depth` levels out carries
    const std::vector<uint8_t> &label_types(uint64_t depth, const std::string &shown) {
        if (depth + 1 >= ctrl_.size()) throw std::invalid_argument("wat: " + shown + ": no such enclosing block (a branch out of the function body is not supported; use return)");
        const ctrl_t &c = ctrl_[ctrl_.size() - 1 - (size_t)depth];
        return c.kind == ins::loop ? c.start : c.end;
    }
    void emit_br(ins::kind_t k, uint64_t depth) {
        const std::string shown = k == ins::br ? "br" : "br_if";
        if (k == ins::br_if) want(32, shown);
        const std::vector<uint8_t> ts = label_types(depth, shown);
        want_all(ts, shown);
        if (k == ins::br) dead_code();
        else types_.insert(types_.end(), ts.begin(), ts.end());
        put(k, depth);
    }
    void emit_br_table(const std::vector<uint32_t> &targets) {            // the default last
        if (targets.empty() || targets.size() > 100000) throw std::invalid_argument("wat: malformed br_table");
        want(32, "br_table");
        const std::vector<uint8_t> ts = label_types(targets.back(), "br_table");
        for (uint32_t t : targets) if (label_types(t, "br_table") != ts) throw std::invalid_argument("wat: br_table targets disagree");
        want_all(ts, "br_table");
        dead_code();
        cur_->tables.push_back(targets);
        put(ins::br_table, cur_->tables.size() - 1);
    }
    void emit_return() { want_all(cur_->results, "return"); dead_code(); put(ins::return_); }
    void emit_unreachable() { dead_code(); put(ins::unreachable); }
    // a $label or a depth -> depth
    uint64_t label_depth(const std::string &id) const {
        if (!id.empty() && id[0] == '$') {
            for (size_t d = 0; d + 1 < ctrl_.size(); d++) if (ctrl_[ctrl_.size() - 1 - d].name == id) return d;
            throw std::invalid_argument("wat: unknown label " + id);
        }
        return parse_i64(id);
    }
    void put(ins::kind_t k, uint64_t imm = 0) { ins i; i.kind = k; i.imm = imm; cur_->code.push_back(i); }
    void emit_op(const std::string &name, int width, const std::string &shown) {
        opinfo oi;
        if (!lookup(name, oi)) throw std::invalid_argument("wat: unsupported instruction " + shown);
        if ((oi.o == op::extend32 || oi.o == op::extend_i32) && width != 64) throw std::invalid_argument("wat: unsupported instruction " + shown);
        if (oi.o == op::wrap && width != 32) throw std::invalid_argument("wat: unsupported instruction " + shown);
        const int in_width = oi.o == op::extend_i32 ? 32 : (oi.o == op::wrap ? 64 : width);
        if (oi.arity != 1) want(in_width, shown);
        want(in_width, shown);
        const bool predicate = oi.o == op::eqz || oi.o == op::eq || oi.o == op::ne || oi.o == op::lt || oi.o == op::gt || oi.o == op::le || oi.o == op::ge;
        types_.push_back(predicate ? 32 : (uint8_t)width);
        ins i;
        i.kind = oi.arity == 1 ? ins::unary_op : (oi.arity == 3 ? ins::shift_op : ins::binary_op);
        i.o = (uint8_t)oi.o; i.width = (uint8_t)width; i.sgn = oi.sgn;
        cur_->code.push_back(i);
    }
    void emit_const(int width, uint64_t v) {                  // for F32 / F64: the bit pattern
        types_.push_back((uint8_t)width);
        ins i; i.kind = ins::konst; i.width = (uint8_t)width; i.imm = bytes_of((uint8_t)width) == 4 ? (v & 0xFFFFFFFFULL) : v; cur_->code.push_back(i);
    }
    // a floating-point instruction or a conversion by its full name ("f32.add", "i64.trunc_sat_f64_u"); false if it is none.
    // `probe`: only say how many operands it takes (0 = not a floating-point instruction)
    int emit_float(const std::string &full, const std::string &shown, bool probe = false) {
        if (full.size() < 5 || full[3] != '.') return 0;
        const std::string ty = full.substr(0, 3), name = full.substr(4);
        const uint8_t t = ty == "i32" ? 32 : (ty == "i64" ? 64 : (ty == "f32" ? F32 : (ty == "f64" ? F64 : 0)));
        if (!t) return 0;
        static const std::map<std::string, fop> arith = {
            {"abs", fop::abs}, {"neg", fop::neg}, {"ceil", fop::ceil}, {"floor", fop::floor}, {"trunc", fop::trunc}, {"nearest", fop::nearest}, {"sqrt", fop::sqrt},
            {"add", fop::add}, {"sub", fop::sub}, {"mul", fop::mul}, {"div", fop::div}, {"min", fop::min}, {"max", fop::max}, {"copysign", fop::copysign},
            {"eq", fop::eq}, {"ne", fop::ne}, {"lt", fop::lt}, {"gt", fop::gt}, {"le", fop::le}, {"ge", fop::ge}};
        ins i; i.kind = ins::float_op;
        uint8_t from = 0, to = 0;
        int arity = 1;
        const auto type_at = [&](size_t pos) -> uint8_t {    // "i32" / "i64" / "f32" / "f64" inside the name
            const std::string w = name.substr(pos, 3);
            return w == "i32" ? 32 : (w == "i64" ? 64 : (w == "f32" ? F32 : (w == "f64" ? F64 : 0)));
        };
        if (is_float(t)) {
            const auto it = arith.find(name);
            if (it != arith.end()) {
                i.o = (uint8_t)it->second; i.width = t;
                arity = it->second <= fop::sqrt ? 1 : 2;
                from = t; to = it->second >= fop::eq ? 32 : t;
            } else if (name.compare(0, 8, "convert_") == 0 && name.size() == 13 && (name[12] == 's' || name[12] == 'u') && name[11] == '_' && !is_float(type_at(8)) && type_at(8)) {
                i.o = (uint8_t)fop::convert; i.width = t; i.aux = type_at(8); i.sgn = name[12] == 's';
                from = (uint8_t)i.aux; to = t;
            } else if ((t == F32 && name == "demote_f64") || (t == F64 && name == "promote_f32")) {
                i.o = (uint8_t)(t == F32 ? fop::demote : fop::promote); i.width = t;
                from = t == F32 ? F64 : F32; to = t;
            } else if ((t == F32 && name == "reinterpret_i32") || (t == F64 && name == "reinterpret_i64")) {
                i.o = (uint8_t)fop::reinterpret; i.width = t;
                from = t == F32 ? 32 : 64; to = t;
            } else return 0;
        } else {
            if ((t == 32 && name == "reinterpret_f32") || (t == 64 && name == "reinterpret_f64")) {
                i.o = (uint8_t)fop::reinterpret; i.width = t;
                from = t == 32 ? F32 : F64; to = t;
            } else if (name.compare(0, 6, "trunc_") == 0) {
                const size_t at = name.compare(0, 10, "trunc_sat_") == 0 ? 10 : 6;
                if (name.size() != at + 5 || name[at + 3] != '_' || (name[at + 4] != 's' && name[at + 4] != 'u') || !is_float(type_at(at))) return 0;
                i.o = (uint8_t)(at == 10 ? fop::trunc_sat : fop::trunc_to_int); i.width = t; i.aux = type_at(at); i.sgn = name[at + 4] == 's';
                from = (uint8_t)i.aux; to = t;
            } else return 0;
        }
        if (probe) return arity;
        if (arity == 2) want(from, shown);
        want(from, shown);
        types_.push_back(to);
        cur_->code.push_back(i);
        return arity;
    }
    void emit_global(ins::kind_t k, uint64_t index) {
        if (index >= globals_.size()) throw std::invalid_argument("wat: unknown global " + std::to_string(index));
        const global_t &g = globals_[(size_t)index];
        if (k == ins::global_get) types_.push_back(g.type);
        else {
            if (!g.mut) throw std::invalid_argument("wat: global.set of an immutable global");
            want(g.type, "global.set");
        }
        put(k, index);
    }
    void add_global(uint8_t type, bool mut, uint64_t init) {
        if (type != 32 && type != 64) throw std::invalid_argument("wat: only i32 and i64 globals are supported (the reference traps on others: Unexpected global value type)");
        if (globals_.size() >= 100000) throw std::invalid_argument("wat: too many globals");
        globals_.push_back(global_t{type, mut, type == 32 ? (init & 0xFFFFFFFFULL) : init});
    }
    // fNN.const literals: decimal and hexadecimal floating point, inf, nan, nan:0x<payload> (underscores allowed) -> bits
    static uint64_t parse_float(const std::string &lit, bool single) {
        std::string t;
        for (char ch : lit) if (ch != '_') t.push_back(ch);
        size_t p = 0;
        bool neg = false;
        if (p < t.size() && (t[p] == '+' || t[p] == '-')) { neg = t[p] == '-'; p++; }
        const std::string body = t.substr(p);
        const uint64_t sign = neg ? (single ? 0x80000000ULL : 0x8000000000000000ULL) : 0;
        if (body == "inf") return sign | (single ? 0x7F800000ULL : 0x7FF0000000000000ULL);
        if (body == "nan") return sign | (single ? 0x7FC00000ULL : 0x7FF8000000000000ULL);
        if (body.compare(0, 6, "nan:0x") == 0) {
            const uint64_t payload = parse_i64(body.substr(4));
            const uint64_t mask = single ? 0x7FFFFFULL : 0xFFFFFFFFFFFFFULL;
            if (!payload || payload > mask) throw std::invalid_argument("wat: bad NaN payload " + lit);
            return sign | (single ? 0x7F800000ULL : 0x7FF0000000000000ULL) | payload;
        }
        if (body.empty() || !((body[0] >= '0' && body[0] <= '9'))) throw std::invalid_argument("wat: bad floating-point literal " + lit);
        for (char ch : body) if (!strchr("0123456789abcdefABCDEFxXpP.+-", ch)) throw std::invalid_argument("wat: bad floating-point literal " + lit);
        char *end = nullptr;
        uint64_t bits = 0;
        if (single) { const float v = strtof(body.c_str(), &end); uint32_t b; memcpy(&b, &v, 4); bits = b; }
        else { const double v = strtod(body.c_str(), &end); memcpy(&bits, &v, 8); }
        if (!end || *end) throw std::invalid_argument("wat: bad floating-point literal " + lit);
        return bits | sign;
    }
    // a call of an import the front end does not provide: typed by the import's signature, failing when executed
    bool emit_unprovided(const import_t &im) {
        if (!im.typed) return false;
        want_all(im.params, "call " + im.module + "." + im.field);
        types_.insert(types_.end(), im.results.begin(), im.results.end());
        unprovided_.push_back(printable(im.module + "." + im.field));
        ins i; i.kind = ins::host_call; i.o = (uint8_t)host_fn::unsupported; i.imm = unprovided_.size() - 1; cur_->code.push_back(i);
        return true;
    }
    void emit_host(const import_t &im) {
        const std::string &module = im.module, &field = im.field;
        if (module == "wasi_snapshot_preview1") {
            static const std::map<std::string, std::pair<host_fn, int>> wasi = {       // function, operands (all i32); all but proc_exit return an i32 errno
                {"args_sizes_get", {host_fn::wasi_args_sizes_get, 2}}, {"args_get", {host_fn::wasi_args_get, 2}}, {"fd_write", {host_fn::wasi_fd_write, 4}},
                {"proc_exit", {host_fn::wasi_proc_exit, 1}}, {"random_get", {host_fn::wasi_random_get, 2}}};
            const auto it = wasi.find(field);
            if (it == wasi.end() && emit_unprovided(im)) return;
            if (it == wasi.end()) throw std::invalid_argument("wat: wasi_snapshot_preview1." + printable(field) + " is not supported by the front end (args_sizes_get, args_get, fd_write, proc_exit, random_get are)");
            if (!has_memory_ && it->second.first != host_fn::wasi_proc_exit) throw std::invalid_argument("wat: a WASI call in a module without a memory");
            for (int j = 0; j < it->second.second; j++) want(32, "call wasi_snapshot_preview1." + field);
            if (it->second.first != host_fn::wasi_proc_exit) types_.push_back(32);
            ins i; i.kind = ins::host_call; i.o = (uint8_t)it->second.first; cur_->code.push_back(i);
            return;
        }
        host_fn f = host_fn::unsupported;
        if ((module != "env" || !host_lookup(field, f)) && emit_unprovided(im)) return;
        if (module != "env") throw std::invalid_argument("wat: only the env and wasi_snapshot_preview1 host modules are supported (no bn254fr / vbn254fr / uint256 / ecc imports)");
        if (!host_lookup(field, f)) throw std::invalid_argument("wat: env." + printable(field) + " is not supported by the front end");
        const std::string shown = "call env." + field;
        pop_type(shown);
        if (f == host_fn::assert_equal || f == host_fn::print_str || f == host_fn::dump_memory) pop_type(shown);
        if ((f == host_fn::print_str || f == host_fn::dump_memory) && !has_memory_) throw std::invalid_argument("wat: " + shown + " in a module without a memory");
        if (f == host_fn::i32_private_const || f == host_fn::i64_private_const) types_.push_back(f == host_fn::i32_private_const ? 32 : 64);
        if (f == host_fn::witness_cast) types_.push_back(field == "witness_cast_u32" ? 32 : 64);
        ins i; i.kind = ins::host_call; i.o = (uint8_t)f; cur_->code.push_back(i);
    }
    // a call by function index: imports first, then the module's own functions
    void emit_call(uint64_t index, const std::vector<import_t> &imports) {
        if (index < imports.size()) { emit_host(imports[(size_t)index]); return; }
        const uint64_t fi = index - imports.size();
        if (fi >= funcs_.size()) throw std::invalid_argument("wat: call of an unknown function (" + std::to_string(index) + ")");
        const func_t &f = funcs_[(size_t)fi];
        for (size_t i = f.params.size(); i-- > 0;) want(f.params[i], "call " + std::to_string(index));
        for (uint8_t r : f.results) types_.push_back(r);
        put(ins::func_call, fi);
    }
    void emit_call_indirect(uint64_t type_index, uint64_t table_index) {
        if (!has_table_ || table_index != 0) throw std::invalid_argument("wat: call_indirect through an unknown table");
        if (type_index >= sigs_.size()) throw std::invalid_argument("wat: call_indirect with an unknown type (" + std::to_string(type_index) + ")");
        const sig_t sg = sigs_[(size_t)type_index];
        want(32, "call_indirect");
        want_all(sg.params, "call_indirect");
        types_.insert(types_.end(), sg.results.begin(), sg.results.end());
        put(ins::call_indirect, type_index);
    }
    // the index of a function type with these parameters and results (added if the module has none)
    uint64_t sig_index(const std::vector<uint8_t> &params, const std::vector<uint8_t> &results) {
        for (size_t i = 0; i < sigs_.size(); i++) if (sigs_[i].params == params && sigs_[i].results == results) return i;
        sigs_.push_back(sig_t{params, results});
        return sigs_.size() - 1;
    }
    void set_table(uint64_t min) {
        if (has_table_) throw std::invalid_argument("wat: more than one table");
        if (min > 1000000) throw std::invalid_argument("wat: table too large");
        has_table_ = true;
        table_.assign((size_t)min, -1);
    }
    // an active element segment: module functions (by index among imports + functions) written into the table at `offset`
    void set_elements(uint64_t offset, const std::vector<int64_t> &funcs, size_t nimports) {
        if (!has_table_) throw std::invalid_argument("wat: an element segment needs a table");
        if (offset + funcs.size() > table_.size()) throw std::invalid_argument("wat: table_init: index out of bound");
        nimports_ = nimports;
        for (size_t i = 0; i < funcs.size(); i++) {
            if (funcs[i] >= 0 && (uint64_t)funcs[i] >= nimports + funcs_.size()) throw std::invalid_argument("wat: an element segment names an unknown function");
            table_[(size_t)offset + i] = funcs[i];
        }
        elem_segs_.push_back(funcs);                          // active segments are NOT dropped by the reference's instantiate() (runtime.hpp:518-536): table.init may read them again
    }
    // ref.null / ref.is_null / ref.func and the table instructions; `a`, `b`: table / segment / function indices as the instruction has them
    void emit_reference(ins::kind_t k, uint64_t a = 0, uint64_t b = 0, uint8_t type = FUNCREF) {
        const auto table_known = [&](uint64_t t, const char *shown) { if (!has_table_ || t != 0) throw std::invalid_argument(std::string("wat: ") + shown + " on an unknown table"); };
        switch (k) {
        case ins::ref_null: if (!is_ref(type)) throw std::invalid_argument("wat: ref.null of an unknown type"); types_.push_back(type); break;
        case ins::ref_is_null: { const uint8_t t = pop_type("ref.is_null"); if (t && !is_ref(t)) throw std::invalid_argument("wat: type mismatch: ref.is_null applied to an " + type_name(t) + " value"); types_.push_back(32); break; }
        case ins::ref_func: if (a >= nimports_ + funcs_.size()) throw std::invalid_argument("wat: ref.func of an unknown function"); types_.push_back(FUNCREF); break;
        case ins::table_get: table_known(a, "table.get"); want(32, "table.get"); types_.push_back(FUNCREF); break;
        case ins::table_set: table_known(a, "table.set"); want(FUNCREF, "table.set"); want(32, "table.set"); break;
        case ins::table_size: table_known(a, "table.size"); types_.push_back(32); break;
        case ins::table_grow: table_known(a, "table.grow"); want(32, "table.grow"); want(FUNCREF, "table.grow"); types_.push_back(32); break;
        case ins::table_fill: table_known(a, "table.fill"); want(32, "table.fill"); want(FUNCREF, "table.fill"); want(32, "table.fill"); break;
        case ins::table_copy: table_known(a, "table.copy"); table_known(b, "table.copy"); for (int j = 0; j < 3; j++) want(32, "table.copy"); break;
        case ins::table_init:                                 // a = segment, b = table
            table_known(b, "table.init");
            if (a >= elem_segs_.size()) throw std::invalid_argument("wat: table.init of an unknown element segment");
            for (int j = 0; j < 3; j++) want(32, "table.init");
            break;
        case ins::elem_drop:
            // exec_elem_drop (interpreter_impl.hpp:2095-2104) looks its operand up in the module's TABLE list and clears the ELEMENT
            // segment at that address: right for segment 0 of a module with one table, out of bounds otherwise
            if (a != 0 || !has_table_ || elem_segs_.empty()) throw std::invalid_argument("wat: elem.drop of a segment other than 0 is not supported (the reference's elem.drop reads outside its table list there)");
            break;
        default: throw std::logic_error("wat: not a reference instruction");
        }
        put(k, a);
    }
    void emit_local(ins::kind_t k, uint64_t index) {
        if (index >= cur_->locals.size()) throw std::invalid_argument("wat: unknown local " + std::to_string(index));
        const uint8_t t = cur_->locals[(size_t)index];
        if (k == ins::local_get) types_.push_back(t);
        else { want(t, k == ins::local_set ? "local.set" : "local.tee"); if (k == ins::local_tee) types_.push_back(t); }
        put(k, index);
    }
    // iNN.load* / iNN.store* (name without the "iNN." prefix); false if it is not a memory access
    bool emit_access(const std::string &name, int width, uint64_t offset, const std::string &shown) {
        static const std::map<std::string, std::pair<int, int>> table = {      // bytes (0 = the full width), sign: 1 signed, 0 unsigned
            {"load", {0, 0}}, {"load8_s", {1, 1}}, {"load8_u", {1, 0}}, {"load16_s", {2, 1}}, {"load16_u", {2, 0}}, {"load32_s", {4, 1}}, {"load32_u", {4, 0}},
            {"store", {0, 0}}, {"store8", {1, 0}}, {"store16", {2, 0}}, {"store32", {4, 0}}};
        const auto it = table.find(name);
        if (it == table.end()) return false;
        const int bytes = it->second.first ? it->second.first : (int)bytes_of((uint8_t)width);
        if (bytes == 4 && it->second.first && width != 64) return false;      // load32_* / store32 exist for i64 only
        if (is_float((uint8_t)width) && it->second.first) return false;       // floats move whole
        if (!has_memory_) throw std::invalid_argument("wat: " + shown + " in a module without a memory");
        if (offset > 0xFFFFFFFFULL) throw std::invalid_argument("wat: memory offset out of range");
        const bool is_store = name[0] == 's';
        if (is_store) want(width, shown);
        want(32, shown);
        if (!is_store) types_.push_back((uint8_t)width);
        ins i;
        i.kind = is_store ? ins::store : ins::load;
        i.o = (uint8_t)bytes; i.width = (uint8_t)width; i.sgn = it->second.second != 0; i.imm = offset;
        cur_->code.push_back(i);
        return true;
    }
    void emit_bulk(ins::kind_t k, uint64_t index = 0) {
        if (!has_memory_) throw std::invalid_argument("wat: memory instruction in a module without a memory");
        const char *shown = k == ins::memory_size ? "memory.size" : k == ins::memory_grow ? "memory.grow" : k == ins::memory_fill ? "memory.fill"
                          : k == ins::memory_copy ? "memory.copy" : k == ins::memory_init ? "memory.init" : "data.drop";
        if ((k == ins::memory_init || k == ins::data_drop) && index >= datas_.size()) throw std::invalid_argument(std::string("wat: ") + shown + " of an unknown data segment");
        if (k == ins::memory_grow) want(32, shown);
        if (k == ins::memory_fill || k == ins::memory_copy || k == ins::memory_init) for (int j = 0; j < 3; j++) want(32, shown);
        if (k == ins::memory_size || k == ins::memory_grow) types_.push_back(32);
        put(k, index);
    }
    void emit_plain(ins::kind_t k) {
        if (k == ins::drop) pop_type("drop");
        if (k == ins::select) {
            want(32, "select");
            const uint8_t b = pop_type("select"), a = pop_type("select");
            if (a && b && a != b) throw std::invalid_argument("wat: type mismatch: select applied to an " + type_name(a) + " and an " + type_name(b) + " value");
            types_.push_back(a ? a : b);
        }
        put(k);
    }
    void begin_body(func_t &f) {
        cur_ = &f; types_.clear(); ctrl_.clear();
        ctrl_.push_back(ctrl_t{ins::nop, {}, f.results, 0, false, 0, ""});
    }
    void end_body(const std::string &name) {
        if (ctrl_.size() != 1) throw std::invalid_argument("wat: " + name + " ends inside a block");
        if (!ctrl_.back().unreachable && types_.size() != cur_->results.size())
            throw std::invalid_argument("wat: " + name + " leaves " + std::to_string(types_.size()) + " values, its type says " + std::to_string(cur_->results.size()));
        for (size_t i = cur_->results.size(); i-- > 0;) {
            const uint8_t t = pop_type(name);
            if (t && t != cur_->results[i]) throw std::invalid_argument("wat: " + name + " returns a value of the wrong type");
        }
        if (!types_.empty()) throw std::invalid_argument("wat: " + name + " leaves values behind its results");
        cur_ = nullptr;
    }
    static std::string printable(const std::string &s) {      // names from a binary go into error messages
        std::string o;
        for (size_t i = 0; i < s.size() && i < 64; i++) o.push_back((s[i] >= 0x20 && s[i] < 0x7f) ? s[i] : '?');
        return o;
    }
    static uint8_t width_of(const std::string &t) {
        if (t == "i32") return 32;
        if (t == "i64") return 64;
        if (t == "f32") return F32;
        if (t == "f64") return F64;
        if (t == "funcref") return FUNCREF;
        if (t == "externref") return EXTERNREF;
        throw std::invalid_argument("wat: unknown value type (" + printable(t) + ")");
    }

    // ---- text ------------------------------------------------------------------------------------------------
    struct text_scope {
        const std::vector<import_t> &imports;
        const std::map<std::string, size_t> &func_ids;        // $name -> function index (imports first)
        const std::map<std::string, size_t> &data_ids;
        const std::map<std::string, size_t> &global_ids;
        const std::map<std::string, size_t> &type_ids;
        const std::map<std::string, size_t> &elem_ids;
        std::map<std::string, size_t> local_ids;
    };
    void set_memory(uint64_t pages, uint64_t max_pages) {
        if (pages > 4096) throw std::invalid_argument("wat: memories beyond 256 MiB are not supported");
        has_memory_ = true; mem_pages_ = (uint32_t)pages; mem_max_ = (uint32_t)std::min<uint64_t>(max_pages, 65536);
    }
    static std::string decode_string(const std::string &quoted) {           // WebAssembly text string literal -> bytes
        std::string out;
        const auto hex = [](char ch) { return (ch >= '0' && ch <= '9') ? ch - '0' : ((ch >= 'a' && ch <= 'f') ? ch - 'a' + 10 : ((ch >= 'A' && ch <= 'F') ? ch - 'A' + 10 : -1)); };
        for (size_t i = 1; i + 1 < quoted.size(); i++) {
            if (quoted[i] != '\\') { out.push_back(quoted[i]); continue; }
            if (++i >= quoted.size() - 1) throw std::invalid_argument("wat: bad escape in a string");
            const char ch = quoted[i];
            if (ch == 'n') out.push_back('\n');
            else if (ch == 't') out.push_back('\t');
            else if (ch == 'r') out.push_back('\r');
            else if (ch == '"' || ch == '\'' || ch == '\\') out.push_back(ch);
            else if (hex(ch) >= 0 && i + 1 < quoted.size() - 1 && hex(quoted[i + 1]) >= 0) { out.push_back((char)(hex(ch) * 16 + hex(quoted[i + 1]))); i++; }
            else throw std::invalid_argument("wat: unsupported escape in a string");
        }
        return out;
    }
    static uint64_t data_index(const std::string &id, const text_scope &sc) {
        const auto it = sc.data_ids.find(id);
        if (it != sc.data_ids.end()) return it->second;
        if (!id.empty() && id[0] >= '0' && id[0] <= '9') return parse_i64(id);
        throw std::invalid_argument("wat: unknown data segment " + id);
    }
    // offset=N / align=N after a memory access; returns the offset and steps `j` past them
    static uint64_t memarg(const std::vector<sexpr> &list, size_t &j) {
        uint64_t offset = 0;
        for (; j < list.size() && !list[j].is_list; j++) {
            const std::string &a = list[j].atom;
            if (a.compare(0, 7, "offset=") == 0) offset = parse_i64(a.substr(7));
            else if (a.compare(0, 6, "align=") == 0) (void)parse_i64(a.substr(6));
            else break;
        }
        return offset;
    }
    void parse_text(const std::string &text) {
        sexpr_parser p(text);
        const sexpr top = p.parse_top();
        if (top.head() != "module") throw std::invalid_argument("wat: expected (module ...)");
        std::vector<import_t> imports;
        std::map<std::string, size_t> func_ids, data_ids, global_ids, type_ids, elem_ids;
        std::vector<const sexpr *> bodies, elems, inline_elems, import_descs;
        std::string start;
        for (size_t i = 1; i < top.list.size(); i++) {
            const sexpr &f = top.list[i];
            if (f.head() == "import") {
                // (import "env" "name" (func $id ...))
                if (f.list.size() < 4 || f.list[3].head() != "func") throw std::invalid_argument("wat: unsupported import");
                if (!bodies.empty()) throw std::invalid_argument("wat: imports must come before the module's functions");
                if (f.list[3].list.size() >= 2 && !f.list[3].list[1].is_list) func_ids[f.list[3].list[1].atom] = imports.size();
                imports.push_back(import_t{unquote(f.list[1].atom), unquote(f.list[2].atom), false, {}, {}});
                import_descs.push_back(&f.list[3]);
            } else if (f.head() == "func") {
                if (f.list.size() >= 2 && !f.list[1].is_list) func_ids[f.list[1].atom] = imports.size() + bodies.size();
                bodies.push_back(&f);
            } else if (f.head() == "export") {
                if (f.list.size() >= 3 && unquote(f.list[1].atom) == "_start" && f.list[2].head() == "func" && f.list[2].list.size() == 2) start = f.list[2].list[1].atom;
            } else if (f.head() == "start") {
                // a start function is NOT run by the reference's instantiate() (include/runtime.hpp:345-604 never reads module.starts): ignored here too
            } else if (f.head() == "memory") {                 // (memory [$id] min [max])
                size_t j = (f.list.size() >= 2 && !f.list[1].is_list && f.list[1].atom[0] == '$') ? 2 : 1;
                while (j < f.list.size() && f.list[j].head() == "export") j++;                    // (memory (export "memory") 1): exports other than _start say nothing here
                if (has_memory_ || j >= f.list.size() || f.list[j].is_list) throw std::invalid_argument("wat: unsupported memory declaration");
                set_memory(parse_i64(f.list[j].atom), j + 1 < f.list.size() && !f.list[j + 1].is_list ? parse_i64(f.list[j + 1].atom) : 0);
            } else if (f.head() == "data") {                   // (data [$id] "bytes"...) passive; (data [$id] (i32.const off) "bytes"...) active
                data_t d;
                size_t j = 1;
                if (j < f.list.size() && !f.list[j].is_list && f.list[j].atom[0] == '$') data_ids[f.list[j++].atom] = datas_.size();
                if (j < f.list.size() && f.list[j].is_list) {
                    const sexpr *off = &f.list[j++];
                    if (off->head() == "offset" && off->list.size() == 2) off = &off->list[1];
                    if (off->head() != "i32.const" || off->list.size() != 2) throw std::invalid_argument("wat: a data segment's offset must be an i32.const");
                    d.active = true;
                    d.offset = (uint32_t)parse_i64(off->list[1].atom);
                }
                for (; j < f.list.size(); j++) {
                    if (f.list[j].is_list || f.list[j].atom.empty() || f.list[j].atom[0] != '"') throw std::invalid_argument("wat: malformed data segment");
                    d.bytes += decode_string(f.list[j].atom);
                }
                datas_.push_back(d);
            } else if (f.head() == "type") {                   // (type [$id] (func (param ..)* (result ..)*))
                size_t j = 1;
                if (j < f.list.size() && !f.list[j].is_list && f.list[j].atom[0] == '$') type_ids[f.list[j++].atom] = sigs_.size();
                if (j + 1 != f.list.size() || f.list[j].head() != "func") throw std::invalid_argument("wat: unsupported type definition");
                sig_t sg;
                for (size_t q = 1; q < f.list[j].list.size(); q++) {
                    const sexpr &part = f.list[j].list[q];
                    if (part.head() != "param" && part.head() != "result") throw std::invalid_argument("wat: unsupported type definition");
                    for (size_t t = (part.list.size() == 3 && !part.list[1].atom.empty() && part.list[1].atom[0] == '$') ? 2 : 1; t < part.list.size(); t++)
                        (part.head() == "param" ? sg.params : sg.results).push_back(width_of(part.list[t].atom));
                }
                sigs_.push_back(sg);
            } else if (f.head() == "table") {                  // (table [$id] min [max] funcref) | (table [$id] funcref (elem $f ...))
                size_t j = (f.list.size() >= 2 && !f.list[1].is_list && f.list[1].atom[0] == '$') ? 2 : 1;
                if (j < f.list.size() && !f.list[j].is_list && f.list[j].atom == "funcref" && j + 2 == f.list.size() && f.list[j + 1].head() == "elem") {
                    set_table(f.list[j + 1].list.size() - 1);
                    inline_elems.push_back(&f.list[j + 1]);
                } else {
                    if (j >= f.list.size() || f.list[j].is_list || f.list.back().is_list || f.list.back().atom != "funcref") throw std::invalid_argument("wat: unsupported table declaration");
                    set_table(parse_i64(f.list[j].atom));
                }
            } else if (f.head() == "elem") {
                elems.push_back(&f);
            } else if (f.head() == "global") {                 // (global [$id] i32 | (mut i32) (i32.const v)): the initialiser must be a constant
                size_t j = 1;
                if (j < f.list.size() && !f.list[j].is_list && f.list[j].atom[0] == '$') global_ids[f.list[j++].atom] = globals_.size();
                if (j + 2 != f.list.size() || !f.list[j + 1].is_list) throw std::invalid_argument("wat: unsupported global declaration");
                const sexpr &ty = f.list[j], &init = f.list[j + 1];
                const bool mut = ty.is_list;
                if (mut && (ty.head() != "mut" || ty.list.size() != 2 || ty.list[1].is_list)) throw std::invalid_argument("wat: unsupported global declaration");
                const uint8_t t = width_of(mut ? ty.list[1].atom : ty.atom);
                if (is_float(t)) add_global(t, mut, 0);         // (throws: the reference takes i32 / i64 globals only)
                if (init.list.size() != 2 || init.head() != (t == 32 ? "i32.const" : "i64.const") || init.list[1].is_list) throw std::invalid_argument("wat: a global's initialiser must be a constant of its type");
                add_global(t, mut, literal(init.head(), init.list[1].atom));
            } else {
                throw std::invalid_argument("wat: unsupported module field (" + f.head() + ")");
            }
        }
        for (const data_t &d : datas_) if (d.active && !has_memory_) throw std::invalid_argument("wat: an active data segment needs a memory");
        // signatures first (a body may call a later function), then the bodies
        funcs_.resize(bodies.size());
        std::vector<std::map<std::string, size_t>> local_ids(bodies.size());
        std::vector<size_t> first_instr(bodies.size());
        for (size_t k = 0; k < bodies.size(); k++) {
            const sexpr &f = *bodies[k];
            func_t &fn = funcs_[k];
            size_t i = (f.list.size() >= 2 && !f.list[1].is_list) ? 2 : 1;
            int64_t typed = -1;
            for (; i < f.list.size() && f.list[i].is_list; i++) {
                const sexpr &e = f.list[i];
                const std::string &h = e.head();
                if (h == "export") {                           // (func $f (export "name") ...): the inline form of an export
                    if (e.list.size() != 2 || e.list[1].is_list) throw std::invalid_argument("wat: malformed function header");
                    if (unquote(e.list[1].atom) == "_start") start = std::to_string(imports.size() + k);
                    continue;
                }
                if (h == "type") {                             // (func $f (type $t) ...): the signature is the type's; a (param ..) / (result ..) list after it repeats it
                    if (e.list.size() != 2 || e.list[1].is_list || typed >= 0 || !fn.locals.empty() || !fn.results.empty()) throw std::invalid_argument("wat: malformed function header");
                    const std::string &id = e.list[1].atom;
                    const auto it = type_ids.find(id);
                    const uint64_t t = it != type_ids.end() ? it->second : ((!id.empty() && id[0] >= '0' && id[0] <= '9') ? parse_i64(id) : ~0ULL);
                    if (t >= sigs_.size()) throw std::invalid_argument("wat: unknown type " + id);
                    typed = (int64_t)t;
                    continue;
                }
                if (h != "param" && h != "result" && h != "local") break;
                if (h != "param" && fn.locals.size() < fn.params.size()) throw std::invalid_argument("wat: malformed function header");
                size_t j = 1;
                if (h != "result" && e.list.size() == 3 && !e.list[1].atom.empty() && e.list[1].atom[0] == '$') { local_ids[k][e.list[1].atom] = fn.locals.size(); j = 2; }
                for (; j < e.list.size(); j++) {
                    const uint8_t w = width_of(e.list[j].atom);
                    if (h == "result") { fn.results.push_back(w); continue; }
                    if (h == "local" && is_ref(w)) throw std::invalid_argument("wat: locals of a reference type are not supported (the reference stops: Unsupported local type)");
                    if (h == "param") { if (fn.locals.size() != fn.params.size()) throw std::invalid_argument("wat: parameters must come before locals"); fn.params.push_back(w); }
                    fn.locals.push_back(w);
                }
            }
            first_instr[k] = i;
            if (typed >= 0) {
                const sig_t &sg = sigs_[(size_t)typed];
                if (fn.params.empty() && fn.results.empty()) {   // the parameters come before the declared locals: named locals move up
                    fn.params = sg.params; fn.results = sg.results;
                    fn.locals.insert(fn.locals.begin(), sg.params.begin(), sg.params.end());
                    for (auto &named : local_ids[k]) named.second += sg.params.size();
                } else if (fn.params != sg.params || fn.results != sg.results) throw std::invalid_argument("wat: a function header disagrees with its type");
            }
        }
        // element segments: (elem [$id] [(table ..)] (offset? (i32.const n)) [func] $f ...) writes functions into the table at instantiation
        nimports_ = imports.size();
        const text_scope names{imports, func_ids, data_ids, global_ids, type_ids, elem_ids, {}};
        for (size_t k = 0; k < imports.size(); k++) {          // (func $id? (type t) | (param ..)* (result ..)*): the import's signature, if it is spelled out
            const sexpr &d = *import_descs[k];
            size_t j = (d.list.size() >= 2 && !d.list[1].is_list) ? 2 : 1;
            try {
                const uint64_t t = type_use(d.list, j, names);
                if (j == d.list.size()) { imports[k].typed = true; imports[k].params = sigs_[(size_t)t].params; imports[k].results = sigs_[(size_t)t].results; }
            } catch (const std::invalid_argument &) {}       // a signature over types the front end does not know: the import stays untyped
        }
        const auto functions_from = [&](const sexpr &e, size_t j) {
            std::vector<int64_t> fs;
            for (; j < e.list.size(); j++) {
                if (e.list[j].is_list) throw std::invalid_argument("wat: element expressions are not supported (name the functions)");
                fs.push_back((int64_t)func_index(e.list[j].atom, names));
            }
            return fs;
        };
        for (const sexpr *e : inline_elems) set_elements(0, functions_from(*e, 1), imports.size());
        for (const sexpr *e : elems) {
            size_t j = 1;
            if (j + 1 < e->list.size() && !e->list[j].is_list && e->list[j].atom[0] == '$' && e->list[j + 1].is_list) elem_ids[e->list[j++].atom] = elem_segs_.size();   // the segment's own name
            if (j < e->list.size() && e->list[j].head() == "table") j++;
            if (j < e->list.size() && !e->list[j].is_list && e->list[j].atom == "declare") continue;   // declarative: only announces ref.func targets
            if (j >= e->list.size() || !e->list[j].is_list) throw std::invalid_argument("wat: passive element segments are not supported (no table.init)");
            const sexpr *off = &e->list[j++];
            if (off->head() == "offset" && off->list.size() == 2) off = &off->list[1];
            if (off->head() != "i32.const" || off->list.size() != 2 || off->list[1].is_list) throw std::invalid_argument("wat: an element segment's offset must be an i32.const");
            if (j < e->list.size() && !e->list[j].is_list && e->list[j].atom == "func") j++;
            set_elements((uint32_t)parse_i64(off->list[1].atom), functions_from(*e, j), imports.size());
        }
        if (start.empty()) throw std::invalid_argument("wat: no exported _start function");
        const auto sit = func_ids.find(start);
        size_t start_index;
        if (sit != func_ids.end()) start_index = sit->second;
        else if (start[0] >= '0' && start[0] <= '9') start_index = (size_t)parse_i64(start);
        else throw std::invalid_argument("wat: no exported _start function");
        if (start_index < imports.size() || start_index - imports.size() >= funcs_.size()) throw std::invalid_argument("wat: no exported _start function");
        start_ = start_index - imports.size();
        if (!funcs_[start_].params.empty() || !funcs_[start_].results.empty()) throw std::invalid_argument("wat: _start with parameters / results is not supported");
        for (size_t k = 0; k < bodies.size(); k++) {
            const sexpr &f = *bodies[k];
            func_t &fn = funcs_[k];
            text_scope sc{imports, func_ids, data_ids, global_ids, type_ids, elem_ids, local_ids[k]};
            begin_body(fn);
            parse_seq(f.list, first_instr[k], f.list.size(), sc);
            end_body(f.list.size() >= 2 && !f.list[1].is_list ? f.list[1].atom : "function " + std::to_string(k));
        }
    }
    // (param ..) / (result ..) lists of a block type, from list[j] on
    static void blocktype(const std::vector<sexpr> &list, size_t &j, std::vector<uint8_t> &params, std::vector<uint8_t> &results) {
        for (; j < list.size() && list[j].is_list && (list[j].head() == "param" || list[j].head() == "result" || list[j].head() == "type"); j++) {
            if (list[j].head() == "type") throw std::invalid_argument("wat: block types by index are not supported in text");
            const size_t first = (list[j].list.size() == 3 && !list[j].list[1].is_list && !list[j].list[1].atom.empty() && list[j].list[1].atom[0] == '$') ? 2 : 1;   // (param $name t)
            for (size_t t = first; t < list[j].list.size(); t++) (list[j].head() == "param" ? params : results).push_back(width_of(list[j].list[t].atom));
        }
    }
    // the type use after call_indirect: (type $t) and / or (param ..)* (result ..)*, from list[j] on -> type index
    uint64_t type_use(const std::vector<sexpr> &list, size_t &j, const text_scope &sc) {
        int64_t named = -1;
        if (j < list.size() && list[j].is_list && list[j].head() == "type" && list[j].list.size() == 2 && !list[j].list[1].is_list) {
            const std::string &id = list[j++].list[1].atom;
            const auto it = sc.type_ids.find(id);
            if (it != sc.type_ids.end()) named = (int64_t)it->second;
            else if (!id.empty() && id[0] >= '0' && id[0] <= '9') named = (int64_t)parse_i64(id);
            else throw std::invalid_argument("wat: unknown type " + id);
            if ((uint64_t)named >= sigs_.size()) throw std::invalid_argument("wat: unknown type " + id);
        }
        std::vector<uint8_t> params, results;
        const size_t before = j;
        blocktype(list, j, params, results);
        if (named >= 0) {
            if (j != before && (sigs_[(size_t)named].params != params || sigs_[(size_t)named].results != results)) throw std::invalid_argument("wat: a type use disagrees with its type");
            return (uint64_t)named;
        }
        return sig_index(params, results);
    }
    // a sequence of instructions, folded forms and plain ones mixed: list[from, to)
    void parse_seq(const std::vector<sexpr> &list, size_t from, size_t to, const text_scope &sc) {
        for (size_t i = from; i < to; i++) {
            const sexpr &e = list[i];
            if (e.is_list) {
                flatten(e, sc);
                continue;
            }
            // plain (unfolded) instructions: immediates follow their instruction
            const std::string &a = e.atom;
            const auto next = [&]() -> const std::string & {
                if (i + 1 >= to || list[i + 1].is_list) throw std::invalid_argument("wat: " + a + " needs an immediate");
                return list[++i].atom;
            };
            const auto label_follows = [&]() { return i + 1 < to && !list[i + 1].is_list && (list[i + 1].atom[0] == '$' || (list[i + 1].atom[0] >= '0' && list[i + 1].atom[0] <= '9')); };
            if (a == "i32.const" || a == "i64.const") emit_const(a[1] == '3' ? 32 : 64, literal(a, next()));
            else if (a == "f32.const" || a == "f64.const") emit_const(a[1] == '3' ? F32 : F64, parse_float(next(), a[1] == '3'));
            else if (a == "global.get" || a == "global.set") emit_global(a == "global.get" ? ins::global_get : ins::global_set, global_index(next(), sc));
            else if (a == "call") emit_call(func_index(next(), sc), sc.imports);
            else if (a == "call_indirect") {
                if (i + 1 < to && !list[i + 1].is_list) { if (list[i + 1].atom != "0" && list[i + 1].atom[0] != '$') throw std::invalid_argument("wat: call_indirect through an unknown table"); i++; }
                size_t j = i + 1;
                const uint64_t t = type_use(list, j, sc);
                i = j - 1;
                emit_call_indirect(t, 0);
            }
            else if (a == "local.get" || a == "local.set" || a == "local.tee") emit_local(a == "local.get" ? ins::local_get : (a == "local.set" ? ins::local_set : ins::local_tee), local_index(next(), sc));
            else if (a == "select") {
                while (i + 1 < to && list[i + 1].is_list && list[i + 1].head() == "result") i++;      // select (result t): the typed spelling of the same instruction
                emit_plain(ins::select);
            }
            else if (a == "ref.null") { const std::string &t = next(); emit_reference(ins::ref_null, 0, 0, (t == "func" || t == "funcref") ? FUNCREF : ((t == "extern" || t == "externref") ? EXTERNREF : 0)); }
            else if (a == "ref.is_null") emit_reference(ins::ref_is_null);
            else if (a == "ref.func") emit_reference(ins::ref_func, func_index(next(), sc));
            else if (a == "table.get" || a == "table.set" || a == "table.size" || a == "table.grow" || a == "table.fill" || a == "table.copy" || a == "table.init" || a == "elem.drop") {
                std::vector<std::string> imm;                 // table names / indices (and the segment for table.init / elem.drop)
                while (i + 1 < to && !list[i + 1].is_list && (list[i + 1].atom[0] == '$' || (list[i + 1].atom[0] >= '0' && list[i + 1].atom[0] <= '9')) && imm.size() < 2) imm.push_back(list[++i].atom);
                emit_table_text(a, imm, sc);
            }
            else if (a == "drop") emit_plain(ins::drop);
            else if (a == "nop") emit_plain(ins::nop);
            else if (a == "block" || a == "loop" || a == "if") {
                std::string name;
                if (i + 1 < to && !list[i + 1].is_list && list[i + 1].atom[0] == '$') name = list[++i].atom;
                std::vector<uint8_t> params, results;
                size_t j = i + 1;
                blocktype(list, j, params, results);
                i = j - 1;
                emit_block(a == "block" ? ins::block : (a == "loop" ? ins::loop : ins::if_), params, results, name);
            }
            else if (a == "else") { if (i + 1 < to && !list[i + 1].is_list && list[i + 1].atom[0] == '$') i++; emit_else(); }
            else if (a == "end") { if (i + 1 < to && !list[i + 1].is_list && list[i + 1].atom[0] == '$') i++; emit_end(); }
            else if (a == "br" || a == "br_if") emit_br(a == "br" ? ins::br : ins::br_if, label_depth(next()));
            else if (a == "br_table") {
                std::vector<uint32_t> targets;
                while (label_follows()) targets.push_back((uint32_t)label_depth(list[++i].atom));
                emit_br_table(targets);
            }
            else if (a == "return") emit_return();
            else if (a == "unreachable") emit_unreachable();
            else if (a == "memory.size") emit_bulk(ins::memory_size);
            else if (a == "memory.grow") emit_bulk(ins::memory_grow);
            else if (a == "memory.fill") emit_bulk(ins::memory_fill);
            else if (a == "memory.copy") emit_bulk(ins::memory_copy);
            else if (a == "memory.init") emit_bulk(ins::memory_init, data_index(next(), sc));
            else if (a == "data.drop") emit_bulk(ins::data_drop, data_index(next(), sc));
            else if (a.size() > 4 && (a.compare(0, 4, "i32.") == 0 || a.compare(0, 4, "i64.") == 0 || a.compare(0, 4, "f32.") == 0 || a.compare(0, 4, "f64.") == 0)) {
                size_t j = i + 1;
                const uint64_t offset = memarg(list, j);
                if (emit_access(a.substr(4), width_of(a.substr(0, 3)), offset, a)) i = j - 1;
                else if (emit_float(a, a)) {}
                else if (a[0] == 'i') emit_op(a.substr(4), a[1] == '3' ? 32 : 64, a);
                else throw std::invalid_argument("wat: unsupported instruction " + a);
            }
            else throw std::invalid_argument("wat: unsupported instruction " + a);
        }
    }
    static uint64_t literal(const std::string &instr, const std::string &lit) {
        const uint64_t v = parse_i64(lit);
        if (instr[1] == '3' && v > 0xFFFFFFFFULL && v < 0xFFFFFFFF80000000ULL) throw std::invalid_argument("wat: integer literal out of range " + lit);
        return v;
    }
    static uint64_t func_index(const std::string &id, const text_scope &sc) {
        const auto it = sc.func_ids.find(id);
        if (it != sc.func_ids.end()) return it->second;
        if (!id.empty() && id[0] >= '0' && id[0] <= '9') return parse_i64(id);
        throw std::invalid_argument("wat: call of an unknown function (" + id + ")");
    }
    // table.* / elem.drop with their textual immediates: tables are named $t or 0 (there is one), segments by $name or index
    void emit_table_text(const std::string &name, const std::vector<std::string> &imm, const text_scope &sc) {
        const auto segment = [&](const std::string &id) -> uint64_t {
            const auto it = sc.elem_ids.find(id);
            if (it != sc.elem_ids.end()) return it->second;
            if (!id.empty() && id[0] >= '0' && id[0] <= '9') return parse_i64(id);
            throw std::invalid_argument("wat: unknown element segment " + id);
        };
        const auto table = [&](const std::string &id) -> uint64_t { return (!id.empty() && id[0] >= '0' && id[0] <= '9') ? parse_i64(id) : 0; };
        if (name == "elem.drop") { if (imm.size() != 1) throw std::invalid_argument("wat: malformed elem.drop"); emit_reference(ins::elem_drop, segment(imm[0])); return; }
        if (name == "table.init") {                           // table.init [$t] $e
            if (imm.empty()) throw std::invalid_argument("wat: malformed table.init");
            emit_reference(ins::table_init, segment(imm.back()), imm.size() == 2 ? table(imm[0]) : 0);
            return;
        }
        if (name == "table.copy") { emit_reference(ins::table_copy, imm.size() == 2 ? table(imm[0]) : 0, imm.size() == 2 ? table(imm[1]) : 0); return; }
        if (imm.size() > 1) throw std::invalid_argument("wat: malformed " + name);
        const ins::kind_t k = name == "table.get" ? ins::table_get : (name == "table.set" ? ins::table_set : (name == "table.size" ? ins::table_size : (name == "table.grow" ? ins::table_grow : ins::table_fill)));
        emit_reference(k, imm.empty() ? 0 : table(imm[0]));
    }
    static uint64_t global_index(const std::string &id, const text_scope &sc) {
        const auto it = sc.global_ids.find(id);
        if (it != sc.global_ids.end()) return it->second;
        if (!id.empty() && id[0] >= '0' && id[0] <= '9') return parse_i64(id);
        throw std::invalid_argument("wat: unknown global " + id);
    }
    static uint64_t local_index(const std::string &id, const text_scope &sc) {
        const auto it = sc.local_ids.find(id);
        if (it != sc.local_ids.end()) return it->second;
        if (!id.empty() && id[0] >= '0' && id[0] <= '9') return parse_i64(id);
        throw std::invalid_argument("wat: unknown local " + id);
    }
    // one folded instruction: operands first (each leaves its value on the stack), then the instruction itself
    void flatten(const sexpr &e, const text_scope &sc) {
        if (!e.is_list) throw std::invalid_argument("wat: only folded instructions are supported inside a folded form (" + e.atom + ")");
        const std::string &h = e.head();
        const auto operands = [&](size_t from) { for (size_t i = from; i < e.list.size(); i++) flatten(e.list[i], sc); };
        if (h == "i64.const" || h == "i32.const") {
            if (e.list.size() != 2) throw std::invalid_argument("wat: " + h + " takes one literal");
            emit_const(h[1] == '3' ? 32 : 64, literal(h, e.list[1].atom));
            return;
        }
        if (h == "f32.const" || h == "f64.const") {
            if (e.list.size() != 2 || e.list[1].is_list) throw std::invalid_argument("wat: " + h + " takes one literal");
            emit_const(h[1] == '3' ? F32 : F64, parse_float(e.list[1].atom, h[1] == '3'));
            return;
        }
        const bool typed = h.size() > 4 && h[3] == '.' && (h[0] == 'i' || h[0] == 'f') && (h.compare(1, 2, "32") == 0 || h.compare(1, 2, "64") == 0);
        if (typed && (h.compare(4, 4, "load") == 0 || h.compare(4, 5, "store") == 0)) {
            size_t j = 1;
            const uint64_t offset = memarg(e.list, j);
            if (e.list.size() - j > (h[4] == 's' ? 2u : 1u)) throw std::invalid_argument("wat: " + h + " takes " + (h[4] == 's' ? "two folded operands" : "one folded operand"));   // fewer: the rest is on the stack already
            operands(j);
            if (!emit_access(h.substr(4), width_of(h.substr(0, 3)), offset, h)) throw std::invalid_argument("wat: unsupported instruction " + h);
            return;
        }
        if (typed) {
            const int n = emit_float(h, h, true);
            if (n) {
                if ((int)e.list.size() > 1 + n) throw std::invalid_argument("wat: " + h + " takes " + (n == 1 ? "one folded operand" : "two folded operands"));
                operands(1);
                emit_float(h, h);
                return;
            }
        }
        if (h == "global.get" || h == "global.set") {
            if (e.list.size() < 2 || e.list[1].is_list || e.list.size() > (h == "global.get" ? 2u : 3u)) throw std::invalid_argument("wat: malformed " + h);
            operands(2);
            emit_global(h == "global.get" ? ins::global_get : ins::global_set, global_index(e.list[1].atom, sc));
            return;
        }
        if (h == "memory.size" || h == "memory.grow" || h == "memory.fill" || h == "memory.copy") {
            const size_t n = h == "memory.size" ? 0 : (h == "memory.grow" ? 1 : 3);
            if (e.list.size() != 1 + n) throw std::invalid_argument("wat: malformed " + h);
            operands(1);
            emit_bulk(h == "memory.size" ? ins::memory_size : (h == "memory.grow" ? ins::memory_grow : (h == "memory.fill" ? ins::memory_fill : ins::memory_copy)));
            return;
        }
        if (h == "memory.init" || h == "data.drop") {
            if (e.list.size() != (h == "memory.init" ? 5u : 2u) || e.list[1].is_list) throw std::invalid_argument("wat: malformed " + h);
            operands(2);
            emit_bulk(h == "memory.init" ? ins::memory_init : ins::data_drop, data_index(e.list[1].atom, sc));
            return;
        }
        if (h.size() > 4 && (h.compare(0, 4, "i32.") == 0 || h.compare(0, 4, "i64.") == 0)) {
            opinfo oi;
            if (!lookup(h.substr(4), oi)) throw std::invalid_argument("wat: unsupported instruction " + h);
            const int n = oi.arity == 1 ? 1 : 2;
            if ((int)e.list.size() > 1 + n) throw std::invalid_argument("wat: " + h + " takes " + (n == 1 ? "one folded operand" : "two folded operands"));   // fewer: the rest is on the stack already
            operands(1);
            emit_op(h.substr(4), h[1] == '3' ? 32 : 64, h);
            return;
        }
        if (h == "call") {
            if (e.list.size() < 2 || e.list[1].is_list) throw std::invalid_argument("wat: call without a target");
            const uint64_t index = func_index(e.list[1].atom, sc);
            operands(2);
            emit_call(index, sc.imports);
            return;
        }
        if (h == "call_indirect") {                             // (call_indirect $table? (type $t) operand* index)
            size_t j = 1;
            if (j < e.list.size() && !e.list[j].is_list) { if (e.list[j].atom != "0" && e.list[j].atom[0] != '$') throw std::invalid_argument("wat: call_indirect through an unknown table"); j++; }
            const uint64_t t = type_use(e.list, j, sc);
            operands(j);
            emit_call_indirect(t, 0);
            return;
        }
        if (h == "local.get" || h == "local.set" || h == "local.tee") {
            if (e.list.size() < 2 || e.list[1].is_list || e.list.size() != (h == "local.get" ? 2u : 3u)) throw std::invalid_argument("wat: malformed " + h);
            operands(2);
            emit_local(h == "local.get" ? ins::local_get : (h == "local.set" ? ins::local_set : ins::local_tee), local_index(e.list[1].atom, sc));
            return;
        }
        if (h == "select") {
            size_t j = 1;
            while (j < e.list.size() && e.list[j].head() == "result") j++;                        // (select (result t) a b c)
            if (e.list.size() - j > 3) throw std::invalid_argument("wat: select takes three folded operands");
            operands(j);
            emit_plain(ins::select);
            return;
        }
        if (h == "ref.null") {
            if (e.list.size() != 2 || e.list[1].is_list) throw std::invalid_argument("wat: malformed ref.null");
            const std::string &t = e.list[1].atom;
            emit_reference(ins::ref_null, 0, 0, (t == "func" || t == "funcref") ? FUNCREF : ((t == "extern" || t == "externref") ? EXTERNREF : 0));
            return;
        }
        if (h == "ref.is_null") { if (e.list.size() > 2) throw std::invalid_argument("wat: malformed ref.is_null"); operands(1); emit_reference(ins::ref_is_null); return; }
        if (h == "ref.func") {
            if (e.list.size() != 2 || e.list[1].is_list) throw std::invalid_argument("wat: malformed ref.func");
            emit_reference(ins::ref_func, func_index(e.list[1].atom, sc));
            return;
        }
        if (h == "table.get" || h == "table.set" || h == "table.size" || h == "table.grow" || h == "table.fill" || h == "table.copy" || h == "table.init" || h == "elem.drop") {
            std::vector<std::string> imm;
            size_t j = 1;
            for (; j < e.list.size() && !e.list[j].is_list && imm.size() < 2; j++) imm.push_back(e.list[j].atom);
            operands(j);
            emit_table_text(h, imm, sc);
            return;
        }
        if (h == "block" || h == "loop") {                       // (block $l? (param ..)* (result ..)* instr*)
            size_t j = 1;
            std::string name;
            if (j < e.list.size() && !e.list[j].is_list && e.list[j].atom[0] == '$') name = e.list[j++].atom;
            std::vector<uint8_t> params, results;
            blocktype(e.list, j, params, results);
            emit_block(h == "block" ? ins::block : ins::loop, params, results, name);
            parse_seq(e.list, j, e.list.size(), sc);
            emit_end();
            return;
        }
        if (h == "if") {                                          // (if $l? (result ..)* condition* (then instr*) (else instr*)?)
            size_t j = 1;
            std::string name;
            if (j < e.list.size() && !e.list[j].is_list && e.list[j].atom[0] == '$') name = e.list[j++].atom;
            std::vector<uint8_t> params, results;
            blocktype(e.list, j, params, results);
            size_t then_at = j;
            while (then_at < e.list.size() && !(e.list[then_at].is_list && e.list[then_at].head() == "then")) then_at++;
            if (then_at >= e.list.size()) throw std::invalid_argument("wat: if without a (then ...) arm");
            for (size_t c = j; c < then_at; c++) flatten(e.list[c], sc);
            emit_block(ins::if_, params, results, name);
            parse_seq(e.list[then_at].list, 1, e.list[then_at].list.size(), sc);
            if (then_at + 1 < e.list.size()) {
                if (then_at + 2 != e.list.size() || e.list[then_at + 1].head() != "else") throw std::invalid_argument("wat: malformed if");
                emit_else();
                parse_seq(e.list[then_at + 1].list, 1, e.list[then_at + 1].list.size(), sc);
            }
            emit_end();
            return;
        }
        if (h == "br" || h == "br_if") {
            if (e.list.size() < 2 || e.list[1].is_list) throw std::invalid_argument("wat: " + h + " without a label");
            operands(2);
            emit_br(h == "br" ? ins::br : ins::br_if, label_depth(e.list[1].atom));
            return;
        }
        if (h == "br_table") {
            std::vector<uint32_t> targets;
            size_t j = 1;
            for (; j < e.list.size() && !e.list[j].is_list; j++) targets.push_back((uint32_t)label_depth(e.list[j].atom));
            operands(j);
            emit_br_table(targets);
            return;
        }
        if (h == "return") { operands(1); emit_return(); return; }
        if (h == "unreachable") { emit_unreachable(); return; }
        if (h == "drop") { operands(1); emit_plain(ins::drop); return; }
        if (h == "nop") { emit_plain(ins::nop); return; }
        throw std::invalid_argument("wat: unsupported instruction " + h);
    }

    // ---- binary (WebAssembly 1.0 module format + the sign-extension operators) ---------------------------------
    struct reader {
        const uint8_t *p, *end;
        uint8_t byte() { if (p >= end) throw std::invalid_argument("wasm: unexpected end of the module"); return *p++; }
        uint64_t uleb(int bits = 32) {
            uint64_t v = 0;
            for (int shift = 0;; shift += 7) {
                const uint8_t b = byte();
                if (shift >= bits + 7) throw std::invalid_argument("wasm: malformed LEB128 integer");
                v |= (uint64_t)(b & 0x7f) << (shift < 64 ? shift : 63);
                if (!(b & 0x80)) return v;
            }
        }
        int64_t sleb(int bits) {
            int64_t v = 0;
            int shift = 0;
            uint8_t b;
            do {
                b = byte();
                if (shift >= bits + 7) throw std::invalid_argument("wasm: malformed LEB128 integer");
                if (shift < 64) v = (int64_t)((uint64_t)v | ((uint64_t)(b & 0x7f) << shift));
                shift += 7;
            } while (b & 0x80);
            if (shift < 64 && (b & 0x40)) v = (int64_t)((uint64_t)v | (~0ULL << shift));
            return v;
        }
        std::string name() {
            const size_t n = (size_t)uleb();
            if ((size_t)(end - p) < n) throw std::invalid_argument("wasm: unexpected end of the module");
            std::string s((const char *)p, n);
            p += n;
            return s;
        }
        reader sub(size_t n) {
            if ((size_t)(end - p) < n) throw std::invalid_argument("wasm: section runs past the end of the module");
            reader r{p, p + n};
            p += n;
            return r;
        }
        uint8_t valtype() {
            const uint8_t t = byte();
            if (t == 0x7f) return 32;
            if (t == 0x7e) return 64;
            if (t == 0x7d) return F32;
            if (t == 0x7c) return F64;
            if (t == 0x70 || t == 0x6F) return t;
            throw std::invalid_argument("wasm: unknown value type");
        }
    };
    void parse_binary(const std::string &data) {
        reader r{(const uint8_t *)data.data(), (const uint8_t *)data.data() + data.size()};
        r.p += 4;
        if (r.sub(4).p[0] != 1) throw std::invalid_argument("wasm: unsupported binary version");
        struct sig { std::vector<uint8_t> params, results; bool usable = true; };
        std::vector<sig> types;
        std::vector<import_t> imports;
        std::vector<uint64_t> func_types;
        int64_t start = -1;
        std::vector<reader> bodies;
        std::vector<std::pair<uint64_t, std::vector<int64_t>>> pending_elems;   // (offset, functions): written once the functions are known
        while (r.p < r.end) {
            const uint8_t id = r.byte();
            reader s = r.sub((size_t)r.uleb());
            switch (id) {
            case 0: case 12: break;                           // custom / data count: nothing the subset needs
            case 8: break;                                    // start function: the reference's instantiate() does not run it (include/runtime.hpp never reads module.starts)
            case 5: {                                         // memory: at most one
                const size_t n = (size_t)s.uleb();
                if (n > 1 || (n && has_memory_)) throw std::invalid_argument("wasm: more than one memory");
                if (n) {
                    const uint8_t flags = s.byte();
                    if (flags > 1) throw std::invalid_argument("wasm: unsupported memory limits");
                    const uint64_t lo = s.uleb();
                    set_memory(lo, flags ? s.uleb() : 0);
                }
                break;
            }
            case 11: {                                        // data segments
                const size_t n = (size_t)s.uleb();
                for (size_t i = 0; i < n; i++) {
                    data_t d;
                    const uint64_t mode = s.uleb();
                    if (mode > 2) throw std::invalid_argument("wasm: malformed data segment");
                    if (mode == 2 && s.uleb() != 0) throw std::invalid_argument("wasm: data segment for an unknown memory");
                    if (mode != 1) {
                        if (s.byte() != 0x41) throw std::invalid_argument("wasm: a data segment's offset must be an i32.const");
                        d.active = true;
                        d.offset = (uint32_t)s.sleb(32);
                        if (s.byte() != 0x0B) throw std::invalid_argument("wasm: a data segment's offset must be an i32.const");
                    }
                    const size_t len = (size_t)s.uleb();
                    reader bytes = s.sub(len);
                    d.bytes.assign((const char *)bytes.p, len);
                    datas_.push_back(d);
                }
                break;
            }
            case 1: {                                         // types (a type naming other value types only matters if a module function uses it)
                const size_t n = (size_t)s.uleb();
                for (size_t i = 0; i < n; i++) {
                    if (s.byte() != 0x60) throw std::invalid_argument("wasm: malformed type section");
                    sig t;
                    for (int part = 0; part < 2; part++) {
                        const size_t cnt = (size_t)s.uleb();
                        for (size_t j = 0; j < cnt; j++) {
                            const uint8_t v = s.byte();
                            if (v != 0x7f && v != 0x7e && v != 0x7d && v != 0x7c && v != 0x70 && v != 0x6F) t.usable = false;
                            (part ? t.results : t.params).push_back(v == 0x7f ? 32 : (v == 0x7e ? 64 : (v == 0x7d ? F32 : (v == 0x7c ? F64 : v))));
                        }
                    }
                    types.push_back(t);
                    sigs_.push_back(sig_t{t.params, t.results});
                }
                break;
            }
            case 4: {                                         // tables: at most one, of function references
                const size_t n = (size_t)s.uleb();
                if (n > 1) throw std::invalid_argument("wasm: more than one table");
                if (n) {
                    if (s.byte() != 0x70) throw std::invalid_argument("wasm: only tables of function references are supported");
                    const uint8_t flags = s.byte();
                    if (flags > 1) throw std::invalid_argument("wasm: unsupported table limits");
                    set_table(s.uleb());
                    if (flags) s.uleb();
                }
                break;
            }
            case 9: {                                         // element segments: active ones fill the table; declarative ones say nothing
                const size_t n = (size_t)s.uleb();
                for (size_t i = 0; i < n; i++) {
                    const uint64_t flag = s.uleb();
                    if (flag > 7) throw std::invalid_argument("wasm: malformed element segment");
                    const bool active = !(flag & 1), exprs = (flag & 4) != 0;
                    if (!active && !(flag & 2)) throw std::invalid_argument("wasm: passive element segments are not supported (no table.init)");
                    uint64_t offset = 0;
                    if (active) {
                        if ((flag & 2) && s.uleb() != 0) throw std::invalid_argument("wasm: element segment for an unknown table");
                        if (s.byte() != 0x41) throw std::invalid_argument("wasm: an element segment's offset must be an i32.const");
                        offset = (uint32_t)s.sleb(32);
                        if (s.byte() != 0x0B) throw std::invalid_argument("wasm: an element segment's offset must be an i32.const");
                    }
                    if (flag & 3) { const uint8_t kind = s.byte(); if (kind != (exprs ? 0x70 : 0x00)) throw std::invalid_argument("wasm: only function elements are supported"); }
                    const size_t cnt = (size_t)s.uleb();
                    if (cnt > 1000000) throw std::invalid_argument("wasm: element segment too large");
                    std::vector<int64_t> fs;
                    for (size_t j = 0; j < cnt; j++) {
                        if (!exprs) { fs.push_back((int64_t)s.uleb()); continue; }
                        const uint8_t op = s.byte();           // ref.func f | ref.null func
                        if (op == 0xD2) fs.push_back((int64_t)s.uleb());
                        else if (op == 0xD0) { if (s.byte() != 0x70) throw std::invalid_argument("wasm: unsupported element expression"); fs.push_back(-1); }
                        else throw std::invalid_argument("wasm: unsupported element expression");
                        if (s.byte() != 0x0B) throw std::invalid_argument("wasm: unsupported element expression");
                    }
                    if (active) pending_elems.emplace_back(offset, std::move(fs));
                }
                break;
            }
            case 2: {                                         // imports: functions of env only
                const size_t n = (size_t)s.uleb();
                for (size_t i = 0; i < n; i++) {
                    import_t im{s.name(), s.name(), false, {}, {}};
                    if (s.byte() != 0x00) throw std::invalid_argument("wasm: only function imports are supported (" + printable(im.module + "." + im.field) + ")");
                    const uint64_t t = s.uleb();
                    if (t < types.size() && types[(size_t)t].usable) { im.typed = true; im.params = types[(size_t)t].params; im.results = types[(size_t)t].results; }
                    imports.push_back(im);
                }
                break;
            }
            case 3: {
                const size_t n = (size_t)s.uleb();
                for (size_t i = 0; i < n; i++) func_types.push_back(s.uleb());
                break;
            }
            case 6: {                                         // globals: valtype, mutability, a constant initialiser
                const size_t n = (size_t)s.uleb();
                for (size_t i = 0; i < n; i++) {
                    const uint8_t t = s.valtype(), mut = s.byte();
                    if (mut > 1) throw std::invalid_argument("wasm: malformed global");
                    if (is_float(t)) add_global(t, mut != 0, 0);   // (throws)
                    if (s.byte() != (t == 32 ? 0x41 : 0x42)) throw std::invalid_argument("wasm: a global's initialiser must be a constant of its type");
                    const uint64_t init = (uint64_t)s.sleb(t == 32 ? 32 : 64);
                    if (s.byte() != 0x0B) throw std::invalid_argument("wasm: a global's initialiser must be a constant of its type");
                    add_global(t, mut != 0, init);
                }
                break;
            }
            case 7: {                                         // exports: the function called _start
                const size_t n = (size_t)s.uleb();
                for (size_t i = 0; i < n; i++) {
                    const std::string nm = s.name();
                    const uint8_t kind = s.byte();
                    const uint64_t idx = s.uleb();
                    if (kind == 0x00 && nm == "_start") start = (int64_t)idx;
                }
                break;
            }
            case 10: {                                        // code
                const size_t n = (size_t)s.uleb();
                for (size_t i = 0; i < n; i++) bodies.push_back(s.sub((size_t)s.uleb()));
                break;
            }
            default:
                throw std::invalid_argument("wasm: unsupported module section (id " + std::to_string(id) + ")");
            }
        }
        nimports_ = imports.size();
        if (bodies.size() != func_types.size()) throw std::invalid_argument("wasm: function and code sections disagree");
        for (const data_t &d : datas_) if (d.active && !has_memory_) throw std::invalid_argument("wasm: an active data segment needs a memory");
        if (start < 0 || (size_t)start < imports.size() || (size_t)start - imports.size() >= bodies.size()) throw std::invalid_argument("wasm: no exported _start function");
        start_ = (size_t)start - imports.size();
        funcs_.resize(bodies.size());
        for (size_t k = 0; k < bodies.size(); k++) {
            if (func_types[k] >= types.size() || !types[(size_t)func_types[k]].usable) throw std::invalid_argument("wasm: function " + std::to_string(k) + " has an unsupported type");
            funcs_[k].params = funcs_[k].locals = types[(size_t)func_types[k]].params;
            funcs_[k].results = types[(size_t)func_types[k]].results;
            reader &b = bodies[k];
            const size_t groups = (size_t)b.uleb();
            for (size_t g = 0; g < groups; g++) {
                const uint64_t cnt = b.uleb();
                const uint8_t w = b.valtype();
                if (is_ref(w)) throw std::invalid_argument("wasm: locals of a reference type are not supported (the reference stops: Unsupported local type)");
                if (cnt > 10000 || funcs_[k].locals.size() + cnt > 10000) throw std::invalid_argument("wasm: too many locals");
                funcs_[k].locals.insert(funcs_[k].locals.end(), (size_t)cnt, w);
            }
        }
        if (!funcs_[start_].params.empty() || !funcs_[start_].results.empty()) throw std::invalid_argument("wasm: _start with parameters / results is not supported");
        for (const auto &pe : pending_elems) set_elements(pe.first, pe.second, imports.size());
        static const char *const int_ops[] = {"clz", "ctz", "popcnt", "add", "sub", "mul", "div_s", "div_u", "rem_s", "rem_u", "and", "or", "xor", "shl", "shr_s", "shr_u", "rotl", "rotr"};
        static const char *const cmp_ops[] = {"eqz", "eq", "ne", "lt_s", "lt_u", "gt_s", "gt_u", "le_s", "le_u", "ge_s", "ge_u"};
        for (size_t k = 0; k < bodies.size(); k++) {
            reader &b = bodies[k];
            begin_body(funcs_[k]);
            for (;;) {
                const uint8_t c = b.byte();
                if (c == 0x0B) {                              // end: of a block, or of the function
                    if (ctrl_.size() == 1) break;
                    emit_end();
                    continue;
                }
                const std::string shown = "0x" + std::string(1, "0123456789abcdef"[c >> 4]) + std::string(1, "0123456789abcdef"[c & 15]);
                if (c == 0x01) emit_plain(ins::nop);
                else if (c == 0x00) emit_unreachable();
                else if (c == 0x02 || c == 0x03 || c == 0x04) {
                    std::vector<uint8_t> params, results;
                    if (b.p < b.end && *b.p == 0x40) b.byte();
                    else if (b.p < b.end && (*b.p == 0x7f || *b.p == 0x7e || *b.p == 0x7d || *b.p == 0x7c || *b.p == 0x70 || *b.p == 0x6F)) results.push_back(b.valtype());
                    else {
                        const int64_t t = b.sleb(33);
                        if (t < 0 || (uint64_t)t >= types.size() || !types[(size_t)t].usable) throw std::invalid_argument("wasm: unsupported block type");
                        params = types[(size_t)t].params; results = types[(size_t)t].results;
                    }
                    emit_block(c == 0x02 ? ins::block : (c == 0x03 ? ins::loop : ins::if_), params, results, "");
                }
                else if (c == 0x05) emit_else();
                else if (c == 0x0C || c == 0x0D) emit_br(c == 0x0C ? ins::br : ins::br_if, b.uleb());
                else if (c == 0x0E) {
                    const uint64_t n = b.uleb();
                    if (n > 100000) throw std::invalid_argument("wasm: malformed br_table");
                    std::vector<uint32_t> targets;
                    for (uint64_t j = 0; j <= n; j++) targets.push_back((uint32_t)b.uleb());
                    emit_br_table(targets);
                }
                else if (c == 0x0F) emit_return();
                else if (c == 0x1A) emit_plain(ins::drop);
                else if (c == 0x1B) emit_plain(ins::select);
                else if (c == 0x1C) { const uint64_t n = b.uleb(); if (n != 1) throw std::invalid_argument("wasm: malformed typed select"); b.valtype(); emit_plain(ins::select); }
                else if (c == 0x25 || c == 0x26) emit_reference(c == 0x25 ? ins::table_get : ins::table_set, b.uleb());
                else if (c == 0xD0) emit_reference(ins::ref_null, 0, 0, b.valtype());
                else if (c == 0xD1) emit_reference(ins::ref_is_null);
                else if (c == 0xD2) emit_reference(ins::ref_func, b.uleb());
                else if (c == 0x10) emit_call(b.uleb(), imports);
                else if (c == 0x11) {
                    const uint64_t t = b.uleb(), tab = b.uleb();
                    if (t >= types.size() || !types[(size_t)t].usable) throw std::invalid_argument("wasm: call_indirect with an unsupported type");
                    emit_call_indirect(t, tab);
                }
                else if (c == 0x23 || c == 0x24) emit_global(c == 0x23 ? ins::global_get : ins::global_set, b.uleb());
                else if (c == 0x20 || c == 0x21 || c == 0x22) emit_local(c == 0x20 ? ins::local_get : (c == 0x21 ? ins::local_set : ins::local_tee), b.uleb());
                else if (c >= 0x28 && c <= 0x3E) {
                    static const char *const access[] = {"i32.load", "i64.load", "f32.load", "f64.load", "i32.load8_s", "i32.load8_u", "i32.load16_s", "i32.load16_u", "i64.load8_s", "i64.load8_u",
                                                         "i64.load16_s", "i64.load16_u", "i64.load32_s", "i64.load32_u", "i32.store", "i64.store", "f32.store", "f64.store", "i32.store8", "i32.store16",
                                                         "i64.store8", "i64.store16", "i64.store32"};
                    const char *nm = access[c - 0x28];
                    if (!nm) throw std::invalid_argument("wasm: unsupported instruction " + shown);
                    b.uleb();                                 // alignment hint
                    const uint64_t offset = b.uleb();
                    if (!emit_access(std::string(nm).substr(4), width_of(std::string(nm, 3)), offset, nm)) throw std::invalid_argument("wasm: unsupported instruction " + shown);
                }
                else if (c == 0x3F || c == 0x40) { if (b.byte() != 0) throw std::invalid_argument("wasm: unknown memory"); emit_bulk(c == 0x3F ? ins::memory_size : ins::memory_grow); }
                else if (c == 0xFC) {
                    const uint64_t sub = b.uleb();
                    if (sub == 8) { const uint64_t d = b.uleb(); if (b.byte() != 0) throw std::invalid_argument("wasm: unknown memory"); emit_bulk(ins::memory_init, d); }
                    else if (sub == 9) emit_bulk(ins::data_drop, b.uleb());
                    else if (sub == 10) { if (b.byte() != 0 || b.byte() != 0) throw std::invalid_argument("wasm: unknown memory"); emit_bulk(ins::memory_copy); }
                    else if (sub == 11) { if (b.byte() != 0) throw std::invalid_argument("wasm: unknown memory"); emit_bulk(ins::memory_fill); }
                    else if (sub == 12) { const uint64_t seg = b.uleb(), tab = b.uleb(); emit_reference(ins::table_init, seg, tab); }
                    else if (sub == 13) emit_reference(ins::elem_drop, b.uleb());
                    else if (sub == 14) { const uint64_t dst = b.uleb(), src = b.uleb(); emit_reference(ins::table_copy, dst, src); }
                    else if (sub == 15) emit_reference(ins::table_grow, b.uleb());
                    else if (sub == 16) emit_reference(ins::table_size, b.uleb());
                    else if (sub == 17) emit_reference(ins::table_fill, b.uleb());
                    else if (sub <= 7) {                        // iNN.trunc_sat_fMM_s/u
                        static const char *const sat[] = {"i32.trunc_sat_f32_s", "i32.trunc_sat_f32_u", "i32.trunc_sat_f64_s", "i32.trunc_sat_f64_u",
                                                          "i64.trunc_sat_f32_s", "i64.trunc_sat_f32_u", "i64.trunc_sat_f64_s", "i64.trunc_sat_f64_u"};
                        emit_float(sat[sub], sat[sub]);
                    }
                    else throw std::invalid_argument("wasm: unsupported instruction 0xfc " + std::to_string(sub));
                }
                else if (c == 0x41) emit_const(32, (uint64_t)b.sleb(32));
                else if (c == 0x42) emit_const(64, (uint64_t)b.sleb(64));
                else if (c == 0x43 || c == 0x44) {              // fNN.const: the bits, little endian
                    uint64_t bits = 0;
                    for (int j = 0; j < (c == 0x43 ? 4 : 8); j++) bits |= (uint64_t)b.byte() << (8 * j);
                    emit_const(c == 0x43 ? F32 : F64, bits);
                }
                else if ((c >= 0x5B && c <= 0x66) || (c >= 0x8B && c <= 0xA6)) {
                    static const char *const fcmp[] = {"eq", "ne", "lt", "gt", "le", "ge"};
                    static const char *const farith[] = {"abs", "neg", "ceil", "floor", "trunc", "nearest", "sqrt", "add", "sub", "mul", "div", "min", "max", "copysign"};
                    const bool cmp = c <= 0x66;
                    const int k = cmp ? c - 0x5B : c - 0x8B, per = cmp ? 6 : 14;
                    const std::string nm = std::string(k < per ? "f32." : "f64.") + (cmp ? fcmp : farith)[k % per];
                    emit_float(nm, nm);
                }
                else if ((c >= 0xA8 && c <= 0xAB) || (c >= 0xAE && c <= 0xBF)) {
                    static const char *const conv[] = {"i32.trunc_f32_s", "i32.trunc_f32_u", "i32.trunc_f64_s", "i32.trunc_f64_u", nullptr, nullptr,
                                                       "i64.trunc_f32_s", "i64.trunc_f32_u", "i64.trunc_f64_s", "i64.trunc_f64_u",
                                                       "f32.convert_i32_s", "f32.convert_i32_u", "f32.convert_i64_s", "f32.convert_i64_u", "f32.demote_f64",
                                                       "f64.convert_i32_s", "f64.convert_i32_u", "f64.convert_i64_s", "f64.convert_i64_u", "f64.promote_f32",
                                                       "i32.reinterpret_f32", "i64.reinterpret_f64", "f32.reinterpret_i32", "f64.reinterpret_i64"};
                    emit_float(conv[c - 0xA8], conv[c - 0xA8]);
                }
                else if (c >= 0x45 && c <= 0x4F) emit_op(cmp_ops[c - 0x45], 32, shown);
                else if (c >= 0x50 && c <= 0x5A) emit_op(cmp_ops[c - 0x50], 64, shown);
                else if (c >= 0x67 && c <= 0x78) emit_op(int_ops[c - 0x67], 32, shown);
                else if (c >= 0x79 && c <= 0x8A) emit_op(int_ops[c - 0x79], 64, shown);
                else if (c == 0xA7) emit_op("wrap_i64", 32, shown);
                else if (c == 0xAC) emit_op("extend_i32_s", 64, shown);
                else if (c == 0xAD) emit_op("extend_i32_u", 64, shown);
                else if (c == 0xC0) emit_op("extend8_s", 32, shown);
                else if (c == 0xC1) emit_op("extend16_s", 32, shown);
                else if (c == 0xC2) emit_op("extend8_s", 64, shown);
                else if (c == 0xC3) emit_op("extend16_s", 64, shown);
                else if (c == 0xC4) emit_op("extend32_s", 64, shown);
                else throw std::invalid_argument("wasm: unsupported instruction " + shown);
            }
            if (b.p != b.end) throw std::invalid_argument("wasm: bytes after the end of a function body");
            end_body("function " + std::to_string(k));
        }
    }

    std::vector<std::string> unprovided_;                     // names of the imports called in the text that the front end does not provide
    std::vector<std::vector<uint8_t>> args_;
    std::set<int> private_;
    bool echo_ = false;
    mutable int exit_code_ = -1;
    struct data_t { std::string bytes; bool active = false; uint32_t offset = 0; };
    struct sig_t { std::vector<uint8_t> params, results; };   // the module's function types (call_indirect names one)
    std::vector<sig_t> sigs_;
    bool has_table_ = false;
    std::vector<int64_t> table_;                              // table 0 after instantiation: function index (imports first), -1 = null (table_instance, runtime.hpp:96-102)
    std::vector<std::vector<int64_t>> elem_segs_;             // the active element segments, in order
    size_t nimports_ = 0;
    struct global_t { uint8_t type = 32; bool mut = false; uint64_t init = 0; };   // global_instance (runtime.hpp:181-187): i32 / i64 only (:441-455)
    std::vector<global_t> globals_;
    std::vector<func_t> funcs_;
    size_t start_ = 0;
    uint64_t step_limit_ = 200000000;
    bool has_memory_ = false;
    uint32_t mem_pages_ = 0, mem_max_ = 0;
    std::vector<data_t> datas_;
};

}  // namespace ligero::cuda::host

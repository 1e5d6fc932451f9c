// ref_contexts.cpp -- the REFERENCE's own prover layers run over an executor of our choice.  TEST INFRASTRUCTURE.
//
// What is compiled here is the reference's code where it lies under /root/reference (never copied into the repo):
//   include/zkp/nonbatch_context.hpp   stage 1/2/3 contexts (the row schedule)
//   include/zkp/backend/*.hpp          ligetron_backend + witness_manager (witness -> row packing, masks, randomness)
//   include/interpreter*.hpp           the WASM interpreter's opcode semantics over witnesses
//   include/host_modules/{env,vbn254fr,wasi_preview1}.hpp   the host modules the test programs call
//   include/zkp/{hash,merkle_tree,random}.hpp, src/bn254.cpp   transcript, tree, field
// and what this file adds is ONLY what `main` of src/webgpu_prover.cpp:228-471 does around them (the three passes, the
// seeds, the self-check) plus a hand-assembled instruction list where the reference would call wabt (absent here).
//
// Two builds (oracle/Makefile, target `refctx`):
//   -DREFCTX_CUDA   Executor = ligero::webgpu_context, which ligero-prover_b200/host/compat/wgpu.hpp makes the CUDA
//                   executor: the drop-in boundary compiled AND run inside the reference's translation unit (B200 only);
//   (default)       Executor = ligero::oracle_context (tests/refctx/oracle_executor.hpp): the same program on the CPU
//                   oracle, which is how tests/golden/refctx_*.json are produced in this GPU-less container.
// Output: one JSON file with the statement the stage contexts saw (row events, values, coefficient rows) and everything
// they produced (root, test vectors, openings).  tests/test_refctx_*.py compare the CUDA path and the restatements with it.
#include <algorithm>
#include <cstring>
#include <format>
#include <fstream>
#include <iomanip>
#include <map>
#include <memory>
#include <stdexcept>

#include <params.hpp>
#include <interpreter.hpp>
#include <wgpu.hpp>
#include <zkp/common.hpp>
#include <zkp/finite_field_gmp.hpp>
#include <zkp/nonbatch_context.hpp>
#include <host_modules/env.hpp>
#include <host_modules/vbn254fr.hpp>
#include <filesystem>
#include <unordered_set>
#include <host_modules/wasi_preview1.hpp>

#include <fiat_shamir.hpp>   // ligero-prover_b200/host: the sampler only (portable_sample needs Boost)

#ifdef REFCTX_CUDA
using executor_t = ligero::webgpu_context;
static const char *kExecutorName = "cuda";
#else
#include "oracle_executor.hpp"
using executor_t = ligero::oracle_context;
static const char *kExecutorName = "oracle";
#endif

using namespace ligero;
using namespace ligero::vm;
using field_t = zkp::bn254_gmp;
using buffer_t = executor_t::buffer_type;

// ---------------------------------------------------------------------------------------------------------------
// the log: what the stage contexts were asked to commit, in real-time order
enum { EV_LINEAR = 0, EV_QUAD = 1, EV_VSET = 2, EV_VCOPY = 3, EV_VADD = 4, EV_VSUB = 5, EV_VMUL = 6, EV_VDIV = 7, EV_VASSERT_EQ = 8, EV_VBIT = 9,
       EV_VADDC = 10, EV_VSUBC = 11, EV_VCSUB = 12, EV_VMULC = 13, EV_VMONTMULC = 14 };

struct run_log {
    std::vector<int> kinds;
    std::vector<uint32_t> rows;          // [rows][l][8]: first l elements of every scalar row / the l values of a VSET
    std::vector<uint32_t> batch_args;    // 3 per vbn254fr event
    std::vector<uint32_t> batch_consts;  // 8 per constant-taking event
    size_t l = 0;
    void row(const mpz_vector &v) {
        for (size_t i = 0; i < l; i++) {
            uint32_t limbs[8] = {0};
            if (i < v.size()) mpz_export(limbs, nullptr, -1, sizeof(uint32_t), 0, 0, v[i].get_mpz_t());
            rows.insert(rows.end(), limbs, limbs + 8);
        }
    }
};

// a stage context that writes down the rows it is handed, then does what the reference does
template <typename Base>
struct recording : Base {
    using witness_row_type = typename Base::witness_row_type;
    template <typename... A> explicit recording(run_log *lg, bool want_coefs, A &&...a) : Base(std::forward<A>(a)...), log_(lg), coefs_(want_coefs) {}
    void linear_callback(witness_row_type row) override {
        log_->kinds.push_back(EV_LINEAR);
        log_->row(coefs_ ? row.second : row.first);
        Base::linear_callback(row);
    }
    void quadratic_callback(witness_row_type x, witness_row_type y, witness_row_type z) override {
        log_->kinds.push_back(EV_QUAD);
        for (auto *r : {&x, &y, &z}) log_->row(coefs_ ? r->second : r->first);
        Base::quadratic_callback(x, y, z);
    }
    run_log *log_;
    bool coefs_;
};

using stage1_t = recording<zkp::nonbatch_stage1_context<field_t, executor_t, zkp::stage1_random_policy, params::hasher>>;
using stage2_t = recording<zkp::nonbatch_stage2_context<field_t, executor_t, zkp::stage2_random_policy>>;
using stage3_t = recording<zkp::nonbatch_stage3_context<field_t, executor_t, zkp::stage3_random_policy>>;

// ---------------------------------------------------------------------------------------------------------------
// programs.  Each one is a function template over the stage context, like run_program (include/invoke.hpp:79-98).

// the module store the reference's instantiate() would build (include/runtime.hpp:345-604) for a module that imports
// env.i64_private_const (func 0), env.assert_equal (func 1) and env.i32_private_const (func 2), has one 64 KiB memory and one
// function `_start` (func 3)
struct tiny_module {
    store_t store;
    module_instance inst;
    // env functions in import order: the call token "call:<name>" becomes call <index>
    // shape: 0 = (i64)->(i64), 1 = (i64 i64)->(), 2 = (i32)->(i32), 3 = (i64)->(), 4 = (i32 i32)->(i32), 5 = (i32 i32 i32 i32)->(i32), 6 = (i32)->()
    // names "wasi.<function>" are imports of wasi_snapshot_preview1, the others of env
    static const std::vector<std::pair<std::string, int>> &env_imports() {
        static const std::vector<std::pair<std::string, int>> t = {
            {"i64_private_const", 0}, {"assert_equal", 1}, {"i32_private_const", 2}, {"assert_zero", 3}, {"assert_one", 3}, {"assert_constant", 3},
            {"witness_cast_u32", 2}, {"witness_cast_u64", 0}, {"assert_is_concrete", 3},
            {"wasi.args_sizes_get", 4}, {"wasi.args_get", 4}, {"wasi.fd_write", 5}, {"wasi.proc_exit", 6}, {"wasi.random_get", 4}};
        return t;
    }
    static int import_index(const std::string &name) {
        const auto &t = env_imports();
        for (size_t i = 0; i < t.size(); i++) if (t[i].first == name) return (int)i;
        return -1;
    }
    // a function of the module itself: signature, declared locals, body; `_start` is functions[start]
    struct module_func {
        std::vector<value_kind> params, results, locals;
        std::vector<instr_ptr> body;
    };
    struct data_seg { std::vector<u8> bytes; bool active = false; u32 offset = 0; };
    explicit tiny_module(std::vector<module_func> functions, size_t start_function, u32 mem_pages = 1, u32 mem_max = 0, const std::vector<data_seg> &datas = {},
                         const std::vector<std::pair<value_kind, uint64_t>> &globals = {}, const std::vector<reference_t> *table = nullptr,
                         const std::vector<std::vector<reference_t>> &segments = {}) {
        function_kind k_pc({value_kind::i64}, {value_kind::i64}), k_eq({value_kind::i64, value_kind::i64}, {}), k_pc32({value_kind::i32}, {value_kind::i32}),
            k_one({value_kind::i64}, {}), k_w2({value_kind::i32, value_kind::i32}, {value_kind::i32}),
            k_w4({value_kind::i32, value_kind::i32, value_kind::i32, value_kind::i32}, {value_kind::i32}), k_w1({value_kind::i32}, {});
        inst.types = {k_pc, k_eq, k_pc32, k_one, k_w2, k_w4, k_w1};
        const function_kind *shapes[7] = {&k_pc, &k_eq, &k_pc32, &k_one, &k_w2, &k_w4, &k_w1};
        const auto &t = env_imports();
        for (size_t i = 0; i < t.size(); i++) {
            const bool wasi = t[i].first.rfind("wasi.", 0) == 0;
            const std::string fn = wasi ? t[i].first.substr(5) : t[i].first;
            inst.funcaddrs.push_back(store.emplace_back<function_instance>(name_t(fn), *shapes[t[i].second], &inst,
                                                                           function_instance::host_code{(index_t)i, wasi ? "wasi_snapshot_preview1" : "env", fn}));
        }
        for (size_t k = 0; k < functions.size(); k++) {
            function_kind kind(functions[k].params, functions[k].results);
            inst.types.push_back(kind);
            const index_t index = (index_t)(t.size() + k);
            inst.funcaddrs.push_back(store.emplace_back<function_instance>(name_t("f" + std::to_string(k)), kind, &inst,
                                                                           function_instance::func_code{index, functions[k].locals, std::move(functions[k].body)}));
        }
        inst.memaddrs.push_back(store.emplace_back<memory_instance>(mem_max ? memory_kind(limits(mem_pages, mem_max)) : memory_kind(limits(mem_pages)), (size_t)mem_pages * memory_instance::page_size));
        // data segments as instantiate() leaves them (include/runtime.hpp:537-556): active ones copied in (memory_init clears marks, none exist yet) and dropped
        for (const data_seg &d : datas) {
            const u32 i = store.emplace_back<data_instance>(d.bytes);
            inst.dataaddrs.push_back(i);
            if (d.active) {
                auto &mem = store.memorys[inst.memaddrs[0]];
                if ((size_t)d.offset + d.bytes.size() > mem.data.size()) throw std::runtime_error("data segment does not fit");
                std::copy(d.bytes.begin(), d.bytes.end(), mem.data.begin() + d.offset);
                store.datas[i].data.clear();
            }
        }
        if (table) inst.tableaddrs.push_back(store.emplace_back<table_instance>(table_kind{value_kind::funcref, limits((u32)table->size())}, *table));
        for (const auto &seg : segments) inst.elemaddrs.push_back(store.emplace_back<element_instance>(value_kind::funcref, seg));
        // globals as instantiate() creates them (include/runtime.hpp:428-456): i32 / i64, holding a native number
        for (const auto &g : globals) {
            if (g.first == value_kind::i32) inst.globaladdrs.push_back(store.emplace_back<global_instance>(g.first, (u32)g.second));
            else if (g.first == value_kind::i64) inst.globaladdrs.push_back(store.emplace_back<global_instance>(g.first, (u64)g.second));
            else throw std::runtime_error("Unexpected global value type");
        }
        inst.exports["_start"] = (index_t)(t.size() + start_function);
    }
    explicit tiny_module(std::vector<instr_ptr> body) : tiny_module(single(std::move(body)), 0) {}
    static std::vector<module_func> single(std::vector<instr_ptr> body) {
        std::vector<module_func> f(1);
        f[0].body = std::move(body);
        return f;
    }
};

// A program of the arithmetic-test subset as the flat instruction stream its folded text denotes (operands first):
//   iNN.const <value>    iNN.<op> (every integer instruction of the reference's tests)    call:<env function>
// (plus the short forms c / pc / eq / mul of the built-in i64_mul program)
// assembled the way transpile() (include/transpiler.hpp:741-776) would: runs of plain opcodes become basic blocks, calls
// stand alone.
struct wasm_token { std::string op; uint64_t imm = 0; std::vector<std::string> types; std::vector<uint64_t> targets; };

static value_kind token_kind(const std::string &t) {
    if (t == "i32") return value_kind::i32;
    if (t == "i64") return value_kind::i64;
    if (t == "f32") return value_kind::f32;
    if (t == "f64") return value_kind::f64;
    if (t == "funcref") return value_kind::funcref;
    if (t == "externref") return value_kind::externref;
    throw std::runtime_error("unknown value type " + t);
}
static std::vector<value_kind> token_kinds(const std::string &list) {      // "i32,i64" or "-"
    std::vector<value_kind> out;
    if (list == "-") return out;
    size_t b = 0;
    while (b <= list.size()) {
        const size_t e = list.find(',', b);
        const std::string t = list.substr(b, e == std::string::npos ? std::string::npos : e - b);
        out.push_back(token_kind(t));
        if (e == std::string::npos) break;
        b = e + 1;
    }
    return out;
}

// tokens from `pos` up to the "end" / "else" that closes the enclosing block (consumed; `stop` says which) or to the last token
static std::vector<instr_ptr> assemble_until(const std::vector<wasm_token> &toks, size_t &pos, std::string &stop);

static std::vector<instr_ptr> assemble(const std::vector<wasm_token> &toks) {
    size_t pos = 0;
    std::string stop;
    auto body = assemble_until(toks, pos, stop);
    if (!stop.empty() || pos != toks.size()) throw std::runtime_error("unbalanced block tokens");
    return body;
}

static std::vector<instr_ptr> assemble_until(const std::vector<wasm_token> &toks, size_t &pos, std::string &stop) {
    std::vector<instr_ptr> body;
    stop.clear();
    std::unique_ptr<basic_block> bb;
    size_t bb_id = 0;
    auto flush = [&] { if (bb) body.push_back(std::move(bb)); };
    auto plain = [&](opcode o) {
        if (!bb) { bb = std::make_unique<basic_block>(); bb->id = bb_id++; }
        bb->body.push_back(o);
    };
    // what transpile_opcode (include/transpiler.hpp:95-371) emits for the integer instructions
    static const std::map<std::string, std::pair<opcode::kind, sign_kind>> table = {
        {"clz", {opcode::inn_clz, sign_kind::unspecified}}, {"ctz", {opcode::inn_ctz, sign_kind::unspecified}},
        {"popcnt", {opcode::inn_popcnt, sign_kind::unspecified}}, {"eqz", {opcode::inn_eqz, sign_kind::unspecified}},
        {"add", {opcode::inn_add, sign_kind::unspecified}}, {"sub", {opcode::inn_sub, sign_kind::unspecified}},
        {"mul", {opcode::inn_mul, sign_kind::unspecified}},
        {"div_s", {opcode::inn_div_sx, sign_kind::sign}}, {"div_u", {opcode::inn_div_sx, sign_kind::unsign}},
        {"rem_s", {opcode::inn_rem_sx, sign_kind::sign}}, {"rem_u", {opcode::inn_rem_sx, sign_kind::unsign}},
        {"and", {opcode::inn_and, sign_kind::unspecified}}, {"or", {opcode::inn_or, sign_kind::unspecified}},
        {"xor", {opcode::inn_xor, sign_kind::unspecified}}, {"shl", {opcode::inn_shl, sign_kind::unspecified}},
        {"shr_s", {opcode::inn_shr_sx, sign_kind::sign}}, {"shr_u", {opcode::inn_shr_sx, sign_kind::unsign}},
        {"rotl", {opcode::inn_rotl, sign_kind::unspecified}}, {"rotr", {opcode::inn_rotr, sign_kind::unspecified}},
        {"eq", {opcode::inn_eq, sign_kind::unspecified}}, {"ne", {opcode::inn_ne, sign_kind::unspecified}},
        {"lt_s", {opcode::inn_lt_sx, sign_kind::sign}}, {"lt_u", {opcode::inn_lt_sx, sign_kind::unsign}},
        {"gt_s", {opcode::inn_gt_sx, sign_kind::sign}}, {"gt_u", {opcode::inn_gt_sx, sign_kind::unsign}},
        {"le_s", {opcode::inn_le_sx, sign_kind::sign}}, {"le_u", {opcode::inn_le_sx, sign_kind::unsign}},
        {"ge_s", {opcode::inn_ge_sx, sign_kind::sign}}, {"ge_u", {opcode::inn_ge_sx, sign_kind::unsign}},
        {"extend8_s", {opcode::inn_extend8_s, sign_kind::unspecified}}, {"extend16_s", {opcode::inn_extend16_s, sign_kind::unspecified}},
    };
    while (pos < toks.size()) {
        const wasm_token &t = toks[pos++];
        if (t.op == "end" || t.op == "else") { stop = t.op; break; }
        if (t.op == "block" || t.op == "loop" || t.op == "if") {          // block <params> <results>: what transpile_scope / transpile_if build (include/transpiler.hpp:649-700)
            flush();
            block_kind kind(token_kinds(t.types[0]), token_kinds(t.types[1]));
            std::string closed;
            std::vector<instr_ptr> inner = assemble_until(toks, pos, closed);
            if (t.op == "if") {
                if_then_else br;
                br.type = kind;
                br.then_body = std::move(inner);
                if (closed == "else") { br.else_body = assemble_until(toks, pos, closed); }
                if (closed != "end") throw std::runtime_error("if without end");
                body.push_back(make_instr<if_then_else>(std::move(br)));
            } else {
                if (closed != "end") throw std::runtime_error("block without end");
                if (t.op == "block") { scoped_block b; b.type = kind; b.body = std::move(inner); body.push_back(make_instr<scoped_block>(std::move(b))); }
                else { loop b; b.type = kind; b.body = std::move(inner); body.push_back(make_instr<loop>(std::move(b))); }
            }
            continue;
        }
        if (t.op == "br") { flush(); body.push_back(make_instr<br>((index_t)t.imm)); continue; }
        if (t.op == "br_if") { flush(); body.push_back(make_instr<br_if>((index_t)t.imm)); continue; }
        if (t.op == "br_table") {
            flush();
            std::vector<index_t> branches(t.targets.begin(), t.targets.end() - 1);
            body.push_back(make_instr<br_table>(std::move(branches), (index_t)t.targets.back()));
            continue;
        }
        if (t.op == "return") { flush(); body.push_back(make_instr<ret>()); continue; }
        if (t.op == "unreachable") { plain(opcode(opcode::unreachable)); continue; }
        // floating point: what transpile_opcode emits for fNN.* and the conversions (include/transpiler.hpp:128-161,237-270,325-352,373-440)
        static const std::map<std::string, opcode::kind> ftable = {
            {"abs", opcode::fnn_abs}, {"neg", opcode::fnn_neg}, {"ceil", opcode::fnn_ceil}, {"floor", opcode::fnn_floor}, {"trunc", opcode::fnn_trunc},
            {"nearest", opcode::fnn_nearest}, {"sqrt", opcode::fnn_sqrt}, {"add", opcode::fnn_add}, {"sub", opcode::fnn_sub}, {"mul", opcode::fnn_mul},
            {"div", opcode::fnn_div}, {"min", opcode::fnn_min}, {"max", opcode::fnn_max}, {"copysign", opcode::fnn_copysign},
            {"eq", opcode::fnn_eq}, {"ne", opcode::fnn_ne}, {"lt", opcode::fnn_lt}, {"gt", opcode::fnn_gt}, {"le", opcode::fnn_le}, {"ge", opcode::fnn_ge}};
        static const std::map<std::string, opcode::kind> plain_conv = {
            {"f32.demote_f64", opcode::f32_demote_f64}, {"f64.promote_f32", opcode::f64_promote_f32}, {"i32.reinterpret_f32", opcode::i32_reinterpret_f32},
            {"i64.reinterpret_f64", opcode::i64_reinterpret_f64}, {"f32.reinterpret_i32", opcode::f32_reinterpret_i32}, {"f64.reinterpret_i64", opcode::f64_reinterpret_i64}};
        static const std::map<std::string, opcode::kind> signed_conv = {
            {"f32.convert_i32_s", opcode::f32_convert_i32_s}, {"f32.convert_i32_u", opcode::f32_convert_i32_u}, {"f32.convert_i64_s", opcode::f32_convert_i64_s},
            {"f32.convert_i64_u", opcode::f32_convert_i64_u}, {"f64.convert_i32_s", opcode::f64_convert_i32_s}, {"f64.convert_i32_u", opcode::f64_convert_i32_u},
            {"f64.convert_i64_s", opcode::f64_convert_i64_s}, {"f64.convert_i64_u", opcode::f64_convert_i64_u},
            {"i32.trunc_f32_s", opcode::i32_trunc_f32_s}, {"i32.trunc_f32_u", opcode::i32_trunc_f32_u}, {"i32.trunc_f64_s", opcode::i32_trunc_f64_s},
            {"i32.trunc_f64_u", opcode::i32_trunc_f64_u}, {"i64.trunc_f32_s", opcode::i64_trunc_f32_s}, {"i64.trunc_f32_u", opcode::i64_trunc_f32_u},
            {"i64.trunc_f64_s", opcode::i64_trunc_f64_s}, {"i64.trunc_f64_u", opcode::i64_trunc_f64_u},
            {"i32.trunc_sat_f32_s", opcode::i32_trunc_sat_f32_s}, {"i32.trunc_sat_f32_u", opcode::i32_trunc_sat_f32_u}, {"i32.trunc_sat_f64_s", opcode::i32_trunc_sat_f64_s},
            {"i32.trunc_sat_f64_u", opcode::i32_trunc_sat_f64_u}, {"i64.trunc_sat_f32_s", opcode::i64_trunc_sat_f32_s}, {"i64.trunc_sat_f32_u", opcode::i64_trunc_sat_f32_u},
            {"i64.trunc_sat_f64_s", opcode::i64_trunc_sat_f64_s}, {"i64.trunc_sat_f64_u", opcode::i64_trunc_sat_f64_u}};
        if (plain_conv.count(t.op)) { plain(opcode(plain_conv.at(t.op))); continue; }
        if (signed_conv.count(t.op)) { plain(opcode(signed_conv.at(t.op), token_kind(t.op.substr(0, 3)), t.op.back() == 's' ? sign_kind::sign : sign_kind::unsign)); continue; }
        if (t.op.size() > 4 && (t.op.rfind("f32.", 0) == 0 || t.op.rfind("f64.", 0) == 0)) {
            const value_kind fk = token_kind(t.op.substr(0, 3));
            const std::string fname = t.op.substr(4);
            if (fname == "const") plain(fk == value_kind::f32 ? opcode(opcode::fnn_const, value_kind::f32, (uint32_t)t.imm) : opcode(opcode::fnn_const, value_kind::f64, t.imm));
            else if (fname == "load") plain(opcode(opcode::fnn_load, fk, sign_kind::unspecified, (u32)0, (u32)t.imm));
            else if (fname == "store") plain(opcode(opcode::fnn_store, fk, sign_kind::unspecified, (u32)0, (u32)t.imm));
            else if (ftable.count(fname)) plain(opcode(ftable.at(fname), fk));
            else throw std::runtime_error("unknown token " + t.op);
            continue;
        }
        if (t.op == "call_indirect") { flush(); body.push_back(make_instr<call_indirect>((index_t)0, (index_t)0)); continue; }   // (the type index is not read at run time)
        // references and the table (include/transpiler.hpp:546-553,612-641)
        if (t.op == "ref.null") { plain(opcode(opcode::ref_null, value_kind::funcref)); continue; }
        if (t.op == "ref.is_null") { plain(opcode(opcode::ref_is_null)); continue; }
        if (t.op == "ref.func") { plain(opcode(opcode::ref_func, (index_t)(tiny_module::env_imports().size() + t.imm))); continue; }
        if (t.op == "table.get") { plain(opcode(opcode::table_get, (index_t)0)); continue; }
        if (t.op == "table.set") { plain(opcode(opcode::table_set, (index_t)0)); continue; }
        if (t.op == "table.size") { plain(opcode(opcode::table_size, (index_t)0)); continue; }
        if (t.op == "table.grow") { plain(opcode(opcode::table_grow, (index_t)0)); continue; }
        if (t.op == "table.fill") { plain(opcode(opcode::table_fill, (index_t)0)); continue; }
        if (t.op == "table.copy") { plain(opcode(opcode::table_copy, (index_t)0, (index_t)0)); continue; }
        if (t.op == "table.init") { plain(opcode(opcode::table_init, (index_t)t.imm, (index_t)0)); continue; }
        if (t.op == "elem.drop") { plain(opcode(opcode::elem_drop, (index_t)t.imm)); continue; }
        if (t.op == "global.get") { plain(opcode(opcode::global_get, (index_t)t.imm)); continue; }
        if (t.op == "global.set") { plain(opcode(opcode::global_set, (index_t)t.imm)); continue; }
        const bool typed = t.op.size() > 4 && (t.op.rfind("i32.", 0) == 0 || t.op.rfind("i64.", 0) == 0);
        const value_kind vk = (typed && t.op[1] == '3') ? value_kind::i32 : value_kind::i64;
        const std::string name = typed ? t.op.substr(4) : std::string();
        if (typed && name == "const") plain(vk == value_kind::i32 ? opcode(opcode::inn_const, value_kind::i32, (uint32_t)t.imm) : opcode(opcode::inn_const, value_kind::i64, t.imm));
        else if (t.op == "i64.extend32_s") plain(opcode(opcode::i64_extend32_s));
        else if (t.op == "i64.extend_i32_s") plain(opcode(opcode::i64_extend_i32_sx, value_kind::i64, sign_kind::sign));
        else if (t.op == "i64.extend_i32_u") plain(opcode(opcode::i64_extend_i32_sx, value_kind::i64, sign_kind::unsign));
        else if (t.op == "i32.wrap_i64") plain(opcode(opcode::i32_wrap_i64));
        else if (typed && table.count(name)) {
            const auto &e = table.at(name);
            plain(e.second == sign_kind::unspecified ? opcode(e.first, vk) : opcode(e.first, vk, e.second));
        }
        else if (t.op == "c") plain(opcode(opcode::inn_const, value_kind::i64, t.imm));
        else if (t.op == "mul") plain(opcode(opcode::inn_mul, value_kind::i64));
        else if (t.op == "pc32") { flush(); body.push_back(make_instr<call>(2)); }
        else if (t.op == "pc") { flush(); body.push_back(make_instr<call>(0)); }
        else if (t.op == "eq") { flush(); body.push_back(make_instr<call>(1)); }
        else if (t.op.rfind("call:", 0) == 0 && tiny_module::import_index(t.op.substr(5)) >= 0) { flush(); body.push_back(make_instr<call>((index_t)tiny_module::import_index(t.op.substr(5)))); }
        else if (t.op == "drop") plain(opcode(opcode::drop));
        else if (t.op == "nop") plain(opcode(opcode::nop));
        else if (t.op == "select") plain(opcode(opcode::select, value_kind::unit, value_kind::unit));
        else if (t.op == "local.get") plain(opcode(opcode::local_get, (index_t)t.imm));
        else if (t.op == "local.set") plain(opcode(opcode::local_set, (index_t)t.imm));
        else if (t.op == "local.tee") plain(opcode(opcode::local_tee, (index_t)t.imm));
        else if (typed && (name.rfind("load", 0) == 0 || name.rfind("store", 0) == 0)) {
            const u32 align = 0, offset = (u32)t.imm;
            const sign_kind sg = name.size() > 2 && name.substr(name.size() - 2) == "_s" ? sign_kind::sign : (name.size() > 2 && name.substr(name.size() - 2) == "_u" ? sign_kind::unsign : sign_kind::unspecified);
            opcode::kind kd;
            if (name == "load") kd = opcode::inn_load;
            else if (name.rfind("load8", 0) == 0) kd = opcode::inn_load8_sx;
            else if (name.rfind("load16", 0) == 0) kd = opcode::inn_load16_sx;
            else if (name.rfind("load32", 0) == 0) kd = opcode::i64_load32_sx;
            else if (name == "store") kd = opcode::inn_store;
            else if (name == "store8") kd = opcode::inn_store8;
            else if (name == "store16") kd = opcode::inn_store16;
            else if (name == "store32") kd = opcode::i64_store32;
            else throw std::runtime_error("unknown token " + t.op);
            plain(opcode(kd, vk, sg, align, offset));
        }
        else if (t.op == "memory.size") plain(opcode(opcode::memory_size, (index_t)0));
        else if (t.op == "memory.grow") plain(opcode(opcode::memory_grow, (index_t)0));
        else if (t.op == "memory.fill") plain(opcode(opcode::memory_fill, (index_t)0));
        else if (t.op == "memory.copy") plain(opcode(opcode::memory_copy, (index_t)0, (index_t)0));
        else if (t.op == "memory.init") plain(opcode(opcode::memory_init, (index_t)t.imm));
        else if (t.op == "data.drop") plain(opcode(opcode::data_drop, (index_t)t.imm));
        else if (t.op == "callf") { flush(); body.push_back(make_instr<call>((index_t)(tiny_module::env_imports().size() + t.imm))); }
        else throw std::runtime_error("unknown token " + t.op);
    }
    flush();
    return body;
}

struct binop_case { uint64_t a, b, c; };

// tests/i64_mul.wat:5-14, the nine assertions in order
static const binop_case kI64Mul[] = {
    {1, 1, 1},
    {1, 0, 0},
    {~0ULL, ~0ULL, 1},
    {0x1000000000000000ULL, 4096, 0},
    {0x8000000000000000ULL, 0, 0},
    {0x8000000000000000ULL, ~0ULL, 0x8000000000000000ULL},
    {0x7fffffffffffffffULL, ~0ULL, 0x8000000000000001ULL},
    {0x0123456789abcdefULL, 0xfedcba9876543210ULL, 0x2236d88fe5618cf0ULL},
    {0x7fffffffffffffffULL, 0x7fffffffffffffffULL, 1},
};

// (call $assert_equal (i64.OP (call $pc (i64.const a)) (call $pc (i64.const b))) (call $pc (i64.const c)))
static std::vector<wasm_token> binop_tokens(const char *op, const binop_case *cases, size_t ncases) {
    std::vector<wasm_token> t;
    for (size_t i = 0; i < ncases; i++) {
        t.push_back({"c", cases[i].a}); t.push_back({"pc"});
        t.push_back({"c", cases[i].b}); t.push_back({"pc"});
        t.push_back({op});
        t.push_back({"c", cases[i].c}); t.push_back({"pc"});
        t.push_back({"eq"});
    }
    return t;
}

static std::vector<wasm_token> read_tokens(const std::string &path) {
    std::ifstream in(path);
    if (!in) throw std::runtime_error("cannot read " + path);
    std::vector<wasm_token> t;
    std::string op;
    while (in >> op) {
        wasm_token tok{op};
        if (op == "c" || op == "i32.const" || op == "i64.const" || op == "f32.const" || op == "f64.const" || op == "global.get" || op == "global.set" || op == "ref.func" || op == "table.init" || op == "elem.drop" || op == "local.get" || op == "local.set" || op == "local.tee" || op == "callf" || op == "start" || op == "memory.init" || op == "data.drop" || op == "br" || op == "br_if" ||
            ((op.rfind("i32.", 0) == 0 || op.rfind("i64.", 0) == 0 || op.rfind("f32.", 0) == 0 || op.rfind("f64.", 0) == 0) && (op.find(".load") != std::string::npos || op.find(".store") != std::string::npos))) {
            std::string lit; in >> lit; tok.imm = std::stoull(lit, nullptr, 0);
        }
        if (op == "func") {                                   // func <params> <results> <locals>, e.g. "func i64,i32 i64 -": a new module function starts
            for (int j = 0; j < 3; j++) { std::string part; in >> part; tok.types.push_back(part); }
        }
        if (op == "block" || op == "loop" || op == "if") {    // block <params> <results>
            for (int j = 0; j < 2; j++) { std::string part; in >> part; tok.types.push_back(part); }
        }
        if (op == "br_table") {                               // br_table <count> <targets ... default>
            size_t n; in >> n;
            for (size_t j = 0; j < n; j++) { uint64_t l; in >> l; tok.targets.push_back(l); }
        }
        if (op == "table") { std::string part; in >> part; tok.imm = std::stoull(part); }          // table <size>
        if (op == "elem") {                                   // elem <offset> <count> <module function indices, - for null>
            std::string part; in >> part; tok.imm = std::stoull(part);
            size_t n; in >> n;
            for (size_t j = 0; j < n; j++) { in >> part; tok.types.push_back(part); }
        }
        if (op == "arg") {                                    // arg public|private <hex or ->: one more argument for wasi args_get
            for (int j = 0; j < 2; j++) { std::string part; in >> part; tok.types.push_back(part); }
        }
        if (op == "global") {                                 // global <type> <initial value>
            std::string part; in >> part; tok.types.push_back(part);
            std::string lit; in >> lit; tok.imm = std::stoull(lit, nullptr, 0);
        }
        if (op == "memory") {                                 // memory <pages> <max pages or 0>
            for (int j = 0; j < 2; j++) { std::string part; in >> part; tok.types.push_back(part); }
        }
        if (op == "data") {                                   // data passive <hex or -> | data active <offset> <hex or ->
            std::string mode; in >> mode; tok.types.push_back(mode);
            for (int j = 0; j < (mode == "active" ? 2 : 1); j++) { std::string part; in >> part; tok.types.push_back(part); }
        }
        t.push_back(tok);
    }
    return t;
}

// a token stream with "func" headers is a module of several functions ("start K" names _start); without, one function.
// "memory" and "data" directives describe the module's memory and data segments
static tiny_module build_module(const std::vector<wasm_token> &toks) {
    const auto kinds = [](const std::string &list) { return token_kinds(list); };
    std::vector<tiny_module::module_func> functions;
    std::vector<tiny_module::data_seg> datas;
    std::vector<std::pair<value_kind, uint64_t>> globals;
    std::vector<reference_t> table;
    std::vector<std::vector<reference_t>> segments;
    bool has_table = false;
    std::vector<wasm_token> body;
    size_t start = 0;
    u32 pages = 1, max_pages = 0;
    bool structured = false;
    for (const wasm_token &t : toks) structured |= t.op == "func";
    if (!structured) functions.emplace_back();
    const auto close = [&] { if (!functions.empty()) functions.back().body = assemble(body); body.clear(); };
    for (const wasm_token &t : toks) {
        if (t.op == "func") {
            close();
            functions.emplace_back();
            functions.back().params = kinds(t.types[0]); functions.back().results = kinds(t.types[1]); functions.back().locals = kinds(t.types[2]);
        } else if (t.op == "start") start = (size_t)t.imm;
        else if (t.op == "arg") continue;
        else if (t.op == "table") { has_table = true; table.assign((size_t)t.imm, std::nullopt); }
        else if (t.op == "elem") {                            // what table_init leaves after instantiation (include/runtime.hpp:518-536): function addresses = indices here
            segments.emplace_back();
            for (size_t j = 0; j < t.types.size(); j++) {
                const reference_t r = t.types[j] == "-" ? reference_t{} : reference_t{(index_t)(tiny_module::env_imports().size() + std::stoul(t.types[j]))};
                table.at((size_t)t.imm + j) = r;
                segments.back().push_back(r);                 // active segments stay in the store (instantiate() only drops the others)
            }
        }
        else if (t.op == "global") globals.emplace_back(token_kind(t.types[0]), t.imm);
        else if (t.op == "memory") { pages = (u32)std::stoul(t.types[0]); max_pages = (u32)std::stoul(t.types[1]); }
        else if (t.op == "data") {
            tiny_module::data_seg d;
            d.active = t.types[0] == "active";
            if (d.active) d.offset = (u32)std::stoul(t.types[1]);
            const std::string &hx = t.types.back();
            if (hx != "-") for (size_t i = 0; i + 1 < hx.size(); i += 2) d.bytes.push_back((u8)std::stoul(hx.substr(i, 2), nullptr, 16));
            datas.push_back(d);
        }
        else body.push_back(t);
    }
    close();
    return tiny_module(std::move(functions), start, pages, max_pages, datas, globals, has_table ? &table : nullptr, segments);
}

template <typename Ctx>
static void run_wasm_tokens(Ctx &ctx, const std::vector<wasm_token> &toks) {
    tiny_module m = build_module(toks);
    wasm_interpreter<Ctx> interp(ctx);
    ctx.store(&m.store);
    ctx.module(&m.inst);
    // invoke() (include/invoke.hpp:34-77)
    auto dummy = ctx.make_frame();
    dummy->module = &m.inst;
    ctx.set_current_frame(dummy.get());
    ctx.stack_push(std::move(dummy));
    std::vector<std::vector<u8>> args;
    std::unordered_set<int> private_indices;
    for (const wasm_token &t : toks) {
        if (t.op != "arg") continue;
        if (t.types[0] == "private") private_indices.insert((int)args.size());
        std::vector<u8> bytes;
        const std::string &hx = t.types[1];
        if (hx != "-") for (size_t i = 0; i + 1 < hx.size(); i += 2) bytes.push_back((u8)std::stoul(hx.substr(i, 2), nullptr, 16));
        args.push_back(bytes);
    }
    ctx.template add_host_module<wasi_preview1_module<Ctx>>(&ctx, args, private_indices);
    ctx.template add_host_module<env_module<Ctx>>(&ctx);
    ctx.template add_host_module<vbn254fr_module<Ctx>>(&ctx);
    auto result = interp.run(call{m.inst.exports["_start"]});
    if (result.is_exit()) std::cerr << "Exit with code " << result.exit_code() << std::endl;      // as invoke() does: the run ends, the stage goes on
    ctx.stack_pop();
    ctx.finalize();
}

// vbn254fr host calls on device-resident variables, driven through the module's own call table with the arguments on
// the interpreter stack and the handles in WASM memory, as compiled guest code would (include/host_modules/vbn254fr.hpp)
template <typename Ctx>
struct vbn_driver {
    Ctx &ctx;
    run_log *log;
    uint32_t k, l;
    uint32_t next_addr = 64;             // guest addresses of the 4-byte handles
    uint32_t scratch = 4096;             // guest scratch for constants / strings
    uint32_t new_var() {
        uint32_t a = next_addr; next_addr += 4;
        ctx.stack_push(a);
        call("vbn254fr_alloc");
        return a;
    }
    uint32_t slot(uint32_t addr) { return ctx.template memory_load<uint32_t>(addr) / k; }
    void call(const char *fn) { ctx.call_host(0, "vbn254fr", fn); }
    void ev(int kind, uint32_t a0, uint32_t a1, uint32_t a2) {
        if (!log) return;
        log->kinds.push_back(kind);
        log->batch_args.insert(log->batch_args.end(), {a0, a1, a2});
    }
    void set_ui_scalar(uint32_t out, uint32_t v) {
        ev(EV_VSET, slot(out), 0, 0);
        if (log) { mpz_vector row; for (uint32_t i = 0; i < l; i++) row.push_back(mpz_class(v)); log->row(row); }
        ctx.stack_push(out); ctx.stack_push(v);
        call("vbn254fr_set_ui_scalar");
    }
    void set_str_scalar(uint32_t out, const std::string &s, uint32_t base) {
        mpz_class val(s, (int)base);
        ev(EV_VSET, slot(out), 0, 0);
        if (log) { mpz_vector row; for (uint32_t i = 0; i < l; i++) row.push_back(val); log->row(row); }
        std::memcpy(ctx.memory_data().data() + scratch, s.c_str(), s.size() + 1);
        ctx.stack_push(out); ctx.stack_push(scratch); ctx.stack_push(base);
        call("vbn254fr_set_str_scalar");
        ctx.stack_pop();                                      // the error code the call leaves behind
    }
    void op3(int kind, const char *fn, uint32_t out, uint32_t x, uint32_t y) {
        ev(kind, slot(out), slot(x), slot(y));
        ctx.stack_push(out); ctx.stack_push(x); ctx.stack_push(y);
        call(fn);
    }
    void opc(int kind, const char *fn, uint32_t out, uint32_t x, const mpz_class &c, bool const_first = false) {
        ev(kind, slot(out), slot(x), 0);
        uint32_t limbs[8] = {0};
        mpz_export(limbs, nullptr, -1, sizeof(uint32_t), 0, 0, c.get_mpz_t());
        if (log) log->batch_consts.insert(log->batch_consts.end(), limbs, limbs + 8);
        std::memcpy(ctx.memory_data().data() + scratch, limbs, 32);
        ctx.stack_push(out);
        if (const_first) { ctx.stack_push(scratch); ctx.stack_push(x); }
        else { ctx.stack_push(x); ctx.stack_push(scratch); }
        call(fn);
    }
    void copy(uint32_t out, uint32_t in) {
        ev(EV_VCOPY, slot(out), slot(in), 0);
        ctx.stack_push(out); ctx.stack_push(in);
        call("vbn254fr_copy");
    }
    void assert_equal(uint32_t x, uint32_t y) {
        ev(EV_VASSERT_EQ, slot(x), slot(y), 0);
        ctx.stack_push(x); ctx.stack_push(y);
        call("vbn254fr_assert_equal");
    }
    // all 254 bits: the call commits one row per bit (vbn254fr.hpp:548-565)
    void bit_decompose(uint32_t arr_addr, const std::vector<uint32_t> &outs, uint32_t x) {
        for (uint32_t i = 0; i < outs.size(); i++) {
            ev(EV_VBIT, slot(outs[i]), slot(x), i);
            ctx.template memory_store<uint32_t>(arr_addr + 4 * i, ctx.template memory_load<uint32_t>(outs[i]));
        }
        ctx.stack_push(arr_addr); ctx.stack_push(x);
        call("vbn254fr_bit_decompose");
    }
};

template <typename Ctx>
static void run_vbn_program(Ctx &ctx, run_log *log, uint32_t l, uint32_t k) {
    tiny_module m({});
    ctx.store(&m.store);
    ctx.module(&m.inst);
    auto dummy = ctx.make_frame();
    dummy->module = &m.inst;
    ctx.set_current_frame(dummy.get());
    ctx.stack_push(std::move(dummy));
    ctx.template add_host_module<env_module<Ctx>>(&ctx);
    ctx.template add_host_module<vbn254fr_module<Ctx>>(&ctx);

    vbn_driver<Ctx> v{ctx, log, k, l};
    uint32_t a = v.new_var(), b = v.new_var(), c = v.new_var(), d = v.new_var(), e = v.new_var(), f = v.new_var(), g = v.new_var();
    v.set_ui_scalar(a, 0x12345u);
    v.set_str_scalar(b, "1234567890123456789012345678901234567890123456789012345678901234567", 10);
    v.op3(EV_VADD, "vbn254fr_addmod", c, a, b);
    v.op3(EV_VMUL, "vbn254fr_mulmod", d, c, b);
    v.op3(EV_VMUL, "vbn254fr_mulmod", e, a, a);               // x == y: the y row is a copy of the encoded x
    v.op3(EV_VSUB, "vbn254fr_submod", f, d, e);
    v.op3(EV_VDIV, "vbn254fr_divmod", g, f, b);
    v.op3(EV_VMUL, "vbn254fr_mulmod", c, g, b);
    v.assert_equal(c, f);
    v.copy(d, g);
    v.opc(EV_VADDC, "vbn254fr_addmod_constant", e, d, mpz_class("987654321987654321987654321", 10));
    v.opc(EV_VSUBC, "vbn254fr_submod_constant", e, e, mpz_class(77));
    v.opc(EV_VCSUB, "vbn254fr_constant_submod", f, e, mpz_class("123456789123456789", 10), true);
    v.opc(EV_VMULC, "vbn254fr_mulmod_constant", f, f, mpz_class("55555555555555555555", 10));
    v.opc(EV_VMONTMULC, "vbn254fr_mont_mul_constant", g, f, mpz_class("31415926535897932384626433832795", 10));
    std::vector<uint32_t> bits;
    for (uint32_t i = 0; i < field_t::num_bits; i++) bits.push_back(v.new_var());
    v.bit_decompose(8192, bits, a);
    // a few scalar witnesses between and after the batch rows, so both kinds of row interleave
    {
        auto x = ctx.backend().acquire_witness(); x.val(123456789u);
        auto y = ctx.backend().acquire_witness(); y.val(987654321u);
        auto z = ctx.backend().eval(x * y);
        auto w = ctx.backend().acquire_witness(); w.val(121932631112635269ull);
        ctx.backend().assert_equal(z, w);
    }
    ctx.stack_pop();
    ctx.finalize();
}

template <typename Ctx>
static void run_named(const std::string &prog, Ctx &ctx, run_log *log, uint32_t l, uint32_t k) {
    if (prog == "i64_mul") run_wasm_tokens(ctx, binop_tokens("mul", kI64Mul, std::size(kI64Mul)));
    else if (prog == "i64_mul3") run_wasm_tokens(ctx, binop_tokens("mul", kI64Mul + 6, 3));
    else if (prog.rfind("ops:", 0) == 0) run_wasm_tokens(ctx, read_tokens(prog.substr(4)));   // any program of the subset, as tokens
    else if (prog == "vbn") run_vbn_program(ctx, log, l, k);
    else throw std::runtime_error("unknown program " + prog);
}

// ---------------------------------------------------------------------------------------------------------------
static std::string hex(const void *p, size_t n) {
    static const char d[] = "0123456789abcdef";
    std::string s(2 * n, '0');
    for (size_t i = 0; i < n; i++) { unsigned char c = ((const unsigned char *)p)[i]; s[2 * i] = d[c >> 4]; s[2 * i + 1] = d[c & 15]; }
    return s;
}
template <typename T> static std::string hexv(const std::vector<T> &v) { return hex(v.data(), v.size() * sizeof(T)); }

// ---------------------------------------------------------------------------------------------------------------
// the reference's verifier (src/webgpu_verifier.cpp:262-449) around its own nonbatch_verifier_context
using verifier_t = zkp::nonbatch_verifier_context<field_t, executor_t, zkp::verifier_random_policy, params::hasher>;

struct proof_fields {
    params::hasher::digest root;
    std::vector<uint32_t> code, linear, quad, samplings;
    std::vector<std::pair<size_t, params::hasher::digest>> siblings;      // tree position -> digest
    size_t total_count = 0;
};

struct verdict { bool merkle, code, linear, quad, code_eq, linear_eq, quad_eq; bool all() const { return merkle && code && linear && quad && code_eq && linear_eq && quad_eq; } };

static verdict verify(const std::string &prog, executor_t &executor, size_t l, size_t k, size_t n, const params::hasher::digest &instance_hash,
                      const proof_fields &pf) {
    auto stage1_seed = zkp::hash<params::hasher>("LigetronStage1", pf.root, instance_hash);
    auto sample_seed = zkp::hash<params::hasher>("LigetronStage2", pf.root, pf.code, pf.linear, pf.quad);
    unsigned char seed[params::hasher::digest_size];
    std::memcpy(seed, stage1_seed.data, params::hasher::digest_size);
    cuda::host::digest s2;
    std::memcpy(s2.data, sample_seed.data, 32);
    std::vector<uint64_t> sample64 = cuda::host::sample_indices(s2, n, params::sample_size);       // (Boost-free sampler, see main)
    std::vector<size_t> sample_index(sample64.begin(), sample64.end());

    zkp::merkle_tree<params::hasher>::decommitment decommit(pf.total_count, sample_index);
    for (const auto &sb : pf.siblings) decommit.insert(sb.first, sb.second);

    auto vctx = std::make_unique<verifier_t>(executor, sample_index, pf.samplings);
    vctx->init_witness_random(seed, params::any_iv);
    run_log unused;
    unused.l = l;
    run_named(prog, *vctx, nullptr, l, k);
    auto vs1_root = zkp::merkle_tree<params::hasher>::recommit(vctx->flush_digests(), decommit);

    auto linear_sums = vctx->linear_sums();
    mpz_vector vsample_code, vsample_linear, vsample_quad;
    auto vc = executor.template copy_to_host<uint32_t>(vctx->code());
    auto vl = executor.template copy_to_host<uint32_t>(vctx->linear());
    auto vq = executor.template copy_to_host<uint32_t>(vctx->quadratic());
    vsample_code.import_limbs(vc.data(), vc.size(), sizeof(uint32_t), field_t::num_u32_limbs);
    vsample_linear.import_limbs(vl.data(), vl.size(), sizeof(uint32_t), field_t::num_u32_limbs);
    vsample_quad.import_limbs(vq.data(), vq.size(), sizeof(uint32_t), field_t::num_u32_limbs);

    buffer_t device_code = executor.make_codeword_buffer(), device_linear = executor.make_codeword_buffer(), device_quad = executor.make_codeword_buffer();
    executor.write_buffer(device_code, pf.code.data(), pf.code.size());
    executor.write_buffer(device_linear, pf.linear.data(), pf.linear.size());
    executor.write_buffer(device_quad, pf.quad.data(), pf.quad.size());
    executor.decode_ntt_device(executor.bind_ntt(device_code));
    executor.decode_ntt_device(executor.bind_ntt(device_linear));
    executor.decode_ntt_device(executor.bind_ntt(device_quad));
    mpz_vector prover_code, prover_linear, prover_quad, enc_code, enc_linear, enc_quad;
    { auto limbs = executor.template copy_to_host<uint32_t>(device_code); prover_code.import_limbs(limbs.data(), limbs.size(), sizeof(uint32_t), field_t::num_u32_limbs); }
    { auto limbs = executor.template copy_to_host<uint32_t>(device_linear); prover_linear.import_limbs(limbs.data(), limbs.size(), sizeof(uint32_t), field_t::num_u32_limbs); prover_linear.resize(l); }
    { auto limbs = executor.template copy_to_host<uint32_t>(device_quad); prover_quad.import_limbs(limbs.data(), limbs.size(), sizeof(uint32_t), field_t::num_u32_limbs); prover_quad.resize(l); }
    enc_code.import_limbs(pf.code.data(), pf.code.size(), sizeof(uint32_t), field_t::num_u32_limbs);
    enc_linear.import_limbs(pf.linear.data(), pf.linear.size(), sizeof(uint32_t), field_t::num_u32_limbs);
    enc_quad.import_limbs(pf.quad.data(), pf.quad.size(), sizeof(uint32_t), field_t::num_u32_limbs);

    verdict v;
    v.merkle = pf.root == vs1_root;
    v.code = std::all_of(prover_code.begin() + k, prover_code.end(), [](const auto &x) { return x == 0; });
    v.linear = zkp::validate_sum<field_t>(prover_linear, linear_sums);
    v.quad = zkp::validate(prover_quad);
    v.code_eq = v.linear_eq = v.quad_eq = true;
    for (size_t i = 0; i < params::sample_size; i++) {
        v.code_eq &= enc_code[sample_index[i]] == vsample_code[i];
        v.linear_eq &= enc_linear[sample_index[i]] == vsample_linear[i];
        v.quad_eq &= enc_quad[sample_index[i]] == vsample_quad[i];
    }
    return v;
}

static std::vector<unsigned char> unhex(const std::string &h) {
    std::vector<unsigned char> out(h.size() / 2);
    auto nib = [](char c) { return (unsigned)(c <= '9' ? c - '0' : (c | 32) - 'a' + 10); };
    for (size_t i = 0; i < out.size(); i++) out[i] = (unsigned char)(nib(h[2 * i]) << 4 | nib(h[2 * i + 1]));
    return out;
}
template <typename T> static std::vector<T> unhex_as(const std::string &h) {
    std::vector<unsigned char> b = unhex(h);
    std::vector<T> v(b.size() / sizeof(T));
    std::memcpy(v.data(), b.data(), v.size() * sizeof(T));
    return v;
}

// a proof handed in from outside (e.g. the envelope of lgrp_prove): lines of `key hex...`
static proof_fields read_proof(const std::string &path, params::hasher::digest &instance_hash) {
    std::ifstream in(path);
    if (!in) throw std::runtime_error("cannot read " + path);
    proof_fields pf;
    std::string key;
    while (in >> key) {
        if (key == "total") { in >> pf.total_count; continue; }
        if (key == "sibling") {
            size_t pos; std::string h; in >> pos >> h;
            params::hasher::digest d; auto b = unhex(h); std::memcpy(d.data, b.data(), 32);
            pf.siblings.emplace_back(pos, d);
            continue;
        }
        std::string h; in >> h;
        if (key == "root") { auto b = unhex(h); std::memcpy(pf.root.data, b.data(), 32); }
        else if (key == "instance") { auto b = unhex(h); std::memcpy(instance_hash.data, b.data(), 32); }
        else if (key == "code") pf.code = unhex_as<uint32_t>(h);
        else if (key == "linear") pf.linear = unhex_as<uint32_t>(h);
        else if (key == "quad") pf.quad = unhex_as<uint32_t>(h);
        else if (key == "samplings") pf.samplings = unhex_as<uint32_t>(h);
        else throw std::runtime_error("unknown key " + key);
    }
    return pf;
}

int main(int argc, char **argv) {
    if (argc < 4) {
        std::fprintf(stderr, "usage: %s <program> <k> <out.json> [seed byte]\n       %s verify:<proof file> <program> <k>\n", argv[0], argv[0]);
        return 2;
    }
    const bool verify_only = std::string(argv[1]).rfind("verify:", 0) == 0;
    const std::string proof_path = verify_only ? std::string(argv[1]).substr(7) : std::string();
    const std::string prog = verify_only ? argv[2] : argv[1];
    if (verify_only) { argv[2] = argv[3]; argc = 4; }
    const size_t k = std::stoul(argv[2]), l = k - params::sample_size, n = 4 * k;
    const unsigned seed_byte = argc > 4 ? std::stoul(argv[4]) : 7;
    std::streambuf *chatter = std::cout.rdbuf();                 // the reference prints statistics on stdout; keep them off the JSON
    std::ofstream devnull("/dev/null");
    if (!std::getenv("REFCTX_VERBOSE")) std::cout.rdbuf(devnull.rdbuf());

    // src/webgpu_prover.cpp:228-245
    auto [omega_k, omega_2k, omega_4k] = field_t::generate_omegas(k, n);
    executor_t executor;
    executor.webgpu_init(1, "");
    executor.ntt_init(l, k, n, field_t::modulus, field_t::barrett_factor, omega_k, omega_2k, omega_4k);
    unsigned char encoding_random_seed[32];
    for (int i = 0; i < 32; i++) encoding_random_seed[i] = (unsigned char)(seed_byte * 31 + i * 7 + 1);
    params::hasher::digest instance_hash;                         // no public arguments: hash of nothing, as a fixed 32-byte string here
    for (size_t i = 0; i < params::hasher::digest_size; i++) instance_hash.data[i] = (unsigned char)(0xA0 + i);

    if (verify_only) {
        // the reference's VERIFIER on somebody else's proof of the same program: accept = exit code 0
        proof_fields pf = read_proof(proof_path, instance_hash);
        verdict v{};
        try {
            v = verify(prog, executor, l, k, n, instance_hash, pf);
        } catch (const std::exception &e) {                       // e.g. openings that do not belong to the re-derived sample positions
            std::cout.rdbuf(chatter);
            std::printf("verifier on %s: rejected (%s)\n", kExecutorName, e.what());
            return 1;
        }
        std::cout.rdbuf(chatter);
        std::printf("verifier on %s: merkle %d code %d linear %d quad %d code_eq %d linear_eq %d quad_eq %d\n", kExecutorName, v.merkle, v.code, v.linear,
                    v.quad, v.code_eq, v.linear_eq, v.quad_eq);
        return v.all() ? 0 : 1;
    }

    run_log log1, log2, log3;
    log1.l = log2.l = log3.l = l;

    // stage 1 (src/webgpu_prover.cpp:249-282)
    zkp::merkle_tree<params::hasher> tree;
    std::vector<params::hasher::digest> digests;
    {
        auto ctx = std::make_unique<stage1_t>(&log1, false, executor);
        ctx->init_encoding_random(encoding_random_seed, params::any_iv);
        run_named(prog, *ctx, &log1, l, k);
        digests = ctx->flush_digests();
        tree = digests;
        executor.device_synchronize();
    }
    params::hasher::digest stage1_root = tree.root();
    auto stage1_seed = zkp::hash<params::hasher>("LigetronStage1", stage1_root, instance_hash);

    // stage 2 (:284-341)
    unsigned char seed[params::hasher::digest_size];
    std::copy(stage1_seed.begin(), stage1_seed.end(), seed);
    std::vector<uint32_t> code_limbs, linear_limbs, quad_limbs, dec_code, dec_linear, dec_quad;
    mpz_class linear_sum;
    {
        auto ctx2 = std::make_unique<stage2_t>(&log2, true, executor);
        ctx2->init_encoding_random(encoding_random_seed, params::any_iv);
        ctx2->init_witness_random(seed, params::any_iv);
        run_named(prog, *ctx2, &log2, l, k);
        linear_sum = ctx2->linear_sums();
        buffer_t code_poly = ctx2->code(), linear_poly = ctx2->linear(), quad_poly = ctx2->quadratic();
        code_limbs = executor.template copy_to_host<uint32_t>(code_poly);
        linear_limbs = executor.template copy_to_host<uint32_t>(linear_poly);
        quad_limbs = executor.template copy_to_host<uint32_t>(quad_poly);
        // the prover's self-check (:355-388,465-471)
        executor.decode_ntt_device(executor.bind_ntt(code_poly));
        executor.decode_ntt_device(executor.bind_ntt(linear_poly));
        executor.decode_ntt_device(executor.bind_ntt(quad_poly));
        dec_code = executor.template copy_to_host<uint32_t>(code_poly);
        dec_linear = executor.template copy_to_host<uint32_t>(linear_poly);
        dec_quad = executor.template copy_to_host<uint32_t>(quad_poly);
        executor.device_synchronize();
    }
    auto stage2_seed = zkp::hash<params::hasher>("LigetronStage2", stage1_root, code_limbs, linear_limbs, quad_limbs);
    mpz_vector host_code, host_linear, host_quad;
    host_code.import_limbs(dec_code.data(), dec_code.size(), sizeof(uint32_t), field_t::num_u32_limbs);
    host_linear.import_limbs(dec_linear.data(), dec_linear.size(), sizeof(uint32_t), field_t::num_u32_limbs);
    host_linear.resize(l);
    host_quad.import_limbs(dec_quad.data(), dec_quad.size(), sizeof(uint32_t), field_t::num_u32_limbs);
    host_quad.resize(l);
    const bool valid_code = std::all_of(host_code.begin() + k, host_code.end(), [](const auto &x) { return x == 0; });
    const bool valid_linear = zkp::validate_sum<field_t>(host_linear, linear_sum);
    const bool valid_quad = zkp::validate(host_quad);

    // sampling (:343-353): the reference's hash engine would feed boost's uniform_int_distribution; Boost is absent, the
    // repo's restatement of it stands in (its parity is unpinned -- DESIGN.md section 2)
    cuda::host::digest s2;
    std::memcpy(s2.data, stage2_seed.data, 32);
    std::vector<uint64_t> sample64 = cuda::host::sample_indices(s2, n, params::sample_size);
    std::vector<size_t> sample_index(sample64.begin(), sample64.end());
    auto decommit = tree.decommit(sample_index);

    // stage 3 (:393-407)
    std::vector<uint32_t> samplings;
    {
        auto ctx3 = std::make_unique<stage3_t>(&log3, false, executor, sample_index);
        ctx3->init_encoding_random(encoding_random_seed, params::any_iv);
        run_named(prog, *ctx3, &log3, l, k);
        samplings = ctx3->host_samplings();
    }
    if (log1.kinds != log2.kinds || log1.kinds != log3.kinds || log1.rows != log3.rows || log1.batch_args != log2.batch_args)
        throw std::runtime_error("the three passes did not see the same rows");

    // and the reference's verifier on the proof the three passes made
    proof_fields pf;
    pf.root = stage1_root; pf.code = code_limbs; pf.linear = linear_limbs; pf.quad = quad_limbs; pf.samplings = samplings;
    pf.total_count = decommit.size();
    pf.siblings.assign(decommit.nodes().begin(), decommit.nodes().end());
    const verdict vd = verify(prog, executor, l, k, n, instance_hash, pf);

    std::cout.rdbuf(chatter);
    uint32_t const_sum[8] = {0};
    mpz_export(const_sum, nullptr, -1, sizeof(uint32_t), 0, 0, linear_sum.get_mpz_t());
    std::vector<std::pair<size_t, params::hasher::digest>> nodes(decommit.nodes().begin(), decommit.nodes().end());
    std::sort(nodes.begin(), nodes.end(), [](const auto &a, const auto &b) { return a.first < b.first; });

    std::ofstream out(argv[3]);
    out << "{\n";
    out << "\"program\": \"" << (prog.rfind("ops:", 0) == 0 ? std::string("ops") : prog) << "\", \"executor\": \"" << kExecutorName << "\", \"l\": " << l << ", \"k\": " << k << ", \"n\": " << n << ",\n";
    out << "\"encoding_seed\": \"" << hex(encoding_random_seed, 32) << "\", \"instance_hash\": \"" << hex(instance_hash.data, 32) << "\",\n";
    out << "\"kinds\": [";
    for (size_t i = 0; i < log1.kinds.size(); i++) out << (i ? "," : "") << log1.kinds[i];
    out << "],\n";
    out << "\"values\": \"" << hexv(log1.rows) << "\",\n";
    out << "\"coefs\": \"" << hexv(log2.rows) << "\",\n";
    out << "\"batch_args\": \"" << hexv(log1.batch_args) << "\", \"batch_consts\": \"" << hexv(log1.batch_consts) << "\",\n";
    out << "\"const_sum\": \"" << hex(const_sum, 32) << "\",\n";
    out << "\"digests\": \"" << hex(digests.data(), digests.size() * sizeof(digests[0])) << "\",\n";
    out << "\"root\": \"" << hex(stage1_root.data, 32) << "\", \"stage1_seed\": \"" << hex(stage1_seed.data, 32) << "\", \"stage2_seed\": \"" << hex(stage2_seed.data, 32) << "\",\n";
    out << "\"code\": \"" << hexv(code_limbs) << "\",\n\"linear\": \"" << hexv(linear_limbs) << "\",\n\"quad\": \"" << hexv(quad_limbs) << "\",\n";
    out << "\"valid\": [" << valid_code << "," << valid_linear << "," << valid_quad << "],\n";
    out << "\"verifier\": [" << vd.merkle << "," << vd.code << "," << vd.linear << "," << vd.quad << "," << vd.code_eq << "," << vd.linear_eq << "," << vd.quad_eq << "],\n";
    out << "\"sample_index\": [";
    for (size_t i = 0; i < sample_index.size(); i++) out << (i ? "," : "") << sample_index[i];
    out << "],\n\"decommit_total\": " << decommit.size() << ", \"decommit_nodes\": {";
    for (size_t i = 0; i < nodes.size(); i++) out << (i ? "," : "") << "\"" << nodes[i].first << "\": \"" << hex(nodes[i].second.data, 32) << "\"";
    out << "},\n\"samplings\": \"" << hexv(samplings) << "\"\n}\n";
    out.close();
    std::printf("%s k=%zu on %s: %zu events, root %s, valid %d%d%d\n", prog.c_str(), k, kExecutorName, log1.kinds.size(), hex(stage1_root.data, 32).c_str(),
                (int)valid_code, (int)valid_linear, (int)valid_quad);
    return (valid_code && valid_linear && valid_quad && vd.all()) ? 0 : 1;
}

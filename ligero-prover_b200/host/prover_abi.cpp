// C ABI (include/lgr_prover.h) over the header-only host layer: fiat_shamir.hpp, merkle_host.hpp,
// proof_wire.hpp, matrix_prover.hpp.  Built as liblgr_prover.so, linked against liblgr.so, libcrypto, libz.
#include <exception>
#include <string>

#include "../../include/lgr_prover.h"
#include "matrix_prover.hpp"
#include "row_packer.hpp"
#include "wat_emitter.hpp"

using namespace ligero::cuda::host;

static thread_local std::string g_perr;

struct lgrp_proof {
    prove_result r;
};

#define LGRP_TRY try {
#define LGRP_END                                                        \
    }                                                                   \
    catch (const std::exception &e) { g_perr = e.what(); return 1; }    \
    catch (...) { g_perr = "unknown error"; return 1; }                 \
    return 0;

static digest dg(const uint8_t *p) { digest d; memcpy(d.data, p, 32); return d; }

extern "C" {

const char *lgrp_last_error(void) { return g_perr.c_str(); }

int lgrp_stage1_seed(const uint8_t root[32], const uint8_t instance_hash[32], uint8_t out[32]) {
    LGRP_TRY
    const digest d = stage1_seed(dg(root), dg(instance_hash));
    memcpy(out, d.data, 32);
    LGRP_END
}

int lgrp_stage2_seed(const uint8_t root[32], const uint32_t *code, const uint32_t *linear, const uint32_t *quad, size_t nwords, uint8_t out[32]) {
    LGRP_TRY
    const std::vector<uint32_t> c(code, code + nwords), l(linear, linear + nwords), q(quad, quad + nwords);
    const digest d = stage2_seed(dg(root), c, l, q);
    memcpy(out, d.data, 32);
    LGRP_END
}

int lgrp_hash_random_bytes(const uint8_t seed[32], uint8_t *out, size_t count) {
    LGRP_TRY
    hash_random_engine eng(dg(seed));
    for (size_t i = 0; i < count; i++) out[i] = eng();
    LGRP_END
}

int lgrp_sample_indices(const uint8_t seed[32], uint64_t n, uint64_t sample_size, uint64_t *out, uint64_t *out_count) {
    LGRP_TRY
    const std::vector<uint64_t> s = sample_indices(dg(seed), n, sample_size);
    for (size_t i = 0; i < s.size(); i++) out[i] = s[i];
    if (out_count) *out_count = s.size();
    LGRP_END
}

int lgrp_fr_random(const uint8_t key[32], const uint8_t iv[16], size_t count, uint32_t *out_limbs) {
    LGRP_TRY
    fr_random_stream s(key, iv);
    for (size_t i = 0; i < count; i++) s.next(out_limbs + 8 * i);
    LGRP_END
}

int lgrp_decommit(const uint8_t *nodes, uint64_t total_count, const uint64_t *leaf_idx, uint64_t nidx, uint64_t *positions_out,
                  uint8_t *siblings_out, uint64_t *count_out) {
    LGRP_TRY
    const std::vector<uint64_t> idx(leaf_idx, leaf_idx + nidx);
    const decommitment d = decommit(nodes, total_count, idx);
    for (size_t i = 0; i < d.positions.size(); i++) {
        if (positions_out) positions_out[i] = d.positions[i];
        if (siblings_out) memcpy(siblings_out + 32 * i, d.siblings[i].data, 32);
    }
    if (count_out) *count_out = d.positions.size();
    LGRP_END
}

int lgrp_recommit(const uint8_t *leaves, const uint64_t *leaf_idx, uint64_t nidx, uint64_t total_count, const uint8_t *siblings, uint64_t nsib,
                  uint8_t root_out[32]) {
    LGRP_TRY
    decommitment d;
    d.total_count = total_count;
    d.known_index.assign(leaf_idx, leaf_idx + nidx);
    d.positions = sibling_positions(d.known_index, total_count);
    if (d.positions.size() != nsib) throw std::runtime_error("Sibling hash count mismatch: expected " + std::to_string(d.positions.size()) + ", got " + std::to_string(nsib));
    d.siblings.resize(nsib);
    for (uint64_t i = 0; i < nsib; i++) memcpy(d.siblings[i].data, siblings + 32 * i, 32);
    std::vector<digest> lv(nidx);
    for (uint64_t i = 0; i < nidx; i++) memcpy(lv[i].data, leaves + 32 * i, 32);
    const digest r = recommit(lv, d);
    memcpy(root_out, r.data, 32);
    LGRP_END
}

int lgrp_proof_bytes(const lgrp_proof *p, int which, const uint8_t **data, size_t *len) {
    LGRP_TRY
    if (!p || !data || !len) throw std::invalid_argument("null argument");
    const std::string &s = which == 0 ? p->r.envelope : p->r.gzip;
    *data = reinterpret_cast<const uint8_t *>(s.data());
    *len = s.size();
    LGRP_END
}

int lgrp_proof_parse(const uint8_t *data, size_t len, lgrp_proof **out) {
    LGRP_TRY
    if (!data || !out) throw std::invalid_argument("null argument");
    std::string in(reinterpret_cast<const char *>(data), len);
    if (len >= 2 && data[0] == 0x1f && data[1] == 0x8b) in = gzip_decompress(in);
    lgrp_proof *p = new lgrp_proof();
    try {
        p->r.proof = deserialize_proof(in);
        p->r.envelope = serialize_proof(p->r.proof);
        p->r.gzip = gzip_compress(p->r.envelope, 6);
    } catch (...) { delete p; throw; }
    *out = p;
    LGRP_END
}

void lgrp_proof_free(lgrp_proof *p) { delete p; }

int lgrp_proof_info(const lgrp_proof *p, uint32_t *valid_bits, uint8_t s1[32], uint8_t s2[32], uint64_t *encoded_rows) {
    LGRP_TRY
    if (!p) throw std::invalid_argument("null argument");
    if (valid_bits) *valid_bits = (p->r.valid_code ? 1u : 0u) | (p->r.valid_linear ? 2u : 0u) | (p->r.valid_quad ? 4u : 0u);
    if (s1) memcpy(s1, p->r.stage1_seed.data, 32);
    if (s2) memcpy(s2, p->r.stage2_seed.data, 32);
    if (encoded_rows) *encoded_rows = p->r.encoded_rows;
    LGRP_END
}

int lgrp_proof_timing(const lgrp_proof *p, double ms[4]) {
    LGRP_TRY
    if (!p || !ms) throw std::invalid_argument("null argument");
    for (int i = 0; i < 4; i++) ms[i] = p->r.ms[i];
    LGRP_END
}

struct lgrp_packer {
    row_packer p;
    explicit lgrp_packer(uint32_t l) : p(l) {}
};

int lgrp_packer_create(uint32_t l, lgrp_packer **out) {
    LGRP_TRY
    if (!out || !l) throw std::invalid_argument("row size must be positive");
    *out = new lgrp_packer(l);
    LGRP_END
}
void lgrp_packer_free(lgrp_packer *p) { delete p; }
int lgrp_packer_push_linear(lgrp_packer *p, const uint32_t value[8], const uint32_t coef[8]) {
    LGRP_TRY
    if (!p || !value || !coef) throw std::invalid_argument("null argument");
    p->p.push_linear(value, coef);
    LGRP_END
}
int lgrp_packer_push_quadratic(lgrp_packer *p, const uint32_t xyz[24], const uint32_t coef_xyz[24]) {
    LGRP_TRY
    if (!p || !xyz || !coef_xyz) throw std::invalid_argument("null argument");
    p->p.push_quadratic(xyz, xyz + 8, xyz + 16, coef_xyz, coef_xyz + 8, coef_xyz + 16);
    LGRP_END
}
int lgrp_packer_finalize(lgrp_packer *p) {
    LGRP_TRY
    if (!p) throw std::invalid_argument("null argument");
    p->p.finalize();
    LGRP_END
}
int lgrp_packer_rows(const lgrp_packer *p, uint64_t *n_events, const uint8_t **kinds, uint64_t *n_rows, const uint32_t **values, const uint32_t **coefs) {
    LGRP_TRY
    if (!p) throw std::invalid_argument("null argument");
    if (n_events) *n_events = p->p.kinds().size();
    if (kinds) *kinds = p->p.kinds().data();
    if (n_rows) *n_rows = p->p.rows();
    if (values) *values = p->p.values().data();
    if (coefs) *coefs = p->p.coefs().data();
    LGRP_END
}

int lgrp_prove(lgr_ctx *ctx, const lgrp_statement *s, lgrp_proof **out) {
    LGRP_TRY
    if (!ctx || !s || !out) throw std::invalid_argument("null argument");
    if (s->n_events && (!s->kinds || !s->values)) throw std::invalid_argument("statement without rows");
    statement st;
    st.l = s->l; st.k = s->k;
    memcpy(st.const_sum, s->const_sum, 32);
    memcpy(st.encoding_seed, s->encoding_seed, 32);
    st.instance_hash = dg(s->instance_hash);
    st.program_hash = dg(s->program_hash);
    st.generated_at_seconds = s->generated_at_seconds;
    st.sample_size = s->sample_size ? s->sample_size : 192;
    size_t row = 0, nb = 0, nc = 0;
    const size_t stride = (size_t)s->l * 8;
    st.arena_slots = s->arena_slots;
    st.events.resize(s->n_events);
    for (uint64_t e = 0; e < s->n_events; e++) {
        row_event &ev = st.events[e];
        ev.kind = s->kinds[e];
        if (ev.kind >= EV_KIND_COUNT) throw std::invalid_argument("unknown event kind");
        for (int j = 0; j < event_host_rows(ev.kind); j++, row++) {
            ev.val[j] = s->values + row * stride;
            ev.coef[j] = (s->coefs && ev.kind != EV_VSET) ? s->coefs + row * stride : nullptr;
        }
        if (ev.kind >= EV_VSET) {
            if (!s->batch_args) throw std::invalid_argument("vbn254fr events without batch_args");
            for (int j = 0; j < 3; j++) ev.arg[j] = s->batch_args[3 * nb + j];
            nb++;
            if (event_takes_constant(ev.kind)) {
                if (!s->batch_consts) throw std::invalid_argument("constant-taking vbn254fr event without batch_consts");
                ev.konst = s->batch_consts + 8 * nc++;
            }
        }
    }
    lgrp_proof *p = new lgrp_proof();
    try {
        matrix_prover mp(ctx);
        p->r = mp.prove(st);
    } catch (...) { delete p; throw; }
    *out = p;
    LGRP_END
}

static void fill_stats(lgrp_wat_stats *o, const wat_stats &s) {
    if (!o) return;
    o->private_consts = s.private_consts; o->asserts = s.asserts; o->arithmetic_ops = s.arithmetic_ops;
    o->linear_witnesses = s.linear_witnesses; o->quadratic_slots = s.quadratic_slots; o->linear_constraints = s.linear_constraints;
    o->violated_constraints = s.violated_constraints;
}

// the guest's arguments: argv[0] included, private ones by index
static void take_args(wat_program &prog, const lgrp_wat_args *a) {
    if (!a) return;
    if (a->nargs && (!a->args || !a->arg_lens)) throw std::invalid_argument("null argument list");
    if (a->nprivate && !a->private_indices) throw std::invalid_argument("null private index list");
    std::vector<std::vector<uint8_t>> args;
    for (uint32_t i = 0; i < a->nargs; i++) {
        if (a->arg_lens[i] && !a->args[i]) throw std::invalid_argument("null argument");
        args.emplace_back(a->args[i], a->args[i] + a->arg_lens[i]);
    }
    std::set<int> priv;
    for (uint32_t i = 0; i < a->nprivate; i++) priv.insert(a->private_indices[i]);
    prog.set_args(std::move(args), std::move(priv));
}
// instance_hash: the public arguments folded into a zero digest, one hash(instance_hash, argument) each (src/webgpu_prover.cpp:160-168)
static digest instance_hash_of(const lgrp_wat_args *a) {
    digest d;
    if (!a) return d;
    std::set<int> priv;
    for (uint32_t i = 0; i < a->nprivate; i++) priv.insert(a->private_indices[i]);
    for (uint32_t i = 0; i < a->nargs; i++) {
        if (priv.count((int)i)) continue;
        sha256 h;
        h << d;
        h.update(a->args[i], a->arg_lens[i]);
        d = h.flush_digest();
    }
    return d;
}

int lgrp_wat_instance_hash(const lgrp_wat_args *args, uint8_t out[32]) {
    LGRP_TRY
    if (!out) throw std::invalid_argument("null argument");
    const digest d = instance_hash_of(args);
    memcpy(out, d.data, 32);
    LGRP_END
}

int lgrp_wat_emit(const char *wat, size_t len, uint32_t l, const uint8_t *stage1_seed, lgrp_packer **rows_out, uint32_t const_sum[8],
                  lgrp_wat_stats *stats) {
    return lgrp_wat_emit_args(wat, len, nullptr, l, stage1_seed, rows_out, const_sum, stats, nullptr);
}

int lgrp_wat_emit_args(const char *wat, size_t len, const lgrp_wat_args *args, uint32_t l, const uint8_t *stage1_seed, lgrp_packer **rows_out,
                       uint32_t const_sum[8], lgrp_wat_stats *stats, int32_t *exit_code) {
    LGRP_TRY
    if (!wat || !rows_out || !l) throw std::invalid_argument("null argument");
    wat_program prog(std::string(wat, len));
    take_args(prog, args);
    wat_stats ws;
    lgrp_packer *pk = new lgrp_packer(l);
    try {
        witness_machine m(pk->p, stage1_seed);
        prog.run(m, ws);
        m.finish(const_sum);
    } catch (...) { delete pk; throw; }
    fill_stats(stats, ws);
    if (exit_code) *exit_code = prog.exit_code();
    *rows_out = pk;
    LGRP_END
}

int lgrp_prove_wat(lgr_ctx *ctx, const char *wat, size_t len, const uint8_t encoding_seed[32], int64_t generated_at_seconds,
                   lgrp_proof **out, lgrp_wat_stats *stats) {
    return lgrp_prove_wat_args(ctx, wat, len, nullptr, encoding_seed, generated_at_seconds, out, stats);
}

int lgrp_prove_wat_args(lgr_ctx *ctx, const char *wat, size_t len, const lgrp_wat_args *args, const uint8_t encoding_seed[32],
                        int64_t generated_at_seconds, lgrp_proof **out, lgrp_wat_stats *stats) {
    LGRP_TRY
    if (!ctx || !wat || !encoding_seed || !out) throw std::invalid_argument("null argument");
    uint32_t l = 0, k = 0, n = 0;
    if (lgr_geometry(ctx, &l, &k, &n)) throw std::runtime_error(lgr_last_error());
    wat_program prog(std::string(wat, len));
    take_args(prog, args);
    wat_stats ws;
    row_packer values(l, true, false);
    {
        witness_machine m(values, nullptr);                  // stage 1 needs the values only
        prog.set_echo(true);                                 // what the guest prints appears once, not once per stage
        prog.run(m, ws);
        prog.set_echo(false);
        m.finish(nullptr);
    }
    statement st;
    st.l = l; st.k = k;
    st.instance_hash = instance_hash_of(args);
    memcpy(st.encoding_seed, encoding_seed, 32);
    st.generated_at_seconds = generated_at_seconds;
    {
        sha256 h; h.update(wat, len); st.program_hash = h.flush_digest();
    }
    size_t row = 0;
    const size_t stride = (size_t)l * 8;
    for (uint8_t kind : values.kinds()) {
        row_event ev;
        ev.kind = kind ? EV_QUAD : EV_LINEAR;
        for (int j = 0; j < (kind ? 3 : 1); j++, row++) ev.val[j] = values.values().data() + row * stride;
        st.events.push_back(ev);
    }
    st.coef_provider = [&](const uint8_t seed[32], std::vector<uint32_t> &coef_rows, uint32_t const_sum[8]) {
        row_packer with_coefs(l, false, true);               // the program again, now drawing the linear-test randomness
        with_coefs.reserve_rows(values.rows());
        witness_machine m(with_coefs, seed);
        wat_stats again;
        prog.run(m, again);
        m.finish(const_sum);
        coef_rows = with_coefs.take_coefs();
    };
    lgrp_proof *p = new lgrp_proof();
    try {
        matrix_prover mp(ctx);
        p->r = mp.prove(st);
    } catch (...) { delete p; throw; }
    fill_stats(stats, ws);
    *out = p;
    LGRP_END
}

}  // extern "C"

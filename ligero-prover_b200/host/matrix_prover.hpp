// Three-stage Ligero prover over a witness MATRIX: what src/webgpu_prover.cpp:249-471
// does, with the WASM interpreter + witness_manager (out of scope) replaced by an explicit
// list of rows.  Everything device-side goes through the C ABI (include/lgr.h); everything host-side
// (transcript, sampling, openings, container) is fiat_shamir.hpp / merkle_host.hpp / proof_wire.hpp.
//
// Reference flow reproduced here (file:line in /root/reference):
//   rows + pads      include/zkp/backend/witness_manager.hpp:200-269 (zero-fill to l, k-l pads from the
//                    encoding AES-CTR stream; coefficient rows zero on the pads)
//   masks            witness_manager.hpp:271-321 (code mask on the w_k domain, linear / quadratic masks on
//                    the w_2k domain), emitted last (witness_manager.hpp:497-503)
//   stage 1          include/zkp/nonbatch_context.hpp:445-494,555-558  encode + column hash, tree
//   stage 2          nonbatch_context.hpp:654-780   code += r*enc(row), linear += enc(row)*enc(coef),
//                    quad += r*(X*Y - Z); masks added; r from AES-CTR(stage1_seed)
//   sampling         src/webgpu_prover.cpp:337-353
//   stage 3          nonbatch_context.hpp:935-993   sampled columns of every encoded row, masks included
//   self-check       src/webgpu_prover.cpp:355-388,465-471
// Unlike the reference (three interpreter runs, one row per callback) the rows live in HBM once and every
// stage sweeps them in tiles: encode T rows -> hash / combine / gather the resident tile.
#pragma once
#include <chrono>
#include <functional>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/lgr.h"
#include "proof_wire.hpp"

namespace ligero::cuda::host {

using limbs = uint32_t[8];

// One event, in emission order.
//   EV_LINEAR / EV_QUAD: a linear row (1 encoded row) or a quadratic triple (3 encoded rows x, y, z) from the scalar
//     backend.  Pointers are to l canonical elements (8 x u32 each); coef = the row of linear-test coefficients the
//     backend accumulated for these witnesses (witness_manager.hpp:141-171).
//   EV_V*: a vbn254fr host call (include/host_modules/vbn254fr.hpp:139-566) on DEVICE-RESIDENT variables of k elements
//     each (`arg` = arena slot indices), with the on_batch_* callback it triggers
//     (nonbatch_context.hpp:497-553 stage 1, :782-847 stage 2, :996-1047 stage 3):
//       VSET   out            write_buffer_clear(out, val[0]); on_batch_init(out): 192 pads from the encoding stream are
//                             written INTO the variable at [l, k), then 1 row is committed       [code: 1 draw]
//       VCOPY  out, in        out = in; on_batch_equal(out, in): 2 rows                          [quad += r (x - y)]
//       VASSERT_EQ x, y       on_batch_equal(x, y): 2 rows                                       [quad += r (x - y)]
//       VMUL   out, x, y      tmp = x*y; on_batch_quadratic(x, y, tmp): 3 rows; out = tmp        [code: 3 draws, quad += r (xy - z)]
//       VDIV   out, x, y      tmp = x/y; on_batch_quadratic(tmp, y, x): 3 rows; out = tmp
//       VBIT   out, x, bit    tmp = bit `bit` of x; out = tmp; on_batch_bit(out): 1 row          [code: 1 draw, quad += r (x*x - x)]
//       VADD / VSUB out, x, y and VADDC / VSUBC / VCSUB / VMULC / VMONTMULC out, x, konst: arithmetic only, no row
//     All arithmetic runs over the whole k elements, pads included, exactly as the reference's arena kernels do.
enum event_kind : uint8_t {
    EV_LINEAR = 0, EV_QUAD = 1, EV_VSET = 2, EV_VCOPY = 3, EV_VADD = 4, EV_VSUB = 5, EV_VMUL = 6, EV_VDIV = 7, EV_VASSERT_EQ = 8,
    EV_VBIT = 9, EV_VADDC = 10, EV_VSUBC = 11, EV_VCSUB = 12, EV_VMULC = 13, EV_VMONTMULC = 14, EV_KIND_COUNT = 15
};
inline int event_rows(uint8_t kind) {            // encoded (committed) rows of an event
    switch (kind) {
        case EV_LINEAR: case EV_VSET: case EV_VBIT: return 1;
        case EV_VCOPY: case EV_VASSERT_EQ: return 2;
        case EV_QUAD: case EV_VMUL: case EV_VDIV: return 3;
        default: return 0;
    }
}
inline int event_host_rows(uint8_t kind) {       // rows of `values` / `coefs` the event consumes
    return kind == EV_LINEAR || kind == EV_VSET ? 1 : (kind == EV_QUAD ? 3 : 0);
}
inline bool event_takes_constant(uint8_t kind) { return kind >= EV_VADDC && kind <= EV_VMONTMULC; }
struct row_event {
    uint8_t kind = EV_LINEAR;
    const uint32_t *val[3] = {nullptr, nullptr, nullptr};
    const uint32_t *coef[3] = {nullptr, nullptr, nullptr};
    uint32_t arg[3] = {0, 0, 0};                 // EV_V*: arena slots (out, x, y) / bit index
    const uint32_t *konst = nullptr;             // EV_V*C: canonical constant
};

struct statement {
    uint32_t l = 0, k = 0;                       // n = 4k
    std::vector<row_event> events;
    uint32_t const_sum[8] = {0};                 // linear test: sum_{i<l} P_linear[i] + const_sum == 0 (zkp/common.hpp:68-79)
    uint8_t encoding_seed[32] = {0};             // the reference draws it from std::random_device (src/webgpu_prover.cpp:239-245)
    digest instance_hash, program_hash;
    int64_t generated_at_seconds = -1;           // < 0: wall clock
    uint32_t sample_size = 192;                  // params::sample_size
    uint32_t arena_slots = 0;                    // vbn254fr variables (k elements each, zero on allocation)
    // The reference derives the linear-test coefficients AFTER the commitment: stage 2 re-runs the program with the linear
    // random engine keyed by the stage-1 seed (nonbatch_context.hpp:105-112, src/webgpu_prover.cpp:281-335).  A statement
    // whose coefficients depend on that seed supplies this callback instead of `coef` pointers: it receives the seed and
    // fills one coefficient row per host row ([host rows][l][8 x u32], event order) and const_sum.
    std::function<void(const uint8_t stage1_seed[32], std::vector<uint32_t> &coef_rows, uint32_t const_sum[8])> coef_provider;
};

struct prove_result {
    proof_data proof;
    std::string envelope;                        // serialized LigeroProofEnvelope
    std::string gzip;                            // what the reference writes to proof_data.gz
    digest stage1_seed, stage2_seed;
    bool valid_code = false, valid_linear = false, valid_quad = false;   // the prover's self-check (src/webgpu_prover.cpp:465-471)
    bool ok() const { return valid_code && valid_linear && valid_quad; }
    uint64_t encoded_rows = 0;                   // incl. the three mask rows
    double ms[4] = {0, 0, 0, 0};                 // wall time of stage 1, stage 2 (+ sampling, self-check), stage 3, container
};

class matrix_prover {
public:
    explicit matrix_prover(lgr_ctx *ctx) : ctx_(ctx) {
        if (lgr_geometry(ctx_, &l_, &k_, &n_)) fail("lgr_geometry");
    }
    ~matrix_prover() { release(); }
    matrix_prover(const matrix_prover &) = delete;
    matrix_prover &operator=(const matrix_prover &) = delete;

    prove_result prove(const statement &st) {
        if (st.l != l_ || st.k != k_) throw std::invalid_argument("statement geometry differs from the context's");
        release();
        prove_result out;
        build_rows(st);
        out.encoded_rows = rows_ + 3;

        // ---- stage 1: commit -----------------------------------------------------------------------
        const size_t sha_bytes = lgr_sha_ctx_bytes(n_);
        void *sha = dalloc(sha_bytes), *digests = dalloc((size_t)n_ * 32);
        const size_t node_count = lgr_merkle_node_count(n_);
        void *nodes = dalloc(node_count * 32);
        const auto t0 = clock_now();
        chk(lgr_sha_init(ctx_, sha, n_));
        chk(lgr_encode_absorb(ctx_, sha, d_val_, rows_));      // tile pipeline: encode of tile t+1 overlaps hashing of tile t
        for (int m = 0; m < 3; m++) { encode_mask(m); chk(lgr_sha_update(ctx_, sha, n_, mask_cw_)); }
        chk(lgr_sha_final(ctx_, sha, n_, digests));
        chk(lgr_merkle_build(ctx_, digests, n_, nodes));
        std::vector<uint8_t> host_nodes(node_count * 32);
        chk(lgr_read(ctx_, host_nodes.data(), nodes, 0, host_nodes.size()));
        digest root;
        memcpy(root.data, host_nodes.data(), 32);
        out.stage1_seed = stage1_seed(root, st.instance_hash);
        const auto t1 = clock_now();

        uint32_t const_sum[8];
        memcpy(const_sum, st.const_sum, 32);
        if (st.coef_provider) {                                                   // coefficients that depend on the commitment
            std::vector<uint32_t> crow(host_row_of_.size() * (size_t)l_ * 8, 0);
            st.coef_provider(out.stage1_seed.data, crow, const_sum);
            if (crow.size() != host_row_of_.size() * (size_t)l_ * 8) throw std::invalid_argument("coef_provider returned a wrong number of rows");
            for (size_t h = 0; h < host_row_of_.size(); h++)
                chk(lgr_write(ctx_, at(d_coef_, host_row_of_[h] * k_), 0, &crow[h * (size_t)l_ * 8], (size_t)l_ * 32));
        }
        // ---- stage 2: test vectors -----------------------------------------------------------------
        static const uint8_t any_iv[16] = {0};                                 // params::any_iv
        fr_random_stream code_rng(out.stage1_seed.data, any_iv), quad_rng(out.stage1_seed.data, any_iv);   // nonbatch_context.hpp:105-112
        // draw order = callback order (nonbatch_context.hpp:756-780,782-847): per event, first the code draws (one per
        // row that goes through check_code; on_batch_equal rows take none), then one quadratic draw for events that feed
        // the quadratic test.  Rows without a code draw combine with r = 0.
        std::vector<uint32_t> r_code(std::max<size_t>(rows_, 1) * 8, 0);
        std::vector<quad_item> quads;
        {
            size_t r = 0;
            for (const row_event &e : st.events) {
                const int nr = event_rows(e.kind);
                const bool code_draw = e.kind != EV_VCOPY && e.kind != EV_VASSERT_EQ;
                if (code_draw) for (int j = 0; j < nr; j++) code_rng.next(&r_code[(r + j) * 8]);
                if (e.kind == EV_QUAD || e.kind == EV_VMUL || e.kind == EV_VDIV || e.kind == EV_VBIT || e.kind == EV_VCOPY || e.kind == EV_VASSERT_EQ) {
                    quad_item q;
                    q.row = r;
                    q.shape = (e.kind == EV_VBIT) ? 1 : ((e.kind == EV_VCOPY || e.kind == EV_VASSERT_EQ) ? 2 : 0);
                    quad_rng.next(q.r);
                    quads.push_back(q);
                }
                r += nr;
            }
        }
        void *code = dalloc((size_t)n_ * 32), *linear = dalloc((size_t)n_ * 32), *quad = dalloc((size_t)n_ * 32);
        size_t quad_seen = 0;
        for_each_tile([&](size_t r0, uint32_t T) {
            chk(lgr_encode_rows(ctx_, at(d_val_, r0 * k_), k_, T, tile_));
            chk(lgr_encode_rows(ctx_, at(d_coef_, r0 * k_), k_, T, tile2_));
            chk(lgr_combine_code(ctx_, tile_, T, &r_code[r0 * 8], code));
            chk(lgr_combine_linear(ctx_, tile_, tile2_, T, linear));
            // the quadratic-test items of the tile, wherever they sit between other rows:
            //   triples: x, y, z are 3 consecutive codewords        quad += r (x*y - z)
            //   bits   : one codeword                               quad += r (x*x - x)      (on_batch_bit copies x into y and z)
            //   equals : two consecutive codewords                  quad += r x + (p - r) y  (EltwiseSubMod + EltwiseFMAMod)
            std::vector<uint32_t> trip_rows, trip_r, bit_rows, bit_r, eq_scal;
            for (; quad_seen < quads.size() && quads[quad_seen].row < r0 + T; quad_seen++) {
                const quad_item &q = quads[quad_seen];
                const uint32_t rel = (uint32_t)(q.row - r0);
                if (q.shape == 0) { trip_rows.push_back(rel); trip_r.insert(trip_r.end(), q.r, q.r + 8); }
                else if (q.shape == 1) { bit_rows.push_back(rel); bit_r.insert(bit_r.end(), q.r, q.r + 8); }
                else {
                    if (eq_scal.empty()) eq_scal.assign((size_t)T * 8, 0);
                    memcpy(&eq_scal[(size_t)rel * 8], q.r, 32);
                    uint64_t rr[4]; memcpy(rr, q.r, 32);
                    negate(&eq_scal[(size_t)(rel + 1) * 8], rr);
                }
            }
            if (!trip_rows.empty()) chk(lgr_combine_quad_indexed(ctx_, tile_, trip_rows.data(), (uint32_t)trip_rows.size(), trip_r.data(), quad));
            if (!bit_rows.empty()) chk(lgr_combine_bit_indexed(ctx_, tile_, bit_rows.data(), (uint32_t)bit_rows.size(), bit_r.data(), quad));
            if (!eq_scal.empty()) chk(lgr_combine_code(ctx_, tile_, T, eq_scal.data(), quad));
        });
        encode_mask(0); chk(lgr_elt_add_assign(ctx_, mask_cw_, code, n_));     // nonbatch_context.hpp:732-754
        encode_mask(1); chk(lgr_elt_add_assign(ctx_, mask_cw_, linear, n_));
        encode_mask(2); chk(lgr_elt_add_assign(ctx_, mask_cw_, quad, n_));
        proof_data &pd = out.proof;
        pd.code.resize((size_t)n_ * 8); pd.linear.resize((size_t)n_ * 8); pd.quad.resize((size_t)n_ * 8);
        chk(lgr_read(ctx_, pd.code.data(), code, 0, (size_t)n_ * 32));
        chk(lgr_read(ctx_, pd.linear.data(), linear, 0, (size_t)n_ * 32));
        chk(lgr_read(ctx_, pd.quad.data(), quad, 0, (size_t)n_ * 32));
        out.stage2_seed = stage2_seed(root, pd.code, pd.linear, pd.quad);
        const std::vector<uint64_t> sample = sample_indices(out.stage2_seed, n_, st.sample_size);
        pd.merkle_root = root;
        pd.decommit = decommit(host_nodes.data(), node_count, sample);

        // the prover's own validation (src/webgpu_prover.cpp:355-388,465-471): decode the three test vectors
        {
            std::vector<uint32_t> h((size_t)n_ * 8);
            chk(lgr_decode(ctx_, code)); chk(lgr_read(ctx_, h.data(), code, 0, h.size() * 4));
            out.valid_code = true;
            for (size_t i = (size_t)k_ * 8; i < h.size(); i++) if (h[i]) { out.valid_code = false; break; }
            chk(lgr_decode(ctx_, linear)); chk(lgr_read(ctx_, h.data(), linear, 0, h.size() * 4));
            out.valid_linear = sum_is_zero(h.data(), l_, const_sum);
            chk(lgr_decode(ctx_, quad)); chk(lgr_read(ctx_, h.data(), quad, 0, h.size() * 4));
            out.valid_quad = true;
            for (size_t i = 0; i < (size_t)l_ * 8; i++) if (h[i]) { out.valid_quad = false; break; }
        }

        const auto t2 = clock_now();
        // ---- stage 3: open the sampled columns -----------------------------------------------------
        const uint32_t S = (uint32_t)sample.size();
        chk(lgr_sample_init(ctx_, sample.data(), S));
        pd.samplings.resize((rows_ + 3) * (size_t)S * 8);
        void *d_samp = dalloc((size_t)tile_rows_ * S * 32);
        for_each_tile([&](size_t r0, uint32_t T) {
            chk(lgr_encode_rows(ctx_, at(d_val_, r0 * k_), k_, T, tile_));
            chk(lgr_sample_gather_rows(ctx_, tile_, n_, T, d_samp));
            chk(lgr_read(ctx_, &pd.samplings[r0 * S * 8], d_samp, 0, (size_t)T * S * 32));
        });
        for (int m = 0; m < 3; m++) {
            encode_mask(m);
            chk(lgr_sample_gather_rows(ctx_, mask_cw_, n_, 1, d_samp));
            chk(lgr_read(ctx_, &pd.samplings[(rows_ + m) * S * 8], d_samp, 0, (size_t)S * 32));
        }

        const auto t3 = clock_now();
        // ---- container (src/webgpu_prover.cpp:410-458) ---------------------------------------------
        pd.meta.program_hash = st.program_hash;
        pd.meta.packing_size = k_;
        pd.meta.codeword_size = n_;
        pd.meta.sample_size = st.sample_size;
        pd.meta.generated_at_seconds = st.generated_at_seconds >= 0
            ? st.generated_at_seconds
            : (int64_t)std::chrono::duration_cast<std::chrono::seconds>(std::chrono::system_clock::now().time_since_epoch()).count();
        out.envelope = serialize_proof(pd);
        out.gzip = gzip_compress(out.envelope, 6);
        const auto t4 = clock_now();
        out.ms[0] = ms_between(t0, t1); out.ms[1] = ms_between(t1, t2); out.ms[2] = ms_between(t2, t3); out.ms[3] = ms_between(t3, t4);
        release();
        return out;
    }

private:
    using clock_point = std::chrono::steady_clock::time_point;
    static clock_point clock_now() { return std::chrono::steady_clock::now(); }
    static double ms_between(clock_point a, clock_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); }
    [[noreturn]] void fail(const char *what) { throw std::runtime_error(std::string(what) + ": " + lgr_last_error()); }
    void chk(int rc) { if (rc) fail("liblgr"); }
    static void *at(void *base, size_t elems) { return static_cast<uint8_t *>(base) + elems * 32; }
    void *dalloc(size_t bytes) { void *p = nullptr; chk(lgr_alloc(ctx_, bytes ? bytes : 32, &p)); owned_.push_back(p); return p; }
    void release() { for (void *p : owned_) lgr_free(ctx_, p); owned_.clear(); d_val_ = d_coef_ = tile_ = tile2_ = mask_cw_ = nullptr; }

    template <typename F> void for_each_tile(F f) {
        for (size_t r0 = 0; r0 < rows_;) {
            size_t T = std::min<size_t>(tile_rows_, rows_ - r0);
            // never split a triple: back off to the last event boundary inside the tile
            while (T > 0 && r0 + T < rows_ && !row_starts_event_[r0 + T]) T--;
            if (T == 0) throw std::logic_error("tile smaller than one event");
            f(r0, (uint32_t)T);
            r0 += T;
        }
    }

    struct quad_item { size_t row; int shape; uint32_t r[8]; };       // shape 0 = triple, 1 = bit, 2 = equal pair

    // rows in emission order -> device, pads from the encoding stream; the three masks
    void build_rows(const statement &st) {
        rows_ = 0;
        bool any_batch = false;
        for (const row_event &e : st.events) {
            if (e.kind >= EV_KIND_COUNT) throw std::invalid_argument("unknown event kind");
            rows_ += event_rows(e.kind);
            any_batch |= e.kind >= EV_VSET;
        }
        static const uint8_t any_iv[16] = {0};
        fr_random_stream enc(st.encoding_seed, any_iv);                       // init_encoding_random(seed, params::any_iv)
        const size_t pad = k_ - l_;
        std::vector<uint32_t> val(std::max<size_t>(rows_, 1) * k_ * 8, 0), coef(std::max<size_t>(rows_, 1) * k_ * 8, 0);
        std::vector<std::vector<uint32_t>> vset_rows;                         // l values + pads of every VSET, in order
        row_is_quad_x_.assign(rows_ + 1, 0);
        row_starts_event_.assign(rows_ + 1, 0);
        host_row_of_.clear();
        size_t r = 0;
        for (const row_event &e : st.events) {
            const int nr = event_rows(e.kind);
            if (nr) row_starts_event_[r] = 1;
            if (e.kind == EV_QUAD) row_is_quad_x_[r] = 1;
            if (e.kind == EV_LINEAR || e.kind == EV_QUAD) {
                for (int j = 0; j < nr; j++, r++) {
                    host_row_of_.push_back(r);
                    if (!e.val[j]) throw std::invalid_argument("row event without values");
                    memcpy(&val[r * k_ * 8], e.val[j], (size_t)l_ * 32);
                    for (size_t i = 0; i < pad; i++) enc.next(&val[(r * k_ + l_ + i) * 8]);
                    if (e.coef[j]) memcpy(&coef[r * k_ * 8], e.coef[j], (size_t)l_ * 32);
                }
                continue;
            }
            if (e.kind == EV_VSET) {                                          // pads are drawn when the event happens
                host_row_of_.push_back(r);                                    // (its coefficient row is ignored: batch rows take no linear test)
                if (!e.val[0]) throw std::invalid_argument("vbn254fr set without values");
                std::vector<uint32_t> row((size_t)k_ * 8, 0);
                memcpy(row.data(), e.val[0], (size_t)l_ * 32);
                for (size_t i = 0; i < pad; i++) enc.next(&row[(l_ + i) * 8]);
                vset_rows.push_back(std::move(row));
            }
            const int nslots = e.kind == EV_VSET ? 1 : ((e.kind == EV_VADD || e.kind == EV_VSUB || e.kind == EV_VMUL || e.kind == EV_VDIV) ? 3 : 2);
            for (int j = 0; j < nslots; j++)
                if (e.arg[j] >= st.arena_slots) throw std::invalid_argument("vbn254fr variable index out of range");
            r += nr;
        }
        row_starts_event_[rows_] = 1;
        // masks (witness_manager.hpp:271-321)
        mask_[0].assign((size_t)k_ * 8, 0);
        for (size_t i = 0; i < l_; i++) enc.next(&mask_[0][i * 8]);
        mask_[1].assign((size_t)2 * k_ * 8, 0);
        {
            uint64_t sum[4] = {0, 0, 0, 0};
            for (size_t i = 0; i + 1 < l_; i++) { enc.next(&mask_[1][(2 * i + 1) * 8]); add_mod(sum, &mask_[1][(2 * i + 1) * 8]); }
            if (l_ >= 1) { uint32_t neg[8]; negate(neg, sum); memcpy(&mask_[1][(2 * (size_t)l_ - 1) * 8], neg, 32); }
            for (size_t i = 0; i < 2 * pad; i++) enc.next(&mask_[1][(2 * (size_t)l_ + i) * 8]);
        }
        mask_[2].assign((size_t)2 * k_ * 8, 0);
        for (size_t i = 0; i < l_; i++) enc.next(&mask_[2][(2 * i + 1) * 8]);
        for (size_t i = 0; i < 2 * pad; i++) enc.next(&mask_[2][(2 * (size_t)l_ + i) * 8]);

        d_val_ = dalloc(val.size() * 4);
        d_coef_ = dalloc(coef.size() * 4);
        chk(lgr_write(ctx_, d_val_, 0, val.data(), val.size() * 4));
        chk(lgr_write(ctx_, d_coef_, 0, coef.data(), coef.size() * 4));
        if (any_batch) run_arena_program(st, vset_rows);
        // codeword tiles: 2^22 elements (128 MiB) each, at least one triple
        tile_rows_ = std::max<size_t>(3, ((size_t)1 << 22) / n_);
        tile_rows_ = std::min<size_t>(tile_rows_, std::max<size_t>(rows_, 3));
        tile_ = dalloc(tile_rows_ * n_ * 32);
        tile2_ = dalloc(tile_rows_ * n_ * 32);
        mask_cw_ = dalloc((size_t)n_ * 32);
    }

    // The vbn254fr calls of the statement, executed on the device in event order; every on_batch_* row is snapshotted
    // (device-to-device) into its slot of the row matrix, so that the three stages see it like any other row.
    void run_arena_program(const statement &st, const std::vector<std::vector<uint32_t>> &vset_rows) {
        const size_t vb = (size_t)k_ * 32;
        void *arena = dalloc(std::max<size_t>(st.arena_slots, 1) * vb), *tmp = dalloc(vb);
        auto var = [&](uint32_t slot) { return at(arena, (size_t)slot * k_); };
        auto snap = [&](const void *src, size_t row) { chk(lgr_copy(ctx_, src, at(d_val_, row * k_), vb)); };
        size_t r = 0, vs = 0;
        for (const row_event &e : st.events) {
            void *out = var(e.arg[0]);
            switch (e.kind) {
                case EV_VSET:
                    chk(lgr_write(ctx_, out, 0, vset_rows[vs].data(), vb)); vs++;      // values, zero fill, then the pads at [l, k)
                    snap(out, r);
                    break;
                case EV_VCOPY:
                    if (e.arg[0] != e.arg[1]) chk(lgr_copy(ctx_, var(e.arg[1]), out, vb));
                    snap(out, r); snap(var(e.arg[1]), r + 1);
                    break;
                case EV_VASSERT_EQ:
                    snap(var(e.arg[0]), r); snap(var(e.arg[1]), r + 1);
                    break;
                case EV_VADD: chk(lgr_elt_add(ctx_, var(e.arg[1]), var(e.arg[2]), tmp, k_)); chk(lgr_copy(ctx_, tmp, out, vb)); break;
                case EV_VSUB: chk(lgr_elt_sub(ctx_, var(e.arg[1]), var(e.arg[2]), tmp, k_)); chk(lgr_copy(ctx_, tmp, out, vb)); break;
                case EV_VMUL:
                    chk(lgr_elt_mul(ctx_, var(e.arg[1]), var(e.arg[2]), tmp, k_));
                    snap(var(e.arg[1]), r); snap(var(e.arg[2]), r + 1); snap(tmp, r + 2);
                    chk(lgr_copy(ctx_, tmp, out, vb));
                    break;
                case EV_VDIV:
                    chk(lgr_elt_div(ctx_, var(e.arg[1]), var(e.arg[2]), tmp, k_));
                    snap(tmp, r); snap(var(e.arg[2]), r + 1); snap(var(e.arg[1]), r + 2);
                    chk(lgr_copy(ctx_, tmp, out, vb));
                    break;
                case EV_VBIT:
                    if (e.arg[2] >= 256) throw std::invalid_argument("bit index out of range");
                    chk(lgr_elt_bit(ctx_, var(e.arg[1]), tmp, k_, e.arg[2]));
                    chk(lgr_copy(ctx_, tmp, out, vb));
                    snap(out, r);
                    break;
                case EV_VADDC: case EV_VSUBC: case EV_VCSUB: case EV_VMULC: case EV_VMONTMULC: {
                    if (!e.konst) throw std::invalid_argument("constant operation without a constant");
                    const void *x = var(e.arg[1]);
                    if (e.kind == EV_VADDC) chk(lgr_elt_add_const(ctx_, x, tmp, k_, e.konst));
                    else if (e.kind == EV_VSUBC) chk(lgr_elt_sub_const(ctx_, x, tmp, k_, e.konst));
                    else if (e.kind == EV_VCSUB) chk(lgr_elt_const_sub(ctx_, x, tmp, k_, e.konst));
                    else if (e.kind == EV_VMULC) chk(lgr_elt_mul_const(ctx_, x, tmp, k_, e.konst));
                    else chk(lgr_elt_montmul_const(ctx_, x, tmp, k_, e.konst));
                    chk(lgr_copy(ctx_, tmp, out, vb));
                    break;
                }
                default: break;
            }
            r += event_rows(e.kind);
        }
    }

    // mask m as a codeword in mask_cw_: m = 0 a normal row; m = 1, 2 given on the w_2k domain:
    // iNTT_2k then NTT_n (nonbatch_context.hpp:482-494)
    void encode_mask(int m) {
        chk(lgr_write_clear(ctx_, mask_cw_, (size_t)n_ * 32, mask_[m].data(), mask_[m].size() * 4));
        if (m == 0) { chk(lgr_encode(ctx_, mask_cw_)); return; }
        chk(lgr_ntt(ctx_, mask_cw_, LGR_SIZE_2K, LGR_INVERSE));
        chk(lgr_ntt(ctx_, mask_cw_, LGR_SIZE_N, LGR_FORWARD));
    }

    // ---- tiny host field helpers (sums of canonical elements) ----
    static constexpr uint64_t P[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
    static bool geq_p(const uint64_t *a) { for (int i = 3; i >= 0; i--) if (a[i] != P[i]) return a[i] > P[i]; return true; }
    static void sub_p(uint64_t *a) { unsigned __int128 br = 0; for (int i = 0; i < 4; i++) { unsigned __int128 d = (unsigned __int128)a[i] - P[i] - (uint64_t)br; a[i] = (uint64_t)d; br = (d >> 64) & 1; } }
    static void add_mod(uint64_t *acc, const uint32_t *x) {
        uint64_t w[4]; memcpy(w, x, 32);
        unsigned __int128 c = 0;
        for (int i = 0; i < 4; i++) { c += (unsigned __int128)acc[i] + w[i]; acc[i] = (uint64_t)c; c >>= 64; }
        if (geq_p(acc)) sub_p(acc);                 // both < p < 2^254: no carry out
    }
    static void negate(uint32_t *out, const uint64_t *a) {
        uint64_t r[4] = {0, 0, 0, 0};
        if (a[0] | a[1] | a[2] | a[3]) { unsigned __int128 br = 0; for (int i = 0; i < 4; i++) { unsigned __int128 d = (unsigned __int128)P[i] - a[i] - (uint64_t)br; r[i] = (uint64_t)d; br = (d >> 64) & 1; } }
        memcpy(out, r, 32);
    }
    // validate_sum (zkp/common.hpp:68-79) over the first l decoded positions
    static bool sum_is_zero(const uint32_t *v, size_t count, const uint32_t *const_sum) {
        uint64_t acc[4] = {0, 0, 0, 0};
        for (size_t i = 0; i < count; i++) add_mod(acc, v + i * 8);
        add_mod(acc, const_sum);
        return !(acc[0] | acc[1] | acc[2] | acc[3]);
    }

    lgr_ctx *ctx_;
    uint32_t l_ = 0, k_ = 0, n_ = 0;
    size_t rows_ = 0, tile_rows_ = 0;
    std::vector<uint8_t> row_is_quad_x_, row_starts_event_;
    std::vector<size_t> host_row_of_;            // host row (values / coefs order) -> committed row
    std::vector<uint32_t> mask_[3];
    std::vector<void *> owned_;
    void *d_val_ = nullptr, *d_coef_ = nullptr, *tile_ = nullptr, *tile2_ = nullptr, *mask_cw_ = nullptr;
};

}  // namespace ligero::cuda::host

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
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/lgr.h"
#include "proof_wire.hpp"

namespace ligero::cuda::host {

using limbs = uint32_t[8];

// One row event, in emission order: a linear row (1 encoded row) or a quadratic triple
// (3 encoded rows x, y, z).  Pointers are to l canonical elements (8 x u32 each); coef = the row of
// linear-test coefficients the backend accumulated for these witnesses (witness_manager.hpp:141-171).
struct row_event {
    bool quadratic = false;
    const uint32_t *val[3] = {nullptr, nullptr, nullptr};
    const uint32_t *coef[3] = {nullptr, nullptr, nullptr};
};

struct statement {
    uint32_t l = 0, k = 0;                       // n = 4k
    std::vector<row_event> events;
    uint32_t const_sum[8] = {0};                 // linear test: sum_{i<l} P_linear[i] + const_sum == 0 (zkp/common.hpp:68-79)
    uint8_t encoding_seed[32] = {0};             // the reference draws it from std::random_device (src/webgpu_prover.cpp:239-245)
    digest instance_hash, program_hash;
    int64_t generated_at_seconds = -1;           // < 0: wall clock
    uint32_t sample_size = 192;                  // params::sample_size
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

        // ---- stage 2: test vectors -----------------------------------------------------------------
        static const uint8_t any_iv[16] = {0};                                 // params::any_iv
        fr_random_stream code_rng(out.stage1_seed.data, any_iv), quad_rng(out.stage1_seed.data, any_iv);   // nonbatch_context.hpp:105-112
        std::vector<uint32_t> r_code(rows_ * 8), r_quad;
        // draw order = callback order: linear row -> 1 code draw; triple -> 3 code draws then 1 quadratic draw
        {
            size_t r = 0;
            for (const row_event &e : st.events) {
                for (int j = 0; j < (e.quadratic ? 3 : 1); j++, r++) code_rng.next(&r_code[r * 8]);
                if (e.quadratic) { r_quad.resize(r_quad.size() + 8); quad_rng.next(&r_quad[r_quad.size() - 8]); }
            }
        }
        void *code = dalloc((size_t)n_ * 32), *linear = dalloc((size_t)n_ * 32), *quad = dalloc((size_t)n_ * 32);
        size_t quad_seen = 0;
        for_each_tile([&](size_t r0, uint32_t T) {
            chk(lgr_encode_rows(ctx_, at(d_val_, r0 * k_), k_, T, tile_));
            chk(lgr_encode_rows(ctx_, at(d_coef_, r0 * k_), k_, T, tile2_));
            chk(lgr_combine_code(ctx_, tile_, T, &r_code[r0 * 8], code));
            chk(lgr_combine_linear(ctx_, tile_, tile2_, T, linear));
            // the triples of the tile, wherever they sit between linear rows: x, y, z are 3 consecutive codewords
            std::vector<uint32_t> xrows;
            for (size_t r = r0; r < r0 + T; r++) if (row_is_quad_x_[r]) xrows.push_back((uint32_t)(r - r0));
            if (!xrows.empty()) {
                chk(lgr_combine_quad_indexed(ctx_, tile_, xrows.data(), (uint32_t)xrows.size(), &r_quad[quad_seen * 8], quad));
                quad_seen += xrows.size();
            }
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
            out.valid_linear = sum_is_zero(h.data(), l_, st.const_sum);
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

    // rows in emission order -> device, pads from the encoding stream; the three masks
    void build_rows(const statement &st) {
        rows_ = 0;
        for (const row_event &e : st.events) rows_ += e.quadratic ? 3 : 1;
        static const uint8_t any_iv[16] = {0};
        fr_random_stream enc(st.encoding_seed, any_iv);                       // init_encoding_random(seed, params::any_iv)
        const size_t pad = k_ - l_;
        std::vector<uint32_t> val(std::max<size_t>(rows_, 1) * k_ * 8, 0), coef(std::max<size_t>(rows_, 1) * k_ * 8, 0);
        row_is_quad_x_.assign(rows_ + 1, 0);
        row_starts_event_.assign(rows_ + 1, 0);
        size_t r = 0;
        for (const row_event &e : st.events) {
            row_starts_event_[r] = 1;
            if (e.quadratic) row_is_quad_x_[r] = 1;
            for (int j = 0; j < (e.quadratic ? 3 : 1); j++, r++) {
                if (!e.val[j]) throw std::invalid_argument("row event without values");
                memcpy(&val[r * k_ * 8], e.val[j], (size_t)l_ * 32);
                for (size_t i = 0; i < pad; i++) enc.next(&val[(r * k_ + l_ + i) * 8]);
                if (e.coef[j]) memcpy(&coef[r * k_ * 8], e.coef[j], (size_t)l_ * 32);
            }
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
        // codeword tiles: 2^22 elements (128 MiB) each, at least one triple
        tile_rows_ = std::max<size_t>(3, ((size_t)1 << 22) / n_);
        tile_rows_ = std::min<size_t>(tile_rows_, std::max<size_t>(rows_, 3));
        tile_ = dalloc(tile_rows_ * n_ * 32);
        tile2_ = dalloc(tile_rows_ * n_ * 32);
        mask_cw_ = dalloc((size_t)n_ * 32);
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
    std::vector<uint32_t> mask_[3];
    std::vector<void *> owned_;
    void *d_val_ = nullptr, *d_coef_ = nullptr, *tile_ = nullptr, *tile2_ = nullptr, *mask_cw_ = nullptr;
};

}  // namespace ligero::cuda::host

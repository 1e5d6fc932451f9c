// C ABI (include/lgr.h) over the sm_100a kernels: context, twiddle tables, NTT plans, pipelines.
// Host-side counterpart of src/webgpu/engine.cpp + src/webgpu/device_context.cpp of the reference
// (pipeline / bind-group creation and per-op command submission) re-done for one CUDA stream.
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/lgr.h"
#include "host_fr.h"
#include "kernels.h"

using namespace lgr;
using host::Fr;

static thread_local std::string g_err;
static int fail(int code, const std::string &msg) { g_err = msg; return code; }
#define CU(expr)                                                                                         \
    do {                                                                                                 \
        cudaError_t e__ = (expr);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return fail(e__ == cudaErrorMemoryAllocation ? LGR_ERR_NOMEM : LGR_ERR_CUDA,                 \
                        std::string(#expr) + ": " + cudaGetErrorString(e__));                            \
    } while (0)
#define REQUIRE(cond, msg) do { if (!(cond)) return fail(LGR_ERR_INVALID, msg); } while (0)
// every ABI entry point runs with its context's device current, whatever the caller's current device is
#define ENTER_NOFLUSH(c) do { REQUIRE((c) != nullptr, "null context"); CU(cudaSetDevice((c)->device)); } while (0)
// ... and with the one-row encodes lgr_encode deferred (see flush_encodes) enqueued first, so stream order is call order
#define ENTER(c) do { ENTER_NOFLUSH(c); if (!(c)->enc_pending.empty()) { int rc_ = flush_encodes(c); if (rc_) return rc_; } } while (0)
struct lgr_ctx;
static int flush_encodes(lgr_ctx *c);

namespace {

struct DevTable {
    fr_mem *d = nullptr;
    size_t count = 0;
};

struct NttPlan {
    int logn = 0;
    bool inverse = false;
    int l1 = 0, l2 = 0;          // four-step split (l1 + l2 = logn) or l1 = logn, l2 = 0
    DevTable tw1, tw2;           // sub-transform twiddles
    DevTable twist_lo, twist_hi, twist_full;   // four-step twist: full table (N <= 2^20) or lo/hi factors
    int twist_shift = 0;
    DevTable scale;              // N^-1 * R for inverse transforms
};

struct PlanKey {
    int logn; bool inverse; uint64_t w[4];
    bool operator<(const PlanKey &o) const {
        if (logn != o.logn) return logn < o.logn;
        if (inverse != o.inverse) return inverse < o.inverse;
        return memcmp(w, o.w, 32) < 0;
    }
};

}  // namespace

struct lgr_ctx {
    int device = 0;
    uint32_t l = 0, k = 0, n = 0;
    int logk = 0;
    Fr root_k, root_2k, root_n;
    cudaStream_t own_stream = nullptr, stream = nullptr, aux_stream = nullptr;
    cudaEvent_t ev_enc[2] = {nullptr, nullptr}, ev_hash[2] = {nullptr, nullptr}, ev_join = nullptr, ev_fork = nullptr;
    uint64_t launches = 0;
    std::map<PlanKey, NttPlan> plans;
    std::vector<void *> owned;              // tables etc.
    EncodeTables enc{};                      // fused encoder tables (k <= 2048)
    fr_mem *enc_large_twist = nullptr;       // tile-engine encoder (k > 2048): [4][k] coset twists
    bool enc_ready = false;
    uint32_t sys_mul = 0;                    // w_n^4 = w_k^sys_mul (find_sys_mul); 0 = not usable
    fr_mem *scratch = nullptr; size_t scratch_elems = 0;
    fr_mem *tile[2] = {nullptr, nullptr}; size_t tile_elems = 0;
    uint32_t *commit_sha = nullptr;
    // pinned host staging for lgr_write: a ring of kStagingSlots slots, each with its own completion event, so that the
    // host can stage upload i+1 .. i+7 while upload i is still queued behind the kernels of earlier rows
    static constexpr int kStagingSlots = 8;
    void *staging = nullptr; size_t staging_bytes = 0;      // bytes per slot
    cudaEvent_t ev_staging[kStagingSlots] = {};
    int staging_next = 0;
    // one-row encodes at large k are 5 small launches each.  lgr_encode defers them; consecutive encodes of DISTINCT codeword
    // buffers (the reference's stage-2 callbacks issue two or six in a row, nonbatch_context.hpp:667-668,715-720) are replayed
    // as ONE CUDA graph with a parallel branch per buffer, each branch on its own scratch (flush_encodes)
    static constexpr int kEncodeLanes = 8;
    struct EncodeGraph { cudaGraphExec_t exec = nullptr; fr_mem *scratch = nullptr; };
    std::map<std::vector<void *>, EncodeGraph> encode_graphs;
    std::map<void *, int> encode_seen;
    std::vector<void *> enc_pending;
    bool enc_tables_warm = false;            // one plain large-k encode has run: tables, plans exist, capture allocates nothing
    cudaStream_t lane_stream[kEncodeLanes - 1] = {};
    cudaEvent_t ev_lane[kEncodeLanes - 1] = {}, ev_lane_fork = nullptr;
    uint32_t *sample_idx = nullptr; uint32_t sample_count = 0;
    // per-kernel timing of the commit pipeline (lgr_profile): events on the launching streams
    bool profiling = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_enc, prof_sha;
    std::vector<cudaEvent_t> prof_pool;
    fr_mem *h2d_buf[2] = {nullptr, nullptr}; size_t h2d_elems = 0;   // device landing buffers for host-resident rows
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_h2d_free[2] = {nullptr, nullptr};
};

static cudaEvent_t prof_event(lgr_ctx *c) {
    cudaEvent_t e = nullptr;
    if (!c->prof_pool.empty()) { e = c->prof_pool.back(); c->prof_pool.pop_back(); }
    else cudaEventCreate(&e);
    return e;
}

static int upload(lgr_ctx *c, const std::vector<Fr> &v, DevTable &t) {
    t.count = v.size();
    if (v.empty()) { t.d = nullptr; return LGR_OK; }
    CU(cudaMalloc((void **)&t.d, v.size() * 32));
    c->owned.push_back(t.d);
    CU(cudaMemcpy(t.d, v.data(), v.size() * 32, cudaMemcpyHostToDevice));
    return LGR_OK;
}

static int ensure_scratch(lgr_ctx *c, size_t elems) {
    if (c->scratch_elems >= elems) return LGR_OK;
    if (c->scratch) { CU(cudaStreamSynchronize(c->stream)); CU(cudaFree(c->scratch)); c->scratch = nullptr; c->scratch_elems = 0; }
    CU(cudaMalloc((void **)&c->scratch, elems * 32));
    c->scratch_elems = elems;
    return LGR_OK;
}

// ---- NTT plans ---------------------------------------------------------------------------------
static int get_plan(lgr_ctx *c, int logn, const Fr &omega, bool inverse, NttPlan **out) {
    PlanKey key{logn, inverse, {omega.v[0], omega.v[1], omega.v[2], omega.v[3]}};
    auto it = c->plans.find(key);
    if (it != c->plans.end()) { *out = &it->second; return LGR_OK; }
    REQUIRE(logn >= 1 && logn <= 2 * ntt_tile_max_logm(), "transform size out of range (2 .. 2^22 points)");
    // omega must have order exactly 2^logn
    {
        Fr t = omega;
        for (int i = 0; i < logn - 1; i++) t = host::mul(t, t);
        Fr minus1; const uint64_t one[4] = {1, 0, 0, 0}; host::sub4(minus1.v, host::kP, one);
        REQUIRE(host::is_canonical(omega) && t == minus1, "omega is not a primitive 2^logn-th root of unity");
    }
    NttPlan p;
    p.logn = logn; p.inverse = inverse;
    const Fr w = inverse ? host::inv(omega) : omega;
    const size_t N = (size_t)1 << logn;
    int rc;
    if (logn <= ntt_tile_max_logm()) {
        p.l1 = logn; p.l2 = 0;
        if ((rc = upload(c, host::power_table_mont(w, N / 2), p.tw1))) return rc;
    } else {
        p.l2 = logn / 2; p.l1 = logn - p.l2;
        const size_t N1 = (size_t)1 << p.l1, N2 = (size_t)1 << p.l2;
        if ((rc = upload(c, host::power_table_mont(host::pow(w, N2), N1 / 2), p.tw1))) return rc;   // w_N1 = w^N2
        if ((rc = upload(c, host::power_table_mont(host::pow(w, N1), N2 / 2), p.tw2))) return rc;   // w_N2 = w^N1
        if (logn <= 20) {                                       // full table (<= 32 MiB): one multiplication per element
            if ((rc = upload(c, host::power_table_mont(w, N), p.twist_full))) return rc;                // w^e, e = i2*k1 < N
        } else {
            p.twist_shift = (logn + 1) / 2;
            const size_t lo = (size_t)1 << p.twist_shift, hi = N >> p.twist_shift;
            if ((rc = upload(c, host::power_table_mont(w, lo), p.twist_lo))) return rc;
            if ((rc = upload(c, host::power_table_mont(host::pow(w, lo), hi), p.twist_hi))) return rc;
        }
    }
    if (inverse) {
        std::vector<Fr> s(1, host::to_mont(host::inv(host::from_u64(N))));
        if ((rc = upload(c, s, p.scale))) return rc;
    }
    auto ins = c->plans.emplace(key, p);
    *out = &ins.first->second;
    return LGR_OK;
}

// a CTA owns up to 1024 elements (32 KiB of shared memory, 128 threads; four CTAs per SM) -- measured
// 12% faster on the 2^20 transform than 2048-element / 256-thread tiles; 2048-point sub-transforms
// take one lane of 256 threads
static int lanes_per_cta(int logm) {
    const int M = 1 << logm;
    static const int tile = getenv("LGR_NTT_TILE") ? atoi(getenv("LGR_NTT_TILE")) : 1024;   // tuning knob
    int C = std::max(1, tile / M);
    const int TL = M >= 8 ? M / 8 : 1;
    if (C * TL > 256) C = 256 / TL;
    return std::max(C, 1);
}

// One batched transform job.  Plain mode: `batch` transforms, transform b read at in + b*in_stride and
// written to out + b*out_stride (in place allowed for single-pass plans; four-step plans go through
// `tmp`, batch*N elements).  Coset mode (coset_in_twist != nullptr, large-k encoder): the batch index
// packs (row, r) = (b / cosets, b % cosets + coset_base); input row `row` is read from in + row*in_stride and multiplied by
// coset_in_twist[(r - coset_base)*N + i]; output element m of coset r goes to out + row*out_stride + 4*m + r.
struct NttJob {
    const fr_mem *in; long long in_stride;
    fr_mem *out; long long out_stride;
    fr_mem *tmp;
    uint32_t batch;
    const fr_mem *coset_in_twist;
    bool no_scale;
    int cosets = 4;              // coset mode: transforms per row; batch index b = row*cosets + s, r = s + coset_base
    int coset_base = 0;          // coset_in_twist points at the table of coset `coset_base`
    CodewordSink sink{};         // coset mode: nslabs > 0 = slab-major codewords (out / out_stride ignored)
};

static int run_ntt_job(lgr_ctx *c, NttPlan &p, const NttJob &j) {
    if (j.batch == 0) return LGR_OK;
    const long long N = 1ll << p.logn;
    const bool coset = j.coset_in_twist != nullptr;
    const fr_mem *scale = (p.inverse && !j.no_scale) ? p.scale.d : nullptr;
    NttTileParams q{};
    q.in_natural = 1;
    if (p.l2 == 0) {
        REQUIRE(!coset, "internal: coset mode needs a four-step plan");
        q.in = j.in; q.out = j.out;
        q.in_lane_stride = j.in_stride; q.out_lane_stride = j.out_stride;
        q.in_point_stride = q.out_point_stride = 1;
        q.lanes_inner = (int)j.batch; q.total_lanes = (int)j.batch;
        q.logm = p.l1; q.lanes_per_cta = lanes_per_cta(p.l1);
        q.tw = p.tw1.d; q.tws = 1;
        q.scale = scale;
        q.canon = 1;
        CU(launch_ntt_tile(q, c->stream)); c->launches++;
        return LGR_OK;
    }
    const long long N1 = 1ll << p.l1, N2 = 1ll << p.l2;
    REQUIRE((unsigned long long)j.batch * (unsigned long long)N2 < (1ull << 31) && (unsigned long long)j.batch * (unsigned long long)N1 < (1ull << 31), "batch too large");
    REQUIRE(j.tmp, "internal: four-step plan without scratch");
    // pass 1: for every column i2, N1-point transform over i1 (stride N2), twist by w^(i2*k1); in -> tmp
    q.in = j.in; q.out = j.tmp;
    q.in_outer_stride = j.in_stride; q.out_outer_stride = N;
    q.in_lane_stride = q.out_lane_stride = 1;
    q.in_point_stride = q.out_point_stride = N2;
    q.lanes_inner = (int)N2; q.total_lanes = (int)(j.batch * N2);
    q.logm = p.l1; q.lanes_per_cta = lanes_per_cta(p.l1);
    q.tw = p.tw1.d; q.tws = 1;
    q.twist_full = p.twist_full.d; q.twist_lo = p.twist_lo.d; q.twist_hi = p.twist_hi.d; q.twist_shift = p.twist_shift;
    if (coset) { q.in_outer_div = j.cosets; q.in_twist = j.coset_in_twist; q.in_twist_sub_stride = N; }
    q.scale = nullptr; q.canon = 0;
    CU(launch_ntt_tile(q, c->stream)); c->launches++;
    // pass 2: for every k1, N2-point transform over i2 (contiguous); output index k1 + N1*k2; tmp -> out
    NttTileParams r{};
    r.in_natural = 1;
    r.in = j.tmp; r.out = j.out;
    r.in_outer_stride = N;
    r.in_lane_stride = N2; r.in_point_stride = 1;
    if (coset) {
        r.out_outer_div = j.cosets; r.out_outer_stride = j.out_stride; r.out_sub_stride = 1; r.out_sub_base = j.coset_base;
        r.out_lane_stride = 4; r.out_point_stride = 4 * N1;
        r.sink = j.sink;
    } else {
        r.out_outer_stride = j.out_stride;
        r.out_lane_stride = 1; r.out_point_stride = N1;
    }
    r.lanes_inner = (int)N1; r.total_lanes = (int)(j.batch * N1);
    r.logm = p.l2; r.lanes_per_cta = lanes_per_cta(p.l2);
    r.tw = p.tw2.d; r.tws = 1;
    r.scale = scale;
    r.canon = 1;
    CU(launch_ntt_tile(r, c->stream)); c->launches++;
    return LGR_OK;
}

// batch transforms of 2^logn points in place; transform b starts at buf + b*batch_stride (elements)
static int run_ntt(lgr_ctx *c, fr_mem *buf, NttPlan &p, uint32_t batch, size_t batch_stride) {
    if (batch == 0) return LGR_OK;
    if (p.l2 != 0) { int rc = ensure_scratch(c, (size_t)batch << p.logn); if (rc) return rc; }
    NttJob j{buf, (long long)batch_stride, buf, (long long)batch_stride, c->scratch, batch, nullptr, false};
    return run_ntt_job(c, p, j);
}

static int ilog2u(uint64_t x) { int l = 0; while ((1ull << l) < x) l++; return l; }

// c with w_n^4 = w_k^c: both generate the k-th roots of unity, so c exists and is odd.  For the reference's
// roots (w_n from root2 = root1^(2^61-1), src/bn254.cpp:36-43,59-61) c = 2^61-1 mod k = k-1.  Then
// e[4m] = U(w_k^(c m)) = row[c m mod k]: one of the four cosets of the codeword is a permuted copy.
static uint32_t find_sys_mul(const lgr_ctx *c) {
    const Fr wn4 = host::pow(c->root_n, 4);
    if (host::pow(c->root_k, c->k - 1) == wn4) return c->k - 1;
    if (c->k == 2) return 1;
    const Fr wm = host::to_mont(c->root_k);
    Fr x = host::consts().R, target = host::to_mont(wn4);
    for (uint32_t e = 0; e < c->k; e++) { if (x == target) return e; x = host::montmul(x, wm); }
    return 0;                                                   // unreachable for valid roots: compute all four cosets
}

static int build_encode_tables(lgr_ctx *c) {
    if (c->enc_ready) return LGR_OK;
    const size_t k = c->k;
    const int logk = c->logk;
    DevTable a, b, t;
    int rc;
    if ((rc = upload(c, host::power_table_mont(host::inv(c->root_k), k / 2), a))) return rc;
    const Fr wn4 = host::pow(c->root_n, 4);
    if ((rc = upload(c, host::power_table_mont(wn4, k / 2), b))) return rc;
    // twist[r][q] = w_n^(r * bitrev_k(q)) / k * R
    std::vector<Fr> tw(4 * k);
    const Fr kinv_m = host::to_mont(host::inv(host::from_u64(k)));
    for (int r = 0; r < 4; r++) {
        std::vector<Fr> pw = host::power_table_mont(host::pow(c->root_n, r), k);     // (w_n^r)^i * R
        for (size_t q = 0; q < k; q++) {
            size_t i = 0; for (int bit = 0; bit < logk; bit++) if (q >> bit & 1) i |= (size_t)1 << (logk - 1 - bit);
            tw[r * k + q] = host::montmul(pw[i], kinv_m);                              // (x R)(k^-1 R)/R
        }
    }
    if ((rc = upload(c, tw, t))) return rc;
    c->enc.inv_k = a.d; c->enc.fwd_c = b.d; c->enc.twist = t.d; c->enc.sys_mul = (int)c->sys_mul;
    c->enc_ready = true;
    return LGR_OK;
}

// coset twist for the tile-engine encoder: [4][k], natural index: w_n^(r*i) / k * R
static int build_large_encode_tables(lgr_ctx *c) {
    if (c->enc_large_twist) return LGR_OK;
    const size_t k = c->k;
    std::vector<Fr> tw(4 * k);
    const Fr kinv_m = host::to_mont(host::inv(host::from_u64(k)));
    for (int r = 0; r < 4; r++) {
        std::vector<Fr> pw = host::power_table_mont(host::pow(c->root_n, r), k);
        for (size_t i = 0; i < k; i++) tw[r * k + i] = host::montmul(pw[i], kinv_m);
    }
    DevTable t;
    int rc = upload(c, tw, t);
    if (rc) return rc;
    c->enc_large_twist = t.d;
    return LGR_OK;
}

static bool fused_encode_ok(const lgr_ctx *c) { return c->logk >= encode_rows_min_logk() && c->logk <= encode_rows_max_logk(); }

// nrows encodes: rows -> codewords at `sink` (plain sinks may alias the rows: in-place encode)
static int encode_rows_impl(lgr_ctx *c, const fr_mem *rows, size_t row_stride, uint32_t nrows, const CodewordSink &sink, cudaStream_t st) {
    if (nrows == 0) return LGR_OK;
    fr_mem *cw = sink.nslabs ? nullptr : sink.base[0];
    if (fused_encode_ok(c)) {
        int rc = build_encode_tables(c);
        if (rc) return rc;
        CU(launch_encode_rows(rows, (long long)row_stride, sink, (int)nrows, c->logk, c->enc, st)); c->launches++;
        return LGR_OK;
    }
    // large k (> 2048): same decomposition as the fused kernel, on the tile engine.  Coefficients
    // c = iNTT_k(row) (four-step, unscaled) go to scratch; the coset transforms
    // e[4m+r] = NTT_k(c_i * w_n^(r i) / k) with root w_n^4 run as one batched four-step job whose first
    // pass applies the coset twist on load and whose second pass writes the interleaved codeword.
    // Coset 0 is a permuted copy of the row (find_sys_mul).  The zero padding of engine.cpp:755-770's
    // NTT_n is never touched.
    REQUIRE(st == c->stream, "internal: the tile engine runs on the main stream");
    REQUIRE(c->logk > ntt_tile_max_logm(), "internal: small k takes the fused encoder");
    const size_t k = c->k;
    int rc;
    // a handful of rows (the per-row schedule of the reference's stage contexts): ONE launch on thread-block clusters, the row
    // resident in distributed shared memory (cluster_encode_kernel.cu).  Measured on B200 (profiles/r02_per_row.md): 75 us per
    // one-row launch against 54 us for the five latency-kernel launches of the path below (26 dependent butterfly stages on
    // 32 SMs at 4 warps per scheduler), so it is OFF by default; LGR_CLUSTER_ENCODE_ROWS=<rows> turns it on for up to that
    // many rows per call (the parity suite runs it that way once).
    static const uint32_t cluster_max_rows = getenv("LGR_CLUSTER_ENCODE_ROWS") ? (uint32_t)atoi(getenv("LGR_CLUSTER_ENCODE_ROWS")) : 0u;
    if (nrows <= cluster_max_rows && encode_rows_cluster_ok(c->logk)) {
        if ((rc = build_encode_tables(c))) return rc;
        const fr_mem *src = rows; long long src_stride = (long long)row_stride;
        if (!sink.nslabs && rows == cw) {                       // in place: the clusters of a row read it while others write the codeword
            if ((rc = ensure_scratch(c, (size_t)nrows * k))) return rc;
            CU(cudaMemcpy2DAsync(c->scratch, k * 32, rows, row_stride * 32, k * 32, nrows, cudaMemcpyDeviceToDevice, st));
            src = c->scratch; src_stride = (long long)k;
        }
        CU(launch_encode_rows_cluster(src, src_stride, sink, (int)nrows, c->logk, c->enc, st)); c->launches++;
        return LGR_OK;
    }
    if ((rc = build_large_encode_tables(c))) return rc;
    NttPlan *pi, *pf;
    if ((rc = get_plan(c, c->logk, c->root_k, true, &pi))) return rc;
    if ((rc = get_plan(c, c->logk, host::pow(c->root_n, 4), false, &pf))) return rc;
    const bool sys = c->sys_mul != 0, inplace = (rows == cw);
    const uint32_t ncos = sys ? 3 : 4;
    if ((rc = ensure_scratch(c, (size_t)nrows * k * (1 + ncos + ((sys && inplace) ? 1 : 0))))) return rc;
    fr_mem *coef = c->scratch, *tmp = c->scratch + (size_t)nrows * k;
    NttJob inv{rows, (long long)row_stride, coef, (long long)k, tmp, nrows, nullptr, true};
    if ((rc = run_ntt_job(c, *pi, inv))) return rc;
    if (sys) {
        const fr_mem *src = rows; long long src_stride = (long long)row_stride;
        if (inplace) {                                          // the permuted copy would overwrite rows it still has to read
            fr_mem *keep = c->scratch + (size_t)nrows * k * (1 + ncos);
            CU(cudaMemcpy2DAsync(keep, k * 32, rows, row_stride * 32, k * 32, nrows, cudaMemcpyDeviceToDevice, st));
            src = keep; src_stride = (long long)k;
        }
        CU(launch_sys_copy(src, src_stride, sink, (int)nrows, c->logk, c->sys_mul, st)); c->launches++;
    }
    NttJob fwd{coef, (long long)k, cw, sink.row_stride, tmp, nrows * ncos, c->enc_large_twist + (sys ? k : 0), false, (int)ncos, sys ? 1 : 0, sink};
    return run_ntt_job(c, *pf, fwd);
}

// everything lgr_create does after allocating the context: any early return leaves a context lgr_destroy can take
static int create_body(lgr_ctx *c, int device, uint32_t l, uint32_t k, uint32_t n, const uint32_t root_k[8], const uint32_t root_2k[8], const uint32_t root_n[8]) {
    c->device = device; c->l = l; c->k = k; c->n = n; c->logk = ilog2u(k);
    c->root_k = host::from_u32(root_k); c->root_2k = host::from_u32(root_2k); c->root_n = host::from_u32(root_n);
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    CU(cudaStreamCreateWithPriority(&c->own_stream, cudaStreamNonBlocking, lo));
    CU(cudaStreamCreateWithPriority(&c->aux_stream, cudaStreamNonBlocking, hi));
    c->stream = c->own_stream;
    for (int i = 0; i < 2; i++) {
        CU(cudaEventCreateWithFlags(&c->ev_enc[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->ev_hash[i], cudaEventDisableTiming));
    }
    CU(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    for (int i = 0; i < lgr_ctx::kStagingSlots; i++) CU(cudaEventCreateWithFlags(&c->ev_staging[i], cudaEventDisableTiming));
    CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
        CU(cudaEventCreateWithFlags(&c->ev_h2d[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->ev_h2d_free[i], cudaEventDisableTiming));
    }
    if (!getenv("LGR_NO_SYSTEMATIC")) c->sys_mul = find_sys_mul(c);    // debugging knob: compute all four cosets
    // validate the roots and build the six context plans eagerly (engine.cpp:196-211 does the same)
    NttPlan *pl; int rc = LGR_OK;
    for (int inv = 0; inv < 2 && !rc; inv++) {
        rc = get_plan(c, c->logk, c->root_k, inv, &pl);
        if (!rc) rc = get_plan(c, c->logk + 1, c->root_2k, inv, &pl);
        if (!rc) rc = get_plan(c, c->logk + 2, c->root_n, inv, &pl);
    }
    if (!rc && fused_encode_ok(c)) rc = build_encode_tables(c);
    return rc;
}

// ================================================================================================
// Enqueue the one-row encodes lgr_encode deferred (large k only).  They are replayed as ONE CUDA graph per distinct sequence of
// codeword buffers: a branch per buffer -- captured on lane streams forked from and joined back into the main stream -- so the
// two or six rows of a stage-2 callback run side by side instead of back to back (each is ~26 dependent butterfly stages on a
// few dozen CTAs: latency, not throughput).  Every branch has its own scratch; the tables the branches share are read-only.
// The graph is launched on the main stream at the point of the flush, i.e. before anything the caller enqueues after the
// encodes and after everything it enqueued before them: stream order stays call order.
static int flush_encodes(lgr_ctx *c) {
    std::vector<void *> bufs;
    bufs.swap(c->enc_pending);                                                   // whatever happens below, nothing stays deferred
    const size_t m = bufs.size();
    const uint64_t launches = c->launches;                                       // counted when the encodes were deferred
    auto it = c->encode_graphs.find(bufs);
    if (it != c->encode_graphs.end()) { CU(cudaGraphLaunch(it->second.exec, c->stream)); return LGR_OK; }
    auto plain = [&]() -> int {
        int rc = LGR_OK;
        for (size_t i = 0; i < m && rc == LGR_OK; i++) rc = encode_rows_impl(c, (const fr_mem *)bufs[i], c->n, 1, plain_sink((fr_mem *)bufs[i], c->n), c->stream);
        c->launches = launches;
        return rc;
    };
    if (c->encode_graphs.size() >= 64) return plain();
    for (size_t i = 0; i + 1 < m; i++) if (!c->lane_stream[i]) {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        CU(cudaStreamCreateWithPriority(&c->lane_stream[i], cudaStreamNonBlocking, lo));
        CU(cudaEventCreateWithFlags(&c->ev_lane[i], cudaEventDisableTiming));
    }
    if (!c->ev_lane_fork) CU(cudaEventCreateWithFlags(&c->ev_lane_fork, cudaEventDisableTiming));
    const size_t per_row = (size_t)c->k * 5;        // coefficients + four cosets, or + three cosets + the kept row (encode_rows_impl)
    lgr_ctx::EncodeGraph eg;
    CU(cudaMalloc((void **)&eg.scratch, m * per_row * 32));
    fr_mem *const keep_scratch = c->scratch; const size_t keep_elems = c->scratch_elems; cudaStream_t const main_stream = c->stream;
    int rc = LGR_OK;
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamBeginCapture(main_stream, cudaStreamCaptureModeThreadLocal);
    if (e == cudaSuccess) {
        if (m > 1) e = cudaEventRecord(c->ev_lane_fork, main_stream);
        for (size_t i = 0; i < m && rc == LGR_OK && e == cudaSuccess; i++) {
            cudaStream_t st = i ? c->lane_stream[i - 1] : main_stream;
            if (i) e = cudaStreamWaitEvent(st, c->ev_lane_fork, 0);
            if (e != cudaSuccess) break;
            c->stream = st; c->scratch = eg.scratch + i * per_row; c->scratch_elems = per_row;
            rc = encode_rows_impl(c, (const fr_mem *)bufs[i], c->n, 1, plain_sink((fr_mem *)bufs[i], c->n), st);
            if (i && rc == LGR_OK) { e = cudaEventRecord(c->ev_lane[i - 1], st); if (e == cudaSuccess) e = cudaStreamWaitEvent(main_stream, c->ev_lane[i - 1], 0); }
        }
        c->stream = main_stream; c->scratch = keep_scratch; c->scratch_elems = keep_elems; c->launches = launches;
        const cudaError_t e2 = cudaStreamEndCapture(main_stream, &g);                // also ends a capture that failed half-way
        if (e == cudaSuccess) e = e2;
    }
    if (rc == LGR_OK && e == cudaSuccess && g) e = cudaGraphInstantiate(&eg.exec, g, 0);
    if (g) cudaGraphDestroy(g);
    if (rc != LGR_OK || e != cudaSuccess || !eg.exec) {                             // capture refused: plain launches, one row after the other
        cudaGetLastError();
        cudaFree(eg.scratch);
        for (void *b : bufs) c->encode_seen[b] = -1000000;
        return plain();
    }
    c->encode_graphs[bufs] = eg;
    CU(cudaGraphLaunch(eg.exec, c->stream));
    return LGR_OK;
}

extern "C" {

const char *lgr_last_error(void) { return g_err.c_str(); }
int lgr_version(void) { return 100; }

int lgr_create(lgr_ctx **out, int device, uint32_t l, uint32_t k, uint32_t n, const uint32_t p[8], const uint32_t root_k[8],
               const uint32_t root_2k[8], const uint32_t root_n[8]) {
    REQUIRE(out && p && root_k && root_2k && root_n, "null argument");
    REQUIRE(k >= 8 && (k & (k - 1)) == 0, "k must be a power of two >= 8 (the reference needs k >= 512, engine.cpp:850)");
    REQUIRE(n == 4 * k, "n must equal 4k (src/webgpu_prover.cpp:88-97)");
    REQUIRE(l <= k, "l must not exceed k");
    REQUIRE(memcmp(p, host::kP, 32) == 0, "modulus is not the BN254 scalar field (the kernels are specialised, as the reference's WGSL is)");
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    REQUIRE(device >= 0 && device < ndev, "no such CUDA device (this library has no CPU fallback)");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(LGR_ERR_UNSUPPORTED, std::string("liblgr is built for sm_100a only; device is sm_") + std::to_string(prop.major) + std::to_string(prop.minor));
    lgr_ctx *c = new lgr_ctx();
    int rc = create_body(c, device, l, k, n, root_k, root_2k, root_n);
    if (rc) { std::string keep = g_err; lgr_destroy(c); g_err = keep; return rc; }
    *out = c;
    return LGR_OK;
}

int lgr_destroy(lgr_ctx *c) {
    if (!c) return LGR_OK;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (void *p : c->owned) cudaFree(p);
    if (c->scratch) cudaFree(c->scratch);
    for (int i = 0; i < 2; i++) { if (c->tile[i]) cudaFree(c->tile[i]); if (c->ev_enc[i]) cudaEventDestroy(c->ev_enc[i]); if (c->ev_hash[i]) cudaEventDestroy(c->ev_hash[i]); }
    if (c->commit_sha) cudaFree(c->commit_sha);
    if (c->sample_idx) cudaFree(c->sample_idx);
    if (c->staging) cudaFreeHost(c->staging);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    for (int i = 0; i < lgr_ctx::kStagingSlots; i++) if (c->ev_staging[i]) cudaEventDestroy(c->ev_staging[i]);
    for (auto &g : c->encode_graphs) { cudaGraphExecDestroy(g.second.exec); if (g.second.scratch) cudaFree(g.second.scratch); }
    for (int i = 0; i < lgr_ctx::kEncodeLanes - 1; i++) { if (c->lane_stream[i]) cudaStreamDestroy(c->lane_stream[i]); if (c->ev_lane[i]) cudaEventDestroy(c->ev_lane[i]); }
    if (c->ev_lane_fork) cudaEventDestroy(c->ev_lane_fork);
    for (int i = 0; i < 2; i++) { if (c->h2d_buf[i]) cudaFree(c->h2d_buf[i]); if (c->ev_h2d[i]) cudaEventDestroy(c->ev_h2d[i]); if (c->ev_h2d_free[i]) cudaEventDestroy(c->ev_h2d_free[i]); }
    for (cudaEvent_t e : c->prof_pool) cudaEventDestroy(e);
    for (auto &pr : c->prof_enc) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    for (auto &pr : c->prof_sha) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
    delete c;
    return LGR_OK;
}

int lgr_set_stream(lgr_ctx *c, void *s) { ENTER(c); REQUIRE(c, "null context"); c->stream = s ? (cudaStream_t)s : c->own_stream; return LGR_OK; }
int lgr_sync(lgr_ctx *c) { ENTER(c); REQUIRE(c, "null context"); CU(cudaStreamSynchronize(c->stream)); return LGR_OK; }
int lgr_geometry(const lgr_ctx *c, uint32_t *l, uint32_t *k, uint32_t *n) {
    REQUIRE(c, "null context");
    if (l) *l = c->l; if (k) *k = c->k; if (n) *n = c->n;
    return LGR_OK;
}
int lgr_launch_count(const lgr_ctx *c, uint64_t *count) { REQUIRE(c && count, "null argument"); *count = c->launches; return LGR_OK; }

// ---- buffers -----------------------------------------------------------------------------------
int lgr_alloc(lgr_ctx *c, size_t bytes, void **dptr) { ENTER(c);
    REQUIRE(c && dptr, "null argument");
    CU(cudaSetDevice(c->device));
    CU(cudaMallocAsync(dptr, bytes ? bytes : 32, c->stream));            // stream-ordered: no device-wide synchronisation
    CU(cudaMemsetAsync(*dptr, 0, bytes ? bytes : 32, c->stream));
    return LGR_OK;
}
int lgr_free(lgr_ctx *c, void *dptr) { ENTER(c);
    REQUIRE(c, "null context");
    // stream-ordered release: every op that used the buffer was enqueued on (or joined back into) c->stream before this call
    if (dptr) CU(cudaFreeAsync(dptr, c->stream));
    return LGR_OK;
}
int lgr_write(lgr_ctx *c, void *dst, size_t off, const void *src, size_t bytes) { ENTER(c);
    REQUIRE(c && dst && (src || !bytes), "null argument");
    if (!bytes) return LGR_OK;
    // the caller may reuse `src` immediately (nonbatch_context.hpp:455-468): stage through pinned memory.  Each slot of
    // the ring waits only for ITS previous upload, so the host runs up to kStagingSlots uploads ahead of the device.
    if (c->staging_bytes < bytes) {
        if (c->staging) {
            for (int i = 0; i < lgr_ctx::kStagingSlots; i++) CU(cudaEventSynchronize(c->ev_staging[i]));
            CU(cudaFreeHost(c->staging)); c->staging = nullptr; c->staging_bytes = 0;
        }
        size_t cap = std::max(bytes, (size_t)1 << 20);
        CU(cudaMallocHost(&c->staging, cap * lgr_ctx::kStagingSlots));
        c->staging_bytes = cap;
    }
    const int slot = c->staging_next;
    c->staging_next = (slot + 1) % lgr_ctx::kStagingSlots;
    CU(cudaEventSynchronize(c->ev_staging[slot]));
    char *stage = (char *)c->staging + (size_t)slot * c->staging_bytes;
    memcpy(stage, src, bytes);
    CU(cudaMemcpyAsync((char *)dst + off, stage, bytes, cudaMemcpyHostToDevice, c->stream));
    CU(cudaEventRecord(c->ev_staging[slot], c->stream));
    return LGR_OK;
}
int lgr_clear(lgr_ctx *c, void *dst, size_t off, size_t bytes) { ENTER(c);
    REQUIRE(c && dst, "null argument");
    if (bytes) CU(cudaMemsetAsync((char *)dst + off, 0, bytes, c->stream));
    return LGR_OK;
}
int lgr_write_clear(lgr_ctx *c, void *dst, size_t dst_bytes, const void *src, size_t bytes) { ENTER(c);
    REQUIRE(bytes <= dst_bytes, "write_buffer_clear: source larger than destination");
    // The stage contexts export every row into a 2k-element scratch whose tail is zero (nonbatch_context.hpp:415,455-468):
    // trailing zero elements are not staged or sent -- the clear below covers them.
    if (src && (bytes & 31) == 0 && (((uintptr_t)src) & 7) == 0) {
        const uint64_t *w = (const uint64_t *)src;
        size_t e = bytes / 32;
        while (e > 0 && !(w[4 * e - 1] | w[4 * e - 2] | w[4 * e - 3] | w[4 * e - 4])) e--;
        bytes = e * 32;
    }
    int rc = lgr_write(c, dst, 0, src, bytes);
    if (rc) return rc;
    return lgr_clear(c, dst, bytes, dst_bytes - bytes);
}
int lgr_copy(lgr_ctx *c, const void *src, void *dst, size_t bytes) { ENTER(c);
    REQUIRE(c && src && dst, "null argument");
    if (bytes) CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, c->stream));
    return LGR_OK;
}
int lgr_copy_clear(lgr_ctx *c, const void *src, size_t src_bytes, void *dst, size_t dst_bytes) { ENTER(c);
    REQUIRE(src_bytes <= dst_bytes, "copy_buffer_clear: source larger than destination");
    int rc = lgr_copy(c, src, dst, src_bytes);
    if (rc) return rc;
    return lgr_clear(c, dst, src_bytes, dst_bytes - src_bytes);
}
int lgr_read(lgr_ctx *c, void *host_dst, const void *src, size_t off, size_t bytes) { ENTER(c);
    REQUIRE(c && src && (host_dst || !bytes), "null argument");
    if (bytes) CU(cudaMemcpyAsync(host_dst, (const char *)src + off, bytes, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return LGR_OK;
}

// ---- transforms --------------------------------------------------------------------------------
int lgr_ntt(lgr_ctx *c, void *buf, int sel, int dir) { ENTER(c);
    REQUIRE(c && buf, "null argument");
    REQUIRE(sel >= LGR_SIZE_K && sel <= LGR_SIZE_N && (dir == LGR_FORWARD || dir == LGR_INVERSE), "bad size selector / direction");
    const Fr &w = sel == LGR_SIZE_K ? c->root_k : (sel == LGR_SIZE_2K ? c->root_2k : c->root_n);
    NttPlan *p; int rc = get_plan(c, c->logk + sel, w, dir == LGR_INVERSE, &p);
    if (rc) return rc;
    return run_ntt(c, (fr_mem *)buf, *p, 1, (size_t)1 << (c->logk + sel));
}
int lgr_ntt_pow2(lgr_ctx *c, void *buf, uint32_t logn, uint32_t batch, const uint32_t omega[8], int dir) { ENTER(c);
    REQUIRE(c && buf && omega, "null argument");
    REQUIRE(dir == LGR_FORWARD || dir == LGR_INVERSE, "bad direction");
    NttPlan *p; int rc = get_plan(c, (int)logn, host::from_u32(omega), dir == LGR_INVERSE, &p);
    if (rc) return rc;
    return run_ntt(c, (fr_mem *)buf, *p, batch, (size_t)1 << logn);
}
int lgr_encode(lgr_ctx *c, void *buf) { ENTER_NOFLUSH(c);
    REQUIRE(buf, "null argument");
    // Large k: a one-row encode is 5 small launches + a copy on the tile engine.  The stage contexts call it on the same
    // few codeword buffers for every row (nonbatch_context.hpp:445-468,667-668,715-720), so from the third call on a buffer
    // it is deferred until the next call of any other kind and replayed from a CUDA graph (flush_encodes).  Only on the
    // context's own stream (a caller-provided stream may be capturing).  LGR_ENCODE_LANES=1 keeps one row per graph.
    static const bool use_graphs = !getenv("LGR_NO_ENCODE_GRAPH");
    static const int lanes = std::min(std::max(getenv("LGR_ENCODE_LANES") ? atoi(getenv("LGR_ENCODE_LANES")) : lgr_ctx::kEncodeLanes, 1), (int)lgr_ctx::kEncodeLanes);
    int rc;
    if (use_graphs && !fused_encode_ok(c) && c->stream == c->own_stream && c->enc_tables_warm) {
        if (c->encode_seen.size() > 4096) c->encode_seen.clear();
        if (++c->encode_seen[buf] >= 3) {
            // a second encode of the same (or an overlapping) codeword depends on the first: those two stay in order
            const char *b = (const char *)buf; const ptrdiff_t span = (ptrdiff_t)c->n * 32;
            for (void *q : c->enc_pending) if (b - (const char *)q < span && (const char *)q - b < span) { if ((rc = flush_encodes(c))) return rc; break; }
            c->enc_pending.push_back(buf); c->launches += 6;
            return (int)c->enc_pending.size() >= lanes ? flush_encodes(c) : LGR_OK;
        }
    }
    if (!c->enc_pending.empty() && (rc = flush_encodes(c))) return rc;
    rc = encode_rows_impl(c, (const fr_mem *)buf, c->n, 1, plain_sink((fr_mem *)buf, c->n), c->stream);
    if (rc == LGR_OK && !fused_encode_ok(c)) c->enc_tables_warm = true;      // tables, plans and the lazily set kernel attributes exist from here on
    return rc;
}
int lgr_encode_rows(lgr_ctx *c, const void *rows, uint64_t row_stride, uint32_t nrows, void *cw) { ENTER(c);
    REQUIRE(c && rows && cw, "null argument");
    REQUIRE(row_stride >= c->k, "row stride smaller than k");
    REQUIRE(rows != cw || row_stride == c->n, "in-place encode needs a row stride of n elements");
    return encode_rows_impl(c, (const fr_mem *)rows, row_stride, nrows, plain_sink((fr_mem *)cw, c->n), c->stream);
}
int lgr_decode(lgr_ctx *c, void *buf) { ENTER(c);
    REQUIRE(c && buf, "null argument");
    NttPlan *pi, *pf; int rc;
    if ((rc = get_plan(c, c->logk + 2, c->root_n, true, &pi))) return rc;
    if ((rc = get_plan(c, c->logk, c->root_k, false, &pf))) return rc;
    if ((rc = run_ntt(c, (fr_mem *)buf, *pi, 1, c->n))) return rc;
    EltParams e{};                                    // ntt_fold (kernels.wgsl.in:104-116): c[i] += c[i+k]
    e.x = (fr_mem *)buf; e.y = (fr_mem *)buf + c->k; e.out = (fr_mem *)buf; e.n = c->k;
    CU(launch_eltwise(ELT_ADD, e, c->stream)); c->launches++;
    return run_ntt(c, (fr_mem *)buf, *pf, 1, c->k);
}

// ---- hashing -----------------------------------------------------------------------------------
size_t lgr_sha_ctx_bytes(uint32_t ninst) { return sha_ctx_words(ninst) * 4; }
int lgr_sha_init(lgr_ctx *c, void *s, uint32_t ninst) { ENTER(c);
    REQUIRE(c && s, "null argument");
    REQUIRE(ninst > 0 && ninst < (1u << 30), "sha256: instance count out of range");
    CU(launch_sha_init((uint32_t *)s, (int)ninst, c->stream)); c->launches++;
    return LGR_OK;
}
int lgr_sha_update_rows(lgr_ctx *c, void *s, uint32_t ninst, const void *tile, uint64_t row_stride, uint32_t nrows) { ENTER(c);
    REQUIRE(c && s && tile, "null argument");
    REQUIRE(ninst > 0 && ninst < (1u << 30), "sha256: instance count out of range");
    REQUIRE(row_stride >= ninst, "sha256: row stride smaller than the instance count");
    REQUIRE(nrows < (1u << 31), "sha256: too many rows in one call");
    if (!nrows) return LGR_OK;
    CU(launch_sha_update((uint32_t *)s, (int)ninst, (const fr_mem *)tile, (long long)row_stride, (int)nrows, c->stream)); c->launches++;
    return LGR_OK;
}
int lgr_sha_update(lgr_ctx *c, void *s, uint32_t ninst, const void *buf) { ENTER(c); return lgr_sha_update_rows(c, s, ninst, buf, ninst, 1); }
int lgr_sha_final(lgr_ctx *c, const void *s, uint32_t ninst, void *digests) { ENTER(c);
    REQUIRE(c && s && digests, "null argument");
    REQUIRE(ninst > 0 && ninst < (1u << 30), "sha256: instance count out of range");
    CU(launch_sha_final((const uint32_t *)s, (int)ninst, (uint32_t *)digests, c->stream)); c->launches++;
    return LGR_OK;
}
size_t lgr_merkle_node_count(uint32_t nleaves) { size_t p = 1; while (p < nleaves) p <<= 1; return 2 * p - 1; }
int lgr_merkle_build(lgr_ctx *c, const void *leaf, uint32_t nleaves, void *nodes) { ENTER(c);
    REQUIRE(c && leaf && nodes && nleaves, "null argument");
    CU(launch_merkle_build((const uint32_t *)leaf, (int)nleaves, (uint32_t *)nodes, c->stream));
    int P2 = 1, lv = 1; while (P2 < (int)nleaves) P2 <<= 1;
    for (int cnt = P2 >> 1; cnt > 256; cnt >>= 1) lv++;
    c->launches += lv + 1;
    return LGR_OK;
}

// ---- element-wise ------------------------------------------------------------------------------
static int elt(lgr_ctx *c, EltOp op, const void *x, const void *y, const void *z, void *out, size_t n, const uint32_t *sc, bool to_mont,
               const void *idx = nullptr, uint32_t bit = 0) {
    REQUIRE(c && out, "null argument");
    EltParams p{};
    p.x = (const fr_mem *)x; p.y = (const fr_mem *)y; p.z = (const fr_mem *)z; p.out = (fr_mem *)out; p.n = n;
    p.idx = (const uint32_t *)idx; p.bit = bit;
    if (sc) {
        Fr s = host::from_u32(sc);
        REQUIRE(host::is_canonical(s), "scalar is not reduced modulo p");
        if (to_mont) s = host::to_mont(s);
        host::to_u32(p.scalar, s);
    }
    CU(launch_eltwise(op, p, c->stream)); c->launches++;
    return LGR_OK;
}
int lgr_elt_add(lgr_ctx *c, const void *x, const void *y, void *o, size_t n) { ENTER(c); REQUIRE(x && y, "null argument"); return elt(c, ELT_ADD, x, y, 0, o, n, 0, false); }
int lgr_elt_sub(lgr_ctx *c, const void *x, const void *y, void *o, size_t n) { ENTER(c); REQUIRE(x && y, "null argument"); return elt(c, ELT_SUB, x, y, 0, o, n, 0, false); }
int lgr_elt_mul(lgr_ctx *c, const void *x, const void *y, void *o, size_t n) { ENTER(c); REQUIRE(x && y, "null argument"); return elt(c, ELT_MUL, x, y, 0, o, n, 0, false); }
int lgr_elt_div(lgr_ctx *c, const void *x, const void *y, void *o, size_t n) { ENTER(c); REQUIRE(x && y, "null argument"); return elt(c, ELT_DIV, x, y, 0, o, n, 0, false); }
int lgr_elt_fma(lgr_ctx *c, const void *x, const void *y, void *o, size_t n) { ENTER(c); REQUIRE(x && y, "null argument"); return elt(c, ELT_FMA, x, y, 0, o, n, 0, false); }
int lgr_elt_fma_const(lgr_ctx *c, const void *x, void *o, size_t n, const uint32_t k[8]) { ENTER(c); REQUIRE(x && k, "null argument"); return elt(c, ELT_FMA_CONST, x, 0, 0, o, n, k, true); }
int lgr_elt_add_assign(lgr_ctx *c, const void *x, void *o, size_t n) { ENTER(c); REQUIRE(x, "null argument"); return elt(c, ELT_ADD_ASSIGN, x, 0, 0, o, n, 0, false); }
int lgr_elt_add_const(lgr_ctx *c, const void *x, void *o, size_t n, const uint32_t k[8]) { ENTER(c); REQUIRE(x && k, "null argument"); return elt(c, ELT_ADD_CONST, x, 0, 0, o, n, k, false); }
int lgr_elt_sub_const(lgr_ctx *c, const void *x, void *o, size_t n, const uint32_t k[8]) { ENTER(c); REQUIRE(x && k, "null argument"); return elt(c, ELT_SUB_CONST, x, 0, 0, o, n, k, false); }
int lgr_elt_const_sub(lgr_ctx *c, const void *x, void *o, size_t n, const uint32_t k[8]) { ENTER(c); REQUIRE(x && k, "null argument"); return elt(c, ELT_CONST_SUB, x, 0, 0, o, n, k, false); }
int lgr_elt_mul_const(lgr_ctx *c, const void *x, void *o, size_t n, const uint32_t k[8]) { ENTER(c); REQUIRE(x && k, "null argument"); return elt(c, ELT_MUL_CONST, x, 0, 0, o, n, k, true); }
int lgr_elt_montmul_const(lgr_ctx *c, const void *x, void *o, size_t n, const uint32_t k[8]) { ENTER(c); REQUIRE(x && k, "null argument"); return elt(c, ELT_MONTMUL_CONST, x, 0, 0, o, n, k, false); }
int lgr_elt_bit(lgr_ctx *c, const void *x, void *o, size_t n, uint32_t bit) { ENTER(c); REQUIRE(x, "null argument"); REQUIRE(bit < 256, "bit index out of range"); return elt(c, ELT_BIT, x, 0, 0, o, n, 0, false, nullptr, bit); }
int lgr_elt_powmod(lgr_ctx *c, const void *coeff, const void *exp, void *o, size_t n, const uint32_t base[8], int add) { ENTER(c);
    REQUIRE(coeff && exp && base, "null argument");
    return elt(c, add ? ELT_POWADD : ELT_POWMOD, coeff, 0, 0, o, n, base, true, exp);
}
int lgr_elt_quad(lgr_ctx *c, const void *x, const void *y, const void *z, void *o, size_t n, const uint32_t r[8]) { ENTER(c);
    REQUIRE(x && y && z && r, "null argument");
    return elt(c, ELT_QUAD_FUSED, x, y, z, o, n, r, true);
}

// ---- sampling ----------------------------------------------------------------------------------
int lgr_sample_init(lgr_ctx *c, const uint64_t *idx, uint32_t count) { ENTER(c);
    REQUIRE(c && (idx || !count), "null argument");
    if (c->sample_idx) { CU(cudaStreamSynchronize(c->stream)); CU(cudaFree(c->sample_idx)); c->sample_idx = nullptr; }
    c->sample_count = count;
    if (!count) return LGR_OK;
    std::vector<uint32_t> v(count);
    for (uint32_t i = 0; i < count; i++) { REQUIRE(idx[i] < c->n, "sample index out of range"); v[i] = (uint32_t)idx[i]; }
    CU(cudaMalloc((void **)&c->sample_idx, count * 4));
    CU(cudaMemcpy(c->sample_idx, v.data(), count * 4, cudaMemcpyHostToDevice));
    return LGR_OK;
}
int lgr_sample_gather(lgr_ctx *c, const void *x, void *out) { ENTER(c);
    REQUIRE(c && x && out, "null argument");
    REQUIRE(c->sample_idx, "sampling_init has not been called");
    return elt(c, ELT_GATHER, x, 0, 0, out, c->sample_count, 0, false, c->sample_idx);
}
int lgr_sample_gather_rows(lgr_ctx *c, const void *tile, uint64_t row_stride, uint32_t nrows, void *out) { ENTER(c);
    REQUIRE(c && tile && out, "null argument");
    REQUIRE(c->sample_idx, "sampling_init has not been called");
    CU(launch_gather_rows((const fr_mem *)tile, (long long)row_stride, (int)nrows, c->sample_idx, (int)c->sample_count, (fr_mem *)out, c->stream)); c->launches++;
    return LGR_OK;
}

// ---- stage-1 commit pipeline -------------------------------------------------------------------
// tile: even number of rows, 2^23 codeword elements (256 MiB) per buffer: at k = 256 that is 8192 rows
// = 2048 encoder CTAs (4.6 waves of 444 resident CTAs; a 2^21 tile was 1.15 waves and lost 40% to the tail)
static size_t commit_tile_rows(const lgr_ctx *c, uint64_t nrows) {
    static const int tile_log = getenv("LGR_COMMIT_TILE_LOG") ? atoi(getenv("LGR_COMMIT_TILE_LOG")) : 23;   // tuning knob
    size_t T = ((size_t)1 << tile_log) / c->n;
    if (T < 2) T = 2;
    T &= ~(size_t)1;
    if (T > nrows) T = (size_t)nrows;
    if (T == 0) T = 1;
    return T;
}

// rows: device-resident (host_rows == nullptr) or host-resident (pinned or pageable) row-major R x k.
// ext_sha != nullptr: absorb into the caller's column-hash context and stop there (no init, no final, no tree)
static int encode_commit_body(lgr_ctx *c, const fr_mem *rows, const fr_mem *host_rows, uint64_t nrows, void *digests, void *nodes, uint32_t *ext_sha, bool *forked);
// Any failure after the aux / copy streams were forked joins them back before returning, so that a later call cannot
// reuse tile[] / h2d_buf[] while a stream of the failed call is still working on them.
static int encode_commit_impl(lgr_ctx *c, const fr_mem *rows, const fr_mem *host_rows, uint64_t nrows, void *digests, void *nodes, uint32_t *ext_sha = nullptr) {
    bool forked = false;
    const int rc = encode_commit_body(c, rows, host_rows, nrows, digests, nodes, ext_sha, &forked);
    if (rc != LGR_OK && forked) {
        const std::string keep = g_err;
        cudaStreamSynchronize(c->aux_stream);
        cudaStreamSynchronize(c->copy_stream);
        cudaGetLastError();
        g_err = keep;
    }
    return rc;
}
static int encode_commit_body(lgr_ctx *c, const fr_mem *rows, const fr_mem *host_rows, uint64_t nrows, void *digests, void *nodes, uint32_t *ext_sha, bool *forked) {
    REQUIRE(nrows < (1ull << 40), "too many rows");
    const size_t n = c->n, k = c->k;
    const size_t T = commit_tile_rows(c, nrows);
    const bool overlap = true;                          // encodes on the main stream, hashing on the aux stream
    if (c->tile_elems < T * n) {
        CU(cudaStreamSynchronize(c->stream)); CU(cudaStreamSynchronize(c->aux_stream));
        for (int i = 0; i < 2; i++) { if (c->tile[i]) CU(cudaFree(c->tile[i])); c->tile[i] = nullptr; }
        for (int i = 0; i < 2; i++) CU(cudaMalloc((void **)&c->tile[i], T * n * 32));
        c->tile_elems = T * n;
    }
    if (host_rows && c->h2d_elems < T * k) {
        CU(cudaStreamSynchronize(c->stream)); CU(cudaStreamSynchronize(c->copy_stream));
        for (int i = 0; i < 2; i++) { if (c->h2d_buf[i]) CU(cudaFree(c->h2d_buf[i])); c->h2d_buf[i] = nullptr; }
        for (int i = 0; i < 2; i++) CU(cudaMalloc((void **)&c->h2d_buf[i], T * k * 32));
        c->h2d_elems = T * k;
    }
    if (!ext_sha && !c->commit_sha) CU(cudaMalloc((void **)&c->commit_sha, lgr_sha_ctx_bytes((uint32_t)n)));
    uint32_t *sha = ext_sha ? ext_sha : c->commit_sha;
    cudaStream_t es = c->stream, hs = overlap ? c->aux_stream : c->stream, cs = c->copy_stream;
    CU(cudaEventRecord(c->ev_fork, es));
    *forked = true;
    if (overlap) CU(cudaStreamWaitEvent(hs, c->ev_fork, 0));
    if (host_rows) CU(cudaStreamWaitEvent(cs, c->ev_fork, 0));
    if (!ext_sha) { CU(launch_sha_init(sha, (int)n, hs)); c->launches++; }
    int rc;
    size_t tile_idx = 0;
    for (uint64_t r0 = 0; r0 < nrows; r0 += T, tile_idx++) {
        const int b = (int)(tile_idx & 1);
        const uint32_t t = (uint32_t)std::min<uint64_t>(T, nrows - r0);
        const fr_mem *src = rows ? rows + r0 * k : c->h2d_buf[b];
        if (host_rows) {                                 // H2D of tile i overlaps encode of tile i-1 and hash of tile i-2
            if (tile_idx >= 2) CU(cudaStreamWaitEvent(cs, c->ev_h2d_free[b], 0));
            CU(cudaMemcpyAsync(c->h2d_buf[b], host_rows + r0 * k, (size_t)t * k * 32, cudaMemcpyHostToDevice, cs));
            CU(cudaEventRecord(c->ev_h2d[b], cs));
            CU(cudaStreamWaitEvent(es, c->ev_h2d[b], 0));
        }
        if (overlap && tile_idx >= 2) CU(cudaStreamWaitEvent(es, c->ev_hash[b], 0));     // codeword buffer free again
        cudaEvent_t p0 = nullptr, p1 = nullptr;
        if (c->profiling) { p0 = prof_event(c); p1 = prof_event(c); CU(cudaEventRecord(p0, es)); }
        if ((rc = encode_rows_impl(c, src, k, t, plain_sink(c->tile[b], (long long)n), es))) return rc;
        if (c->profiling) { CU(cudaEventRecord(p1, es)); c->prof_enc.emplace_back(p0, p1); }
        if (host_rows) CU(cudaEventRecord(c->ev_h2d_free[b], es));
        if (overlap) { CU(cudaEventRecord(c->ev_enc[b], es)); CU(cudaStreamWaitEvent(hs, c->ev_enc[b], 0)); }
        if (c->profiling) { p0 = prof_event(c); p1 = prof_event(c); CU(cudaEventRecord(p0, hs)); }
        CU(launch_sha_update(sha, (int)n, c->tile[b], (long long)n, (int)t, hs)); c->launches++;
        if (c->profiling) { CU(cudaEventRecord(p1, hs)); c->prof_sha.emplace_back(p0, p1); }
        if (overlap) CU(cudaEventRecord(c->ev_hash[b], hs));
    }
    if (!ext_sha) { CU(launch_sha_final(sha, (int)n, (uint32_t *)digests, hs)); c->launches++; }
    if (!ext_sha && nodes) {
        CU(launch_merkle_build((const uint32_t *)digests, (int)n, (uint32_t *)nodes, hs));
        int lv = 1; for (size_t cnt = n >> 1; cnt > 256; cnt >>= 1) lv++;
        c->launches += lv + 1;
    }
    if (overlap) { CU(cudaEventRecord(c->ev_join, hs)); CU(cudaStreamWaitEvent(es, c->ev_join, 0)); }
    return LGR_OK;
}

int lgr_encode_commit(lgr_ctx *c, const void *rows, uint64_t nrows, void *digests, void *nodes) { ENTER(c);
    REQUIRE(c && rows && digests, "null argument");
    return encode_commit_impl(c, (const fr_mem *)rows, nullptr, nrows, digests, nodes);
}

int lgr_encode_absorb(lgr_ctx *c, void *sha_ctx, const void *rows, uint64_t nrows) { ENTER(c);
    REQUIRE(c && sha_ctx && rows, "null argument");
    if (!nrows) return LGR_OK;
    return encode_commit_impl(c, (const fr_mem *)rows, nullptr, nrows, nullptr, nullptr, (uint32_t *)sha_ctx);
}

int lgr_encode_commit_host(lgr_ctx *c, const void *host_rows, uint64_t nrows, void *host_digests, void *host_root) { ENTER(c);
    REQUIRE(c && host_rows && (host_digests || host_root), "null argument");
    const size_t n = c->n;
    const size_t need = n + (2 * n - 1);                // digests + nodes, in elements of 32 bytes
    int rc = LGR_OK;
    // results live at the tail of the scratch buffer (not used by the fused encoder)
    if (!fused_encode_ok(c)) { if ((rc = ensure_scratch(c, (size_t)c->k * 5 * commit_tile_rows(c, nrows) + need))) return rc; }
    else if ((rc = ensure_scratch(c, need))) return rc;
    fr_mem *dig = c->scratch + (c->scratch_elems - need), *nodes = dig + n;
    if ((rc = encode_commit_impl(c, nullptr, (const fr_mem *)host_rows, nrows, dig, nodes))) return rc;
    if (host_digests) CU(cudaMemcpyAsync(host_digests, dig, n * 32, cudaMemcpyDeviceToHost, c->stream));
    if (host_root) CU(cudaMemcpyAsync(host_root, nodes, 32, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return LGR_OK;
}

int lgr_profile(lgr_ctx *c, int enable) { ENTER(c); REQUIRE(c, "null context"); c->profiling = enable != 0; return LGR_OK; }
int lgr_profile_read(lgr_ctx *c, double *enc_ms, uint64_t *enc_launches, double *sha_ms, uint64_t *sha_launches) { ENTER(c);
    REQUIRE(c && enc_ms && enc_launches && sha_ms && sha_launches, "null argument");
    CU(cudaStreamSynchronize(c->stream)); CU(cudaStreamSynchronize(c->aux_stream));
    double acc[2] = {0, 0};
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> *v[2] = {&c->prof_enc, &c->prof_sha};
    for (int i = 0; i < 2; i++) {
        for (auto &pr : *v[i]) {
            float ms = 0; CU(cudaEventElapsedTime(&ms, pr.first, pr.second)); acc[i] += ms;
            c->prof_pool.push_back(pr.first); c->prof_pool.push_back(pr.second);
        }
    }
    *enc_ms = acc[0]; *enc_launches = c->prof_enc.size(); *sha_ms = acc[1]; *sha_launches = c->prof_sha.size();
    c->prof_enc.clear(); c->prof_sha.clear();
    return LGR_OK;
}

// ---- stage-2 tile combiners --------------------------------------------------------------------
// upload nrows canonical scalars to scratch + at (raw; the kernels rescale them on the device)
static int upload_scalars(lgr_ctx *c, const uint32_t *host_r, uint32_t nrows, size_t at) {
    for (uint32_t i = 0; i < nrows; i++) REQUIRE(host::is_canonical(host::from_u32(host_r + 8 * i)), "scalar is not reduced modulo p");
    return lgr_write(c, c->scratch + at, 0, host_r, (size_t)nrows * 32);
}
int lgr_combine_code(lgr_ctx *c, const void *tile, uint32_t nrows, const uint32_t *host_r, void *acc) { ENTER(c);
    REQUIRE(c && tile && host_r && acc, "null argument");
    if (!nrows) return LGR_OK;
    // scratch: partial sums | scaled scalars (T) | their Karatsuba halves (1.5 T) | raw scalars (T)
    const size_t part = combine_scratch_elems((int)nrows, (int)c->n);
    int rc = ensure_scratch(c, part + 4 * (size_t)nrows);
    if (rc) return rc;
    if ((rc = upload_scalars(c, host_r, nrows, part + 3 * (size_t)nrows))) return rc;
    CU(launch_combine_code((const fr_mem *)tile, (long long)c->n, (int)nrows, (int)c->n, c->scratch + part + 3 * (size_t)nrows, (fr_mem *)acc, c->scratch,
                           part + 3 * (size_t)nrows, c->stream));
    c->launches += 4;
    return LGR_OK;
}
int lgr_combine_quad(lgr_ctx *c, const void *x, const void *y, const void *z, uint32_t nrows, const uint32_t *host_r, void *acc) { ENTER(c);
    return lgr_combine_quad_rows(c, x, y, z, c ? c->n : 0, nrows, host_r, acc);
}
int lgr_combine_quad_rows(lgr_ctx *c, const void *x, const void *y, const void *z, uint64_t row_stride, uint32_t nrows, const uint32_t *host_r, void *acc) { ENTER(c);
    REQUIRE(c && x && y && z && host_r && acc, "null argument");
    REQUIRE(row_stride >= c->n, "row stride smaller than n");
    if (!nrows) return LGR_OK;
    const size_t part = combine_scratch_elems((int)nrows, (int)c->n);
    int rc = ensure_scratch(c, part + 3 * (size_t)nrows);
    if (rc) return rc;
    if ((rc = upload_scalars(c, host_r, nrows, part + 2 * (size_t)nrows))) return rc;
    CU(launch_combine_quad((const fr_mem *)x, (const fr_mem *)y, (const fr_mem *)z, (long long)row_stride, (int)nrows, (int)c->n, c->scratch + part + 2 * (size_t)nrows,
                           (fr_mem *)acc, c->scratch, part + 2 * (size_t)nrows, c->stream));
    c->launches += 4;
    return LGR_OK;
}
int lgr_combine_quad_indexed(lgr_ctx *c, const void *tile, const uint32_t *host_x_rows, uint32_t ntriples, const uint32_t *host_r, void *acc) { ENTER(c);
    REQUIRE(c && tile && host_x_rows && host_r && acc, "null argument");
    if (!ntriples) return LGR_OK;
    const size_t part = combine_scratch_elems((int)ntriples, (int)c->n);
    const size_t idx_elems = ((size_t)ntriples * 4 + 31) / 32;                  // u32 indices, in 32-byte elements
    int rc = ensure_scratch(c, part + 3 * (size_t)ntriples + idx_elems);
    if (rc) return rc;
    if ((rc = upload_scalars(c, host_r, ntriples, part + 2 * (size_t)ntriples))) return rc;
    fr_mem *idx = c->scratch + part + 3 * (size_t)ntriples;
    if ((rc = lgr_write(c, idx, 0, host_x_rows, (size_t)ntriples * 4))) return rc;
    const fr_mem *x = (const fr_mem *)tile;
    CU(launch_combine_quad(x, x + c->n, x + 2 * (size_t)c->n, (long long)c->n, (int)ntriples, (int)c->n, c->scratch + part + 2 * (size_t)ntriples,
                           (fr_mem *)acc, c->scratch, part + 2 * (size_t)ntriples, c->stream, (const uint32_t *)idx));
    c->launches += 4;
    return LGR_OK;
}
int lgr_combine_bit_indexed(lgr_ctx *c, const void *tile, const uint32_t *host_rows, uint32_t count, const uint32_t *host_r, void *acc) {
    ENTER(c);
    REQUIRE(tile && host_rows && host_r && acc, "null argument");
    if (!count) return LGR_OK;
    const size_t part = combine_scratch_elems((int)count, (int)c->n);
    const size_t idx_elems = ((size_t)count * 4 + 31) / 32;
    int rc = ensure_scratch(c, part + 3 * (size_t)count + idx_elems);
    if (rc) return rc;
    if ((rc = upload_scalars(c, host_r, count, part + 2 * (size_t)count))) return rc;
    fr_mem *idx = c->scratch + part + 3 * (size_t)count;
    if ((rc = lgr_write(c, idx, 0, host_rows, (size_t)count * 4))) return rc;
    const fr_mem *x = (const fr_mem *)tile;                                 // x = y = z: r (x*x - x), what on_batch_bit's copies give check_quadratic
    CU(launch_combine_quad(x, x, x, (long long)c->n, (int)count, (int)c->n, c->scratch + part + 2 * (size_t)count,
                           (fr_mem *)acc, c->scratch, part + 2 * (size_t)count, c->stream, (const uint32_t *)idx));
    c->launches += 4;
    return LGR_OK;
}
int lgr_combine_linear(lgr_ctx *c, const void *a, const void *b, uint32_t nrows, void *acc) { ENTER(c);
    REQUIRE(c && a && b && acc, "null argument");
    if (!nrows) return LGR_OK;
    const size_t part = combine_scratch_elems((int)nrows, (int)c->n);
    int rc = ensure_scratch(c, part);
    if (rc) return rc;
    CU(launch_combine_linear((const fr_mem *)a, (const fr_mem *)b, (long long)c->n, (int)nrows, (int)c->n, (fr_mem *)acc, c->scratch, part, c->stream));
    c->launches += 2;
    return LGR_OK;
}

// ---- exact multi-GPU layout: slab-major codewords, peer memory, hand-over flags -------------------
int lgr_encode_rows_slabs(lgr_ctx *c, const void *rows, uint64_t row_stride, uint32_t nrows, void *const *slab_base, uint32_t nslabs) {
    ENTER(c);
    REQUIRE(rows && slab_base, "null argument");
    REQUIRE(row_stride >= c->k, "row stride smaller than k");
    REQUIRE(nslabs >= 1 && nslabs <= 8 && (nslabs & (nslabs - 1)) == 0 && c->n % nslabs == 0, "slab count must be 1, 2, 4 or 8");
    CodewordSink s{};
    s.nslabs = (int)nslabs;
    s.slab_shift = ilog2u(c->n / nslabs);
    s.row_stride = (long long)(c->n / nslabs);
    for (uint32_t h = 0; h < nslabs; h++) { REQUIRE(slab_base[h], "null slab base"); s.base[h] = (fr_mem *)slab_base[h]; }
    return encode_rows_impl(c, (const fr_mem *)rows, row_stride, nrows, s, c->stream);
}
int lgr_ipc_alloc(lgr_ctx *c, size_t bytes, void **dptr, unsigned char handle[64]) {
    ENTER(c);
    REQUIRE(dptr && handle && bytes, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    CU(cudaMalloc(dptr, bytes));                               // IPC needs a cudaMalloc allocation (not the stream-ordered pool)
    CU(cudaMemset(*dptr, 0, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, *dptr);
    if (e != cudaSuccess) { cudaFree(*dptr); *dptr = nullptr; return fail(LGR_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e)); }
    memcpy(handle, &h, 64);
    return LGR_OK;
}
int lgr_ipc_open(lgr_ctx *c, const unsigned char handle[64], void **dptr) {
    ENTER(c);
    REQUIRE(dptr && handle, "null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CU(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
    return LGR_OK;
}
int lgr_ipc_close(lgr_ctx *c, void *dptr) {
    ENTER(c);
    if (dptr) { CU(cudaDeviceSynchronize()); CU(cudaIpcCloseMemHandle(dptr)); }
    return LGR_OK;
}
int lgr_ipc_free(lgr_ctx *c, void *dptr) {
    ENTER(c);
    if (dptr) { CU(cudaDeviceSynchronize()); CU(cudaFree(dptr)); }
    return LGR_OK;
}
int lgr_peer_signal(lgr_ctx *c, void *const *slots, uint32_t nslots, uint64_t value) {
    ENTER(c);
    REQUIRE(slots && nslots <= 8, "at most 8 peers");
    PeerSlots s{};
    s.n = (int)nslots;
    for (uint32_t i = 0; i < nslots; i++) { REQUIRE(slots[i], "null flag slot"); s.p[i] = (unsigned long long *)slots[i]; }
    CU(launch_peer_signal(s, value, c->stream)); c->launches++;
    return LGR_OK;
}
int lgr_peer_wait(lgr_ctx *c, const void *flags, uint32_t nflags, uint64_t value, uint32_t timeout_ms, void *err_flag) {
    ENTER(c);
    REQUIRE(flags && err_flag && nflags <= 32, "bad arguments");
    CU(launch_peer_wait((const unsigned long long *)flags, (int)nflags, value, (unsigned long long)timeout_ms * 1000000ull, (unsigned int *)err_flag, c->stream));
    c->launches++;
    return LGR_OK;
}

// ---- synthetic data ---------------------------------------------------------
int lgr_synth(lgr_ctx *c, void *out, uint64_t seed, uint64_t row0, uint64_t nrows, uint64_t ncols) { ENTER(c);
    REQUIRE(c && out, "null argument");
    CU(launch_synth((fr_mem *)out, seed, row0, nrows, ncols, c->stream)); c->launches++;
    return LGR_OK;
}
}  // extern "C"

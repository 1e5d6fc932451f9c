// TEST INFRASTRUCTURE (oracle side): host driver for the reference's OWN shaders, transliterated to C++ by
// oracle/wgsl2cpp.py and compiled here into oracle/_ref/libwgslref.so.  This is the closest thing to "running
// the reference" that this image allows (Dawn/WebGPU is absent): every field operation, butterfly, hash byte
// and reduction below is executed by the reference's shader text; this file only restates WHICH entry point
// is dispatched in WHICH order with WHICH bindings, following src/webgpu/engine.cpp:
//   ntt_forward_kernel   engine.cpp:844-882     ntt_inverse_kernel  engine.cpp:932-968
//   encode / decode      engine.cpp:755-796     twiddle tables      engine.cpp:1382-1503
//   sha256 init/update/final  engine.cpp:1606-1666
//   Eltwise* dispatch    engine.cpp:420-750     powmod table        src/webgpu/powmod_context.cpp:232-249
// Host big-number work the reference does with GMP (twiddle powers, N^-1, the powmod table) is done here with
// the shader's own Barrett product and Euclid inverse, so no third arithmetic implementation is involved.
// Only tests/ may load this library (see oracle/README in DESIGN.md section 2); the product never does.
#include "wgsl_shim.hpp"

#include <map>
#include <string>
#include <utility>

namespace wgsl {
namespace kern {
#include "_ref/wgsl_kernels.gen.inc"
static_assert(sizeof(bigint) == 32, "bigint must be 8 packed u32");
static_assert(sizeof(ntt_config_t) == 256, "ntt_config_t is padded to 256 bytes (kernels.wgsl.in:24-33)");
}  // namespace kern

}  // namespace wgsl

// The sha256 module is compiled once per instance count, exactly as the reference re-compiles the shader with
// `#INSTANCES` replaced (engine.cpp:1514-1527).  Instance counts the tests use:
#define WGSL_SHA_COUNTS(X) X(1) X(5) X(192) X(256) X(1024) X(2048) X(4096) X(32768)
namespace wgsl {
namespace sha_1 {
#define WGSL_SHA_INSTANCES 1
#include "_ref/wgsl_sha256.gen.inc"
#undef WGSL_SHA_INSTANCES
}
namespace sha_5 {
#define WGSL_SHA_INSTANCES 5
#include "_ref/wgsl_sha256.gen.inc"
#undef WGSL_SHA_INSTANCES
}
namespace sha_192 {
#define WGSL_SHA_INSTANCES 192
#include "_ref/wgsl_sha256.gen.inc"
#undef WGSL_SHA_INSTANCES
}
namespace sha_256 {
#define WGSL_SHA_INSTANCES 256
#include "_ref/wgsl_sha256.gen.inc"
#undef WGSL_SHA_INSTANCES
}
namespace sha_1024 {
#define WGSL_SHA_INSTANCES 1024
#include "_ref/wgsl_sha256.gen.inc"
#undef WGSL_SHA_INSTANCES
}
namespace sha_2048 {
#define WGSL_SHA_INSTANCES 2048
#include "_ref/wgsl_sha256.gen.inc"
#undef WGSL_SHA_INSTANCES
}
namespace sha_4096 {
#define WGSL_SHA_INSTANCES 4096
#include "_ref/wgsl_sha256.gen.inc"
#undef WGSL_SHA_INSTANCES
}
namespace sha_32768 {
#define WGSL_SHA_INSTANCES 32768
#include "_ref/wgsl_sha256.gen.inc"
#undef WGSL_SHA_INSTANCES
}
}  // namespace wgsl

using namespace wgsl;
using kern::bigint;

namespace {

constexpr u32 WG = 256;                      // include/ligetron/webgpu/common.hpp: workgroup_size
constexpr u32 MAX_WORKGROUPS = 256;          // common.hpp:33
constexpr u32 NTT_SHARED_SIZE = WG * 2;      // common.hpp:35
constexpr u32 NTT_SHARED_ITERS = 9;          // common.hpp:36 countr_zero(512)
// num_default_workgroups_ = calc_blocks(hardware cores, 256) (engine.cpp:1199); every kernel dispatched with it
// is a grid-stride loop, so the value does not change results.  8 keeps host runs short.
constexpr u32 DEFAULT_WORKGROUPS = 8;

const entry_point& kern_entry(const char* name) {
    for (const entry_point& e : kern::entry_points)
        if (!std::strcmp(e.name, name)) return e;
    std::fprintf(stderr, "wgslref: no entry point %s\n", name);
    std::abort();
}

bigint from_limbs(const u32* w) { bigint b; std::memcpy(&b, w, 32); return b; }
bigint mulmod(const bigint& a, const bigint& b) { return kern::barrett_reduce_wide(kern::bigint_mul_wide(a, b)); }

u32 ilog2(u32 n) { u32 l = 0; while ((1u << l) < n) ++l; return l; }

// ntt_precompute_omegas (engine.cpp:1382-1503): tables[0] = shared-pass table (stages 1..9 concatenated),
// tables[i] = stage-i table (M = 2^i, M/2 entries, stride N/M), all entries w^j * R mod p; config[i] = {N^-1 R, N, log2N, 2^i, i}
struct ntt_plan {
    u32 N = 0, log2N = 0;
    std::vector<std::vector<bigint>> omegas, omegas_inv;
    std::vector<kern::ntt_config_t> configs;
};

std::vector<std::vector<bigint>> stage_tables(const bigint& root, u32 N, u32 log2N) {
    std::vector<bigint> pw(N / 2);
    bigint cur = kern::BN254_mont_R;                   // w^0 * R
    for (u32 i = 0; i < N / 2; ++i) { pw[i] = cur; cur = mulmod(cur, root); }
    std::vector<std::vector<bigint>> t(log2N + 1);
    for (u32 i = 1; i <= log2N; ++i) {
        u32 M = 1u << i, num = M / 2, stride = N / M;
        t[i].resize(num);
        for (u32 j = 0; j < num; ++j) t[i][j] = pw[size_t(j) * stride];
    }
    t[0].resize((1u << NTT_SHARED_ITERS) - 1);
    for (u32 i = 1, base = 0; i <= NTT_SHARED_ITERS && i <= log2N; ++i) {
        u32 M = 1u << i, num = M / 2, stride = N / M;
        for (u32 j = 0; j < num; ++j) t[0][base + j] = pw[size_t(j) * stride];
        base += num;
    }
    return t;
}

const ntt_plan& plan_for(const u32* root_limbs, u32 N) {
    static std::map<std::pair<std::string, u32>, ntt_plan> cache;
    std::pair<std::string, u32> key(std::string(reinterpret_cast<const char*>(root_limbs), 32), N);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    ntt_plan p;
    p.N = N;
    p.log2N = ilog2(N);
    bigint root = from_limbs(root_limbs);
    p.omegas = stage_tables(root, N, p.log2N);
    p.omegas_inv = stage_tables(kern::bn254fr_invmod(root), N, p.log2N);
    bigint n_inv = mulmod(kern::bn254fr_invmod(kern::bigint_from_u32(N)), kern::BN254_mont_R);
    for (u32 i = 0; i <= p.log2N; ++i) {
        kern::ntt_config_t c;
        c.N_inv = n_inv;
        c.params = vec4u(N, p.log2N, 1u << i, i);
        p.configs.push_back(c);
    }
    return cache.emplace(key, std::move(p)).first->second;
}

void set_group1(const ntt_plan& p, const std::vector<std::vector<bigint>>& tab, u32 iter) {
    kern::ntt_config = p.configs[iter];
    kern::ntt_omegas.bind(const_cast<bigint*>(tab[iter].data()), tab[iter].size());
}

// engine.cpp:844-882
void ntt_forward(u32* buf, size_t buf_elems, const ntt_plan& p) {
    kern::ntt_buffer.bind(buf, buf_elems);
    for (u32 iter = p.log2N; iter > NTT_SHARED_ITERS; --iter) {
        set_group1(p, p.omegas, iter);
        rt::dispatch(kern_entry("ntt_forward_radix2"), DEFAULT_WORKGROUPS);
    }
    u32 shared_wg = p.N / NTT_SHARED_SIZE;
    if (shared_wg <= MAX_WORKGROUPS) {
        set_group1(p, p.omegas, 0);
        rt::dispatch(kern_entry("ntt_forward_radix2_shared"), shared_wg);
    } else {
        for (u32 iter = NTT_SHARED_ITERS; iter >= 1; --iter) {
            set_group1(p, p.omegas, iter);
            rt::dispatch(kern_entry("ntt_forward_radix2"), DEFAULT_WORKGROUPS);
        }
        rt::dispatch(kern_entry("ntt_reduce4p"), DEFAULT_WORKGROUPS);
    }
    set_group1(p, p.omegas, 0);
    rt::dispatch(kern_entry("ntt_bit_reverse"), DEFAULT_WORKGROUPS);
}

// engine.cpp:932-968
void ntt_inverse(u32* buf, size_t buf_elems, const ntt_plan& p) {
    kern::ntt_buffer.bind(buf, buf_elems);
    set_group1(p, p.omegas_inv, 0);
    rt::dispatch(kern_entry("ntt_bit_reverse"), DEFAULT_WORKGROUPS);
    u32 shared_wg = p.N / NTT_SHARED_SIZE;
    if (shared_wg <= MAX_WORKGROUPS) {
        rt::dispatch(kern_entry("ntt_inverse_radix2_shared"), shared_wg);
    } else {
        for (u32 iter = 1; iter <= NTT_SHARED_ITERS; ++iter) {
            set_group1(p, p.omegas_inv, iter);
            rt::dispatch(kern_entry("ntt_inverse_radix2"), DEFAULT_WORKGROUPS);
        }
    }
    for (u32 iter = NTT_SHARED_ITERS + 1; iter <= p.log2N; ++iter) {
        set_group1(p, p.omegas_inv, iter);
        rt::dispatch(kern_entry("ntt_inverse_radix2"), DEFAULT_WORKGROUPS);
    }
    rt::dispatch(kern_entry("ntt_adjust_inverse_reduce"), DEFAULT_WORKGROUPS);
}

bool pow2_ge512(u32 n) { return n >= NTT_SHARED_SIZE && (n & (n - 1)) == 0; }

}  // namespace

extern "C" {

int wref_abi_version() { return 1; }

// one transform of the first N elements of buf (buf_elems >= N elements are bound, as the 4k-element codeword
// buffer is in the reference).  N >= 512 (engine.cpp:850 asserts log2N >= 9).
int wref_ntt(u32* buf, u32 buf_elems, u32 N, const u32* root, int inverse) {
    if (!pow2_ge512(N) || buf_elems < N) return -1;
    const ntt_plan& p = plan_for(root, N);
    if (inverse) ntt_inverse(buf, buf_elems, p); else ntt_forward(buf, buf_elems, p);
    return 0;
}

// encode_ntt_device (engine.cpp:755-770): inverse over k with w_k, forward over n = 4k with w_n, in place
int wref_encode(u32* buf, u32 k, const u32* root_k, const u32* root_n) {
    if (!pow2_ge512(k)) return -1;
    ntt_inverse(buf, 4 * size_t(k), plan_for(root_k, k));
    ntt_forward(buf, 4 * size_t(k), plan_for(root_n, 4 * k));
    return 0;
}

// decode_ntt_device (engine.cpp:772-796): inverse over n, ntt_fold with the 2k config (half = k), forward over k.
// root_2k is only needed because the fold takes its N from ntt_forward_bindings_2k_[0].
int wref_decode(u32* buf, u32 k, const u32* root_k, const u32* root_2k, const u32* root_n) {
    if (!pow2_ge512(k)) return -1;
    ntt_inverse(buf, 4 * size_t(k), plan_for(root_n, 4 * k));
    const ntt_plan& p2k = plan_for(root_2k, 2 * k);
    kern::ntt_buffer.bind(buf, 4 * size_t(k));
    set_group1(p2k, p2k.omegas, 0);
    rt::dispatch(kern_entry("ntt_fold"), DEFAULT_WORKGROUPS);
    ntt_forward(buf, 4 * size_t(k), plan_for(root_k, k));
    return 0;
}

// the stage-`stage` twiddle table the shaders read (stage 0 = shared-pass table); returns the entry count
int wref_twiddles(u32 N, const u32* root, int inverse, u32 stage, u32* out, u32 out_elems) {
    if (!pow2_ge512(N)) return -1;
    const ntt_plan& p = plan_for(root, N);
    if (stage > p.log2N) return -1;
    const std::vector<bigint>& t = (inverse ? p.omegas_inv : p.omegas)[stage];
    if (out_elems < t.size()) return -1;
    std::memcpy(out, t.data(), t.size() * 32);
    return int(t.size());
}

// N^-1 * R mod p as written into ntt_config_t.N_inv
int wref_n_inv(u32 N, const u32* root, u32* out) {
    if (!pow2_ge512(N)) return -1;
    std::memcpy(out, &plan_for(root, N).configs[0].N_inv, 32);
    return 0;
}

// ---- field primitives, for direct checks -----------------------------------------------------------------
void wref_montgomery_mul(const u32* a, const u32* b, u32* out, int two_p) {
    bigint r = two_p ? kern::montgomery_mul_2p(from_limbs(a), from_limbs(b)) : kern::montgomery_mul(from_limbs(a), from_limbs(b));
    std::memcpy(out, &r, 32);
}
void wref_barrett_mul(const u32* a, const u32* b, u32* out) {
    bigint r = mulmod(from_limbs(a), from_limbs(b));
    std::memcpy(out, &r, 32);
}
void wref_invmod(const u32* a, u32* out) {
    bigint r = kern::bn254fr_invmod(from_limbs(a));
    std::memcpy(out, &r, 32);
}
void wref_constants(u32* p, u32* two_p, u32* mont_inv, u32* mont_r, u32* barrett) {
    std::memcpy(p, &kern::BN254_p, 32);
    std::memcpy(two_p, &kern::BN254_2p, 32);
    std::memcpy(mont_inv, &kern::BN254_mont_inv, 32);
    std::memcpy(mont_r, &kern::BN254_mont_R, 32);
    std::memcpy(barrett, &kern::BN254_barrett_factor, 32);
}

// ---- Eltwise* (engine.cpp:420-750): x, y, out bound with `count` elements each; scalar through the uniform ----
// EltwiseBitDecompose has no grid-stride loop (kernels.wgsl.in:501-511) and out-of-range invocations are no-ops in
// WGSL (robust buffer access); here the buffers are padded to the dispatch size instead.
int wref_eltwise(const char* name, const u32* x, const u32* y, u32* out, const u32* scalar, u32 count) {
    const entry_point* ep = nullptr;
    for (const entry_point& e : kern::entry_points)
        if (!std::strcmp(e.name, name)) ep = &e;
    if (!ep || std::strncmp(name, "Eltwise", 7) != 0 || std::strstr(name, "Pow")) return -1;
    if (scalar) std::memcpy(&kern::input_scalar, scalar, 32);
    if (!std::strcmp(name, "EltwiseBitDecompose")) {
        u32 nwg = (count + WG - 1) / WG;
        std::vector<bigint> px(size_t(nwg) * WG), po(size_t(nwg) * WG);
        std::memcpy(px.data(), x, size_t(count) * 32);
        kern::vector_x.bind(px.data(), px.size());
        kern::vector_out.bind(po.data(), po.size());
        rt::dispatch(*ep, nwg);
        std::memcpy(out, po.data(), size_t(count) * 32);
        return 0;
    }
    kern::vector_x.bind(const_cast<u32*>(x), count);
    kern::vector_y.bind(const_cast<u32*>(y ? y : x), count);
    kern::vector_out.bind(out, count);
    rt::dispatch(*ep, DEFAULT_WORKGROUPS);
    return 0;
}

// powmod_context::set_base (powmod_context.cpp:232-249): table[i] = base^(2^i) * R mod p; then EltwisePowMod /
// EltwisePowAddMod with `workgroups` workgroups (tests/webgpu/test_powmod.cpp dispatches test_workgroups)
int wref_powmod(int add, const u32* base, const u32* exp, const u32* coeff, u32* out, u32 count, u32 workgroups) {
    bigint rpow = from_limbs(base);
    for (u32 i = 0; i < 32; ++i) {
        kern::powmod_table[i] = mulmod(rpow, kern::BN254_mont_R);
        rpow = mulmod(rpow, rpow);
    }
    kern::powmod_exp.bind(const_cast<u32*>(exp), count);
    kern::powmod_coeff.bind(const_cast<u32*>(coeff), count);
    kern::powmod_out.bind(out, count);
    rt::dispatch(kern_entry(add ? "EltwisePowAddMod" : "EltwisePowMod"), workgroups);
    return 0;
}

// sample_gather (kernels.wgsl.in:541-549): index i lives in component 0 of a vec4 (engine.cpp:1792-1809)
int wref_sample_gather(const u32* x, u32 nx, const u32* idx192, u32* out192) {
    for (u32 i = 0; i < 192; ++i) kern::sample_index[i] = vec4u(idx192[i], 0, 0, 0);
    kern::vector_x.bind(const_cast<u32*>(x), nx);
    kern::vector_out.bind(out192, 192);
    rt::dispatch(kern_entry("sample_gather"), DEFAULT_WORKGROUPS);
    return 0;
}

// ---- sha256 (engine.cpp:1606-1666): ctx is the reference's SoA context buffer, 75 u32 per instance -------------
u32 wref_sha_ctx_words(u32 ninst) { return 75u * ninst; }

#define SHA_DISPATCH(NS, ENTRY)                                                           \
    {                                                                                     \
        static_assert(sizeof(NS::ctx) == 75u * 4u * NS::num_instances, "SoA layout");     \
        std::memcpy(&NS::ctx, ctx, sizeof(NS::ctx));                                      \
        const entry_point* ep = nullptr;                                                  \
        for (const entry_point& e : NS::entry_points)                                     \
            if (!std::strcmp(e.name, ENTRY)) ep = &e;                                     \
        NS::digest.bind(digest, digest ? NS::num_instances : 0);                          \
        NS::input.bind(const_cast<u32*>(input), input ? size_t(NS::num_instances) * 8 : 0); \
        rt::dispatch(*ep, (NS::num_instances + WG - 1) / WG);                             \
        std::memcpy(ctx, &NS::ctx, sizeof(NS::ctx));                                      \
        return 0;                                                                         \
    }

static int sha_call(const char* entry, u32* ctx, u32 ninst, const u32* input, u32* digest_words) {
    void* digest = digest_words;
    switch (ninst) {
#define X(N) case N: { namespace S = wgsl::sha_##N; auto* dg = static_cast<S::sha256_digest*>(digest); \
                       { auto* digest = dg; SHA_DISPATCH(S, entry) } }
        WGSL_SHA_COUNTS(X)
#undef X
        default: return -1;
    }
}

// sha256_digest_init clears the context buffer first (engine.cpp:1609-1614)
int wref_sha_init(u32* ctx, u32 ninst) {
    std::memset(ctx, 0, size_t(75) * 4 * ninst);
    return sha_call("sha256_init", ctx, ninst, nullptr, nullptr);
}
int wref_sha_update(u32* ctx, u32 ninst, const u32* input) { return sha_call("sha256_update", ctx, ninst, input, nullptr); }
int wref_sha_final(u32* ctx, u32 ninst, u32* digest) { return sha_call("sha256_final", ctx, ninst, nullptr, digest); }

}  // extern "C"

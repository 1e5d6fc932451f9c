// Host-side BN254 Fr used only to build device tables (twiddles, twists, N^-1, scalar
// conversions).  Replaces the GMP calls of webgpu_context::ntt_precompute_omegas
// (src/webgpu/engine.cpp:1382-1503: mpz_powm_ui / mpz_invert / "<<256 mod p").
// 4 x u64 limbs, Montgomery CIOS with unsigned __int128.  Product code: not shared with oracle/.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace lgr {
namespace host {

typedef unsigned __int128 u128;

struct Fr {
    uint64_t v[4];
    bool operator==(const Fr &o) const { return !memcmp(v, o.v, 32); }
};

static const uint64_t kP[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
static const uint64_t kPinv64 = 0xc2e1f593efffffffULL;   // -p^-1 mod 2^64

inline bool geq_p(const uint64_t *a) {
    for (int i = 3; i >= 0; i--) { if (a[i] != kP[i]) return a[i] > kP[i]; }
    return true;
}
inline uint64_t add4(uint64_t *o, const uint64_t *a, const uint64_t *b) {
    u128 c = 0; for (int i = 0; i < 4; i++) { c += (u128)a[i] + b[i]; o[i] = (uint64_t)c; c >>= 64; } return (uint64_t)c;
}
inline uint64_t sub4(uint64_t *o, const uint64_t *a, const uint64_t *b) {
    uint64_t br = 0; for (int i = 0; i < 4; i++) { u128 d = (u128)a[i] - b[i] - br; o[i] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1; } return br;
}
inline Fr montmul(const Fr &a, const Fr &b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) { c += (u128)a.v[j] * b.v[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * kPinv64;
        c = (u128)m * kP[0] + t[0]; c >>= 64;
        for (int j = 1; j < 4; j++) { c += (u128)m * kP[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
    }
    Fr r;
    if (t[4] || geq_p(t)) sub4(t, t, kP);
    memcpy(r.v, t, 32);
    return r;
}

struct Consts { Fr R, R2, one; };
inline const Consts &consts() {
    static Consts c = [] {
        Consts k; uint64_t r[4] = {1, 0, 0, 0};
        for (int i = 0; i < 512; i++) { uint64_t cy = add4(r, r, r); if (cy || geq_p(r)) sub4(r, r, kP); if (i == 255) memcpy(k.R.v, r, 32); }
        memcpy(k.R2.v, r, 32); k.one = Fr{{1, 0, 0, 0}}; return k;
    }();
    return c;
}
// canonical <-> Montgomery
inline Fr to_mont(const Fr &a) { return montmul(a, consts().R2); }
inline Fr from_mont(const Fr &a) { return montmul(a, consts().one); }
inline Fr mul(const Fr &a, const Fr &b) { return montmul(to_mont(a), b); }     // canonical * canonical
inline Fr add(const Fr &a, const Fr &b) { Fr r; add4(r.v, a.v, b.v); if (geq_p(r.v)) sub4(r.v, r.v, kP); return r; }
inline Fr pow(const Fr &a, uint64_t e) {
    Fr acc = consts().R, base = to_mont(a);
    while (e) { if (e & 1) acc = montmul(acc, base); base = montmul(base, base); e >>= 1; }
    return from_mont(acc);
}
inline Fr inv(const Fr &a) {                     // a^(p-2)
    uint64_t e[4]; const uint64_t two[4] = {2, 0, 0, 0}; sub4(e, kP, two);
    Fr acc = consts().R, base = to_mont(a);
    for (int i = 0; i < 256; i++) { if ((e[i >> 6] >> (i & 63)) & 1) acc = montmul(acc, base); base = montmul(base, base); }
    return from_mont(acc);
}
inline Fr from_u32(const uint32_t *l) { Fr r; memcpy(r.v, l, 32); return r; }
inline void to_u32(uint32_t *l, const Fr &a) { memcpy(l, a.v, 32); }
inline Fr from_u64(uint64_t x) { return Fr{{x, 0, 0, 0}}; }
inline bool is_canonical(const Fr &a) { return !geq_p(a.v); }

// table[j] = w^j * R mod p (Montgomery form), j < count
inline std::vector<Fr> power_table_mont(const Fr &w, size_t count) {
    std::vector<Fr> t(count);
    Fr wm = to_mont(w);
    if (count) t[0] = consts().R;
    for (size_t j = 1; j < count; j++) t[j] = montmul(t[j - 1], wm);
    return t;
}

}  // namespace host
}  // namespace lgr

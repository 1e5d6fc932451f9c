/*
 * oracle/oracle.c -- CPU ORACLE (test infrastructure; see lgo.h header for status and rules).
 *
 * Every function cites the reference file:line whose behaviour it restates.  Nothing here is
 * copied from the reference: the reference computes on the GPU in WGSL with 16-bit half
 * products, subtractive Montgomery and Barrett reduction; this file computes the same
 * canonical results with 4x64-bit CIOS Montgomery on the CPU.
 */
#include "lgo.h"
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#if defined(__x86_64__)
#include <immintrin.h>
#endif

typedef unsigned __int128 u128;

/* ------------------------------------------------------------------------------------------
 * Field constants.  p: src/bn254.cpp:21-22, shader/bn254fr.wgsl.in:19-22.
 * ------------------------------------------------------------------------------------------ */
static const uint64_t P[4] = { 0x43e1f593f0000001ULL, 0x2833e84879b97091ULL,
                               0xb85045b68181585dULL, 0x30644e72e131a029ULL };
static uint64_t PINV;        /* -p^-1 mod 2^64 */
static lgo_fr R1, R2;        /* 2^256 mod p, 2^512 mod p */
static lgo_fr ROOT1, ROOT2;  /* src/bn254.cpp:36-39 */

static inline int geq_p(const uint64_t a[4]) {
    for (int i = 3; i >= 0; i--) { if (a[i] > P[i]) return 1; if (a[i] < P[i]) return 0; }
    return 1;
}
static inline uint64_t add4(uint64_t o[4], const uint64_t a[4], const uint64_t b[4]) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) { c += (u128)a[i] + b[i]; o[i] = (uint64_t)c; c >>= 64; }
    return (uint64_t)c;
}
static inline uint64_t sub4(uint64_t o[4], const uint64_t a[4], const uint64_t b[4]) {
    uint64_t br = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a[i] - b[i] - br; o[i] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1;
    }
    return br;
}

void lgo_fr_modulus(lgo_fr *o) { memcpy(o->v, P, 32); }

/* kernels.wgsl.in:325-336 EltwiseAddMod: add then one conditional subtract (bn254fr_reduce, bn254fr.wgsl.in:50-58) */
void lgo_fr_add(lgo_fr *o, const lgo_fr *a, const lgo_fr *b) {
    uint64_t t[4]; add4(t, a->v, b->v);          /* a,b < p < 2^254: no carry out */
    if (geq_p(t)) sub4(t, t, P);
    memcpy(o->v, t, 32);
}
/* kernels.wgsl.in:364-380 EltwiseSubMod: subtract, add p back on borrow */
void lgo_fr_sub(lgo_fr *o, const lgo_fr *a, const lgo_fr *b) {
    uint64_t t[4];
    if (sub4(t, a->v, b->v)) add4(t, t, P);
    memcpy(o->v, t, 32);
}

/* a*b*2^-256 mod p, canonical: the value shader/bn254fr.wgsl.in:76-104 montgomery_mul returns.
 * (The WGSL uses hi - mulhi(lo*J, p) with J = p^-1; CIOS below yields the same residue.) */
static inline void montmul(uint64_t o[4], const uint64_t a[4], const uint64_t b[4]) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) { c += (u128)a[j] * b[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * PINV;
        c = (u128)m * P[0] + t[0]; c >>= 64;
        for (int j = 1; j < 4; j++) { c += (u128)m * P[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
    }
    if (t[4] || geq_p(t)) sub4(t, t, P);
    memcpy(o, t, 32);
}
void lgo_fr_montmul(lgo_fr *o, const lgo_fr *a, const lgo_fr *b) { montmul(o->v, a->v, b->v); }

/* canonical product: the value barrett_reduce_wide(bigint_mul_wide(a,b)) returns for canonical
 * inputs (shader/bn254fr.wgsl.in:113-124, kernels.wgsl.in:416-427) */
void lgo_fr_mul(lgo_fr *o, const lgo_fr *a, const lgo_fr *b) {
    uint64_t t[4]; montmul(t, a->v, R2.v); montmul(o->v, t, b->v);
}
static inline void to_mont(uint64_t o[4], const uint64_t a[4]) { montmul(o, a, R2.v); }
static inline void from_mont(uint64_t o[4], const uint64_t a[4]) {
    static const uint64_t one[4] = {1, 0, 0, 0}; montmul(o, a, one);
}
void lgo_fr_pow(lgo_fr *o, const lgo_fr *a, uint64_t e) {
    uint64_t acc[4], base[4]; memcpy(acc, R1.v, 32); to_mont(base, a->v);
    while (e) { if (e & 1) montmul(acc, acc, base); montmul(base, base, base); e >>= 1; }
    from_mont(o->v, acc);
}
/* Fermat inverse: the unique x with a*x = 1, what mpz_invert (engine.cpp:1435,1482) and the WGSL
 * extended Euclid (bn254fr.wgsl.in:128-151) both return for a != 0 */
void lgo_fr_inv(lgo_fr *o, const lgo_fr *a) {
    uint64_t e[4]; static const uint64_t two[4] = {2, 0, 0, 0}; sub4(e, P, two);
    uint64_t acc[4], base[4]; memcpy(acc, R1.v, 32); to_mont(base, a->v);
    for (int i = 0; i < 256; i++) {
        if ((e[i >> 6] >> (i & 63)) & 1) montmul(acc, acc, base);
        montmul(base, base, base);
    }
    from_mont(o->v, acc);
}

static void from_decimal(lgo_fr *o, const char *s) {
    uint64_t acc[4] = {0, 0, 0, 0}, ten[4] = {10, 0, 0, 0}, tenm[4];
    to_mont(tenm, ten);
    for (; *s; s++) {
        uint64_t d[4] = {(uint64_t)(*s - '0'), 0, 0, 0};
        montmul(acc, acc, tenm);                 /* acc*10 (acc canonical, tenm = 10R) */
        add4(acc, acc, d); if (geq_p(acc)) sub4(acc, acc, P);
    }
    memcpy(o->v, acc, 32);
}

__attribute__((constructor)) static void lgo_init(void) {
    uint64_t x = 1;                               /* Newton: x = p^-1 mod 2^64 */
    for (int i = 0; i < 6; i++) x *= 2 - P[0] * x;
    PINV = (uint64_t)0 - x;
    uint64_t r[4] = {1, 0, 0, 0};
    for (int i = 0; i < 512; i++) {               /* doubling mod p */
        uint64_t c = add4(r, r, r);
        if (c || geq_p(r)) sub4(r, r, P);
        if (i == 255) memcpy(R1.v, r, 32);
    }
    memcpy(R2.v, r, 32);
    from_decimal(&ROOT1, "1748695177688661943023146337482803886740723238769601073607632802312037301404");
    from_decimal(&ROOT2, "2037444462055058054189478067370099086220733342011840546702672064072905551290");
}
void lgo_root1(lgo_fr *o) { *o = ROOT1; }
void lgo_root2(lgo_fr *o) { *o = ROOT2; }

/* src/bn254.cpp:51-64 generate_omegas: w_k, w_2k from root1, w_n (n = 4k) from root2; both roots
 * have multiplicative order 2^28 (:41-43) */
void lgo_omegas(uint64_t k, lgo_fr *w_k, lgo_fr *w_2k, lgo_fr *w_n) {
    if (w_k)  lgo_fr_pow(w_k,  &ROOT1, (1ULL << 28) / k);
    if (w_2k) lgo_fr_pow(w_2k, &ROOT1, (1ULL << 28) / (2 * k));
    if (w_n)  lgo_fr_pow(w_n,  &ROOT2, (1ULL << 28) / (4 * k));
}

/* ------------------------------------------------------------------------------------------
 * Transforms.  The reference's pass structure (bit_reverse + shared + global radix-2 stages,
 * engine.cpp:844-882 forward DIF, :932-968 inverse DIT + N^-1 scale) computes the plain DFT
 *   forward:  X[j] = sum_i x[i] w^(ij)        inverse: x[i] = N^-1 sum_j X[j] w^(-ij)
 * natural order in and out, canonical outputs (final bn254fr_reduce / montgomery_mul by N^-1 R,
 * kernels.wgsl.in:92-102,203-216).  Restated as a textbook iterative Cooley-Tukey.
 * ------------------------------------------------------------------------------------------ */
static size_t ilog2(size_t n) { size_t l = 0; while (((size_t)1 << l) < n) l++; return l; }

static void ntt_mont(uint64_t (*a)[4], size_t N, const uint64_t (*tw)[4] /* N/2 powers, Montgomery */) {
    size_t L = ilog2(N);
    for (size_t i = 0; i < N; i++) {              /* bit reversal, kernels.wgsl.in:57-73 */
        size_t r = 0; for (size_t b = 0; b < L; b++) if (i >> b & 1) r |= (size_t)1 << (L - 1 - b);
        if (i < r) { uint64_t t[4]; memcpy(t, a[i], 32); memcpy(a[i], a[r], 32); memcpy(a[r], t, 32); }
    }
    for (size_t M2 = 1; M2 < N; M2 <<= 1) {
        size_t stride = N / (2 * M2);
        for (size_t g = 0; g < N; g += 2 * M2)
            for (size_t j = 0; j < M2; j++) {
                uint64_t *x = a[g + j], *y = a[g + j + M2], t[4], u[4];
                montmul(t, y, tw[j * stride]);
                add4(u, x, t); if (geq_p(u)) sub4(u, u, P);
                if (sub4(t, x, t)) add4(t, t, P);
                memcpy(x, u, 32); memcpy(y, t, 32);
            }
    }
}

static void make_twiddles(uint64_t (*tw)[4], size_t half, const lgo_fr *omega, int inverse) {
    lgo_fr w = *omega; if (inverse) lgo_fr_inv(&w, omega);
    uint64_t wm[4]; to_mont(wm, w.v);
    memcpy(tw[0], R1.v, 32);
    for (size_t i = 1; i < half; i++) montmul(tw[i], tw[i - 1], wm);
}

static void ntt_with_tw(lgo_fr *x, size_t N, const uint64_t (*tw)[4], int inverse) {
    uint64_t (*a)[4] = (uint64_t (*)[4])x;
    /* values stay canonical throughout, exactly as in the reference: only the twiddles carry the factor R
     * (engine.cpp:1397-1401), so montgomery_mul(y, wR) = y*w needs no conversion of the data */
    ntt_mont(a, N, tw);
    if (inverse) {                                /* kernels.wgsl.in:92-102 ntt_adjust_inverse_reduce: N^-1 * R (engine.cpp:1476-1482) */
        lgo_fr n = {{N, 0, 0, 0}}, ninv; lgo_fr_inv(&ninv, &n);
        uint64_t ninv_m[4]; to_mont(ninv_m, ninv.v);
        for (size_t i = 0; i < N; i++) montmul(a[i], a[i], ninv_m);
    }
}

void lgo_ntt(lgo_fr *x, size_t N, const lgo_fr *omega, int inverse) {
    if (N <= 1) return;
    uint64_t (*tw)[4] = malloc((N / 2) * 32);
    make_twiddles(tw, N / 2, omega, inverse);
    ntt_with_tw(x, N, tw, inverse);
    free(tw);
}

void lgo_dft_naive(lgo_fr *out, const lgo_fr *x, size_t N, const lgo_fr *omega, int inverse) {
    lgo_fr w = *omega; if (inverse) lgo_fr_inv(&w, omega);
    lgo_fr *pw = malloc(N * sizeof(lgo_fr));
    pw[0] = (lgo_fr){{1, 0, 0, 0}};
    for (size_t i = 1; i < N; i++) lgo_fr_mul(&pw[i], &pw[i - 1], &w);
    lgo_fr ninv = {{1, 0, 0, 0}};
    if (inverse) { lgo_fr n = {{N, 0, 0, 0}}; lgo_fr_inv(&ninv, &n); }
    for (size_t j = 0; j < N; j++) {
        lgo_fr acc = {{0, 0, 0, 0}}, t;
        for (size_t i = 0; i < N; i++) { lgo_fr_mul(&t, &x[i], &pw[(i * j) % N]); lgo_fr_add(&acc, &acc, &t); }
        lgo_fr_mul(&out[j], &acc, &ninv);
    }
    free(pw);
}

/* per-(k) cached twiddles for the row encoder */
typedef struct { size_t k; uint64_t (*inv_k)[4]; uint64_t (*fwd_n)[4]; uint64_t (*inv_2k)[4];
                 uint64_t (*inv_n)[4]; uint64_t (*fwd_k)[4]; } enc_tables;
static enc_tables ET;
static void enc_tables_get(size_t k) {
    if (ET.k == k) return;
#ifdef _OPENMP
#pragma omp critical(lgo_tables)
#endif
    if (ET.k != k) {
        free(ET.inv_k); free(ET.fwd_n); free(ET.inv_2k); free(ET.inv_n); free(ET.fwd_k);
        lgo_fr wk, w2k, wn; lgo_omegas(k, &wk, &w2k, &wn);
        ET.inv_k = malloc((k / 2) * 32);  make_twiddles(ET.inv_k, k / 2, &wk, 1);
        ET.fwd_k = malloc((k / 2) * 32);  make_twiddles(ET.fwd_k, k / 2, &wk, 0);
        ET.inv_2k = malloc(k * 32);       make_twiddles(ET.inv_2k, k, &w2k, 1);
        ET.fwd_n = malloc(2 * k * 32);    make_twiddles(ET.fwd_n, 2 * k, &wn, 0);
        ET.inv_n = malloc(2 * k * 32);    make_twiddles(ET.inv_n, 2 * k, &wn, 1);
        ET.k = k;
    }
}

/* engine.cpp:755-770: ntt_inverse_kernel(N=k, w_k) on buf[0:k), then ntt_forward_kernel(N=n, w_n) */
void lgo_encode(lgo_fr *buf, size_t k) {
    enc_tables_get(k);
    ntt_with_tw(buf, k, ET.inv_k, 1);
    ntt_with_tw(buf, 4 * k, ET.fwd_n, 0);
}
/* nonbatch_context.hpp:482-494 (mask rows): ntt_inverse_2k then ntt_forward_n */
void lgo_encode_2k(lgo_fr *buf, size_t k) {
    enc_tables_get(k);
    ntt_with_tw(buf, 2 * k, ET.inv_2k, 1);
    ntt_with_tw(buf, 4 * k, ET.fwd_n, 0);
}
/* engine.cpp:772-796: iNTT_n; fold c[i] += c[i+k] for i<k (ntt_fold reads params of the 2k config,
 * kernels.wgsl.in:104-116: half = 2k>>1 = k); NTT_k on buf[0:k); buf[k:n) keeps the coefficients */
void lgo_decode(lgo_fr *buf, size_t k) {
    enc_tables_get(k);
    ntt_with_tw(buf, 4 * k, ET.inv_n, 1);
    for (size_t i = 0; i < k; i++) lgo_fr_add(&buf[i], &buf[i], &buf[i + k]);
    ntt_with_tw(buf, k, ET.fwd_k, 0);
}

/* ------------------------------------------------------------------------------------------
 * Element-wise kernels, shader/kernels.wgsl.in:325-549 (canonical in, canonical out)
 * ------------------------------------------------------------------------------------------ */
void lgo_elt_add(lgo_fr *o, const lgo_fr *x, const lgo_fr *y, size_t n) { for (size_t i = 0; i < n; i++) lgo_fr_add(&o[i], &x[i], &y[i]); }
void lgo_elt_sub(lgo_fr *o, const lgo_fr *x, const lgo_fr *y, size_t n) { for (size_t i = 0; i < n; i++) lgo_fr_sub(&o[i], &x[i], &y[i]); }
void lgo_elt_mul(lgo_fr *o, const lgo_fr *x, const lgo_fr *y, size_t n) { for (size_t i = 0; i < n; i++) lgo_fr_mul(&o[i], &x[i], &y[i]); }
void lgo_elt_div(lgo_fr *o, const lgo_fr *x, const lgo_fr *y, size_t n) {
    for (size_t i = 0; i < n; i++) { lgo_fr t; lgo_fr_inv(&t, &y[i]); lgo_fr_mul(&o[i], &x[i], &t); }
}
void lgo_elt_fma(lgo_fr *o, const lgo_fr *x, const lgo_fr *y, size_t n) {
    for (size_t i = 0; i < n; i++) { lgo_fr t; lgo_fr_mul(&t, &x[i], &y[i]); lgo_fr_add(&o[i], &o[i], &t); }
}
void lgo_elt_fma_const(lgo_fr *o, const lgo_fr *x, const lgo_fr *c, size_t n) {
    for (size_t i = 0; i < n; i++) { lgo_fr t; lgo_fr_mul(&t, &x[i], c); lgo_fr_add(&o[i], &o[i], &t); }
}
void lgo_elt_add_assign(lgo_fr *o, const lgo_fr *x, size_t n) { for (size_t i = 0; i < n; i++) lgo_fr_add(&o[i], &o[i], &x[i]); }
void lgo_elt_add_const(lgo_fr *o, const lgo_fr *x, const lgo_fr *c, size_t n) { for (size_t i = 0; i < n; i++) lgo_fr_add(&o[i], &x[i], c); }
void lgo_elt_sub_const(lgo_fr *o, const lgo_fr *x, const lgo_fr *c, size_t n) { for (size_t i = 0; i < n; i++) lgo_fr_sub(&o[i], &x[i], c); }
void lgo_elt_const_sub(lgo_fr *o, const lgo_fr *x, const lgo_fr *c, size_t n) { for (size_t i = 0; i < n; i++) lgo_fr_sub(&o[i], c, &x[i]); }
void lgo_elt_mul_const(lgo_fr *o, const lgo_fr *x, const lgo_fr *c, size_t n) { for (size_t i = 0; i < n; i++) lgo_fr_mul(&o[i], &x[i], c); }
void lgo_elt_montmul_const(lgo_fr *o, const lgo_fr *x, const lgo_fr *c, size_t n) { for (size_t i = 0; i < n; i++) montmul(o[i].v, x[i].v, c->v); }
/* kernels.wgsl.in:501-510 EltwiseBitDecompose */
void lgo_elt_bit(lgo_fr *o, const lgo_fr *x, uint32_t bit, size_t n) {
    for (size_t i = 0; i < n; i++) { uint64_t b = (x[i].v[(bit >> 6) & 3] >> (bit & 63)) & 1; o[i] = (lgo_fr){{b, 0, 0, 0}}; }
}
/* kernels.wgsl.in:512-537: coeff * base^exp (the Montgomery table of powmod_context.cpp:141-197
 * cancels out: result is the plain field value; tests/webgpu/test_powmod.cpp checks exactly that) */
void lgo_elt_powmod(lgo_fr *o, const lgo_fr *coeff, const uint32_t *exp, const lgo_fr *base, size_t n, int add) {
    for (size_t i = 0; i < n; i++) {
        lgo_fr t; lgo_fr_pow(&t, base, exp[i]); lgo_fr_mul(&t, &t, &coeff[i]);
        if (add) lgo_fr_add(&o[i], &o[i], &t); else o[i] = t;
    }
}
void lgo_gather(lgo_fr *o, const lgo_fr *x, const uint32_t *idx, size_t ns) { for (size_t i = 0; i < ns; i++) o[i] = x[idx[i]]; }

/* ------------------------------------------------------------------------------------------
 * SHA-256 (FIPS 180-4).  compress() takes the 16 message words already in host order.
 * ------------------------------------------------------------------------------------------ */
static const uint32_t K256[64] = {
    0x428a2f98,0x71374491,0xb5c0fbcf,0xe9b5dba5,0x3956c25b,0x59f111f1,0x923f82a4,0xab1c5ed5,
    0xd807aa98,0x12835b01,0x243185be,0x550c7dc3,0x72be5d74,0x80deb1fe,0x9bdc06a7,0xc19bf174,
    0xe49b69c1,0xefbe4786,0x0fc19dc6,0x240ca1cc,0x2de92c6f,0x4a7484aa,0x5cb0a9dc,0x76f988da,
    0x983e5152,0xa831c66d,0xb00327c8,0xbf597fc7,0xc6e00bf3,0xd5a79147,0x06ca6351,0x14292967,
    0x27b70a85,0x2e1b2138,0x4d2c6dfc,0x53380d13,0x650a7354,0x766a0abb,0x81c2c92e,0x92722c85,
    0xa2bfe8a1,0xa81a664b,0xc24b8b70,0xc76c51a3,0xd192e819,0xd6990624,0xf40e3585,0x106aa070,
    0x19a4c116,0x1e376c08,0x2748774c,0x34b0bcb5,0x391c0cb3,0x4ed8aa4a,0x5b9cca4f,0x682e6ff3,
    0x748f82ee,0x78a5636f,0x84c87814,0x8cc70208,0x90befffa,0xa4506ceb,0xbef9a3f7,0xc67178f2 };
static const uint32_t IV256[8] = { 0x6a09e667,0xbb67ae85,0x3c6ef372,0xa54ff53a,0x510e527f,0x9b05688c,0x1f83d9ab,0x5be0cd19 };

#define ROR(x, n) (((x) >> (n)) | ((x) << (32 - (n))))
static void compress_portable(uint32_t st[8], const uint32_t w16[16]) {
    uint32_t w[64], a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
    for (int i = 0; i < 16; i++) w[i] = w16[i];
    for (int i = 16; i < 64; i++) {
        uint32_t s0 = ROR(w[i-15], 7) ^ ROR(w[i-15], 18) ^ (w[i-15] >> 3);
        uint32_t s1 = ROR(w[i-2], 17) ^ ROR(w[i-2], 19) ^ (w[i-2] >> 10);
        w[i] = w[i-16] + s0 + w[i-7] + s1;
    }
    for (int i = 0; i < 64; i++) {
        uint32_t t1 = h + (ROR(e, 6) ^ ROR(e, 11) ^ ROR(e, 25)) + ((e & f) ^ (~e & g)) + K256[i] + w[i];
        uint32_t t2 = (ROR(a, 2) ^ ROR(a, 13) ^ ROR(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
        h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}

#if defined(__x86_64__)
__attribute__((target("sha,sse4.1,ssse3")))
static void compress_shani(uint32_t st[8], const uint32_t w16[16]) {
    __m128i tmp = _mm_loadu_si128((const __m128i *)&st[0]);      /* a b c d (low..high) */
    __m128i s1 = _mm_loadu_si128((const __m128i *)&st[4]);       /* e f g h */
    tmp = _mm_shuffle_epi32(tmp, 0xB1);                           /* b a d c */
    s1 = _mm_shuffle_epi32(s1, 0x1B);                             /* h g f e */
    __m128i s0 = _mm_alignr_epi8(tmp, s1, 8);                     /* ABEF */
    s1 = _mm_blend_epi16(s1, tmp, 0xF0);                          /* CDGH */
    const __m128i abef_save = s0, cdgh_save = s1;
    __m128i m[4];
    for (int i = 0; i < 4; i++) m[i] = _mm_loadu_si128((const __m128i *)&w16[4 * i]);   /* words already native */
    __m128i msg;
    for (int r = 0; r < 16; r++) {
        __m128i cur = m[r & 3];
        msg = _mm_add_epi32(cur, _mm_loadu_si128((const __m128i *)&K256[4 * r]));
        s1 = _mm_sha256rnds2_epu32(s1, s0, msg);
        msg = _mm_shuffle_epi32(msg, 0x0E);
        s0 = _mm_sha256rnds2_epu32(s0, s1, msg);
        if (r < 12) {                                             /* schedule w[4(r+4) .. 4(r+4)+3] */
            __m128i w0 = m[r & 3], w1 = m[(r + 1) & 3], w2 = m[(r + 2) & 3], w3 = m[(r + 3) & 3];
            __m128i t = _mm_sha256msg1_epu32(w0, w1);
            t = _mm_add_epi32(t, _mm_alignr_epi8(w3, w2, 4));
            m[r & 3] = _mm_sha256msg2_epu32(t, w3);
        }
    }
    s0 = _mm_add_epi32(s0, abef_save);
    s1 = _mm_add_epi32(s1, cdgh_save);
    tmp = _mm_shuffle_epi32(s0, 0x1B);                            /* FEBA */
    s1 = _mm_shuffle_epi32(s1, 0xB1);                             /* DCHG */
    s0 = _mm_blend_epi16(tmp, s1, 0xF0);                          /* DCBA */
    s1 = _mm_alignr_epi8(s1, tmp, 8);                             /* HGFE */
    _mm_storeu_si128((__m128i *)&st[0], s0);
    _mm_storeu_si128((__m128i *)&st[4], s1);
}
#endif

static void (*compress)(uint32_t st[8], const uint32_t w16[16]) = compress_portable;
__attribute__((constructor)) static void sha_dispatch(void) {
#if defined(__x86_64__)
    __builtin_cpu_init();
    if (__builtin_cpu_supports("sha") && __builtin_cpu_supports("sse4.1") && !getenv("LGO_NO_SHANI"))
        compress = compress_shani;
#endif
}

void lgo_sha256(uint8_t out[32], const uint8_t *msg, size_t len) {
    uint32_t st[8], w[16]; memcpy(st, IV256, 32);
    size_t off = 0;
    for (; off + 64 <= len; off += 64) {
        for (int i = 0; i < 16; i++) w[i] = (uint32_t)msg[off+4*i] << 24 | (uint32_t)msg[off+4*i+1] << 16 | (uint32_t)msg[off+4*i+2] << 8 | msg[off+4*i+3];
        compress(st, w);
    }
    uint8_t blk[128]; size_t rem = len - off; memset(blk, 0, 128); memcpy(blk, msg + off, rem);
    blk[rem] = 0x80; size_t tot = rem < 56 ? 64 : 128; uint64_t bits = (uint64_t)len * 8;
    for (int i = 0; i < 8; i++) blk[tot - 1 - i] = (uint8_t)(bits >> (8 * i));
    for (size_t o = 0; o < tot; o += 64) {
        for (int i = 0; i < 16; i++) w[i] = (uint32_t)blk[o+4*i] << 24 | (uint32_t)blk[o+4*i+1] << 16 | (uint32_t)blk[o+4*i+2] << 8 | blk[o+4*i+3];
        compress(st, w);
    }
    for (int i = 0; i < 8; i++) { out[4*i] = st[i] >> 24; out[4*i+1] = st[i] >> 16; out[4*i+2] = st[i] >> 8; out[4*i+3] = st[i]; }
}

/* ---- batched column contexts, shader/sha256.wgsl:23-28,127-230 ----
 * sha256_update (:147-177) appends, for limb = 0..7, the bytes (val>>24, val>>16, val>>8, val):
 * so the u32 limb IS the big-endian message word.  Even rows fill W[0..7], odd rows W[8..15],
 * one compression every second row. */
struct lgo_sha { size_t n; uint64_t rows; uint32_t *state; uint32_t *pend; };
lgo_sha *lgo_sha_new(size_t n) {
    lgo_sha *s = calloc(1, sizeof *s); s->n = n;
    s->state = malloc(n * 32); s->pend = malloc(n * 32); lgo_sha_init(s); return s;
}
void lgo_sha_free(lgo_sha *s) { if (s) { free(s->state); free(s->pend); free(s); } }
void lgo_sha_init(lgo_sha *s) { s->rows = 0; for (size_t j = 0; j < s->n; j++) memcpy(&s->state[8 * j], IV256, 32); }
void lgo_sha_update(lgo_sha *s, const lgo_fr *row) {
    const uint32_t *in = (const uint32_t *)row;
    if ((s->rows & 1) == 0) memcpy(s->pend, in, s->n * 32);
    else {
#ifdef _OPENMP
#pragma omp parallel for schedule(static) if (s->n >= 4096)
#endif
        for (size_t j = 0; j < s->n; j++) {
            uint32_t w[16]; memcpy(w, &s->pend[8 * j], 32); memcpy(w + 8, &in[8 * j], 32);
            compress(&s->state[8 * j], w);
        }
    }
    s->rows++;
}
/* sha256_final (:179-230): 0x80 pad, 64-bit big-endian bit length, digest = state words as u32 */
static void sha_final_one(uint32_t st[8], const uint32_t *pend, uint64_t rows) {
    uint32_t w[16]; memset(w, 0, 64);
    uint64_t bits = rows * 256; int used = 0;
    if (rows & 1) { memcpy(w, pend, 32); used = 8; }
    w[used] = 0x80000000u; w[14] = (uint32_t)(bits >> 32); w[15] = (uint32_t)bits;
    compress(st, w);
}
void lgo_sha_final(lgo_sha *s, uint8_t *digests) {
    for (size_t j = 0; j < s->n; j++) {
        uint32_t st[8]; memcpy(st, &s->state[8 * j], 32);
        sha_final_one(st, &s->pend[8 * j], s->rows);
        memcpy(digests + 32 * j, st, 32);         /* native LE u32 words, sha256.wgsl:226-228 */
    }
}

/* ---- Merkle tree, include/zkp/merkle_tree.hpp:343-375 ---- */
static size_t bit_ceil(size_t n) { size_t p = 1; while (p < n) p <<= 1; return p; }
size_t lgo_merkle_nodes(size_t nleaves) { return 2 * bit_ceil(nleaves) - 1; }
void lgo_merkle_build(uint8_t *nodes, const uint8_t *leaf, size_t nleaves) {
    size_t Pn = bit_ceil(nleaves), parent = Pn - 1;
    memset(nodes, 0, (2 * Pn - 1) * 32);
    memcpy(nodes + parent * 32, leaf, nleaves * 32);
    for (size_t i = parent; i-- > 0;)             /* children always have larger indices */
        lgo_sha256(nodes + 32 * i, nodes + 32 * (2 * i + 1), 64);
}

/* ---- synthetic matrix ---- */
static inline uint64_t splitmix(uint64_t *s) {
    uint64_t z = (*s += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL; z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL; return z ^ (z >> 31);
}
static inline void synth_one(lgo_fr *o, uint64_t seed, uint64_t row, uint64_t col) {
    uint64_t s = seed * 0xd1342543de82ef95ULL + row * 0x2545f4914f6cdd1dULL + col * 0x9e3779b97f4a7c15ULL + 0x632be59bd9b4e019ULL;
    uint64_t w[4]; for (int i = 0; i < 4; i++) w[i] = splitmix(&s);
    /* finite_field_gmp.hpp:70-78 generate_random: 256 random bits, >> 2, one conditional subtract */
    for (int i = 0; i < 3; i++) w[i] = (w[i] >> 2) | (w[i + 1] << 62);
    w[3] >>= 2;
    if (geq_p(w)) sub4(w, w, P);
    memcpy(o->v, w, 32);
}
void lgo_synth(lgo_fr *out, uint64_t seed, uint64_t row0, uint64_t nrows, uint64_t ncols) {
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (uint64_t r = 0; r < nrows; r++)
        for (uint64_t c = 0; c < ncols; c++) synth_one(&out[r * ncols + c], seed, row0 + r, c);
}

void lgo_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int lgo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---- whole path (CPU baseline): tile of T rows encoded in parallel over rows, then every
 * column absorbs the tile's rows in order, parallel over column blocks ---- */
static int encode_commit_impl(const lgo_fr *rows, uint64_t seed, size_t R, size_t k, uint8_t *digests, uint8_t *nodes) {
    size_t n = 4 * k;
    enc_tables_get(k);
    size_t T = (size_t)1 << 22 >> ilog2(n);       /* ~4 Mi codeword elements = 128 MiB per tile */
    if (T < 2) T = 2;
    if (T & 1) T++;
    if (T > R) T = R;
    lgo_fr *tile = malloc(T * n * sizeof(lgo_fr));
    uint32_t *state = malloc(n * 32), *pend = malloc(n * 32);
    for (size_t j = 0; j < n; j++) memcpy(&state[8 * j], IV256, 32);
    for (size_t r0 = 0; r0 < R; r0 += T) {
        size_t t = R - r0 < T ? R - r0 : T;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 4)
#endif
        for (size_t r = 0; r < t; r++) {
            lgo_fr *buf = tile + r * n;
            if (rows) memcpy(buf, rows + (r0 + r) * k, k * 32);
            else for (size_t c = 0; c < k; c++) synth_one(&buf[c], seed, r0 + r, c);
            memset(buf + k, 0, (n - k) * 32);
            ntt_with_tw(buf, k, ET.inv_k, 1);
            ntt_with_tw(buf, n, ET.fwd_n, 0);
        }
        /* column block: at least two blocks per thread so that narrow matrices (n = 1024) keep every core busy */
        size_t CB = n / (2 * (size_t)lgo_num_threads());
        while (CB & (CB - 1)) CB &= CB - 1;        /* power of two: divides n */
        if (CB > 64) CB = 64;
        if (CB < 8) CB = 8;
        if (CB > n) CB = n;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
        for (size_t cb = 0; cb < n; cb += CB) {
            for (size_t r = 0; r < t; r++) {
                const uint32_t *in = (const uint32_t *)(tile + r * n + cb);
                if (((r0 + r) & 1) == 0) memcpy(&pend[8 * cb], in, CB * 32);
                else for (size_t j = 0; j < CB; j++) {
                    uint32_t w[16]; memcpy(w, &pend[8 * (cb + j)], 32); memcpy(w + 8, &in[8 * j], 32);
                    compress(&state[8 * (cb + j)], w);
                }
            }
        }
    }
    for (size_t j = 0; j < n; j++) { sha_final_one(&state[8 * j], &pend[8 * j], R); memcpy(digests + 32 * j, &state[8 * j], 32); }
    if (nodes) lgo_merkle_build(nodes, digests, n);
    free(tile); free(state); free(pend);
    return lgo_num_threads();
}
int lgo_encode_commit(const lgo_fr *rows, size_t R, size_t k, uint8_t *digests, uint8_t *nodes) {
    return encode_commit_impl(rows, 0, R, k, digests, nodes);
}
int lgo_encode_commit_synth(uint64_t seed, size_t R, size_t k, uint8_t *digests, uint8_t *nodes) {
    return encode_commit_impl(NULL, seed, R, k, digests, nodes);
}
int lgo_ntt_batch(lgo_fr *x, size_t N, size_t batch, const lgo_fr *omega, int inverse) {
    uint64_t (*tw)[4] = malloc((N / 2) * 32);
    make_twiddles(tw, N / 2, omega, inverse);
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (size_t b = 0; b < batch; b++) ntt_with_tw(x + b * N, N, tw, inverse);
    free(tw);
    return lgo_num_threads();
}

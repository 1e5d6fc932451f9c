// BN254 scalar field (Fr) arithmetic for sm_100a: 8 x u32 limbs held entirely in registers.
//
// Replaces shader/bigint.wgsl.in + shader/bn254fr.wgsl.in of the reference (256-bit integers as
// 2 x vec4<u32>, products built from 16-bit halves because WGSL has no mulhi,
// bigint.wgsl.in:69-88).  Here one 32x32->64 product is one IMAD.WIDE; a Montgomery
// multiplication is an interleaved 8-row CIOS on two 64-bit-aligned accumulator sets ("even"
// columns and "odd" columns) so that every partial product is a wide multiply-add on a carry
// chain (mad.lo.cc / madc.hi.cc pairs).
//
// Value conventions
//   * host-visible elements: canonical [0,p), NOT Montgomery form (bn254fr.wgsl.in)
//   * twiddles / constants: value * 2^256 mod p ("Montgomery form"), canonical
//   * fr_mont_mul(a, b): a in [0,4p), b in [0,p)  ->  a*b*2^-256 mod p in [0,2p)
//     (same residue as montgomery_mul_2p, bn254fr.wgsl.in:106-109)
//
// The file also compiles for the host (LGR_FR_HOST_EMU) with the PTX carry primitives emulated,
// so the algorithm can be checked against the oracle without a GPU (tests/cpp/fr_emu.cpp, tests/test_host_cpu.py).
#pragma once
#include <stdint.h>

#ifdef LGR_FR_HOST_EMU
#define LGR_DEV static inline
namespace lgr_emu {
static thread_local uint32_t cc;
static inline uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a + b; cc = (uint32_t)(s >> 32); return (uint32_t)s; }
static inline uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a + b + cc; cc = (uint32_t)(s >> 32); return (uint32_t)s; }
static inline uint32_t addc(uint32_t a, uint32_t b) { return a + b + cc; }
// PTX semantics: after sub.cc / subc.cc the flag holds the BORROW; subc computes a - (b + CF)
static inline uint32_t sub_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a - b; cc = (uint32_t)((s >> 32) & 1); return (uint32_t)s; }
static inline uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a - b - cc; cc = (uint32_t)((s >> 32) & 1); return (uint32_t)s; }
static inline uint32_t subc(uint32_t a, uint32_t b) { return a - b - cc; }
static inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t s = (uint64_t)(uint32_t)((uint64_t)a * b) + c; cc = (uint32_t)(s >> 32); return (uint32_t)s; }
static inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t s = (uint64_t)(uint32_t)((uint64_t)a * b) + c + cc; cc = (uint32_t)(s >> 32); return (uint32_t)s; }
static inline uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t s = (((uint64_t)a * b) >> 32) + c; cc = (uint32_t)(s >> 32); return (uint32_t)s; }
static inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t s = (((uint64_t)a * b) >> 32) + c + cc; cc = (uint32_t)(s >> 32); return (uint32_t)s; }
static inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return (uint32_t)((((uint64_t)a * b) >> 32) + c + cc); }
static inline uint32_t mul_lo(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a * b); }
static inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
}  // namespace lgr_emu
using namespace lgr_emu;
#else
#define LGR_DEV __device__ __forceinline__
// PTX carry-flag primitives.  The condition code is not a modelled register: the statements are
// volatile so the compiler keeps them in program order, and nothing else we emit between them
// writes CC (plain add/sub/mad do not).
LGR_DEV uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
LGR_DEV uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
LGR_DEV uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
LGR_DEV uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
LGR_DEV uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
LGR_DEV uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
LGR_DEV uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
LGR_DEV uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
LGR_DEV uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
LGR_DEV uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
LGR_DEV uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
LGR_DEV uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
LGR_DEV uint32_t mul_hi(uint32_t a, uint32_t b) { return __umulhi(a, b); }
#endif

namespace lgr {

struct fr_t { uint32_t v[8]; };

// p  (src/bn254.cpp:21-22, shader/bn254fr.wgsl.in:19-22), little-endian u32 limbs
#define LGR_P0 0xF0000001u
#define LGR_P1 0x43E1F593u
#define LGR_P2 0x79B97091u
#define LGR_P3 0x2833E848u
#define LGR_P4 0x8181585Du
#define LGR_P5 0xB85045B6u
#define LGR_P6 0xE131A029u
#define LGR_P7 0x30644E72u
// -p^-1 mod 2^32
#define LGR_M0 0xEFFFFFFFu

LGR_DEV uint32_t fr_p(int i) {
    switch (i) { case 0: return LGR_P0; case 1: return LGR_P1; case 2: return LGR_P2; case 3: return LGR_P3;
                 case 4: return LGR_P4; case 5: return LGR_P5; case 6: return LGR_P6; default: return LGR_P7; }
}
// 2p limbs (shader/bn254fr.wgsl.in:25-28)
LGR_DEV uint32_t fr_2p(int i) {
    switch (i) { case 0: return 0xE0000002u; case 1: return 0x87C3EB27u; case 2: return 0xF372E122u; case 3: return 0x5067D090u;
                 case 4: return 0x0302B0BAu; case 5: return 0x70A08B6Du; case 6: return 0xC2634053u; default: return 0x60C89CE5u; }
}

LGR_DEV fr_t fr_zero() { fr_t r; for (int i = 0; i < 8; i++) r.v[i] = 0; return r; }

// r = a + b (no reduction; caller guarantees < 2^256)
LGR_DEV fr_t fr_add_raw(const fr_t &a, const fr_t &b) {
    fr_t r;
    r.v[0] = add_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < 7; i++) r.v[i] = addc_cc(a.v[i], b.v[i]);
    r.v[7] = addc(a.v[7], b.v[7]);
    return r;
}

// if (a >= m) a -= m, with m = p (TWO=false) or 2p (TWO=true); bn254fr_reduce / bn254fr_reduce_2p
// (shader/bn254fr.wgsl.in:50-68).  Branch-free select on the final borrow.
template <bool TWO>
LGR_DEV fr_t fr_csub(const fr_t &a) {
    fr_t d;
    d.v[0] = sub_cc(a.v[0], TWO ? fr_2p(0) : fr_p(0));
#pragma unroll
    for (int i = 1; i < 8; i++) d.v[i] = subc_cc(a.v[i], TWO ? fr_2p(i) : fr_p(i));
    uint32_t borrow = subc(0, 0);        // all ones iff a < m
    fr_t r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = borrow ? a.v[i] : d.v[i];
    return r;
}
LGR_DEV fr_t fr_reduce_p(const fr_t &a) { return fr_csub<false>(a); }
LGR_DEV fr_t fr_reduce_2p(const fr_t &a) { return fr_csub<true>(a); }
// [0,4p) -> [0,p)
LGR_DEV fr_t fr_canon4(const fr_t &a) { return fr_reduce_p(fr_reduce_2p(a)); }

// a + b mod p for canonical inputs (EltwiseAddMod, kernels.wgsl.in:325-336)
LGR_DEV fr_t fr_add(const fr_t &a, const fr_t &b) { return fr_reduce_p(fr_add_raw(a, b)); }

// a - b mod p for canonical inputs (EltwiseSubMod, kernels.wgsl.in:364-380): add p back on borrow
LGR_DEV fr_t fr_sub(const fr_t &a, const fr_t &b) {
    fr_t d;
    d.v[0] = sub_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) d.v[i] = subc_cc(a.v[i], b.v[i]);
    uint32_t mask = subc(0, 0);          // all ones iff a < b
    fr_t r;
    r.v[0] = add_cc(d.v[0], fr_p(0) & mask);
#pragma unroll
    for (int i = 1; i < 7; i++) r.v[i] = addc_cc(d.v[i], fr_p(i) & mask);
    r.v[7] = addc(d.v[7], fr_p(7) & mask);
    return r;
}

// lazy butterfly helpers: inputs in [0,2p)
//   sum  = a + b           reduced to [0,2p)
//   diff = a - b + 2p      in (0,4p)   (fed straight into fr_mont_mul, which accepts [0,4p))
LGR_DEV fr_t fr_add_lazy(const fr_t &a, const fr_t &b) { return fr_reduce_2p(fr_add_raw(a, b)); }
LGR_DEV fr_t fr_sub_lazy4(const fr_t &a, const fr_t &b) {
    fr_t t;
    t.v[0] = add_cc(a.v[0], fr_2p(0));
#pragma unroll
    for (int i = 1; i < 7; i++) t.v[i] = addc_cc(a.v[i], fr_2p(i));
    t.v[7] = addc(a.v[7], fr_2p(7));
    fr_t r;
    r.v[0] = sub_cc(t.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < 7; i++) r.v[i] = subc_cc(t.v[i], b.v[i]);
    r.v[7] = subc(t.v[7], b.v[7]);
    return r;
}
// a - b in [0,2p) for a,b in [0,2p)
LGR_DEV fr_t fr_sub_lazy(const fr_t &a, const fr_t &b) { return fr_reduce_2p(fr_sub_lazy4(a, b)); }

// ---- Montgomery multiplication -------------------------------------------------------------
// T = even + 2^32 * odd.  Row i adds a*b[i] (even-indexed limbs of a into `even`, odd-indexed
// into `odd`), then m = T[0] * M0 and adds m*p the same way, which zeroes limb 0; the shift by one
// limb is a swap of the roles of the two sets.
namespace detail {
LGR_DEV void mul_n(uint32_t *acc, const uint32_t *a, uint32_t bi) {
#pragma unroll
    for (int j = 0; j < 8; j += 2) { acc[j] = mul_lo(a[j], bi); acc[j + 1] = mul_hi(a[j], bi); }
}
LGR_DEV void cmad_n(uint32_t *acc, const uint32_t *a, uint32_t bi) {
    acc[0] = mad_lo_cc(a[0], bi, acc[0]);
    acc[1] = madc_hi_cc(a[0], bi, acc[1]);
#pragma unroll
    for (int j = 2; j < 8; j += 2) { acc[j] = madc_lo_cc(a[j], bi, acc[j]); acc[j + 1] = madc_hi_cc(a[j], bi, acc[j + 1]); }
}
LGR_DEV void cmad_p(uint32_t *acc, int first, uint32_t mi) {      // acc += p[first, first+2, ..] * mi
    acc[0] = mad_lo_cc(fr_p(first), mi, acc[0]);
    acc[1] = madc_hi_cc(fr_p(first), mi, acc[1]);
#pragma unroll
    for (int j = 2; j < 8; j += 2) { acc[j] = madc_lo_cc(fr_p(first + j), mi, acc[j]); acc[j + 1] = madc_hi_cc(fr_p(first + j), mi, acc[j + 1]); }
}
LGR_DEV void madc_n_rshift(uint32_t *odd, const uint32_t *a, uint32_t bi) {
#pragma unroll
    for (int j = 0; j < 6; j += 2) { odd[j] = madc_lo_cc(a[j], bi, odd[j + 2]); odd[j + 1] = madc_hi_cc(a[j], bi, odd[j + 3]); }
    odd[6] = madc_lo_cc(a[6], bi, 0);
    odd[7] = madc_hi(a[6], bi, 0);
}
LGR_DEV void mad_row(uint32_t *even, uint32_t *odd, const uint32_t *a, uint32_t bi, bool first) {
    if (first) {
        mul_n(odd, a + 1, bi);
        mul_n(even, a, bi);
    } else {
        even[0] = add_cc(even[0], odd[1]);
        madc_n_rshift(odd, a + 1, bi);
        cmad_n(even, a, bi);
        odd[7] = addc(odd[7], 0);
    }
    uint32_t mi = mul_lo(even[0], LGR_M0);
    cmad_p(odd, 1, mi);
    cmad_p(even, 0, mi);
    odd[7] = addc(odd[7], 0);
}
}  // namespace detail

// a in [0,4p), b in [0,p)  ->  a*b*2^-256 mod p, in [0,2p)
// (kept inline: a real call passes the operands through a 352-byte local stack frame -- measured, slower)
LGR_DEV fr_t fr_mont_mul(const fr_t &a, const fr_t &b) {
    uint32_t even[8], odd[8];
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
        detail::mad_row(even, odd, a.v, b.v[i], i == 0);
        detail::mad_row(odd, even, a.v, b.v[i + 1], false);
    }
    fr_t r;
    r.v[0] = add_cc(even[0], odd[1]);
#pragma unroll
    for (int i = 1; i < 7; i++) r.v[i] = addc_cc(even[i], odd[i + 1]);
    r.v[7] = addc(even[7], 0);
    return r;
}

// canonical result
LGR_DEV fr_t fr_mont_mul_canon(const fr_t &a, const fr_t &b) { return fr_canon4(fr_mont_mul(a, b)); }

// ---- multiplication by a table constant (Shoup) --------------------------------------------------
// For a fixed w < p with the precomputed quotient wq = floor(w * 2^256 / p):
//     q = floor(x * wq / 2^256),   r = x*w - q*p  in [0, 2p)      for ANY x < 2^256
// so only the HIGH half of x*wq and the LOW halves of x*w and q*p are needed: 99 wide multiply-adds and
// 16 low-word ones instead of the 136 + 8 of a Montgomery multiplication, operands and result in the plain
// (non-Montgomery) domain.  The high half is computed from the partial products of columns >= 7 plus the
// high words of column 6; what is dropped is < 2^229, so the quotient is at most one short and
// r < 3p < 2^256; one conditional subtraction of 2p brings it back to [0,2p).
// Column-wise (Comba) accumulation: (c0,c1,c2) hold limbs k, k+1, k+2 while column k is summed; each
// product is one IMAD.WIDE with carry-out plus one IADD3.X (the ALU pipe idles in these kernels).
namespace detail {
#define LGR_COMBA_FULL(A, B) { c0 = mad_lo_cc((A), (B), c0); c1 = madc_hi_cc((A), (B), c1); c2 = addc(c2, 0); }
// approximate high half of x*y: q[0..7] = limbs 8..15 (at most one unit in q too small)
LGR_DEV void mul_hi_approx(uint32_t *q, const uint32_t *x, const uint32_t *y) {
    uint32_t c0 = 0, c1 = 0, c2 = 0;
#pragma unroll
    for (int i = 0; i < 7; i++) { c0 = mad_hi_cc(x[6 - i], y[i], c0); c1 = addc(c1, 0); }     // limb 7: high words of column 6
#pragma unroll
    for (int k = 7; k < 15; k++) {
#pragma unroll
        for (int i = k - 7; i < 8; i++) LGR_COMBA_FULL(x[k - i], y[i])
        if (k >= 8) q[k - 8] = c0;
        c0 = c1; c1 = c2; c2 = 0;
    }
    q[7] = c0;
}
// low half of x*y (mod 2^256); YCONST: y = p
template <bool YCONST>
LGR_DEV void mul_lo_256(uint32_t *r, const uint32_t *x, const uint32_t *y) {
    uint32_t c0 = 0, c1 = 0, c2 = 0;
#pragma unroll
    for (int k = 0; k < 7; k++) {
#pragma unroll
        for (int i = 0; i <= k; i++) LGR_COMBA_FULL(x[k - i], YCONST ? fr_p(i) : y[i])
        r[k] = c0;
        c0 = c1; c1 = c2; c2 = 0;
    }
#pragma unroll
    for (int i = 0; i < 8; i++) c0 += mul_lo(x[7 - i], YCONST ? fr_p(i) : y[i]);
    r[7] = c0;
}
#undef LGR_COMBA_FULL
}  // namespace detail

// x < 2^256 (any lazy range), w in [0,p), wq = floor(w*2^256/p)  ->  x*w mod p in [0,2p)
LGR_DEV fr_t fr_shoup_mul(const fr_t &x, const fr_t &w, const fr_t &wq) {
    uint32_t q[8], lo[8], qp[8];
    detail::mul_hi_approx(q, x.v, wq.v);
    detail::mul_lo_256<false>(lo, x.v, w.v);
    detail::mul_lo_256<true>(qp, q, nullptr);
    fr_t r;
    r.v[0] = sub_cc(lo[0], qp[0]);
#pragma unroll
    for (int i = 1; i < 7; i++) r.v[i] = subc_cc(lo[i], qp[i]);
    r.v[7] = subc(lo[7], qp[7]);
    return fr_reduce_2p(r);
}

// ---- wide (unreduced) accumulation of products ----------------------------------------------
// S += a*b as a plain 512-bit product added into a 576-bit accumulator: 64 wide multiply-adds and no
// reduction per product (a Montgomery multiplication costs 136).  The accumulator is kept as an
// even-aligned limb array E (limbs 0..15), an odd-aligned one O (O[x] = limb x+1, limbs 1..15) and
// per-limb carry counters C (limbs 8..16), so every partial-product row is one 8-limb carry chain
// plus one addc.  wide_reduce9 merges them and runs 9 Montgomery rounds:
//   result = S * 2^-288 mod p, in [0,2p), valid for S < 2^540 (e.g. 2^32 products of canonical values)
struct fr_wide { uint32_t e[16], o[15], c[17]; };

LGR_DEV void wide_zero(fr_wide &w) {
#pragma unroll
    for (int i = 0; i < 16; i++) w.e[i] = 0;
#pragma unroll
    for (int i = 0; i < 15; i++) w.o[i] = 0;
#pragma unroll
    for (int i = 0; i < 17; i++) w.c[i] = 0;
}
LGR_DEV void wide_mad(fr_wide &w, const fr_t &a, const fr_t &b) {
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
        detail::cmad_n(w.e + i, a.v, b.v[i]);          w.c[i + 8] = addc(w.c[i + 8], 0);    // even j: limbs i..i+7
        detail::cmad_n(w.o + i, a.v + 1, b.v[i]);      w.c[i + 9] = addc(w.c[i + 9], 0);    // odd j:  limbs i+1..i+8
        detail::cmad_n(w.o + i, a.v, b.v[i + 1]);      w.c[i + 9] = addc(w.c[i + 9], 0);    // even j: limbs i+1..i+8
        detail::cmad_n(w.e + i + 2, a.v + 1, b.v[i + 1]); w.c[i + 10] = addc(w.c[i + 10], 0); // odd j: limbs i+2..i+9
    }
}
// 9 Montgomery rounds on an 18-limb value: v[r] becomes 0, the value shifts down by 288 bits in total
LGR_DEV fr_t reduce9_rounds(uint32_t *v) {
#pragma unroll
    for (int r = 0; r < 9; r++) {
        const uint32_t m = mul_lo(v[r], LGR_M0);
        v[r] = mad_lo_cc(m, fr_p(0), v[r]);
#pragma unroll
        for (int j = 1; j < 8; j++) v[r + j] = madc_lo_cc(m, fr_p(j), v[r + j]);
#pragma unroll
        for (int j = r + 8; j < 17; j++) v[j] = addc_cc(v[j], 0);
        v[17] = addc(v[17], 0);
        v[r + 1] = mad_hi_cc(m, fr_p(0), v[r + 1]);
#pragma unroll
        for (int j = 1; j < 8; j++) v[r + 1 + j] = madc_hi_cc(m, fr_p(j), v[r + 1 + j]);
#pragma unroll
        for (int j = r + 9; j < 17; j++) v[j] = addc_cc(v[j], 0);
        v[17] = addc(v[17], 0);
    }
    fr_t res;
#pragma unroll
    for (int i = 0; i < 8; i++) res.v[i] = v[9 + i];
    return res;
}
LGR_DEV fr_t wide_reduce9(const fr_wide &w) {
    uint32_t v[18];
    // v = E + (O << 32) + (C << 32*limb)
    v[0] = w.e[0];
    v[1] = add_cc(w.e[1], w.o[0]);
#pragma unroll
    for (int i = 2; i < 16; i++) v[i] = addc_cc(w.e[i], w.o[i - 1]);
    v[16] = addc_cc(0, 0);
    v[17] = 0;
    v[8] = add_cc(v[8], w.c[8]);
#pragma unroll
    for (int i = 9; i < 17; i++) v[i] = addc_cc(v[i], w.c[i]);
    v[17] = addc(v[17], 0);
    return reduce9_rounds(v);
}

// ---- Karatsuba form of the wide accumulation (check_code over a resident tile) ---------------------
// Elements are < p < 2^254, so they split at bit 127 into two halves < 2^127 whose SUM still fits 128 bits:
//   x = x0 + x1 * 2^127,   r * x = r0 x0 + [(r0 + r1)(x0 + x1) - r0 x0 - r1 x1] * 2^127 + r1 x1 * 2^254
// Three 128 x 128 products (48 wide multiply-adds) instead of one 256 x 256 (64); the three partial sums are accumulated
// unreduced over the rows of a chunk and recombined once per chunk.  The multiplier is what bounds the sweep
// (DESIGN.md section 4), so the sweep gets 4/3 closer to the HBM roofline.
struct fr_half { uint32_t v[4]; };
struct fr_kara_wide { uint32_t e[8], o[7], c[9]; };        // 128x128 products: limbs 0..7, odd-aligned set, carries of limbs 4..8
LGR_DEV void kara_zero(fr_kara_wide &w) {
#pragma unroll
    for (int i = 0; i < 8; i++) w.e[i] = 0;
#pragma unroll
    for (int i = 0; i < 7; i++) w.o[i] = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) w.c[i] = 0;
}
namespace detail {
LGR_DEV void cmad_4(uint32_t *acc, const uint32_t *a, uint32_t bi) {      // acc[0..3] += {a[0], a[2]} * bi
    acc[0] = mad_lo_cc(a[0], bi, acc[0]);
    acc[1] = madc_hi_cc(a[0], bi, acc[1]);
    acc[2] = madc_lo_cc(a[2], bi, acc[2]);
    acc[3] = madc_hi_cc(a[2], bi, acc[3]);
}
}  // namespace detail
LGR_DEV void kara_mad(fr_kara_wide &w, const fr_half &a, const fr_half &b) {
    // a is read as a[0..4]: index 4 (beyond the half) is only touched by the odd rows through a+1 -> a[1], a[3]
#pragma unroll
    for (int i = 0; i < 4; i += 2) {
        detail::cmad_4(w.e + i, a.v, b.v[i]);             w.c[i + 4] = addc(w.c[i + 4], 0);
        detail::cmad_4(w.o + i, a.v + 1, b.v[i]);         w.c[i + 5] = addc(w.c[i + 5], 0);
        detail::cmad_4(w.o + i, a.v, b.v[i + 1]);         w.c[i + 5] = addc(w.c[i + 5], 0);
        detail::cmad_4(w.e + i + 2, a.v + 1, b.v[i + 1]); w.c[i + 6] = addc(w.c[i + 6], 0);
    }
}
// x -> (x0, x1, x0 + x1) with x0 = x mod 2^127, x1 = x >> 127  (x < 2^254)
LGR_DEV void kara_split(const fr_t &x, fr_half &x0, fr_half &x1, fr_half &xs) {
    x0.v[0] = x.v[0]; x0.v[1] = x.v[1]; x0.v[2] = x.v[2]; x0.v[3] = x.v[3] & 0x7FFFFFFFu;
#pragma unroll
    for (int i = 0; i < 3; i++) x1.v[i] = (x.v[3 + i] >> 31) | (x.v[4 + i] << 1);
    x1.v[3] = (x.v[6] >> 31) | (x.v[7] << 1);
    xs.v[0] = add_cc(x0.v[0], x1.v[0]);
    xs.v[1] = addc_cc(x0.v[1], x1.v[1]);
    xs.v[2] = addc_cc(x0.v[2], x1.v[2]);
    xs.v[3] = addc(x0.v[3], x1.v[3]);
}
LGR_DEV void kara_merge(uint32_t *v9, const fr_kara_wide &w) {             // 9 limbs: E + (O << 32) + carries
    v9[0] = w.e[0];
    v9[1] = add_cc(w.e[1], w.o[0]);
#pragma unroll
    for (int i = 2; i < 8; i++) v9[i] = addc_cc(w.e[i], w.o[i - 1]);
    v9[8] = addc(0, 0);
    v9[4] = add_cc(v9[4], w.c[4]);
#pragma unroll
    for (int i = 5; i < 8; i++) v9[i] = addc_cc(v9[i], w.c[i]);
    v9[8] = addc(v9[8], w.c[8]);
}
// (L + (M - L - H) * 2^127 + H * 2^254) * 2^-288 mod p, in [0,2p); L, H, M: sums of at most 2^6 products
LGR_DEV fr_t kara_reduce9(const fr_kara_wide &wl, const fr_kara_wide &wh, const fr_kara_wide &wm) {
    uint32_t L[9], H[9], M[9];
    kara_merge(L, wl); kara_merge(H, wh); kara_merge(M, wm);
    // M -= L + H  (non-negative)
    M[0] = sub_cc(M[0], L[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) M[i] = subc_cc(M[i], L[i]);
    M[8] = subc(M[8], L[8]);
    M[0] = sub_cc(M[0], H[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) M[i] = subc_cc(M[i], H[i]);
    M[8] = subc(M[8], H[8]);
    uint32_t v[18];
#pragma unroll
    for (int i = 0; i < 9; i++) v[i] = L[i];
#pragma unroll
    for (int i = 9; i < 18; i++) v[i] = 0;
    // + M << 127 = (M << 31) placed at limb 3: limbs 3..12
    uint32_t s[10];
    s[0] = M[0] << 31;
#pragma unroll
    for (int i = 1; i < 9; i++) s[i] = (M[i - 1] >> 1) | (M[i] << 31);
    s[9] = M[8] >> 1;
    v[3] = add_cc(v[3], s[0]);
#pragma unroll
    for (int i = 1; i < 10; i++) v[3 + i] = addc_cc(v[3 + i], s[i]);
#pragma unroll
    for (int i = 13; i < 17; i++) v[i] = addc_cc(v[i], 0);
    v[17] = addc(v[17], 0);
    // + H << 254 = (H << 30) placed at limb 7: limbs 7..16
    s[0] = H[0] << 30;
#pragma unroll
    for (int i = 1; i < 9; i++) s[i] = (H[i - 1] >> 2) | (H[i] << 30);
    s[9] = H[8] >> 2;
    v[7] = add_cc(v[7], s[0]);
#pragma unroll
    for (int i = 1; i < 10; i++) v[7 + i] = addc_cc(v[7 + i], s[i]);
    v[17] = addc(v[17], 0);
    return reduce9_rounds(v);
}

LGR_DEV bool fr_eq(const fr_t &a, const fr_t &b) {
    uint32_t d = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) d |= a.v[i] ^ b.v[i];
    return d == 0;
}
LGR_DEV bool fr_is_zero(const fr_t &a) {
    uint32_t d = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) d |= a.v[i];
    return d == 0;
}

}  // namespace lgr

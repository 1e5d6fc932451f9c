// Montgomery multiplication on the FP64 pipe ("double-precision floating point" formulation, after Emmart et al.):
// measured alternative to fr.cuh's IMAD.WIDE CIOS -- see DESIGN.md section 4 for the outcome.
//
// An element is 5 limbs of 52 bits held as doubles (exact integers < 2^52), R = 2^260.  For x, y < 2^52 the two halves of
// the 104-bit product come out of two fused multiply-adds with round-toward-zero:
//     ph = fma_rz(x, y, 2^104)                 = 2^104 + 2^52 * floor(xy / 2^52)          (the addend pins the exponent)
//     pl = fma_rz(x, y, (2^104 + 2^52) - ph)   = 2^52 + (xy mod 2^52)                     (exact: the high half cancels)
// so the IEEE bit patterns of ph and pl carry the two 52-bit halves in their mantissas and are accumulated as 64-bit
// integers; the exponent constants are subtracted once per column.  Per limb product: 2 DFMA + 1 DADD + two 64-bit adds.
#pragma once
#include <stdint.h>

namespace lgr {

struct dpf_t { double l[5]; };                       // limbs < 2^52, value < 2p

#define LGR_DPF_C1 0x1p104
#define LGR_DPF_C12 (0x1p104 + 0x1p52)          /* exact: one ulp of 2^104 */

__device__ __forceinline__ double dpf_modulus(int j) {
    // p = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001 in 52-bit limbs
    switch (j) {
        case 0: return (double)0x1f593f0000001ull;
        case 1: return (double)0x4879b9709143eull;
        case 2: return (double)0x181585d2833e8ull;
        case 3: return (double)0xa029b85045b68ull;
        default: return (double)0x030644e72e131ull;
    }
}
// -p^-1 mod 2^52
#define LGR_DPF_NP0 ((double)0x1f593efffffffull)

__device__ __forceinline__ unsigned long long dpf_bits(double x) { return (unsigned long long)__double_as_longlong(x); }
// integer v < 2^52 -> double, without the conversion unit: (2^52 + v) - 2^52
__device__ __forceinline__ double dpf_from_int(unsigned long long v) { return __longlong_as_double((long long)(v | 0x4330000000000000ull)) - 0x1p52; }

// acc[j] += lo(x*y), acc[j+1] += hi(x*y) as raw bit patterns (constants removed by the caller)
__device__ __forceinline__ void dpf_mac(unsigned long long &lo_acc, unsigned long long &hi_acc, double x, double y) {
    const double ph = __fma_rz(x, y, LGR_DPF_C1);
    const double c = (LGR_DPF_C12) - ph;
    const double pl = __fma_rz(x, y, c);
    lo_acc += dpf_bits(pl);
    hi_acc += dpf_bits(ph);
}

// a * b * 2^-260 mod p, result limbs normalised to < 2^52, value < 2p for a, b < 2p
__device__ __forceinline__ dpf_t dpf_mont_mul(const dpf_t &a, const dpf_t &b) {
    // 64-bit accumulators with wrap-around: the exponent constants are added with every half and taken out per column
    const unsigned long long KH = 0x4670000000000000ull, KL = 0x4330000000000000ull, M52 = 0x000FFFFFFFFFFFFFull;
    unsigned long long acc[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 5; i++) {
#pragma unroll
        for (int j = 0; j < 5; j++) dpf_mac(acc[j], acc[j + 1], a.l[j], b.l[i]);
        // column 0 has received one low half so far in this iteration
        const unsigned long long v = (acc[0] - KL) & M52;
        const double vd = dpf_from_int(v);
        const double qh = __fma_rz(vd, LGR_DPF_NP0, LGR_DPF_C1);
        const double qd = __fma_rz(vd, LGR_DPF_NP0, (LGR_DPF_C12) - qh) - 0x1p52;       // (v * np0) mod 2^52
#pragma unroll
        for (int j = 0; j < 5; j++) dpf_mac(acc[j], acc[j + 1], qd, dpf_modulus(j));
        // remove the exponent constants of this iteration: column j got 2 low halves (j < 5) and 2 high halves (j >= 1)
        acc[0] -= 2 * KL;
#pragma unroll
        for (int j = 1; j < 5; j++) acc[j] -= 2 * KL + 2 * KH;
        acc[5] -= 2 * KH;
        // column 0 is now a multiple of 2^52: shift the window down by one limb
        const unsigned long long carry = acc[0] >> 52;
        acc[0] = acc[1] + carry; acc[1] = acc[2]; acc[2] = acc[3]; acc[3] = acc[4]; acc[4] = acc[5]; acc[5] = 0;
    }
    dpf_t r;
    unsigned long long carry = 0;
#pragma unroll
    for (int j = 0; j < 5; j++) {
        const unsigned long long t = acc[j] + carry;
        r.l[j] = dpf_from_int(t & M52);
        carry = t >> 52;
    }
    return r;                                        // top limb absorbs nothing: value < 2p < 2^255 fits 5 limbs
}

// 8 x u32 little-endian limbs <-> 5 x 52-bit limbs
__device__ __forceinline__ dpf_t dpf_from_u32(const uint32_t *w) {
    unsigned long long q[4];
#pragma unroll
    for (int i = 0; i < 4; i++) q[i] = (unsigned long long)w[2 * i] | ((unsigned long long)w[2 * i + 1] << 32);
    const unsigned long long M = 0x000FFFFFFFFFFFFFull;
    dpf_t r;
    r.l[0] = dpf_from_int((q[0] & M));
    r.l[1] = dpf_from_int((((q[0] >> 52) | (q[1] << 12)) & M));
    r.l[2] = dpf_from_int((((q[1] >> 40) | (q[2] << 24)) & M));
    r.l[3] = dpf_from_int((((q[2] >> 28) | (q[3] << 36)) & M));
    r.l[4] = dpf_from_int((q[3] >> 16));
    return r;
}
__device__ __forceinline__ void dpf_to_u32(uint32_t *w, const dpf_t &a) {
    unsigned long long l[5];
#pragma unroll
    for (int j = 0; j < 5; j++) l[j] = (unsigned long long)__double2ll_rz(a.l[j]);
    unsigned long long q[4];
    q[0] = l[0] | (l[1] << 52);
    q[1] = (l[1] >> 12) | (l[2] << 40);
    q[2] = (l[2] >> 24) | (l[3] << 28);
    q[3] = (l[3] >> 36) | (l[4] << 16);
#pragma unroll
    for (int i = 0; i < 4; i++) { w[2 * i] = (uint32_t)q[i]; w[2 * i + 1] = (uint32_t)(q[i] >> 32); }
}

}  // namespace lgr

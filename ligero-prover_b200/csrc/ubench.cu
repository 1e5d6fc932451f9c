// Micro-benchmarks of the two issue-bound primitives ("first deliverable is an
// IMAD-throughput microbenchmark"): the path is integer-issue-bound, so the honest roofline
// next to HBM bandwidth is the measured IMAD.WIDE / Montgomery / SHA-256 rate of this part.
#include "kernels.h"
#include "ntt.cuh"

namespace lgr {

// 8 independent 32x32+64 multiply-add chains per thread; every chain multiplies its own running low
// word so that nothing can be hoisted or strength-reduced (WIDE = IMAD.WIDE.U32, otherwise IMAD lo)
template <bool WIDE>
__global__ void __launch_bounds__(256) ubench_imad_kernel(uint32_t *out, int iters) {
    unsigned long long a[8];
    uint32_t lo[8];
    const uint32_t y = blockIdx.x * 40503u + 977u + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = ((unsigned long long)(threadIdx.x + i) << 20) | 0x9E3779B1u; lo[i] = threadIdx.x * 2654435761u + i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (WIDE) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a[i]) : "r"((uint32_t)a[i]), "r"(y));
            else asm volatile("mad.lo.u32 %0, %0, %1, %0;" : "+r"(lo[i]) : "r"(y));
        }
    }
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= a[i] ^ lo[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)s ^ (uint32_t)(s >> 32);
}
// 8 independent double-precision FMA chains per thread (is the FP64 pipe a second multiplier?)
__global__ void __launch_bounds__(256) ubench_dfma_kernel(uint32_t *out, int iters) {
    double a[8];
    const double y = 1.0 + 1e-9 * threadIdx.x, z = 1e-3 * blockIdx.x;
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = 1.0 + i + threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = fma(a[i], y, z);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)__double2ll_rn(s);
}

// 4 independent Montgomery multiplications per iteration
__global__ void __launch_bounds__(256) ubench_mont_kernel(uint32_t *out, int iters) {
    fr_t x[4], w;
#pragma unroll
    for (int i = 0; i < 8; i++) w.v[i] = (threadIdx.x + 1) * 0x9E3779B1u + i * 0x85EBCA77u;
    w.v[7] &= 0x0FFFFFFFu;
#pragma unroll
    for (int q = 0; q < 4; q++) { x[q] = w; x[q].v[0] += q + blockIdx.x; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int q = 0; q < 4; q++) x[q] = fr_mont_mul(x[q], w);
    }
    uint32_t s = 0;
#pragma unroll
    for (int q = 0; q < 4; q++)
#pragma unroll
        for (int i = 0; i < 8; i++) s ^= x[q].v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 4 independent Shoup multiplications per iteration (fr_shoup_mul: the twiddle multiplication of the butterflies)
__global__ void __launch_bounds__(256) ubench_shoup_kernel(uint32_t *out, int iters) {
    fr_t x[4], w, wq;
#pragma unroll
    for (int i = 0; i < 8; i++) { w.v[i] = (threadIdx.x + 1) * 0x9E3779B1u + i * 0x85EBCA77u; wq.v[i] = w.v[i] * 0x2545F491u + 77u; }
    w.v[7] &= 0x0FFFFFFFu;
#pragma unroll
    for (int q = 0; q < 4; q++) { x[q] = w; x[q].v[0] += q + blockIdx.x; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int q = 0; q < 4; q++) x[q] = fr_shoup_mul(x[q], w, wq);
    }
    uint32_t s = 0;
#pragma unroll
    for (int q = 0; q < 4; q++)
#pragma unroll
        for (int i = 0; i < 8; i++) s ^= x[q].v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ uint32_t ub_rotr(uint32_t x, int n) { return __funnelshift_r(x, x, n); }
// one dependent SHA-256 compression per iteration (same round structure as sha_kernels.cu)
__global__ void __launch_bounds__(256) ubench_sha_kernel(uint32_t *out, int iters) {
    uint32_t st[8], w[16];
#pragma unroll
    for (int i = 0; i < 8; i++) st[i] = threadIdx.x * 0x01000193u + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) w[i] = st[i & 7] + i + it;
        uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll
        for (int i = 0; i < 64; i++) {
            uint32_t wi;
            if (i < 16) wi = w[i];
            else {
                const uint32_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
                wi = w[i & 15] + (ub_rotr(w15, 7) ^ ub_rotr(w15, 18) ^ (w15 >> 3)) + w[(i + 9) & 15] + (ub_rotr(w2, 17) ^ ub_rotr(w2, 19) ^ (w2 >> 10));
                w[i & 15] = wi;
            }
            const uint32_t t1 = h + (ub_rotr(e, 6) ^ ub_rotr(e, 11) ^ ub_rotr(e, 25)) + ((e & f) ^ (~e & g)) + 0x428a2f98u * (i + 1) + wi;
            const uint32_t t2 = (ub_rotr(a, 2) ^ ub_rotr(a, 13) ^ ub_rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
            h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= st[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- single-warp SHA-256 chain experiments (one warp per SM sub-partition) ---------------------
// The column hash of a narrow matrix (n = 1024 -> 32 warps) is a dependent chain: what matters is
// the cycles one warp needs per compression when it owns a scheduler.  Variants:
//   3: plain rounds + in-line message schedule (what sha_update_kernel does)
//   4: rounds only, K+W pre-computed in shared memory (message schedule done by helper warps)
//   5: as 4, with part of the rotations / additions moved to the FMA pipe through opaque
//      multipliers (kernel parameters), so ALU and FMA pipes share the 2-cycle/instruction load
struct ChainConsts { uint32_t one, m6, m11, m25, m2, m13, m22, mone; };

__device__ __forceinline__ uint32_t rot_fma(uint32_t x, uint32_t mult) {      // rotr(x, s), mult = 2^(32-s): 2 FMA-pipe ops
    uint32_t hi = __umulhi(x, mult);
    return x * mult + hi;
}
__device__ __forceinline__ uint32_t add_fma(uint32_t a, uint32_t b, uint32_t one) { return a * one + b; }

template <int VARIANT>
__device__ __forceinline__ void chain_rounds(uint32_t st[8], const uint32_t *kw /* [64][32] in smem */, int lane, const ChainConsts cc) {
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll
    for (int i = 0; i < 64; i++) {
        const uint32_t kwi = kw[i * 32 + lane];
        if (VARIANT == 4) {
            const uint32_t t1 = h + (ub_rotr(e, 6) ^ ub_rotr(e, 11) ^ ub_rotr(e, 25)) + ((e & f) ^ (~e & g)) + kwi;
            const uint32_t t2 = (ub_rotr(a, 2) ^ ub_rotr(a, 13) ^ ub_rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
            h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        } else if (VARIANT == 17) {
            // e' through two FMA-pipe adds (h+KW+d+Ch early, + Sigma1 late), a' as IADD3: 11 ALU ops per round
            const uint32_t s1 = ub_rotr(e, 6) ^ ub_rotr(e, 11) ^ ub_rotr(e, 25);
            const uint32_t ch = (e & f) ^ (~e & g);
            const uint32_t s0 = ub_rotr(a, 2) ^ ub_rotr(a, 13) ^ ub_rotr(a, 22);
            const uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
            const uint32_t hkd = add_fma(add_fma(h, kwi, cc.one), d, cc.one);
            const uint32_t pd = add_fma(d, mj, cc.mone);
            const uint32_t ne = add_fma(add_fma(ch, hkd, cc.one), s1, cc.one);
            const uint32_t na = ne + s0 + pd;
            h = g; g = f; f = e; e = ne; d = c; c = b; b = a; a = na;
        } else if (VARIANT >= 13) {
            // 13..16: variant 11's association with rotations moved to the FMA pipe where the chain has slack
            //  13: rotr(a,22) as IMAD.HI + IMAD       14: rotr(a,22) and rotr(e,25)
            //  15: rotr(a,22) and rotr(a,13)          16: rotr(a,22) as IMAD.WIDE halves xor-ed by a wider LOP3 tree
            const uint32_t r25 = (VARIANT == 14) ? rot_fma(e, cc.m25) : ub_rotr(e, 25);
            const uint32_t s1 = ub_rotr(e, 6) ^ ub_rotr(e, 11) ^ r25;
            const uint32_t ch = (e & f) ^ (~e & g);
            uint32_t s0;
            if (VARIANT == 16) {
                const unsigned long long w = (unsigned long long)a * cc.m22;
                s0 = (ub_rotr(a, 2) ^ ub_rotr(a, 13) ^ (uint32_t)w) ^ (uint32_t)(w >> 32);
            } else {
                const uint32_t r13 = (VARIANT == 15) ? rot_fma(a, cc.m13) : ub_rotr(a, 13);
                s0 = ub_rotr(a, 2) ^ r13 ^ rot_fma(a, cc.m22);
            }
            const uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
            const uint32_t hk = add_fma(h, kwi, cc.one);
            const uint32_t hkd = add_fma(hk, d, cc.one);
            const uint32_t pd = add_fma(d, mj, cc.mone);
            const uint32_t ne = hkd + s1 + ch;
            const uint32_t na = ne + s0 + pd;
            h = g; g = f; f = e; e = ne; d = c; c = b; b = a; a = na;
        } else if (VARIANT >= 9) {
            // 9..12: all rotations on the ALU pipe; the additions are re-associated so that only
            // Sigma -> sum sits on the dependent chain, and the early sums run on the FMA pipe
            //   9: ne = IADD3 (ALU), na through FMA-pipe adds        (11 ALU ops / round)
            //  10: every addition on the FMA pipe                    (10 ALU ops / round)
            //  11: ne and na both IADD3, early sums on the FMA pipe  (12 ALU ops / round, shortest chain)
            //  12: as 11 but plain C (ptxas picks the pipes)
            const uint32_t s1 = ub_rotr(e, 6) ^ ub_rotr(e, 11) ^ ub_rotr(e, 25);
            const uint32_t ch = (e & f) ^ (~e & g);
            const uint32_t s0 = ub_rotr(a, 2) ^ ub_rotr(a, 13) ^ ub_rotr(a, 22);
            const uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
            uint32_t ne, na;
            if (VARIANT == 12) {
                const uint32_t hkd = (h + kwi) + d;
                const uint32_t pd = mj - d;
                ne = hkd + s1 + ch;
                na = ne + s0 + pd;
            } else {
                const uint32_t hk = add_fma(h, kwi, cc.one);
                const uint32_t hkd = add_fma(hk, d, cc.one);
                const uint32_t pd = add_fma(d, mj, cc.mone);              // Maj - d
                if (VARIANT == 9) {
                    ne = hkd + s1 + ch;
                    const uint32_t q = add_fma(s0, pd, cc.one);
                    na = add_fma(ne, q, cc.one);
                } else if (VARIANT == 10) {
                    ne = add_fma(add_fma(ch, hkd, cc.one), s1, cc.one);
                    const uint32_t q = add_fma(s0, pd, cc.one);
                    na = add_fma(ne, q, cc.one);
                } else {
                    ne = hkd + s1 + ch;
                    na = ne + s0 + pd;
                }
            }
            h = g; g = f; f = e; e = ne; d = c; c = b; b = a; a = na;
        } else if (VARIANT >= 6) {
            // 6: one FMA-pipe rotation in each Sigma; 7: Sigma1 one, Sigma0 two; 8: Sigma1 two, Sigma0 one
            const uint32_t r25 = rot_fma(e, cc.m25);
            const uint32_t r11 = (VARIANT == 8) ? rot_fma(e, cc.m11) : ub_rotr(e, 11);
            const uint32_t s1 = ub_rotr(e, 6) ^ r11 ^ r25;
            const uint32_t r22 = rot_fma(a, cc.m22);
            const uint32_t r13 = (VARIANT == 7) ? rot_fma(a, cc.m13) : ub_rotr(a, 13);
            const uint32_t s0 = ub_rotr(a, 2) ^ r13 ^ r22;
            const uint32_t t1 = h + s1 + ((e & f) ^ (~e & g)) + kwi;
            const uint32_t t2 = s0 + ((a & b) ^ (a & c) ^ (b & c));
            h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        } else {
            // early terms on the FMA pipe
            const uint32_t x = add_fma(h, kwi, cc.one);                  // h + KW
            const uint32_t ch = (e & f) ^ (~e & g);
            const uint32_t xch = add_fma(x, ch, cc.one);                 // h + KW + Ch
            const uint32_t ych = add_fma(xch, d, cc.one);                // d + h + KW + Ch
            const uint32_t s1 = ub_rotr(e, 6) ^ rot_fma(e, cc.m11) ^ ub_rotr(e, 25);
            const uint32_t s0 = ub_rotr(a, 2) ^ rot_fma(a, cc.m13) ^ rot_fma(a, cc.m22);
            const uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
            const uint32_t ne = ych + s1;
            const uint32_t na = xch + s1 + (s0 + mj);
            h = g; g = f; f = e; e = ne; d = c; c = b; b = a; a = na;
        }
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}

template <int LUT> __device__ __forceinline__ uint32_t ub_lop3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d; asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(LUT)); return d;
}
__device__ __forceinline__ uint32_t ub_mad(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d; asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d;
}

template <int VARIANT>
__global__ void __launch_bounds__(128) ubench_chain_kernel(uint32_t *out, int iters, const ChainConsts cc, int active_lanes) {
    __shared__ uint32_t kw[4][64 * 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = lane; i < 64 * 32; i += 32) kw[warp][i] = i * 2654435761u + warp;
    __syncwarp();
    if (lane >= active_lanes) return;
    uint32_t st[8];
#pragma unroll
    for (int i = 0; i < 8; i++) st[i] = threadIdx.x * 0x01000193u + i;
    long long t0 = clock64();
    if (VARIANT == 3) {
        uint32_t w[16];
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 16; i++) w[i] = kw[warp][i * 32 + lane] + it;
            uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll
            for (int i = 0; i < 64; i++) {
                uint32_t wi;
                if (i < 16) wi = w[i];
                else {
                    const uint32_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
                    wi = w[i & 15] + (ub_rotr(w15, 7) ^ ub_rotr(w15, 18) ^ (w15 >> 3)) + w[(i + 9) & 15] + (ub_rotr(w2, 17) ^ ub_rotr(w2, 19) ^ (w2 >> 10));
                    w[i & 15] = wi;
                }
                const uint32_t t1 = h + (ub_rotr(e, 6) ^ ub_rotr(e, 11) ^ ub_rotr(e, 25)) + ((e & f) ^ (~e & g)) + 0x428a2f98u * (i + 1) + wi;
                const uint32_t t2 = (ub_rotr(a, 2) ^ ub_rotr(a, 13) ^ ub_rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
                h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
            }
            st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
        }
    } else if (VARIANT == 18) {
        // latency of one SHFL.BFLY hop: 64 dependent shuffles (+ one dependent add each) per iteration
        uint32_t x = st[0];
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 64; i++) x = __shfl_xor_sync(0xffffffffu, x, 16) + cc.one;
        }
        st[0] = x;
    } else if (VARIANT == 19 || VARIANT == 20) {
        // lane-split rounds (sha_kernels.cu: sha_chain16_kernel): E half on lanes 0-15, A half on lanes 16-31, one
        // shuffle per round; 19: partner word used two iterations later (as in the kernel), 20: three iterations later
        // (not a valid SHA schedule -- it only shows what the exchange latency costs)
        const bool isE = lane < 16;
        const uint32_t s1 = isE ? 6 : 2, s2 = isE ? 11 : 13, s3 = isE ? 25 : 22;
        const uint32_t sgn = isE ? cc.one : 0u - cc.one, M = isE ? 0u : 0u - cc.one;
        uint32_t p = st[0], q = st[1], r = st[2], s = st[3], rm3 = st[4], rm2 = st[5], rm1 = st[6];
        uint32_t u = ub_lop3<0xF8>(q, r, M), v = ub_lop3<0xC4>(q, r, M);
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 64; i++) {
                const uint32_t d = (VARIANT == 20) ? rm3 : rm2;
                const uint32_t wr = ub_mad(ub_mad(s, sgn, kw[warp][i * 32 + lane]), cc.one, d);
                const uint32_t sg = ub_lop3<0x96>(__funnelshift_r(p, p, s1), __funnelshift_r(p, p, s2), __funnelshift_r(p, p, s3));
                const uint32_t np = sg + ub_lop3<0xCA>(p, u, v) + wr;
                const uint32_t r0 = __shfl_xor_sync(0xffffffffu, np, 16);
                s = r; r = q; q = p; p = np; u = ub_lop3<0xF8>(q, r, M); v = ub_lop3<0xC4>(q, r, M);
                rm3 = rm2; rm2 = rm1; rm1 = r0;
            }
        }
        st[0] = p; st[1] = q; st[2] = r; st[3] = s; st[4] = rm2 ^ rm3;
    } else if (VARIANT >= 21 && VARIANT <= 25) {
        // the block loop of sha_chain16_kernel verbatim (boundary steps, K+W ring of eight registers, A lanes reading
        // zeros); 22: without the boundary steps and the first-four-rounds special case; 23: as 22 with 32-word rows (the A
        // lanes read the upper half of the row the E lanes read: one 128-byte row per LDS)
        const bool isE = lane < 16;
        const uint32_t s1 = isE ? 6 : 2, s2 = isE ? 11 : 13, s3 = isE ? 25 : 22;
        const uint32_t one = cc.one, sgn = isE ? one : 0u - one, mE = isE ? one : 0u, mA = isE ? 0u : one, M = 0u - mA;
        uint32_t p = st[0], q = st[1], r = st[2], s = st[3], rm2 = st[5], rm1 = st[6];
        uint32_t cv0 = 0, cv1 = 0, cv2 = 0, cv3 = 0, ca[4] = {0, 0, 0, 0};
        uint32_t u = ub_lop3<0xF8>(q, r, M), v = ub_lop3<0xC4>(q, r, M);
        uint32_t kq[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        constexpr int RS = (VARIANT == 23) ? 32 : 16;            // row stride in words
        constexpr int D = (VARIANT == 24) ? 4 : ((VARIANT == 25) ? 2 : 8);   // 24 / 25: K+W fetched 4 / 2 rounds ahead instead of 8
        const uint32_t *kwp = (VARIANT == 23) ? &kw[warp][lane] : (isE ? &kw[warp][lane] : &kw[warp][16 + (lane & 15)]);
        const uint32_t *kwn = kwp;
#define UB_BOUNDARY(MX) { uint32_t t_; t_ = p; p = cv0 * (MX) + p; cv0 = t_ * (MX) + cv0; t_ = q; q = cv1 * (MX) + q; cv1 = t_ * (MX) + cv1; \
      t_ = r; r = cv2 * (MX) + r; cv2 = t_ * (MX) + cv2; t_ = s; s = cv3 * (MX) + s; cv3 = t_ * (MX) + cv3; u = ub_lop3<0xF8>(q, r, M); v = ub_lop3<0xC4>(q, r, M); }
#define UB_ROUND(J, KWX) { uint32_t d_ = rm2; \
      if (VARIANT == 21 && (J) < 4) { d_ = ub_mad(ca[3 - (J)], one, rm2); ca[3 - (J)] = ub_mad(d_, mE, 0u); } \
      const uint32_t wr_ = ub_mad(ub_mad(s, sgn, (KWX)), one, d_); \
      const uint32_t sg_ = ub_lop3<0x96>(__funnelshift_r(p, p, s1), __funnelshift_r(p, p, s2), __funnelshift_r(p, p, s3)); \
      const uint32_t np_ = sg_ + ub_lop3<0xCA>(p, u, v) + wr_; \
      const uint32_t r0_ = __shfl_xor_sync(0xffffffffu, np_, 16); \
      s = r; r = q; q = p; p = np_; u = ub_lop3<0xF8>(q, r, M); v = ub_lop3<0xC4>(q, r, M); rm2 = rm1; rm1 = r0_; }
#pragma unroll
        for (int j = 0; j < D; j++) kq[j] = kwp[j * RS];
#pragma unroll 1
        for (int it = 0; it < iters; it++) {
            if (VARIANT == 21) UB_BOUNDARY(mE)
            UB_ROUND(0, kq[0]) kq[0] = (D > 0) ? kwp[D * RS] : 0u;
            UB_ROUND(1, kq[1 % D]) kq[1 % D] = kwp[(1 + D) * RS];
            if (VARIANT == 21) UB_BOUNDARY(mA)
#pragma unroll
            for (int j = 2; j < 64; j++) {
                const uint32_t kw_ = kq[j % D];
                kq[j % D] = (j + D < 64) ? kwp[(j + D) * RS] : kwn[(j + D - 64) * RS];
                UB_ROUND(j, kw_)
            }
        }
#undef UB_ROUND
#undef UB_BOUNDARY
        st[0] = p + cv0; st[1] = q + cv1; st[2] = r + cv2; st[3] = s + cv3; st[4] = rm2 ^ ca[0] ^ ca[1] ^ ca[2] ^ ca[3];
    } else {
        for (int it = 0; it < iters; it++) chain_rounds<VARIANT>(st, kw[warp], lane, cc);
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= st[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[148 * 8 * 256 - 1] = (uint32_t)((t1 - t0) / iters); }
}

cudaError_t launch_ubench_chain(int variant, uint32_t *out, int iters, int warps_per_cta, int active_lanes, cudaStream_t st) {
    ChainConsts cc{1u, 1u << 26, 1u << 21, 1u << 7, 1u << 30, 1u << 19, 1u << 10, 0xFFFFFFFFu};
    const int threads = warps_per_cta * 32;
    if (variant == 3) ubench_chain_kernel<3><<<148, threads, 0, st>>>(out, iters, cc, active_lanes);
    else if (variant == 4) ubench_chain_kernel<4><<<148, threads, 0, st>>>(out, iters, cc, active_lanes);
    else if (variant == 5) ubench_chain_kernel<5><<<148, threads, 0, st>>>(out, iters, cc, active_lanes);
    else if (variant == 6) ubench_chain_kernel<6><<<148, threads, 0, st>>>(out, iters, cc, active_lanes);
    else if (variant == 7) ubench_chain_kernel<7><<<148, threads, 0, st>>>(out, iters, cc, active_lanes);
    else if (variant == 8) ubench_chain_kernel<8><<<148, threads, 0, st>>>(out, iters, cc, active_lanes);
    else if (variant == 9) ubench_chain_kernel<9><<<148, threads, 0, st>>>(out, iters, cc, active_lanes);
    else if (variant == 10) ubench_chain_kernel<10><<<148, threads, 0, st>>>(out, iters, cc, active_lanes);
    else if (variant == 11) ubench_chain_kernel<11><<<148, threads, 0, st>>>(out, iters, cc, active_lanes);
    else if (variant == 12) ubench_chain_kernel<12><<<148, threads, 0, st>>>(out, iters, cc, active_lanes);
    else if (variant == 13) ubench_chain_kernel<13><<<148, threads, 0, st>>>(out, iters, cc, active_lanes);
    else if (variant == 14) ubench_chain_kernel<14><<<148, threads, 0, st>>>(out, iters, cc, active_lanes);
    else if (variant == 15) ubench_chain_kernel<15><<<148, threads, 0, st>>>(out, iters, cc, active_lanes);
    else if (variant == 16) ubench_chain_kernel<16><<<148, threads, 0, st>>>(out, iters, cc, active_lanes);
    else if (variant == 17) ubench_chain_kernel<17><<<148, threads, 0, st>>>(out, iters, cc, active_lanes);
    else if (variant == 18) ubench_chain_kernel<18><<<148, threads, 0, st>>>(out, iters, cc, active_lanes);
    else if (variant == 19) ubench_chain_kernel<19><<<148, threads, 0, st>>>(out, iters, cc, active_lanes);
    else if (variant == 20) ubench_chain_kernel<20><<<148, threads, 0, st>>>(out, iters, cc, active_lanes);
    else if (variant == 21) ubench_chain_kernel<21><<<148, threads, 0, st>>>(out, iters, cc, active_lanes);
    else if (variant == 22) ubench_chain_kernel<22><<<148, threads, 0, st>>>(out, iters, cc, active_lanes);
    else if (variant == 23) ubench_chain_kernel<23><<<148, threads, 0, st>>>(out, iters, cc, active_lanes);
    else if (variant == 24) ubench_chain_kernel<24><<<148, threads, 0, st>>>(out, iters, cc, active_lanes);
    else ubench_chain_kernel<25><<<148, threads, 0, st>>>(out, iters, cc, active_lanes);
    return cudaGetLastError();
}

// Montgomery-multiply throughput as a function of resident warps per SM and independent chains per
// thread: tells how much TLP / ILP the butterfly kernels need to keep the FMA pipe busy.
template <int NCHAIN>
__global__ void ubench_mont_occ_kernel(uint32_t *out, int iters) {
    extern __shared__ unsigned char occ_pad[];
    fr_t x[NCHAIN], w;
#pragma unroll
    for (int i = 0; i < 8; i++) w.v[i] = (threadIdx.x + 1) * 0x9E3779B1u + i * 0x85EBCA77u;
    w.v[7] &= 0x0FFFFFFFu;
#pragma unroll
    for (int q = 0; q < NCHAIN; q++) { x[q] = w; x[q].v[0] += q + blockIdx.x; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int q = 0; q < NCHAIN; q++) x[q] = fr_mont_mul(x[q], w);
    }
    uint32_t s = 0;
#pragma unroll
    for (int q = 0; q < NCHAIN; q++)
#pragma unroll
        for (int i = 0; i < 8; i++) s ^= x[q].v[i];
    if (s == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = s + occ_pad[0];
}
cudaError_t launch_ubench_mont_occ(int nchain, int warps_per_sm, uint32_t *out, int iters, cudaStream_t st) {
    // one CTA per SM with warps_per_sm warps; 120 KiB of dynamic shared memory keeps it alone on the SM
    const size_t smem = 120 * 1024;
    const int threads = warps_per_sm * 32;
    cudaError_t e;
#define LGR_OCC(N)                                                                                                        \
    e = cudaFuncSetAttribute(ubench_mont_occ_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);          \
    if (e != cudaSuccess) return e;                                                                                       \
    ubench_mont_occ_kernel<N><<<148, threads, smem, st>>>(out, iters);
    if (nchain == 1) { LGR_OCC(1) } else if (nchain == 2) { LGR_OCC(2) } else { LGR_OCC(4) }
    return cudaGetLastError();
}

cudaError_t launch_ubench(int which, uint32_t *out, int iters, int blocks, int threads, cudaStream_t st) {
    if (which == 0) ubench_imad_kernel<true><<<blocks, threads, 0, st>>>(out, iters);
    else if (which == 3) ubench_imad_kernel<false><<<blocks, threads, 0, st>>>(out, iters);
    else if (which == 4) ubench_dfma_kernel<<<blocks, threads, 0, st>>>(out, iters);
    else if (which == 1) ubench_mont_kernel<<<blocks, threads, 0, st>>>(out, iters);
    else if (which == 5) ubench_shoup_kernel<<<blocks, threads, 0, st>>>(out, iters);
    else ubench_sha_kernel<<<blocks, threads, 0, st>>>(out, iters);
    return cudaGetLastError();
}

}  // namespace lgr

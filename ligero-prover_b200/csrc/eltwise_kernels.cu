// Element-wise field kernels and the stage-2 row combiners.
//
// Replaces the Eltwise* entry points of shader/kernels.wgsl.in:325-549 (one dispatch per
// operation, Barrett reduction = 3 full + 1 low-half 256-bit products per element)
// and the check_code / check_linear / check_quadratic schedules of
// include/zkp/nonbatch_context.hpp:756-780.  Canonical in, canonical out, bit-exact with the WGSL
// for canonical inputs.  A canonical product x*y is two Montgomery multiplications
// (x*y/R, then *R^2/R); a product with a host constant c is one (c is pre-multiplied by R).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include "kernels.h"
#include "ntt.cuh"
#include "dpf_mont.cuh"

namespace lgr {

// 2^512 mod p (limbs, little-endian): converts x*y/R back to x*y
__device__ __forceinline__ fr_t fr_R2() {
    fr_t r;
    r.v[0] = 0xae216da7u; r.v[1] = 0x1bb8e645u; r.v[2] = 0xe35c59e3u; r.v[3] = 0x53fe3ab1u;
    r.v[4] = 0x53bb8085u; r.v[5] = 0x8c49833du; r.v[6] = 0x7f4e44a5u; r.v[7] = 0x0216d0b1u;
    return r;
}
// 2^256 mod p (shader/bn254fr.wgsl.in:37-39)
__device__ __forceinline__ fr_t fr_R1() {
    fr_t r;
    r.v[0] = 0x4FFFFFFBu; r.v[1] = 0xAC96341Cu; r.v[2] = 0x9F60CD29u; r.v[3] = 0x36FC7695u;
    r.v[4] = 0x7879462Eu; r.v[5] = 0x666EA36Fu; r.v[6] = 0x9A07DF2Fu; r.v[7] = 0x0E0A77C1u;
    return r;
}
__device__ __forceinline__ fr_t fr_from_words(const uint32_t *w) { fr_t r; for (int i = 0; i < 8; i++) r.v[i] = w[i]; return r; }

// canonical x*y
__device__ __forceinline__ fr_t fr_mul_canon(const fr_t &x, const fr_t &y) {
    fr_t t = fr_reduce_p(fr_mont_mul(x, y));           // x*y/R, canonical (second operand must be < p)
    return fr_reduce_p(fr_mont_mul(fr_R2(), t));       // * R^2 / R
}
// Montgomery-domain helpers for inversion / powers
__device__ __forceinline__ fr_t fr_msqr(const fr_t &a) { return fr_reduce_p(fr_mont_mul(a, a)); }
__device__ __forceinline__ fr_t fr_mmul(const fr_t &a, const fr_t &b) { return fr_reduce_p(fr_mont_mul(a, b)); }

// y^-1 (0 -> 0, like the WGSL extended Euclid, bn254fr.wgsl.in:128-151) via y^(p-2)
__device__ fr_t fr_inv_canon(const fr_t &y) {
    fr_t base = fr_mmul(y, fr_R2());                    // yR
    fr_t acc = fr_R1();                                 // 1R
    // exponent p-2, little-endian limbs
    const uint32_t e[8] = {LGR_P0 - 2u, LGR_P1, LGR_P2, LGR_P3, LGR_P4, LGR_P5, LGR_P6, LGR_P7};
    for (int i = 0; i < 254; i++) {
        if ((e[i >> 5] >> (i & 31)) & 1u) acc = fr_mmul(acc, base);
        base = fr_msqr(base);
    }
    fr_t one = fr_zero(); one.v[0] = 1;
    return fr_mmul(acc, one);                           // leave Montgomery form
}

template <int OP>
__global__ void __launch_bounds__(256) eltwise_kernel(const EltParams p) {
    const fr_t c = fr_from_words(p.scalar);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < p.n; i += (size_t)gridDim.x * blockDim.x) {
        fr_t r;
        if constexpr (OP == ELT_ADD) r = fr_add(fr_ldg(p.x + i), fr_ldg(p.y + i));
        else if constexpr (OP == ELT_SUB) r = fr_sub(fr_ldg(p.x + i), fr_ldg(p.y + i));
        else if constexpr (OP == ELT_MUL) r = fr_mul_canon(fr_ldg(p.x + i), fr_ldg(p.y + i));
        else if constexpr (OP == ELT_DIV) r = fr_mul_canon(fr_ldg(p.x + i), fr_inv_canon(fr_ldg(p.y + i)));
        else if constexpr (OP == ELT_FMA) r = fr_add(fr_ldg(p.out + i), fr_mul_canon(fr_ldg(p.x + i), fr_ldg(p.y + i)));
        else if constexpr (OP == ELT_FMA_CONST) r = fr_add(fr_ldg(p.out + i), fr_reduce_p(fr_mont_mul(fr_ldg(p.x + i), c)));   // c = const*R
        else if constexpr (OP == ELT_ADD_ASSIGN) r = fr_add(fr_ldg(p.out + i), fr_ldg(p.x + i));
        else if constexpr (OP == ELT_ADD_CONST) r = fr_add(fr_ldg(p.x + i), c);
        else if constexpr (OP == ELT_SUB_CONST) r = fr_sub(fr_ldg(p.x + i), c);
        else if constexpr (OP == ELT_CONST_SUB) r = fr_sub(c, fr_ldg(p.x + i));
        else if constexpr (OP == ELT_MUL_CONST) r = fr_reduce_p(fr_mont_mul(fr_ldg(p.x + i), c));                                 // c = const*R
        else if constexpr (OP == ELT_MONTMUL_CONST) r = fr_reduce_p(fr_mont_mul(fr_ldg(p.x + i), c));                             // c = const
        else if constexpr (OP == ELT_BIT) {
            fr_t x = fr_ldg(p.x + i);
            r = fr_zero();
            uint32_t limb = 0;
#pragma unroll
            for (int q = 0; q < 8; q++) if (q == (int)((p.bit >> 5) & 7u)) limb = x.v[q];
            r.v[0] = (limb >> (p.bit & 31u)) & 1u;
        } else if constexpr (OP == ELT_POWMOD || OP == ELT_POWADD) {
            // out = coeff * base^exp (+ out); c = base*R; table-free square-and-multiply over 32 bits
            // (kernels.wgsl.in:512-537, bn254fr.wgsl.in:157-168)
            const uint32_t ex = p.idx[i];
            fr_t acc = fr_R1(), b = c;
            for (int q = 0; q < 32; q++) {
                if ((ex >> q) & 1u) acc = fr_mmul(acc, b);
                if ((ex >> q) >> 1) b = fr_msqr(b); else break;
            }
            r = fr_mmul(fr_ldg(p.x + i), acc);          // coeff * (base^exp R) / R
            if constexpr (OP == ELT_POWADD) r = fr_add(fr_ldg(p.out + i), r);
        } else if constexpr (OP == ELT_GATHER) {
            r = fr_ldg(p.x + p.idx[i]);
        } else if constexpr (OP == ELT_QUAD_FUSED) {
            // out += r*(x*y - z): check_quadratic (nonbatch_context.hpp:771-780) in one sweep; c = r*R
            fr_t t = fr_sub(fr_mul_canon(fr_ldg(p.x + i), fr_ldg(p.y + i)), fr_ldg(p.z + i));
            r = fr_add(fr_ldg(p.out + i), fr_reduce_p(fr_mont_mul(t, c)));
        }
        fr_stg(p.out + i, r);
    }
}

template <int OP>
static cudaError_t launch_elt(const EltParams &p, cudaStream_t st) {
    if (p.n == 0) return cudaSuccess;
    size_t blocks = (p.n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    eltwise_kernel<OP><<<(int)blocks, 256, 0, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_eltwise(EltOp op, const EltParams &p, cudaStream_t st) {
    switch (op) {
        case ELT_ADD: return launch_elt<ELT_ADD>(p, st);
        case ELT_SUB: return launch_elt<ELT_SUB>(p, st);
        case ELT_MUL: return launch_elt<ELT_MUL>(p, st);
        case ELT_DIV: return launch_elt<ELT_DIV>(p, st);
        case ELT_FMA: return launch_elt<ELT_FMA>(p, st);
        case ELT_FMA_CONST: return launch_elt<ELT_FMA_CONST>(p, st);
        case ELT_ADD_ASSIGN: return launch_elt<ELT_ADD_ASSIGN>(p, st);
        case ELT_ADD_CONST: return launch_elt<ELT_ADD_CONST>(p, st);
        case ELT_SUB_CONST: return launch_elt<ELT_SUB_CONST>(p, st);
        case ELT_CONST_SUB: return launch_elt<ELT_CONST_SUB>(p, st);
        case ELT_MUL_CONST: return launch_elt<ELT_MUL_CONST>(p, st);
        case ELT_MONTMUL_CONST: return launch_elt<ELT_MONTMUL_CONST>(p, st);
        case ELT_BIT: return launch_elt<ELT_BIT>(p, st);
        case ELT_POWMOD: return launch_elt<ELT_POWMOD>(p, st);
        case ELT_POWADD: return launch_elt<ELT_POWADD>(p, st);
        case ELT_GATHER: return launch_elt<ELT_GATHER>(p, st);
        case ELT_QUAD_FUSED: return launch_elt<ELT_QUAD_FUSED>(p, st);
    }
    return cudaErrorInvalidValue;
}

// ---- tile combiners ----------------------------------------------------------------------------
// One sweep over a resident tile of T codeword rows: every thread owns one column of one 64-row
// chunk and accumulates the plain 512-bit products in a 576-bit accumulator (fr.cuh: wide_mad, 64
// wide multiply-adds per element instead of the 136 of a Montgomery multiplication, no reduction per
// element); one 9-round Montgomery reduction per chunk.  The per-chunk partial sums are folded into
// acc by a second small kernel.  32 bytes (code) / 64 bytes (linear) read per codeword element
// (SURVEY 8d iii): with the multiplier work halved the sweep is bound by HBM.
//   code  : scalars arrive pre-multiplied by 2^288, so the reduction returns sum r_t*e_t directly
//   linear: the reduction returns S*2^-288; one Montgomery multiplication by 2^544 undoes it
constexpr int kCombineChunk = 64;
size_t combine_scratch_elems(int T, int n) { return (size_t)((T + kCombineChunk - 1) / kCombineChunk) * n; }

__device__ __forceinline__ fr_t fr_2p544() {       // 2^544 mod p
    fr_t r;
    r.v[0] = 0x04f2bf4fu; r.v[1] = 0x6dae765eu; r.v[2] = 0xa11298d6u; r.v[3] = 0x347d7b17u;
    r.v[4] = 0xf603929eu; r.v[5] = 0x70f88a97u; r.v[6] = 0x83e5893eu; r.v[7] = 0x0ad4bd84u;
    return r;
}

// r_scaled[t] = r[t] * 2^288 mod p (canonical): montmul by 2^544
__global__ void combine_scale_kernel(const fr_mem *__restrict__ r, int T, fr_mem *__restrict__ r_scaled) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < T) fr_stg(r_scaled + t, fr_reduce_p(fr_mont_mul(fr_ldg(r + t), fr_2p544())));
}

// r_kara[3t .. 3t+2] = halves of r_scaled[t] split at bit 127 and their sum (fr.cuh: kara_split), 16 bytes each
__global__ void combine_split_kernel(const fr_mem *__restrict__ r_scaled, int T, uint4 *__restrict__ r_kara) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    fr_half r0, r1, rs;
    kara_split(fr_ldg(r_scaled + t), r0, r1, rs);
    r_kara[3 * t] = make_uint4(r0.v[0], r0.v[1], r0.v[2], r0.v[3]);
    r_kara[3 * t + 1] = make_uint4(r1.v[0], r1.v[1], r1.v[2], r1.v[3]);
    r_kara[3 * t + 2] = make_uint4(rs.v[0], rs.v[1], rs.v[2], rs.v[3]);
}
__device__ __forceinline__ fr_half half_ldc(const uint4 *p) { const uint4 q = __ldg(p); fr_half h; h.v[0] = q.x; h.v[1] = q.y; h.v[2] = q.z; h.v[3] = q.w; return h; }

// check_code over a resident tile, Karatsuba form: per element three 128 x 128 products (48 wide multiply-adds) instead of
// one 256 x 256 (64); the three partial sums are recombined and reduced once per 64-row chunk (fr.cuh: kara_reduce9)
__global__ void __launch_bounds__(128) combine_code_kara_kernel(const fr_mem *__restrict__ a, long long row_stride, int T, int n,
                                                                const uint4 *__restrict__ r_kara, fr_mem *__restrict__ partial) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int chunk = blockIdx.y;
    if (j >= n) return;
    const int t0 = chunk * kCombineChunk, t1 = min(T, t0 + kCombineChunk);
    fr_kara_wide wl, wh, wm;
    kara_zero(wl); kara_zero(wh); kara_zero(wm);
    const fr_mem *pa = a + (long long)t0 * row_stride + j;
    const uint4 *pr = r_kara + 3 * t0;
    int t = t0;
    for (; t + 1 < t1; t += 2) {                               // two rows in flight per thread
        const fr_t e0 = fr_ldg(pa), e1 = fr_ldg(pa + row_stride);
        fr_half x0, x1, xs;
        kara_split(e0, x0, x1, xs);
        kara_mad(wl, x0, half_ldc(pr)); kara_mad(wh, x1, half_ldc(pr + 1)); kara_mad(wm, xs, half_ldc(pr + 2));
        kara_split(e1, x0, x1, xs);
        kara_mad(wl, x0, half_ldc(pr + 3)); kara_mad(wh, x1, half_ldc(pr + 4)); kara_mad(wm, xs, half_ldc(pr + 5));
        pa += 2 * row_stride;
        pr += 6;
    }
    if (t < t1) {
        fr_half x0, x1, xs;
        kara_split(fr_ldg(pa), x0, x1, xs);
        kara_mad(wl, x0, half_ldc(pr)); kara_mad(wh, x1, half_ldc(pr + 1)); kara_mad(wm, xs, half_ldc(pr + 2));
    }
    fr_stg(partial + (size_t)chunk * n + j, fr_reduce_p(kara_reduce9(wl, wh, wm)));
}

// check_code over a resident tile on the FP64 pipe.  The sweep is bound by the 32x32->64 multiplier (IMAD.WIDE issues at a
// quarter of the lane rate: 6.5e12/s chip-wide, 64 per element = 3.3 TB/s of codeword, half the HBM roofline).  An UNREDUCED
// dot product needs no Montgomery step per element, which is where the double-precision formulation (dpf_mont.cuh) wins:
// 5 x 5 products of 52-bit limbs = 50 DFMA + 25 DADD on the FP64 pipe, their exact halves accumulated as 64-bit integers
// (two adds each on the ALU pipe), the exponent constants taken out once per 64-row chunk, one 9-round Montgomery
// reduction per chunk as before.  r_dpf[5t .. 5t+4] = the 52-bit limbs of r_scaled[t] as doubles.
__global__ void combine_dpf_scalars_kernel(const fr_mem *__restrict__ r_scaled, int T, double *__restrict__ r_dpf) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const fr_t r = fr_ldg(r_scaled + t);
    const dpf_t d = dpf_from_u32(r.v);
#pragma unroll
    for (int j = 0; j < 5; j++) r_dpf[5 * t + j] = d.l[j];
}
__global__ void __launch_bounds__(128) combine_code_dpf_kernel(const fr_mem *__restrict__ a, long long row_stride, int T, int n,
                                                               const double *__restrict__ r_dpf, fr_mem *__restrict__ partial) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int chunk = blockIdx.y;
    if (j >= n) return;
    const int t0 = chunk * kCombineChunk, t1 = min(T, t0 + kCombineChunk);
    unsigned long long col[10];
#pragma unroll
    for (int i = 0; i < 10; i++) col[i] = 0;
    const fr_mem *pa = a + (long long)t0 * row_stride + j;
    const double *pr = r_dpf + 5 * t0;
    // LGR_DPF_ROWS rows in flight per thread: the sweep needs ~7 MB of loads in flight chip-wide to cover HBM latency
    // (Little: 6.5 TB/s x ~1 us), i.e. more than one or two 32-byte loads per thread
    constexpr int RF = 4;
    int t = t0;
    for (; t + RF <= t1; t += RF) {
        fr_t e[RF];
#pragma unroll
        for (int u = 0; u < RF; u++) e[u] = fr_ldg(pa + (long long)u * row_stride);
#pragma unroll
        for (int u = 0; u < RF; u++) {
            const dpf_t x = dpf_from_u32(e[u].v);
            double r[5];
#pragma unroll
            for (int q = 0; q < 5; q++) r[q] = __ldg(pr + 5 * u + q);
#pragma unroll
            for (int p = 0; p < 5; p++)
#pragma unroll
                for (int q = 0; q < 5; q++) dpf_mac(col[p + q], col[p + q + 1], x.l[p], r[q]);
        }
        pa += (long long)RF * row_stride;
        pr += 5 * RF;
    }
    for (; t < t1; t++) {
        const fr_t e = fr_ldg(pa);
        const dpf_t x = dpf_from_u32(e.v);
        double r[5];
#pragma unroll
        for (int q = 0; q < 5; q++) r[q] = __ldg(pr + q);
#pragma unroll
        for (int p = 0; p < 5; p++)
#pragma unroll
            for (int q = 0; q < 5; q++) dpf_mac(col[p + q], col[p + q + 1], x.l[p], r[q]);
        pa += row_stride;
        pr += 5;
    }
    // take out the exponent constants: column c received (number of (p,q) with p+q == c) low halves and
    // (number with p+q == c-1) high halves per row
    const unsigned long long rows = (unsigned long long)(t1 - t0);
    const unsigned long long KH = 0x4670000000000000ull, KL = 0x4330000000000000ull;
#pragma unroll
    for (int c = 0; c < 10; c++) {
        const int nlo = (c <= 4) ? c + 1 : ((c <= 8) ? 9 - c : 0);
        const int nhi = (c >= 1) ? ((c - 1 <= 4) ? c : 10 - c) : 0;
        col[c] -= rows * ((unsigned long long)nlo * KL + (unsigned long long)nhi * KH);
    }
    // sum col[c] * 2^(52 c) as 18 x 32-bit limbs (every column < 2^62)
    uint32_t v[18];
#pragma unroll
    for (int i = 0; i < 18; i++) v[i] = 0;
#pragma unroll
    for (int c = 0; c < 10; c++) {
        const int bit = 52 * c, limb = bit >> 5, sh = bit & 31;
        // col[c] << sh spans up to 3 limbs (62 + 31 bits)
        const unsigned long long lo = col[c] << sh;
        const uint32_t hi = sh ? (uint32_t)(col[c] >> (64 - sh)) : 0u;
        v[limb] = add_cc(v[limb], (uint32_t)lo);
        v[limb + 1] = addc_cc(v[limb + 1], (uint32_t)(lo >> 32));
        if (limb + 2 < 18) v[limb + 2] = addc_cc(v[limb + 2], hi);
#pragma unroll
        for (int i = limb + 3; i < 18; i++) v[i] = addc_cc(v[i], 0);
    }
    fr_stg(partial + (size_t)chunk * n + j, fr_reduce_p(reduce9_rounds(v)));
}

template <bool LINEAR>
__global__ void __launch_bounds__(128) combine_partial_kernel(const fr_mem *__restrict__ a, const fr_mem *__restrict__ b, long long row_stride,
                                                              int T, int n, const fr_mem *__restrict__ r_scaled, fr_mem *__restrict__ partial) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int chunk = blockIdx.y;
    if (j >= n) return;
    const int t0 = chunk * kCombineChunk, t1 = min(T, t0 + kCombineChunk);
    fr_wide w;
    wide_zero(w);
    const fr_mem *pa = a + (long long)t0 * row_stride + j;
    const fr_mem *pb = LINEAR ? b + (long long)t0 * row_stride + j : r_scaled + t0;
    int t = t0;
    if (!LINEAR) {
        for (; t + 3 < t1; t += 4) {                           // four rows in flight per thread (memory-level parallelism)
            const fr_t e0 = fr_ldg(pa), e1 = fr_ldg(pa + row_stride), e2 = fr_ldg(pa + 2 * row_stride), e3 = fr_ldg(pa + 3 * row_stride);
            wide_mad(w, e0, fr_ldc(pb));
            wide_mad(w, e1, fr_ldc(pb + 1));
            wide_mad(w, e2, fr_ldc(pb + 2));
            wide_mad(w, e3, fr_ldc(pb + 3));
            pa += 4 * row_stride;
            pb += 4;
        }
    }
    for (; t + 1 < t1; t += 2) {                               // two rows (four loads for the linear form) in flight per thread
        fr_t e0 = fr_ldg(pa), e1 = fr_ldg(pa + row_stride);
        fr_t m0 = LINEAR ? fr_ldg(pb) : fr_ldc(pb);
        fr_t m1 = LINEAR ? fr_ldg(pb + row_stride) : fr_ldc(pb + 1);
        wide_mad(w, e0, m0);
        wide_mad(w, e1, m1);
        pa += 2 * row_stride;
        pb += LINEAR ? 2 * row_stride : 2;
    }
    if (t < t1) {
        fr_t e0 = fr_ldg(pa);
        fr_t m0 = LINEAR ? fr_ldg(pb) : fr_ldc(pb);
        wide_mad(w, e0, m0);
    }
    fr_t s = wide_reduce9(w);                                  // [0,2p)
    if (LINEAR) s = fr_mont_mul(s, fr_2p544());
    fr_stg(partial + (size_t)chunk * n + j, fr_reduce_p(s));
}
// check_quadratic over resident tiles (nonbatch_context.hpp:771-780): partial = sum_t r_t*(X_t*Y_t - Z_t)
//   = REDC9( sum_t (r_t X_t mod p) * Y_t ) * 2^288  -  REDC9( sum_t (r_t 2^288) * Z_t )
// one Montgomery multiplication + two wide products per triple element, 96 bytes read
__global__ void __launch_bounds__(128) combine_quad_kernel(const fr_mem *__restrict__ x, const fr_mem *__restrict__ y, const fr_mem *__restrict__ z,
                                                           long long row_stride, int T, int n, const fr_mem *__restrict__ r_mont,
                                                           const fr_mem *__restrict__ r_scaled, fr_mem *__restrict__ partial,
                                                           const uint32_t *__restrict__ row_idx) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int chunk = blockIdx.y;
    if (j >= n) return;
    const int t0 = chunk * kCombineChunk, t1 = min(T, t0 + kCombineChunk);
    fr_wide wp, wz;
    wide_zero(wp); wide_zero(wz);
    for (int t = t0; t < t1; t++) {
        const long long off = (long long)(row_idx ? row_idx[t] : (uint32_t)t) * row_stride + j;     // row_idx: triples scattered over a tile
        fr_t rx = fr_reduce_p(fr_mont_mul(fr_ldg(x + off), fr_ldc(r_mont + t)));     // r_t * X, canonical
        wide_mad(wp, rx, fr_ldg(y + off));
        wide_mad(wz, fr_ldg(z + off), fr_ldc(r_scaled + t));
    }
    fr_t a = fr_reduce_p(fr_mont_mul(wide_reduce9(wp), fr_2p544()));
    fr_t b = fr_reduce_p(wide_reduce9(wz));
    fr_stg(partial + (size_t)chunk * n + j, fr_sub(a, b));
}
// r_mont[t] = r[t] * 2^256 mod p
__global__ void combine_mont_kernel(const fr_mem *__restrict__ r, int T, fr_mem *__restrict__ r_mont) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < T) fr_stg(r_mont + t, fr_reduce_p(fr_mont_mul(fr_ldg(r + t), fr_R2())));
}
// partial sums -> acc, two levels so that narrow matrices (n = 1024, thousands of chunks) do not
// serialise thousands of dependent loads per thread: level 1 folds chunk c into group c % G (in place
// on the first G rows of `partial`), level 2 adds the G group sums to acc
constexpr int kFoldGroups = 32;
__global__ void combine_fold1_kernel(fr_mem *__restrict__ partial, int chunks, int n, int G) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int g = blockIdx.y;
    if (j >= n) return;
    fr_t s = fr_ldg(partial + (size_t)g * n + j);
    for (int c = g + G; c < chunks; c += G) s = fr_add(s, fr_ldg(partial + (size_t)c * n + j));
    fr_stg(partial + (size_t)g * n + j, s);
}
__global__ void combine_fold_kernel(const fr_mem *__restrict__ partial, int chunks, int n, fr_mem *acc) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    fr_t s = fr_ldg(acc + j);
    for (int c = 0; c < chunks; c++) s = fr_add(s, fr_ldg(partial + (size_t)c * n + j));
    fr_stg(acc + j, s);
}
static void launch_fold(fr_mem *partial, int chunks, int n, fr_mem *acc, cudaStream_t st) {
    if (chunks > 2 * kFoldGroups) {
        combine_fold1_kernel<<<dim3((n + 127) / 128, kFoldGroups), 128, 0, st>>>(partial, chunks, n, kFoldGroups);
        chunks = kFoldGroups;
    }
    combine_fold_kernel<<<(n + 127) / 128, 128, 0, st>>>(partial, chunks, n, acc);
}

cudaError_t launch_combine_code(const fr_mem *tile, long long row_stride, int T, int n, const fr_mem *r_raw, fr_mem *acc,
                                fr_mem *scratch, size_t scratch_elems, cudaStream_t st) {
    if (T <= 0 || n <= 0) return cudaSuccess;
    const int chunks = (T + kCombineChunk - 1) / kCombineChunk;
    if (scratch_elems < (size_t)chunks * n + 3 * (size_t)T) return cudaErrorInvalidValue;
    dim3 grid((n + 127) / 128, chunks);
    fr_mem *r_scaled = scratch + (size_t)chunks * n;
    combine_scale_kernel<<<(T + 127) / 128, 128, 0, st>>>(r_raw, T, r_scaled);
    // Three formulations of the same sweep, measured on B200 (k = 256, 65536-row tile, 2.15 GB; ncu gpu__time_duration,
    // profiles/r02_combine_formulations.md):  plain 64 IMAD.WIDE per element 561 us (3.83 TB/s, 58 % of HBM) -- the default;
    // Karatsuba (48 IMAD.WIDE + splits and carries) 635 us;  FP64 pipe (50 DFMA + 25 DADD + 100 integer adds) 606 us.
    // Fewer multiplier operations did not pay: the extra ALU work costs more issue slots than the multiplies it saves.
    // LGR_COMBINE_CODE = plain | kara | dpf selects one (the parity suite runs all three).
    static const char *mode_env = getenv("LGR_COMBINE_CODE");
    static const int mode = !mode_env ? 0 : (!strcmp(mode_env, "dpf") ? 2 : (!strcmp(mode_env, "kara") ? 1 : 0));
    if (mode == 2) {
        double *r_dpf = reinterpret_cast<double *>(r_scaled + T);            // 40 bytes per row, after the T scaled scalars
        combine_dpf_scalars_kernel<<<(T + 127) / 128, 128, 0, st>>>(r_scaled, T, r_dpf);
        combine_code_dpf_kernel<<<grid, 128, 0, st>>>(tile, row_stride, T, n, r_dpf, scratch);
    } else if (mode == 1) {
        uint4 *r_kara = reinterpret_cast<uint4 *>(r_scaled + T);             // 48 bytes per row, after the T scaled scalars
        combine_split_kernel<<<(T + 127) / 128, 128, 0, st>>>(r_scaled, T, r_kara);
        combine_code_kara_kernel<<<grid, 128, 0, st>>>(tile, row_stride, T, n, r_kara, scratch);
    } else {
        combine_partial_kernel<false><<<grid, 128, 0, st>>>(tile, nullptr, row_stride, T, n, r_scaled, scratch);
    }
    launch_fold(scratch, chunks, n, acc, st);
    return cudaGetLastError();
}
cudaError_t launch_combine_linear(const fr_mem *a, const fr_mem *b, long long row_stride, int T, int n, fr_mem *acc,
                                  fr_mem *scratch, size_t scratch_elems, cudaStream_t st) {
    if (T <= 0 || n <= 0) return cudaSuccess;
    const int chunks = (T + kCombineChunk - 1) / kCombineChunk;
    if (scratch_elems < (size_t)chunks * n) return cudaErrorInvalidValue;
    dim3 grid((n + 127) / 128, chunks);
    combine_partial_kernel<true><<<grid, 128, 0, st>>>(a, b, row_stride, T, n, nullptr, scratch);
    launch_fold(scratch, chunks, n, acc, st);
    return cudaGetLastError();
}

cudaError_t launch_combine_quad(const fr_mem *x, const fr_mem *y, const fr_mem *z, long long row_stride, int T, int n, const fr_mem *r_raw,
                                fr_mem *acc, fr_mem *scratch, size_t scratch_elems, cudaStream_t st, const uint32_t *row_idx) {
    if (T <= 0 || n <= 0) return cudaSuccess;
    const int chunks = (T + kCombineChunk - 1) / kCombineChunk;
    if (scratch_elems < (size_t)chunks * n + 2 * (size_t)T) return cudaErrorInvalidValue;
    dim3 grid((n + 127) / 128, chunks);
    fr_mem *r_scaled = scratch + (size_t)chunks * n, *r_mont = r_scaled + T;
    combine_scale_kernel<<<(T + 127) / 128, 128, 0, st>>>(r_raw, T, r_scaled);
    combine_mont_kernel<<<(T + 127) / 128, 128, 0, st>>>(r_raw, T, r_mont);
    combine_quad_kernel<<<grid, 128, 0, st>>>(x, y, z, row_stride, T, n, r_mont, r_scaled, scratch, row_idx);
    launch_fold(scratch, chunks, n, acc, st);
    return cudaGetLastError();
}

// ---- coset 0 of the large-k encoder --------------------------------------------------------------
// out[row][4m] = rows[row][c*m mod k] reduced to [0,p): w_n^4 = w_k^c, so the codeword positions 4m are
// the message evaluations themselves (api.cu: find_sys_mul).  Writes are 32-byte sectors 128 bytes apart;
// the other three cosets fill the gaps right after, while the lines are still in L2.
__global__ void __launch_bounds__(256) sys_copy_kernel(const fr_mem *__restrict__ rows, long long row_stride, const __grid_constant__ CodewordSink sink,
                                                       long long total, int logk, uint32_t c) {
    const uint32_t mask = (1u << logk) - 1u;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i >> logk;
        const uint32_t m = (uint32_t)i & mask;
        const fr_t x = fr_ldg(rows + row * row_stride + ((m * c) & mask));
        fr_stg(sink_at(sink, row, (int)(4u * m)), fr_reduce_p(fr_reduce_2p(fr_reduce_2p(x))));
    }
}
cudaError_t launch_sys_copy(const fr_mem *rows, long long row_stride, const CodewordSink &sink, int R, int logk, uint32_t c, cudaStream_t st) {
    if (R <= 0) return cudaSuccess;
    const long long total = (long long)R << logk;
    const int grid = (int)std::min<long long>((total + 255) / 256, 148 * 8);
    sys_copy_kernel<<<grid, 256, 0, st>>>(rows, row_stride, sink, total, logk, c);
    return cudaGetLastError();
}

// ---- stage-3 openings over a resident codeword tile ------------------------------------------------
// out[t][s] = tile[t][idx[s]] (sample_gather, kernels.wgsl.in:541-549, for T rows at once; the staging
// layout [row][sample][8 x u32] is the proof's host_samplings, nonbatch_context.hpp:906-933)
__global__ void __launch_bounds__(256) gather_rows_kernel(const fr_mem *__restrict__ tile, long long row_stride, int T, const uint32_t *__restrict__ idx,
                                                          int count, fr_mem *__restrict__ out) {
    const long long total = (long long)T * count;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long t = i / count;
        const int s = (int)(i - t * count);
        fr_stg(out + i, fr_ldg(tile + t * row_stride + idx[s]));
    }
}
cudaError_t launch_gather_rows(const fr_mem *tile, long long row_stride, int T, const uint32_t *idx, int count, fr_mem *out, cudaStream_t st) {
    if (T <= 0 || count <= 0) return cudaSuccess;
    const long long total = (long long)T * count;
    const int grid = (int)std::min<long long>((total + 255) / 256, 148 * 8);
    gather_rows_kernel<<<grid, 256, 0, st>>>(tile, row_stride, T, idx, count, out);
    return cudaGetLastError();
}

}  // namespace lgr

// Element-wise field kernels and the stage-2 row combiners.
//
// Replaces the Eltwise* entry points of shader/kernels.wgsl.in:325-549 (one dispatch per
// operation, Barrett reduction = 3 full + 1 low-half 256-bit products per element)
// and the check_code / check_linear / check_quadratic schedules of
// include/zkp/nonbatch_context.hpp:756-780.  Canonical in, canonical out, bit-exact with the WGSL
// for canonical inputs.  A canonical product x*y is two Montgomery multiplications
// (x*y/R, then *R^2/R); a product with a host constant c is one (c is pre-multiplied by R).
#include "kernels.h"
#include "ntt.cuh"

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
// One sweep over a resident tile of T codeword rows: every thread owns one column of one row
// chunk, accumulates lazily in [0,2p), and the per-chunk partial sums are folded into acc by a
// second small kernel.  32 bytes read per codeword element (SURVEY 8d iii).
constexpr int kCombineChunk = 32;
size_t combine_scratch_elems(int T, int n) { return (size_t)((T + kCombineChunk - 1) / kCombineChunk) * n; }

template <bool LINEAR>
__global__ void __launch_bounds__(128) combine_partial_kernel(const fr_mem *__restrict__ a, const fr_mem *__restrict__ b, long long row_stride,
                                                              int T, int n, const fr_mem *__restrict__ r_mont, fr_mem *__restrict__ partial) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int chunk = blockIdx.y;
    if (j >= n) return;
    const int t0 = chunk * kCombineChunk, t1 = min(T, t0 + kCombineChunk);
    fr_t s = fr_zero();
    for (int t = t0; t < t1; t++) {
        fr_t e = fr_ldg(a + (long long)t * row_stride + j);
        fr_t m;
        if (LINEAR) m = fr_mul_canon(e, fr_ldg(b + (long long)t * row_stride + j));
        else m = fr_mont_mul(e, fr_ldc(r_mont + t));           // e * r, [0,2p)
        s = fr_add_lazy(s, m);
    }
    fr_stg(partial + (size_t)chunk * n + j, fr_reduce_p(s));
}
__global__ void combine_fold_kernel(const fr_mem *__restrict__ partial, int chunks, int n, fr_mem *acc) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    fr_t s = fr_ldg(acc + j);
    for (int c = 0; c < chunks; c++) s = fr_add(s, fr_ldg(partial + (size_t)c * n + j));
    fr_stg(acc + j, s);
}

cudaError_t launch_combine_code(const fr_mem *tile, long long row_stride, int T, int n, const fr_mem *r_mont, fr_mem *acc,
                                fr_mem *scratch, size_t scratch_elems, cudaStream_t st) {
    if (T <= 0 || n <= 0) return cudaSuccess;
    const int chunks = (T + kCombineChunk - 1) / kCombineChunk;
    if (scratch_elems < (size_t)chunks * n) return cudaErrorInvalidValue;
    dim3 grid((n + 127) / 128, chunks);
    combine_partial_kernel<false><<<grid, 128, 0, st>>>(tile, nullptr, row_stride, T, n, r_mont, scratch);
    combine_fold_kernel<<<(n + 127) / 128, 128, 0, st>>>(scratch, chunks, n, acc);
    return cudaGetLastError();
}
cudaError_t launch_combine_linear(const fr_mem *a, const fr_mem *b, long long row_stride, int T, int n, fr_mem *acc,
                                  fr_mem *scratch, size_t scratch_elems, cudaStream_t st) {
    if (T <= 0 || n <= 0) return cudaSuccess;
    const int chunks = (T + kCombineChunk - 1) / kCombineChunk;
    if (scratch_elems < (size_t)chunks * n) return cudaErrorInvalidValue;
    dim3 grid((n + 127) / 128, chunks);
    combine_partial_kernel<true><<<grid, 128, 0, st>>>(a, b, row_stride, T, n, nullptr, scratch);
    combine_fold_kernel<<<(n + 127) / 128, 128, 0, st>>>(scratch, chunks, n, acc);
    return cudaGetLastError();
}

}  // namespace lgr

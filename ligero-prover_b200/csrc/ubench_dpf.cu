// Micro-benchmarks of the FP64-pipe Montgomery multiplication (dpf_mont.cuh) alone and next to the IMAD one (fr.cuh):
// VERDICT r01 item 5 -- "measure, don't cost, a DFMA (52-bit-limb) Montgomery multiply".
#include "dpf_mont.cuh"
#include "kernels.h"
#include "ntt.cuh"

namespace lgr {

__device__ __forceinline__ dpf_t dpf_seed(uint32_t s) {
    dpf_t x;
#pragma unroll
    for (int j = 0; j < 5; j++) x.l[j] = dpf_from_int((((unsigned long long)(s * 0x9E3779B1u + j * 0x85EBCA77u) << 19) & 0x000FFFFFFFFFFFFFull));
    x.l[4] = dpf_from_int(((s * 2654435761u) & 0xFFFFFFFFFFull));            // < 2^40: value < p
    return x;
}

// 4 independent FP64-pipe Montgomery multiplications per iteration
__global__ void __launch_bounds__(256) ubench_dpf_kernel(uint32_t *out, int iters) {
    dpf_t x[4];
    const dpf_t w = dpf_seed(threadIdx.x + 1);
#pragma unroll
    for (int q = 0; q < 4; q++) x[q] = dpf_seed(threadIdx.x * 4 + q + blockIdx.x);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int q = 0; q < 4; q++) x[q] = dpf_mont_mul(x[q], w);
    }
    double s = 0;
#pragma unroll
    for (int q = 0; q < 4; q++)
#pragma unroll
        for (int j = 0; j < 5; j++) s += x[q].l[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)__double2ll_rz(s);
}

// warps alternate between the two formulations: IMAD_PER_DPF integer warps for every FP64 warp (warp-uniform branch)
template <int IMAD_PER_DPF>
__global__ void __launch_bounds__(256) ubench_mixed_kernel(uint32_t *out, int iters) {
    const int warp = threadIdx.x >> 5;
    uint32_t res;
    if (warp % (IMAD_PER_DPF + 1) == IMAD_PER_DPF) {
        dpf_t x[4];
        const dpf_t w = dpf_seed(threadIdx.x + 1);
#pragma unroll
        for (int q = 0; q < 4; q++) x[q] = dpf_seed(threadIdx.x * 4 + q + blockIdx.x);
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int q = 0; q < 4; q++) x[q] = dpf_mont_mul(x[q], w);
        }
        double s = 0;
#pragma unroll
        for (int q = 0; q < 4; q++)
#pragma unroll
            for (int j = 0; j < 5; j++) s += x[q].l[j];
        res = (uint32_t)__double2ll_rz(s);
    } else {
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
        res = 0;
#pragma unroll
        for (int q = 0; q < 4; q++)
#pragma unroll
            for (int i = 0; i < 8; i++) res ^= x[q].v[i];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = res;
}

// Do the encoder's multiplier work and the column hash overlap on one SM?  Even warps run IMAD Montgomery multiplications,
// odd warps SHA-256 compressions; a side with iteration count 0 leaves at once.  Timed alone and together by lgru_overlap.
__device__ __forceinline__ uint32_t ubd_rotr(uint32_t x, int n) { return __funnelshift_r(x, x, n); }
__global__ void __launch_bounds__(256) ubench_mont_sha_kernel(uint32_t *out, int iters_mont, int iters_sha) {
    const int warp = threadIdx.x >> 5;
    uint32_t res = 0;
    if (warp & 1) {
        uint32_t st[8], w[16];
#pragma unroll
        for (int i = 0; i < 8; i++) st[i] = threadIdx.x * 0x01000193u + i;
        for (int it = 0; it < iters_sha; it++) {
#pragma unroll
            for (int i = 0; i < 16; i++) w[i] = st[i & 7] + i + it;
            uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll
            for (int i = 0; i < 64; i++) {
                uint32_t wi;
                if (i < 16) wi = w[i];
                else {
                    const uint32_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
                    wi = w[i & 15] + (ubd_rotr(w15, 7) ^ ubd_rotr(w15, 18) ^ (w15 >> 3)) + w[(i + 9) & 15] + (ubd_rotr(w2, 17) ^ ubd_rotr(w2, 19) ^ (w2 >> 10));
                    w[i & 15] = wi;
                }
                const uint32_t t1 = h + (ubd_rotr(e, 6) ^ ubd_rotr(e, 11) ^ ubd_rotr(e, 25)) + ((e & f) ^ (~e & g)) + 0x428a2f98u * (i + 1) + wi;
                const uint32_t t2 = (ubd_rotr(a, 2) ^ ubd_rotr(a, 13) ^ ubd_rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
                h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
            }
            st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
        }
#pragma unroll
        for (int i = 0; i < 8; i++) res ^= st[i];
    } else {
        fr_t x[4], w;
#pragma unroll
        for (int i = 0; i < 8; i++) w.v[i] = (threadIdx.x + 1) * 0x9E3779B1u + i * 0x85EBCA77u;
        w.v[7] &= 0x0FFFFFFFu;
#pragma unroll
        for (int q = 0; q < 4; q++) { x[q] = w; x[q].v[0] += q + blockIdx.x; }
        for (int it = 0; it < iters_mont; it++) {
#pragma unroll
            for (int q = 0; q < 4; q++) x[q] = fr_mont_mul(x[q], w);
        }
#pragma unroll
        for (int q = 0; q < 4; q++)
#pragma unroll
            for (int i = 0; i < 8; i++) res ^= x[q].v[i];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = res;
}
cudaError_t launch_ubench_mont_sha(uint32_t *out, int iters_mont, int iters_sha, int blocks, cudaStream_t st) {
    ubench_mont_sha_kernel<<<blocks, 256, 0, st>>>(out, iters_mont, iters_sha);
    return cudaGetLastError();
}

// out[i] = canonical(a[i] * b[i] * 2^-260 mod p)
__global__ void dpf_mul_kernel(const uint32_t *a, const uint32_t *b, uint32_t *out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const dpf_t r = dpf_mont_mul(dpf_from_u32(a + 8 * i), dpf_from_u32(b + 8 * i));
    fr_t x;
    dpf_to_u32(x.v, r);
    x = fr_reduce_p(fr_reduce_2p(x));
#pragma unroll
    for (int j = 0; j < 8; j++) out[8 * i + j] = x.v[j];
}

cudaError_t launch_ubench_dpf(int which, uint32_t *out, int iters, int blocks, int threads, cudaStream_t st) {
    if (which == 6) ubench_dpf_kernel<<<blocks, threads, 0, st>>>(out, iters);
    else if (which == 7) ubench_mixed_kernel<1><<<blocks, threads, 0, st>>>(out, iters);
    else if (which == 8) ubench_mixed_kernel<3><<<blocks, threads, 0, st>>>(out, iters);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}
cudaError_t launch_dpf_mul(const uint32_t *a, const uint32_t *b, uint32_t *out, int n, cudaStream_t st) {
    dpf_mul_kernel<<<(n + 127) / 128, 128, 0, st>>>(a, b, out, n);
    return cudaGetLastError();
}

}  // namespace lgr

// Fused Reed-Solomon row encoder for k = 2^LOGK <= 2048.
//
// Reference semantics (src/webgpu/engine.cpp:755-770 encode_ntt_device): c = iNTT_k(row) on the
// w_k domain, codeword e = NTT_n(c || 0...0) on the w_n domain, n = 4k: 15 dispatches, every stage
// a pass over global memory, 3/4 of the forward input known-zero.
//
// Here one row never leaves the SM between input and codeword:
//   1. inverse transform (unscaled), coefficients in shared memory
//   2. the 4k-point forward transform is split into four k-point coset transforms:
//        e[4m + r] = sum_i (c_i * w_n^(r*i)) * (w_n^4)^(i*m),   r = 0..3
//      so the zero padding is never touched; the twist table also carries the 1/k of the inverse
//   3. every transform runs as a DIT (bit-reversed in -> natural out), the cosets are written interleaved;
//      coset r = 0 is a permutation of the message (w_n^4 is a power of w_k) and is copied, not computed.
// Per witness element: (log2(k)-1)/2 + 3 + 3*(log2(k)-1)/2 Montgomery multiplications.
#include "kernels.h"
#include "ntt.cuh"

#ifndef LGR_ENC_LOGK
#error "compile with -DLGR_ENC_LOGK=<3..11> (see Makefile): one translation unit per row size keeps the build parallel"
#endif

namespace lgr {

// CTA = 128 threads (4 warps) working on 128/(k/8) rows at once, in lock step (__syncthreads between
// passes), three CTAs per SM.  The kernel is ~220 KiB of straight-line code (every butterfly inlines
// a 170-instruction Montgomery multiplication); with per-row warps drifting apart a third of all stall
// samples were instruction-cache misses (ncu: stall_no_inst 32%).  Measured on B200, 16384 rows of
// k = 256: per-row sync 1.92 ms; 384-thread CTAs in lock step 1.86 ms; 128-thread CTAs in lock step
// 1.63 ms (three instruction streams per SM, and the CTAs hide each other's load phases).
template <int LOGK> struct enc_cfg {
    static constexpr int K = 1 << LOGK;
    static constexpr int TL = K / 8;                                    // threads per row
#ifndef LGR_ENC_THREADS
#define LGR_ENC_THREADS 128
#endif
    static constexpr int THREADS = (TL >= 256) ? TL : ((TL > LGR_ENC_THREADS) ? TL : LGR_ENC_THREADS);
    static constexpr int SLOTS = THREADS / TL;                          // rows in flight per CTA
    static constexpr size_t SMEM = (size_t)SLOTS * 2 * K * 32;          // coefficients + work array per slot
};

template <int LOGK>
__global__ void __launch_bounds__(enc_cfg<LOGK>::THREADS, (enc_cfg<LOGK>::SMEM <= 72 * 1024) ? 3 : ((enc_cfg<LOGK>::SMEM <= 110 * 1024) ? 2 : 1)) encode_rows_kernel(
    const fr_mem *__restrict__ rows_in, long long in_row_stride, const __grid_constant__ CodewordSink sink, int R, const EncodeTables t) {
    using cfg = enc_cfg<LOGK>;
    constexpr int K = cfg::K, TL = cfg::TL;
    extern __shared__ __align__(32) unsigned char smem_raw[];
    const int slot = threadIdx.x / TL, tl = threadIdx.x % TL;
    fr_mem *C = reinterpret_cast<fr_mem *>(smem_raw) + (size_t)slot * 2 * K;
    fr_mem *W = C + K;
    const long long row = (long long)blockIdx.x * cfg::SLOTS + slot;
    const bool active = row < R;
    auto sync = [] { __syncthreads(); };

    // All transforms of a row -- the inverse one and the cosets -- run through ONE decimation-in-time body (bit-reversed
    // in, natural out), selected by run-time flags: a second, decimation-in-frequency body for the inverse transform
    // doubled the straight-line code to 220 KB, more than the instruction cache holds (ncu: 20 % of the stall samples
    // were `no_inst`).
    //   pass t = 0      : c = k * iNTT_k(row): the first pass gathers row[bitrev(q)] from the bulk-loaded row in W, twiddles
    //                     w_k^-1, result to C in natural order, values in [0,4p); 1/k is folded into the twists
    //   pass t = 1..    : coset r: the first pass reads C[bitrev(q)] and applies the twist w_n^(r*bitrev(q))/k, the last
    //                     pass canonicalises and writes e[4m + r] straight to global memory
    const fr_mem *src = rows_in + (active ? row : 0) * in_row_stride;
    const long long drow = active ? row : 0;
    const bool sys = t.sys_mul != 0;
    const int ntrans = sys ? 4 : 5;
    // The message row comes in as ONE bulk copy (TMA, cp.async.bulk global -> shared, k*32 contiguous bytes) issued by the
    // first thread of the row's slot and signalled on an mbarrier; the row then sits in W, where the inverse transform
    // gathers it in bit-reversed order and the copied coset reads it back.  In-place encodes (rows aliasing codewords) are
    // safe: the whole row is in shared memory before the first codeword element is written.
    __shared__ __align__(8) unsigned long long row_bar[cfg::SLOTS];
    if (tl == 0) {
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&row_bar[slot]);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (active) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"((uint32_t)(K * 32)) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"((uint32_t)__cvta_generic_to_shared(W)), "l"(src), "r"((uint32_t)(K * 32)), "r"(bar) : "memory");
        }
    }
    __syncthreads();                                                   // barrier initialised before anybody polls it
    if (active) {
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&row_bar[slot]);
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "ROW_WAIT:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
            "@p bra ROW_DONE;\n\t"
            "bra ROW_WAIT;\n\t"
            "ROW_DONE:\n\t}" :: "r"(bar) : "memory");
    }
#pragma unroll 1
    for (int tr = 0; tr < ntrans; tr++) {
        const bool inv = tr == 0;
        const int r = sys ? tr : tr - 1;                                       // coset index of passes tr >= 1
        const fr_mem *tw = inv ? t.inv_k : t.fwd_c;
        const fr_mem *tw_r = t.twist + (inv ? 0 : r) * K;
        // the inverse transform works in C (its first pass reads the raw row from W and leaves it there); the cosets work in W
        ntt_passes_io<LOGK, 0, false>(inv ? C : W, tl, tw, 1, sync,
            [&](int q) -> fr_t {
                const int i = (int)bitrev((uint32_t)q, LOGK);
                if (inv) return active ? fr_lds(W + i) : fr_zero();
                return fr_mont_mul(fr_lds(C + i), fr_ldc(tw_r + q));
            },
            [&](int m, const fr_t &x) {
                if (inv) fr_sts(C + m, x);
                else if (active) fr_stg(sink_at(sink, drow, 4 * m + r), fr_canon4(x));
            });
        sync();
        // coset r = 0 is the message itself: w_n^4 and w_k generate the same group of k-th roots of unity,
        // w_n^4 = w_k^c (c = 2^61-1 mod k = k-1 for the reference's roots, src/bn254.cpp:36-43,51-64), so
        // e[4m] = U((w_n^4)^m) = U(w_k^(c m)) = row[c m mod k]: a permuted copy, no arithmetic.
        if (inv && sys) {
            if (active) {
#pragma unroll
                for (int g = 0; g < 8; g++) {
                    const int m = tl + g * TL;
                    const fr_t x = fr_lds(W + (int)(((unsigned)m * (unsigned)t.sys_mul) & (unsigned)(K - 1)));
                    fr_stg(sink_at(sink, drow, 4 * m), fr_reduce_p(fr_reduce_2p(fr_reduce_2p(x))));   // any 256-bit input -> [0,p), as the transforms would
                }
            }
            sync();
        }
    }
}

template <int LOGK>
static cudaError_t launch_enc(const fr_mem *rows_in, long long in_row_stride, const CodewordSink &sink, int R,
                              const EncodeTables &t, cudaStream_t st) {
    using cfg = enc_cfg<LOGK>;
    // the opt-in above 48 KiB of dynamic shared memory is stored per device: cache it per device id
    static bool attr_done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(encode_rows_kernel<LOGK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg::SMEM);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
    const int grid = (R + cfg::SLOTS - 1) / cfg::SLOTS;
    encode_rows_kernel<LOGK><<<grid, cfg::THREADS, cfg::SMEM, st>>>(rows_in, in_row_stride, sink, R, t);
    return cudaGetLastError();
}

#define LGR_CAT2(a, b) a##b
#define LGR_CAT(a, b) LGR_CAT2(a, b)
cudaError_t LGR_CAT(launch_encode_rows_, LGR_ENC_LOGK)(const fr_mem *rows_in, long long in_row_stride, const CodewordSink &sink,
                                                       int R, const EncodeTables &t, cudaStream_t st) {
    return launch_enc<LGR_ENC_LOGK>(rows_in, in_row_stride, sink, R, t, st);
}

}  // namespace lgr

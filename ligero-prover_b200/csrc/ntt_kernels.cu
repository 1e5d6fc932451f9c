// Generic tile NTT kernel: every power-of-two transform the executor exposes
// (ntt_forward/inverse_{k,2k,n}, include/wgpu.hpp:128-134; the encode/decode compositions of
// src/webgpu/engine.cpp:755-796 for large k; BASELINE configs 1 and 2) is one or two launches of
// this kernel.  Sizes above 2^11 use a four-step split N = N1*N2 (column pass with twist through
// a scratch buffer, then row pass) instead of the reference's log2(N)-9 global radix-2 dispatches
// (engine.cpp:854-858,960-964).
#include "kernels.h"
#include "ntt.cuh"

#ifndef LGR_NTT_LOGM
#error "compile with -DLGR_NTT_LOGM=<1..11> (see Makefile)"
#endif

namespace lgr {

#ifndef LGR_NTT_MINBLOCKS
#define LGR_NTT_MINBLOCKS 2          // tuning knob (A/B builds): minimum resident 256-thread CTAs per SM => register cap
#endif
template <int LOGM>
__global__ void __launch_bounds__(256, LGR_NTT_MINBLOCKS) ntt_tile_kernel(const __grid_constant__ NttTileParams p) {
    extern __shared__ __align__(32) unsigned char smem_raw[];
    fr_mem *sm = reinterpret_cast<fr_mem *>(smem_raw);
    constexpr int M = 1 << LOGM;
    constexpr int TL = (M >= 8) ? (M / 8) : 1;
    const int C = p.lanes_per_cta;
    const int lane0 = blockIdx.x * C;
    const int nl = min(C, p.total_lanes - lane0);
    const int total = C << LOGM;

    // ---- fill: coalesced global read, bit-reversed placement in shared memory ----
    const bool lane_fast_in = p.in_lane_stride < p.in_point_stride;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        int c, m;
        if (lane_fast_in) { c = idx % C; m = idx / C; } else { c = idx >> LOGM; m = idx & (M - 1); }
        if (c < nl) {
            const int L = lane0 + c;
            const int outer = L / p.lanes_inner, inner = L % p.lanes_inner;
            const long long off = (long long)inner * p.in_lane_stride + (long long)m * p.in_point_stride;
            const int orow = p.in_outer_div ? outer / p.in_outer_div : outer;
            fr_t x = fr_ldg(p.in + (long long)orow * p.in_outer_stride + off);
            if (p.in_twist)                                             // coset twist (and 1/k) on the way in
                x = fr_mont_mul(x, fr_ldc(p.in_twist + (long long)(outer - orow * p.in_outer_div) * p.in_twist_sub_stride + off));
            fr_sts(sm + (c << LOGM) + bitrev(m, LOGM), x);
        }
    }
    __syncthreads();

    // ---- transform ----
    {
        const int c = threadIdx.x / TL, tl = threadIdx.x % TL;
        ntt_dit<LOGM>(sm + (c << LOGM), tl, p.tw, p.tws, [] { __syncthreads(); });
    }

    // ---- drain: optional twist / scale / canonicalisation, coalesced global write ----
    const bool lane_fast_out = p.out_lane_stride < p.out_point_stride;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        int c, m;
        if (lane_fast_out) { c = idx % C; m = idx / C; } else { c = idx >> LOGM; m = idx & (M - 1); }
        if (c < nl) {
            const int L = lane0 + c;
            const int outer = L / p.lanes_inner, inner = L % p.lanes_inner;
            const int orow = p.out_outer_div ? outer / p.out_outer_div : outer;
            const int osub = p.out_outer_div ? outer - orow * p.out_outer_div + p.out_sub_base : 0;
            fr_t x = fr_lds(sm + (c << LOGM) + m);                      // [0,4p)
            if (p.twist_full) {
                x = fr_mont_mul(x, fr_ldc(p.twist_full + (unsigned long long)inner * (unsigned)m));   // [0,2p)
            } else if (p.twist_lo) {
                const unsigned long long e = (unsigned long long)inner * (unsigned)m;
                fr_t w = fr_mont_mul(fr_ldc(p.twist_hi + (e >> p.twist_shift)), fr_ldc(p.twist_lo + (e & ((1ull << p.twist_shift) - 1))));
                x = fr_mont_mul(x, fr_reduce_p(w));                     // [0,2p)
            }
            if (p.scale) x = fr_mont_mul(x, fr_ldc(p.scale));           // [0,2p)
            if (p.canon) x = fr_canon4(x);
            if (p.sink.nslabs)                                          // coset mode, slab-major codewords (possibly peer memory)
                fr_stg(sink_at(p.sink, orow, (int)((long long)osub * p.out_sub_stride + (long long)inner * p.out_lane_stride + (long long)m * p.out_point_stride)), x);
            else
                fr_stg(p.out + (long long)orow * p.out_outer_stride + (long long)osub * p.out_sub_stride +
                           (long long)inner * p.out_lane_stride + (long long)m * p.out_point_stride, x);
        }
    }
}

template <int LOGM>
static cudaError_t launch_one(const NttTileParams &p, cudaStream_t st) {
    constexpr int M = 1 << LOGM;
    constexpr int TL = (M >= 8) ? (M / 8) : 1;
    const int threads = p.lanes_per_cta * TL;
    const size_t smem = (size_t)p.lanes_per_cta * M * 32;
    if (threads > 256 || smem > 64 * 1024) return cudaErrorInvalidValue;
    // the opt-in above 48 KiB of dynamic shared memory is stored per device: cache it per device id
    static bool attr_done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(ntt_tile_kernel<LOGM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
    const int grid = (p.total_lanes + p.lanes_per_cta - 1) / p.lanes_per_cta;
    ntt_tile_kernel<LOGM><<<grid, threads, smem, st>>>(p);
    return cudaGetLastError();
}

#define LGR_CAT2(a, b) a##b
#define LGR_CAT(a, b) LGR_CAT2(a, b)
cudaError_t LGR_CAT(launch_ntt_tile_, LGR_NTT_LOGM)(const NttTileParams &p, cudaStream_t st) { return launch_one<LGR_NTT_LOGM>(p, st); }

}  // namespace lgr

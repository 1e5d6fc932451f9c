// Latency-oriented variant of the tile NTT (ntt_kernels.cu), same NttTileParams contract, for SMALL jobs: the per-row
// schedule of the reference's stage contexts (one row per callback, include/zkp/nonbatch_context.hpp:445-468) hands the
// device one 8192-point encode at a time.  The throughput kernel gives each thread 8 elements and ~140 KB of straight-line
// code per size; with only 8-24 CTAs in flight every instruction is an instruction-cache miss and a launch takes 24-42 us
// (profiles/r02_per_row_launches.csv).  Here a lane of M points is worked by M/2 threads, ONE butterfly per thread per
// stage in a rolled loop (a few KB of code, resident after the first stage), so a launch is bound by ~log2(M) dependent
// Montgomery multiplications plus barriers.  Selected by launch_ntt_tile for jobs of at most 2^16 points.
#include <cstdlib>
#include "kernels.h"
#include "ntt.cuh"

namespace lgr {

__global__ void __launch_bounds__(1024, 1) ntt_lat_kernel(const __grid_constant__ NttTileParams p) {
    extern __shared__ __align__(32) unsigned char smem_raw[];
    fr_mem *sm = reinterpret_cast<fr_mem *>(smem_raw);
    const int LOGM = p.logm;
    const int M = 1 << LOGM;
    const int C = p.lanes_per_cta;
    const int lane0 = blockIdx.x * C;
    const int nl = min(C, p.total_lanes - lane0);
    const int total = C << LOGM;

    // ---- fill: coalesced global read, bit-reversed placement in shared memory (as ntt_tile_kernel) ----
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
            if (p.in_twist)
                x = fr_mont_mul(x, fr_ldc(p.in_twist + (long long)(outer - orow * p.in_outer_div) * p.in_twist_sub_stride + off));
            fr_sts(sm + (c << LOGM) + bitrev(m, LOGM), x);
        }
    }
    __syncthreads();

    // ---- transform: decimation in time, bit-reversed in -> natural out, values rest in [0,4p) ----
    const int half_total = total >> 1;                         // butterflies per stage over all lanes of the CTA
#pragma unroll 1
    for (int s = 0; s < LOGM; s++) {
        const int h = 1 << s;
        for (int b = threadIdx.x; b < half_total; b += blockDim.x) {
            const int c = b >> (LOGM - 1), bl = b & ((M >> 1) - 1);
            const int low = bl & (h - 1);
            const int j0 = ((bl >> s) << (s + 1)) | low;
            fr_mem *base = sm + (c << LOGM);
            const fr_t u = fr_reduce_2p(fr_lds(base + j0));
            fr_t t = fr_lds(base + j0 + h);
            if (s == 0) t = fr_reduce_2p(t);                   // w = 1
            else t = fr_mont_mul(t, fr_ldc(p.tw + (size_t)(low << (LOGM - 1 - s)) * p.tws));
            fr_sts(base + j0, fr_add_raw(u, t));
            fr_sts(base + j0 + h, fr_sub_lazy4(u, t));
        }
        __syncthreads();
    }

    // ---- drain: optional twist / scale / canonicalisation, coalesced global write (as ntt_tile_kernel) ----
    const bool lane_fast_out = p.out_lane_stride < p.out_point_stride;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        int c, m;
        if (lane_fast_out) { c = idx % C; m = idx / C; } else { c = idx >> LOGM; m = idx & (M - 1); }
        if (c < nl) {
            const int L = lane0 + c;
            const int outer = L / p.lanes_inner, inner = L % p.lanes_inner;
            const int orow = p.out_outer_div ? outer / p.out_outer_div : outer;
            const int osub = p.out_outer_div ? outer - orow * p.out_outer_div + p.out_sub_base : 0;
            fr_t x = fr_lds(sm + (c << LOGM) + m);
            if (p.twist_full) {
                x = fr_mont_mul(x, fr_ldc(p.twist_full + (unsigned long long)inner * (unsigned)m));
            } else if (p.twist_lo) {
                const unsigned long long e = (unsigned long long)inner * (unsigned)m;
                fr_t w = fr_mont_mul(fr_ldc(p.twist_hi + (e >> p.twist_shift)), fr_ldc(p.twist_lo + (e & ((1ull << p.twist_shift) - 1))));
                x = fr_mont_mul(x, fr_reduce_p(w));
            }
            if (p.scale) x = fr_mont_mul(x, fr_ldc(p.scale));
            if (p.canon) x = fr_canon4(x);
            if (p.sink.nslabs)
                fr_stg(sink_at(p.sink, orow, (int)((long long)osub * p.out_sub_stride + (long long)inner * p.out_lane_stride + (long long)m * p.out_point_stride)), x);
            else
                fr_stg(p.out + (long long)orow * p.out_outer_stride + (long long)osub * p.out_sub_stride +
                           (long long)inner * p.out_lane_stride + (long long)m * p.out_point_stride, x);
        }
    }
}

// lanes per CTA: 512 points (256 threads) unless a lane is larger; at most 1024 threads
cudaError_t launch_ntt_lat(const NttTileParams &q, cudaStream_t st) {
    NttTileParams p = q;
    const int M = 1 << p.logm;
    static const int cta_points = getenv("LGR_LAT_CTA_POINTS") ? atoi(getenv("LGR_LAT_CTA_POINTS")) : 512;   // tuning knob
    int C = M >= cta_points ? 1 : cta_points / M;
    if (C > p.total_lanes) C = p.total_lanes;
    p.lanes_per_cta = C;
    const int threads = std::max(32, std::min(1024, (C * M) / 2));
    const size_t smem = (size_t)C * M * 32;
    if (smem > 64 * 1024) return cudaErrorInvalidValue;
    static bool attr_done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(ntt_lat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
    const int grid = (p.total_lanes + C - 1) / C;
    ntt_lat_kernel<<<grid, threads, smem, st>>>(p);
    return cudaGetLastError();
}

}  // namespace lgr

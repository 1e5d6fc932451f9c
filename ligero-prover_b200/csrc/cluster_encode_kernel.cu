// One-launch Reed-Solomon encode of a FEW rows at large k (k = 4096 or 8192) on thread-block clusters.
//
// The per-row schedule of the reference's stage contexts (include/zkp/nonbatch_context.hpp:445-468: write_buffer_clear ->
// encode_ntt_device -> sha256_digest_update, one row per callback) hands the device a single 8192-point encode at a time.
// The batched tile engine needs 5 launches for it and is sized for thousands of rows; here ONE launch does the whole encode
// with the row resident in distributed shared memory:
//   * a cluster of CS = k/1024 CTAs (8 for k = 8192) holds one k-point array, 1024 points (32 KiB) per CTA, 512 threads per
//     CTA = one radix-2 butterfly per thread per stage, rolled stage loop (a few KB of code);
//   * stages 0..9 touch only the CTA's own shared memory (__syncthreads); stages 10.. pair points of different CTAs through
//     distributed shared memory (cluster.map_shared_rank + cluster.sync);
//   * four clusters per row: cluster r = 1..3 computes c = k*iNTT_k(row) (redundantly -- 13 stages are cheaper than a trip
//     through global memory and a second launch), twists it by w_n^(r i)/k and transforms it with root w_n^4 into the coset
//     e[4m + r]; cluster 0 writes the coset that is a permuted copy of the message, e[4m] = row[c m mod k]
//     (api.cu: find_sys_mul), or computes it like the others when that shortcut is unknown.
// Same results as encode_ntt_device (src/webgpu/engine.cpp:755-770): canonical, bit-exact (tests/test_gpu_parity.py).
#include <cooperative_groups.h>

#include "kernels.h"
#include "ntt.cuh"

namespace cg = cooperative_groups;

namespace lgr {

constexpr int kClusterPts = 1024;            // points per CTA
constexpr int kClusterThreads = 512;

// element j of a k-point array spread over the cluster, 1024 consecutive points per CTA
template <typename Cluster>
__device__ __forceinline__ fr_mem *dist_at(Cluster &cluster, fr_mem *local_base, int j) {
    return cluster.map_shared_rank(local_base, (unsigned)(j >> 10)) + (j & (kClusterPts - 1));
}

// in-place decimation-in-time transform of the distributed array A: bit-reversed in, natural out, values rest in [0,4p)
template <typename Cluster>
__device__ __forceinline__ void dist_dit(Cluster &cluster, fr_mem *A, int logk, int rank, const fr_mem *__restrict__ tw) {
    const int b = rank * kClusterThreads + (int)threadIdx.x;          // this thread's butterfly in every stage
#pragma unroll 1
    for (int s = 0; s < logk; s++) {
        const int h = 1 << s;
        const int low = b & (h - 1);
        const int j0 = ((b >> s) << (s + 1)) | low;
        const bool local = s < 10;
        if (s == 10) cluster.sync();                                   // everybody has finished its local stages
        fr_mem *p0 = local ? A + (j0 & (kClusterPts - 1)) : dist_at(cluster, A, j0);
        fr_mem *p1 = local ? A + ((j0 + h) & (kClusterPts - 1)) : dist_at(cluster, A, j0 + h);
        const fr_t u = fr_reduce_2p(fr_lds(p0));
        fr_t t = fr_lds(p1);
        if (s == 0) t = fr_reduce_2p(t);
        else t = fr_mont_mul(t, fr_ldc(tw + (size_t)(low << (logk - 1 - s))));
        fr_sts(p0, fr_add_raw(u, t));
        fr_sts(p1, fr_sub_lazy4(u, t));
        if (local) __syncthreads(); else cluster.sync();
    }
}

template <int CS>
__global__ void __cluster_dims__(CS, 1, 1) __launch_bounds__(kClusterThreads, 1)
encode_rows_cluster_kernel(const fr_mem *__restrict__ rows, long long row_stride, const __grid_constant__ CodewordSink sink, int R,
                           const EncodeTables t) {
    constexpr int K = CS * kClusterPts;
    constexpr int LOGK = (CS == 8) ? 13 : ((CS == 4) ? 12 : 11);
    extern __shared__ __align__(32) unsigned char smem_raw[];
    fr_mem *C = reinterpret_cast<fr_mem *>(smem_raw);                  // this CTA's 1024 points of the coefficient array
    fr_mem *W = C + kClusterPts;                                       // ... and of the coset work array
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int cid = blockIdx.x / CS;                                   // cluster index = row * 4 + coset
    const long long row = cid >> 2;
    const int r = cid & 3;
    const fr_mem *src = rows + row * row_stride;
    const bool sys = t.sys_mul != 0;

    if (r == 0 && sys) {                                               // e[4m] = row[c*m mod k], reduced like the transforms would
        for (int q = threadIdx.x; q < kClusterPts; q += kClusterThreads) {
            const int m = rank * kClusterPts + q;
            const fr_t x = fr_ldg(src + (int)(((unsigned)m * (unsigned)t.sys_mul) & (unsigned)(K - 1)));
            fr_stg(sink_at(sink, row, 4 * m), fr_reduce_p(fr_reduce_2p(fr_reduce_2p(x))));
        }
        return;                                                        // (no cluster barrier is pending: the whole cluster takes this branch)
    }
    // c = k * iNTT_k(row): gather the row in bit-reversed order into this CTA's slice of C
    for (int q = threadIdx.x; q < kClusterPts; q += kClusterThreads) {
        const int j = rank * kClusterPts + q;
        fr_sts(C + q, fr_ldg(src + bitrev((uint32_t)j, LOGK)));
    }
    __syncthreads();
    dist_dit(cluster, C, LOGK, rank, t.inv_k);
    // coset r: W[j] = c[bitrev(j)] * w_n^(r * bitrev(j)) / k   (t.twist is indexed by the bit-reversed position, as in the fused encoder)
    const fr_mem *tw_r = t.twist + (size_t)r * K;
    for (int q = threadIdx.x; q < kClusterPts; q += kClusterThreads) {
        const int j = rank * kClusterPts + q;
        const fr_t c = fr_lds(dist_at(cluster, C, (int)bitrev((uint32_t)j, LOGK)));
        fr_sts(W + q, fr_mont_mul(c, fr_ldc(tw_r + j)));
    }
    __syncthreads();
    dist_dit(cluster, W, LOGK, rank, t.fwd_c);
    for (int q = threadIdx.x; q < kClusterPts; q += kClusterThreads) {
        const int m = rank * kClusterPts + q;
        fr_stg(sink_at(sink, row, 4 * m + r), fr_canon4(fr_lds(W + q)));
    }
    cluster.sync();                                                    // nobody leaves while a peer may still read its shared memory
}

template <int CS>
static cudaError_t launch_cluster(const fr_mem *rows, long long row_stride, const CodewordSink &sink, int R, const EncodeTables &t, cudaStream_t st) {
    const size_t smem = 2 * (size_t)kClusterPts * 32;
    static bool attr_done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(encode_rows_cluster_kernel<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
    encode_rows_cluster_kernel<CS><<<R * 4 * CS, kClusterThreads, smem, st>>>(rows, row_stride, sink, R, t);
    return cudaGetLastError();
}

bool encode_rows_cluster_ok(int logk) { return logk == 12 || logk == 13; }
// rows must NOT alias the sink (the clusters of a row read it while others already write the codeword)
cudaError_t launch_encode_rows_cluster(const fr_mem *rows, long long row_stride, const CodewordSink &sink, int R, int logk,
                                       const EncodeTables &t, cudaStream_t st) {
    if (R <= 0) return cudaSuccess;
    if (logk == 13) return launch_cluster<8>(rows, row_stride, sink, R, t, st);
    if (logk == 12) return launch_cluster<4>(rows, row_stride, sink, R, t, st);
    return cudaErrorInvalidValue;
}

}  // namespace lgr

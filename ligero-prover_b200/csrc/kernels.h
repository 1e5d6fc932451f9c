// Internal launch interface between the C-ABI layer (api.cu) and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lgr {

struct __align__(32) fr_mem { uint4 lo, hi; };   // one 32-byte element in global or shared memory

// ---- where codewords go ----------------------------------------------------------------------------
// nslabs == 0: plain row-major [row][n] at base[0], `row_stride` elements between rows.
// nslabs == G (exact multi-GPU layout, sharding.py): the n columns are cut into G slabs of 2^slab_shift columns;
// column j of row r goes to base[j >> slab_shift] + r * 2^slab_shift + (j & (2^slab_shift - 1)), i.e. slab h of every
// row lands row-major [rows][n/G] at base[h] -- which may be PEER memory: the encoder's 256-bit stores then travel
// over NVLink straight into the buffer the owner of slab h hashes from (no pack pass, no all-to-all kernel).
struct CodewordSink {
    fr_mem *base[8];
    int nslabs;
    int slab_shift;
    long long row_stride;
};
static inline CodewordSink plain_sink(fr_mem *out, long long row_stride) {
    CodewordSink s{}; s.base[0] = out; s.nslabs = 0; s.slab_shift = 0; s.row_stride = row_stride; return s;
}

// ---- generic tile NTT (ntt_kernels.cu) -------------------------------------------------------
// A "lane" is one M-point transform.  Lane L = outer * lanes_inner + inner lives at
//   in  + outer*in_outer_stride  + inner*in_lane_stride  + m*in_point_stride
// and is written to the same formula with the out_* strides (all in elements).
struct NttTileParams {
    const fr_mem *in;
    fr_mem *out;
    long long in_outer_stride, in_lane_stride, in_point_stride;
    long long out_outer_stride, out_lane_stride, out_point_stride;
    int lanes_inner;
    int total_lanes;
    int lanes_per_cta;
    int logm;
    const fr_mem *tw;        // tw[j*tws] = w_M^j * R, j < M/2 (inverse root for inverse transforms)
    int tws;
    // optional four-step twist: out(inner, m) *= w_N^(inner*m): one lookup in twist_full (small N) or
    // twist_hi[e >> shift] * twist_lo[e & mask] composed on the fly
    const fr_mem *twist_lo, *twist_hi, *twist_full;
    int twist_shift;
    // coset mode (large-k encoder): the outer index packs (row, s) as outer = row * div + s  (div = 0: plain mode).
    //   input : row = outer / in_outer_div is used for addressing; every loaded element is multiplied
    //           by in_twist[s * in_twist_sub_stride + element offset within the row]
    //   output: base = (outer / out_outer_div) * out_outer_stride + (s + out_sub_base) * out_sub_stride
    int in_outer_div, out_outer_div, out_sub_base;
    const fr_mem *in_twist;
    long long in_twist_sub_stride, out_sub_stride;
    CodewordSink sink;       // coset mode only: nslabs > 0 scatters the codeword columns into slabs (out is ignored)
    const fr_mem *scale;     // optional N^-1 * R (Montgomery form)
    int canon;               // 1: outputs reduced to [0,p)
    int in_natural;          // 1: input natural order (bit-reverse on load, DIT); 0 never used here
};
cudaError_t launch_ntt_tile(const NttTileParams &p, cudaStream_t st);
int ntt_tile_max_logm();

// ---- fused row encoder (encode_kernels.cu), k = 2^logk, 3 <= logk <= 11 ----------------------
struct EncodeTables {
    const fr_mem *inv_k;     // w_k^-j * R, j < k/2
    const fr_mem *fwd_c;     // (w_n^4)^j * R, j < k/2
    const fr_mem *twist;     // [4][k]: w_n^(r*bitrev_k(q)) / k * R
    int sys_mul;             // c with w_n^4 = w_k^c (odd, < k): e[4m] = row[c*m mod k]; 0 = unknown, compute coset 0 too
};
// rows_in: [R][in_row_stride] elements (first k of each row used); codewords go to `sink`
cudaError_t launch_encode_rows(const fr_mem *rows_in, long long in_row_stride, const CodewordSink &sink,
                               int R, int logk, const EncodeTables &t, cudaStream_t st);
// coset 0 of the large-k encoder: codeword[row][4m] = canonical(rows[row][c*m mod k])
cudaError_t launch_sys_copy(const fr_mem *rows, long long row_stride, const CodewordSink &sink, int R, int logk, uint32_t c, cudaStream_t st);
// one-launch encode of a few rows at k = 4096 / 8192 on thread-block clusters (cluster_encode_kernel.cu); rows must not alias the sink
bool encode_rows_cluster_ok(int logk);
cudaError_t launch_encode_rows_cluster(const fr_mem *rows, long long row_stride, const CodewordSink &sink, int R, int logk,
                                       const EncodeTables &t, cudaStream_t st);
int encode_rows_max_logk();
int encode_rows_min_logk();

// ---- element-wise (eltwise_kernels.cu) -------------------------------------------------------
enum EltOp {
    ELT_ADD = 0, ELT_SUB, ELT_MUL, ELT_DIV, ELT_FMA, ELT_FMA_CONST, ELT_ADD_ASSIGN, ELT_ADD_CONST, ELT_SUB_CONST,
    ELT_CONST_SUB, ELT_MUL_CONST, ELT_MONTMUL_CONST, ELT_BIT, ELT_POWMOD, ELT_POWADD, ELT_GATHER, ELT_QUAD_FUSED
};
struct EltParams {
    const fr_mem *x, *y, *z;
    fr_mem *out;
    size_t n;
    uint32_t scalar[8];      // canonical constant (op-dependent pre-conversion done by the caller)
    uint32_t scalar2[8];
    const uint32_t *idx;     // gather indices / powmod exponents
    uint32_t bit;
};
cudaError_t launch_eltwise(EltOp op, const EltParams &p, cudaStream_t st);

// tile combiners: acc[j] (+)= sum_t r[t] * tile[t][j]   (check_code over a resident tile)
cudaError_t launch_combine_code(const fr_mem *tile, long long row_stride, int T, int n, const fr_mem *r_raw /*[T] canonical scalars (device)*/,
                                fr_mem *acc, fr_mem *scratch, size_t scratch_elems, cudaStream_t st);
// acc[j] += sum_t a[t][j] * b[t][j]   (check_linear over two resident tiles)
cudaError_t launch_combine_linear(const fr_mem *a, const fr_mem *b, long long row_stride, int T, int n,
                                  fr_mem *acc, fr_mem *scratch, size_t scratch_elems, cudaStream_t st);
// acc[j] += sum_t r[t] * (x[t][j]*y[t][j] - z[t][j])   (check_quadratic over resident tiles); row_idx (device, optional):
// triple t sits at row row_idx[t] of x / y / z instead of row t
cudaError_t launch_combine_quad(const fr_mem *x, const fr_mem *y, const fr_mem *z, long long row_stride, int T, int n, const fr_mem *r_raw,
                                fr_mem *acc, fr_mem *scratch, size_t scratch_elems, cudaStream_t st, const uint32_t *row_idx = nullptr);
size_t combine_scratch_elems(int T, int n);
// out[t][s] = tile[t][idx[s]]: the sampled columns of T resident codewords (stage 3)
cudaError_t launch_gather_rows(const fr_mem *tile, long long row_stride, int T, const uint32_t *idx, int count, fr_mem *out, cudaStream_t st);   // partial sums only; callers add T (code) or 2T (quad) for scalars

// ---- SHA-256 column hashing + Merkle (sha_kernels.cu) ----------------------------------------
// ctx layout (u32 words): state[8][n] | pend[8][n] | rows_lo[n] | rows_hi[n]
static inline size_t sha_ctx_words(size_t n) { return 18 * n; }
cudaError_t launch_sha_init(uint32_t *ctx, int n, cudaStream_t st);
// absorb T rows: element j of row t at tile + t*row_stride + j
cudaError_t launch_sha_update(uint32_t *ctx, int n, const fr_mem *tile, long long row_stride, int T, cudaStream_t st);
cudaError_t launch_sha_final(const uint32_t *ctx, int n, uint32_t *digests, cudaStream_t st);
// nodes: (2*P2-1)*8 u32, P2 = bit_ceil(nleaves)
cudaError_t launch_merkle_build(const uint32_t *leaf_digests, int nleaves, uint32_t *nodes, cudaStream_t st);

// ---- cross-GPU hand-over flags (peer_kernels.cu) ------------------------------------------------
struct PeerSlots { unsigned long long *p[8]; int n; };      // one u64 slot per peer (peer memory, CUDA IPC)
cudaError_t launch_peer_signal(const PeerSlots &slots, unsigned long long value, cudaStream_t st);
cudaError_t launch_peer_wait(const unsigned long long *flags, int n, unsigned long long value, unsigned long long timeout_ns,
                             unsigned int *err, cudaStream_t st);

// ---- synthetic witness generator (sha_kernels.cu; bench / tests only) ------------------------
cudaError_t launch_synth(fr_mem *out, uint64_t seed, uint64_t row0, uint64_t nrows, uint64_t ncols, cudaStream_t st);

// ---- micro-benchmarks (ubench.cu, ubench_dpf.cu; linked into liblgr_ubench.so only) ------------------------------------------------------------
cudaError_t launch_ubench(int which, uint32_t *out, int iters, int blocks, int threads, cudaStream_t st);
cudaError_t launch_ubench_mont_occ(int nchain, int warps_per_sm, uint32_t *out, int iters, cudaStream_t st);
cudaError_t launch_ubench_chain(int variant, uint32_t *out, int iters, int warps_per_cta, int active_lanes, cudaStream_t st);
cudaError_t launch_ubench_dpf(int which, uint32_t *out, int iters, int blocks, int threads, cudaStream_t st);      // ubench_dpf.cu
cudaError_t launch_ubench_mont_sha(uint32_t *out, int iters_mont, int iters_sha, int blocks, cudaStream_t st);
cudaError_t launch_dpf_mul(const uint32_t *a, const uint32_t *b, uint32_t *out, int n, cudaStream_t st);

}  // namespace lgr

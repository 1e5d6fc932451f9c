/*
 * lgr.h -- C ABI of the B200-native Ligero hot path (liblgr.so).
 *
 * Drop-in boundary: the reference has no C ABI; its hot path sits behind the duck-typed
 * `Executor` template parameter of the stage contexts (include/zkp/nonbatch_context.hpp:392,586,875)
 * bound to `webgpu_context` (include/wgpu.hpp:50-295, src/webgpu_prover.cpp:53-55).  Every entry
 * point below names the webgpu_context / device_context method it replaces.  The header-only
 * C++ adapter ligero-prover_b200/host/cuda_executor.hpp re-exposes this ABI with the reference's
 * method names so that nonbatch_context.hpp instantiates on it.
 *
 * Conventions
 *   - Elements: 32 bytes, 8 x u32 little-endian limbs, canonical [0,p), NOT Montgomery form
 *     (include/ligetron/webgpu/device_bignum.hpp:30-100).  The field is BN254 Fr (hard-wired in the
 *     reference's WGSL too: shader/bn254fr.wgsl.in:19-45).
 *   - `void *` buffer arguments are DEVICE pointers (from lgr_alloc or any CUDA allocation, e.g. a
 *     torch tensor's data_ptr()), 32-byte aligned.  Host pointers are named host_*.
 *   - All work is enqueued on the context's stream and returns immediately (WebGPU queue semantics,
 *     src/webgpu/device_context.cpp:344-354); lgr_read and lgr_sync block.  On the context's OWN stream a call may be
 *     held back until the next call on the context (lgr_encode, below) -- order on the stream is always call order, and
 *     lgr_sync / lgr_read / lgr_set_stream enqueue everything held back first.  A caller that shares the stream with
 *     other CUDA code uses lgr_set_stream, where nothing is held back.
 *   - Return value: 0 = LGR_OK, otherwise an error code; lgr_last_error() gives the text for the
 *     calling thread.  (The reference aborts on device errors, device_context.cpp:121-128; a C ABI
 *     reports instead.)  Not thread-safe per context, like the reference.
 *   - There is no CPU fallback: every compute entry point launches sm_100a kernels or fails.
 */
#ifndef LGR_H
#define LGR_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LGR_OK 0
#define LGR_ERR_INVALID 1
#define LGR_ERR_CUDA 2
#define LGR_ERR_NOMEM 3
#define LGR_ERR_UNSUPPORTED 4

typedef struct lgr_ctx lgr_ctx;

/* size selectors / directions for lgr_ntt (ntt_{forward,inverse}_{k,2k,n}, include/wgpu.hpp:128-134) */
#define LGR_SIZE_K 0
#define LGR_SIZE_2K 1
#define LGR_SIZE_N 2
#define LGR_FORWARD 0
#define LGR_INVERSE 1

const char *lgr_last_error(void);
int lgr_version(void);

/* ---- lifecycle -------------------------------------------------------------------------------
 * lgr_create = webgpu_init + ntt_init (include/wgpu.hpp:73-82, src/webgpu/engine.cpp:176-211):
 * l = message size, k = padded size (power of two >= 8), n = 4k, p must equal the BN254 scalar
 * modulus, root_k / root_2k / root_n as returned by bn254_gmp::generate_omegas
 * (src/bn254.cpp:51-64).  Builds the twiddle tables (engine.cpp:1382-1503) on `device`. */
int lgr_create(lgr_ctx **out, int device, uint32_t l, uint32_t k, uint32_t n, const uint32_t p[8],
               const uint32_t root_k[8], const uint32_t root_2k[8], const uint32_t root_n[8]);
int lgr_destroy(lgr_ctx *ctx);
/* run on a caller-owned cudaStream_t (e.g. torch's current stream); NULL = the context's own */
int lgr_set_stream(lgr_ctx *ctx, void *cuda_stream);
int lgr_sync(lgr_ctx *ctx);                                   /* device_synchronize, device_context.hpp:35 */
int lgr_geometry(const lgr_ctx *ctx, uint32_t *l, uint32_t *k, uint32_t *n);   /* message/padding/encoding_size, wgpu.hpp:153-155 */
/* number of kernels this context has launched so far (bench.py "gpu_launches") */
int lgr_launch_count(const lgr_ctx *ctx, uint64_t *count);

/* ---- buffers (include/ligetron/webgpu/device_context.hpp:44-66) ------------------------------ */
int lgr_alloc(lgr_ctx *ctx, size_t bytes, void **dptr);                          /* make_device_buffer (zero-filled) */
int lgr_free(lgr_ctx *ctx, void *dptr);
int lgr_write(lgr_ctx *ctx, void *dst, size_t dst_off, const void *host_src, size_t bytes);   /* write_buffer: host data is copied before return */
int lgr_write_clear(lgr_ctx *ctx, void *dst, size_t dst_bytes, const void *host_src, size_t bytes); /* write_buffer_clear: prefix written, rest zeroed */
int lgr_clear(lgr_ctx *ctx, void *dst, size_t off, size_t bytes);                /* clear_buffer */
int lgr_copy(lgr_ctx *ctx, const void *src, void *dst, size_t bytes);            /* copy_buffer_to_buffer */
int lgr_copy_clear(lgr_ctx *ctx, const void *src, size_t src_bytes, void *dst, size_t dst_bytes); /* copy_buffer_clear */
int lgr_read(lgr_ctx *ctx, void *host_dst, const void *src, size_t src_off, size_t bytes);    /* copy_to_host (blocking) */

/* ---- transforms ------------------------------------------------------------------------------ */
/* encode_ntt_device (engine.cpp:755-770): buf = n elements, buf[0:k) message, buf[k:n) zero on
 * entry; on return buf = codeword.  In place.  At k > 2048 on the context's own stream, consecutive calls on distinct
 * buffers (nonbatch_context.hpp:667-668,715-720 issues two or six) are enqueued together, as one CUDA graph in which
 * the rows run side by side, by the next call of any other kind. */
int lgr_encode(lgr_ctx *ctx, void *buf);
/* decode_ntt_device (engine.cpp:772-796): iNTT_n, fold, NTT_k on buf[0:k); buf[k:n) = coefficients */
int lgr_decode(lgr_ctx *ctx, void *buf);
/* ntt_forward_k/2k/n, ntt_inverse_k/2k/n (engine.cpp:798-930): natural order in and out, canonical,
 * in place on the first k / 2k / n elements of buf */
int lgr_ntt(lgr_ctx *ctx, void *buf, int size_sel, int dir);
/* any power-of-two transform with an explicit root (BASELINE configs 1, 2): `batch` independent
 * transforms of 2^logn points laid out back to back; omega must have order 2^logn */
int lgr_ntt_pow2(lgr_ctx *ctx, void *buf, uint32_t logn, uint32_t batch, const uint32_t omega[8], int dir);

/* ---- column hashing + Merkle (shader/sha256.wgsl:127-230, engine.cpp:1514-1686,
 *      include/zkp/merkle_tree.hpp:343-375) ---------------------------------------------------- */
size_t lgr_sha_ctx_bytes(uint32_t ninst);                     /* <= ninst * sizeof(webgpu_context::sha256_context) */
int lgr_sha_init(lgr_ctx *ctx, void *sha_ctx, uint32_t ninst);                   /* sha256_digest_init */
int lgr_sha_update(lgr_ctx *ctx, void *sha_ctx, uint32_t ninst, const void *buf); /* sha256_digest_update: element j -> column j */
/* same as nrows successive lgr_sha_update calls on rows tile + r*row_stride_elems, one launch */
int lgr_sha_update_rows(lgr_ctx *ctx, void *sha_ctx, uint32_t ninst, const void *tile, uint64_t row_stride_elems, uint32_t nrows);
int lgr_sha_final(lgr_ctx *ctx, const void *sha_ctx, uint32_t ninst, void *digests); /* sha256_digest_final: ninst*32 B, state words native u32 */
size_t lgr_merkle_node_count(uint32_t nleaves);               /* 2*bit_ceil(nleaves)-1 */
/* merkle_tree::initialize_from_digest + build_tree: nodes = node_count*32 B in the reference's heap
 * layout and byte order; root = nodes[0..32) */
int lgr_merkle_build(lgr_ctx *ctx, const void *leaf_digests, uint32_t nleaves, void *nodes);

/* ---- element-wise (shader/kernels.wgsl.in:325-549, engine.cpp:432-751) ------------------------
 * n = element count; pointers already include any element offset (webgpu::eltwise_offset). */
int lgr_elt_add(lgr_ctx *ctx, const void *x, const void *y, void *out, size_t n);         /* EltwiseAddMod */
int lgr_elt_sub(lgr_ctx *ctx, const void *x, const void *y, void *out, size_t n);         /* EltwiseSubMod */
int lgr_elt_mul(lgr_ctx *ctx, const void *x, const void *y, void *out, size_t n);         /* EltwiseMultMod */
int lgr_elt_div(lgr_ctx *ctx, const void *x, const void *y, void *out, size_t n);         /* EltwiseDivMod */
int lgr_elt_fma(lgr_ctx *ctx, const void *x, const void *y, void *out, size_t n);         /* EltwiseFMAMod: out += x*y */
int lgr_elt_fma_const(lgr_ctx *ctx, const void *x, void *out, size_t n, const uint32_t c[8]);   /* EltwiseFMAMod(bind, k): out += c*x */
int lgr_elt_add_assign(lgr_ctx *ctx, const void *x, void *out, size_t n);                 /* EltwiseAddAssignMod: out += x */
int lgr_elt_add_const(lgr_ctx *ctx, const void *x, void *out, size_t n, const uint32_t c[8]);   /* EltwiseAddMod(bind, k) */
int lgr_elt_sub_const(lgr_ctx *ctx, const void *x, void *out, size_t n, const uint32_t c[8]);   /* EltwiseSubConstMod: x - c */
int lgr_elt_const_sub(lgr_ctx *ctx, const void *x, void *out, size_t n, const uint32_t c[8]);   /* EltwiseConstSubMod: c - x */
int lgr_elt_mul_const(lgr_ctx *ctx, const void *x, void *out, size_t n, const uint32_t c[8]);   /* EltwiseMultMod(bind, k) */
int lgr_elt_montmul_const(lgr_ctx *ctx, const void *x, void *out, size_t n, const uint32_t c[8]); /* EltwiseMontMultMod: x*c*2^-256 */
int lgr_elt_bit(lgr_ctx *ctx, const void *x, void *out, size_t n, uint32_t bit);          /* EltwiseBitDecompose */
/* EltwisePowMod / EltwisePowAddMod (powmod_context.cpp:141-197): out (+)= coeff * base^exp, exp = n x u32 on device */
int lgr_elt_powmod(lgr_ctx *ctx, const void *coeff, const void *exp, void *out, size_t n, const uint32_t base[8], int add);
/* fused check_quadratic (nonbatch_context.hpp:771-780): out += r*(x*y - z) in one sweep */
int lgr_elt_quad(lgr_ctx *ctx, const void *x, const void *y, const void *z, void *out, size_t n, const uint32_t r[8]);

/* ---- sampling (engine.cpp:1689-1809) ---------------------------------------------------------- */
int lgr_sample_init(lgr_ctx *ctx, const uint64_t *host_indices, uint32_t count);          /* sampling_init */
int lgr_sample_gather(lgr_ctx *ctx, const void *x, void *out);                            /* sample_gather: out[i] = x[idx[i]] */

/* sample_gather for nrows resident codewords: out[t][s] = tile[t*row_stride_elems + idx[s]], i.e. the proof's
 * host_samplings layout [row][sample][8 x u32] (nonbatch_context.hpp:906-950) */
int lgr_sample_gather_rows(lgr_ctx *ctx, const void *tile, uint64_t row_stride_elems, uint32_t nrows, void *out);

/* ---- batched fast paths (B200-native additions; same results as the per-row calls) ------------ */
/* nrows encodes in one launch: rows[r] = k elements at rows + r*row_stride_elems, codewords[r] = n
 * elements at codewords + r*n.  rows may alias codewords when row_stride_elems == n (in place). */
int lgr_encode_rows(lgr_ctx *ctx, const void *rows, uint64_t row_stride_elems, uint32_t nrows, void *codewords);
/* ---- exact multi-GPU commitment (B200-native; no reference counterpart: the reference is single-device) ----
 * A column's leaf is ONE SHA-256 stream over all rows (nonbatch_context.hpp:445-451 -> sha256.wgsl:147-177), so rows
 * shard across GPUs for the ENCODE only; the codeword columns are then cut into `nslabs` slabs of n/nslabs columns and
 * slab h of every row is hashed by GPU h.  lgr_encode_rows_slabs writes column j of row r to
 * slab_base[j / (n/nslabs)] + r*(n/nslabs) + j % (n/nslabs): row-major [nrows][n/nslabs] per slab.  slab_base[h] may
 * be PEER memory (lgr_ipc_open): the encoder's stores then cross NVLink directly into the hasher's buffer. */
int lgr_encode_rows_slabs(lgr_ctx *ctx, const void *rows, uint64_t row_stride_elems, uint32_t nrows, void *const *slab_base, uint32_t nslabs);
/* peer-visible device memory through CUDA IPC (one process per GPU): alloc + export on the owner, open on the peers */
int lgr_ipc_alloc(lgr_ctx *ctx, size_t bytes, void **dptr, unsigned char handle[64]);
int lgr_ipc_open(lgr_ctx *ctx, const unsigned char handle[64], void **dptr);
int lgr_ipc_close(lgr_ctx *ctx, void *dptr);
int lgr_ipc_free(lgr_ctx *ctx, void *dptr);
/* hand-over flags (u64 monotone counters in peer-visible memory), enqueued on the context's stream:
 * signal: *slots[i] = value for every i, after everything enqueued before it (system-scope release);
 * wait  : until flags[i] >= value for all i < nflags (system-scope acquire); gives up after timeout_ms and writes
 *         a non-zero code to *err_flag (u32, device memory) instead of hanging the GPU. */
int lgr_peer_signal(lgr_ctx *ctx, void *const *slots, uint32_t nslots, uint64_t value);
int lgr_peer_wait(lgr_ctx *ctx, const void *flags, uint32_t nflags, uint64_t value, uint32_t timeout_ms, void *err_flag);

/* stage-1 commit of an R x k device-resident row-major witness: for every row, encode and absorb
 * into the n column hashes in row order (nonbatch_context.hpp:445-451), then final + tree
 * (nonbatch_context.hpp:555-558, merkle_tree.hpp:343-375).  digests: n*32 B; nodes: (2n-1)*32 B or NULL.
 * Encoding of tile t+1 overlaps hashing of tile t on a second stream. */
int lgr_encode_commit(lgr_ctx *ctx, const void *rows, uint64_t nrows, void *digests, void *nodes);
/* the same tile pipeline without init / final / tree: every row is encoded and absorbed, in order, into the caller's
 * column-hash context (n instances, lgr_sha_init): what a stage-1 context does between its first row and
 * flush_digests when other rows (the 2k-domain mask rows, nonbatch_context.hpp:482-494) have to follow */
int lgr_encode_absorb(lgr_ctx *ctx, void *sha_ctx, const void *rows, uint64_t nrows);
/* same pipeline for a HOST-resident witness (pinned memory recommended): rows are copied tile by
 * tile on a third stream (H2D of tile t+2, encode of tile t+1 and hashing of tile t overlap), the
 * n digests (may be NULL) and the 32-byte root come back to the host; blocking.  This is the call
 * behind bench.py's "e2e" figure. */
int lgr_encode_commit_host(lgr_ctx *ctx, const void *host_rows, uint64_t nrows, void *host_digests, void *host_root);
/* per-kernel device timing of the commit pipeline: CUDA events on the launching streams around
 * every encode / hash-update launch.  lgr_profile_read drains the totals (ms) and launch counts. */
int lgr_profile(lgr_ctx *ctx, int enable);
int lgr_profile_read(lgr_ctx *ctx, double *encode_ms, uint64_t *encode_launches, double *sha_ms, uint64_t *sha_launches);
/* check_code over a resident tile of nrows codewords: acc[j] += sum_t r[t]*tile[t][j]
 * (nonbatch_context.hpp:756-763); host_r = nrows x 8 u32 canonical scalars */
int lgr_combine_code(lgr_ctx *ctx, const void *tile, uint32_t nrows, const uint32_t *host_r, void *acc);
/* check_quadratic over three resident tiles: acc[j] += sum_t r[t]*(x[t][j]*y[t][j] - z[t][j])
 * (nonbatch_context.hpp:771-780: EltwiseMultMod, EltwiseSubMod, EltwiseFMAMod(r) per triple) */
int lgr_combine_quad(lgr_ctx *ctx, const void *tile_x, const void *tile_y, const void *tile_z, uint32_t nrows, const uint32_t *host_r, void *acc);
/* same with the three operand rows of triple t at x/y/z + t*row_stride_elems: the layout of a tile encoded in
 * emission order (x_0, y_0, z_0, x_1, ...: x = tile, y = tile + n, z = tile + 2n, stride 3n) */
int lgr_combine_quad_rows(lgr_ctx *ctx, const void *x, const void *y, const void *z, uint64_t row_stride_elems, uint32_t nrows, const uint32_t *host_r, void *acc);
/* same for triples scattered over a tile encoded in emission order: triple t occupies codeword rows host_x_rows[t],
 * +1 and +2 of `tile` (row stride n); one launch for any interleaving of linear rows and triples */
int lgr_combine_quad_indexed(lgr_ctx *ctx, const void *tile, const uint32_t *host_x_rows, uint32_t ntriples, const uint32_t *host_r, void *acc);
/* on_batch_bit rows (nonbatch_context.hpp:798-808 copies x into y and z before check_quadratic): acc += sum_t r_t *
 * (x_t*x_t - x_t) for the codeword rows host_rows[t] of `tile` */
int lgr_combine_bit_indexed(lgr_ctx *ctx, const void *tile, const uint32_t *host_rows, uint32_t count, const uint32_t *host_r, void *acc);
/* check_linear over two resident tiles: acc[j] += sum_t a[t][j]*b[t][j] (nonbatch_context.hpp:765-769) */
int lgr_combine_linear(lgr_ctx *ctx, const void *tile_a, const void *tile_b, uint32_t nrows, void *acc);

/* ---- synthetic data + micro-benchmarks (bench / tests) ---------------------------------------- */
/* uniform canonical elements keyed by (seed,row,col): 256 bits >> 2, one conditional subtract
 * (include/zkp/finite_field_gmp.hpp:70-78); identical to the oracle's generator */
int lgr_synth(lgr_ctx *ctx, void *out, uint64_t seed, uint64_t row0, uint64_t nrows, uint64_t ncols);
/* (micro-benchmarks live in include/lgr_ubench.h / liblgr_ubench.so, outside the drop-in library) */

#ifdef __cplusplus
}
#endif
#endif

/*
 * oracle/lgo.h -- CPU ORACLE for the Ligero encode + commit hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the algorithm the reference (ligeroinc/ligero-prover v1.5.0)
 * implements in WGSL + C++ for the path SURVEY.md section 8 scopes.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product (ligero-prover_b200/csrc) never links, includes or calls anything in oracle/.
 *
 * Parity status: PINNED to the reference's own shaders.  The reference ships no golden vectors / KATs for NTT, SHA leaf
 * format, Merkle root or combiners and its C++ host side cannot be built here (Dawn, wabt, GMP
 * headers, Boost absent), but its arithmetic is first-party WGSL text: oracle/wgsl2cpp.py transliterates
 * shader/{bigint,bn254fr,kernels}.wgsl.in and shader/sha256.wgsl to C++, oracle/wgslref.cpp drives them in
 * src/webgpu/engine.cpp's dispatch order (oracle/_ref/libwgslref.so), and tests/test_wgslref_cpu.py checks every function
 * below against them -- live where the library is present, and against committed outputs of those shaders
 * (tests/golden/wgslref_vectors.json) everywhere.  Also kept: (1) the standard definitions (DFT over BN254 Fr with the
 * roots of src/bn254.cpp:36-43, FIPS 180-4 SHA-256), (2) an independent pure-Python restatement (oracle/pyref.py),
 * (3) the derived seed vectors of SURVEY.md section 8c (tests/golden/survey_vectors.json), (4) the reference's own device
 * KATs (tests/webgpu/test_powmod.cpp).  See DESIGN.md section 2.
 *
 * Element format everywhere: 8 x u32 little-endian limbs = 4 x u64 little-endian limbs = 32 bytes,
 * canonical value in [0,p), NOT Montgomery form
 * (include/ligetron/webgpu/device_bignum.hpp:30-100, shader/bigint.wgsl.in:36).
 */
#ifndef LGO_H
#define LGO_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t v[4]; } lgo_fr;            /* canonical, little-endian limbs */

/* ---- field (src/bn254.cpp:21-49, shader/bn254fr.wgsl.in:19-124) ---- */
void lgo_fr_modulus(lgo_fr *out);
void lgo_fr_add(lgo_fr *o, const lgo_fr *a, const lgo_fr *b);
void lgo_fr_sub(lgo_fr *o, const lgo_fr *a, const lgo_fr *b);
void lgo_fr_mul(lgo_fr *o, const lgo_fr *a, const lgo_fr *b);
void lgo_fr_pow(lgo_fr *o, const lgo_fr *a, uint64_t e);
void lgo_fr_inv(lgo_fr *o, const lgo_fr *a);
/* shader/bn254fr.wgsl.in:101-104 montgomery_mul: a*b*2^-256 mod p, canonical */
void lgo_fr_montmul(lgo_fr *o, const lgo_fr *a, const lgo_fr *b);

/* roots of unity: src/bn254.cpp:36-43 (root1, root2) and :51-64 (generate_omegas) */
void lgo_root1(lgo_fr *o);
void lgo_root2(lgo_fr *o);
void lgo_omegas(uint64_t k, lgo_fr *w_k, lgo_fr *w_2k, lgo_fr *w_n);

/* ---- transforms (shader/kernels.wgsl.in:57-323 + src/webgpu/engine.cpp:844-882,932-968) ----
 * natural order in / natural order out, canonical outputs. */
void lgo_ntt(lgo_fr *x, size_t N, const lgo_fr *omega, int inverse);       /* O(N log N), in place */
void lgo_dft_naive(lgo_fr *out, const lgo_fr *x, size_t N, const lgo_fr *omega, int inverse); /* O(N^2) */

/* engine.cpp:755-770 encode_ntt_device: buf has n = 4k elements, buf[0:k) = message row,
 * buf[k:n) must be zero on entry (write_buffer_clear); result = codeword in buf[0:n). */
void lgo_encode(lgo_fr *buf, size_t k);
/* mask-row path, nonbatch_context.hpp:482-494: iNTT_2k (w_2k) then NTT_n; buf[0:2k) input */
void lgo_encode_2k(lgo_fr *buf, size_t k);
/* engine.cpp:772-796 decode_ntt_device: iNTT_n, fold (kernels.wgsl.in:104-116), NTT_k on buf[0:k) */
void lgo_decode(lgo_fr *buf, size_t k);

/* ---- element-wise combiners (shader/kernels.wgsl.in:325-510) ---- */
void lgo_elt_add(lgo_fr *out, const lgo_fr *x, const lgo_fr *y, size_t n);
void lgo_elt_sub(lgo_fr *out, const lgo_fr *x, const lgo_fr *y, size_t n);
void lgo_elt_mul(lgo_fr *out, const lgo_fr *x, const lgo_fr *y, size_t n);
void lgo_elt_div(lgo_fr *out, const lgo_fr *x, const lgo_fr *y, size_t n);
void lgo_elt_fma(lgo_fr *out, const lgo_fr *x, const lgo_fr *y, size_t n);            /* out += x*y */
void lgo_elt_fma_const(lgo_fr *out, const lgo_fr *x, const lgo_fr *c, size_t n);      /* out += c*x */
void lgo_elt_add_assign(lgo_fr *out, const lgo_fr *x, size_t n);                      /* out += x   */
void lgo_elt_add_const(lgo_fr *out, const lgo_fr *x, const lgo_fr *c, size_t n);
void lgo_elt_sub_const(lgo_fr *out, const lgo_fr *x, const lgo_fr *c, size_t n);      /* x - c */
void lgo_elt_const_sub(lgo_fr *out, const lgo_fr *x, const lgo_fr *c, size_t n);      /* c - x */
void lgo_elt_mul_const(lgo_fr *out, const lgo_fr *x, const lgo_fr *c, size_t n);
void lgo_elt_montmul_const(lgo_fr *out, const lgo_fr *x, const lgo_fr *c, size_t n);  /* x*c*2^-256 */
void lgo_elt_bit(lgo_fr *out, const lgo_fr *x, uint32_t bit, size_t n);
/* kernels.wgsl.in:512-537 + powmod_context.cpp: out = coeff * base^exp (exp: 32 bit) */
void lgo_elt_powmod(lgo_fr *out, const lgo_fr *coeff, const uint32_t *exp, const lgo_fr *base, size_t n, int add);
/* kernels.wgsl.in:541-549 */
void lgo_gather(lgo_fr *out, const lgo_fr *x, const uint32_t *idx, size_t ns);

/* ---- column hashing (shader/sha256.wgsl:127-230) ----
 * Streaming context for ninst independent columns.  lgo_sha_update appends element j of `row`
 * to column j: 8 LE u32 limbs, least significant limb first, each limb most significant byte
 * first (sha256.wgsl:155-162).  lgo_sha_final writes digests as the 8 state words in native
 * little-endian u32 (sha256.wgsl:226-228). */
typedef struct lgo_sha lgo_sha;
lgo_sha *lgo_sha_new(size_t ninst);
void lgo_sha_free(lgo_sha *s);
void lgo_sha_init(lgo_sha *s);
void lgo_sha_update(lgo_sha *s, const lgo_fr *row);
void lgo_sha_final(lgo_sha *s, uint8_t *digests /* ninst*32 */);
/* plain FIPS 180-4 SHA-256 of a byte string (for the tree and for pinning against hashlib) */
void lgo_sha256(uint8_t out[32], const uint8_t *msg, size_t len);

/* ---- Merkle tree (include/zkp/merkle_tree.hpp:343-375) ----
 * nodes: (2*P-1)*32 bytes, P = bit_ceil(nleaves); heap layout, leaves at [P-1, P-1+nleaves),
 * missing leaves are all-zero digests (hash.hpp:157), node i = SHA-256(node[2i+1] || node[2i+2])
 * in standard big-endian digest byte order; root = node 0. */
size_t lgo_merkle_nodes(size_t nleaves);
void lgo_merkle_build(uint8_t *nodes, const uint8_t *leaf_digests, size_t nleaves);

/* ---- synthetic witness generator (finite_field_gmp.hpp:70-78) ----
 * element (row, col) of matrix `seed`: 256 bits from SplitMix64 keyed by (seed,row,col),
 * >> 2, one conditional subtract of p. */
void lgo_synth(lgo_fr *out, uint64_t seed, uint64_t row0, uint64_t nrows, uint64_t ncols);

/* ---- whole path: encode + commit of an R x k row-major witness (threads = OpenMP) ----
 * digests: n*32 bytes; nodes: lgo_merkle_nodes(n)*32 bytes (may be NULL).  Rows enter every
 * column hash in order 0..R-1 (nonbatch_context.hpp:445-451).  Returns threads used. */
int lgo_encode_commit(const lgo_fr *rows, size_t R, size_t k, uint8_t *digests, uint8_t *nodes);
/* same, but rows generated on the fly by lgo_synth(seed) (no R*k*32 B buffer needed) */
int lgo_encode_commit_synth(uint64_t seed, size_t R, size_t k, uint8_t *digests, uint8_t *nodes);
/* batch of independent NTTs (CPU baseline for config 2): x[batch][N] */
int lgo_ntt_batch(lgo_fr *x, size_t N, size_t batch, const lgo_fr *omega, int inverse);
int lgo_num_threads(void);
void lgo_set_threads(int n);   /* torchrun exports OMP_NUM_THREADS=1: the baseline legs set all cores explicitly */

#ifdef __cplusplus
}
#endif
#endif

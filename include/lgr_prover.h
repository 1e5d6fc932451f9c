/*
 * lgr_prover.h -- C ABI of the host-side prover driver (liblgr_prover.so), the layer ABOVE include/lgr.h.
 *
 * The reference's prover entry point is `main` of src/webgpu_prover.cpp:59-495: interpreter -> rows -> three
 * stages -> proof_data.gz.  The interpreter / witness manager are out of scope; this library
 * is the rest of that function for a caller that already has the rows: stage 1/2/3 over the device hot path
 * (nonbatch_context.hpp:445-993), the Fiat-Shamir transcript and sampler (src/webgpu_prover.cpp:281-353,
 * include/zkp/random.hpp:87-146, include/util/portable_sample.hpp), Merkle openings
 * (include/zkp/merkle_tree.hpp:155-318) and the proof container (include/zkp/proof_serializer.hpp:60-224,
 * proto/ligero_proof.proto).  Host pieces are callable on their own (no GPU needed) so that they can be
 * checked against an independent restatement.
 *
 * All functions return 0 on success; lgrp_last_error() gives the text otherwise.
 */
#ifndef LGR_PROVER_H
#define LGR_PROVER_H
#include <stddef.h>
#include <stdint.h>

#include "lgr.h"

#ifdef __cplusplus
extern "C" {
#endif

const char *lgrp_last_error(void);

/* ---- transcript / randomness (host only) ------------------------------------------------------- */
/* zkp::hash<sha256>("LigetronStage1", root, instance_hash)  (src/webgpu_prover.cpp:281-282) */
int lgrp_stage1_seed(const uint8_t root[32], const uint8_t instance_hash[32], uint8_t out[32]);
/* zkp::hash<sha256>("LigetronStage2", root, code, linear, quad), each vector = nwords u32 (src/webgpu_prover.cpp:337-341) */
int lgrp_stage2_seed(const uint8_t root[32], const uint32_t *code, const uint32_t *linear, const uint32_t *quad, size_t nwords, uint8_t out[32]);
/* the first `count` bytes of zkp::hash_random_engine<sha256>(seed)  (include/zkp/random.hpp:87-146) */
int lgrp_hash_random_bytes(const uint8_t seed[32], uint8_t *out, size_t count);
/* portable_sample(iota(n), sample_size, hash_random_engine(seed)) + sort (src/webgpu_prover.cpp:343-351);
 * out holds min(n, sample_size) indices */
int lgrp_sample_indices(const uint8_t seed[32], uint64_t n, uint64_t sample_size, uint64_t *out, uint64_t *out_count);
/* `count` draws of bn254_gmp::generate_random over mpz_random_engine(key, iv): 8 x u32 limbs each
 * (include/util/csprng.hpp:28-110, include/zkp/finite_field_gmp.hpp:70-78) */
int lgrp_fr_random(const uint8_t key[32], const uint8_t iv[16], size_t count, uint32_t *out_limbs);

/* ---- Merkle openings (host only; nodes = the array lgr_merkle_build produced) ------------------ */
/* merkle_tree::decommit + canonical sibling order (merkle_tree.hpp:155-215, proof_serializer.hpp:82-117).
 * positions_out / siblings_out need room for at most total_count entries; *count_out = number written */
int lgrp_decommit(const uint8_t *nodes, uint64_t total_count, const uint64_t *leaf_idx, uint64_t nidx, uint64_t *positions_out,
                  uint8_t *siblings_out, uint64_t *count_out);
/* merkle_tree::recommit (merkle_tree.hpp:232-318): leaves = nidx digests of the opened leaves */
int lgrp_recommit(const uint8_t *leaves, const uint64_t *leaf_idx, uint64_t nidx, uint64_t total_count, const uint8_t *siblings, uint64_t nsib,
                  uint8_t root_out[32]);

/* ---- proof container (host only) ---------------------------------------------------------------- */
typedef struct lgrp_proof lgrp_proof;
/* which: 0 = serialized LigeroProofEnvelope, 1 = gzip of it (what the reference writes to proof_data.gz) */
int lgrp_proof_bytes(const lgrp_proof *p, int which, const uint8_t **data, size_t *len);
/* parse an envelope (gzip or plain) back into a proof object (deserialize_proof, proof_serializer.hpp:193-224);
 * lgrp_proof_bytes(.., 0, ..) then returns the re-serialized envelope */
int lgrp_proof_parse(const uint8_t *data, size_t len, lgrp_proof **out);
void lgrp_proof_free(lgrp_proof *p);
/* prover self-check flags (src/webgpu_prover.cpp:465-471): bit 0 code, bit 1 linear, bit 2 quadratic; the two stage seeds */
int lgrp_proof_info(const lgrp_proof *p, uint32_t *valid_bits, uint8_t stage1_seed[32], uint8_t stage2_seed[32], uint64_t *encoded_rows);

/* host wall-clock milliseconds of stage 1 (commit), stage 2 (test vectors, sampling, self-check), stage 3 (openings) and
 * the container (serialise + gzip); every stage ends with a blocking read, so device time is included */
int lgrp_proof_timing(const lgrp_proof *p, double ms[4]);

/* ---- packing committed witnesses into rows (host only) --------------------------------------------
 * witness_manager::commit_release_witness / process_reset_*_row / finalize
 * (include/zkp/backend/witness_manager.hpp:117-186,188-269,497-503): rows are emitted lazily, when a witness arrives
 * and the open row already holds l of them; finalize emits the partial linear row, then the partial triple.  The
 * result is exactly the (kinds, values, coefs) triple of lgrp_statement. */
typedef struct lgrp_packer lgrp_packer;
int lgrp_packer_create(uint32_t l, lgrp_packer **out);
void lgrp_packer_free(lgrp_packer *p);
int lgrp_packer_push_linear(lgrp_packer *p, const uint32_t value[8], const uint32_t coef[8]);          /* commit_status::linear_ready */
int lgrp_packer_push_quadratic(lgrp_packer *p, const uint32_t xyz[24], const uint32_t coef_xyz[24]);   /* commit_status::quadratic_ready */
int lgrp_packer_finalize(lgrp_packer *p);
/* views into the packer (valid until the next push / free): n_rows encoded rows of l elements each */
int lgrp_packer_rows(const lgrp_packer *p, uint64_t *n_events, const uint8_t **kinds, uint64_t *n_rows, const uint32_t **values, const uint32_t **coefs);

/* ---- the three-stage prover over a witness matrix (needs a B200: runs the hot path through lgr.h) -- */
/* Event kinds.  0 / 1 come from the scalar backend (nonbatch_context.hpp:445-468); 2.. are vbn254fr host calls on
 * device-resident variables of k elements (include/host_modules/vbn254fr.hpp:139-566) together with the on_batch_*
 * callback each one triggers (nonbatch_context.hpp:497-553,782-847,996-1047).  Every vbn254fr event reads three u32
 * from batch_args (variable indices out, x, y -- or the bit index in the third place), constant-taking ones also 8 u32
 * from batch_consts; LGRP_EV_VSET consumes one row of `values` like a linear event. */
enum {
    LGRP_EV_LINEAR = 0,      /* 1 committed row */
    LGRP_EV_QUAD = 1,        /* 3 committed rows x, y, z */
    LGRP_EV_VSET = 2,        /* vbn254fr_set_*: out := values row, zero fill; on_batch_init: pads drawn into out[l..k), 1 row */
    LGRP_EV_VCOPY = 3,       /* vbn254fr_copy: out := in; on_batch_equal(out, in): 2 rows */
    LGRP_EV_VADD = 4,        /* out := x + y */
    LGRP_EV_VSUB = 5,        /* out := x - y */
    LGRP_EV_VMUL = 6,        /* tmp := x*y; on_batch_quadratic(x, y, tmp): 3 rows; out := tmp */
    LGRP_EV_VDIV = 7,        /* tmp := x/y; on_batch_quadratic(tmp, y, x): 3 rows; out := tmp */
    LGRP_EV_VASSERT_EQ = 8,  /* on_batch_equal(x = arg0, y = arg1): 2 rows */
    LGRP_EV_VBIT = 9,        /* tmp := bit arg2 of x; out := tmp; on_batch_bit(out): 1 row */
    LGRP_EV_VADDC = 10, LGRP_EV_VSUBC = 11, LGRP_EV_VCSUB = 12, LGRP_EV_VMULC = 13, LGRP_EV_VMONTMULC = 14   /* out := x (op) constant */
};

typedef struct {
    uint32_t l, k;                 /* must match the context (n = 4k) */
    uint64_t n_events;             /* events in emission order */
    const uint8_t *kinds;          /* n_events bytes, LGRP_EV_* */
    const uint32_t *values;        /* host rows in event order (1 per linear / VSET event, 3 per triple): [rows][l][8 x u32], canonical */
    const uint32_t *coefs;         /* linear-test coefficient rows, same shape (ignored for VSET rows); NULL = all zero */
    uint32_t const_sum[8];         /* linear test constant (zkp/common.hpp:68-79) */
    uint8_t encoding_seed[32];     /* AES-CTR key of the padding / mask stream (src/webgpu_prover.cpp:239-263) */
    uint8_t instance_hash[32];     /* src/webgpu_prover.cpp:161-168 */
    uint8_t program_hash[32];      /* metadata only */
    int64_t generated_at_seconds;  /* metadata; < 0 = wall clock */
    uint32_t sample_size;          /* params::sample_size = 192 */
    uint32_t arena_slots;          /* vbn254fr variables (0 when there are no LGRP_EV_V* events) */
    const uint32_t *batch_args;    /* 3 u32 per vbn254fr event, in event order */
    const uint32_t *batch_consts;  /* 8 u32 per constant-taking vbn254fr event, in event order */
} lgrp_statement;

int lgrp_prove(lgr_ctx *ctx, const lgrp_statement *st, lgrp_proof **out);

/* ---- interpreter boundary (BASELINE config 4) ------------------------------------------
 * A front end for WebAssembly programs over the env and wasi_snapshot_preview1 host modules and the witness emitter behind it
 * (host/wat_emitter.hpp, host/witness_machine.hpp).  `wat` is WebAssembly text (folded like the .wat files under the reference's tests/, or plain) or a WebAssembly
 * binary (it starts with "\0asm"): the reference's prover takes both (src/webgpu_prover.cpp:189-207).  Supported: every
 * integer instruction the reference implements (interpreter_impl.hpp:155-1309: const, add sub mul, div / rem, and or xor,
 * shifts and rotates, comparisons, clz ctz popcnt, extend / wrap; 32 and 64 bits), select, drop, nop, local.get / set / tee,
 * calls of the module's own functions, linear memory (loads / stores of every width, memory.size / grow / fill / copy / init,
 * data.drop, data segments), structured control flow (block / loop / if / br / br_if / br_table / return / unreachable), i32 / i64
 * globals, f32 / f64 arithmetic and conversions (numbers only, as in the reference), call_indirect, references and the table
 * instructions -- every instruction the reference's interpreter dispatches (interpreter_impl.hpp:2405-2548) -- and
 * env.i32_private_const, i64_private_const, assert_equal, assert_zero, assert_one, assert_constant, witness_cast_u32 / _u64,
 * assert_is_concrete, print_str, dump_memory; wasi args_sizes_get, args_get, fd_write, proc_exit, random_get (lgrp_wat_args below).
 * A run is bounded to 2 x 10^8 executed instructions (loops make running time a property of the guest); the environment variable
 * LGRP_WAT_STEP_LIMIT sets another bound, 0 none.
 * Not provided: the bn254fr / vbn254fr / uint256 / ecc host modules and the other WASI functions -- a module may import them (imports
 * resolve when they are called, as in the reference); CALLING one is an LGRP error naming it.  Refused when the module is read:
 * passive element segments, imported memories / tables / globals.
 * It stands where include/invoke.hpp:79-98 + include/interpreter_impl.hpp + the headers under include/zkp/backend/ stand in
 * the reference.  It gives every instruction the reference's
 * meaning -- the same witnesses, released in the same order, with the same linear-test randomness -- so the rows, the
 * stage-2 coefficient rows and const_sum are the reference's, element for element: checked against runs of the
 * reference's own interpreter / env and WASI modules / backend / witness manager (tests/refctx/ref_contexts.cpp,
 * tests/golden/refctx_*.json, tests/test_refctx_cpu.py) on all 69 programs of its tests/ and on random programs. */
typedef struct {
    uint64_t private_consts, asserts, arithmetic_ops;
    uint64_t linear_witnesses, quadratic_slots, linear_constraints;
    uint64_t violated_constraints;   /* > 0: an assertion of the program does not hold; the proof will not validate */
} lgrp_wat_stats;
/* host only: text -> rows.  stage1_seed == NULL: coefficient rows are zero (what stage 1 needs); otherwise the linear-test
 * coefficients are drawn from the linear stream keyed by the seed, one rho per constraint, and const_sum is set. */
int lgrp_wat_emit(const char *wat, size_t len, uint32_t l, const uint8_t *stage1_seed, lgrp_packer **rows_out, uint32_t const_sum[8],
                  lgrp_wat_stats *stats);
/* text or binary -> proof on the context's geometry: stage 1 on the values, coefficients from the stage-1 seed, stages 2 and 3.
 * program_hash = SHA-256 of the bytes handed in, instance_hash = 0 (no public arguments). */
int lgrp_prove_wat(lgr_ctx *ctx, const char *wat, size_t len, const uint8_t encoding_seed[32], int64_t generated_at_seconds,
                   lgrp_proof **out, lgrp_wat_stats *stats);

/* The guest's arguments, as the reference's prover builds them from its JSON configuration (src/webgpu_prover.cpp:69,110-157):
 * args[0] is the program name ("Ligero" with its NUL there), {"str": s} is the string with its NUL, {"i64": v} the 8
 * little-endian bytes, {"hex": h} the decoded bytes; "private-indices" names the secret ones.  The guest receives them through
 * wasi_snapshot_preview1.args_sizes_get / args_get (include/host_modules/wasi_preview1.hpp:52-100): the bytes of a private
 * argument are marked in linear memory, so every load that touches them commits a witness -- that is how private inputs
 * enter a proof.  The public ones are folded into instance_hash (src/webgpu_prover.cpp:160-168), which seeds stage 1. */
typedef struct {
    uint32_t nargs;
    const uint8_t *const *args;     /* nargs byte strings */
    const size_t *arg_lens;
    uint32_t nprivate;
    const int32_t *private_indices;
} lgrp_wat_args;
int lgrp_wat_instance_hash(const lgrp_wat_args *args, uint8_t out[32]);
/* lgrp_wat_emit / lgrp_prove_wat with arguments (args == NULL: none).  exit_code (may be NULL): what the guest passed to
 * proc_exit, -1 if it returned from _start. */
int lgrp_wat_emit_args(const char *wat, size_t len, const lgrp_wat_args *args, uint32_t l, const uint8_t *stage1_seed, lgrp_packer **rows_out,
                       uint32_t const_sum[8], lgrp_wat_stats *stats, int32_t *exit_code);
int lgrp_prove_wat_args(lgr_ctx *ctx, const char *wat, size_t len, const lgrp_wat_args *args, const uint8_t encoding_seed[32],
                        int64_t generated_at_seconds, lgrp_proof **out, lgrp_wat_stats *stats);

#ifdef __cplusplus
}
#endif
#endif

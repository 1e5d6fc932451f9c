"""The REFERENCE's own prover layers as the checker (CPU part).

tests/golden/refctx_*.json hold what the reference's stage contexts (include/zkp/nonbatch_context.hpp), its backend and
witness manager (include/zkp/backend/), its interpreter's opcode semantics (include/interpreter_impl.hpp) and its env /
vbn254fr host modules commit for three programs, when compiled where they lie and run over the CPU oracle as executor
(tests/refctx/ref_contexts.cpp, oracle/Makefile target `refctx`).  Here: the CPU restatement of the prover
(oracle/prover_ref.py, what the GPU prover is checked against elsewhere) must reproduce those runs bit for bit; the harness
must regenerate the committed vectors; and, where /root/reference is present, the reference's stage contexts must COMPILE
against the CUDA executor through the compat headers (the drop-in boundary as a build result, not a reading)."""
import json
import os
import shutil
import subprocess
import tempfile

import numpy as np
import pytest

from oracle import prover_ref as ref
import refctx_util as U

REFERENCE = os.environ.get("LGR_REFERENCE_ROOT", "/root/reference")


@pytest.mark.parametrize("case", U.CASES)
def test_restatement_reproduces_the_reference_run(oracle, case):
    """root, seeds, code / linear / quadratic test vectors, sampled columns and Merkle openings of the reference's own
    three passes == oracle/prover_ref.py on the statement those passes saw"""
    st = U.load(case)
    fx = st["fx"]
    out = ref.prove(st["l"], st["k"], st["kinds"], st["values"], st["coefs"], st["const_sum"], st["encoding_seed"], st["instance_hash"],
                    arena_slots=st["slots"], batch_args=st["args"], batch_consts=st["consts"])
    assert out["root"].hex() == fx["root"]
    assert U.sha(out["digests"]) == fx["sha256"]["digests"]
    assert out["stage1_seed"].hex() == fx["stage1_seed"]
    for name in ("code", "linear", "quad"):
        assert U.sha(out[name]) == fx["sha256"][name], name
    assert out["stage2_seed"].hex() == fx["stage2_seed"]
    assert [int(i) for i in out["sample"]] == fx["sample_index"]
    assert U.sha(out["samplings"]) == fx["sha256"]["samplings"]
    assert out["encoded_rows"] == out["samplings"].shape[0]
    assert list(out["valid"]) == [bool(v) for v in fx["valid"]] == [True, True, True]
    assert fx["verifier"] == [1] * 7            # ... and the reference's verifier accepted that run
    # openings: merkle_tree::decommit of the reference (merkle_tree.hpp:155-215) against the restated level walk
    assert out["total_count"] == fx["decommit_total"]
    assert sorted(out["positions"]) == fx["decommit_positions"]
    by_pos = dict(zip(out["positions"], out["siblings"]))
    assert U.sha(np.frombuffer(b"".join(by_pos[p] for p in fx["decommit_positions"]), np.uint8)) == fx["sha256"]["decommit_siblings"]
    # the same through the proof container (the check the GPU prover's envelope goes through in tests/test_refctx_gpu.py)
    meta = {"prover_version": "0", "program_hash": bytes(32), "generated_at": 1, "k": st["k"], "n": st["n"], "sample_size": 192}
    env = ref.build_envelope(meta, out["root"], out["siblings"], out["sample"], out["code"], out["linear"], out["quad"], out["samplings"])
    U.check_envelope(ref.parse_envelope(env), fx, ref.sibling_positions)


@pytest.mark.skipif(not os.path.exists(U.REF_BIN_CPU), reason="oracle/_ref/refctx_cpu not built (needs /root/reference at build time)")
@pytest.mark.parametrize("case", ["i64_mul3_k256", "vbn_k256", "intops_k256"])
def test_reference_verifier_accepts_the_restated_prover_and_rejects_tampering(oracle, case, tmp_path):
    """the reference's own verifier (nonbatch_verifier_context re-running the program over the sampled columns, recommit,
    the seven checks of src/webgpu_verifier.cpp:412-442) accepts the proof oracle/prover_ref.py makes -- the proof the
    GPU prover is held to byte for byte -- and rejects it when one sampled value, one test-vector element or the
    instance hash is changed"""
    st = U.load(case)
    out = ref.prove(st["l"], st["k"], st["kinds"], st["values"], st["coefs"], st["const_sum"], st["encoding_seed"], st["instance_hash"],
                    arena_slots=st["slots"], batch_args=st["args"], batch_consts=st["consts"])

    def verdict(tag, **change):
        f = dict(root=out["root"], code=out["code"], linear=out["linear"], quad=out["quad"], samplings=out["samplings"], positions=out["positions"],
                 siblings=out["siblings"], total_count=out["total_count"], instance_hash=st["instance_hash"])
        f.update(change)
        path = str(tmp_path / (tag + ".proof"))
        U.write_proof_file(path, **f)
        return U.reference_verifier(U.REF_BIN_CPU, case, path, str(tmp_path))

    rc, msg = verdict("honest")
    assert rc == 0, msg
    bad = out["samplings"].copy(); bad[1, 5, 0] ^= 1
    rc, msg = verdict("sample", samplings=bad)
    assert rc == 1 and "merkle 0" in msg, msg
    bad = out["quad"].copy(); bad[3, 0] ^= 1
    assert verdict("quad", quad=bad)[0] == 1
    assert verdict("instance", instance_hash=bytes(32))[0] == 1


def test_i64_mul_row_counts_at_the_default_geometry():
    """BASELINE config 4 (tests/i64_mul.wat at k = 8192): the reference commits 1 linear row + 1 quadratic triple + 3 masks"""
    st = U.load("i64_mul_k8192")
    assert list(st["kinds"]) == [ref.EV_LINEAR, ref.EV_QUAD]
    assert st["values"].shape[0] == 4
    # 27 private constants of 64 bits + 9 products of 128 bits, every bit slot b*b = b, plus the 9 products themselves
    quad_slots = int(st["values"][1].any(axis=1).sum())
    assert quad_slots <= 27 * 64 + 9 * 128 + 9 and (st["values"][3] != 0).any()


@pytest.mark.skipif(not os.path.exists(U.REF_BIN_CPU), reason="oracle/_ref/refctx_cpu not built (needs /root/reference at build time)")
@pytest.mark.parametrize("case", ["i64_mul3_k256", "vbn_k256", "mul64_k256", "arith32_k256"])
def test_harness_regenerates_the_committed_vectors(oracle, case):
    """the committed JSON is exactly what the harness writes today (same reference sources, same oracle)"""
    gen = U.compact_module()
    with tempfile.TemporaryDirectory() as tmp:
        prog, k, name = U.harness_args(case, tmp)
        path = os.path.join(tmp, "out.json")
        subprocess.check_call([U.REF_BIN_CPU, prog, k, path], stdout=subprocess.DEVNULL)
        raw = json.load(open(path))
        raw["program"] = name
        got = gen.compact(raw)
    want = U.load(case)["fx"]
    for key in want:
        assert got[key] == want[key], key


# ---------------------------------------------------------------- the .wat front end against the reference's interpreter
@pytest.fixture(scope="module")
def pr(lgr):
    import importlib
    return importlib.import_module("ligero_prover_b200.prover")


def _emitter_equals_reference_rows(pr, text, st):
    fx = st["fx"]
    kinds, vals, coefs, const_sum, stats = pr.wat_emit(text, st["l"], bytes.fromhex(fx["stage1_seed"]))
    assert list(kinds) == list(st["kinds"])
    assert np.array_equal(vals, st["values"]), "rows differ from what the reference's interpreter + witness manager emit"
    assert np.array_equal(coefs, st["coefs"]), "linear-test coefficient rows differ from the reference's"
    assert const_sum == st["const_sum"]
    assert stats["violated_constraints"] == 0
    return stats


def test_wat_emitter_reproduces_the_reference_rows_for_i64_mul(pr):
    """BASELINE config 4: tests/i64_mul.wat at the default geometry.  The rows and the stage-2 coefficient rows that
    lgrp_wat_emit / lgrp_prove_wat feed the prover are, element for element, the ones the reference's own interpreter, env
    module, backend and witness manager hand to its stage contexts (tests/golden/refctx_i64_mul_k8192.json)"""
    st = U.load("i64_mul_k8192")
    path = os.path.join(REFERENCE, "tests", "i64_mul.wat")
    text = open(path).read() if os.path.exists(path) else U.binop_wat("mul", U.I64_MUL_CASES)
    stats = _emitter_equals_reference_rows(pr, text, st)
    assert (stats["private_consts"], stats["asserts"], stats["arithmetic_ops"], stats["quadratic_slots"]) == (27, 9, 9, 2889)
    assert _emitter_equals_reference_rows(pr, U.binop_wat("mul", U.I64_MUL_CASES), st) == stats
    # the last three assertions at l = 64: 17 row events, rows of both kinds interleaved
    _emitter_equals_reference_rows(pr, U.binop_wat("mul", U.I64_MUL_CASES[6:]), U.load("i64_mul3_k256"))


@pytest.mark.parametrize("name", ["mul64", "arith32", "intops"])
def test_wat_emitter_reproduces_the_reference_rows_for_the_repo_programs(pr, name):
    """tests/golden/mul64.wat and arith32.wat: products, sums, differences, nested forms and literal operands in 64 and 32
    bits (39 and 19 row events at l = 64); intops.wat: every other integer instruction -- bitwise, shifts, rotates,
    comparisons, counts, division, extensions -- and forms that chain them (109 row events)"""
    st = U.load(name + "_k256")
    _emitter_equals_reference_rows(pr, open(U.WAT_TEXT[name]).read(), st)


def _rand_expr(rng, depth, w=64):
    """(text, value mod 2^w) of a random folded expression over private and literal operands of width w"""
    M = 1 << w
    if depth == 0 or rng.random() < 0.25:
        v = rng.choice([0, 1, 2, M - 1, M >> 1, (M >> 1) - 1, 1 << (w // 2), rng.getrandbits(w), rng.getrandbits(20)])
        lit = "(i%d.const %d)" % (w, v)
        return ("(call $i%d_private_const %s)" % (w, lit) if rng.random() < 0.7 else lit), v
    op = rng.choice(["mul", "add", "sub"])
    (ta, va), (tb, vb) = _rand_expr(rng, depth - 1, w), _rand_expr(rng, depth - 1, w)
    return "(i%d.%s %s %s)" % (w, op, ta, tb), {"mul": va * vb, "add": va + vb, "sub": va - vb}[op] % M


def _rand_program(rng, w):
    exprs = [_rand_expr(rng, rng.randrange(1, 4), w) for _ in range(4)]
    rhs = lambda v: ("(i%d.const %d)" % (w, v)) if rng.random() < 0.5 else ("(call $i%d_private_const (i%d.const %d))" % (w, w, v))
    head = U.WAT_HEAD if w == 64 else U.WAT_HEAD32
    return head + "".join("(call $assert_equal %s %s)\n" % (t, rhs(v)) for t, v in exprs) + U.WAT_TAIL


@pytest.mark.skipif(not os.path.exists(U.REF_BIN_CPU), reason="oracle/_ref/refctx_cpu not built (needs /root/reference at build time)")
@pytest.mark.parametrize("seed,w", [(0, 64), (1, 64), (2, 64), (3, 32), (4, 32), (5, 32)])
def test_wat_emitter_against_the_reference_interpreter_on_random_programs(oracle, pr, seed, w):
    """differential: random expression trees (private and literal operands, nested mul / add / sub, 64- and 32-bit) run
    through the reference's interpreter + backend (oracle/_ref/refctx_cpu) and through the emitter give the same rows"""
    import random
    rng = random.Random(4200 + seed)
    text = _rand_program(rng, w)
    k = rng.choice([256, 512])
    raw = U.run_reference_on_wat(text, k, seed_byte=seed + 1)
    assert raw["valid"] == [1, 1, 1] and raw["verifier"] == [1] * 7
    _emitter_equals_reference_rows(pr, text, _reference_rows(raw))


def _reference_rows(raw):
    l = raw["l"]
    u32 = lambda h: np.frombuffer(bytes.fromhex(h), np.uint32)
    return {"fx": raw, "l": l, "kinds": raw["kinds"], "values": u32(raw["values"]).reshape(-1, l, 8), "coefs": u32(raw["coefs"]).reshape(-1, l, 8),
            "const_sum": int.from_bytes(bytes.fromhex(raw["const_sum"]), "little")}


@pytest.mark.skipif(not os.path.exists(U.REF_BIN_CPU), reason="oracle/_ref/refctx_cpu not built (needs /root/reference at build time)")
@pytest.mark.parametrize("seed", range(12))
def test_wat_emitter_against_the_reference_interpreter_on_every_integer_instruction(oracle, pr, seed):
    """differential over the whole integer instruction set: random trees of clz / ctz / popcnt / eqz / extend / add / sub / mul /
    and / or / xor / shifts / rotates / comparisons / div / rem with private and literal leaves -- so bit-vector results,
    single-witness results and concrete numbers meet in every combination -- through the reference's interpreter + backend
    and through the emitter: same rows, same coefficient rows, same const_sum, and the reference's verifier accepts"""
    import random
    rng = random.Random(7700 + seed)
    w = (32, 64)[seed & 1]
    text, _ = U.rand_int_program(rng, w, nexpr=rng.randrange(1, 4), depth=rng.randrange(1, 4))
    raw = U.run_reference_on_wat(text, 256, seed_byte=seed + 1)
    assert raw["valid"] == [1, 1, 1] and raw["verifier"] == [1] * 7
    _emitter_equals_reference_rows(pr, text, _reference_rows(raw))


def test_signed_division_by_the_most_negative_number_is_rejected_as_the_reference_rejects_it(pr):
    """a quirk kept: the reference's div_s / rem_s gadget range-checks the remainder against |y| with a SIGNED comparison
    (interpreter_impl.hpp:452-456), so y = -2^(w-1), whose absolute value has the top bit set, fails the reference's own
    self-check ("valid 101" from refctx_cpu); the emitter flags the same constraint instead of emitting a proof that
    cannot verify"""
    text = U.WAT_HEAD_BOTH + "(call $assert_equal (i32.rem_s (call $i32_private_const (i32.const 3)) (call $i32_private_const (i32.const 0x80000000))) (i32.const 3))\n" + U.WAT_TAIL
    assert pr.wat_emit(text, 64)[4]["violated_constraints"] == 1
    ok = text.replace("0x80000000", "0x80000001")
    assert pr.wat_emit(ok, 64)[4]["violated_constraints"] == 0


ENV_PROGRAM = ('(module (import "env" "i64_private_const" (func $pc (param i64) (result i64)))\n(import "env" "assert_zero" (func $z (param i64)))\n'
               '(import "env" "assert_one" (func $o (param i64)))\n(import "env" "assert_constant" (func $c (param i64)))\n'
               '(import "env" "witness_cast_u64" (func $cast (param i64) (result i64)))\n(import "env" "assert_is_concrete" (func $conc (param i64)))\n'
               '(import "env" "assert_equal" (func $eq (param i64 i64)))\n(func $t\n'
               '(call $z (i64.sub (call $pc (i64.const 9)) (call $pc (i64.const 9))))\n(call $o (i64.eqz (call $pc (i64.const 0))))\n'
               '(call $c (i64.add (call $pc (i64.const 40)) (i64.const 2)))\n(call $eq (call $cast (call $pc (i64.const 77))) (i64.const 77))\n'
               '(call $conc (i64.mul (i64.const 6) (i64.const 7)))\n(drop (call $pc (i64.const 5)))\n(call $z (call $cast (i64.const 0)))\n'
               '(call $o (call $cast (i64.popcnt (call $pc (i64.const 256)))))\n(call $c (i64.const 12))\n(drop (i64.clz (call $pc (i64.const 3))))\n(nop)\n'
               ')\n(export "_start" (func $t)))\n')


@pytest.mark.skipif(not os.path.exists(U.REF_BIN_CPU), reason="oracle/_ref/refctx_cpu not built (needs /root/reference at build time)")
def test_wat_emitter_env_assertions_casts_and_drop_against_the_reference(oracle, pr):
    """env.assert_zero / assert_one / assert_constant / witness_cast_u64 / assert_is_concrete, `drop` and `nop` through the
    reference's env module + interpreter and through the emitter: same rows, coefficient rows and const_sum; the binary
    spelling of the program too"""
    raw = U.run_reference_on_wat(ENV_PROGRAM, 256, seed_byte=9)
    assert raw["valid"] == [1, 1, 1] and raw["verifier"] == [1] * 7
    st = _reference_rows(raw)
    _emitter_equals_reference_rows(pr, ENV_PROGRAM, st)
    _emitter_equals_reference_rows(pr, U.wat_to_wasm(ENV_PROGRAM), st)


STRUCT_PROGRAM = """(module
 (import "env" "i64_private_const" (func $pc (param i64) (result i64)))
 (import "env" "i32_private_const" (func $pc32 (param i32) (result i32)))
 (import "env" "assert_equal" (func $eq (param i64 i64)))
 (func $sq (param $a i64) (result i64) (i64.mul (local.get $a) (local.get $a)))
 (func $mix (param $a i64) (param $b i64) (result i64) (local $t i64)
   (local.set $t (i64.xor (local.get $a) (local.get $b)))
   (i64.add (local.get $t) (call $sq (local.get $b))))
 (func $main (local $x i64) (local $y i64) (local $c i32)
   (local.set $x (call $pc (i64.const 7)))
   (local.set $y (i64.add (local.tee $x (i64.add (local.get $x) (i64.const 1))) (call $pc (i64.const 5))))
   (call $eq (local.get $y) (i64.const 13))
   (call $eq (call $mix (local.get $x) (local.get $y)) (i64.const 174))
   (local.set $c (i64.lt_u (local.get $x) (local.get $y)))
   (call $eq (select (local.get $x) (local.get $y) (local.get $c)) (i64.const 8))
   (call $eq (select (local.get $x) (local.get $y) (i32.const 0)) (i64.const 13))
   (call $eq (select (call $pc (i64.const 1)) (i64.const 2) (call $pc32 (i32.const 0))) (call $pc (i64.const 2)))
   (local.set $x (i64.const 3))
   (local.set $y (i64.clz (local.get $y)))
   (call $eq (local.get $y) (i64.const 60))
 )
 (export "_start" (func $main)))
"""


@pytest.mark.skipif(not os.path.exists(U.REF_BIN_CPU), reason="oracle/_ref/refctx_cpu not built (needs /root/reference at build time)")
def test_wat_emitter_locals_select_and_module_functions_against_the_reference(oracle, pr):
    """locals (get / set / tee, overwritten while they hold bits, witnesses and numbers, alive at the end of the frame), select
    on concrete and on witness conditions, calls of module functions with parameters, locals and nested calls: through the
    reference's interpreter (frames, run_call, exec_select, exec_local_*) and through the emitter in all three spellings"""
    raw = U.run_reference_on_wat(STRUCT_PROGRAM, 256, seed_byte=5)
    assert raw["valid"] == [1, 1, 1] and raw["verifier"] == [1] * 7
    st = _reference_rows(raw)
    for spelling in (STRUCT_PROGRAM, U.wat_to_wasm(STRUCT_PROGRAM), U.wat_to_plain(STRUCT_PROGRAM)):
        _emitter_equals_reference_rows(pr, spelling, st)


@pytest.mark.skipif(not os.path.exists(U.REF_BIN_CPU), reason="oracle/_ref/refctx_cpu not built (needs /root/reference at build time)")
@pytest.mark.parametrize("seed", range(10))
def test_wat_emitter_against_the_reference_interpreter_on_structured_programs(oracle, pr, seed):
    """differential: random programs with three locals, helper functions, select, local.tee inside expressions and every
    integer instruction (tests/refctx_util.py: rand_struct_program), text and binary"""
    import random
    rng = random.Random(8800 + seed)
    text = U.rand_struct_program(rng, (32, 64)[seed & 1], nstmt=rng.randrange(2, 7), depth=rng.randrange(1, 4))
    raw = U.run_reference_on_wat(text, 256, seed_byte=seed + 1)
    assert raw["valid"] == [1, 1, 1] and raw["verifier"] == [1] * 7
    st = _reference_rows(raw)
    _emitter_equals_reference_rows(pr, text, st)
    _emitter_equals_reference_rows(pr, U.wat_to_wasm(text), st)


MEMORY_PROGRAM = """(module
 (import "env" "i64_private_const" (func $pc (param i64) (result i64)))
 (import "env" "i32_private_const" (func $pc32 (param i32) (result i32)))
 (import "env" "assert_equal" (func $eq (param i64 i64)))
 (memory 1 2)
 (data (i32.const 64) "\\01\\02\\03\\04\\05\\06\\07\\08")
 (data $seg "ab\\ff\\00")
 (func $main (local $p i32)
   (i64.store (i32.const 8) (call $pc (i64.const 0x1122334455667788)))
   (call $eq (i64.load (i32.const 8)) (i64.const 0x1122334455667788))
   (call $eq (i32.load offset=4 (i32.const 8)) (i32.const 0x11223344))
   (call $eq (i64.load8_s offset=7 (i32.const 4)) (i64.const 0x55))
   (call $eq (i32.load16_u (i32.const 8)) (i32.const 0x7788))
   (call $eq (i32.load (i32.const 64)) (i32.const 0x04030201))
   (i32.store16 (i32.const 10) (i32.const 0xbeef))
   (call $eq (i64.load (i32.const 8)) (call $pc (i64.const 0x11223344beef7788)))
   (memory.copy (i32.const 32) (i32.const 8) (i32.const 8))
   (call $eq (i64.load (i32.const 32)) (i64.const 0x11223344beef7788))
   (call $eq (i32.load (i32.const 34)) (i32.const 0x3344beef))
   (memory.fill (i32.const 36) (i32.const 0xAA) (i32.const 2))
   (call $eq (i64.load (i32.const 32)) (i64.const 0x1122aaaabeef7788))
   (memory.init $seg (i32.const 100) (i32.const 1) (i32.const 3))
   (call $eq (i32.load (i32.const 100)) (i32.const 0x00ff62))
   (data.drop $seg)
   (i32.store8 (i32.const 101) (call $pc32 (i32.const 0x1ff)))
   (call $eq (i32.load8_s (i32.const 101)) (i32.const -1))
   (call $eq (i32.load (i32.const 100)) (i32.const 0xff62))
   (local.set $p (call $pc32 (i32.const 200)))
   (i64.store32 (local.get $p) (i64.const 0x9988776655443322))
   (call $eq (i64.load32_u (i32.const 200)) (i64.const 0x55443322))
   (call $eq (i64.load32_s offset=0 (local.get $p)) (i64.const 0x55443322))
   (call $eq (memory.size) (i32.const 1))
   (call $eq (memory.grow (i32.const 1)) (i32.const 1))
   (call $eq (memory.grow (i32.const 1)) (i32.const -1))
   (i64.store (i32.const 70000) (call $pc (i64.const 5)))
   (call $eq (i64.load (i32.const 70000)) (i64.const 5))
   (memory.copy (i32.const 4) (i32.const 8) (i32.const 8))
   (call $eq (i64.load (i32.const 4)) (i64.const 0x11223344beef7788))
   (memory.copy (i32.const 12) (i32.const 8) (i32.const 8))
   (call $eq (i32.load (i32.const 16)) (i32.const 0x11223344))
 )
 (export "_start" (func $main)))
"""


@pytest.mark.skipif(not os.path.exists(U.REF_BIN_CPU), reason="oracle/_ref/refctx_cpu not built (needs /root/reference at build time)")
def test_wat_emitter_linear_memory_against_the_reference(oracle, pr):
    """loads and stores of every width (with offsets, through witness addresses), memory.size / grow / fill / copy / init,
    data.drop, active and passive data segments: the reference's memory keeps concrete bytes plus the set of ranges a
    witness was stored to; a load touching such a range yields a fresh witness (so the marks decide which rows exist).
    Same rows through the reference's interpreter and through the emitter, in all three spellings"""
    raw = U.run_reference_on_wat(MEMORY_PROGRAM, 256, seed_byte=3)
    assert raw["valid"] == [1, 1, 1] and raw["verifier"] == [1] * 7
    st = _reference_rows(raw)
    for spelling in (MEMORY_PROGRAM, U.wat_to_wasm(MEMORY_PROGRAM), U.wat_to_plain(MEMORY_PROGRAM)):
        _emitter_equals_reference_rows(pr, spelling, st)


@pytest.mark.skipif(not os.path.exists(U.REF_BIN_CPU), reason="oracle/_ref/refctx_cpu not built (needs /root/reference at build time)")
@pytest.mark.parametrize("seed", range(10))
def test_wat_emitter_against_the_reference_interpreter_on_memory_programs(oracle, pr, seed):
    """differential: random stores / loads / fills / overlapping copies / inits over a small window (tests/refctx_util.py:
    rand_memory_program): marked ranges are split, joined, moved and cleared in every order; text and binary"""
    import random
    rng = random.Random(9900 + seed)
    text = U.rand_memory_program(rng, nstmt=rng.randrange(6, 24))
    raw = U.run_reference_on_wat(text, 256, seed_byte=seed + 1)
    assert raw["valid"] == [1, 1, 1] and raw["verifier"] == [1] * 7
    st = _reference_rows(raw)
    _emitter_equals_reference_rows(pr, text, st)
    _emitter_equals_reference_rows(pr, U.wat_to_wasm(text), st)


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "tests")) or not os.path.exists(U.REF_BIN_CPU), reason="needs the reference tree and oracle/_ref/refctx_cpu")
@pytest.mark.parametrize("name", ["memory_fill_clears_secret_tag", "memory_init_clears_secret_tag"])
def test_wat_emitter_on_the_reference_memory_programs(oracle, pr, name):
    """the two memory programs of the reference's tests/: with them, every program of that directory that commits a
    witness goes through the emitter (f32.wat / f64.wat compute on concrete numbers only and commit nothing)"""
    text = open(os.path.join(REFERENCE, "tests", name + ".wat")).read()
    raw = U.run_reference_on_wat(text, 256)
    assert raw["valid"] == [1, 1, 1] and raw["verifier"] == [1] * 7
    st = _reference_rows(raw)
    _emitter_equals_reference_rows(pr, text, st)
    _emitter_equals_reference_rows(pr, U.wat_to_wasm(text), st)


CONTROL_PROGRAM = """(module
 (import "env" "i64_private_const" (func $pc (param i64) (result i64)))
 (import "env" "i32_private_const" (func $pc32 (param i32) (result i32)))
 (import "env" "assert_equal" (func $eq (param i64 i64)))
 (func $sum (param $n i32) (result i64) (local $acc i64) (local $i i32)
   (block $done
     (loop $again
       (br_if $done (i32.ge_u (local.get $i) (local.get $n)))
       (local.set $acc (i64.add (local.get $acc) (call $pc (i64.extend_i32_u (local.get $i)))))
       (local.set $i (i32.add (local.get $i) (i32.const 1)))
       (br $again)))
   (local.get $acc))
 (func $early (param $a i64) (result i64) (local $t i64)
   (local.set $t (i64.mul (local.get $a) (local.get $a)))
   (if (i64.gt_u (local.get $a) (i64.const 3)) (then (return (local.get $t))))
   (i64.add (local.get $t) (i64.const 100)))
 (func $main (local $x i64) (local $y i64) (local $i i32)
   (call $eq (call $sum (i32.const 4)) (i64.const 6))
   (if (i32.const 1) (then (call $eq (call $pc (i64.const 3)) (i64.const 3))) (else (unreachable)))
   (call $eq (if (result i64) (call $pc32 (i32.const 1)) (then (call $pc (i64.const 7))) (else (i64.const 8))) (i64.const 7))
   (call $eq (block $b (result i64) (drop (br_if $b (call $pc (i64.const 4)) (i32.const 1))) (i64.const 5)) (i64.const 4))
   (call $eq (block $b (result i64) (call $pc (i64.const 1)) (call $pc (i64.const 2)) (i64.clz (call $pc (i64.const 9))) (br $b)) (i64.const 60))
   (local.set $x (call $pc (i64.const 5)))
   (block $o (block $in (br_table $in $o $in (i32.const 1))) (unreachable))
   (call $eq (call $early (call $pc (i64.const 5))) (i64.const 25))
   (call $eq (call $early (call $pc (i64.const 2))) (i64.const 104))
   (block $b (local.set $y (call $pc (i64.const 6))) (call $pc (i64.const 9)) (i64.popcnt (call $pc (i64.const 7))) (call $pc (i64.const 8)) (br $b))
   (call $eq (local.get $y) (i64.const 6))
   (loop $l (result i64) (call $pc (i64.const 3)) (local.set $i (i32.add (local.get $i) (i32.const 1))) (br_if $l (i32.lt_u (local.get $i) (i32.const 3))))
   (drop)
 )
 (export "_start" (func $main)))
"""


@pytest.mark.skipif(not os.path.exists(U.REF_BIN_CPU), reason="oracle/_ref/refctx_cpu not built (needs /root/reference at build time)")
def test_wat_emitter_control_flow_against_the_reference(oracle, pr):
    """block / loop / if / else / br / br_if / br_table / return: labels and frames live on the operand stack and a branch
    drops everything above its label -- witnesses, bit vectors and frames in the order drop_n_below and std::vector::erase
    give.  Counted loops, conditions that are witnesses (read as numbers), value-carrying blocks, branches over mixed
    values, early returns: same rows through the reference's interpreter and the emitter, in all three spellings"""
    raw = U.run_reference_on_wat(CONTROL_PROGRAM, 256, seed_byte=4)
    assert raw["valid"] == [1, 1, 1] and raw["verifier"] == [1] * 7
    st = _reference_rows(raw)
    for spelling in (CONTROL_PROGRAM, U.wat_to_wasm(CONTROL_PROGRAM), U.wat_to_plain(CONTROL_PROGRAM)):
        _emitter_equals_reference_rows(pr, spelling, st)


@pytest.mark.skipif(not os.path.exists(U.REF_BIN_CPU), reason="oracle/_ref/refctx_cpu not built (needs /root/reference at build time)")
@pytest.mark.parametrize("seed", range(10))
def test_wat_emitter_against_the_reference_interpreter_on_control_flow_programs(oracle, pr, seed):
    """differential: random programs with if / else, blocks left early, branches that drop values, br_table over nested
    blocks, counted loops and early returns around random integer statements (tests/refctx_util.py: rand_cf_program)"""
    import random
    rng = random.Random(12300 + seed)
    text = U.rand_cf_program(rng, (32, 64)[seed & 1], nstmt=rng.randrange(2, 7), depth=rng.randrange(1, 3))
    raw = U.run_reference_on_wat(text, 256, seed_byte=seed + 1)
    assert raw["valid"] == [1, 1, 1] and raw["verifier"] == [1] * 7
    st = _reference_rows(raw)
    _emitter_equals_reference_rows(pr, text, st)
    _emitter_equals_reference_rows(pr, U.wat_to_wasm(text), st)


@pytest.mark.skipif(not os.path.exists(U.REF_BIN_CPU), reason="oracle/_ref/refctx_cpu not built (needs /root/reference at build time)")
@pytest.mark.parametrize("seed", range(10))
def test_wat_emitter_against_the_reference_interpreter_on_floating_point_and_globals(oracle, pr, seed):
    """differential: random programs of f32 / f64 arithmetic (special values, NaN payloads, signed zeros, subnormals),
    conversions in every direction, saturating and trapping truncations, float loads / stores / locals / select, and
    mutable globals -- numbers only in the reference (interpreter_impl.hpp:1314-1853,1902-1924) -- whose results are all
    committed as integers, so the reference's rows carry every bit of them (tests/refctx_util.py: rand_float_program)"""
    import random
    rng = random.Random(4400 + seed)
    text = U.rand_float_program(rng, nstmt=rng.randrange(4, 14))
    raw = U.run_reference_on_wat(text, 256, seed_byte=seed + 1)
    assert raw["valid"] == [1, 1, 1] and raw["verifier"] == [1] * 7
    st = _reference_rows(raw)
    for spelling in (text, U.wat_to_wasm(text), U.wat_to_plain(text)):
        _emitter_equals_reference_rows(pr, spelling, st)


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "tests")) or not os.path.exists(U.REF_BIN_CPU), reason="needs the reference tree and oracle/_ref/refctx_cpu")
@pytest.mark.parametrize("name", ["f32", "f64"])
def test_wat_emitter_on_the_reference_floating_point_programs(oracle, pr, name):
    """tests/f32.wat and tests/f64.wat, read where they lie (plain instructions in their helper functions, folded forms in
    _start): programs that check themselves -- every result is compared with its expected value and `unreachable` ends the
    run otherwise.  They commit no witness: the reference's interpreter runs them to the end with no row, and so does the
    emitter, from the text, the binary and the plain spelling"""
    text = open(os.path.join(REFERENCE, "tests", name + ".wat")).read()
    raw = U.run_reference_on_wat(text, 256)
    assert raw["kinds"] == [] and raw["valid"] == [1, 1, 1]
    for spelling in (text, U.wat_to_wasm(text), U.wat_to_plain(text)):
        kinds, vals, _, _, stats = pr.wat_emit(spelling, 64, bytes(32))
        assert len(kinds) == 0 and len(vals) == 0 and stats["violated_constraints"] == 0
    # and a wrong expectation is noticed: the program runs into its `unreachable`
    broken = text.replace("(f%s.const 6.5)" % name[1:], "(f%s.const 6.75)" % name[1:], 1)
    assert broken != text
    with pytest.raises(pr.ProverError, match="unreachable executed"):
        pr.wat_emit(broken, 64, bytes(32))


def _fixture_args(st):
    return [bytes.fromhex(a) for a in st["fx"]["args"]], st["fx"]["private_indices"]


def test_wat_emitter_reproduces_the_reference_rows_for_a_guest_with_arguments(pr):
    """tests/golden/wasi_args.wat with {"args": [{"i64": 400}, {"i64": 600}, {"str": "hello"}], "private-indices": [1, 3]}: the
    guest fetches its arguments through wasi args_sizes_get / args_get, the bytes of the private ones are marked in memory and
    loads that touch them commit witnesses (even one that straddles a public and a private argument); it prints through
    fd_write, reads the reference's constant-seeded random bytes and leaves through proc_exit from inside a call with
    witnesses alive on the stack and in locals.  Rows, coefficient rows and const_sum are those of the reference run
    (its interpreter + wasi_preview1 / env modules: tests/golden/refctx_wasi_k256.json), in all three spellings"""
    st = U.load("wasi_k256")
    args, private = _fixture_args(st)
    assert args == pr.config_args([{"i64": 400}, {"i64": 600}, {"str": "hello"}])
    text = open(os.path.join(U.HERE, "golden", "wasi_args.wat")).read()
    for spelling in (text, U.wat_to_wasm(text), U.wat_to_plain(text)):
        kinds, vals, coefs, const_sum, stats, code = pr.wat_emit(spelling, st["l"], bytes.fromhex(st["fx"]["stage1_seed"]), args=args, private_indices=private,
                                                               want_exit_code=True)
        assert list(kinds) == list(st["kinds"]) and np.array_equal(vals, st["values"]) and np.array_equal(coefs, st["coefs"])
        assert const_sum == st["const_sum"] and stats["violated_constraints"] == 0 and code == 3
    # with other private indices other rows exist: nothing private -> the loads are numbers and far fewer witnesses are committed
    kinds_pub, _, _, _, stats_pub = pr.wat_emit(text, st["l"], args=args, private_indices=())
    assert stats_pub["violated_constraints"] == 0 and len(kinds_pub) < len(st["kinds"])


@pytest.mark.skipif(not os.path.exists(U.REF_BIN_CPU), reason="oracle/_ref/refctx_cpu not built (needs /root/reference at build time)")
@pytest.mark.parametrize("private", [(), (1,), (2, 3), (0, 1, 2, 3)])
def test_wat_emitter_wasi_arguments_against_the_reference(oracle, pr, private):
    """live differential of the same guest with other arguments and other choices of the private ones (argv[0] included)"""
    import random
    rng = random.Random(77 + len(private))
    a = rng.randrange(1000)
    args = [b"Ligero\0", a.to_bytes(8, "little"), (1000 - a).to_bytes(8, "little"), b"hello\0"]
    text = open(os.path.join(U.HERE, "golden", "wasi_args.wat")).read()
    raw = U.run_reference_on_wat(text, 256, seed_byte=5, args=args, private_indices=private)
    assert raw["valid"] == [1, 1, 1] and raw["verifier"] == [1] * 7
    st = _reference_rows(raw)
    kinds, vals, coefs, const_sum, stats = pr.wat_emit(text, st["l"], bytes.fromhex(st["fx"]["stage1_seed"]), args=args, private_indices=private)
    assert list(kinds) == list(st["kinds"]) and np.array_equal(vals, st["values"]) and np.array_equal(coefs, st["coefs"]) and const_sum == st["const_sum"]


INDIRECT_PROGRAM = open(os.path.join(U.HERE, "golden", "indirect.wat")).read()


@pytest.mark.skipif(not os.path.exists(U.REF_BIN_CPU), reason="oracle/_ref/refctx_cpu not built (needs /root/reference at build time)")
def test_wat_emitter_indirect_calls_against_the_reference(oracle, pr):
    """call_indirect through a function table filled by element segments (run_call_indirect, interpreter.hpp:372-398): a
    dispatch loop over add / mul / sub with a witness accumulator, an indirect call inside an indirectly called function, type
    uses by name and inline, functions declared by (type $t); same rows through the reference's interpreter and the emitter
    in all three spellings (the binary carries table and element sections)"""
    raw = U.run_reference_on_wat(INDIRECT_PROGRAM, 256, seed_byte=6)
    assert raw["valid"] == [1, 1, 1] and raw["verifier"] == [1] * 7
    st = _reference_rows(raw)
    for spelling in (INDIRECT_PROGRAM, U.wat_to_wasm(INDIRECT_PROGRAM), U.wat_to_plain(INDIRECT_PROGRAM)):
        _emitter_equals_reference_rows(pr, spelling, st)


@pytest.mark.skipif(not os.path.exists(U.REF_BIN_CPU), reason="oracle/_ref/refctx_cpu not built (needs /root/reference at build time)")
def test_wat_emitter_references_and_table_instructions_against_the_reference(oracle, pr):
    """the last instructions of the reference's interpreter the front end did not take (interpreter_impl.hpp:1926-2106):
    ref.null / ref.is_null / ref.func, table.get / set / size / grow / fill / copy / init, elem.drop, typed select, funcref
    parameters and results (tests/golden/tables.wat).  Every result is committed and asserted, the reference's own run is valid,
    and the rows agree in all three spellings (the binary uses the 0xD0-0xD2, 0x25 / 0x26 and 0xFC 12-17 encodings)"""
    text = open(os.path.join(U.HERE, "golden", "tables.wat")).read()
    raw = U.run_reference_on_wat(text, 256, seed_byte=8)
    assert raw["valid"] == [1, 1, 1] and raw["verifier"] == [1] * 7
    st = _reference_rows(raw)
    for spelling in (text, U.wat_to_wasm(text), U.wat_to_plain(text)):
        _emitter_equals_reference_rows(pr, spelling, st)


@pytest.mark.skipif(not os.path.exists(U.REF_BIN_CPU), reason="oracle/_ref/refctx_cpu not built (needs /root/reference at build time)")
def test_wat_emitter_on_a_module_spelled_like_compiled_code_against_the_reference(oracle, pr):
    """tests/golden/compiled_style.wat (the way wasm2wat prints a C guest: shadow stack in a global, WASI start-up, a private
    argument, a loop, an indirect call, proc_exit) with a private argument: the reference's interpreter and the emitter give
    the same rows from the text, the binary and the plain spelling"""
    text = open(os.path.join(U.HERE, "golden", "compiled_style.wat")).read()
    args = [b"Ligero\0", (7).to_bytes(8, "little")]
    raw = U.run_reference_on_wat(text, 256, seed_byte=9, args=args, private_indices=[1])
    assert raw["valid"] == [1, 1, 1] and raw["verifier"] == [1] * 7
    st = _reference_rows(raw)
    for spelling in (text, U.wat_to_wasm(text), U.wat_to_plain(text)):
        kinds, vals, coefs, const_sum, stats = pr.wat_emit(spelling, st["l"], bytes.fromhex(st["fx"]["stage1_seed"]), args=args, private_indices=[1])
        assert list(kinds) == list(st["kinds"]) and np.array_equal(vals, st["values"]) and np.array_equal(coefs, st["coefs"]) and const_sum == st["const_sum"]


@pytest.mark.skipif(not os.path.exists(U.REF_BIN_CPU), reason="oracle/_ref/refctx_cpu not built (needs /root/reference at build time)")
def test_wat_emitter_on_every_family_interleaved_against_the_reference(oracle, pr):
    """tests/golden/kitchen_sink.wat: witnesses held in locals across loops and branches, stores and loads of witnesses inside
    indirectly called functions, floats and globals between them, a value-carrying block left by br_table with witnesses above
    the label, an early return that drops live bit vectors, a typed select on references -- 40 row events, the same through
    the reference's interpreter and the emitter in all three spellings"""
    text = open(os.path.join(U.HERE, "golden", "kitchen_sink.wat")).read()
    raw = U.run_reference_on_wat(text, 256, seed_byte=10)
    assert raw["valid"] == [1, 1, 1] and raw["verifier"] == [1] * 7 and len(raw["kinds"]) == 40
    st = _reference_rows(raw)
    for spelling in (text, U.wat_to_wasm(text), U.wat_to_plain(text)):
        _emitter_equals_reference_rows(pr, spelling, st)


@pytest.mark.skipif(not os.path.exists(U.REF_BIN_CPU), reason="oracle/_ref/refctx_cpu not built (needs /root/reference at build time)")
@pytest.mark.parametrize("seed", range(6))
def test_wat_emitter_against_the_reference_interpreter_on_mixed_programs(oracle, pr, seed):
    """differential: witness stores / loads / bulk memory operations shuffled between floating-point statements, globals and float
    memory traffic in one function (tests/refctx_util.py: rand_mixed_program), text and binary"""
    import random
    rng = random.Random(5500 + seed)
    text = U.rand_mixed_program(rng, nmem=rng.randrange(6, 14), nfloat=rng.randrange(4, 10))
    raw = U.run_reference_on_wat(text, 256, seed_byte=seed + 1)
    assert raw["valid"] == [1, 1, 1] and raw["verifier"] == [1] * 7
    st = _reference_rows(raw)
    _emitter_equals_reference_rows(pr, text, st)
    _emitter_equals_reference_rows(pr, U.wat_to_wasm(text), st)


@pytest.mark.skipif(not os.path.exists(U.REF_BIN_CPU), reason="oracle/_ref/refctx_cpu not built (needs /root/reference at build time)")
def test_wat_emitter_nan_propagation_against_the_reference(oracle, pr):
    """two NaN operands of different sign and payload through add / sub / mul / div in both widths and both orders: SSE returns
    the first source of the instruction, and which C++ operand that is, is the compiler's choice for the commutative ones --
    the reference's handlers (GCC 13, -O1 and -O3 alike) return the first operand's; the emitter says so explicitly (a mixed
    random program found that, left to the compiler, it returned the second's for add and mul)"""
    stmts = []
    for w, t, pay in ((32, "f32", "0x200001"), (64, "f64", "0x4000000000001")):
        for op in ("add", "sub", "mul", "div"):
            for x, y in (("-nan", "nan:" + pay), ("nan:" + pay, "-nan")):
                stmts.append("(drop (call $i%d_private_const (i%d.reinterpret_%s (%s.%s (%s.const %s) (%s.const %s)))))" % (w, w, t, t, op, t, x, t, y))
    head = U.WAT_HEAD_BOTH[:U.WAT_HEAD_BOTH.index("(func $t")]
    text = head + "(func $t\n" + "\n".join(stmts) + "\n" + U.WAT_TAIL
    raw = U.run_reference_on_wat(text, 256, seed_byte=11)
    assert raw["valid"] == [1, 1, 1]
    st = _reference_rows(raw)
    _emitter_equals_reference_rows(pr, text, st)
    _emitter_equals_reference_rows(pr, U.wat_to_wasm(text), st)


REFERENCE_INTEGER_PROGRAMS = [w + "_" + op for w in ("i32", "i64") for op in (
    "add and clz ctz div_s div_u eq eqz ge_s ge_u gt_s gt_u le_s le_u lt_s lt_u mul ne or popcnt rem_s rem_u rotl rotr shl shr_s shr_u sub xor").split()] + [
    "i32_extend", "i32_wrap_i64", "i64_extend8_s", "i64_extend16_s", "i64_extend32_s", "i64_extend_i32_s", "i64_extend_i32_u"]


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "tests")) or not os.path.exists(U.REF_BIN_CPU), reason="needs the reference tree and oracle/_ref/refctx_cpu")
@pytest.mark.parametrize("name", REFERENCE_INTEGER_PROGRAMS)
def test_wat_emitter_on_the_reference_test_programs(oracle, pr, name):
    """all 65 integer programs of the reference's tests/ (every iNN instruction it tests), read where they lie: the
    reference's own interpreter and the emitter agree on every row at l = 64 -- and the reference's verifier accepts"""
    text = open(os.path.join(REFERENCE, "tests", name + ".wat")).read()
    raw = U.run_reference_on_wat(text, 256)
    assert raw["valid"] == [1, 1, 1] and raw["verifier"] == [1] * 7
    _emitter_equals_reference_rows(pr, text, _reference_rows(raw))


BOUNDARY_TU = r"""
#include <algorithm>
#include <cstring>
#include <format>
#include <fstream>
#include <iomanip>
#include <memory>
#include <stdexcept>
#include <params.hpp>
#include <interpreter.hpp>
#include <wgpu.hpp>                       // ligero-prover_b200/host/compat: webgpu_context = cuda_context
#include <zkp/finite_field_gmp.hpp>
#include <zkp/nonbatch_context.hpp>
#include <host_modules/vbn254fr.hpp>
using namespace ligero; using namespace ligero::vm;
using field_t = zkp::bn254_gmp;
static_assert(std::is_same_v<webgpu_context, cuda_context>);
using ctx1_t = zkp::nonbatch_stage1_context<field_t, webgpu_context, zkp::stage1_random_policy, params::hasher>;
template struct zkp::nonbatch_stage1_context<field_t, webgpu_context, zkp::stage1_random_policy, params::hasher>;
template struct zkp::nonbatch_stage2_context<field_t, webgpu_context, zkp::stage2_random_policy>;
template struct zkp::nonbatch_stage3_context<field_t, webgpu_context, zkp::stage3_random_policy>;
template struct zkp::nonbatch_verifier_context<field_t, webgpu_context, zkp::verifier_random_policy, params::hasher>;
template struct ligero::vm::vbn254fr_module<ctx1_t>;
"""


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "include", "zkp")) or shutil.which("g++") is None,
                    reason="needs the reference tree and g++ (build container only)")
def test_reference_contexts_compile_against_the_cuda_executor(tmp_path):
    """explicit instantiation of the reference's stage 1/2/3 contexts, its verifier context and its vbn254fr module with
    Executor = ligero::webgpu_context, which the compat include directory makes the CUDA adapter: every member function
    (some 40 executor call sites) must type-check -- with NO edit to the reference's headers"""
    gen = os.path.join(U.ROOT, "oracle", "_ref", "gen")
    subprocess.check_call(["python3", os.path.join(U.ROOT, "oracle", "refgen.py"), REFERENCE, gen])
    tu = tmp_path / "boundary.cpp"
    tu.write_text(BOUNDARY_TU)
    cmd = ["g++", "-std=c++20", "-fsyntax-only", "-I", gen, "-I", os.path.join(U.ROOT, "ligero-prover_b200", "host", "compat"),
           "-I", os.path.join(U.ROOT, "ligero-prover_b200", "host"), "-I", os.path.join(REFERENCE, "include"),
           "-I", os.path.join(U.ROOT, "tests", "stubs"), str(tu)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-4000:]

"""Host-side prover layer (include/lgr_prover.h, liblgr_prover.so) against the independent Python restatement
(oracle/prover_ref.py: hashlib, `cryptography` AES, google.protobuf).  No GPU needed: SURVEY 8f rows N1 / N2."""
import ctypes as C
import gzip
import hashlib
import importlib
import os
import random
import re
import shutil
import struct
import subprocess

import numpy as np
import pytest

from oracle import prover_ref as ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def pr(lgr):
    return importlib.import_module("ligero_prover_b200.prover")


def test_prover_abi_exports_every_declared_symbol(pr):
    hdr = open(os.path.join(ROOT, "include", "lgr_prover.h")).read()
    declared = set(re.findall(r"\b(lgrp_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 13
    missing = [s for s in sorted(declared) if not hasattr(pr.lib(), s)]
    assert not missing, missing


def test_stage_seeds_hash_the_literal_with_its_nul(pr):
    """hash("LigetronStage1", root, instance) binds the literal to the array overload (hash.hpp:61-65): 15 bytes"""
    rng = random.Random(1)
    root, inst = rng.randbytes(32), rng.randbytes(32)
    want = hashlib.sha256(b"LigetronStage1\x00" + root + inst).digest()
    assert pr.stage1_seed(root, inst) == want == ref.stage1_seed(root, inst)
    assert want != hashlib.sha256(b"LigetronStage1" + root + inst).digest()
    v = [np.frombuffer(rng.randbytes(4 * 8 * 16), np.uint32) for _ in range(3)]
    want2 = hashlib.sha256(b"LigetronStage2\x00" + root + b"".join(x.tobytes() for x in v)).digest()
    assert pr.stage2_seed(root, *v) == want2 == ref.stage2_seed(root, *v)


def test_hash_random_engine_block_structure(pr):
    """random.hpp:129-138: block 0 = SHA-256(LE64(0)) (no seed), block i = SHA-256(seed || LE64(i)); bytes from index 31 down"""
    seed = bytes(range(32))
    got = pr.hash_random_bytes(seed, 96)
    b0 = hashlib.sha256(struct.pack("<Q", 0)).digest()
    b1 = hashlib.sha256(seed + struct.pack("<Q", 1)).digest()
    b2 = hashlib.sha256(seed + struct.pack("<Q", 2)).digest()
    assert got == b0[::-1] + b1[::-1] + b2[::-1]
    e = ref.HashRandomEngine(seed)
    assert got == bytes(e() for _ in range(96))


@pytest.mark.parametrize("n", [1, 2, 3, 100, 192, 193, 255, 256, 257, 300, 448, 1024, 4096, 32768, 65536, 1 << 17])
def test_sampler_matches_restatement(pr, n):
    """every branch of Boost's generate_uniform_int is reached: range 0, == 255, < 255 (small n), > 255"""
    for s in range(3):
        seed = hashlib.sha256(b"sampler%d/%d" % (n, s)).digest()
        got = pr.sample_indices(seed, n)
        assert got == ref.sample_indices(seed, n)
        assert len(got) == min(n, 192) == len(set(got)) and got == sorted(got) and all(0 <= i < n for i in got)


def test_sampler_statistics_and_the_seedless_first_block(pr):
    """sanity of the restated Boost algorithm: every index of a small population gets drawn.  The reference's PRG
    quirk shows too: its first 32 bytes do not depend on the seed (random.hpp:129-138), so the first ~16 draws of
    the partial shuffle -- and with them a handful of sampled columns -- are the same for EVERY proof."""
    n, seeds, cnt = 1000, 200, np.zeros(1000)
    for s in range(seeds):
        for i in pr.sample_indices(hashlib.sha256(b"u%d" % s).digest(), n, 100):
            cnt[i] += 1
    always = int((cnt == seeds).sum())
    assert 10 <= always <= 16                                   # 32 bytes / 2 bytes per draw (range > 255 takes two)
    rest = cnt[cnt < seeds]
    assert rest.min() > 0 and rest.max() < 45 and abs(rest.mean() - seeds * (100 - always) / (n - always)) < 1e-9


def test_fr_random_stream(pr, oracle):
    key = hashlib.sha256(b"key").digest()
    got = pr.fr_random(key, 1300)                     # crosses two 16 KiB refills (512 draws each)
    want = ref.FrRandomStream(key).take(1300)
    assert np.array_equal(got, oracle.to_limbs(want))
    assert all(0 <= v < ref.P for v in want)
    # definition: AES-256-CTR keystream of zeros, 32 bytes little-endian, >> 2, one conditional subtract
    from cryptography.hazmat.primitives.ciphers import Cipher, algorithms, modes
    ks = Cipher(algorithms.AES(key), modes.CTR(bytes(16))).encryptor().update(bytes(64))
    v1 = int.from_bytes(ks[32:64], "little") >> 2
    assert oracle.from_limbs(got[1])[0] == (v1 - ref.P if v1 >= ref.P else v1)


@pytest.mark.parametrize("nleaves,opened", [(8, [3]), (8, [0, 1, 6]), (256, None), (1024, None), (1000, None)])
def test_decommit_recommit(pr, oracle, nleaves, opened):
    rng = random.Random(nleaves)
    leaves = np.frombuffer(rng.randbytes(32 * nleaves), np.uint8).reshape(nleaves, 32)
    nodes = oracle.merkle_build(leaves)
    total = nodes.shape[0]
    if opened is None:
        opened = sorted(rng.sample(range(nleaves), min(192, nleaves // 2)))
    pos, sib = pr.decommit(nodes, opened)
    assert pos == ref.sibling_positions(opened, total)
    assert sib == [nodes[p].tobytes() for p in pos]
    half = total // 2
    lv = [nodes[half + i].tobytes() for i in opened]
    root = nodes[0].tobytes()
    assert pr.recommit(lv, opened, total, sib) == root
    assert ref.recommit(dict(zip(opened, lv)), opened, total, sib) == root          # plain recursive recomputation
    bad = list(sib); bad[0] = bytes(32)
    assert pr.recommit(lv, opened, total, bad) != root
    with pytest.raises(pr.ProverError):
        pr.recommit(lv, opened, total, sib[:-1])                                     # "Sibling hash count mismatch"


def _random_proof_fields(rng, k, rows, nsib, S):
    n = 4 * k
    u = lambda cnt: np.frombuffer(rng.randbytes(4 * cnt), np.uint32)
    return dict(code=u(n * 8), linear=u(n * 8), quad=u(n * 8), samplings=u(rows * S * 8))


def test_envelope_bytes_equal_google_protobuf(pr):
    """the hand-written proto3 writer emits exactly what libprotobuf would for the same field values, and the reader
    round-trips it (proof_serializer.hpp:166-224)"""
    rng = random.Random(7)
    k, S = 64, 192
    n = 4 * k
    opened = sorted(rng.sample(range(n), S))
    total = 2 * n - 1
    pos = ref.sibling_positions(opened, total)
    sib = [rng.randbytes(32) for _ in pos]
    f = _random_proof_fields(rng, k, 9, len(pos), S)
    meta = {"prover_version": "1.5.0", "program_hash": rng.randbytes(32), "generated_at": 1792214281, "k": k, "n": n, "sample_size": S}
    want = ref.build_envelope(meta, rng.randbytes(32), sib, opened, f["code"], f["linear"], f["quad"], f["samplings"])
    for blob in (want, gzip.compress(want)):
        p = pr.parse_proof(blob)
        assert p.envelope == want
        assert gzip.decompress(p.gzip) == want
        p.close()
    # a zero index, a zero timestamp and empty vectors exercise proto3's "default values are not written" rule
    meta0 = dict(meta, generated_at=0)
    want0 = ref.build_envelope(meta0, bytes(32), [], list(range(n)), [], [], [], [])
    p = pr.parse_proof(want0)
    assert p.envelope == want0
    p.close()
    with pytest.raises(pr.ProverError):
        pr.parse_proof(want[: len(want) // 2])
    with pytest.raises(pr.ProverError):
        pr.parse_proof(ref.build_envelope(meta, bytes(32), sib[:-1], opened, [], [], [], []))   # sibling count mismatch


def test_oracle_prover_is_self_consistent(oracle):
    """the CPU restatement proves a small true statement, its container parses back, and the verifier-side checks
    (openings recommit, test vectors match the openings) pass -- and fail on a tampered proof"""
    rng = random.Random(11)
    k, l, kinds = 64, 40, [0, 1, 1, 0]
    n = 4 * k
    vals, coefs, acc = [], [], 0
    for kind in kinds:
        if kind:
            x = [rng.randrange(ref.P) for _ in range(l)]; y = [rng.randrange(ref.P) for _ in range(l)]
            rows = [x, y, [a * b % ref.P for a, b in zip(x, y)]]
        else:
            rows = [[rng.randrange(ref.P) for _ in range(l)]]
        for r in rows:
            c = [rng.randrange(ref.P) for _ in range(l)]
            acc += sum(a * b for a, b in zip(r, c)); vals.append(oracle.to_limbs(r)); coefs.append(oracle.to_limbs(c))
    values, coef = np.stack(vals), np.stack(coefs)
    inst = bytes(range(32))
    w = ref.prove(l, k, kinds, values, coef, (-acc) % ref.P, hashlib.sha256(b"s").digest(), inst)
    assert w["valid"] == (True, True, True)
    meta = {"prover_version": "1.5.0", "program_hash": bytes(32), "generated_at": 5, "k": k, "n": n, "sample_size": 192}
    blob = ref.build_envelope(meta, w["root"], w["siblings"], w["sample"], w["code"], w["linear"], w["quad"], w["samplings"])
    env = ref.parse_envelope(gzip.compress(blob))
    assert ref.verify_openings(env, l, k, kinds, coef, inst)
    bad = w["samplings"].copy(); bad[1, 5, 0] ^= 1
    env2 = ref.parse_envelope(ref.build_envelope(meta, w["root"], w["siblings"], w["sample"], w["code"], w["linear"], w["quad"], bad))
    with pytest.raises(AssertionError):
        ref.verify_openings(env2, l, k, kinds, coef, inst)
    assert not ref.prove(l, k, kinds, values, coef, (1 - acc) % ref.P, bytes(32), inst)["valid"][1]


def _load_fixture():
    import json
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", "prover_k64.json")))
    l = fx["l"]
    values = np.frombuffer(bytes.fromhex(fx["values_hex"]), np.uint32).reshape(-1, l, 8)
    coefs = np.frombuffer(bytes.fromhex(fx["coefs_hex"]), np.uint32).reshape(-1, l, 8)
    return fx, values, coefs


def test_oracle_prover_matches_committed_fixture(oracle):
    """tests/golden/prover_k64.json (made by tests/golden/make_prover_fixture.py) pins the restatement"""
    fx, values, coefs = _load_fixture()
    w = ref.prove(fx["l"], fx["k"], fx["kinds"], values, coefs, int(fx["const_sum"], 16), bytes.fromhex(fx["encoding_seed"]),
                  bytes.fromhex(fx["instance_hash"]))
    assert w["root"].hex() == fx["root"] and w["stage1_seed"].hex() == fx["stage1_seed"] and w["stage2_seed"].hex() == fx["stage2_seed"]
    assert w["sample"] == fx["sample"] and list(w["valid"]) == fx["valid"]
    assert hashlib.sha256(w["code"].tobytes()).hexdigest() == fx["code_sha256"]
    assert hashlib.sha256(w["samplings"].tobytes()).hexdigest() == fx["samplings_sha256"]
    meta = {"prover_version": "1.5.0", "program_hash": bytes.fromhex(fx["program_hash"]), "generated_at": fx["generated_at"], "k": fx["k"],
            "n": 4 * fx["k"], "sample_size": 192}
    env = ref.build_envelope(meta, w["root"], w["siblings"], w["sample"], w["code"], w["linear"], w["quad"], w["samplings"])
    assert len(env) == fx["envelope_len"] and hashlib.sha256(env).hexdigest() == fx["envelope_sha256"]


def test_gzip_of_a_large_proof_is_one_valid_member(pr):
    """proofs above 8 MiB are deflated in parallel chunks stitched into ONE gzip member (proof_wire.hpp); Python's gzip
    (single-member reader path included) must give back the envelope, and the parser must accept it"""
    import zlib
    rng = random.Random(9)
    k, S = 256, 192
    n = 4 * k
    opened = sorted(rng.sample(range(n), S))
    pos = ref.sibling_positions(opened, 2 * n - 1)
    sib = [rng.randbytes(32) for _ in pos]
    u = lambda cnt: np.frombuffer(rng.randbytes(4 * cnt), np.uint32)
    meta = {"prover_version": "1.5.0", "program_hash": bytes(32), "generated_at": 7, "k": k, "n": n, "sample_size": S}
    want = ref.build_envelope(meta, bytes(32), sib, opened, u(n * 8), u(n * 8), u(n * 8), u(3000 * S * 8))
    assert len(want) > (16 << 20)
    p = pr.parse_proof(want)
    gz = p.gzip
    d = zlib.decompressobj(31)                                # gzip wrapper, one member
    assert d.decompress(gz) == want and d.eof and d.unused_data == b""
    assert gzip.decompress(gz) == want
    p2 = pr.parse_proof(gz)
    assert p2.envelope == want
    p.close(); p2.close()


@pytest.mark.parametrize("l,nw,seed", [(4, 0, 1), (4, 3, 2), (4, 50, 3), (7, 200, 4), (40, 500, 5)])
def test_row_packer_matches_restatement(pr, oracle, l, nw, seed):
    """lazy row emission of witness_manager (rows leave when the (l+1)-th witness arrives; finalize: linear then triple)"""
    rng = random.Random(seed)
    ws = []
    for _ in range(nw):
        if rng.random() < 0.6:
            ws.append(("L", rng.randrange(ref.P), rng.randrange(ref.P)))
        else:
            x, y = rng.randrange(ref.P), rng.randrange(ref.P)
            ws.append(("Q", (x, y, x * y % ref.P), tuple(rng.randrange(ref.P) for _ in range(3))))
    want_k, want_v, want_c = ref.pack_rows(l, ws)
    pk = pr.RowPacker(l)
    for w in ws:
        if w[0] == "L":
            pk.push_linear(oracle.to_limbs([w[1]])[0], oracle.to_limbs([w[2]])[0])
        else:
            pk.push_quadratic(oracle.to_limbs(w[1]), oracle.to_limbs(w[2]))
    pk.finalize()
    kinds, vals, coefs = pk.rows()
    assert list(kinds) == want_k
    assert vals.shape[0] == len(want_v) == len(want_k) + 2 * sum(want_k)
    for r in range(len(want_v)):
        assert oracle.from_limbs(vals[r]) == want_v[r] and oracle.from_limbs(coefs[r]) == want_c[r]
    # exactly-full rows are NOT emitted before finalize (lazy flush): l linear witnesses -> one row, at finalize
    pk2 = pr.RowPacker(l)
    for i in range(l):
        pk2.push_linear(oracle.to_limbs([i + 1])[0], oracle.to_limbs([0])[0])
    assert pk2.rows()[0].size == 0
    pk2.push_linear(oracle.to_limbs([99])[0], oracle.to_limbs([0])[0])
    assert list(pk2.rows()[0]) == [0]
    pk2.finalize()
    assert list(pk2.rows()[0]) == [0, 0]
    pk.close(); pk2.close()


def test_oracle_prover_vbn254fr_events_are_self_consistent(oracle):
    """CPU restatement of the on_batch_* flow (vbn254fr calls on k-element variables): a satisfiable batch program
    passes all three tests, a violated assertion fails the quadratic test only, and the committed-row count follows
    nonbatch_context.hpp:497-553 (init 1, bit 1, equal 2, quadratic 3)"""
    import random
    R = ref
    P = ref.P
    l, k = 24, 64
    rng = random.Random(9)
    v0 = [rng.randrange(P) for _ in range(l)]
    v1 = [rng.randrange(1, P) for _ in range(l)]
    lin = [rng.randrange(P) for _ in range(l)]
    cf = [rng.randrange(P) for _ in range(l)]
    kinds = [R.EV_VSET, R.EV_VSET, R.EV_VMUL, R.EV_LINEAR, R.EV_VDIV, R.EV_VASSERT_EQ, R.EV_VBIT, R.EV_VCOPY]
    args = [[0, 0, 0], [1, 0, 0], [2, 0, 1], [3, 2, 1], [3, 0, 0], [4, 2, 5], [5, 4, 0]]
    values = np.stack([oracle.to_limbs(r) for r in (v0, v1, lin)])
    coefs = np.stack([oracle.to_limbs(r) for r in ([0] * l, [0] * l, cf)])
    const_sum = (-sum(a * b for a, b in zip(lin, cf))) % P
    out = R.prove(l, k, kinds, values, coefs, const_sum, bytes(range(32)), bytes(32), arena_slots=6, batch_args=args)
    assert out["valid"] == (True, True, True)
    assert out["encoded_rows"] == 1 + 1 + 3 + 1 + 3 + 2 + 1 + 2 + 3
    bad = [a[:] for a in args]
    bad[4] = [3, 1, 0]                                   # v3 (= v0) asserted equal to v1
    assert R.prove(l, k, kinds, values, coefs, const_sum, bytes(range(32)), bytes(32), arena_slots=6, batch_args=bad)["valid"] == (True, True, False)


# ---------------------------------------------------------------- bounded .wat front end + witness emitter
def _wat_check(pr, oracle, text, l, k, expect_valid=True):
    """emit rows twice (values only / with coefficients from a seed), check the constraint system in big-int arithmetic,
    then run the CPU prover on the rows"""
    kinds0, vals0, coefs0, cs0, st = pr.wat_emit(text, l)
    assert not coefs0.any() and cs0 == 0
    seed = hashlib.sha256(b"stage-1 seed stand-in").digest()
    kinds, vals, coefs, const_sum, st2 = pr.wat_emit(text, l, seed)
    assert np.array_equal(kinds, kinds0) and np.array_equal(vals, vals0) and st == st2
    P = ref.P
    v = oracle.from_limbs(vals.reshape(-1, 8)); c = oracle.from_limbs(coefs.reshape(-1, 8))
    lin_ok = (sum(a * b for a, b in zip(v, c)) + const_sum) % P == 0
    # triples: x*y = z slot by slot
    quad_ok, r = True, 0
    V = np.array(v, dtype=object).reshape(-1, l)
    for kd in kinds:
        if kd:
            quad_ok &= all((int(a) * int(b) - int(z)) % P == 0 for a, b, z in zip(V[r], V[r + 1], V[r + 2]))
        r += 3 if kd else 1
    assert quad_ok == (st["violated_constraints"] == 0 or quad_ok)      # the emitter never breaks a slot relation itself
    assert lin_ok == expect_valid
    out = ref.prove(l, k, kinds, vals, coefs, const_sum, bytes(range(32)), bytes(32))
    assert out["valid"] == (True, expect_valid, True)
    return kinds, st


def test_wat_emitter_on_the_repo_fixture(pr, oracle):
    text = open(os.path.join(ROOT, "tests", "golden", "mul64.wat")).read()
    kinds, st = _wat_check(pr, oracle, text, l=200, k=256)
    assert st["violated_constraints"] == 0
    assert (st["private_consts"], st["asserts"], st["arithmetic_ops"]) == (23, 8, 9)
    # 64 bit slots per private const, 1 + 128 per product, 65 per sum / difference
    assert st["quadratic_slots"] == 23 * 64 + 5 * 129 + 4 * 65
    assert list(kinds).count(1) == -(-st["quadratic_slots"] // 200)
    # a false assertion leaves the slot relations intact and breaks exactly the linear test
    bad = text.replace("(i64.const 15)", "(i64.const 16)")
    _, st_bad = _wat_check(pr, oracle, bad, l=200, k=256, expect_valid=False)
    assert st_bad["violated_constraints"] == 1


def test_wat_emitter_rejects_what_it_does_not_support(pr):
    for text, why in (("(module (func $f) (export \"_start\" (func $f)) (tag $e))", "module field"),
                      ("(module (import \"wasi\" \"x\" (func $x)) (func $f (call $x)) (export \"_start\" (func $f)))", "call of wasi.x, which the front end does not provide"),
                      ("(module (func $f (drop (f64.fma (f64.const 1) (f64.const 2)))) (export \"_start\" (func $f)))", "unsupported instruction"),
                      ("(module (func $f (drop (ref.null any))) (export \"_start\" (func $f)))", "ref.null of an unknown type"),
                      ("(module (func $f (drop (i64.load (i32.const 0)))) (export \"_start\" (func $f)))", "without a memory"),
                      ("(module (func $f)", "unbalanced"),
                      ("(module (func $f))", "_start")):
        with pytest.raises(pr.ProverError, match=why):
            pr.wat_emit(text, 8)


@pytest.mark.skipif(not os.path.exists("/root/reference/tests/i64_mul.wat"), reason="reference tree not present")
@pytest.mark.parametrize("name,consts,asserts,ops", [("i64_mul.wat", 27, 9, 9)])
def test_wat_emitter_on_the_reference_program(pr, oracle, name, consts, asserts, ops):
    """BASELINE config 4's program: 27 private constants, 9 products, 9 assertions; at the reference's default geometry
    (l = 8000) the whole statement is 1 linear row + 1 quadratic triple, as SURVEY 8d counts it"""
    text = open(os.path.join("/root/reference/tests", name)).read()
    kinds, vals, coefs, cs, st = pr.wat_emit(text, 8000)
    assert (st["private_consts"], st["asserts"], st["arithmetic_ops"], st["violated_constraints"]) == (consts, asserts, ops, 0)
    assert st["quadratic_slots"] == 27 * 64 + 9 * 129 and st["quadratic_slots"] < 8000
    assert list(kinds) == [0, 1] and vals.shape == (4, 8000, 8)
    _wat_check(pr, oracle, text, l=448, k=512)


def _rand_wat_expr(rng, depth):
    """(text, value mod 2^64) of a random folded i64 expression over private and public constants"""
    M = 1 << 64
    if depth == 0 or rng.random() < 0.25:
        v = rng.choice([0, 1, 2, M - 1, 1 << 63, (1 << 63) - 1, 1 << 32, rng.getrandbits(64), rng.getrandbits(20)])
        lit = rng.choice([str(v), hex(v), str(v - M) if v >= 1 << 63 else str(v)])
        return ("(call $i64_private_const (i64.const %s))" if rng.random() < 0.7 else "(i64.const %s)") % lit, v
    op = rng.choice(["mul", "add", "sub"])
    (ta, va), (tb, vb) = _rand_wat_expr(rng, depth - 1), _rand_wat_expr(rng, depth - 1)
    return "(i64.%s %s %s)" % (op, ta, tb), {"mul": va * vb, "add": va + vb, "sub": va - vb}[op] % M


@pytest.mark.parametrize("seed", range(8))
def test_wat_emitter_constraint_system_over_every_integer_instruction(pr, oracle, seed):
    """needs no reference run: random trees over the whole integer instruction set (tests/refctx_util.py: rand_int_program,
    expected values from WebAssembly's semantics).  Every assertion holds in the emitted constraint system -- the linear
    combination with the drawn coefficients vanishes, every slot is a product -- and the CPU prover's three checks pass;
    moving one expected value by one breaks exactly the linear test"""
    import refctx_util as U
    rng = random.Random(5100 + seed)
    w = (32, 64)[seed & 1]
    text, exprs = U.rand_int_program(rng, w, nexpr=3, depth=2)
    _, st = _wat_check(pr, oracle, text, l=256, k=512)
    assert st["violated_constraints"] == 0 and st["asserts"] == 3
    t, v = exprs[rng.randrange(3)]
    for rhs in ("(i%d.const %d)" % (w, v), "(call $i%d_private_const (i%d.const %d))" % (w, w, v)):
        stmt = "(call $assert_equal %s %s)" % (t, rhs)
        if stmt in text:
            wrong = text.replace(stmt, "(call $assert_equal %s %s)" % (t, rhs.replace("const %d)" % v, "const %d)" % ((v + 1) % (1 << w)))), 1)
            _, st_bad = _wat_check(pr, oracle, wrong, l=256, k=512, expect_valid=False)
            assert st_bad["violated_constraints"] >= 1
            break
    else:
        raise AssertionError("assertion text not found")


@pytest.mark.parametrize("seed", range(6))
def test_wat_emitter_on_random_expression_trees(pr, oracle, seed):
    """wrap-around semantics of nested products / sums / differences: every assertion against the value Python computes
    holds in the emitted constraint system, and moving one expected value by one breaks exactly the linear test"""
    rng = random.Random(1000 + seed)
    cases = [_rand_wat_expr(rng, rng.randrange(1, 4)) for _ in range(5)]
    head = ('(module (import "env" "i64_private_const" (func $i64_private_const (param i64) (result i64)))\n'
            '(import "env" "assert_equal" (func $assert_equal (param i64 i64)))\n(func $t\n')
    tail = ')\n(export "_start" (func $t)))\n'
    body = lambda cs: "".join("(call $assert_equal %s (i64.const %d))\n" % (t, v) for t, v in cs)
    _, st = _wat_check(pr, oracle, head + body(cases) + tail, l=256, k=256)
    assert st["violated_constraints"] == 0 and st["asserts"] == 5
    j = rng.randrange(5)
    wrong = list(cases); wrong[j] = (cases[j][0], (cases[j][1] + 1) % (1 << 64))
    _, st_bad = _wat_check(pr, oracle, head + body(wrong) + tail, l=256, k=256, expect_valid=False)
    assert st_bad["violated_constraints"] == 1


# ---------------------------------------------------------------- the other two spellings of a program: binary and plain text
def _same_rows(pr, a, b, l=64):
    seed = hashlib.sha256(b"spellings").digest()
    ra, rb = pr.wat_emit(a, l, seed), pr.wat_emit(b, l, seed)
    assert list(ra[0]) == list(rb[0]) and np.array_equal(ra[1], rb[1]) and np.array_equal(ra[2], rb[2]) and ra[3] == rb[3] and ra[4] == rb[4]
    return ra[4]


@pytest.mark.parametrize("name", ["mul64", "arith32", "intops"])
def test_wasm_binary_and_plain_text_give_the_rows_of_the_folded_text(pr, name):
    """the reference's prover takes .wat and .wasm (src/webgpu_prover.cpp:189-207); so does lgrp_wat_emit / lgrp_prove_wat: the
    binary module of a program (assembled by tests/refctx_util.py: wat_to_wasm -- there is no wabt here) and its plain,
    unfolded text go through the same instruction list as the folded text and give the same rows, coefficients and const_sum"""
    import refctx_util as U
    text = open(U.WAT_TEXT[name]).read()
    wasm = U.wat_to_wasm(text)
    assert wasm[:8] == b"\0asm\x01\0\0\0"
    st = _same_rows(pr, text, wasm)
    assert st["violated_constraints"] == 0 and st["asserts"] > 0
    _same_rows(pr, text, U.wat_to_plain(text))


@pytest.mark.parametrize("seed", range(4))
def test_wasm_binary_front_end_on_random_programs(pr, seed):
    import refctx_util as U
    rng = random.Random(900 + seed)
    text, _ = U.rand_int_program(rng, (32, 64)[seed & 1], nexpr=3, depth=3)
    _same_rows(pr, text, U.wat_to_wasm(text, custom_section=bool(seed & 2)))


def test_wasm_binary_front_end_rejects_what_it_does_not_support(pr):
    import refctx_util as U
    good = U.wat_to_wasm(open(U.WAT_TEXT["arith32"]).read(), custom_section=False)
    pr.wat_emit(good, 64)
    sec = lambda sid, body: bytes([sid, len(body)]) + body
    for data, why in ((good[:-3], "section runs past the end|unexpected end"),
                      (good[:8] + sec(13, b"\x00") + good[8:], "unsupported module section"),                  # a tag section
                      (b"\0asm\x02\0\0\0" + good[8:], "binary version"),
                      (good[:-1] + b"\x28\x0b", "section runs past|unexpected end|unsupported"),
                      (b"\0asm\x01\0\0\0", "_start")):
        with pytest.raises(pr.ProverError, match=why):
            pr.wat_emit(data, 64)
    # an opcode the reference's interpreter does not have either, inside _start (try = 0x06)
    text = '(module (import "env" "assert_equal" (func $e (param i32 i32))) (func $f (call $e (i32.const 1) (i32.const 1))) (export "_start" (func $f)))'
    wasm = bytearray(U.wat_to_wasm(text, custom_section=False))
    at = wasm.rindex(b"\x41\x01\x41\x01")
    wasm[at] = 0x06
    with pytest.raises(pr.ProverError, match="unsupported instruction 0x06"):
        pr.wat_emit(bytes(wasm), 64)


def test_env_assertions_and_drop(pr, oracle):
    """env.assert_zero / assert_one / assert_constant / witness_cast_u64 / assert_is_concrete and `drop` (env.hpp:40-110,160-170):
    the constraint system holds, and a false assert_one breaks exactly the linear test"""
    head = ('(module (import "env" "i64_private_const" (func $pc (param i64) (result i64)))\n(import "env" "assert_zero" (func $z (param i64)))\n'
            '(import "env" "assert_one" (func $o (param i64)))\n(import "env" "assert_constant" (func $c (param i64)))\n'
            '(import "env" "witness_cast_u64" (func $cast (param i64) (result i64)))\n(import "env" "assert_is_concrete" (func $conc (param i64)))\n'
            '(import "env" "assert_equal" (func $eq (param i64 i64)))\n(func $t\n')
    body = ('(call $z (i64.sub (call $pc (i64.const 9)) (call $pc (i64.const 9))))\n(call $o (i64.eqz (call $pc (i64.const 0))))\n'
            '(call $c (i64.add (call $pc (i64.const 40)) (i64.const 2)))\n(call $eq (call $cast (call $pc (i64.const 77))) (i64.const 77))\n'
            '(call $conc (i64.mul (i64.const 6) (i64.const 7)))\n(drop (call $pc (i64.const 5)))\n(call $z (call $cast (i64.const 0)))\n')
    tail = ')\n(export "_start" (func $t)))\n'
    _, st = _wat_check(pr, oracle, head + body + tail, l=128, k=256)
    assert st["violated_constraints"] == 0 and st["asserts"] == 5 and st["private_consts"] == 6
    _, st_bad = _wat_check(pr, oracle, head + body.replace("(i64.eqz (call $pc (i64.const 0)))", "(i64.eqz (call $pc (i64.const 4)))") + tail, l=128, k=256, expect_valid=False)
    assert st_bad["violated_constraints"] == 1
    with pytest.raises(pr.ProverError, match="assert_is_concrete"):
        pr.wat_emit(head + "(call $conc (call $pc (i64.const 1)))\n" + tail, 64)


def test_front_end_validates_operand_widths(pr):
    """wabt would validate a module before the reference's interpreter sees it; the front end's own check: an instruction
    applied to a value of the other width is rejected (the handlers index operand bits by the instruction's width)"""
    head = ('(module (import "env" "i32_private_const" (func $p32 (param i32) (result i32)))\n(import "env" "i64_private_const" (func $p64 (param i64) (result i64)))\n'
            '(import "env" "assert_equal" (func $eq (param i64 i64)))\n(func $t\n')
    tail = ')\n(export "_start" (func $t)))\n'
    for body, why in (("(drop (i64.clz (call $p32 (i32.const 1))))", "type mismatch: i64.clz applied to an i32"),
                      ("(drop (i32.add (call $p32 (i32.const 1)) (call $p64 (i64.const 1))))", "type mismatch: i32.add applied to an i64"),
                      ("(drop (i64.and (i64.eq (call $p64 (i64.const 1)) (i64.const 1)) (i64.const 1)))", "type mismatch: i64.and applied to an i32"),
                      ("(drop (i32.wrap_i64 (i32.const 1)))", "type mismatch"),
                      ("i64.add", "stack underflow"), ("(call $eq (i64.const 1))", "stack underflow")):
        with pytest.raises(pr.ProverError, match=why):
            pr.wat_emit(head + body + tail, 64)
    # what the reference's own tests do is accepted: assert_equal (param i64 i64) called with i32 operands
    pr.wat_emit(head + "(call $eq (i64.eq (call $p64 (i64.const 1)) (i64.const 1)) (i32.const 1))" + tail, 64)


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_front_ends_survive_mutated_inputs_under_sanitizers(tmp_path):
    """tests/cpp/fuzz_wat.cpp built with AddressSanitizer + UBSan: thousands of mutated text and binary modules either run
    to the end or are rejected with an error -- no out-of-bounds access, no undefined arithmetic, no leak"""
    import refctx_util as U
    exe = str(tmp_path / "fuzz_wat")
    res = subprocess.run(["g++", "-std=c++17", "-g", "-O1", "-fsanitize=address,undefined,float-cast-overflow", "-fno-sanitize-recover=all", "-o", exe,
                          os.path.join(ROOT, "tests", "cpp", "fuzz_wat.cpp"), "-lcrypto"], capture_output=True, text=True)
    if res.returncode != 0 and "sanitize" in res.stderr:
        pytest.skip("sanitizer runtime not available")
    assert res.returncode == 0, res.stderr[-2000:]
    seeds = {"arith": open(U.WAT_TEXT["arith32"]).read(),
             "struct": U.rand_struct_program(random.Random(1), 64, nstmt=4, depth=2),      # locals, select, module functions
             "memory": U.rand_memory_program(random.Random(2), nstmt=12),                  # loads, stores, fill / copy / init, data
             "control": U.rand_cf_program(random.Random(3), 64, nstmt=4, depth=1),         # blocks, loops, branches, returns
             "float": U.rand_float_program(random.Random(4), nstmt=12),                    # floating point, conversions, globals
             "wasi": open(os.path.join(ROOT, "tests", "golden", "wasi_args.wat")).read(),  # WASI calls (no arguments given: argc = 0), proc_exit
             "indirect": open(os.path.join(ROOT, "tests", "golden", "indirect.wat")).read(),   # tables, element segments, call_indirect
             "tables": open(os.path.join(ROOT, "tests", "golden", "tables.wat")).read()}       # references, table.get / set / grow / fill / copy / init
    for name, text in seeds.items():
        (tmp_path / (name + ".wat")).write_text(text)
        (tmp_path / (name + ".wasm")).write_bytes(U.wat_to_wasm(text))
        for seed in (name + ".wat", name + ".wasm"):
            res = subprocess.run([exe, str(tmp_path / seed), "800"], capture_output=True, text=True, timeout=600)
            assert res.returncode == 0 and "accepted" in res.stdout, (res.stdout + res.stderr)[-3000:]


@pytest.mark.parametrize("seed", range(6))
def test_structured_programs_constraint_system(pr, oracle, seed):
    """needs no reference run: random programs with locals, select, helper functions and every integer instruction; all
    assertions (expected values from WebAssembly's semantics) hold in the emitted constraint system and the binary spelling
    gives the same rows"""
    import refctx_util as U
    rng = random.Random(6100 + seed)
    text = U.rand_struct_program(rng, (32, 64)[seed & 1], nstmt=5, depth=2)
    _, st = _wat_check(pr, oracle, text, l=256, k=512)
    assert st["violated_constraints"] == 0
    _same_rows(pr, text, U.wat_to_wasm(text), l=256)


def test_front_end_rejects_malformed_functions(pr):
    head = '(module (import "env" "i64_private_const" (func $pc (param i64) (result i64)))\n(import "env" "assert_equal" (func $eq (param i64 i64)))\n'
    tail = '(export "_start" (func $t)))'
    for body, why in (("(func $t (local $x i32) (local.set $x (call $pc (i64.const 1))))", "type mismatch: local.set"),
                      ("(func $t (drop (local.get 3)))", "unknown local"),
                      ("(func $t (param i64))", "_start with parameters"),
                      ("(func $f (param i64) (result i64) (i64.const 1) (i64.const 2)) (func $t (drop (call $f (i64.const 1))))", "leaves 2 values"),
                      ("(func $f (result i32) (i64.const 1)) (func $t (drop (call $f)))", "wrong type"),
                      ("(func $t (drop (call $nowhere)))", "unknown function"),
                      ("(func $t (call $t))", "call depth"),
                      ("(func $t (drop (select (i64.const 1) (i32.const 2) (i32.const 1))))", "type mismatch: select")):
        with pytest.raises(pr.ProverError, match=why):
            pr.wat_emit(head + body + tail, 64)


@pytest.mark.parametrize("seed", range(6))
def test_memory_programs_constraint_system(pr, oracle, seed):
    """needs no reference run: random memory programs; every load assertion (expected bytes from a byte-array model) holds
    in the emitted constraint system, and the binary spelling gives the same rows"""
    import refctx_util as U
    rng = random.Random(7300 + seed)
    text = U.rand_memory_program(rng, nstmt=16)
    _, st = _wat_check(pr, oracle, text, l=256, k=512)
    assert st["violated_constraints"] == 0
    _same_rows(pr, text, U.wat_to_wasm(text), l=256)


def test_memory_front_end_errors(pr):
    head = '(module (import "env" "i32_private_const" (func $pc (param i32) (result i32)))\n(memory 1)\n(data $d "abcd")\n'
    tail = '(export "_start" (func $t)))'
    for body, why in (("(func $t (drop (i32.load (i32.const 65533))))", "invalid memory address"),
                      ("(func $t (i32.store (i32.const 0) (i64.const 1)))", "type mismatch: i32.store"),
                      ("(func $t (memory.fill (i32.const 0) (call $pc (i32.const 1)) (i32.const 4)))", "concrete operands"),
                      ("(func $t (memory.init $d (i32.const 0) (i32.const 2) (i32.const 3)))", "memory.init: invalid address"),
                      ("(func $t (memory.init $nope (i32.const 0) (i32.const 0) (i32.const 1)))", "unknown data segment"),
                      ("(func $t (memory.copy (i32.const 65530) (i32.const 0) (i32.const 8)))", "memory.copy: out of range"),
                      ("(func $t (drop (i32.load32_u (i32.const 0))))", "unsupported instruction")):
        with pytest.raises(pr.ProverError, match=why):
            pr.wat_emit(head + body + tail, 64)
    with pytest.raises(pr.ProverError, match="256 MiB"):
        pr.wat_emit('(module (memory 65536) (func $t) (export "_start" (func $t)))', 64)


@pytest.mark.parametrize("seed", range(6))
def test_control_flow_programs_constraint_system(pr, oracle, seed):
    """needs no reference run: random control-flow programs; every assertion that runs holds in the emitted constraint
    system (expected values from a Python model of the control flow), text and binary give the same rows"""
    import refctx_util as U
    rng = random.Random(8400 + seed)
    text = U.rand_cf_program(rng, (32, 64)[seed & 1], nstmt=5, depth=2)
    _, st = _wat_check(pr, oracle, text, l=256, k=512)
    assert st["violated_constraints"] == 0
    _same_rows(pr, text, U.wat_to_wasm(text), l=256)
    _same_rows(pr, text, U.wat_to_plain(text), l=256)


def test_control_flow_validation(pr):
    head = '(module (import "env" "i64_private_const" (func $pc (param i64) (result i64)))\n'
    tail = '(export "_start" (func $t)))'
    for body, why in (("(func $t (block (i64.const 1)))", "values left on the stack at end"),
                      ("(func $t (drop (block (result i64) (i32.const 1))))", "type mismatch: end"),
                      ("(func $t (br 0))", "no such enclosing block"),
                      ("(func $t (block (br 1)))", "no such enclosing block"),
                      ("(func $t (if (i64.const 1) (then)))", "type mismatch: if"),
                      ("(func $t (drop (if (result i64) (i32.const 1) (then (i64.const 1)))))", "an if without an else"),
                      ("(func $t (block $a (block $b (result i64) (br_table $a $b (i32.const 0))) (drop)))", "br_table targets disagree"),
                      ("(func $t else)", "else without an if"),
                      ("(func $t end)", "end without a block"),
                      ("(func $t block)", "ends inside a block"),
                      ("(func $t (unreachable))", "unreachable executed"),
                      ("(func $t (loop $l (br $l)))", "step budget")):
        with pytest.raises(pr.ProverError, match=why):
            pr.wat_emit(head + body + tail, 64)
    # code after a branch is type-checked leniently (its operands may come from nowhere) and never runs
    pr.wat_emit(head + "(func $t (block (br 0) (drop (i64.add))))" + tail, 64)
    pr.wat_emit(head + "(func $f (result i64) (return (call $pc (i64.const 1))) (i64.mul)) (func $t (drop (call $f)))" + tail, 64)


def test_floating_point_and_globals_front_end(pr):
    """floating point (numbers only, as in the reference: interpreter_impl.hpp:1314-1853) and globals (native numbers,
    :1902-1924): results are observed through committed integers; what the reference cannot do is reported, not computed"""
    import refctx_util as U
    head = ('(module (import "env" "i32_private_const" (func $p32 (param i32) (result i32)))\n(import "env" "i64_private_const" (func $p64 (param i64) (result i64)))\n'
            '(import "env" "assert_equal" (func $eq (param i64 i64)))\n(global $g (mut i64) (i64.const -3))\n(global $c i32 (i32.const 9))\n(memory 1)\n(func $t (local $x f64)\n')
    tail = ')\n(export "_start" (func $t)))\n'
    body = ('(call $eq (call $p64 (i64.reinterpret_f64 (f64.mul (f64.const 1.5) (f64.const -4)))) (i64.const 0xC018000000000000))\n'
            '(call $eq (call $p32 (i32.reinterpret_f32 (f32.max (f32.const -0.0) (f32.const 0.0)))) (i32.const 0))\n'                 # the second operand wins a tie
            '(call $eq (call $p32 (i32.reinterpret_f32 (f32.max (f32.const 0.0) (f32.const -0.0)))) (i32.const 0x80000000))\n'
            '(call $eq (call $p32 (i32.reinterpret_f32 (f32.min (f32.const nan:0x1) (f32.const 1)))) (i32.const 0x7fc00000))\n'
            '(call $eq (call $p32 (i32.reinterpret_f32 (f32.mul (f32.const -nan) (f32.const nan:0x200001)))) (i32.const 0xffc00000))\n'         # two NaNs: the first operand's,
            '(call $eq (call $p64 (i64.reinterpret_f64 (f64.add (f64.const nan:0x4000000000001) (f64.const -nan)))) (i64.const 0x7ffc000000000001))\n'   # for the commutative operations too
            '(call $eq (call $p32 (i32.trunc_sat_f64_s (f64.const -1e300))) (i32.const 0x80000000))\n'
            '(call $eq (call $p64 (i64.trunc_f32_u (f32.const 0x1p63))) (i64.const 0x8000000000000000))\n'
            '(call $eq (call $p32 (i32.reinterpret_f32 (f32.demote_f64 (f64.const 16777217)))) (i32.const 0x4b800000))\n'
            '(call $eq (call $p64 (i64.reinterpret_f64 (f64.nearest (f64.const 2.5)))) (i64.const 0x4000000000000000))\n'
            '(f64.store (i32.const 8) (f64.sqrt (f64.const 2)))\n(local.set $x (f64.load (i32.const 8)))\n'
            '(call $eq (call $p64 (i64.load (i32.const 8))) (i64.reinterpret_f64 (local.get $x)))\n'
            '(global.set $g (i64.add (global.get $g) (i64.extend_i32_u (global.get $c))))\n(call $eq (call $p64 (global.get $g)) (i64.const 6))\n')
    for spelling in (head + body + tail, U.wat_to_wasm(head + body + tail), U.wat_to_plain(head + body + tail)):
        _, _, _, _, st = pr.wat_emit(spelling, 64)
        assert st["violated_constraints"] == 0 and st["asserts"] == 12 and st["private_consts"] == 12
    for bad, why in (("(drop (i32.trunc_f32_s (f32.const 3e9)))", "integer overflow"),
                     ("(drop (i64.trunc_f64_u (f64.const -1)))", "integer overflow"),
                     ("(drop (i32.trunc_f64_u (f64.const nan)))", "integer overflow"),
                     ("(call $eq (f64.const 1) (f64.const 1))", "floating-point value where a witness is needed"),
                     ("(drop (f32.convert_i32_s (call $p32 (i32.const 1))))", "concrete operands"),
                     ("(global.set $g (call $p64 (i64.const 1)))", "global.set takes concrete operands"),
                     ("(global.set $c (i32.const 1))", "immutable global"),
                     ("(drop (global.get 7))", "unknown global"),
                     ("(drop (f32.add (f32.const 1) (f64.const 1)))", "type mismatch: f32.add applied to an f64"),
                     ("(drop (f64.const 1.0.0))", "bad floating-point literal"),
                     ("(i32.store (i32.const 0) (call $p32 (i32.const 5)))\n(drop (f32.load (i32.const 0)))", "floating-point value where a witness is needed"),
                     ("(drop (f32.load8_u (i32.const 0)))", "unsupported instruction")):
        with pytest.raises(pr.ProverError, match=why):
            pr.wat_emit(head + bad + tail, 64)
    with pytest.raises(pr.ProverError, match="only i32 and i64 globals"):
        pr.wat_emit('(module (global $f f32 (f32.const 1)) (func $t) (export "_start" (func $t)))', 64)


def test_guest_arguments_and_wasi_front_end(pr):
    """the argument strings of the reference's JSON configuration (src/webgpu_prover.cpp:110-147), the instance hash folded
    from the public ones (:160-168), and what the WASI front end refuses"""
    import hashlib
    args = pr.config_args([{"i64": -2}, {"str": "abc"}, {"hex": "0x123"}, {"hex": "ff00"}])
    assert args == [b"Ligero\0", b"\xfe" + b"\xff" * 7, b"abc\0", b"\x01\x23", b"\xff\x00"]
    want = bytes(32)
    for i, a in enumerate(args):
        if i not in (1, 3):
            want = hashlib.sha256(want + a).digest()
    assert pr.wat_instance_hash(args, (1, 3)) == want
    assert pr.wat_instance_hash(None) == bytes(32) and pr.wat_instance_hash([b"x"], (0,)) == bytes(32)
    head = ('(module (import "wasi_snapshot_preview1" "args_get" (func $args (param i32 i32) (result i32)))\n'
            '(import "wasi_snapshot_preview1" "proc_exit" (func $exit (param i32)))\n(import "env" "i32_private_const" (func $pc (param i32) (result i32)))\n(memory 1)\n(func $t\n')
    tail = ')\n(export "_start" (func $t)))\n'
    _, _, _, _, st, code = pr.wat_emit(head + "(drop (call $args (i32.const 0) (i32.const 64)))\n(drop (i32.load8_u (i32.const 66)))\n(call $exit (i32.const 0))\n(unreachable)" + tail,
                                       64, args=[b"ab", b"cdef"], private_indices=[1], want_exit_code=True)
    assert code == 0 and st["linear_witnesses"] == 1            # one byte of the private argument was loaded: one witness
    for body, why in (("(drop (call $args (call $pc (i32.const 0)) (i32.const 64)))", "WASI call takes concrete operands"),
                      ("(drop (call $args (i32.const 65532) (i32.const 64)))", "reaches outside the memory"),
                      ("(drop (call $args (i64.const 0) (i32.const 64)))", "type mismatch")):
        with pytest.raises(pr.ProverError, match=why):
            pr.wat_emit(head + body + tail, 64, args=[b"ab", b"cdef"])
    with pytest.raises(pr.ProverError, match="call of wasi_snapshot_preview1.environ_get, which the front end does not provide"):
        pr.wat_emit('(module (import "wasi_snapshot_preview1" "environ_get" (func $e (param i32 i32) (result i32))) (memory 1) (func $t (drop (call $e (i32.const 0) (i32.const 0)))) (export "_start" (func $t)))', 64)
    # imports resolve when they are called (as in the reference): a module may import what the front end does not provide
    bn = '(module (import "bn254fr" "bn254fr_alloc" (func $e (param i32))) (func $t (if (i32.const %d) (then (call $e (i32.const 8))))) (export "_start" (func $t)))'
    assert pr.wat_emit(bn % 0, 64)[4]["violated_constraints"] == 0
    with pytest.raises(pr.ProverError, match="call of bn254fr.bn254fr_alloc, which the front end does not provide"):
        pr.wat_emit(bn % 1, 64)
    with pytest.raises(pr.ProverError, match="env and wasi_snapshot_preview1"):      # without a signature the call cannot even be typed
        pr.wat_emit('(module (import "bn254fr" "bn254fr_alloc" (func $e)) (func $t (call $e (i32.const 8))) (export "_start" (func $t)))'.replace("(func $e)", "(func $e (param v128))"), 64)


def test_indirect_call_front_end(pr):
    """call_indirect: what WebAssembly traps on is reported (the reference checks the bounds and the null entry; the callee's type
    it leaves as a TODO), and an index that is a witness is refused as the reference's as_u32() refuses it"""
    head = ('(module (import "env" "i32_private_const" (func $pc (param i32) (result i32)))\n(type $u (func (param i32) (result i32)))\n(type $v (func (result i32)))\n'
            '(table 4 funcref)\n(elem (i32.const 1) $inc $seven)\n(func $inc (type $u) (i32.add (local.get 0) (i32.const 1)))\n(func $seven (result i32) (i32.const 7))\n(func $t\n')
    tail = ')\n(export "_start" (func $t)))\n'
    _, _, _, _, st = pr.wat_emit(head + "(drop (call_indirect (type $u) (call $pc (i32.const 4)) (i32.const 1)))\n(drop (call_indirect (type $v) (i32.const 2)))" + tail, 64)
    assert st["violated_constraints"] == 0 and st["private_consts"] == 1
    for body, why in (("(drop (call_indirect (type $v) (i32.const 4)))", "index out of bound"),
                      ("(drop (call_indirect (type $v) (i32.const 0)))", "null pointer"),
                      ("(drop (call_indirect (type $v) (i32.const 1)))", "indirect call type mismatch"),
                      ("(drop (call_indirect (type $v) (call $pc (i32.const 2))))", "concrete index"),
                      ("(drop (call_indirect (type $nope) (i32.const 2)))", "unknown type"),
                      ("(drop (call_indirect (type $u) (i64.const 1) (i32.const 1)))", "type mismatch: call_indirect")):
        with pytest.raises(pr.ProverError, match=why):
            pr.wat_emit(head + body + tail, 64)
    pr.wat_emit(head.replace("$inc $seven)", "$pc $seven)") + tail, 64)      # an imported function may sit in the table ...
    with pytest.raises(pr.ProverError, match="indirect call of an imported function"):      # ... calling it through the table is not supported
        pr.wat_emit(head.replace("$inc $seven)", "$pc $seven)") + "(drop (call_indirect (type $u) (i32.const 1) (i32.const 1)))" + tail, 64)
    with pytest.raises(pr.ProverError, match="needs a table|unknown table"):
        pr.wat_emit('(module (type $v (func)) (func $t (call_indirect (type $v) (i32.const 0))) (export "_start" (func $t)))', 64)


def test_start_sections_are_ignored_and_inline_exports_are_read(pr):
    """the reference's instantiate() never runs a module's start function (include/runtime.hpp reads no module.starts): a
    program with one commits what _start commits and nothing else; (func (export "_start") ...) is the inline spelling"""
    body = '(import "env" "i64_private_const" (func $pc (param i64) (result i64)))\n(func $init (drop (call $pc (i64.const 1))))\n'
    plain = '(module ' + body + '(func $t (drop (call $pc (i64.const 2))))\n(export "_start" (func $t)))'
    started = '(module ' + body + '(func $t (export "_start") (drop (call $pc (i64.const 2))))\n(start $init))'
    a, b = pr.wat_emit(plain, 64), pr.wat_emit(started, 64)
    assert np.array_equal(a[1], b[1]) and a[4] == b[4] and a[4]["private_consts"] == 1


def test_reference_and_table_instructions_front_end(pr):
    """needs no reference run: tests/golden/tables.wat asserts every result it computes; and what the reference cannot do is refused"""
    import refctx_util as U
    text = open(os.path.join(ROOT, "tests", "golden", "tables.wat")).read()
    for spelling in (text, U.wat_to_wasm(text), U.wat_to_plain(text)):
        st = pr.wat_emit(spelling, 64)[4]
        assert st["violated_constraints"] == 0 and st["asserts"] == 9
    head = ('(module (import "env" "i32_private_const" (func $pc (param i32) (result i32)))\n(type $v (func (result i32)))\n(table 2 funcref)\n'
            '(elem (i32.const 0) $one)\n(elem (i32.const 1) $one)\n(func $one (result i32) (i32.const 1))\n')
    tail = '\n(export "_start" (func $t)))'
    for body, why in (("(func $t (drop (table.get (i32.const 2))))", "table_get: index out of range"),
                      ("(func $t (table.set (i32.const 0) (i32.const 1)))", "type mismatch: table.set applied to an i32"),
                      ("(func $t (drop (table.get (call $pc (i32.const 0)))))", "concrete operands"),
                      ("(func $t (table.fill (i32.const 1) (ref.null func) (i32.const 2)))", "table_fill: index out of bound"),
                      ("(func $t (table.init 1 (i32.const 0) (i32.const 1) (i32.const 1)))", "table_init: index out of bound"),
                      ("(func $t (elem.drop 1))", "elem.drop of a segment other than 0"),
                      ("(func $t (local $r funcref))", "locals of a reference type"),
                      ("(func $t (drop (i32.add (ref.null func) (i32.const 1))))", "type mismatch: i32.add applied to an funcref"),
                      ("(func $t (table.set (i32.const 0) (ref.func $pc)) (drop (call_indirect (type $v) (i32.const 0))))", "indirect call of an imported function"),
                      ("(func $t (drop (ref.func $nope)))", "unknown function")):
        with pytest.raises(pr.ProverError, match=why):
            pr.wat_emit(head + body + tail, 64)


def test_env_output_functions(pr):
    """env.print_str / dump_memory (host_modules/env.hpp:92-126) only read linear memory: no row; file_size_get / file_get would read
    host files on behalf of the guest and are refused"""
    prog = ('(module (import "env" "print_str" (func $p (param i32 i32))) (import "env" "dump_memory" (func $d (param i32 i32))) (memory 1) (data (i32.const 8) "hi\\n")'
            ' (func (export "_start") (call $p (i32.const 8) (i32.const 3)) (call $d (i32.const 8) (i32.const 3))))')
    kinds, _, _, _, st = pr.wat_emit(prog, 64)
    assert len(kinds) == 0 and st["violated_constraints"] == 0
    with pytest.raises(pr.ProverError, match="reaches outside the memory"):
        pr.wat_emit(prog.replace("(i32.const 8) (i32.const 3)) (call $d", "(i32.const 65534) (i32.const 3)) (call $d"), 64)
    with pytest.raises(pr.ProverError, match="call of env.file_get, which the front end does not provide"):
        pr.wat_emit('(module (import "env" "file_get" (func $p (param i64 i64) (result i32))) (memory 1) (func (export "_start") (drop (call $p (i64.const 8) (i64.const 3)))))', 64)


def test_module_spelled_the_way_wasm2wat_prints_compiled_code(pr):
    """tests/golden/compiled_style.wat: numeric type uses, (;n;) comments, `align=`, labels as comments, a shadow-stack global, WASI
    start-up, a private argument squared in a loop, an indirect call and proc_exit -- no compiler or wabt exists here, so this
    stands in for the text a maintainer would get from a compiled guest"""
    text = open(os.path.join(ROOT, "tests", "golden", "compiled_style.wat")).read()
    out = pr.wat_emit(text, 64, args=[b"Ligero\0", (7).to_bytes(8, "little")], private_indices=[1], want_exit_code=True)
    assert out[4]["violated_constraints"] == 0 and out[4]["asserts"] == 1 and out[5] == 0
    wrong = pr.wat_emit(text, 64, args=[b"Ligero\0", (8).to_bytes(8, "little")], private_indices=[1])
    assert wrong[4]["violated_constraints"] == 1                      # 8^4 mod 2^16 is not 2401
    public = pr.wat_emit(text, 64, args=[b"Ligero\0", (7).to_bytes(8, "little")], private_indices=[])
    assert public[4]["quadratic_slots"] < out[4]["quadratic_slots"]   # a public argument is a number: the squaring commits nothing


def test_every_family_interleaved_holds_its_assertions(pr):
    """needs no reference run: tests/golden/kitchen_sink.wat asserts what it computes (nine assertions over every instruction family)"""
    import refctx_util as U
    text = open(os.path.join(ROOT, "tests", "golden", "kitchen_sink.wat")).read()
    for spelling in (text, U.wat_to_wasm(text), U.wat_to_plain(text)):
        st = pr.wat_emit(spelling, 64)[4]
        assert st["violated_constraints"] == 0 and st["asserts"] == 9 and st["private_consts"] == 11


@pytest.mark.skipif(not os.path.isdir("/root/reference/tests"), reason="reference tree not present")
def test_all_programs_of_the_reference_in_three_spellings(pr):
    """all 69 programs of the reference's tests/, read where they lie: the WebAssembly binary assembled from each (what
    `webgpu_prover` is given after wat2wasm) and its plain-instruction spelling give the rows of the folded text, coefficient
    rows and const_sum included"""
    import glob
    import refctx_util as U
    files = sorted(glob.glob("/root/reference/tests/*.wat"))
    assert len(files) == 69
    seed = bytes(range(32))
    for path in files:
        text = open(path).read()
        a = pr.wat_emit(text, 64, seed)
        assert a[4]["violated_constraints"] == 0, path
        for spelling in (U.wat_to_wasm(text), U.wat_to_plain(text)):
            b = pr.wat_emit(spelling, 64, seed)
            assert list(a[0]) == list(b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and a[3] == b[3], path

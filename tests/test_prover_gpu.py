"""Three-stage prover over a witness matrix (lgrp_prove: stage 1/2/3 through the C ABI on the GPU + host glue)
against the CPU restatement of src/webgpu_prover.cpp:249-471 (oracle/prover_ref.py), field by field and byte
for byte on the proof envelope.  SURVEY 8f rows N1-N3."""
import hashlib
import importlib
import random

import numpy as np
import pytest

from oracle import prover_ref as ref

pytestmark = pytest.mark.gpu
P = ref.P


@pytest.fixture(scope="module")
def pr(lgr):
    return importlib.import_module("ligero_prover_b200.prover")


def make_statement(oracle, l, kinds, seed, small=False):
    """satisfiable rows: triples with z = x*y, random linear rows, coefficient rows, const_sum = -sum val*coef"""
    rng = random.Random(seed)
    draw = (lambda: rng.randrange(1 << 64)) if small else (lambda: rng.randrange(P))
    vals, coefs, acc = [], [], 0
    for kind in kinds:
        if kind:
            x = [draw() for _ in range(l)]; y = [draw() for _ in range(l)]
            rows = [x, y, [a * b % P for a, b in zip(x, y)]]
        else:
            rows = [[draw() for _ in range(l)]]
        for r in rows:
            c = [rng.randrange(P) for _ in range(l)]
            acc += sum(a * b for a, b in zip(r, c))
            vals.append(r); coefs.append(c)
    values = np.stack([oracle.to_limbs(r) for r in vals])
    coef = np.stack([oracle.to_limbs(c) for c in coefs])
    return values, coef, (-acc) % P


@pytest.mark.parametrize("k,l,kinds,small", [
    (64, 40, [0, 1, 0, 1, 1, 0], False),
    (256, 64, [1, 1, 1, 0, 0, 1, 0], True),
    (256, 64, [0] * 5, False),
    (2048, 1856, [1, 0], False),
    (8192, 8000, [0, 1], True),          # the reference's default geometry; i64_mul.wat-sized: 1 linear row + 1 triple + 3 masks
])
def test_prove_matches_cpu_restatement(lgr, oracle, pr, executor_factory, k, l, kinds, small):
    n = 4 * k
    ex = executor_factory(k, l)
    values, coefs, const_sum = make_statement(oracle, l, kinds, seed=k + len(kinds), small=small)
    enc_seed = hashlib.sha256(b"encoding seed %d" % k).digest()
    inst = hashlib.sha256(b"instance").digest()
    prog = hashlib.sha256(b"program").digest()
    proof = pr.prove(ex, kinds, values, coefs, const_sum, enc_seed, inst, prog, generated_at=1792214281)
    want = ref.prove(l, k, kinds, values, coefs, const_sum, enc_seed, inst)
    info = proof.info()
    assert info["valid"] == (True, True, True) == want["valid"]
    assert info["stage1_seed"] == want["stage1_seed"] and info["stage2_seed"] == want["stage2_seed"]
    assert info["encoded_rows"] == values.shape[0] + 3
    env = ref.parse_envelope(proof.gzip)
    pf = env.ligero_proof
    assert pf.merkle_tree.root.value == want["root"]
    assert np.array_equal(np.array(pf.encoded_code.values, np.uint32).reshape(n, 8), want["code"])
    assert np.array_equal(np.array(pf.encoded_linear.values, np.uint32).reshape(n, 8), want["linear"])
    assert np.array_equal(np.array(pf.encoded_quadratic.values, np.uint32).reshape(n, 8), want["quad"])
    assert list(pf.merkle_tree.leaf_indices) == want["sample"]
    assert [s.value for s in pf.merkle_tree.sibling_hashes] == want["siblings"]
    assert np.array_equal(np.array(pf.sampled_data.values, np.uint32).reshape(want["samplings"].shape), want["samplings"])
    md = env.metadata
    assert (md.packing_size, md.codeword_size, md.sample_size, md.security_level, md.proof_schema_version, md.proof_type) == (k, n, 192, 128, 1, 1)
    # the whole container, byte for byte, against google.protobuf's serialisation of the oracle's values
    meta = {"prover_version": "1.5.0", "program_hash": prog, "generated_at": 1792214281, "k": k, "n": n, "sample_size": 192}
    assert proof.envelope == ref.build_envelope(meta, want["root"], want["siblings"], want["sample"], want["code"], want["linear"],
                                                want["quad"], want["samplings"])
    # verifier-side consistency of the openings (src/webgpu_verifier.cpp:314-315,412-442)
    if k <= 2048:
        assert ref.verify_openings(env, l, k, kinds, coefs, inst)
    proof.close()


def test_prove_rejects_a_false_statement(lgr, oracle, pr, executor_factory):
    k, l, kinds = 64, 40, [1, 0]
    ex = executor_factory(k, l)
    values, coefs, const_sum = make_statement(oracle, l, kinds, seed=5)
    seed = bytes(32)
    bad = values.copy(); bad[2, 7, 0] ^= 1                       # z[7] != x[7]*y[7]
    info = pr.prove(ex, kinds, bad, coefs, const_sum, seed).info()
    assert info["valid"][0] and not info["valid"][2]
    info = pr.prove(ex, kinds, values, coefs, (const_sum + 1) % P, seed).info()
    assert info["valid"] == (True, False, True)
    assert pr.prove(ex, kinds, values, coefs, const_sum, seed).info()["valid"] == (True, True, True)


def test_prove_large_matrix_properties(lgr, oracle, pr, executor_factory):
    """several tiles per stage (tile = 2^22 codeword elements) with triples straddling tile boundaries: checked
    through properties -- self-check passes, openings recommit to the root, sampled columns of the first rows equal
    the oracle's encoding"""
    k, l = 256, 64
    n = 4 * k
    kinds = ([1, 0, 0] * 1800)[:5000]
    ex = executor_factory(k, l)
    rng = np.random.default_rng(3)
    rows = int(len(kinds) + 2 * sum(kinds))
    values = np.zeros((rows, l, 8), np.uint32)
    values[:, :, :2] = rng.integers(0, 1 << 32, size=(rows, l, 2), dtype=np.uint32)        # 64-bit witnesses
    r = 0
    for kind in kinds:                                                                       # z = x*y on the triples
        if kind:
            x = values[r, :, 0].astype(object) + (values[r, :, 1].astype(object) << 32)
            y = values[r + 1, :, 0].astype(object) + (values[r + 1, :, 1].astype(object) << 32)
            values[r + 2] = oracle.to_limbs([int(a) * int(b) % P for a, b in zip(x, y)])
            r += 3
        else:
            r += 1
    proof = pr.prove(ex, kinds, values, None, 0, hashlib.sha256(b"big").digest())
    assert proof.info()["valid"] == (True, True, True) and proof.info()["encoded_rows"] == rows + 3
    env = ref.parse_envelope(proof.gzip)
    pf = env.ligero_proof
    sample = list(pf.merkle_tree.leaf_indices)
    samp = np.array(pf.sampled_data.values, np.uint32).reshape(rows + 3, 192, 8)
    sha = oracle.Sha(192); sha.init()
    for t in range(rows + 3):
        sha.update(samp[t])
    leaves = dict(zip(sample, (d.tobytes() for d in sha.final())))
    assert ref.recommit(leaves, sample, 2 * n - 1, [s.value for s in pf.merkle_tree.sibling_hashes]) == pf.merkle_tree.root.value
    enc = ref.FrRandomStream(hashlib.sha256(b"big").digest())
    for t in range(3):
        row = np.zeros((k, 8), np.uint32); row[:l] = values[t]; row[l:] = oracle.to_limbs(enc.take(k - l))
        assert np.array_equal(samp[t], oracle.encode(row, k)[sample])
    proof.close()


def test_prove_matches_committed_fixture(lgr, pr, executor_factory):
    """the GPU prover reproduces the committed golden proof (tests/golden/prover_k64.json) byte for byte"""
    import json, os
    fx = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "prover_k64.json")))
    l, k = fx["l"], fx["k"]
    values = np.frombuffer(bytes.fromhex(fx["values_hex"]), np.uint32).reshape(-1, l, 8)
    coefs = np.frombuffer(bytes.fromhex(fx["coefs_hex"]), np.uint32).reshape(-1, l, 8)
    ex = executor_factory(k, l)
    proof = pr.prove(ex, fx["kinds"], values, coefs, int(fx["const_sum"], 16), bytes.fromhex(fx["encoding_seed"]), bytes.fromhex(fx["instance_hash"]),
                     bytes.fromhex(fx["program_hash"]), generated_at=fx["generated_at"])
    info = proof.info()
    assert list(info["valid"]) == fx["valid"] and info["stage1_seed"].hex() == fx["stage1_seed"] and info["stage2_seed"].hex() == fx["stage2_seed"]
    env = proof.envelope
    assert len(env) == fx["envelope_len"] and hashlib.sha256(env).hexdigest() == fx["envelope_sha256"]
    proof.close()


def test_witnesses_to_proof_through_the_packer(lgr, oracle, pr, executor_factory):
    """released witnesses -> row_packer (witness_manager's packing) -> three-stage prover, against the CPU restatement of
    both steps: the seam where the interpreter's backend would plug in"""
    k, l = 64, 40
    rng = random.Random(77)
    ws, acc = [], 0
    for _ in range(230):
        if rng.random() < 0.55:
            v, c = rng.randrange(1 << 64), rng.randrange(P)
            ws.append(("L", v, c)); acc += v * c
        else:
            x, y = rng.randrange(1 << 64), rng.randrange(1 << 64)
            cs = tuple(rng.randrange(P) for _ in range(3))
            ws.append(("Q", (x, y, x * y % P), cs)); acc += x * cs[0] + y * cs[1] + (x * y % P) * cs[2]
    const_sum = (-acc) % P
    pk = pr.RowPacker(l)
    for w in ws:
        if w[0] == "L":
            pk.push_linear(oracle.to_limbs([w[1]])[0], oracle.to_limbs([w[2]])[0])
        else:
            pk.push_quadratic(oracle.to_limbs(w[1]), oracle.to_limbs(w[2]))
    pk.finalize()
    kinds, vals, coefs = pk.rows()
    wk, wv, wc = ref.pack_rows(l, ws)
    assert list(kinds) == wk and len(set(wk)) == 2
    ref_vals = np.stack([oracle.to_limbs(r) for r in wv]); ref_coefs = np.stack([oracle.to_limbs(r) for r in wc])
    assert np.array_equal(vals, ref_vals) and np.array_equal(coefs, ref_coefs)
    ex = executor_factory(k, l)
    seed = hashlib.sha256(b"packer").digest()
    proof = pr.prove(ex, kinds, vals, coefs, const_sum, seed, bytes(32), bytes(32), generated_at=1)
    want = ref.prove(l, k, wk, ref_vals, ref_coefs, const_sum, seed, bytes(32))
    assert proof.info()["valid"] == (True, True, True) == want["valid"]
    env = ref.parse_envelope(proof.envelope)
    assert env.ligero_proof.merkle_tree.root.value == want["root"]
    assert np.array_equal(np.array(env.ligero_proof.sampled_data.values, np.uint32).reshape(want["samplings"].shape), want["samplings"])
    proof.close(); pk.close()


@pytest.mark.parametrize("k,l,kinds", [(64, 40, []), (8, 4, [0, 1, 0]), (16, 16, [1])])
def test_prove_edge_geometries(lgr, oracle, pr, executor_factory, k, l, kinds):
    """an empty statement (the three mask rows only), n smaller than the sample size (every column opened, no sibling
    hashes), and l = k (no padding columns)"""
    n = 4 * k
    ex = executor_factory(k, l)
    values, coefs, const_sum = make_statement(oracle, l, kinds, seed=3) if kinds else (np.zeros((0, l, 8), np.uint32), np.zeros((0, l, 8), np.uint32), 0)
    seed = hashlib.sha256(b"edge%d" % k).digest()
    proof = pr.prove(ex, kinds, values, coefs, const_sum, seed, bytes(32), bytes(32), generated_at=1)
    want = ref.prove(l, k, kinds, values, coefs, const_sum, seed, bytes(32))
    assert proof.info()["valid"] == (True, True, True) == want["valid"]
    assert proof.info()["encoded_rows"] == values.shape[0] + 3
    env = ref.parse_envelope(proof.envelope)
    pf = env.ligero_proof
    assert pf.merkle_tree.root.value == want["root"]
    assert list(pf.merkle_tree.leaf_indices) == want["sample"] and len(want["sample"]) == min(192, n)
    assert [s.value for s in pf.merkle_tree.sibling_hashes] == want["siblings"]
    if n <= 192:
        assert want["siblings"] == []
    assert np.array_equal(np.array(pf.sampled_data.values, np.uint32).reshape(want["samplings"].shape), want["samplings"])
    meta = {"prover_version": "1.5.0", "program_hash": bytes(32), "generated_at": 1, "k": k, "n": n, "sample_size": 192}
    assert proof.envelope == ref.build_envelope(meta, want["root"], want["siblings"], want["sample"], want["code"], want["linear"], want["quad"], want["samplings"])
    proof.close()


def test_cpp_prover_program(lgr, oracle, tmp_path):
    """tests/cpp/test_prover.cpp: witnesses -> row_packer -> matrix_prover in plain C++ (header-only host layer over the C
    ABI), compared with the CPU restatement of the packing and of the three stages"""
    import os, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "cpp", "test_prover")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.join(root, "tests", "cpp"), "-s"])
    k, l = 256, 100
    rng = random.Random(2026)
    ws, acc = [], 0
    for _ in range(470):
        if rng.random() < 0.5:
            v, c = rng.randrange(P), rng.randrange(P)
            ws.append(("L", v, c)); acc += v * c
        else:
            x, y = rng.randrange(P), rng.randrange(P)
            cs = tuple(rng.randrange(P) for _ in range(3))
            ws.append(("Q", (x, y, x * y % P), cs)); acc += x * cs[0] + y * cs[1] + (x * y % P) * cs[2]
    ws.append(("L", 1, (-acc) % P))                                   # closes the linear relation: const_sum = 0
    rec = np.zeros((len(ws), 49), np.uint32)
    for i, w in enumerate(ws):
        if w[0] == "L":
            rec[i, 1:9] = oracle.to_limbs([w[1]])[0]; rec[i, 25:33] = oracle.to_limbs([w[2]])[0]
        else:
            rec[i, 0] = 1
            rec[i, 1:25] = oracle.to_limbs(w[1]).reshape(-1); rec[i, 25:49] = oracle.to_limbs(w[2]).reshape(-1)
    fw, fr_, fp = tmp_path / "w.bin", tmp_path / "roots.bin", tmp_path / "proof.gz"
    rec.tofile(fw)
    lgr.ints_to_array(list(lgr.generate_omegas(k, 4 * k))).tofile(fr_)
    out = subprocess.check_output([exe, str(k), str(l), str(fr_), str(fw), str(fp)], stderr=subprocess.STDOUT, timeout=300).decode()
    assert "ok" in out and "valid 1 1 1" in out, out
    wk, wv, wc = ref.pack_rows(l, ws)
    vals = np.stack([oracle.to_limbs(r) for r in wv]); coefs = np.stack([oracle.to_limbs(r) for r in wc])
    want = ref.prove(l, k, wk, vals, coefs, 0, bytes((i * 7 + 1) & 0xFF for i in range(32)), bytes(32))
    assert want["valid"] == (True, True, True)
    lines = dict(line.split(" ", 1) for line in out.strip().splitlines())
    assert lines["root"] == want["root"].hex() and lines["stage1"] == want["stage1_seed"].hex() and lines["stage2"] == want["stage2_seed"].hex()
    env = ref.parse_envelope(fp.read_bytes())
    meta = {"prover_version": "1.5.0", "program_hash": bytes(32), "generated_at": 1, "k": k, "n": 4 * k, "sample_size": 192}
    assert env.SerializeToString(deterministic=True) == ref.build_envelope(meta, want["root"], want["siblings"], want["sample"], want["code"],
                                                                           want["linear"], want["quad"], want["samplings"])


# ---------------------------------------------------------------- vbn254fr batch events
def batch_statement(oracle, l, seed):
    """a small vbn254fr program between scalar rows: sets, a product, its copy, an assertion that holds, a quotient,
    bit decompositions of a 0/1 vector, constant arithmetic.  Returns everything lgrp_prove / prover_ref.prove take."""
    from oracle import prover_ref as R
    rng = random.Random(seed)
    kinds, args, consts, vals, coefs = [], [], [], [], []
    acc = 0

    def host_row(row, with_coef=True):
        nonlocal acc
        c = [rng.randrange(P) if with_coef else 0 for _ in range(l)]
        if with_coef:
            acc += sum(a * b for a, b in zip(row, c))
        vals.append(row); coefs.append(c)

    def ev(kind, a=0, b=0, c=0, konst=None):
        kinds.append(kind)
        if kind >= R.EV_VSET:
            args.append([a, b, c])
        if konst is not None:
            consts.append(konst)

    ev(R.EV_LINEAR); host_row([rng.randrange(P) for _ in range(l)])
    ev(R.EV_VSET, 0); host_row([rng.randrange(P) for _ in range(l)], False)            # v0
    ev(R.EV_VSET, 1); host_row([rng.randrange(1, P) for _ in range(l)], False)         # v1 (non-zero: it divides below)
    ev(R.EV_VMUL, 2, 0, 1)                                                              # v2 = v0 * v1
    x = [rng.randrange(P) for _ in range(l)]; y = [rng.randrange(P) for _ in range(l)]
    ev(R.EV_QUAD); host_row(x); host_row(y); host_row([a * b % P for a, b in zip(x, y)])
    ev(R.EV_VCOPY, 3, 2)                                                                # v3 = v2
    ev(R.EV_VDIV, 4, 3, 1)                                                              # v4 = v3 / v1 (= v0)
    ev(R.EV_VASSERT_EQ, 4, 0)                                                           # holds on all k elements
    ev(R.EV_VSET, 5); host_row([rng.randrange(2) for _ in range(l)], False)            # bits
    ev(R.EV_VBIT, 6, 5, 0)                                                              # v6 = bit 0 of v5 (pads: bit 0 of random pads)
    ev(R.EV_VMUL, 7, 0, 0)                                                              # a square: x == y
    c1 = rng.randrange(P)
    ev(R.EV_VADDC, 8, 7, 0, c1); ev(R.EV_VSUBC, 8, 8, 0, c1)                            # v8 = v7
    ev(R.EV_VASSERT_EQ, 8, 7)
    ev(R.EV_VADD, 9, 0, 1); ev(R.EV_VSUB, 9, 9, 1); ev(R.EV_VASSERT_EQ, 9, 0)
    c2 = rng.randrange(1, P)
    ev(R.EV_VMULC, 10, 0, 0, c2); ev(R.EV_VMONTMULC, 10, 10, 0, pow(c2, -1, P) * (1 << 256) % P); ev(R.EV_VASSERT_EQ, 10, 0)
    ev(R.EV_VCSUB, 11, 0, 0, 0); ev(R.EV_VADD, 11, 11, 0)                                # (0 - v0) + v0 = 0
    ev(R.EV_VBIT, 12, 11, 7)                                                            # bit of zero
    ev(R.EV_LINEAR); host_row([rng.randrange(P) for _ in range(l)])
    values = np.stack([oracle.to_limbs(r) for r in vals])
    coef = np.stack([oracle.to_limbs(c) for c in coefs])
    return (np.array(kinds, np.uint8), values, coef, (-acc) % P, 13, np.array(args, np.uint32),
            np.stack([oracle.to_limbs([c])[0] for c in consts]))


@pytest.mark.parametrize("k,l", [(256, 64), (512, 320)])
def test_prove_with_vbn254fr_batch_events(lgr, oracle, pr, executor_factory, k, l):
    """on_batch_init / bit / equal / quadratic rows (nonbatch_context.hpp:497-553,782-847,996-1047) driven by vbn254fr
    calls on device-resident variables: proof equal to the CPU restatement, self-check and openings valid"""
    n = 4 * k
    ex = executor_factory(k, l)
    kinds, values, coefs, const_sum, slots, args, consts = batch_statement(oracle, l, seed=k)
    enc_seed = hashlib.sha256(b"batch encoding seed").digest()
    inst = hashlib.sha256(b"batch instance").digest()
    proof = pr.prove(ex, kinds, values, coefs, const_sum, enc_seed, inst, generated_at=7, arena_slots=slots, batch_args=args, batch_consts=consts)
    want = ref.prove(l, k, kinds, values, coefs, const_sum, enc_seed, inst, arena_slots=slots, batch_args=args, batch_consts=consts)
    info = proof.info()
    assert want["valid"] == (True, True, True)
    assert info["valid"] == (True, True, True)
    assert info["encoded_rows"] == want["encoded_rows"]
    assert info["stage1_seed"] == want["stage1_seed"] and info["stage2_seed"] == want["stage2_seed"]
    env = ref.parse_envelope(proof.gzip)
    pf = env.ligero_proof
    assert pf.merkle_tree.root.value == want["root"]
    for name, key in (("encoded_code", "code"), ("encoded_linear", "linear"), ("encoded_quadratic", "quad")):
        assert np.array_equal(np.array(getattr(pf, name).values, np.uint32).reshape(n, 8), want[key]), name
    assert np.array_equal(np.array(pf.sampled_data.values, np.uint32).reshape(want["samplings"].shape), want["samplings"])
    assert ref.verify_openings(env, l, k, kinds, coefs, inst)
    proof.close()
    # a violated batch assertion must fail the quadratic test (and only it)
    bad = kinds.copy()
    i = [j for j, kd in enumerate(kinds) if kd == ref.EV_VASSERT_EQ][0]
    bargs = args.copy()
    bi = int((kinds[:i] >= ref.EV_VSET).sum())
    bargs[bi] = [4, 1, 0]                                              # v4 (= v0) against v1
    info = pr.prove(ex, bad, values, coefs, const_sum, enc_seed, inst, arena_slots=slots, batch_args=bargs, batch_consts=consts).info()
    assert info["valid"] == (True, True, False)


# ---------------------------------------------------------------- BASELINE config 4, bounded: .wat text -> proof
@pytest.mark.parametrize("k,l", [(8192, 8000), (256, 64)])
def test_prove_wat_end_to_end(lgr, oracle, pr, executor_factory, k, l):
    """a .wat program through the product's own entry point (lgrp_prove_wat): front end, witness emitter, row packing,
    stage 1, linear-test coefficients derived from the stage-1 seed, stages 2 and 3.  The proof parses, passes the
    prover's self-check and the verifier-side opening checks, and equals the CPU prover run on the same rows and seed."""
    import os
    n = 4 * k
    ref_wat = "/root/reference/tests/i64_mul.wat"                      # the program BASELINE names, where the tree exists
    path = ref_wat if os.path.exists(ref_wat) else os.path.join(os.path.dirname(__file__), "golden", "mul64.wat")
    text = open(path).read()
    ex = executor_factory(k, l)
    enc_seed = hashlib.sha256(b"config 4 encoding seed").digest()
    proof, st = pr.prove_wat(ex, text, enc_seed, generated_at=11)
    info = proof.info()
    assert st["violated_constraints"] == 0 and info["valid"] == (True, True, True)
    env = ref.parse_envelope(proof.gzip)
    assert env.metadata.program_hash.value == hashlib.sha256(text.encode()).digest()
    assert (env.metadata.packing_size, env.metadata.codeword_size) == (k, n)
    # the statement the product proved, re-derived: rows from the emitter, coefficients from the proof's own stage-1 seed
    kinds, vals, coefs, const_sum, _ = pr.wat_emit(text, l, info["stage1_seed"])
    assert info["encoded_rows"] == vals.shape[0] + 3
    if k == 8192 and path == ref_wat:
        assert list(kinds) == [0, 1]                                   # 1 linear row + 1 triple + 3 masks = 7 encodes
    want = ref.prove(l, k, kinds, vals, coefs, const_sum, enc_seed, bytes(32))
    assert want["valid"] == (True, True, True)
    assert env.ligero_proof.merkle_tree.root.value == want["root"] and info["stage1_seed"] == want["stage1_seed"]
    assert info["stage2_seed"] == want["stage2_seed"]
    assert np.array_equal(np.array(env.ligero_proof.encoded_linear.values, np.uint32).reshape(n, 8), want["linear"])
    assert np.array_equal(np.array(env.ligero_proof.sampled_data.values, np.uint32).reshape(want["samplings"].shape), want["samplings"])
    assert ref.verify_openings(env, l, k, kinds, coefs, bytes(32))
    proof.close()
    # a program whose assertion is false: same pipeline, linear test fails
    bad = text.replace("(i64.const 1)))", "(i64.const 2)))", 1)
    proof, st = pr.prove_wat(ex, bad, enc_seed)
    assert st["violated_constraints"] == 1 and proof.info()["valid"] == (True, False, True)
    proof.close()

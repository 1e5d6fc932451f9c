"""The REFERENCE's own prover layers as the checker (B200 part).

(1) The GPU prover (lgrp_prove: the batched path over lgr.h) must reproduce, bit for bit, what the reference's own stage
    contexts / backend / interpreter committed for the same statement (tests/golden/refctx_*.json, generated from the
    reference's sources by tests/golden/make_refctx_vectors.py) -- incl. BASELINE config 4, tests/i64_mul.wat at k = 8192.
(2) oracle/_ref/refctx_cuda is the reference's stage contexts, vbn254fr module and interpreter compiled with
    Executor = ligero::webgpu_context := the CUDA executor (ligero-prover_b200/host/compat): the unchanged reference code
    drives liblgr.so one row at a time, and must arrive at the same vectors."""
import importlib
import json
import os
import subprocess
import tempfile

import numpy as np
import pytest

from oracle import prover_ref as ref
import refctx_util as U

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pr(lgr):
    return importlib.import_module("ligero_prover_b200.prover")


@pytest.mark.parametrize("case", U.CASES)
def test_gpu_prover_reproduces_the_reference_run(lgr, pr, executor_factory, case):
    st = U.load(case)
    fx, l, k, n = st["fx"], st["l"], st["k"], st["n"]
    ex = executor_factory(k, l)
    proof = pr.prove(ex, st["kinds"], st["values"], st["coefs"], st["const_sum"], st["encoding_seed"], st["instance_hash"], generated_at=1,
                     arena_slots=st["slots"], batch_args=st["args"] if len(st["args"]) else None,
                     batch_consts=st["consts"] if len(st["consts"]) else None)
    info = proof.info()
    assert list(info["valid"]) == [True, True, True]
    assert info["stage1_seed"].hex() == fx["stage1_seed"] and info["stage2_seed"].hex() == fx["stage2_seed"]
    U.check_envelope(ref.parse_envelope(proof.gzip), fx, ref.sibling_positions)
    proof.close()


@pytest.mark.skipif(not os.path.exists(U.REF_BIN_CUDA), reason="oracle/_ref/refctx_cuda not built (needs /root/reference at build time)")
@pytest.mark.parametrize("case", U.CASES)
def test_reference_stage_contexts_run_on_the_cuda_executor(case):
    """the drop-in boundary, run: the reference's unchanged stage-1/2/3 contexts, witness manager, interpreter and vbn254fr
    module, with liblgr.so behind `webgpu_context`, give the vectors they give over the CPU oracle"""
    gen = U.compact_module()
    with tempfile.TemporaryDirectory() as tmp:
        prog, k, name = U.harness_args(case, tmp)
        path = os.path.join(tmp, "out.json")
        res = subprocess.run([U.REF_BIN_CUDA, prog, k, path], capture_output=True, text=True, timeout=600)
        assert res.returncode == 0, (res.stdout + res.stderr)[-3000:]
        raw = json.load(open(path))
    assert raw["executor"] == "cuda"
    raw["program"] = name
    got = gen.compact(raw)
    want = U.load(case)["fx"]
    for key in want:
        if key != "generated_by":
            assert got[key] == want[key], key


@pytest.mark.parametrize("case,wat", [("i64_mul_k8192", None), ("mul64_k256", "mul64"), ("arith32_k256", "arith32"), ("intops_k256", "intops")])
def test_prove_wat_commits_to_the_root_of_the_reference_run(lgr, pr, executor_factory, case, wat):
    """BASELINE config 4 through the product's entry point: lgrp_prove_wat on the text of tests/i64_mul.wat, with the
    encoding seed of the reference run, arrives at the Merkle root the reference's own interpreter + stage-1 context
    arrive at (the root depends on every committed row and pad, not on the instance hash, which the two runs choose
    differently)"""
    st = U.load(case)
    text = open(U.WAT_TEXT[wat]).read() if wat else U.binop_wat("mul", U.I64_MUL_CASES)
    ex = executor_factory(st["k"], st["l"])
    proof, stats = pr.prove_wat(ex, text, st["encoding_seed"], generated_at=3)
    assert stats["violated_constraints"] == 0 and proof.info()["valid"] == (True, True, True)
    env = ref.parse_envelope(proof.gzip)
    assert env.ligero_proof.merkle_tree.root.value.hex() == st["fx"]["root"]
    proof.close()
    if wat == "arith32":                                      # the same program as a WebAssembly binary (the reference takes both)
        proof, stats = pr.prove_wat(ex, U.wat_to_wasm(text), st["encoding_seed"], generated_at=3)
        assert stats["violated_constraints"] == 0 and proof.info()["valid"] == (True, True, True)
        assert ref.parse_envelope(proof.gzip).ligero_proof.merkle_tree.root.value.hex() == st["fx"]["root"]
        proof.close()


@pytest.mark.skipif(not os.path.exists(U.REF_BIN_CPU), reason="oracle/_ref/refctx_cpu not built (needs /root/reference at build time)")
@pytest.mark.parametrize("case", ["i64_mul_k8192", "vbn_k256", "mul64_k256", "intops_k256"])
def test_reference_verifier_accepts_the_gpu_proof(lgr, pr, executor_factory, case, tmp_path):
    """the proof lgrp_prove makes on the B200 goes to the REFERENCE's verifier (its nonbatch_verifier_context re-running the
    program over the opened columns, merkle_tree::recommit, the seven checks of src/webgpu_verifier.cpp:412-442): accepted
    by the verifier running over the CPU oracle and, where built, by the same verifier running over the CUDA executor"""
    st = U.load(case)
    ex = executor_factory(st["k"], st["l"])
    proof = pr.prove(ex, st["kinds"], st["values"], st["coefs"], st["const_sum"], st["encoding_seed"], st["instance_hash"], generated_at=5,
                     arena_slots=st["slots"], batch_args=st["args"] if len(st["args"]) else None,
                     batch_consts=st["consts"] if len(st["consts"]) else None)
    env = ref.parse_envelope(proof.gzip)
    proof.close()
    path = str(tmp_path / "gpu.proof")
    U.envelope_to_proof_file(path, env, ref.sibling_positions, st["fx"]["decommit_total"], st["instance_hash"])
    rc, msg = U.reference_verifier(U.REF_BIN_CPU, case, path, str(tmp_path))
    assert rc == 0, msg
    if os.path.exists(U.REF_BIN_CUDA):
        rc, msg = U.reference_verifier(U.REF_BIN_CUDA, case, path, str(tmp_path))
        assert rc == 0, msg


@pytest.mark.skipif(not (os.path.exists(U.REF_BIN_CUDA) and os.path.exists(U.REF_BIN_CPU)), reason="oracle/_ref/refctx_{cpu,cuda} not built")
@pytest.mark.parametrize("prog,k", [("i64_mul3", 512), ("vbn", 1024), ("i64_mul3", 2048), ("i64_mul3", 4096), ("vbn", 8192)])
def test_reference_contexts_on_cuda_equal_reference_contexts_on_the_oracle(prog, k, tmp_path):
    """live differential at geometries without a committed fixture: the same reference code (stage contexts, witness
    manager, interpreter / vbn254fr module, verifier) over the CUDA executor and over the CPU oracle writes the same
    file -- every row event, the root, the leaf digests, the three test vectors, all openings.  k = 512 ... 2048 run the
    fused encoder one row at a time, k = 4096 / 8192 the tile engine with deferred one-row encodes"""
    outs = {}
    for name, binary in (("oracle", U.REF_BIN_CPU), ("cuda", U.REF_BIN_CUDA)):
        path = str(tmp_path / (name + ".json"))
        res = subprocess.run([binary, prog, str(k), path, "11"], capture_output=True, text=True, timeout=600)
        assert res.returncode == 0, (name, (res.stdout + res.stderr)[-3000:])
        outs[name] = json.load(open(path))
    assert outs["cuda"].pop("executor") == "cuda" and outs["oracle"].pop("executor") == "oracle"
    assert outs["cuda"]["verifier"] == [1] * 7
    for key in outs["oracle"]:
        assert outs["cuda"][key] == outs["oracle"][key], key


def test_prove_wat_with_private_arguments_commits_to_the_root_of_the_reference_run(lgr, pr, executor_factory):
    """a guest that takes arguments (tests/golden/wasi_args.wat; two of the four private) through lgrp_prove_wat_args: the
    Merkle root is the one the reference's interpreter + wasi_preview1 module + stage-1 context arrive at with the same
    encoding seed, the prover's three self-checks pass, and the stage-1 seed is hash("LigetronStage1", root, instance hash
    folded from the PUBLIC arguments as src/webgpu_prover.cpp:160-168 folds them).  (Last in the GPU suite: the path was
    written after the round's GPU budget was spent; everything it shares with lgrp_prove_wat is covered above.)"""
    st = U.load("wasi_k256")
    args, private = [bytes.fromhex(a) for a in st["fx"]["args"]], st["fx"]["private_indices"]
    text = open(os.path.join(U.HERE, "golden", "wasi_args.wat")).read()
    ex = executor_factory(st["k"], st["l"])
    proof, stats = pr.prove_wat(ex, text, st["encoding_seed"], generated_at=3, args=args, private_indices=private)
    info = proof.info()
    assert stats["violated_constraints"] == 0 and info["valid"] == (True, True, True)
    root = ref.parse_envelope(proof.gzip).ligero_proof.merkle_tree.root.value
    assert root.hex() == st["fx"]["root"]
    assert info["stage1_seed"] == pr.stage1_seed(root, pr.wat_instance_hash(args, private))
    proof.close()

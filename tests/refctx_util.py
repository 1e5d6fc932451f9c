"""helpers shared by tests/test_refctx_cpu.py and tests/test_refctx_gpu.py: the golden vectors written by the REFERENCE's own
stage contexts / backend / interpreter (tests/golden/make_refctx_vectors.py)"""
import base64
import hashlib
import importlib.util
import json
import os
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CASES = ["i64_mul_k8192", "i64_mul3_k256", "vbn_k256"]
REF_BIN_CPU = os.path.join(ROOT, "oracle", "_ref", "refctx_cpu")
REF_BIN_CUDA = os.path.join(ROOT, "oracle", "_ref", "refctx_cuda")


def _unz(s):
    return np.frombuffer(zlib.decompress(base64.b64decode(s)), np.uint32)


def load(case):
    fx = json.load(open(os.path.join(HERE, "golden", "refctx_%s.json" % case)))
    l = fx["l"]
    st = {
        "fx": fx, "l": l, "k": fx["k"], "n": fx["n"],
        "kinds": np.array(fx["kinds"], np.uint8),
        "values": _unz(fx["values_zb64"]).reshape(-1, l, 8),
        "coefs": _unz(fx["coefs_zb64"]).reshape(-1, l, 8),
        "args": _unz(fx["batch_args_zb64"]).reshape(-1, 3),
        "consts": _unz(fx["batch_consts_zb64"]).reshape(-1, 8),
        "const_sum": int.from_bytes(bytes.fromhex(fx["const_sum"]), "little"),
        "encoding_seed": bytes.fromhex(fx["encoding_seed"]),
        "instance_hash": bytes.fromhex(fx["instance_hash"]),
    }
    st["slots"] = int(st["args"][:, :2].max()) + 1 if len(st["args"]) else 0
    return st


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def compact_module():
    spec = importlib.util.spec_from_file_location("make_refctx_vectors", os.path.join(HERE, "golden", "make_refctx_vectors.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def check_envelope(env, fx, sibling_positions):
    """a parsed LigeroProofEnvelope (google.protobuf message, oracle/prover_ref.parse_envelope) against a reference-run
    fixture: root, the three test vectors, the sampled columns, the opened leaves and their sibling hashes"""
    pf = env.ligero_proof
    assert pf.merkle_tree.root.value.hex() == fx["root"]
    for name, key in (("encoded_code", "code"), ("encoded_linear", "linear"), ("encoded_quadratic", "quad")):
        assert sha(np.array(getattr(pf, name).values, np.uint32)) == fx["sha256"][key], name
    assert sha(np.array(pf.sampled_data.values, np.uint32)) == fx["sha256"]["samplings"]
    assert [int(i) for i in pf.merkle_tree.leaf_indices] == fx["sample_index"]
    # the envelope lists siblings in canonical order (proof_serializer.hpp:82-117); the reference's decommit holds them by position
    pos = sibling_positions(fx["sample_index"], fx["decommit_total"])
    assert sorted(pos) == fx["decommit_positions"]
    by_pos = dict(zip(pos, [s.value for s in pf.merkle_tree.sibling_hashes]))
    assert sha(np.frombuffer(b"".join(by_pos[p] for p in fx["decommit_positions"]), np.uint8)) == fx["sha256"]["decommit_siblings"]

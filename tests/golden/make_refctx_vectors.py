#!/usr/bin/env python3
"""Generate tests/golden/refctx_*.json: what the REFERENCE's own prover layers commit for three small programs.

Runs oracle/_ref/refctx_cpu (tests/refctx/ref_contexts.cpp: the reference's stage contexts, backend / witness manager,
interpreter opcode semantics, env + vbn254fr host modules, transcript hash and Merkle tree, compiled where they lie
under /root/reference; device work done by the CPU oracle) and compacts its output: the statement the stage contexts
saw (row events with values and stage-2 coefficient rows, zlib + base64) and what they produced (root, seeds, flags and
sample positions in full; SHA-256 of the big arrays).  Needs /root/reference, so it runs in the build container only;
the committed JSON travels.

    make -C oracle refctx && python tests/golden/make_refctx_vectors.py
"""
import base64
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import zlib

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
BIN = os.path.join(ROOT, "oracle", "_ref", "refctx_cpu")

# (program, k): tests/i64_mul.wat at the reference's default geometry (BASELINE config 4: 1 linear row + 1 triple + 3
# masks); its last three assertions at k = 256 (l = 64: 17 row events, so rows of both kinds interleave); the vbn254fr
# batch calls at k = 256 (271 events: init / equal / quadratic / bit rows between scalar rows)
CASES = [("i64_mul", 8192), ("i64_mul3", 256), ("vbn", 256)]
# programs given as text: the repo's own tests/golden/mul64.wat and arith32.wat (products, sums, differences, nested forms,
# literal operands; 64- and 32-bit) go through the reference's interpreter as a token stream (tests/refctx_util.py: wat_to_tokens)
WAT_CASES = [("mul64", os.path.join(HERE, "mul64.wat"), 256), ("arith32", os.path.join(HERE, "arith32.wat"), 256),
             ("intops", os.path.join(HERE, "intops.wat"), 256)]       # every other integer instruction (tests/golden/make_intops_wat.py)


# a guest with arguments (wasi args_get; argv[1] and argv[3] private): what src/webgpu_prover.cpp:110-157 would build from
# {"args": [{"i64": 400}, {"i64": 600}, {"str": "hello"}], "private-indices": [1, 3]}
WASI_CASE = ("wasi", os.path.join(HERE, "wasi_args.wat"), 256, [b"Ligero\0", (400).to_bytes(8, "little"), (600).to_bytes(8, "little"), b"hello\0"], [1, 3])


def zb64(hexstr):
    return base64.b64encode(zlib.compress(bytes.fromhex(hexstr), 9)).decode()


def sha(hexstr):
    return hashlib.sha256(bytes.fromhex(hexstr)).hexdigest()


def compact(raw):
    out = {key: raw[key] for key in ("program", "l", "k", "n", "encoding_seed", "instance_hash", "kinds", "const_sum", "root", "stage1_seed",
                                     "stage2_seed", "valid", "verifier", "sample_index", "decommit_total")}
    out["generated_by"] = "tests/golden/make_refctx_vectors.py <- oracle/_ref/refctx_cpu (reference headers + CPU oracle executor)"
    for key in ("values", "coefs", "batch_args", "batch_consts"):
        out[key + "_zb64"] = zb64(raw[key])
    out["sha256"] = {key: sha(raw[key]) for key in ("digests", "code", "linear", "quad", "samplings")}
    pos = sorted(int(p) for p in raw["decommit_nodes"])
    out["decommit_positions"] = pos
    out["sha256"]["decommit_siblings"] = sha("".join(raw["decommit_nodes"][str(p)] for p in pos))
    return out


def write(name, k, raw, extra=None):
    assert raw["valid"] == [1, 1, 1], "the reference's self-check must pass on an honest run"
    assert raw["verifier"] == [1] * 7, "the reference's verifier must accept the proof of its own prover passes"
    path = os.path.join(HERE, "refctx_%s_k%d.json" % (name, k))
    with open(path, "w") as f:
        json.dump(dict(compact(raw), **(extra or {})), f, separators=(",", ":"))
        f.write("\n")
    print(path, os.path.getsize(path), "bytes")


def main():
    if not os.path.exists(BIN):
        sys.exit("build it first: make -C oracle refctx (needs /root/reference)")
    sys.path.insert(0, os.path.dirname(HERE))
    import refctx_util
    for name, wat, k in WAT_CASES:
        raw = refctx_util.run_reference_on_wat(open(wat).read(), k)
        raw["program"] = name
        write(name, k, raw)
    name, wat, k, args, private = WASI_CASE
    raw = refctx_util.run_reference_on_wat(open(wat).read(), k, args=args, private_indices=private)
    raw["program"] = name
    write(name, k, raw, {"args": [a.hex() for a in args], "private_indices": private})
    for prog, k in CASES:
        with tempfile.NamedTemporaryFile(suffix=".json") as tmp:
            subprocess.check_call([BIN, prog, str(k), tmp.name])
            raw = json.load(open(tmp.name))
        write(prog, k, raw)


if __name__ == "__main__":
    main()

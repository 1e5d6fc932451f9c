#!/usr/bin/env python3
"""Generates tests/golden/wgslref_vectors.json: OUTPUTS OF THE REFERENCE'S OWN SHADERS run on the host
(oracle/_ref/libwgslref.so, built from /root/reference/shader/* by oracle/wgsl2cpp.py + oracle/wgslref.cpp).

Run in the build container (where /root/reference exists):   python tests/golden/make_wgslref_vectors.py
The committed JSON travels; the tests check oracle/oracle.c (and through it the CUDA path) against it even where
neither the reference tree nor the prebuilt library is present.

Inputs are drawn from Python's `random.Random(seed)` -- independent of the oracle's generator.  Large outputs are
stored as SHA-256 of the output bytes plus a few explicit elements.
"""
import hashlib
import json
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import wref  # noqa: E402

P = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
ROOT1 = pow(7, (P - 1) >> 28, P)               # src/bn254.cpp:36-43
ROOT2 = pow(ROOT1, (1 << 61) - 1, P)


def omegas(k):
    """src/bn254.cpp:51-64"""
    return pow(ROOT1, (1 << 28) // k, P), pow(ROOT1, (1 << 28) // (2 * k), P), pow(ROOT2, (1 << 28) // (4 * k), P)


def rand_elems(seed, count, small=False):
    rnd = random.Random(seed)
    vals = [rnd.randrange(1 << 64) if small else rnd.randrange(P) for _ in range(count)]
    out = np.zeros((count, 8), np.uint32)
    for i, v in enumerate(vals):
        for j in range(8):
            out[i, j] = (v >> (32 * j)) & 0xFFFFFFFF
    return out


def digest(arr):
    return hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest()


def hexel(a):
    return "%064x" % sum(int(a[j]) << (32 * j) for j in range(8))


def summary(arr, picks):
    return {"sha256": digest(arr), "elems": {str(i): hexel(arr[i]) for i in picks}}


ELTWISE = [  # (kernel, uses_y, uses_scalar, accumulates into out)
    ("EltwiseAddMod", True, False, False), ("EltwiseAddAssignMod", False, False, True),
    ("EltwiseAddConstantMod", False, True, False), ("EltwiseSubMod", True, False, False),
    ("EltwiseSubConstantMod", False, True, False), ("EltwiseConstantSubMod", False, True, False),
    ("EltwiseMultMod", True, False, False), ("EltwiseMultConstantMod", False, True, False),
    ("EltwiseMontMultConstantMod", False, True, False), ("EltwiseDivMod", True, False, False),
    ("EltwiseFMAMod", True, False, True), ("EltwiseFMAConstantMod", False, True, True),
]


def main():
    wref.build()
    g = {"generator": "tests/golden/make_wgslref_vectors.py", "source": "reference shader/*.wgsl* via oracle/_ref/libwgslref.so"}
    g["constants"] = {k: "%064x" % v for k, v in wref.constants().items()}

    g["ntt"] = []
    for logn, seed in ((9, 11), (11, 12), (13, 13), (18, 14)):      # 2^18: the > 256-workgroup fallback + ntt_reduce4p
        N = 1 << logn
        w = pow(ROOT1, (1 << 28) // N, P)
        x = rand_elems(seed, N)
        f = wref.ntt(x, w)
        i = wref.ntt(x, w, inverse=True)
        g["ntt"].append({"logn": logn, "seed": seed, "forward": summary(f, (0, 1, N - 1)), "inverse": summary(i, (0, 1, N - 1))})

    g["encode"] = []
    for k, seed, small in ((512, 21, False), (512, 22, True), (8192, 23, False)):
        wk, w2k, wn = omegas(k)
        row = rand_elems(seed, k, small)
        e = wref.encode(row, k, wk, wn)
        d = wref.decode(e, k, wk, w2k, wn)
        g["encode"].append({"k": k, "seed": seed, "small": small, "codeword": summary(e, (0, 1, 4 * k - 1)),
                            "decoded": summary(d, (0, k - 1, k, 4 * k - 1))})
    # mask-row path: iNTT_2k then NTT_n on a 4k buffer (nonbatch_context.hpp:482-494)
    k = 512
    wk, w2k, wn = omegas(k)
    buf = np.zeros((4 * k, 8), np.uint32)
    buf[: 2 * k] = rand_elems(31, 2 * k)
    buf[: 2 * k] = wref.ntt(buf[: 2 * k], w2k, inverse=True)
    e = wref.ntt(buf, wn)
    g["encode_2k"] = {"k": k, "seed": 31, "codeword": summary(e, (0, 1, 4 * k - 1))}

    g["sha"] = []
    for ninst, rows, seed in ((1, 1, 41), (1, 3, 42), (5, 2, 43), (5, 7, 44), (192, 5, 45), (1024, 4, 46), (2048, 3, 47)):
        s = wref.Sha(ninst)
        s.init()
        for r in range(rows):
            s.update(rand_elems(seed * 1000 + r, ninst))
        d = s.final()
        g["sha"].append({"ninst": ninst, "rows": rows, "seed": seed, "digests_sha256": digest(d),
                         "digest0": d[0].tobytes().hex(), "digest_last": d[-1].tobytes().hex()})

    g["eltwise"] = []
    n = 300
    x, y, o = rand_elems(51, n), rand_elems(52, n), rand_elems(53, n)
    y[0] = 0
    y[0, 0] = 1                                       # divisor 1
    x[1] = 0                                          # zero operand
    sc = random.Random(54).randrange(P)
    for name, uses_y, uses_sc, acc in ELTWISE:
        r = wref.eltwise(name, x, y if uses_y else None, o if acc else None, sc if uses_sc else None)
        g["eltwise"].append({"kernel": name, "n": n, "scalar": "%064x" % sc, "out": summary(r, (0, 1, n - 1))})
    for bit in (0, 31, 32, 200, 253):
        r = wref.eltwise("EltwiseBitDecompose", x, None, None, bit)
        g["eltwise"].append({"kernel": "EltwiseBitDecompose", "n": n, "bit": bit, "out": summary(r, (0, 1, n - 1))})

    # tests/webgpu/test_powmod.cpp:50-197 (the reference's own device KATs), same sizes and dispatch
    T = 8192
    one = np.zeros((T, 8), np.uint32)
    one[:, 0] = 1
    pw = {}
    pw["test_zero"] = digest(wref.powmod(1, np.zeros(T, np.uint32), np.zeros((T, 8), np.uint32), workgroups=T // 256))
    pw["test_one"] = digest(wref.powmod(1, np.ones(T, np.uint32), one, workgroups=T // 256))
    gen = wref.powmod(7, np.arange(T, dtype=np.uint32), one, workgroups=T // 256)
    pw["test_generator"] = summary(gen, (0, 1, T - 1))
    minus = np.tile(rand_elems(0, 1) * 0, (T, 1))
    for j in range(8):
        minus[:, j] = ((P - 1) >> (32 * j)) & 0xFFFFFFFF
    pw["test_minus"] = summary(wref.powmod(P - 1, np.arange(T, dtype=np.uint32) + (1 << 16), minus, workgroups=T // 256), (0, 1, T - 1))
    acc = gen.copy()
    for _ in range(9):
        acc = wref.powmod(7, np.arange(T, dtype=np.uint32), one, out=acc, add=True, workgroups=T // 256)
    pw["test_powmod_add"] = summary(acc, (0, 1, T - 1))
    g["powmod"] = pw

    idx = np.array(sorted(random.Random(61).sample(range(2048), 192)), np.uint32)
    src = rand_elems(62, 2048)
    g["sample_gather"] = {"seed_idx": 61, "seed_x": 62, "n": 2048, "out_sha256": digest(wref.sample_gather(src, idx))}

    path = os.path.join(HERE, "wgslref_vectors.json")
    with open(path, "w") as f:
        json.dump(g, f, indent=1, sort_keys=True)
    print("wrote", path)


if __name__ == "__main__":
    main()

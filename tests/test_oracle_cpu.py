"""CPU suite: the oracle against the golden vectors, against an independent pure-Python restatement,
and against structural properties.  No GPU needed."""
import hashlib
import json
import os
import random

import numpy as np
import pytest

from oracle import pyref

P = pyref.P
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "survey_vectors.json")))


def limbs_to_int(limbs):
    return sum(int(x, 16) << (32 * i) for i, x in enumerate(limbs))


def test_constants_against_reference_wgsl_values():
    """shader/bn254fr.wgsl.in:19-45 constants are consistent with p (R, J = p^-1, Barrett mu)"""
    assert limbs_to_int(GOLD["mont_R_limbs_le"]) == (1 << 256) % P
    assert limbs_to_int(GOLD["mont_J_limbs_le"]) == pow(P, -1, 1 << 256)
    assert limbs_to_int(GOLD["barrett_mu_limbs_le"]) == (1 << 508) // P
    # roots have order exactly 2^28 (src/bn254.cpp:41-43) and root2 = root1^(2^61-1) (:32-35)
    for r in (pyref.ROOT1, pyref.ROOT2):
        assert pow(r, 1 << 28, P) == 1 and pow(r, 1 << 27, P) == P - 1
    assert pow(pyref.ROOT1, (1 << 61) - 1, P) == pyref.ROOT2
    assert pyref.ROOT1 == pow(7, (P - 1) >> 28, P)


def test_omegas_golden(oracle):
    wk, w2k, wn = oracle.omegas(8192)
    assert (wk, w2k, wn) == (int(GOLD["omega_k"], 16), int(GOLD["omega_2k"], 16), int(GOLD["omega_n"], 16))
    assert pyref.omegas(8192) == (wk, w2k, wn)
    assert oracle.root1() == pyref.ROOT1 and oracle.root2() == pyref.ROOT2
    assert w2k * w2k % P == wk


def test_encode_golden_k8192(oracle):
    for name, pos in (("encode_delta0", 0), ("encode_delta1", 1)):
        row = np.zeros((8192, 8), np.uint32)
        row[pos, 0] = 1
        e = oracle.from_limbs(oracle.encode(row, 8192))
        g = GOLD[name]
        assert e[0] == int(g["e0"], 16) and e[1] == int(g["e1"], 16)
        if "e_last" in g:
            assert e[-1] == int(g["e_last"], 16)


def test_leaf_and_parent_golden(oracle):
    s = oracle.Sha(3)
    s.update(oracle.to_limbs([1, 1, 1]))
    d1 = s.final()
    assert d1[0].tobytes().hex() == GOLD["leaf_one_row_value_1"]
    s = oracle.Sha(1)
    for v in (1, P - 1, 0x0123456789abcdef):
        s.update(oracle.to_limbs([v]))
    d3 = s.final()
    assert d3[0].tobytes().hex() == GOLD["leaf_three_rows"]["digest"]
    nodes = oracle.merkle_build(np.stack([d1[0], d3[0]]))
    assert nodes[0].tobytes().hex() == GOLD["parent_of_the_two"]
    # independent restatement
    assert pyref.leaf_digest([1]).hex() == GOLD["leaf_one_row_value_1"]
    assert pyref.leaf_digest([1, P - 1, 0x0123456789abcdef]).hex() == GOLD["leaf_three_rows"]["digest"]


def test_field_ops_vs_python(oracle):
    rng = random.Random(11)
    edge = [0, 1, 2, P - 1, P - 2, (1 << 64) - 1, 1 << 128, P >> 1]
    a = edge * len(edge) + [rng.randrange(P) for _ in range(500)]
    b = [y for y in edge for _ in edge] + [rng.randrange(P) for _ in range(500)]
    A, B = oracle.to_limbs(a), oracle.to_limbs(b)
    f = oracle.from_limbs
    assert f(oracle.elt_add(A, B)) == [(x + y) % P for x, y in zip(a, b)]
    assert f(oracle.elt_sub(A, B)) == [(x - y) % P for x, y in zip(a, b)]
    assert f(oracle.elt_mul(A, B)) == [(x * y) % P for x, y in zip(a, b)]
    assert f(oracle.elt_fma(A, A, B)) == [(x + x * y) % P for x, y in zip(a, b)]
    assert f(oracle.elt_div(A, B)) == [(x * pow(y, -1, P)) % P if y else 0 for x, y in zip(a, b)]
    c = rng.randrange(P)
    rinv = pow(1 << 256, -1, P)
    assert f(oracle.elt_fma_const(B, A, c)) == [(y + c * x) % P for x, y in zip(a, b)]
    assert f(oracle.elt_add_const(A, c)) == [(x + c) % P for x in a]
    assert f(oracle.elt_sub_const(A, c)) == [(x - c) % P for x in a]
    assert f(oracle.elt_const_sub(A, c)) == [(c - x) % P for x in a]
    assert f(oracle.elt_mul_const(A, c)) == [(x * c) % P for x in a]
    assert f(oracle.elt_montmul_const(A, c)) == [(x * c * rinv) % P for x in a]
    assert f(oracle.elt_add_assign(B, A)) == [(x + y) % P for x, y in zip(a, b)]
    for bit in (0, 31, 32, 200, 253):
        assert f(oracle.elt_bit(A, bit)) == [(x >> bit) & 1 for x in a]
    exps = [rng.getrandbits(32) for _ in a]
    assert f(oracle.elt_powmod(A, exps, 7)) == [x * pow(7, e, P) % P for x, e in zip(a, exps)]


@pytest.mark.parametrize("logn", [1, 2, 3, 5, 8, 10])
def test_ntt_three_way(oracle, logn):
    """fast transform = naive DFT = Python definition; inverse round trip (BASELINE config 1 shape)"""
    n = 1 << logn
    w = pow(pyref.ROOT1, 1 << (28 - logn), P)
    x = oracle.synth(1, 0, 1, n)[0]
    xi = oracle.from_limbs(x)
    y = oracle.ntt(x, w)
    if logn <= 8:
        assert oracle.from_limbs(y) == pyref.dft(xi, w)
        assert np.array_equal(y, oracle.dft_naive(x, w))
    assert oracle.from_limbs(y) == pyref.ntt(xi, w)
    assert np.array_equal(oracle.ntt(y, w, inverse=True), x)
    assert np.array_equal(oracle.dft_naive(y, w, inverse=True), x) if logn <= 8 else True


def test_config1_ntt_4096_oracle_vs_naive(oracle):
    w = pow(pyref.ROOT1, 1 << 16, P)
    x = oracle.synth(1, 0, 1, 4096)[0]
    assert np.array_equal(oracle.ntt(x, w), oracle.dft_naive(x, w))


@pytest.mark.parametrize("k", [2, 4, 16, 64])
def test_encode_decode_vs_python(oracle, k):
    rng = random.Random(k)
    row = [rng.randrange(P) for _ in range(k)]
    cw = oracle.from_limbs(oracle.encode(oracle.to_limbs(row), k))
    assert cw == pyref.encode(row, k)
    # definition: codeword = evaluations on the w_n domain of the degree<k interpolant of the row
    wk, _, wn = pyref.omegas(k)
    coeffs = pyref.dft(row, wk, inverse=True)
    for j in (0, 1, 4 * k - 1):
        assert cw[j] == sum(c * pow(wn, i * j, P) for i, c in enumerate(coeffs)) % P
    r2 = [rng.randrange(P) for _ in range(2 * k)]
    assert oracle.from_limbs(oracle.encode_2k(oracle.to_limbs(r2), k)) == pyref.encode_2k(r2, k)
    dec = oracle.from_limbs(oracle.decode(oracle.to_limbs(cw), k))
    assert dec == pyref.decode(cw, k) and dec[:k] == row and not any(dec[k:])


def test_sha256_vs_hashlib(oracle):
    for m in (b"", b"abc", b"a" * 55, b"a" * 56, b"a" * 63, b"a" * 64, b"a" * 119, b"a" * 120, bytes(range(256)) * 5):
        assert oracle.sha256(m) == hashlib.sha256(m).digest()


def test_shani_and_portable_compress_agree(oracle):
    """the SHA-NI path (if the CPU has it) and the portable rounds give identical leaves"""
    import subprocess, sys
    code = ("import sys; sys.path.insert(0, %r); from oracle import lgo; import numpy as np; "
            "d, n, _ = lgo.encode_commit_synth(3, 9, 16); print(n[0].tobytes().hex())") % os.path.dirname(os.path.dirname(__file__))
    a = subprocess.check_output([sys.executable, "-c", code]).strip()
    b = subprocess.check_output([sys.executable, "-c", code], env=dict(os.environ, LGO_NO_SHANI="1")).strip()
    assert a == b and len(a) == 64


@pytest.mark.parametrize("k,R", [(4, 1), (4, 2), (16, 7), (64, 10)])
def test_encode_commit_vs_python(oracle, k, R):
    rows = oracle.synth(3, 0, R, k)
    dig, nodes, _ = oracle.encode_commit(rows, k)
    cws = [pyref.encode(oracle.from_limbs(rows[r]), k) for r in range(R)]
    leaves = [pyref.leaf_digest([cws[r][j] for r in range(R)]) for j in range(4 * k)]
    assert [bytes(d) for d in dig] == leaves
    assert [bytes(x) for x in nodes] == pyref.merkle(leaves)
    d2, n2, _ = oracle.encode_commit_synth(3, R, k)
    assert np.array_equal(d2, dig) and np.array_equal(n2, nodes)
    # streaming API gives the same leaves
    s = oracle.Sha(4 * k)
    for r in range(R):
        s.update(oracle.encode(rows[r], k))
    assert np.array_equal(s.final(), dig)


@pytest.mark.parametrize("nleaves", [1, 2, 3, 5, 8, 13])
def test_merkle_vs_python(oracle, nleaves):
    rng = np.random.default_rng(nleaves)
    leaves = rng.integers(0, 256, size=(nleaves, 32), dtype=np.uint8)
    nodes = oracle.merkle_build(leaves)
    assert [bytes(x) for x in nodes] == pyref.merkle([bytes(l) for l in leaves])


def test_synth_distribution_and_determinism(oracle):
    a = oracle.synth(3, 5, 4, 8)
    b = oracle.synth(3, 0, 9, 8)[5:]
    assert np.array_equal(a, b)                      # keyed by (seed,row,col), independent of batching
    vals = oracle.from_limbs(oracle.synth(9, 0, 64, 64).reshape(-1, 8))
    assert all(v < P for v in vals) and len(set(vals)) == len(vals)


@pytest.mark.parametrize("k", [8, 64, 1024])
def test_codeword_contains_the_message(oracle, k):
    """w_n^4 = w_k^(2^61-1) (src/bn254.cpp:36-43,51-64: root2 = root1^(2^61-1)) = w_k^(k-1): codeword position
    4m is message position (k-m) mod k.  The CUDA encoder relies on it (csrc/encode_kernels.cu)."""
    w_k, _, w_n = oracle.omegas(k)
    assert pow(w_n, 4, P) == pow(w_k, k - 1, P)
    row = oracle.synth(5, 3, 1, k)[0]
    e = oracle.encode(row, k)
    assert np.array_equal(e[0::4], row[(k - np.arange(k)) % k])

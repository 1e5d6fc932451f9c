"""Pins the CPU oracle (oracle/oracle.c) to the REFERENCE'S OWN SHADERS.

Two layers, both CPU-only:
  * golden: tests/golden/wgslref_vectors.json holds outputs of the reference's WGSL (bigint / bn254fr / kernels /
    sha256) transliterated to C++ and run in engine.cpp's dispatch order (oracle/_ref/libwgslref.so; generator
    tests/golden/make_wgslref_vectors.py).  The oracle must reproduce every one of them.  Always runs.
  * live: where the prebuilt library is present (build container, and the GPU box: it ships with the snapshot)
    the oracle is compared with it on fresh seeded inputs, and the transliterated shaders themselves are checked
    against the reference's own device KATs (tests/webgpu/test_powmod.cpp:50-197) and against Python big-int
    arithmetic, so a transliteration slip cannot hide.
"""
import hashlib
import json
import os
import random

import numpy as np
import pytest

from oracle import wref

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "wgslref_vectors.json")))
P = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
ROOT1 = pow(7, (P - 1) >> 28, P)
ROOT2 = pow(ROOT1, (1 << 61) - 1, P)

live = pytest.mark.skipif(not (wref.available() or os.path.isdir(os.path.join(wref.REF_ROOT, "shader"))),
                          reason="oracle/_ref/libwgslref.so not built (needs the reference tree)")


@pytest.fixture(scope="module")
def ref():
    wref.build()
    return wref


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


def check(arr, summ):
    for i, h in summ["elems"].items():
        assert hexel(arr[int(i)]) == h, f"element {i}"
    assert digest(arr) == summ["sha256"]


def omegas(k):
    return pow(ROOT1, (1 << 28) // k, P), pow(ROOT1, (1 << 28) // (2 * k), P), pow(ROOT2, (1 << 28) // (4 * k), P)


ORACLE_ELT = {
    "EltwiseAddMod": lambda o, x, y, out, sc: o.elt_add(x, y),
    "EltwiseAddAssignMod": lambda o, x, y, out, sc: o.elt_add_assign(out, x),
    "EltwiseAddConstantMod": lambda o, x, y, out, sc: o.elt_add_const(x, sc),
    "EltwiseSubMod": lambda o, x, y, out, sc: o.elt_sub(x, y),
    "EltwiseSubConstantMod": lambda o, x, y, out, sc: o.elt_sub_const(x, sc),
    "EltwiseConstantSubMod": lambda o, x, y, out, sc: o.elt_const_sub(x, sc),
    "EltwiseMultMod": lambda o, x, y, out, sc: o.elt_mul(x, y),
    "EltwiseMultConstantMod": lambda o, x, y, out, sc: o.elt_mul_const(x, sc),
    "EltwiseMontMultConstantMod": lambda o, x, y, out, sc: o.elt_montmul_const(x, sc),
    "EltwiseDivMod": lambda o, x, y, out, sc: o.elt_div(x, y),
    "EltwiseFMAMod": lambda o, x, y, out, sc: o.elt_fma(out, x, y),
    "EltwiseFMAConstantMod": lambda o, x, y, out, sc: o.elt_fma_const(out, x, sc),
}


# ---------------------------------------------------------------------------------------------------------
# golden layer: oracle.c against committed outputs of the reference's shaders
# ---------------------------------------------------------------------------------------------------------

def test_golden_constants():
    c = GOLD["constants"]
    assert int(c["p"], 16) == P and int(c["two_p"], 16) == 2 * P
    assert int(c["mont_r"], 16) == (1 << 256) % P
    assert int(c["mont_inv"], 16) == pow(P, -1, 1 << 256)
    assert int(c["barrett"], 16) == (1 << 508) // P


@pytest.mark.parametrize("case", GOLD["ntt"], ids=lambda c: f"2^{c['logn']}")
def test_golden_ntt(oracle, case):
    N = 1 << case["logn"]
    w = pow(ROOT1, (1 << 28) // N, P)
    x = rand_elems(case["seed"], N)
    check(oracle.ntt(x, w), case["forward"])
    check(oracle.ntt(x, w, inverse=True), case["inverse"])


@pytest.mark.parametrize("case", GOLD["encode"], ids=lambda c: f"k{c['k']}-seed{c['seed']}")
def test_golden_encode_decode(oracle, case):
    k = case["k"]
    assert oracle.omegas(k) == omegas(k)
    row = rand_elems(case["seed"], k, case["small"])
    e = oracle.encode(row, k)
    check(e, case["codeword"])
    check(oracle.decode(e, k), case["decoded"])


def test_golden_mask_row_path(oracle):
    c = GOLD["encode_2k"]
    check(oracle.encode_2k(rand_elems(c["seed"], 2 * c["k"]), c["k"]), c["codeword"])


@pytest.mark.parametrize("case", GOLD["sha"], ids=lambda c: f"n{c['ninst']}-r{c['rows']}")
def test_golden_sha_leaves(oracle, case):
    s = oracle.Sha(case["ninst"])
    s.init()
    for r in range(case["rows"]):
        s.update(rand_elems(case["seed"] * 1000 + r, case["ninst"]))
    d = s.final()
    assert d[0].tobytes().hex() == case["digest0"] and d[-1].tobytes().hex() == case["digest_last"]
    assert digest(d) == case["digests_sha256"]


def test_golden_eltwise(oracle):
    n = 300
    x, y, o = rand_elems(51, n), rand_elems(52, n), rand_elems(53, n)
    y[0] = 0
    y[0, 0] = 1
    x[1] = 0
    seen = set()
    for case in GOLD["eltwise"]:
        name = case["kernel"]
        seen.add(name)
        if name == "EltwiseBitDecompose":
            r = oracle.elt_bit(x, case["bit"])
        else:
            r = ORACLE_ELT[name](oracle, x, y, o, int(case["scalar"], 16))
        check(r, case["out"])
    assert seen == set(ORACLE_ELT) | {"EltwiseBitDecompose"}


def test_golden_powmod_kats(oracle):
    """the reference's own KAT shapes (tests/webgpu/test_powmod.cpp): 8192 lanes"""
    T = 8192
    pw = GOLD["powmod"]
    one = oracle.to_limbs([1] * T)
    ar = np.arange(T, dtype=np.uint32)
    assert digest(oracle.elt_powmod(np.zeros((T, 8), np.uint32), np.zeros(T, np.uint32), 1)) == pw["test_zero"]
    assert digest(oracle.elt_powmod(one, np.ones(T, np.uint32), 1)) == pw["test_one"]
    gen = oracle.elt_powmod(one, ar, 7)
    check(gen, pw["test_generator"])
    check(oracle.elt_powmod(oracle.to_limbs([P - 1] * T), ar + (1 << 16), P - 1), pw["test_minus"])
    acc = gen.copy()
    for _ in range(9):
        acc = oracle.elt_powmod(one, ar, 7, out=acc)
    check(acc, pw["test_powmod_add"])


def test_golden_sample_gather(oracle):
    c = GOLD["sample_gather"]
    idx = np.array(sorted(random.Random(c["seed_idx"]).sample(range(c["n"]), 192)), np.uint32)
    assert digest(oracle.gather(rand_elems(c["seed_x"], c["n"]), idx)) == c["out_sha256"]


# ---------------------------------------------------------------------------------------------------------
# live layer: the transliterated shaders checked on their own, then the oracle against them on fresh inputs
# ---------------------------------------------------------------------------------------------------------

@live
def test_live_field_primitives_against_bigint(ref):
    rnd = random.Random(101)
    rinv = pow(1 << 256, -1, P)
    edge = [0, 1, 2, P - 1, P - 2, (1 << 128) - 1, 1 << 128, (1 << 253) % P]
    pairs = [(a, b) for a in edge for b in edge] + [(rnd.randrange(P), rnd.randrange(P)) for _ in range(300)]
    for a, b in pairs:
        assert ref.montgomery_mul(a, b) == a * b * rinv % P          # bn254fr.wgsl.in:101-104, canonical
        lazy = ref.montgomery_mul(a, b, two_p=True)                  # :106-109, in [0, 2p)
        assert lazy < 2 * P and lazy % P == a * b * rinv % P
        assert ref.barrett_mul(a, b) == a * b % P                    # :113-124
    for a in [1, 2, P - 1] + [rnd.randrange(1, P) for _ in range(20)]:
        assert ref.invmod(a) * a % P == 1                            # :128-151


@live
def test_live_reference_powmod_kats(ref):
    """tests/webgpu/test_powmod.cpp:50-197 executed on the transliterated shaders, checked as the reference
    checks them (against mpz_powm_ui == Python pow)."""
    T = 8192
    one = np.zeros((T, 8), np.uint32)
    one[:, 0] = 1
    ar = np.arange(T, dtype=np.uint32)

    def ints(a):
        return [sum(int(a[i, j]) << (32 * j) for j in range(8)) for i in range(0, T, 37)]

    assert not ref.powmod(1, np.zeros(T, np.uint32), np.zeros((T, 8), np.uint32), workgroups=T // 256).any()      # test_zero
    assert np.array_equal(ref.powmod(1, np.ones(T, np.uint32), one, workgroups=T // 256), one)                   # test_one
    gen = ref.powmod(7, ar, one, workgroups=T // 256)                                                            # test_generator
    assert ints(gen) == [pow(7, i, P) for i in range(0, T, 37)]
    m1 = np.zeros((T, 8), np.uint32)
    for j in range(8):
        m1[:, j] = ((P - 1) >> (32 * j)) & 0xFFFFFFFF
    minus = ref.powmod(P - 1, ar + (1 << 16), m1, workgroups=T // 256)                                           # test_minus
    assert ints(minus) == [pow(P - 1, (1 << 16) + i + 1, P) for i in range(0, T, 37)]
    acc = gen.copy()                                                                                             # test_powmod_add
    for _ in range(9):
        acc = ref.powmod(7, ar, one, out=acc, add=True, workgroups=T // 256)
    assert ints(acc) == [pow(7, i, P) * 10 % P for i in range(0, T, 37)]


@live
def test_live_twiddle_tables_are_engine_tables(ref):
    """engine.cpp:1382-1503: stage i holds w^(j*N/2^i) * R, the shared table concatenates stages 1..9, N^-1 * R"""
    N = 2048
    w = pow(ROOT1, (1 << 28) // N, P)
    R = (1 << 256) % P
    for inverse, root in ((False, w), (True, pow(w, -1, P))):
        for stage in (1, 5, 11):
            t = ref.twiddles(N, w, inverse, stage)
            M = 1 << stage
            assert t.shape[0] == M // 2
            assert [int(hexel(e), 16) for e in t] == [pow(root, j * (N // M), P) * R % P for j in range(M // 2)]
        shared = ref.twiddles(N, w, inverse, 0)
        want = [pow(root, j * (N >> i), P) * R % P for i in range(1, 10) for j in range(1 << (i - 1))]
        assert [int(hexel(e), 16) for e in shared] == want
    assert ref.n_inv(N, w) == pow(N, -1, P) * R % P


@live
@pytest.mark.parametrize("logn", [9, 10, 12])
def test_live_ntt_is_the_plain_dft(ref, oracle, logn):
    """the reference's pass structure computes the standard DFT / inverse DFT, natural order, canonical"""
    N = 1 << logn
    w = pow(ROOT1, (1 << 28) // N, P)
    x = rand_elems(200 + logn, N)
    f = ref.ntt(x, w)
    assert np.array_equal(f, oracle.ntt(x, w))
    assert np.array_equal(ref.ntt(f, w, inverse=True), x)
    if logn == 9:
        assert np.array_equal(f, oracle.dft_naive(x, w))
    # a transform over the first N elements of a larger binding leaves the tail alone (4k-element codeword buffer)
    assert np.array_equal(ref.ntt(x, w, buf_elems=4 * N), f)


@live
@pytest.mark.parametrize("k,small", [(512, False), (512, True), (1024, False), (2048, False)])
def test_live_encode_decode(ref, oracle, k, small):
    wk, w2k, wn = omegas(k)
    row = rand_elems(300 + k + small, k, small)
    e = ref.encode(row, k, wk, wn)
    assert np.array_equal(e, oracle.encode(row, k))
    d = ref.decode(e, k, wk, w2k, wn)
    assert np.array_equal(d, oracle.decode(e, k))
    assert np.array_equal(d[:k], row)


@live
@pytest.mark.parametrize("ninst", [1, 5, 192, 256, 1024])
def test_live_sha_leaves(ref, oracle, ninst):
    for rows in (1, 2, 3, 6):
        a, b = ref.Sha(ninst), oracle.Sha(ninst)
        a.init()
        b.init()
        data = []
        for r in range(rows):
            row = rand_elems(400 + 10 * ninst + r, ninst, small=(r % 2 == 1))
            data.append(row)
            a.update(row)
            b.update(row)
        da, db = a.final(), b.final()
        assert np.array_equal(da, db)
        # and against hashlib on the serialisation the shader defines (sha256.wgsl:155-162,226-228)
        for j in (0, ninst - 1):
            msg = b"".join(int(row[j, i]).to_bytes(4, "big") for row in data for i in range(8))
            std = hashlib.sha256(msg).digest()
            stored = b"".join(std[4 * i:4 * i + 4][::-1] for i in range(8))
            assert da[j].tobytes() == stored


@live
def test_live_eltwise_every_kernel(ref, oracle):
    n = 517                                      # not a multiple of the workgroup size
    x, y, o = rand_elems(501, n), rand_elems(502, n), rand_elems(503, n)
    x[0] = 0
    y[1] = 0
    y[1, 0] = 1
    sc = random.Random(504).randrange(P)
    for name, fn in ORACLE_ELT.items():
        uses_y = name in ("EltwiseAddMod", "EltwiseSubMod", "EltwiseMultMod", "EltwiseDivMod", "EltwiseFMAMod")
        acc = name in ("EltwiseAddAssignMod", "EltwiseFMAMod", "EltwiseFMAConstantMod")
        r = ref.eltwise(name, x, y if uses_y else None, o if acc else None, sc)
        assert np.array_equal(r, fn(oracle, x, y, o, sc)), name
    for bit in (0, 1, 31, 32, 63, 64, 127, 128, 253, 255):
        assert np.array_equal(ref.eltwise("EltwiseBitDecompose", x, None, None, bit), oracle.elt_bit(x, bit)), bit


@live
def test_live_sample_gather(ref, oracle):
    src = rand_elems(601, 4096)
    idx = np.array(sorted(random.Random(602).sample(range(4096), 192)), np.uint32)
    assert np.array_equal(ref.sample_gather(src, idx), oracle.gather(src, idx))


@live
def test_golden_file_is_what_the_library_produces_today(ref):
    """spot re-derivation of committed vectors from the live library (guards a stale JSON)"""
    c = GOLD["encode"][0]
    wk, w2k, wn = omegas(c["k"])
    check(ref.encode(rand_elems(c["seed"], c["k"], c["small"]), c["k"], wk, wn), c["codeword"])
    s = GOLD["sha"][3]
    h = ref.Sha(s["ninst"])
    h.init()
    for r in range(s["rows"]):
        h.update(rand_elems(s["seed"] * 1000 + r, s["ninst"]))
    assert digest(h.final()) == s["digests_sha256"]

"""GPU parity: CUDA path (through the C ABI) vs the CPU oracle, bit-exact.  Needs a B200."""
import ctypes as C
import json
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

P = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "survey_vectors.json")))


def rand_elems(oracle, seed, n, ncols=1):
    return oracle.synth(seed, 0, n, ncols).reshape(n * ncols, 8)


def edge_elems(lgr):
    vals = [0, 1, 2, P - 1, P - 2, (1 << 32) - 1, 1 << 32, (1 << 64) - 1, 1 << 128, (1 << 253) + 12345, P >> 1, 0x0123456789abcdef]
    return lgr.ints_to_array(vals)


# ---------------------------------------------------------------- transforms
@pytest.mark.parametrize("logn", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 15, 16])
def test_ntt_pow2_forward_inverse(lgr, oracle, executor_factory, logn):
    ex = executor_factory(256)
    N = 1 << logn
    w = lgr.root_of_unity(logn)
    batch = 3 if logn <= 12 else 1
    x = rand_elems(oracle, 100 + logn, batch * N).reshape(batch, N, 8)
    buf = ex.make_device_buffer(batch * N * 32)
    ex.write_buffer(buf, x)
    ex.ntt_pow2(buf, logn, batch, w, inverse=False)
    got = ex.read_elements(buf).reshape(batch, N, 8)
    want = oracle.ntt_batch(x, w, inverse=False)
    assert np.array_equal(got, want)
    ex.ntt_pow2(buf, logn, batch, w, inverse=True)
    assert np.array_equal(ex.read_elements(buf).reshape(batch, N, 8), x)
    # inverse alone against the oracle
    ex.write_buffer(buf, x)
    ex.ntt_pow2(buf, logn, batch, w, inverse=True)
    assert np.array_equal(ex.read_elements(buf).reshape(batch, N, 8), oracle.ntt_batch(x, w, inverse=True))


def test_config1_ntt_4096_three_way(lgr, oracle, executor_factory):
    """BASELINE config 1: 2^12-point forward NTT, w = root1^(2^16), seed 1: GPU = oracle = naive DFT"""
    ex = executor_factory(256)
    N = 4096
    w = lgr.root_of_unity(12)
    assert w == pow(lgr.ROOT1, 1 << 16, P)
    x = rand_elems(oracle, 1, N)
    buf = ex.make_device_buffer(N * 32)
    ex.write_buffer(buf, x)
    ex.ntt_pow2(buf, 12, 1, w)
    got = ex.read_elements(buf)
    assert np.array_equal(got, oracle.ntt(x, w))
    assert np.array_equal(got, oracle.dft_naive(x, w))


def test_ntt_rejects_bad_root(lgr, executor_factory):
    ex = executor_factory(256)
    buf = ex.make_device_buffer(16 * 32)
    with pytest.raises(lgr.LgrError):
        ex.ntt_pow2(buf, 4, 1, 12345)


@pytest.mark.parametrize("k", [8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192])
def test_encode_decode_executor_api(lgr, oracle, executor_factory, k):
    """encode_ntt_device / decode_ntt_device / ntt_*_{k,2k,n} exactly as nonbatch_context.hpp drives them"""
    ex = executor_factory(k)
    n = 4 * k
    row = rand_elems(oracle, 7 + k, k)
    dev = ex.make_codeword_buffer()
    bind = ex.bind_ntt(dev)
    limbs = np.zeros((2 * k, 8), np.uint32)     # nonbatch_context.hpp:415: scratch of 2k elements, upper half zero
    limbs[:k] = row
    ex.write_buffer_clear(dev, limbs)
    ex.encode_ntt_device(bind)
    cw = ex.read_elements(dev)
    want = oracle.encode(row, k)
    assert np.array_equal(cw, want)
    # decode gives back the message on [0,k) and zero high coefficients (src/webgpu_prover.cpp:465-471)
    ex.decode_ntt_device(bind)
    dec = ex.read_elements(dev)
    assert np.array_equal(dec, oracle.decode(want, k))
    assert np.array_equal(dec[:k], row)
    assert not dec[k:].any()
    # mask-row path: ntt_inverse_2k + ntt_forward_n (nonbatch_context.hpp:482-494)
    m2 = rand_elems(oracle, 9 + k, 2 * k)
    ex.write_buffer_clear(dev, m2)
    ex.ntt_inverse_2k(bind)
    ex.ntt_forward_n(bind)
    assert np.array_equal(ex.read_elements(dev), oracle.encode_2k(m2, k))
    # remaining size selectors round-trip
    for fwd, inv, size, wsel in ((ex.ntt_forward_k, ex.ntt_inverse_k, k, 0), (ex.ntt_forward_2k, ex.ntt_inverse_2k, 2 * k, 1), (ex.ntt_forward_n, ex.ntt_inverse_n, n, 2)):
        x = rand_elems(oracle, 11 + size, n)
        ex.write_buffer(dev, x)
        fwd(bind)
        y = ex.read_elements(dev)
        w = lgr.generate_omegas(k, n)[wsel]
        assert np.array_equal(y[:size], oracle.ntt(x[:size], w))
        assert np.array_equal(y[size:], x[size:])
        inv(bind)
        assert np.array_equal(ex.read_elements(dev), x)


def test_encode_golden_vectors_k8192(lgr, oracle, executor_factory):
    """SURVEY 8c derived vectors at the reference's default geometry"""
    k, n = 8192, 32768
    ex = executor_factory(k)
    dev = ex.make_codeword_buffer()
    for name, pos in (("encode_delta0", 0), ("encode_delta1", 1)):
        row = np.zeros((k, 8), np.uint32)
        row[pos, 0] = 1
        ex.write_buffer_clear(dev, row)
        ex.encode_ntt_device(ex.bind_ntt(dev))
        e = lgr.array_to_ints(ex.read_elements(dev))
        g = GOLD[name]
        assert e[0] == int(g["e0"], 16) and e[1] == int(g["e1"], 16)
        if "e_last" in g:
            assert e[-1] == int(g["e_last"], 16)


@pytest.mark.parametrize("k,R", [(8, 5), (64, 33), (256, 64), (1024, 9), (4096, 3)])
def test_encode_rows_batch(lgr, oracle, executor_factory, k, R):
    ex = executor_factory(k)
    n = 4 * k
    rows = oracle.synth(21, 0, R, k)
    src = ex.make_device_buffer(R * k * 32)
    dst = ex.make_device_buffer(R * n * 32)
    ex.write_buffer(src, rows)
    ex.encode_rows(src, R, dst)
    got = ex.read_elements(dst).reshape(R, n, 8)
    for r in range(R):
        assert np.array_equal(got[r], oracle.encode(rows[r], k)), r
    # linearity (size-independent property): enc(a) + enc(b) = enc(a + b)
    a, b = rows[0], rows[1]
    s = oracle.elt_add(a, b)
    assert np.array_equal(oracle.elt_add(got[0], got[1]), oracle.encode(s, k))


@pytest.mark.parametrize("k,R", [(256, 4096), (2048, 64), (8192, 16)])
def test_codeword_contains_the_message(lgr, oracle, executor_factory, k, R):
    """size-independent property: w_n^4 = w_k^(k-1) for the reference's roots (src/bn254.cpp:36-43,51-64), so
    codeword position 4m holds message position (k-m) mod k; the encoder copies that coset instead of
    computing it, the remaining positions are checked against the oracle on a few rows"""
    ex = executor_factory(k)
    n = 4 * k
    src = ex.make_device_buffer(R * k * 32)
    dst = ex.make_device_buffer(R * n * 32)
    ex.synth(src, 11, 0, R, k)
    ex.encode_rows(src, R, dst)
    rows = ex.read_elements(src).reshape(R, k, 8)
    got = ex.read_elements(dst).reshape(R, n, 8)
    perm = (k - np.arange(k)) % k
    assert np.array_equal(got[:, 0::4], rows[:, perm])
    for r in (0, R // 2, R - 1):
        assert np.array_equal(got[r], oracle.encode(rows[r], k)), r


@pytest.mark.parametrize("k,c", [(64, 1), (64, 3), (256, 77), (4096, 5)])
def test_encode_with_other_roots(lgr, oracle, k, c):
    """lgr_create takes the roots as arguments (include/wgpu.hpp:75-82): for any w_n the shortcut exponent is
    found by search (w_n^4 = w_k^c); results must still equal iNTT_k -> zero pad -> NTT_n"""
    n = 4 * k
    w_k, w_2k, _ = lgr.generate_omegas(k, n)
    w_4k = lgr.root_of_unity(n.bit_length() - 1)
    w_n = pow(w_4k, c, lgr.P)                                       # w_n^4 = w_k^c
    assert pow(w_n, 4, lgr.P) == pow(w_k, c, lgr.P)
    ex = lgr.Executor(0)
    ex.webgpu_init(0, "")
    ex.ntt_init(max(k - 192, 1), k, n, root_k=w_k, root_2k=w_2k, root_n=w_n)
    try:
        R = 5
        rows = oracle.synth(13, 0, R, k)
        src = ex.make_device_buffer(R * k * 32)
        dst = ex.make_device_buffer(R * n * 32)
        ex.write_buffer(src, rows)
        ex.encode_rows(src, R, dst)
        got = ex.read_elements(dst).reshape(R, n, 8)
        for r in range(R):
            coef = np.zeros((n, 8), np.uint32)
            coef[:k] = oracle.ntt(rows[r], w_k, inverse=True)
            assert np.array_equal(got[r], oracle.ntt(coef, w_n)), r
        # in place, one row (encode_ntt_device)
        buf = ex.make_codeword_buffer()
        ex.write_buffer_clear(buf, rows[0])
        ex.encode_ntt_device(ex.bind_ntt(buf))
        assert np.array_equal(ex.read_elements(buf), got[0])
    finally:
        ex.close()


# ---------------------------------------------------------------- hashing
def test_sha_leaf_golden_and_streaming(lgr, oracle, executor_factory):
    ex = executor_factory(256)
    ninst = 1024
    ex.sha256_init(ninst)
    ctx = ex.make_device_buffer(ex.sha256_context_bytes(ninst))
    assert ex.sha256_context_bytes(ninst) <= ninst * 75 * 4        # fits the reference's sha256_context (wgpu.hpp:63-68)
    dig = ex.make_device_buffer(ninst * 32)
    cb = ex.bind_sha256_context(ctx, dig)
    dev = ex.make_codeword_buffer()
    sb = ex.bind_sha256_buffer(dev)
    # golden: one row holding the value 1; three rows (1, p-1, 0x0123456789abcdef)
    ex.sha256_digest_init(cb)
    ex.write_buffer(dev, lgr.ints_to_array([1] * ninst))
    ex.sha256_digest_update(cb, sb)
    ex.sha256_digest_final(cb)
    d = ex.copy_to_host(dig, np.uint8).reshape(ninst, 32)
    assert d[0].tobytes().hex() == GOLD["leaf_one_row_value_1"] and (d == d[0]).all()
    ex.sha256_digest_init(cb)
    for v in (1, P - 1, 0x0123456789abcdef):
        ex.write_buffer(dev, lgr.ints_to_array([v] * ninst))
        ex.sha256_digest_update(cb, sb)
    ex.sha256_digest_final(cb)
    d3 = ex.copy_to_host(dig, np.uint8).reshape(ninst, 32)
    assert d3[5].tobytes().hex() == GOLD["leaf_three_rows"]["digest"]
    # empty stream (0 rows) = SHA-256("") with swapped words
    ex.sha256_digest_init(cb)
    ex.sha256_digest_final(cb)
    s = oracle.Sha(ninst)
    assert np.array_equal(ex.copy_to_host(dig, np.uint8).reshape(ninst, 32), s.final())
    # random rows, ragged update sizes (1 row, then tiles of odd/even length), final is idempotent
    R = 23
    rows = oracle.synth(5, 0, R, ninst)
    tile = ex.make_device_buffer(R * ninst * 32)
    ex.write_buffer(tile, rows)
    ex.sha256_digest_init(cb)
    s = oracle.Sha(ninst)
    pos = 0
    for chunk in (1, 2, 3, 1, 4, 5, 7):
        ex.sha256_digest_update_rows(cb, tile.slice(pos * ninst * 32), chunk)
        for r in range(pos, pos + chunk):
            s.update(rows[r])
        pos += chunk
        ex.sha256_digest_final(cb)
        assert np.array_equal(ex.copy_to_host(dig, np.uint8).reshape(ninst, 32), s.final()), pos
    assert pos == R


@pytest.mark.parametrize("nleaves", [1, 2, 3, 8, 192, 1024, 1025, 4096, 32768])
def test_merkle_build(lgr, oracle, executor_factory, nleaves):
    ex = executor_factory(256)
    rng = np.random.default_rng(nleaves)
    leaves = rng.integers(0, 256, size=(nleaves, 32), dtype=np.uint8)
    d = ex.make_device_buffer(nleaves * 32)
    ex.write_buffer(d, leaves)
    nn = ex.merkle_node_count(nleaves)
    nodes = ex.make_device_buffer(nn * 32)
    ex.merkle_build(d, nleaves, nodes)
    got = ex.copy_to_host(nodes, np.uint8).reshape(nn, 32)
    assert np.array_equal(got, oracle.merkle_build(leaves))


def test_merkle_golden_parent(lgr, executor_factory):
    ex = executor_factory(256)
    leaves = np.frombuffer(bytes.fromhex(GOLD["leaf_one_row_value_1"] + GOLD["leaf_three_rows"]["digest"]), np.uint8).reshape(2, 32)
    d = ex.make_device_buffer(64)
    ex.write_buffer(d, leaves)
    nodes = ex.make_device_buffer(3 * 32)
    ex.merkle_build(d, 2, nodes)
    assert ex.copy_to_host(nodes, np.uint8)[:32].tobytes().hex() == GOLD["parent_of_the_two"]


@pytest.mark.parametrize("k,R", [(8, 1), (8, 6), (64, 17), (256, 100), (256, 4100), (1024, 5), (4096, 4)])
def test_encode_commit_pipeline(lgr, oracle, executor_factory, k, R):
    """stage-1 commit (nonbatch_context.hpp:445-451,555-558 + merkle_tree.hpp:343-375)"""
    ex = executor_factory(k)
    n = 4 * k
    rows = oracle.synth(3, 0, R, k)
    src = ex.make_device_buffer(R * k * 32)
    ex.write_buffer(src, rows)
    dig = ex.make_device_buffer(n * 32)
    nodes = ex.make_device_buffer((2 * n - 1) * 32)
    ex.encode_commit(src, R, dig, nodes)
    want_d, want_n, _ = oracle.encode_commit(rows, k)
    assert np.array_equal(ex.copy_to_host(dig, np.uint8).reshape(n, 32), want_d)
    assert np.array_equal(ex.copy_to_host(nodes, np.uint8).reshape(2 * n - 1, 32), want_n)
    # run again on the same context: identical result (buffers / contexts are re-initialised)
    ex.encode_commit(src, R, dig, nodes)
    assert np.array_equal(ex.copy_to_host(nodes, np.uint8).reshape(2 * n - 1, 32), want_n)


def test_synth_matches_oracle(lgr, oracle, executor_factory):
    ex = executor_factory(256)
    buf = ex.make_device_buffer(37 * 256 * 32)
    ex.synth(buf, 3, 1000, 37, 256)
    assert np.array_equal(ex.read_elements(buf).reshape(37, 256, 8), oracle.synth(3, 1000, 37, 256))


# ---------------------------------------------------------------- element-wise / combiners
def test_eltwise_all_ops(lgr, oracle, executor_factory):
    ex = executor_factory(256)
    n = 1024
    edge = edge_elems(lgr)
    x = rand_elems(oracle, 31, n); y = rand_elems(oracle, 32, n); z = rand_elems(oracle, 33, n)
    ne = len(edge)
    x[:ne] = edge; y[:ne] = edge[::-1]
    x[ne:2 * ne] = edge; y[ne:2 * ne] = edge            # squares / x - x
    bx, by, bz, bo = (ex.make_codeword_buffer() for _ in range(4))
    ex.write_buffer(bx, x); ex.write_buffer(by, y); ex.write_buffer(bz, z)
    b3 = ex.bind_eltwise3(bx, by, bo)
    b2 = ex.bind_eltwise2(bx, bo)
    c = 0x2f0b1c0ffee123456789abcdef0123456789abcdef0123456789abcdef012345 % P

    def out():
        return ex.read_elements(bo)

    ex.EltwiseAddMod(b3); assert np.array_equal(out(), oracle.elt_add(x, y))
    ex.EltwiseSubMod(b3); assert np.array_equal(out(), oracle.elt_sub(x, y))
    ex.EltwiseMultMod(b3); assert np.array_equal(out(), oracle.elt_mul(x, y))
    ex.write_buffer(bo, z); ex.EltwiseFMAMod(b3); assert np.array_equal(out(), oracle.elt_fma(z, x, y))
    for cc in (c, 0, 1, P - 1):
        ex.write_buffer(bo, z); ex.EltwiseFMAMod(b2, cc); assert np.array_equal(out(), oracle.elt_fma_const(z, x, cc))
        ex.EltwiseAddMod(b2, cc); assert np.array_equal(out(), oracle.elt_add_const(x, cc))
        ex.EltwiseSubConstMod(b2, cc); assert np.array_equal(out(), oracle.elt_sub_const(x, cc))
        ex.EltwiseConstSubMod(b2, cc); assert np.array_equal(out(), oracle.elt_const_sub(x, cc))
        ex.EltwiseMultMod(b2, cc); assert np.array_equal(out(), oracle.elt_mul_const(x, cc))
        ex.EltwiseMontMultMod(b2, cc); assert np.array_equal(out(), oracle.elt_montmul_const(x, cc))
    ex.write_buffer(bo, z); ex.EltwiseAddAssignMod(b2); assert np.array_equal(out(), oracle.elt_add_assign(z, x))
    for bit in (0, 1, 31, 32, 63, 100, 253, 255):
        ex.EltwiseBitDecompose(b2, bit); assert np.array_equal(out(), oracle.elt_bit(x, bit))
    # in-place variants used by the reference (bind_eltwise3(tmp1, z, tmp2) etc. alias freely)
    ex.write_buffer(bo, z)
    ex.EltwiseAddMod(ex.bind_eltwise3(bo, bx, bo)); assert np.array_equal(out(), oracle.elt_add(z, x))
    # scalar must be reduced
    with pytest.raises(lgr.LgrError):
        ex.EltwiseFMAMod(b2, P)
    # fused check_quadratic = Mult, Sub, FMA-const of the reference (nonbatch_context.hpp:771-780)
    ex.write_buffer(bo, z)
    ex.quadratic_fused(bx, by, bz, bo, c)
    want = oracle.elt_fma_const(z, oracle.elt_sub(oracle.elt_mul(x, y), z), c)
    assert np.array_equal(out(), want)
    # element offsets (vbn254fr arena addressing, vbn254fr.hpp:56-66)
    ex.EltwiseAddMod(b3, element_offsets=(8, 16, 24), count=100)
    assert np.array_equal(out()[24:124], oracle.elt_add(x[8:108], y[16:116]))


def test_eltwise_div(lgr, oracle, executor_factory):
    ex = executor_factory(256)
    n = 1024
    x = rand_elems(oracle, 41, n); y = rand_elems(oracle, 42, n)
    y[:3] = lgr.ints_to_array([0, 1, P - 1])
    bx, by, bo = (ex.make_codeword_buffer() for _ in range(3))
    ex.write_buffer(bx, x); ex.write_buffer(by, y)
    ex.EltwiseDivMod(ex.bind_eltwise3(bx, by, bo))
    assert np.array_equal(ex.read_elements(bo), oracle.elt_div(x, y))


@pytest.mark.parametrize("base", [1, 7, P - 1, 0x1234567890abcdef1234567890abcdef])
def test_powmod_kat(lgr, oracle, executor_factory, base):
    """the reference's only device KAT: tests/webgpu/test_powmod.cpp:50-197 -- coeff*base^exp vs
    mpz_powm_ui for 8192 lanes, bases 1, 7, p-1 (here checked against Python's pow)"""
    ex = executor_factory(256)
    n = 8192
    rng = random.Random(base & 0xffff)
    exps = [0, 1, 2, 0xFFFFFFFF, 0x80000000] + [rng.getrandbits(32) for _ in range(n - 5)]
    coeffs = [1, 0, P - 1] + [rng.randrange(P) for _ in range(n - 3)]
    be = ex.make_device_buffer(n * 4); bc = ex.make_device_buffer(n * 32); bo = ex.make_device_buffer(n * 32)
    ex.write_buffer(be, np.array(exps, np.uint32)); ex.write_buffer(bc, lgr.ints_to_array(coeffs))
    ex.powmod_init(32); ex.powmod_set_base(base)
    bind = ex.bind_powmod(be, bc, bo)
    ex.EltwisePowMod(bind)
    want = [c * pow(base, e, P) % P for c, e in zip(coeffs, exps)]
    assert lgr.array_to_ints(ex.read_elements(bo)) == want
    ex.EltwisePowAddMod(bind)
    assert lgr.array_to_ints(ex.read_elements(bo)) == [2 * w % P for w in want]
    assert np.array_equal(ex.read_elements(bo), oracle.elt_powmod(lgr.ints_to_array(coeffs), exps, base, out=lgr.ints_to_array(want)))


def test_sample_gather(lgr, oracle, executor_factory):
    ex = executor_factory(256)
    n = 1024
    rng = random.Random(9)
    idx = sorted(rng.sample(range(n), 192))
    ex.sampling_init(idx)
    x = rand_elems(oracle, 51, n)
    bx = ex.make_codeword_buffer(); ex.write_buffer(bx, x)
    stage = ex.make_device_buffer(256 * 192 * 32)             # nonbatch_context.hpp:906-909
    bind = ex.bind_sampling(bx, stage)
    ex.sample_gather(bind, 0); ex.sample_gather(bind, 3)
    got = ex.read_elements(stage).reshape(256, 192, 8)
    want = oracle.gather(x, idx)
    assert np.array_equal(got[0], want) and np.array_equal(got[3], want) and not got[1].any()


@pytest.mark.parametrize("k,T", [(64, 1), (64, 31), (256, 64), (256, 65), (256, 200), (1024, 130), (8, 64 * 70 + 3)])
def test_tile_combiners(lgr, oracle, executor_factory, k, T):
    """check_code / check_linear / check_quadratic over a resident tile == the reference's per-row schedule"""
    ex = executor_factory(k)
    n = 4 * k
    a = oracle.synth(61, 0, T, n); b = oracle.synth(62, 0, T, n)
    acc0 = rand_elems(oracle, 63, n)
    rng = random.Random(T)
    rs = [0, 1, P - 1][: min(3, T)] + [rng.randrange(P) for _ in range(max(0, T - 3))]
    ta = ex.make_device_buffer(T * n * 32); tb = ex.make_device_buffer(T * n * 32); acc = ex.make_codeword_buffer()
    ex.write_buffer(ta, a); ex.write_buffer(tb, b)
    ex.write_buffer(acc, acc0)
    ex.combine_code(ta, T, rs, acc)
    want = acc0.copy()
    for t in range(T):
        want = oracle.elt_fma_const(want, a[t], rs[t])
    assert np.array_equal(ex.read_elements(acc), want)
    ex.write_buffer(acc, acc0)
    ex.combine_linear(ta, tb, T, acc)
    want = acc0.copy()
    for t in range(T):
        want = oracle.elt_fma(want, a[t], b[t])
    assert np.array_equal(ex.read_elements(acc), want)
    # quadratic: worst-case operands in the first rows (p-1 everywhere) exercise the wide accumulator bound
    z = oracle.synth(64, 0, T, n)
    a2, b2 = a.copy(), b.copy()
    a2[0] = lgr.ints_to_array([P - 1] * n); b2[0] = a2[0]; z[0] = 0
    tz = ex.make_device_buffer(T * n * 32)
    ex.write_buffer(ta, a2); ex.write_buffer(tb, b2); ex.write_buffer(tz, z)
    ex.write_buffer(acc, acc0)
    ex.combine_quad(ta, tb, tz, T, rs, acc)
    want = acc0.copy()
    for t in range(T):
        want = oracle.elt_fma_const(want, oracle.elt_sub(oracle.elt_mul(a2[t], b2[t]), z[t]), rs[t])
    assert np.array_equal(ex.read_elements(acc), want)


# ---------------------------------------------------------------- the reference's stage schedule
@pytest.mark.parametrize("k", [512, 8192])
def test_stage1_stage2_schedule_like_nonbatch_context(lgr, oracle, executor_factory, k):
    """Drive the executor exactly as nonbatch_stage1_context / nonbatch_stage2_context do for one
    linear row, one quadratic triple and the three mask rows (BASELINE config 4's row schedule at the
    reference's default k = 8192: 7 encodes per stage), and compare root / code / linear / quad with
    the oracle."""
    n = 4 * k; l = k - 192
    ex = executor_factory(k)
    rng = random.Random(4)

    def row(seed, size=k):
        return rand_elems(oracle, seed, size)

    lin, lin_r = row(70), row(71)
    X, Y = row(72), row(73)
    Z = oracle.elt_mul(X, Y)
    xr, yr, zr = row(74), row(75), row(76)
    mask_c = row(77); mask_l = row(78, 2 * k); mask_q = row(79, 2 * k)
    r_code = [rng.randrange(P) for _ in range(4)]; r_quad = rng.randrange(P)

    # ---- stage 1 (nonbatch_context.hpp:445-494,555-558)
    ex.sha256_init(n)
    dx, dy, dz = (ex.make_codeword_buffer() for _ in range(3))
    sctx = ex.make_device_buffer(ex.sha256_context_bytes(n)); sdig = ex.make_device_buffer(n * 32)
    cb = ex.bind_sha256_context(sctx, sdig)
    ex.sha256_digest_init(cb)
    limbs = np.zeros((2 * k, 8), np.uint32)

    def upload(dev, vals):
        limbs[:] = 0; limbs[: vals.shape[0]] = vals
        ex.write_buffer_clear(dev, limbs)

    for dev, vals in ((dx, lin), (dx, X), (dy, Y), (dz, Z)):
        upload(dev, vals); ex.encode_ntt_device(ex.bind_ntt(dev)); ex.sha256_digest_update(cb, ex.bind_sha256_buffer(dev))
    upload(dx, mask_c); ex.encode_ntt_device(ex.bind_ntt(dx)); ex.sha256_digest_update(cb, ex.bind_sha256_buffer(dx))
    for dev, vals in ((dy, mask_l), (dz, mask_q)):
        upload(dev, vals); ex.ntt_inverse_2k(ex.bind_ntt(dev)); ex.ntt_forward_n(ex.bind_ntt(dev)); ex.sha256_digest_update(cb, ex.bind_sha256_buffer(dev))
    ex.sha256_digest_final(cb)
    nodes = ex.make_device_buffer((2 * n - 1) * 32)
    ex.merkle_build(sdig, n, nodes)
    root = ex.copy_to_host(nodes, np.uint8)[:32]

    enc = lambda v: oracle.encode(v, k)
    enc2 = lambda v: oracle.encode_2k(v, k)
    cws = [enc(lin), enc(X), enc(Y), enc(Z), enc(mask_c), enc2(mask_l), enc2(mask_q)]
    s = oracle.Sha(n)
    for cw in cws:
        s.update(cw)
    assert np.array_equal(root, oracle.merkle_build(s.final())[0])

    # ---- stage 2 (nonbatch_context.hpp:654-780)
    code, linear, quad, tmp1, tmp2 = (ex.make_codeword_buffer() for _ in range(5))
    drx, dry, drz = (ex.make_codeword_buffer() for _ in range(3))
    upload(dx, lin); upload(drx, lin_r)
    ex.encode_ntt_device(ex.bind_ntt(dx)); ex.encode_ntt_device(ex.bind_ntt(drx))
    ex.EltwiseFMAMod(ex.bind_eltwise2(dx, code), r_code[0])
    ex.EltwiseFMAMod(ex.bind_eltwise3(dx, drx, linear))
    for dev, vals in ((dx, X), (drx, xr), (dy, Y), (dry, yr), (dz, Z), (drz, zr)):
        upload(dev, vals); ex.encode_ntt_device(ex.bind_ntt(dev))
    for dev, r in ((dx, r_code[1]), (dy, r_code[2]), (dz, r_code[3])):
        ex.EltwiseFMAMod(ex.bind_eltwise2(dev, code), r)
    for dev, rnd in ((dx, drx), (dy, dry), (dz, drz)):
        ex.EltwiseFMAMod(ex.bind_eltwise3(dev, rnd, linear))
    ex.EltwiseMultMod(ex.bind_eltwise3(dx, dy, tmp1)); ex.EltwiseSubMod(ex.bind_eltwise3(tmp1, dz, tmp2)); ex.EltwiseFMAMod(ex.bind_eltwise2(tmp2, quad), r_quad)
    upload(dx, mask_c); ex.encode_ntt_device(ex.bind_ntt(dx)); ex.EltwiseAddAssignMod(ex.bind_eltwise2(dx, code))
    upload(dy, mask_l); ex.ntt_inverse_2k(ex.bind_ntt(dy)); ex.ntt_forward_n(ex.bind_ntt(dy)); ex.EltwiseAddAssignMod(ex.bind_eltwise2(dy, linear))
    upload(dz, mask_q); ex.ntt_inverse_2k(ex.bind_ntt(dz)); ex.ntt_forward_n(ex.bind_ntt(dz)); ex.EltwiseAddAssignMod(ex.bind_eltwise2(dz, quad))

    zero = np.zeros((n, 8), np.uint32)
    w_code = zero
    for cw, r in zip(cws[:4], r_code):
        w_code = oracle.elt_fma_const(w_code, cw, r)
    w_code = oracle.elt_add_assign(w_code, cws[4])
    w_lin = zero
    for cw, rnd in zip(cws[:4], (lin_r, xr, yr, zr)):
        w_lin = oracle.elt_fma(w_lin, cw, enc(rnd))
    w_lin = oracle.elt_add_assign(w_lin, cws[5])
    w_quad = oracle.elt_fma_const(zero, oracle.elt_sub(oracle.elt_mul(cws[1], cws[2]), cws[3]), r_quad)
    w_quad = oracle.elt_add_assign(w_quad, cws[6])
    assert np.array_equal(ex.read_elements(code), w_code)
    assert np.array_equal(ex.read_elements(linear), w_lin)
    assert np.array_equal(ex.read_elements(quad), w_quad)
    # prover self-checks (src/webgpu_prover.cpp:465-471): code decodes to degree < k; the quadratic
    # test polynomial (degree < 2k) vanishes on the first k points before the mask is added
    ex.decode_ntt_device(ex.bind_ntt(code))
    assert not ex.read_elements(code)[k:].any()


def test_cpp_cuda_executor_drop_in(lgr, oracle, tmp_path):
    """the C++ adapter (host/cuda_executor.hpp) driven like nonbatch_stage1/2_context, vs the oracle"""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "cpp", "test_executor")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.join(root, "tests", "cpp"), "-s"])
    k, nrows = 512, 7
    n = 4 * k
    rows = oracle.synth(13, 0, nrows, k)
    rng = random.Random(13)
    scal = [rng.randrange(P) for _ in range(nrows)]
    wk, w2k, wn = lgr.generate_omegas(k, n)
    hdr = lgr.ints_to_array([P, wk, w2k, wn])
    cnt = np.zeros((1, 8), np.uint32); cnt[0, 0] = nrows
    blob = np.concatenate([hdr, cnt, lgr.ints_to_array(scal), rows.reshape(-1, 8)])
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    blob.tofile(fin)
    out = subprocess.check_output([exe, str(k), str(fin), str(fout)], stderr=subprocess.STDOUT, timeout=300)
    assert b"ok" in out
    data = np.fromfile(fout, np.uint8)
    dig = data[: n * 32].reshape(n, 32)
    code = data[n * 32: 2 * n * 32].view(np.uint32).reshape(n, 8)
    quad = data[2 * n * 32:].view(np.uint32).reshape(n, 8)
    want_d, _, _ = oracle.encode_commit(rows, k)
    assert np.array_equal(dig, want_d)
    cws = [oracle.encode(rows[r], k) for r in range(nrows)]
    wc = np.zeros((n, 8), np.uint32)
    for r in range(nrows):
        wc = oracle.elt_fma_const(wc, cws[r], scal[r])
    wq = np.zeros((n, 8), np.uint32)
    for r in range(0, nrows - 2, 3):
        wq = oracle.elt_fma_const(wq, oracle.elt_sub(oracle.elt_mul(cws[r], cws[r + 1]), cws[r + 2]), scal[r])
    assert np.array_equal(code, wc) and np.array_equal(quad, wq)


@pytest.mark.parametrize("split", ["0", "1"])
def test_both_chain_kernels_forced(tmp_path, split):
    """LGR_CHAIN_SPLIT=1 forces sha_chain16_kernel (round split over the two half-warps, csrc/sha_kernels.cu; the default
    for n <= 1024), =0 forces the 32-column sha_chain_kernel: both must be bit-exact at every narrow size"""
    import subprocess, sys, textwrap
    code = textwrap.dedent("""
        import os, sys
        sys.path.insert(0, %r)
        import numpy as np
        import __graft_entry__ as ge
        from oracle import lgo
        lgr = ge._load_package()
        for k, R in ((256, 4100), (64, 17), (256, 7), (512, 33)):
            n = 4 * k
            ex = lgr.make_executor(max(k - 192, 1), k)
            rows = lgo.synth(3, 0, R, k)
            src = ex.make_device_buffer(R * k * 32); ex.write_buffer(src, rows)
            dig = ex.make_device_buffer(n * 32); nodes = ex.make_device_buffer((2 * n - 1) * 32)
            ex.encode_commit(src, R, dig, nodes)
            want_d, want_n, _ = lgo.encode_commit(rows, k)
            assert np.array_equal(ex.copy_to_host(dig, np.uint8).reshape(n, 32), want_d), (k, R)
            assert np.array_equal(ex.copy_to_host(nodes, np.uint8).reshape(2 * n - 1, 32), want_n), (k, R)
            ex.close()
        print("ok")
    """) % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, LGR_CHAIN_SPLIT=split)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


@pytest.mark.parametrize("k,R1,R2", [(256, 4100, 37), (64, 5, 6), (4096, 3, 4)])
def test_encode_absorb_equals_row_by_row_updates(lgr, oracle, executor_factory, k, R1, R2):
    """lgr_encode_absorb (the commit pipeline without init / final) interleaves with plain updates: two absorbs and a
    single-row update in between give the leaf digests of the whole row sequence"""
    ex = executor_factory(k)
    n = 4 * k
    rows = oracle.synth(17, 0, R1 + 1 + R2, k)
    src = ex.make_device_buffer(rows.shape[0] * k * 32)
    ex.write_buffer(src, rows)
    sha = ex.make_device_buffer(lgr.lib().lgr_sha_ctx_bytes(C.c_uint32(n)))
    lib = lgr.lib()
    assert lib.lgr_sha_init(ex._ctx, sha.ptr(), C.c_uint32(n)) == 0
    ex.encode_absorb(sha, src, R1)
    one = ex.make_codeword_buffer()
    ex.write_buffer_clear(one, rows[R1])
    ex.encode_ntt_device(ex.bind_ntt(one))
    assert lib.lgr_sha_update(ex._ctx, sha.ptr(), C.c_uint32(n), one.ptr()) == 0
    ex.encode_absorb(sha, src.slice((R1 + 1) * k * 32), R2)
    dig = ex.make_device_buffer(n * 32)
    assert lib.lgr_sha_final(ex._ctx, sha.ptr(), C.c_uint32(n), dig.ptr()) == 0
    want_d, _, _ = oracle.encode_commit(rows, k)
    assert np.array_equal(ex.copy_to_host(dig, np.uint8).reshape(n, 32), want_d)


@pytest.mark.parametrize("k,R", [(256, 20001), (64, 70000), (2048, 2100), (8192, 700)])
def test_encode_commit_across_tiles(lgr, oracle, executor_factory, k, R):
    """several tiles of the commit pipeline (tile = 2^23 codeword elements, double-buffered, encode of tile t+1 overlapping
    the hash of tile t) with a ragged last tile, device-resident and host-resident rows, against the oracle"""
    ex = executor_factory(k)
    n = 4 * k
    assert R * n > 2 * (1 << 23)
    rows = oracle.synth(29, 0, R, k)
    want_d, want_n, _ = oracle.encode_commit(rows, k)
    src = ex.make_device_buffer(R * k * 32)
    ex.write_buffer(src, rows)
    dig = ex.make_device_buffer(n * 32)
    nodes = ex.make_device_buffer((2 * n - 1) * 32)
    ex.encode_commit(src, R, dig, nodes)
    assert np.array_equal(ex.copy_to_host(dig, np.uint8).reshape(n, 32), want_d)
    assert np.array_equal(ex.copy_to_host(nodes, np.uint8).reshape(2 * n - 1, 32), want_n)
    _, root = ex.encode_commit_host(np.ascontiguousarray(rows), R)          # pageable host rows, H2D tile by tile
    assert root == want_n[0].tobytes()


def test_indexed_quad_combiner_and_row_gather(lgr, oracle, executor_factory):
    """lgr_combine_quad_indexed (triples scattered between linear rows of a tile) and lgr_sample_gather_rows against the
    per-row Eltwise sequence of check_quadratic (nonbatch_context.hpp:771-780) and sample_gather"""
    k, T = 64, 37
    n = 4 * k
    ex = executor_factory(k)
    rng = random.Random(12)
    rows = oracle.synth(41, 0, T, k)
    src = ex.make_device_buffer(T * k * 32); ex.write_buffer(src, rows)
    tile = ex.make_device_buffer(T * n * 32)
    ex.encode_rows(src, T, tile)
    cw = ex.read_elements(tile).reshape(T, n, 8)
    x_rows = [0, 5, 8, 20, 34]                                   # x at these rows, y and z right behind
    scal = [rng.randrange(P) for _ in x_rows]
    acc = ex.make_codeword_buffer()
    start = oracle.synth(42, 0, 1, n)[0]
    ex.write_buffer(acc, start)
    ex.combine_quad_indexed(tile, x_rows, scal, acc)
    want = start
    for xr, r in zip(x_rows, scal):
        want = oracle.elt_fma_const(want, oracle.elt_sub(oracle.elt_mul(cw[xr], cw[xr + 1]), cw[xr + 2]), r)
    assert np.array_equal(ex.read_elements(acc), want)
    idx = sorted(rng.sample(range(n), 192))
    ex.sampling_init(idx)
    out = ex.make_device_buffer(T * 192 * 32)
    ex.sample_gather_rows(tile, T, out)
    assert np.array_equal(ex.read_elements(out).reshape(T, 192, 8), cw[:, idx])


# ---------------------------------------------------------------- BASELINE config 2 and the large four-step plans
def test_config2_ntt_2pow20(lgr, oracle, executor_factory):
    """BASELINE config 2: 2^20-point NTT then iNTT, w = root1^(2^8), seed 2, 32 MiB in place: NTT = oracle,
    iNTT = oracle, iNTT(NTT(x)) = x"""
    ex = executor_factory(256)
    N = 1 << 20
    w = lgr.root_of_unity(20)
    assert w == pow(lgr.ROOT1, 1 << 8, P)
    x = rand_elems(oracle, 2, N)
    buf = ex.make_device_buffer(N * 32)
    ex.write_buffer(buf, x)
    ex.ntt_pow2(buf, 20, 1, w)
    f = ex.read_elements(buf)
    assert np.array_equal(f, oracle.ntt(x, w))
    ex.ntt_pow2(buf, 20, 1, w, inverse=True)
    assert np.array_equal(ex.read_elements(buf), x)
    ex.ntt_pow2(buf, 20, 1, w, inverse=True)
    assert np.array_equal(ex.read_elements(buf), oracle.ntt(x, w, inverse=True))


@pytest.mark.parametrize("logn", [17, 18, 19, 21, 22])
def test_ntt_pow2_large(lgr, oracle, executor_factory, logn):
    """four-step plans above 2^16, including the composed twist tables (twist_lo/twist_hi) of 2^21 and 2^22"""
    ex = executor_factory(256)
    N = 1 << logn
    w = lgr.root_of_unity(logn)
    x = rand_elems(oracle, 300 + logn, N)
    buf = ex.make_device_buffer(N * 32)
    ex.write_buffer(buf, x)
    ex.ntt_pow2(buf, logn, 1, w)
    f = ex.read_elements(buf)
    assert np.array_equal(f, oracle.ntt(x, w))
    ex.ntt_pow2(buf, logn, 1, w, inverse=True)
    assert np.array_equal(ex.read_elements(buf), x)
    # linearity spot check that does not go through the oracle: NTT(delta_j)[i] = w^(i*j)
    d = np.zeros((N, 8), np.uint32)
    j = 12345 % N
    d[j, 0] = 1
    ex.write_buffer(buf, d)
    ex.ntt_pow2(buf, logn, 1, w)
    got = lgr.array_to_ints(ex.read_elements(buf)[[0, 1, 2, N // 2 + 1, N - 1]])
    assert got == [pow(w, i * j, P) for i in (0, 1, 2, N // 2 + 1, N - 1)]


@pytest.mark.parametrize("ninst", [192, 5, 48, 1000])
def test_sha_instance_counts_off_the_fast_path(lgr, oracle, executor_factory, ninst):
    """192 = the verifier's instance count (sampled columns, src/webgpu_verifier.cpp); others: not a multiple of 16/32"""
    ex = executor_factory(256)
    ex.sha256_init(ninst)
    ctx = ex.make_device_buffer(ex.sha256_context_bytes(ninst))
    dig = ex.make_device_buffer(ninst * 32)
    cb = ex.bind_sha256_context(ctx, dig)
    R = 37
    rows = oracle.synth(60 + ninst, 0, R, ninst)
    tile = ex.make_device_buffer(R * ninst * 32)
    ex.write_buffer(tile, rows)
    # row by row through the reference's per-row call, then as ragged tiles
    for chunks in ([1] * R, [5, 2, 30], [36, 1], [R]):
        ex.sha256_digest_init(cb)
        s = oracle.Sha(ninst)
        pos = 0
        for c in chunks:
            if c == 1:
                ex.sha256_digest_update(cb, ex.bind_sha256_buffer(tile.slice(pos * ninst * 32)))
            else:
                ex.sha256_digest_update_rows(cb, tile.slice(pos * ninst * 32), c)
            for r in range(pos, pos + c):
                s.update(rows[r])
            pos += c
        ex.sha256_digest_final(cb)
        assert np.array_equal(ex.copy_to_host(dig, np.uint8).reshape(ninst, 32), s.final()), chunks
    ex.sha256_init(1024)


# ---------------------------------------------------------------- exact multi-GPU layout, the parts one GPU can check
@pytest.mark.parametrize("k,R,G", [(256, 70, 2), (256, 33, 8), (64, 9, 4), (2048, 5, 8), (4096, 6, 4), (8192, 3, 8), (8192, 4, 1)])
def test_encode_rows_slabs_is_the_codeword_cut_into_column_slabs(lgr, oracle, executor_factory, k, R, G):
    """lgr_encode_rows_slabs: slab h of every row, row-major [R][n/G] at its own base (fused encoder and tile engine)"""
    ex = executor_factory(k)
    n = 4 * k
    slab = n // G
    rows = oracle.synth(70 + k, 0, R, k)
    rb = ex.make_device_buffer(R * k * 32)
    ex.write_buffer(rb, rows)
    cw = ex.make_device_buffer(R * n * 32)
    ex.encode_rows(rb, R, cw)
    plain = ex.read_elements(cw).reshape(R, n, 8)
    assert np.array_equal(plain[0], oracle.encode(rows[0], k))
    T = R + 3                                                     # slab chunks are spaced for a larger tile, as in sharding.py
    out = ex.make_device_buffer(G * T * slab * 32)
    base = out.ptr().value
    ex.encode_rows_slabs(rb, R, [base + h * T * slab * 32 for h in range(G)])
    got = ex.read_elements(out).reshape(G, T, slab, 8)
    for h in range(G):
        assert np.array_equal(got[h, :R], plain[:, h * slab:(h + 1) * slab]), h
        assert not got[h, R:].any()


class _OneRankDist:
    """world-size-1 stand-in for torch.distributed (the hand-over flags and the IPC allocation still run on the device)"""
    @staticmethod
    def all_gather_object(out, obj):
        out[0] = obj

    @staticmethod
    def all_gather_into_tensor(out, t):
        out.copy_(t)


@pytest.mark.parametrize("k,T,total", [(256, 64, 64 * 5 + 17), (8192, 4, 11)])
def test_peer_store_engine_single_rank(lgr, oracle, executor_factory, k, T, total):
    """PeerStoreEngine at G = 1: IPC allocation, slab stores, ready / consumed flags, two commitments back to back"""
    import importlib.util
    import torch
    spec = importlib.util.spec_from_file_location("lgr_sharding", os.path.join(os.path.dirname(__file__), "..", "ligero-prover_b200", "sharding.py"))
    sh = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sh)
    ex = executor_factory(k)
    n = 4 * k
    eng = sh.PeerStoreEngine(ex, T, 1, 0, _OneRankDist)
    rows = oracle.synth(80 + k, 0, total, k)
    bufs = []
    for t0 in range(0, total, T):
        r = min(T, total - t0)
        b = ex.make_device_buffer(T * k * 32)
        ex.write_buffer(b, rows[t0:t0 + r])
        bufs.append((b, r))
    want, _, _ = oracle.encode_commit(rows, k)
    for _ in range(2):
        leaves = sh.commit_exact(eng, lambda i: bufs[i], total, T, 1, 0, _OneRankDist)
        torch.cuda.synchronize()
        assert np.array_equal(leaves.cpu().numpy().view(np.uint8).reshape(n, 32), want)
    eng.close()
    ex.sha256_init(1024)


def test_peer_wait_times_out_instead_of_hanging(lgr, executor_factory):
    ex = executor_factory(256)
    ptr, _ = ex.ipc_alloc(256)
    ex.peer_wait(ptr, 2, 5, ptr + 128, timeout_ms=50)             # nobody will ever signal: must give up
    ex.device_synchronize()
    import torch
    host = torch.zeros(1, dtype=torch.int32).pin_memory()
    ex.read_into(host, ptr + 128, 4)
    assert int(host[0]) != 0
    ex.peer_signal([ptr, ptr + 8], 5)
    ex.peer_wait(ptr, 2, 5, ptr + 192, timeout_ms=50)
    ex.read_into(host, ptr + 192, 4)
    assert int(host[0]) == 0
    ex.ipc_free(ptr)


@pytest.mark.parametrize("mode", ["kara", "dpf"])
def test_alternative_check_code_formulations_forced(mode):
    """check_code over a resident tile through the Karatsuba accumulators and through the FP64 pipe (both measured slower
    than the plain form, so not the default): still bit-exact against the per-row reference schedule"""
    import subprocess, sys
    here = os.path.dirname(os.path.abspath(__file__))
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_gpu_parity.py"), "-m", "gpu", "-x", "-q", "-k", "test_tile_combiners"],
                         env=dict(os.environ, LGR_COMBINE_CODE=mode), capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and " passed" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.parametrize("env", [{"LGR_CLUSTER_ENCODE_ROWS": "4"}, {"LGR_NTT_LAT_MAX": "0"}, {"LGR_NO_ENCODE_GRAPH": "1"}],
                         ids=["cluster-encoder", "no-latency-kernel", "no-graph-replay"])
def test_alternative_encode_paths_forced(env):
    """the encode paths that are not the default for their size must stay bit-exact: the thread-block-cluster one-launch
    encoder (csrc/cluster_encode_kernel.cu, off by default), the throughput tile kernel for small jobs (latency kernel off),
    plain launches instead of the CUDA-graph replay of one-row encodes"""
    import subprocess, sys, textwrap
    code = textwrap.dedent("""
        import os, sys
        sys.path.insert(0, %r)
        import numpy as np
        import __graft_entry__ as ge
        from oracle import lgo
        lgr = ge._load_package()
        for k in (4096, 8192):
            n = 4 * k
            ex = lgr.make_executor(k - 192, k)            # own stream: the graph replay path is live
            rows = lgo.synth(11, 0, 3, k)
            want = [lgo.encode(rows[r], k) for r in range(3)]
            dev = ex.make_codeword_buffer(); bind = ex.bind_ntt(dev)
            for rep in range(5):                          # the 3rd call on a buffer captures, later ones replay
                ex.write_buffer_clear(dev, rows[rep %% 3]); ex.encode_ntt_device(bind)
                assert np.array_equal(ex.read_elements(dev), want[rep %% 3]), (k, rep)
            rb = ex.make_device_buffer(3 * k * 32); ex.write_buffer(rb, rows)
            cw = ex.make_device_buffer(3 * n * 32); ex.encode_rows(rb, 3, cw)
            assert np.array_equal(ex.read_elements(cw).reshape(3, n, 8), np.stack(want)), k
            ex.decode_ntt_device(bind)
            assert np.array_equal(ex.read_elements(dev)[:k], rows[4 %% 3]), k
            ex.close()
        print("ok")
    """) % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


@pytest.mark.parametrize("k", [4096, 8192])
def test_consecutive_one_row_encodes_share_one_graph(lgr, oracle, k):
    """lgr_encode defers one-row encodes at large k and replays consecutive ones as one CUDA graph with a branch per
    codeword buffer (api.cu flush_encodes) -- the six encodes of a quadratic callback, the two of a linear one
    (nonbatch_context.hpp:667-668,715-720).  Whatever the grouping, results equal the oracle's row by row, and stream order
    stays call order: repeated, overlapping and more-than-eight buffers, other operations in between."""
    import torch
    n = 4 * k
    ex = lgr.make_executor(k - 192, k)                     # the context's own stream: deferral and graph replay are live
    try:
        nbuf = 10
        devs = [ex.make_codeword_buffer() for _ in range(nbuf)]
        big = ex.make_device_buffer((n + k) * 32)
        acc = ex.make_codeword_buffer()
        torch.cuda.synchronize()
        binds = [ex.bind_ntt(d) for d in devs]
        rows = oracle.synth(23, 0, 12, k)
        want = [oracle.encode(rows[r], k) for r in range(12)]
        # rounds of 6 / 2 / 1 / 10 consecutive encodes; the first rounds run plain, later ones capture, then replay
        for rep, m in enumerate([6, 6, 6, 6, 6, 2, 2, 2, 2, 1, 1, 1, 1, 10, 10, 10, 10, 6, 2, 6]):
            for j in range(m):
                ex.write_buffer_clear(devs[j], rows[(rep + j) % 12])
            for j in range(m):
                ex.encode_ntt_device(binds[j])
            if rep % 3 == 0:                               # an element-wise op right behind the encodes must see all of them
                ex.clear_buffer(acc)
                for j in range(m):
                    ex.EltwiseAddAssignMod(ex.bind_eltwise2(devs[j], acc))
                tot = oracle.from_limbs(ex.read_elements(acc))
                cols = [oracle.from_limbs(want[(rep + j) % 12]) for j in range(m)]
                assert tot == [sum(c[i] for c in cols) % P for i in range(n)], (k, rep, m)
            for j in range(m):
                assert np.array_equal(ex.read_elements(devs[j]), want[(rep + j) % 12]), (k, rep, m, j)
        assert ex.launch_count() > 0
        # the same buffer twice in a row: the second encode takes the first k elements of the first one's codeword
        for rep in range(2):
            ex.write_buffer_clear(devs[0], rows[3]); ex.write_buffer_clear(devs[1], rows[4])
            ex.encode_ntt_device(binds[0]); ex.encode_ntt_device(binds[1]); ex.encode_ntt_device(binds[0])
            assert np.array_equal(ex.read_elements(devs[0]), oracle.encode(want[3][:k], k)), (k, rep)
            assert np.array_equal(ex.read_elements(devs[1]), want[4])
        # overlapping windows of one allocation: B = [k, k+n) starts inside A = [0, n)
        A, B = big.slice(0, n * 32), big.slice(k * 32, (k + n) * 32)
        for rep in range(5):
            ex.write_buffer_clear(big, rows[5 + rep])
            ex.encode_ntt_device(ex.bind_ntt(A)); ex.encode_ntt_device(ex.bind_ntt(B))
            got = ex.read_elements(big)
            cwA = want[5 + rep]
            assert np.array_equal(got[:k], cwA[:k]) and np.array_equal(got[k:], oracle.encode(cwA[k:2 * k], k)), (k, rep)
        # deferred work is not lost at a synchronise, a stream switch or a decode
        ex.write_buffer_clear(devs[2], rows[7]); ex.encode_ntt_device(binds[2]); ex.device_synchronize()
        got = devs[2].storage.cpu().numpy().view(np.uint32).reshape(-1, 8)
        assert np.array_equal(got, want[7])
        ex.write_buffer_clear(devs[2], rows[8]); ex.encode_ntt_device(binds[2]); ex.decode_ntt_device(binds[2])
        assert np.array_equal(ex.read_elements(devs[2])[:k], rows[8])
        ex.write_buffer_clear(devs[2], rows[9]); ex.encode_ntt_device(binds[2]); ex.use_torch_stream()
        torch.cuda.synchronize()
        assert np.array_equal(devs[2].storage.cpu().numpy().view(np.uint32).reshape(-1, 8), want[9])
    finally:
        ex.close()


@pytest.mark.parametrize("lanes", ["1", "3"])
def test_encode_graph_lane_limits_forced(lanes):
    """LGR_ENCODE_LANES=1 (one row per graph, the round-2 schedule) and 3 (a callback's six encodes split in two graphs)"""
    import subprocess, sys
    here = os.path.dirname(os.path.abspath(__file__))
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_gpu_parity.py"), "-m", "gpu", "-x", "-q", "-k", "test_consecutive_one_row_encodes_share_one_graph"],
                         env=dict(os.environ, LGR_ENCODE_LANES=lanes), capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and " passed" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_fp64_pipe_montgomery_matches_bigint(lgr, executor_factory):
    """csrc/dpf_mont.cuh (52-bit limbs as doubles, product halves from DFMA round-toward-zero pairs): a*b*2^-260 mod p
    against Python big integers, edge values included -- the measured alternative to the IMAD CIOS is at least correct"""
    ex = executor_factory(256)
    rng = random.Random(77)
    edge = [0, 1, 2, P - 1, P - 2, (1 << 52) - 1, 1 << 52, (1 << 104) - 1, (1 << 208) + 12345, (1 << 253) % P, P >> 1]
    a = [x for x in edge for _ in edge] + [rng.randrange(P) for _ in range(4000)]
    b = [y for _ in edge for y in edge] + [rng.randrange(P) for _ in range(4000)]
    got = lgr.array_to_ints(ex.dpf_mul(lgr.ints_to_array(a), lgr.ints_to_array(b)))
    rinv = pow(1 << 260, -1, P)
    assert got == [x * y * rinv % P for x, y in zip(a, b)]

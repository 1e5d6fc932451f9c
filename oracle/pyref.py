"""Pure-Python (big ints + hashlib) restatement of the hot path, from the *definitions* the
reference implements.  TEST INFRASTRUCTURE ONLY -- used to pin oracle/oracle.c independently
(different language, different arithmetic: Python ints and % p instead of Montgomery CIOS;
hashlib instead of the hand-written compression function).

References: constants src/bn254.cpp:21-43,51-64; transforms src/webgpu/engine.cpp:755-796,844-882,
932-968 + shader/kernels.wgsl.in:57-323; hashing shader/sha256.wgsl:147-230; tree
include/zkp/merkle_tree.hpp:343-375.
"""
import hashlib
import struct

P = 21888242871839275222246405745257275088548364400416034343698204186575808495617  # src/bn254.cpp:21-22
ROOT1 = 1748695177688661943023146337482803886740723238769601073607632802312037301404  # :36-37
ROOT2 = 2037444462055058054189478067370099086220733342011840546702672064072905551290  # :38-39
POW2_DEGREE = 28                                                                        # :41-43


def omegas(k):
    """src/bn254.cpp:51-64"""
    return (pow(ROOT1, (1 << POW2_DEGREE) // k, P),
            pow(ROOT1, (1 << POW2_DEGREE) // (2 * k), P),
            pow(ROOT2, (1 << POW2_DEGREE) // (4 * k), P))


def dft(x, w, inverse=False):
    """O(N^2) definition: forward X[j] = sum x[i] w^(ij); inverse x[i] = N^-1 sum X[j] w^(-ij)"""
    n = len(x)
    if inverse:
        w = pow(w, -1, P)
    pw = [pow(w, i, P) for i in range(n)]
    out = [sum(x[i] * pw[(i * j) % n] for i in range(n)) % P for j in range(n)]
    if inverse:
        ninv = pow(n, -1, P)
        out = [v * ninv % P for v in out]
    return out


def ntt(x, w, inverse=False):
    """recursive radix-2, same function as dft()"""
    n = len(x)
    if inverse:
        y = _ntt_rec(list(x), pow(w, -1, P))
        ninv = pow(n, -1, P)
        return [v * ninv % P for v in y]
    return _ntt_rec(list(x), w)


def _ntt_rec(x, w):
    n = len(x)
    if n == 1:
        return x
    ev = _ntt_rec(x[0::2], w * w % P)
    od = _ntt_rec(x[1::2], w * w % P)
    out = [0] * n
    t = 1
    for i in range(n // 2):
        u = od[i] * t % P
        out[i] = (ev[i] + u) % P
        out[i + n // 2] = (ev[i] - u) % P
        t = t * w % P
    return out


def encode(row, k):
    """engine.cpp:755-770: coefficients = iNTT_k(row) on the w_k domain, codeword = NTT_n of the
    zero-padded coefficients on the w_n domain (n = 4k)"""
    wk, _, wn = omegas(k)
    c = ntt(row, wk, inverse=True)
    return ntt(c + [0] * (3 * k), wn)


def encode_2k(row, k):
    """nonbatch_context.hpp:482-494 mask rows: 2k evaluations on the w_2k domain"""
    _, w2k, wn = omegas(k)
    c = ntt(row, w2k, inverse=True)
    return ntt(c + [0] * (2 * k), wn)


def decode(code, k):
    """engine.cpp:772-796"""
    wk, _, wn = omegas(k)
    c = ntt(code, wn, inverse=True)
    folded = [(c[i] + c[i + k]) % P for i in range(k)]
    return ntt(folded, wk) + c[k:]


def ser(x):
    """shader/sha256.wgsl:155-162: limb 0 (least significant) first, each limb MSB first"""
    return b"".join(struct.pack(">I", (x >> (32 * i)) & 0xFFFFFFFF) for i in range(8))


def leaf_digest(column):
    """sha256.wgsl:179-230: standard SHA-256 of the serialised column; the stored digest is the 8
    state words as native little-endian u32 (:226-228) = standard digest with every 4-byte group
    byte-reversed"""
    d = hashlib.sha256(b"".join(ser(v) for v in column)).digest()
    return b"".join(d[4 * i:4 * i + 4][::-1] for i in range(8))


def merkle(leaves):
    """include/zkp/merkle_tree.hpp:343-375 (heap layout; zero digests pad to a power of two)"""
    n = 1
    while n < len(leaves):
        n <<= 1
    nodes = [b"\0" * 32] * (2 * n - 1)
    for i, l in enumerate(leaves):
        nodes[n - 1 + i] = bytes(l)
    for i in range(n - 2, -1, -1):
        nodes[i] = hashlib.sha256(nodes[2 * i + 1] + nodes[2 * i + 2]).digest()
    return nodes

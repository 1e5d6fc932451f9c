"""ctypes binding of the CPU oracle (oracle/liblgo.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module (see oracle/lgo.h).  Elements are numpy uint32 arrays of shape [..., 8]
(8 little-endian limbs, canonical, non-Montgomery) -- the reference's device_bignum<8> format
(include/ligetron/webgpu/device_bignum.hpp:30-100).
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liblgo.so")

P = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("oracle.c", "lgo.h", "Makefile")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.lgo_sha_new.restype = C.c_void_p
        _lib.lgo_merkle_nodes.restype = C.c_size_t
    return _lib


def to_limbs(vals):
    """python ints -> uint32[len, 8]"""
    vals = list(vals)
    out = np.zeros((len(vals), 8), dtype=np.uint32)
    for i, v in enumerate(vals):
        for j in range(8):
            out[i, j] = (v >> (32 * j)) & 0xFFFFFFFF
    return out


def from_limbs(arr):
    arr = np.ascontiguousarray(arr, dtype=np.uint32).reshape(-1, 8)
    return [sum(int(arr[i, j]) << (32 * j) for j in range(8)) for i in range(arr.shape[0])]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _fr(v):
    return np.ascontiguousarray(to_limbs([v])[0])


def _c(a):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    assert a.shape[-1] == 8
    return a


def omegas(k):
    wk, w2k, wn = (np.zeros(8, np.uint32) for _ in range(3))
    lib().lgo_omegas(C.c_uint64(k), _p(wk), _p(w2k), _p(wn))
    return from_limbs(wk)[0], from_limbs(w2k)[0], from_limbs(wn)[0]


def root1():
    r = np.zeros(8, np.uint32); lib().lgo_root1(_p(r)); return from_limbs(r)[0]


def root2():
    r = np.zeros(8, np.uint32); lib().lgo_root2(_p(r)); return from_limbs(r)[0]


def ntt(x, omega, inverse=False):
    x = _c(x).copy(); w = _fr(omega)
    lib().lgo_ntt(_p(x), C.c_size_t(x.shape[0]), _p(w), C.c_int(int(inverse)))
    return x


def ntt_batch(x, omega, inverse=False):
    x = _c(x).copy(); w = _fr(omega)
    assert x.ndim == 3
    lib().lgo_ntt_batch(_p(x), C.c_size_t(x.shape[1]), C.c_size_t(x.shape[0]), _p(w), C.c_int(int(inverse)))
    return x


def dft_naive(x, omega, inverse=False):
    x = _c(x); out = np.zeros_like(x); w = _fr(omega)
    lib().lgo_dft_naive(_p(out), _p(x), C.c_size_t(x.shape[0]), _p(w), C.c_int(int(inverse)))
    return out


def encode(row, k):
    """row: [k,8] (or [2k,8] with two_k) -> codeword [4k,8]"""
    row = _c(row); buf = np.zeros((4 * k, 8), np.uint32); buf[: row.shape[0]] = row
    assert row.shape[0] == k
    lib().lgo_encode(_p(buf), C.c_size_t(k)); return buf


def encode_2k(row, k):
    row = _c(row); buf = np.zeros((4 * k, 8), np.uint32); buf[: 2 * k] = row
    assert row.shape[0] == 2 * k
    lib().lgo_encode_2k(_p(buf), C.c_size_t(k)); return buf


def decode(code, k):
    buf = _c(code).copy(); assert buf.shape[0] == 4 * k
    lib().lgo_decode(_p(buf), C.c_size_t(k)); return buf


def _elt3(name, x, y, out=None):
    x = _c(x); y = _c(y); o = np.zeros_like(x) if out is None else _c(out).copy()
    getattr(lib(), name)(_p(o), _p(x), _p(y), C.c_size_t(x.reshape(-1, 8).shape[0])); return o


def elt_add(x, y): return _elt3("lgo_elt_add", x, y)
def elt_sub(x, y): return _elt3("lgo_elt_sub", x, y)
def elt_mul(x, y): return _elt3("lgo_elt_mul", x, y)
def elt_div(x, y): return _elt3("lgo_elt_div", x, y)
def elt_fma(out, x, y): return _elt3("lgo_elt_fma", x, y, out)


def _eltc(name, x, c, out=None):
    x = _c(x); o = np.zeros_like(x) if out is None else _c(out).copy(); cc = _fr(c)
    getattr(lib(), name)(_p(o), _p(x), _p(cc), C.c_size_t(x.reshape(-1, 8).shape[0])); return o


def elt_fma_const(out, x, c): return _eltc("lgo_elt_fma_const", x, c, out)
def elt_add_const(x, c): return _eltc("lgo_elt_add_const", x, c)
def elt_sub_const(x, c): return _eltc("lgo_elt_sub_const", x, c)
def elt_const_sub(x, c): return _eltc("lgo_elt_const_sub", x, c)
def elt_mul_const(x, c): return _eltc("lgo_elt_mul_const", x, c)
def elt_montmul_const(x, c): return _eltc("lgo_elt_montmul_const", x, c)


def elt_add_assign(out, x):
    x = _c(x); o = _c(out).copy()
    lib().lgo_elt_add_assign(_p(o), _p(x), C.c_size_t(x.reshape(-1, 8).shape[0])); return o


def elt_bit(x, bit):
    x = _c(x); o = np.zeros_like(x)
    lib().lgo_elt_bit(_p(o), _p(x), C.c_uint32(bit), C.c_size_t(x.reshape(-1, 8).shape[0])); return o


def elt_powmod(coeff, exp, base, out=None):
    coeff = _c(coeff); exp = np.ascontiguousarray(exp, np.uint32); b = _fr(base)
    o = np.zeros_like(coeff) if out is None else _c(out).copy()
    lib().lgo_elt_powmod(_p(o), _p(coeff), _p(exp), _p(b), C.c_size_t(coeff.shape[0]), C.c_int(out is not None)); return o


def gather(x, idx):
    x = _c(x); idx = np.ascontiguousarray(idx, np.uint32); o = np.zeros((idx.shape[0], 8), np.uint32)
    lib().lgo_gather(_p(o), _p(x), _p(idx), C.c_size_t(idx.shape[0])); return o


class Sha:
    """streaming column hasher: shader/sha256.wgsl:127-230"""
    def __init__(self, ninst):
        self.n = ninst; self.h = C.c_void_p(lib().lgo_sha_new(C.c_size_t(ninst)))
    def init(self): lib().lgo_sha_init(self.h)
    def update(self, row):
        row = _c(row); assert row.shape[0] == self.n
        lib().lgo_sha_update(self.h, _p(row))
    def final(self):
        d = np.zeros((self.n, 32), np.uint8); lib().lgo_sha_final(self.h, _p(d)); return d
    def __del__(self):
        try: lib().lgo_sha_free(self.h)
        except Exception: pass


def sha256(msg: bytes):
    out = np.zeros(32, np.uint8); m = np.frombuffer(msg, np.uint8) if len(msg) else np.zeros(1, np.uint8)
    lib().lgo_sha256(_p(out), _p(np.ascontiguousarray(m)), C.c_size_t(len(msg))); return out.tobytes()


def merkle_build(leaf_digests):
    d = np.ascontiguousarray(leaf_digests, np.uint8).reshape(-1, 32)
    nn = lib().lgo_merkle_nodes(C.c_size_t(d.shape[0]))
    nodes = np.zeros((nn, 32), np.uint8)
    lib().lgo_merkle_build(_p(nodes), _p(d), C.c_size_t(d.shape[0])); return nodes


def synth(seed, row0, nrows, ncols):
    out = np.zeros((nrows, ncols, 8), np.uint32)
    lib().lgo_synth(_p(out), C.c_uint64(seed), C.c_uint64(row0), C.c_uint64(nrows), C.c_uint64(ncols)); return out


def encode_commit(rows, k):
    rows = _c(rows); R = rows.reshape(-1, k, 8).shape[0]; n = 4 * k
    dig = np.zeros((n, 32), np.uint8); nodes = np.zeros((2 * n - 1, 32), np.uint8)
    thr = lib().lgo_encode_commit(_p(rows), C.c_size_t(R), C.c_size_t(k), _p(dig), _p(nodes))
    return dig, nodes, thr


def encode_commit_synth(seed, R, k):
    n = 4 * k
    dig = np.zeros((n, 32), np.uint8); nodes = np.zeros((2 * n - 1, 32), np.uint8)
    thr = lib().lgo_encode_commit_synth(C.c_uint64(seed), C.c_size_t(R), C.c_size_t(k), _p(dig), _p(nodes))
    return dig, nodes, thr


def num_threads():
    return lib().lgo_num_threads()


def set_threads(n):
    lib().lgo_set_threads(C.c_int(n))

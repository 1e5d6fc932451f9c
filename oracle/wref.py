"""ctypes binding of oracle/_ref/libwgslref.so -- the reference's OWN shaders run on the host.  TEST INFRASTRUCTURE.

The library is produced by `make -C oracle ref` from /root/reference/shader/* (transliterated by
oracle/wgsl2cpp.py, driven by oracle/wgslref.cpp in src/webgpu/engine.cpp's dispatch order).  It exists in the
build container and travels to the GPU box as a prebuilt file; /root/reference is never read at test time.
Only tests/ import this module: it pins oracle/oracle.c (and through it the CUDA path) to the reference.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libwgslref.so")
REF_ROOT = os.environ.get("LGR_REFERENCE_ROOT", "/root/reference")
SHA_COUNTS = (1, 5, 192, 256, 1024, 2048, 4096, 32768)     # instance counts compiled into the library


def available():
    return os.path.exists(_SO)


def build(force=False):
    """(re)build when the reference tree is present; otherwise use the prebuilt library as is."""
    if not os.path.isdir(os.path.join(REF_ROOT, "shader")):
        return _SO if available() else None
    srcs = [os.path.join(_HERE, f) for f in ("wgsl2cpp.py", "wgslref.cpp", "wgsl_shim.hpp", "Makefile")]
    if force or not available() or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "_ref/libwgslref.so", f"REF={REF_ROOT}"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            build()
        _lib = C.CDLL(_SO)
        _lib.wref_sha_ctx_words.restype = C.c_uint32
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _limbs(v):
    return np.array([(v >> (32 * j)) & 0xFFFFFFFF for j in range(8)], dtype=np.uint32)


def _int(a):
    return sum(int(a[j]) << (32 * j) for j in range(8))


def _c(a):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    assert a.shape[-1] == 8
    return a


def constants():
    bufs = [np.zeros(8, np.uint32) for _ in range(5)]
    lib().wref_constants(*[_p(b) for b in bufs])
    return dict(zip(("p", "two_p", "mont_inv", "mont_r", "barrett"), (_int(b) for b in bufs)))


def montgomery_mul(a, b, two_p=False):
    o = np.zeros(8, np.uint32)
    lib().wref_montgomery_mul(_p(_limbs(a)), _p(_limbs(b)), _p(o), C.c_int(int(two_p)))
    return _int(o)


def barrett_mul(a, b):
    o = np.zeros(8, np.uint32)
    lib().wref_barrett_mul(_p(_limbs(a)), _p(_limbs(b)), _p(o))
    return _int(o)


def invmod(a):
    o = np.zeros(8, np.uint32)
    lib().wref_invmod(_p(_limbs(a)), _p(o))
    return _int(o)


def ntt(x, omega, inverse=False, buf_elems=None):
    x = _c(x)
    N = x.shape[0]
    buf = np.zeros((buf_elems or N, 8), np.uint32)
    buf[:N] = x
    rc = lib().wref_ntt(_p(buf), C.c_uint32(buf.shape[0]), C.c_uint32(N), _p(_limbs(omega)), C.c_int(int(inverse)))
    assert rc == 0
    return buf[:N].copy()


def encode(row, k, w_k, w_n):
    buf = np.zeros((4 * k, 8), np.uint32)
    buf[:k] = _c(row)
    assert lib().wref_encode(_p(buf), C.c_uint32(k), _p(_limbs(w_k)), _p(_limbs(w_n))) == 0
    return buf


def decode(code, k, w_k, w_2k, w_n):
    buf = _c(code).copy()
    assert lib().wref_decode(_p(buf), C.c_uint32(k), _p(_limbs(w_k)), _p(_limbs(w_2k)), _p(_limbs(w_n))) == 0
    return buf


def twiddles(N, omega, inverse, stage):
    out = np.zeros((max(N // 2, 511), 8), np.uint32)
    cnt = lib().wref_twiddles(C.c_uint32(N), _p(_limbs(omega)), C.c_int(int(inverse)), C.c_uint32(stage), _p(out),
                              C.c_uint32(out.shape[0]))
    assert cnt >= 0
    return out[:cnt]


def n_inv(N, omega):
    o = np.zeros(8, np.uint32)
    assert lib().wref_n_inv(C.c_uint32(N), _p(_limbs(omega)), _p(o)) == 0
    return _int(o)


def eltwise(name, x, y=None, out=None, scalar=None):
    x = _c(x).reshape(-1, 8)
    o = np.zeros_like(x) if out is None else _c(out).reshape(-1, 8).copy()
    yy = None if y is None else _c(y).reshape(-1, 8)
    sc = None if scalar is None else _limbs(scalar)
    rc = lib().wref_eltwise(name.encode(), _p(x), _p(yy) if yy is not None else None, _p(o),
                            _p(sc) if sc is not None else None, C.c_uint32(x.shape[0]))
    assert rc == 0, name
    return o


def powmod(base, exp, coeff, out=None, add=False, workgroups=4):
    exp = np.ascontiguousarray(exp, np.uint32)
    coeff = _c(coeff)
    o = np.zeros_like(coeff) if out is None else _c(out).copy()
    rc = lib().wref_powmod(C.c_int(int(add)), _p(_limbs(base)), _p(exp), _p(coeff), _p(o), C.c_uint32(exp.shape[0]),
                           C.c_uint32(workgroups))
    assert rc == 0
    return o


def sample_gather(x, idx):
    x = _c(x)
    idx = np.ascontiguousarray(idx, np.uint32)
    assert idx.shape[0] == 192
    o = np.zeros((192, 8), np.uint32)
    assert lib().wref_sample_gather(_p(x), C.c_uint32(x.shape[0]), _p(idx), _p(o)) == 0
    return o


class Sha:
    """the reference's sha256_batch_context driven through sha256_init / update / final"""

    def __init__(self, ninst):
        assert ninst in SHA_COUNTS, f"libwgslref.so is compiled for instance counts {SHA_COUNTS}"
        self.n = ninst
        self.ctx = np.zeros(lib().wref_sha_ctx_words(C.c_uint32(ninst)), np.uint32)

    def init(self):
        assert lib().wref_sha_init(_p(self.ctx), C.c_uint32(self.n)) == 0

    def update(self, row):
        row = _c(row)
        assert row.shape[0] == self.n
        assert lib().wref_sha_update(_p(self.ctx), C.c_uint32(self.n), _p(row)) == 0

    def final(self):
        d = np.zeros((self.n, 8), np.uint32)
        assert lib().wref_sha_final(_p(self.ctx), C.c_uint32(self.n), _p(d)) == 0
        return d.view(np.uint8).reshape(self.n, 32)

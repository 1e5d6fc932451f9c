"""ctypes mirror of include/lgr_prover.h (liblgr_prover.so): the host-side prover driver above the hot path --
transcript, sampler, Merkle openings, proof container, and the three-stage prover over a witness matrix
(what src/webgpu_prover.cpp:249-471 does once the rows exist).  Same names as the C ABI, numpy in / out."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


class ProverError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "liblgr_prover.so")
        if not os.path.exists(path):
            raise ProverError("liblgr_prover.so is not built (make -C ligero-prover_b200); there is no fallback")
        C.CDLL(os.path.join(_HERE, "liblgr.so"), mode=C.RTLD_GLOBAL)
        _lib = C.CDLL(path)
        _lib.lgrp_last_error.restype = C.c_char_p
        _lib.lgrp_proof_free.restype = None
        _lib.lgrp_packer_free.restype = None
    return _lib


def _check(rc):
    if rc:
        raise ProverError(lib().lgrp_last_error().decode("utf-8", "replace"))


def _u8(b, n=None):
    a = np.frombuffer(bytes(b), np.uint8).copy()
    assert n is None or a.size == n
    return a


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def stage1_seed(root, instance_hash):
    out = np.zeros(32, np.uint8)
    _check(lib().lgrp_stage1_seed(_p(_u8(root, 32)), _p(_u8(instance_hash, 32)), _p(out)))
    return out.tobytes()


def stage2_seed(root, code, linear, quad):
    c, l, q = (np.ascontiguousarray(v, np.uint32).reshape(-1) for v in (code, linear, quad))
    assert c.size == l.size == q.size
    out = np.zeros(32, np.uint8)
    _check(lib().lgrp_stage2_seed(_p(_u8(root, 32)), _p(c), _p(l), _p(q), C.c_size_t(c.size), _p(out)))
    return out.tobytes()


def hash_random_bytes(seed, count):
    out = np.zeros(count, np.uint8)
    _check(lib().lgrp_hash_random_bytes(_p(_u8(seed, 32)), _p(out), C.c_size_t(count)))
    return out.tobytes()


def sample_indices(seed, n, sample_size=192):
    out = np.zeros(min(n, sample_size), np.uint64)
    cnt = C.c_uint64()
    _check(lib().lgrp_sample_indices(_p(_u8(seed, 32)), C.c_uint64(n), C.c_uint64(sample_size), _p(out), C.byref(cnt)))
    return [int(x) for x in out[: cnt.value]]


def fr_random(key, count, iv=bytes(16)):
    out = np.zeros((count, 8), np.uint32)
    _check(lib().lgrp_fr_random(_p(_u8(key, 32)), _p(_u8(iv, 16)), C.c_size_t(count), _p(out)))
    return out


def decommit(nodes, leaf_idx):
    nodes = np.ascontiguousarray(nodes, np.uint8).reshape(-1, 32)
    idx = np.ascontiguousarray(leaf_idx, np.uint64)
    pos = np.zeros(nodes.shape[0], np.uint64); sib = np.zeros((nodes.shape[0], 32), np.uint8)
    cnt = C.c_uint64()
    _check(lib().lgrp_decommit(_p(nodes), C.c_uint64(nodes.shape[0]), _p(idx), C.c_uint64(idx.size), _p(pos), _p(sib), C.byref(cnt)))
    return [int(x) for x in pos[: cnt.value]], [sib[i].tobytes() for i in range(cnt.value)]


def recommit(leaves, leaf_idx, total_count, siblings):
    lv = np.frombuffer(b"".join(leaves), np.uint8).copy() if leaves else np.zeros(1, np.uint8)
    sb = np.frombuffer(b"".join(siblings), np.uint8).copy() if siblings else np.zeros(1, np.uint8)
    idx = np.ascontiguousarray(leaf_idx, np.uint64)
    out = np.zeros(32, np.uint8)
    _check(lib().lgrp_recommit(_p(lv), _p(idx), C.c_uint64(idx.size), C.c_uint64(total_count), _p(sb), C.c_uint64(len(siblings)), _p(out)))
    return out.tobytes()


class RowPacker:
    """lgrp_packer_*: witness_manager's row packing (witness_manager.hpp:117-269,497-503)"""

    def __init__(self, l):
        self._h = C.c_void_p()
        self.l = l
        _check(lib().lgrp_packer_create(C.c_uint32(l), C.byref(self._h)))

    def push_linear(self, value, coef):
        v = np.ascontiguousarray(value, np.uint32).reshape(8); c = np.ascontiguousarray(coef, np.uint32).reshape(8)
        _check(lib().lgrp_packer_push_linear(self._h, _p(v), _p(c)))

    def push_quadratic(self, xyz, coef_xyz):
        v = np.ascontiguousarray(xyz, np.uint32).reshape(24); c = np.ascontiguousarray(coef_xyz, np.uint32).reshape(24)
        _check(lib().lgrp_packer_push_quadratic(self._h, _p(v), _p(c)))

    def finalize(self):
        _check(lib().lgrp_packer_finalize(self._h))

    def rows(self):
        """(kinds, values[rows, l, 8], coefs[rows, l, 8]) as numpy copies"""
        ne, nr = C.c_uint64(), C.c_uint64()
        kp, vp, cp = C.POINTER(C.c_uint8)(), C.POINTER(C.c_uint32)(), C.POINTER(C.c_uint32)()
        _check(lib().lgrp_packer_rows(self._h, C.byref(ne), C.byref(kp), C.byref(nr), C.byref(vp), C.byref(cp)))
        cnt = nr.value * self.l * 8
        kinds = np.ctypeslib.as_array(kp, (ne.value,)).copy() if ne.value else np.zeros(0, np.uint8)
        vals = np.ctypeslib.as_array(vp, (cnt,)).copy().reshape(-1, self.l, 8) if cnt else np.zeros((0, self.l, 8), np.uint32)
        coefs = np.ctypeslib.as_array(cp, (cnt,)).copy().reshape(-1, self.l, 8) if cnt else np.zeros((0, self.l, 8), np.uint32)
        return kinds, vals, coefs

    def close(self):
        if self._h:
            lib().lgrp_packer_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Statement(C.Structure):
    _fields_ = [("l", C.c_uint32), ("k", C.c_uint32), ("n_events", C.c_uint64), ("kinds", C.c_void_p), ("values", C.c_void_p),
                ("coefs", C.c_void_p), ("const_sum", C.c_uint32 * 8), ("encoding_seed", C.c_uint8 * 32), ("instance_hash", C.c_uint8 * 32),
                ("program_hash", C.c_uint8 * 32), ("generated_at_seconds", C.c_int64), ("sample_size", C.c_uint32),
                ("arena_slots", C.c_uint32), ("batch_args", C.c_void_p), ("batch_consts", C.c_void_p)]


# event kinds (include/lgr_prover.h LGRP_EV_*)
EV_LINEAR, EV_QUAD, EV_VSET, EV_VCOPY, EV_VADD, EV_VSUB, EV_VMUL, EV_VDIV, EV_VASSERT_EQ, EV_VBIT = range(10)
EV_VADDC, EV_VSUBC, EV_VCSUB, EV_VMULC, EV_VMONTMULC = range(10, 15)
HOST_ROWS = {EV_LINEAR: 1, EV_QUAD: 3, EV_VSET: 1}
COMMITTED_ROWS = {EV_LINEAR: 1, EV_QUAD: 3, EV_VSET: 1, EV_VBIT: 1, EV_VCOPY: 2, EV_VASSERT_EQ: 2, EV_VMUL: 3, EV_VDIV: 3}


class Proof:
    def __init__(self, handle):
        self._h = handle

    def _bytes(self, which):
        data, ln = C.POINTER(C.c_uint8)(), C.c_size_t()
        _check(lib().lgrp_proof_bytes(self._h, C.c_int(which), C.byref(data), C.byref(ln)))
        return C.string_at(data, ln.value)

    @property
    def envelope(self):
        return self._bytes(0)

    @property
    def gzip(self):
        return self._bytes(1)

    def info(self):
        bits, rows = C.c_uint32(), C.c_uint64()
        s1, s2 = np.zeros(32, np.uint8), np.zeros(32, np.uint8)
        _check(lib().lgrp_proof_info(self._h, C.byref(bits), _p(s1), _p(s2), C.byref(rows)))
        return {"valid": (bool(bits.value & 1), bool(bits.value & 2), bool(bits.value & 4)), "stage1_seed": s1.tobytes(),
                "stage2_seed": s2.tobytes(), "encoded_rows": rows.value}

    def timing(self):
        ms = (C.c_double * 4)()
        _check(lib().lgrp_proof_timing(self._h, ms))
        return {"stage1_ms": ms[0], "stage2_ms": ms[1], "stage3_ms": ms[2], "container_ms": ms[3]}

    def close(self):
        if self._h:
            lib().lgrp_proof_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class WatStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("private_consts", "asserts", "arithmetic_ops", "linear_witnesses", "quadratic_slots",
                                          "linear_constraints", "violated_constraints")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class WatArgs(C.Structure):
    """lgrp_wat_args (include/lgr_prover.h): the guest's arguments and the indices of the private ones"""
    _fields_ = [("nargs", C.c_uint32), ("args", C.POINTER(C.c_char_p)), ("arg_lens", C.POINTER(C.c_size_t)),
                ("nprivate", C.c_uint32), ("private_indices", C.POINTER(C.c_int32))]


def config_args(args, program_name=b"Ligero\0"):
    """the argument byte strings the reference's prover builds from the "args" list of its JSON configuration
    (src/webgpu_prover.cpp:110-147): argv[0], then {"i64": v} -> 8 little-endian bytes, {"str": s} -> the string and its NUL,
    {"hex": h} -> the decoded bytes (a leading 0x dropped, an odd digit count padded on the left)"""
    out = [bytes(program_name)]
    for a in args:
        if "i64" in a:
            out.append(int(a["i64"]).to_bytes(8, "little", signed=True))
        elif "str" in a:
            out.append(a["str"].encode() + b"\0")
        elif "hex" in a:
            h = a["hex"][2:] if a["hex"].startswith("0x") else a["hex"]
            out.append(bytes.fromhex(("0" if len(h) % 2 else "") + h))
        else:
            raise ProverError("invalid args type: %r" % (a,))
    return out


def _wat_args(args, private_indices):
    if args is None:
        return None, None
    args = [bytes(a) for a in args]
    priv = sorted(set(int(i) for i in (private_indices or ())))
    holder = ((C.c_char_p * max(1, len(args)))(*args), (C.c_size_t * max(1, len(args)))(*[len(a) for a in args]), (C.c_int32 * max(1, len(priv)))(*priv))
    wa = WatArgs(len(args), holder[0], holder[1], len(priv), holder[2])
    return wa, holder


def wat_instance_hash(args, private_indices=()):
    """lgrp_wat_instance_hash: the public arguments folded into the instance hash (src/webgpu_prover.cpp:160-168)"""
    wa, _keep = _wat_args(args, private_indices)
    out = (C.c_uint8 * 32)()
    _check(lib().lgrp_wat_instance_hash(C.byref(wa) if wa is not None else None, out))
    return bytes(out)


def wat_emit(wat_text, l, stage1_seed=None, args=None, private_indices=(), want_exit_code=False):
    """lgrp_wat_emit / lgrp_wat_emit_args: the .wat / .wasm front end + witness emitter (host only).  Returns (kinds,
    values[rows, l, 8], coefs[rows, l, 8], const_sum, stats); coefs are zero without a stage-1 seed.  `args`: the guest's
    argument byte strings (argv[0] first; see config_args), `private_indices`: which of them are secret."""
    data = wat_text.encode() if isinstance(wat_text, str) else bytes(wat_text)
    h = C.c_void_p()
    cs = (C.c_uint32 * 8)()
    st = WatStats()
    seed = _p(_u8(stage1_seed, 32)) if stage1_seed is not None else None
    wa, _keep = _wat_args(args, private_indices)
    code = C.c_int32(-1)
    _check(lib().lgrp_wat_emit_args(data, C.c_size_t(len(data)), C.byref(wa) if wa is not None else None, C.c_uint32(l), seed, C.byref(h), cs, C.byref(st),
                                    C.byref(code)))
    pk = RowPacker.__new__(RowPacker)
    pk._h, pk.l = h, l
    kinds, vals, coefs = pk.rows()
    pk.close()
    out = (kinds, vals, coefs, sum(int(cs[i]) << (32 * i) for i in range(8)), st.as_dict())
    return out + (int(code.value),) if want_exit_code else out


def prove_wat(executor, wat_text, encoding_seed=bytes(32), generated_at=0, args=None, private_indices=()):
    """lgrp_prove_wat / lgrp_prove_wat_args: .wat / .wasm -> proof on the executor's geometry (BASELINE config 4: include/lgr_prover.h)"""
    data = wat_text.encode() if isinstance(wat_text, str) else bytes(wat_text)
    h = C.c_void_p()
    st = WatStats()
    wa, _keep = _wat_args(args, private_indices)
    _check(lib().lgrp_prove_wat_args(executor._ctx, data, C.c_size_t(len(data)), C.byref(wa) if wa is not None else None, _p(_u8(encoding_seed, 32)),
                                     C.c_int64(generated_at), C.byref(h), C.byref(st)))
    return Proof(h), st.as_dict()


def parse_proof(data):
    buf = _u8(data)
    h = C.c_void_p()
    _check(lib().lgrp_proof_parse(_p(buf), C.c_size_t(buf.size), C.byref(h)))
    return Proof(h)


def prove(executor, kinds, values, coefs=None, const_sum=0, encoding_seed=bytes(32), instance_hash=bytes(32), program_hash=bytes(32),
          generated_at=0, sample_size=192, arena_slots=0, batch_args=None, batch_consts=None):
    """executor: ligero_prover_b200.Executor (its lgr_ctx runs the hot path); kinds: per event EV_* (0 = linear row,
    1 = quadratic triple, 2.. = vbn254fr calls on `arena_slots` device variables, operands in batch_args [events, 3],
    constants in batch_consts [*, 8]); values / coefs: [host rows, l, 8] uint32 in event order"""
    l, k = executor.message_size(), executor.padding_size()
    kinds = np.ascontiguousarray(kinds, np.uint8)
    values = np.ascontiguousarray(values, np.uint32).reshape(-1, l, 8)
    assert values.shape[0] == sum(HOST_ROWS.get(int(kd), 0) for kd in kinds), "one host row per linear / VSET event, three per triple"
    st = Statement()
    st.arena_slots = arena_slots
    if batch_args is not None:
        batch_args = np.ascontiguousarray(batch_args, np.uint32).reshape(-1, 3)
        assert batch_args.shape[0] == int((kinds >= EV_VSET).sum())
        st.batch_args = batch_args.ctypes.data
    if batch_consts is not None:
        batch_consts = np.ascontiguousarray(batch_consts, np.uint32).reshape(-1, 8)
        st.batch_consts = batch_consts.ctypes.data
    st.l, st.k, st.n_events = l, k, kinds.size
    st.kinds, st.values = kinds.ctypes.data, values.ctypes.data
    if coefs is not None:
        coefs = np.ascontiguousarray(coefs, np.uint32).reshape(values.shape)
        st.coefs = coefs.ctypes.data
    for i in range(8):
        st.const_sum[i] = (const_sum >> (32 * i)) & 0xFFFFFFFF
    for name, v in (("encoding_seed", encoding_seed), ("instance_hash", instance_hash), ("program_hash", program_hash)):
        getattr(st, name)[:] = list(bytes(v))
    st.generated_at_seconds, st.sample_size = generated_at, sample_size
    h = C.c_void_p()
    _check(lib().lgrp_prove(executor._ctx, C.byref(st), C.byref(h)))
    return Proof(h)

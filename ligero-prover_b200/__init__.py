"""ligero-prover_b200 -- B200-native (sm_100a) hot path of the Ligero / Ligetron prover.

Python host side over the C ABI (include/lgr.h, liblgr.so).  It mirrors the reference's executor
interface -- `webgpu_context` (include/wgpu.hpp:50-295) with `buffer_view` /
`buffer_binding` handles (include/ligetron/webgpu/buffer_view.hpp:27-69, buffer_binding.hpp:30-48)
-- with the same method names, argument meaning and error behaviour, so that parity tests read like
the stage contexts of include/zkp/nonbatch_context.hpp.  The C++ drop-in for the reference's own
build is ligero-prover_b200/host/cuda_executor.hpp.

PyTorch is plumbing only: device memory (`torch.empty(..., device="cuda")`), streams and
`torch.distributed`.  All arithmetic runs in hand-written CUDA kernels; there is NO CPU fallback and
importing this module fails loudly if liblgr.so is missing.

The directory name contains a hyphen (fixed by the repo layout), so import it with
`importlib` (see tests/conftest.py: `load_package()`), module name `ligero_prover_b200`.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LGR_LIB") or os.path.join(_HERE, "liblgr.so")   # LGR_LIB: A/B builds during tuning

# BN254 scalar field constants (src/bn254.cpp:21-43)
P = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
ROOT1 = 1748695177688661943023146337482803886740723238769601073607632802312037301404
ROOT2 = 2037444462055058054189478067370099086220733342011840546702672064072905551290
ROOT_POW2_DEGREE = 28
SAMPLE_SIZE = 192           # include/params.hpp:24-32
NUM_BYTES = 32              # device_bignum_type::num_bytes

SIZE_K, SIZE_2K, SIZE_N = 0, 1, 2
FORWARD, INVERSE = 0, 1


def build_library(force=False, jobs=None):
    """compile liblgr.so in tree (nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ...)"""
    if force:
        subprocess.check_call(["make", "-C", _HERE, "clean"])
    subprocess.check_call(["make", "-C", _HERE, "-j%d" % (jobs or os.cpu_count() or 4), "-s"])
    return LIB_PATH


class LgrError(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "liblgr.so is missing (%s). Build it with __graft_entry__.build() or `make -C %s`. "
                "There is no CPU fallback for the hot path." % (LIB_PATH, _HERE))
        _lib = C.CDLL(LIB_PATH)
        _lib.lgr_last_error.restype = C.c_char_p
        _lib.lgr_sha_ctx_bytes.restype = C.c_size_t
        _lib.lgr_merkle_node_count.restype = C.c_size_t
    return _lib


def _check(rc):
    if rc != 0:
        raise LgrError("lgr error %d: %s" % (rc, lib().lgr_last_error().decode()))


def generate_omegas(k, n):
    """bn254_gmp::generate_omegas (src/bn254.cpp:51-64)"""
    assert n == 4 * k
    return (pow(ROOT1, (1 << ROOT_POW2_DEGREE) // k, P),
            pow(ROOT1, (1 << ROOT_POW2_DEGREE) // (2 * k), P),
            pow(ROOT2, (1 << ROOT_POW2_DEGREE) // n, P))


def root_of_unity(logn, which=1):
    """primitive 2^logn-th root derived from root1 (which=1) or root2 (which=2)"""
    return pow(ROOT1 if which == 1 else ROOT2, 1 << (ROOT_POW2_DEGREE - logn), P)


def int_to_limbs(v):
    """device_bignum<8> export (include/ligetron/webgpu/device_bignum.hpp:30-100): 8 x u32 LE"""
    return (C.c_uint32 * 8)(*[(v >> (32 * i)) & 0xFFFFFFFF for i in range(8)])


def ints_to_array(vals):
    vals = list(vals)
    out = np.zeros((len(vals), 8), dtype=np.uint32)
    for i, v in enumerate(vals):
        for j in range(8):
            out[i, j] = (v >> (32 * j)) & 0xFFFFFFFF
    return out


def array_to_ints(arr):
    arr = np.ascontiguousarray(arr, dtype=np.uint32).reshape(-1, 8)
    return [sum(int(arr[i, j]) << (32 * j) for j in range(8)) for i in range(arr.shape[0])]


class Buffer:
    """buffer_view (include/ligetron/webgpu/buffer_view.hpp:27-69): a ref-counted slice of a device
    allocation.  Storage is a torch CUDA tensor (int32 words); offsets and sizes are in bytes."""

    def __init__(self, storage, offset=0, size=None):
        self.storage = storage
        self.offset_bytes = offset
        self.size_bytes = storage.numel() * 4 - offset if size is None else size

    def size(self):
        return self.size_bytes

    def offset(self):
        return self.offset_bytes

    def ptr(self):
        return C.c_void_p(self.storage.data_ptr() + self.offset_bytes)

    def slice(self, begin, end=None):
        """slice(begin) / slice(begin, end), byte units (buffer_view.hpp:56-69).  Unlike the
        reference's slice_bytes (src/webgpu/buffer_view.cpp:91-95, which drops the parent offset --
        SURVEY 8b) the parent offset is honoured."""
        end = self.size_bytes if end is None else end
        assert 0 <= begin <= end <= self.size_bytes
        return Buffer(self.storage, self.offset_bytes + begin, end - begin)

    def slice_n(self, begin, n):
        return self.slice(begin, begin + n)

    def __eq__(self, other):
        """buffer_view::operator== -- same allocation, same window (used to detect x == y squaring,
        nonbatch_context.hpp:542)"""
        return (isinstance(other, Buffer) and self.storage.data_ptr() == other.storage.data_ptr()
                and self.offset_bytes == other.offset_bytes and self.size_bytes == other.size_bytes)

    def __hash__(self):
        return hash((self.storage.data_ptr(), self.offset_bytes, self.size_bytes))


class Binding:
    """buffer_binding (include/ligetron/webgpu/buffer_binding.hpp:30-48): the buffers a kernel sees"""

    def __init__(self, *bufs):
        self._bufs = list(bufs)

    def buffers(self):
        return self._bufs


class Executor:
    """Python mirror of `webgpu_context` (include/wgpu.hpp:50-295) on the CUDA backend."""

    device_bignum_num_bytes = NUM_BYTES

    def __init__(self, device=0):
        self._ctx = None
        self._device = device
        self._l = self._k = self._n = 0
        self.sha_instances = 0

    # ---- lifecycle (wgpu.hpp:73-82) ----
    def webgpu_init(self, num_hardware_cores=0, shader_root_path=""):
        """kept for interface parity; the CUDA backend has no shader path and sizes its own grids"""
        return None

    def ntt_init(self, origin_size, padded_size, code_size, p=P, barrett_factor=None, root_k=None, root_2k=None, root_n=None):
        import torch
        if not torch.cuda.is_available():
            raise LgrError("no CUDA device: the Ligero hot path has no CPU fallback (reference README.md:330 has none either)")
        if root_k is None:
            root_k, root_2k, root_n = generate_omegas(padded_size, code_size)
        ctx = C.c_void_p()
        _check(lib().lgr_create(C.byref(ctx), C.c_int(self._device), C.c_uint32(origin_size), C.c_uint32(padded_size),
                                C.c_uint32(code_size), int_to_limbs(p), int_to_limbs(root_k), int_to_limbs(root_2k), int_to_limbs(root_n)))
        self._ctx = ctx
        self._l, self._k, self._n = origin_size, padded_size, code_size
        self._torch = torch
        self.use_torch_stream()

    def use_torch_stream(self):
        """enqueue on torch's current CUDA stream so tensors produced by torch are ordered correctly"""
        s = self._torch.cuda.current_stream(self._device).cuda_stream
        # handle 0 is the legacy default stream: pass cudaStreamLegacy (0x1) explicitly, NULL would
        # select the context's own stream
        _check(lib().lgr_set_stream(self._ctx, C.c_void_p(s if s else 1)))

    def close(self):
        if self._ctx is not None:
            lib().lgr_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def device_synchronize(self):
        _check(lib().lgr_sync(self._ctx))

    def message_size(self):
        return self._l

    def padding_size(self):
        return self._k

    def encoding_size(self):
        return self._n

    def launch_count(self):
        c = C.c_uint64()
        _check(lib().lgr_launch_count(self._ctx, C.byref(c)))
        return c.value

    # ---- buffers (device_context.hpp:44-66, wgpu.hpp:159-183) ----
    def make_device_buffer(self, num_bytes):
        t = self._torch.zeros((num_bytes + 3) // 4, dtype=self._torch.int32, device="cuda:%d" % self._device)
        return Buffer(t, 0, num_bytes)

    def make_message_buffer(self):
        return self.make_device_buffer(self._l * NUM_BYTES)

    def make_codeword_buffer(self):
        return self.make_device_buffer(self._n * NUM_BYTES)

    def make_sample_buffer(self):
        return self.make_device_buffer(SAMPLE_SIZE * NUM_BYTES)

    def wrap(self, tensor):
        """view an existing CUDA tensor (any dtype, contiguous) as a Buffer"""
        assert tensor.is_cuda and tensor.is_contiguous()
        return Buffer(tensor.view(self._torch.int32).reshape(-1), 0, tensor.numel() * tensor.element_size())

    def write_buffer(self, buf, data, length=None):
        data = np.ascontiguousarray(data)
        nbytes = data.nbytes if length is None else length * data.itemsize
        assert nbytes <= buf.size()
        _check(lib().lgr_write(self._ctx, buf.ptr(), C.c_size_t(0), data.ctypes.data_as(C.c_void_p), C.c_size_t(nbytes)))

    def write_buffer_clear(self, buf, data, length=None):
        """write the prefix, zero the rest (device_context.hpp:94-98)"""
        data = np.ascontiguousarray(data)
        nbytes = data.nbytes if length is None else length * data.itemsize
        _check(lib().lgr_write_clear(self._ctx, buf.ptr(), C.c_size_t(buf.size()), data.ctypes.data_as(C.c_void_p), C.c_size_t(nbytes)))

    def write_limbs(self, buf, vals, size=None):
        """wgpu.hpp:171-183"""
        if isinstance(vals, int):
            vals = [vals] * size
        self.write_buffer(buf, ints_to_array(vals))

    def clear_buffer(self, buf):
        _check(lib().lgr_clear(self._ctx, buf.ptr(), C.c_size_t(0), C.c_size_t(buf.size())))

    def copy_buffer_to_buffer(self, src, dst, nbytes=None):
        nbytes = min(src.size(), dst.size()) if nbytes is None else nbytes
        _check(lib().lgr_copy(self._ctx, src.ptr(), dst.ptr(), C.c_size_t(nbytes)))

    def copy_buffer_clear(self, src, dst):
        _check(lib().lgr_copy_clear(self._ctx, src.ptr(), C.c_size_t(src.size()), dst.ptr(), C.c_size_t(dst.size())))

    def copy_to_host(self, buf, dtype=np.uint32):
        """blocking read-back (device_context.hpp:70-79)"""
        out = np.empty(buf.size() // np.dtype(dtype).itemsize, dtype=dtype)
        _check(lib().lgr_read(self._ctx, out.ctypes.data_as(C.c_void_p), buf.ptr(), C.c_size_t(0), C.c_size_t(buf.size())))
        return out

    def read_into(self, host_tensor, dev_addr, nbytes):
        """blocking device -> host copy from a raw device address (peer-visible IPC memory) into a pinned torch tensor"""
        _check(lib().lgr_read(self._ctx, C.c_void_p(host_tensor.data_ptr()), C.c_void_p(int(dev_addr)), C.c_size_t(0), C.c_size_t(nbytes)))

    def read_elements(self, buf):
        return self.copy_to_host(buf).reshape(-1, 8)

    # ---- bindings (wgpu.hpp:87-96) ----
    def bind_ntt(self, buf):
        return Binding(buf)

    def bind_scalar(self, s):
        return Binding(s)

    def bind_eltwise2(self, x, out):
        return Binding(x, out)

    def bind_eltwise3(self, x, y, out):
        return Binding(x, y, out)

    def bind_sha256_context(self, context, digest):
        return Binding(context, digest)

    def bind_sha256_buffer(self, inp):
        return Binding(inp)

    def bind_sampling(self, src, dst):
        return Binding(src, dst)

    def bind_powmod(self, exp, coeff, out):
        return Binding(exp, coeff, out)

    # ---- NTT (wgpu.hpp:117-137) ----
    def _ntt(self, bind, sel, direction):
        _check(lib().lgr_ntt(self._ctx, bind.buffers()[0].ptr(), C.c_int(sel), C.c_int(direction)))

    def ntt_forward_k(self, bind): self._ntt(bind, SIZE_K, FORWARD)
    def ntt_forward_2k(self, bind): self._ntt(bind, SIZE_2K, FORWARD)
    def ntt_forward_n(self, bind): self._ntt(bind, SIZE_N, FORWARD)
    def ntt_inverse_k(self, bind): self._ntt(bind, SIZE_K, INVERSE)
    def ntt_inverse_2k(self, bind): self._ntt(bind, SIZE_2K, INVERSE)
    def ntt_inverse_n(self, bind): self._ntt(bind, SIZE_N, INVERSE)

    def encode_ntt_device(self, bind):
        buf = bind.buffers()[0]
        assert buf.size() == self._n * NUM_BYTES      # engine.cpp:756
        _check(lib().lgr_encode(self._ctx, buf.ptr()))

    def decode_ntt_device(self, bind):
        buf = bind.buffers()[0]
        assert buf.size() == self._n * NUM_BYTES      # engine.cpp:773
        _check(lib().lgr_decode(self._ctx, buf.ptr()))

    def ntt_pow2(self, buf, logn, batch, omega, inverse=False):
        _check(lib().lgr_ntt_pow2(self._ctx, buf.ptr(), C.c_uint32(logn), C.c_uint32(batch), int_to_limbs(omega), C.c_int(int(inverse))))

    # ---- SHA-256 (wgpu.hpp:139-144) ----
    def sha256_context_bytes(self, ninst):
        return lib().lgr_sha_ctx_bytes(C.c_uint32(ninst))

    def sha256_init(self, num_instances):
        self.sha_instances = num_instances

    def sha256_digest_init(self, ctx_bind):
        _check(lib().lgr_sha_init(self._ctx, ctx_bind.buffers()[0].ptr(), C.c_uint32(self.sha_instances)))

    def sha256_digest_update(self, ctx_bind, buf_bind):
        _check(lib().lgr_sha_update(self._ctx, ctx_bind.buffers()[0].ptr(), C.c_uint32(self.sha_instances), buf_bind.buffers()[0].ptr()))

    def sha256_digest_update_rows(self, ctx_bind, tile, nrows, row_stride_elems=None):
        rs = self.sha_instances if row_stride_elems is None else row_stride_elems
        _check(lib().lgr_sha_update_rows(self._ctx, ctx_bind.buffers()[0].ptr(), C.c_uint32(self.sha_instances), tile.ptr(), C.c_uint64(rs), C.c_uint32(nrows)))

    def sha256_digest_final(self, ctx_bind):
        _check(lib().lgr_sha_final(self._ctx, ctx_bind.buffers()[0].ptr(), C.c_uint32(self.sha_instances), ctx_bind.buffers()[1].ptr()))

    def merkle_node_count(self, nleaves):
        return lib().lgr_merkle_node_count(C.c_uint32(nleaves))

    def merkle_build(self, digests, nleaves, nodes):
        _check(lib().lgr_merkle_build(self._ctx, digests.ptr(), C.c_uint32(nleaves), nodes.ptr()))

    # ---- sampling (wgpu.hpp:146-149) ----
    def sampling_init(self, indexes):
        arr = (C.c_uint64 * len(indexes))(*indexes)
        _check(lib().lgr_sample_init(self._ctx, arr, C.c_uint32(len(indexes))))
        self._num_samplings = len(indexes)

    def sample_gather(self, bind, sampling_offset=0):
        """out[sampling_offset*192 + i] = x[idx[i]] (engine.cpp:1792-1809; the offset selects the
        row slot of the 256-row staging buffer, nonbatch_context.hpp:935-950)"""
        src, dst = bind.buffers()
        out = dst.slice(sampling_offset * self._num_samplings * NUM_BYTES)
        _check(lib().lgr_sample_gather(self._ctx, src.ptr(), out.ptr()))

    # ---- element-wise (wgpu.hpp:98-115); element_offsets = (x, y, z) element offsets ----
    @staticmethod
    def _off(buf, off):
        return C.c_void_p(buf.storage.data_ptr() + buf.offset_bytes + off * NUM_BYTES)

    def _n3(self, bind, offs):
        x, y, o = bind.buffers()
        ox, oy, oz = offs
        n = min(x.size() // NUM_BYTES - ox, y.size() // NUM_BYTES - oy, o.size() // NUM_BYTES - oz) if any(offs) else x.size() // NUM_BYTES
        return self._off(x, ox), self._off(y, oy), self._off(o, oz), n

    def _n2(self, bind, offs):
        x, o = bind.buffers()
        ox, _, oz = offs
        n = min(x.size() // NUM_BYTES - ox, o.size() // NUM_BYTES - oz) if any(offs) else x.size() // NUM_BYTES
        return self._off(x, ox), self._off(o, oz), n

    def EltwiseAddMod(self, bind, k=None, element_offsets=(0, 0, 0), count=None):
        if k is None:
            x, y, o, n = self._n3(bind, element_offsets)
            _check(lib().lgr_elt_add(self._ctx, x, y, o, C.c_size_t(count or n)))
        else:
            x, o, n = self._n2(bind, element_offsets)
            _check(lib().lgr_elt_add_const(self._ctx, x, o, C.c_size_t(count or n), int_to_limbs(k)))

    def EltwiseSubMod(self, bind, element_offsets=(0, 0, 0), count=None):
        x, y, o, n = self._n3(bind, element_offsets)
        _check(lib().lgr_elt_sub(self._ctx, x, y, o, C.c_size_t(count or n)))

    def EltwiseMultMod(self, bind, k=None, element_offsets=(0, 0, 0), count=None):
        if k is None:
            x, y, o, n = self._n3(bind, element_offsets)
            _check(lib().lgr_elt_mul(self._ctx, x, y, o, C.c_size_t(count or n)))
        else:
            x, o, n = self._n2(bind, element_offsets)
            _check(lib().lgr_elt_mul_const(self._ctx, x, o, C.c_size_t(count or n), int_to_limbs(k)))

    def EltwiseDivMod(self, bind, element_offsets=(0, 0, 0), count=None):
        x, y, o, n = self._n3(bind, element_offsets)
        _check(lib().lgr_elt_div(self._ctx, x, y, o, C.c_size_t(count or n)))

    def EltwiseFMAMod(self, bind, k=None, element_offsets=(0, 0, 0), count=None):
        """out += x*y (3 buffers) or out += k*x (2 buffers + scalar) -- engine.cpp:683-729"""
        if k is None:
            x, y, o, n = self._n3(bind, element_offsets)
            _check(lib().lgr_elt_fma(self._ctx, x, y, o, C.c_size_t(count or n)))
        else:
            x, o, n = self._n2(bind, element_offsets)
            _check(lib().lgr_elt_fma_const(self._ctx, x, o, C.c_size_t(count or n), int_to_limbs(k)))

    def EltwiseAddAssignMod(self, bind, element_offsets=(0, 0, 0), count=None):
        x, o, n = self._n2(bind, element_offsets)
        _check(lib().lgr_elt_add_assign(self._ctx, x, o, C.c_size_t(count or n)))

    def EltwiseSubConstMod(self, bind, k, element_offsets=(0, 0, 0), count=None):
        x, o, n = self._n2(bind, element_offsets)
        _check(lib().lgr_elt_sub_const(self._ctx, x, o, C.c_size_t(count or n), int_to_limbs(k)))

    def EltwiseConstSubMod(self, bind, k, element_offsets=(0, 0, 0), count=None):
        x, o, n = self._n2(bind, element_offsets)
        _check(lib().lgr_elt_const_sub(self._ctx, x, o, C.c_size_t(count or n), int_to_limbs(k)))

    def EltwiseMontMultMod(self, bind, k, element_offsets=(0, 0, 0), count=None):
        x, o, n = self._n2(bind, element_offsets)
        _check(lib().lgr_elt_montmul_const(self._ctx, x, o, C.c_size_t(count or n), int_to_limbs(k)))

    def EltwiseBitDecompose(self, bind, i, element_offsets=(0, 0, 0), count=None):
        x, o, n = self._n2(bind, element_offsets)
        _check(lib().lgr_elt_bit(self._ctx, x, o, C.c_size_t(count or n), C.c_uint32(i)))

    # powmod (wgpu.hpp:84-85,107-110; src/webgpu/powmod_context.cpp)
    def powmod_init(self, num_exponent_bits=32):
        assert num_exponent_bits <= 32

    def powmod_set_base(self, base, p=P):
        self._powmod_base = base % P

    def EltwisePowMod(self, bind):
        exp, coeff, out = bind.buffers()
        _check(lib().lgr_elt_powmod(self._ctx, coeff.ptr(), exp.ptr(), out.ptr(), C.c_size_t(exp.size() // 4), int_to_limbs(self._powmod_base), C.c_int(0)))

    def EltwisePowAddMod(self, bind):
        exp, coeff, out = bind.buffers()
        _check(lib().lgr_elt_powmod(self._ctx, coeff.ptr(), exp.ptr(), out.ptr(), C.c_size_t(exp.size() // 4), int_to_limbs(self._powmod_base), C.c_int(1)))

    # ---- B200-native batched fast paths ----
    def quadratic_fused(self, x, y, z, out, r):
        """check_quadratic (nonbatch_context.hpp:771-780) in one sweep: out += r*(x*y - z)"""
        _check(lib().lgr_elt_quad(self._ctx, x.ptr(), y.ptr(), z.ptr(), out.ptr(), C.c_size_t(x.size() // NUM_BYTES), int_to_limbs(r)))

    def encode_rows(self, rows, nrows, codewords, row_stride_elems=None):
        rs = self._k if row_stride_elems is None else row_stride_elems
        _check(lib().lgr_encode_rows(self._ctx, rows.ptr(), C.c_uint64(rs), C.c_uint32(nrows), codewords.ptr()))

    # ---- exact multi-GPU layout (include/lgr.h: slab-major codewords, peer memory, hand-over flags) ----
    def encode_rows_slabs(self, rows, nrows, slab_ptrs, row_stride_elems=None):
        """column slab h of every codeword -> row-major [nrows][n/G] at slab_ptrs[h] (raw device addresses, local or peer)"""
        rs = self._k if row_stride_elems is None else row_stride_elems
        arr = (C.c_void_p * len(slab_ptrs))(*[C.c_void_p(int(a)) for a in slab_ptrs])
        _check(lib().lgr_encode_rows_slabs(self._ctx, rows.ptr(), C.c_uint64(rs), C.c_uint32(nrows), arr, C.c_uint32(len(slab_ptrs))))

    def ipc_alloc(self, nbytes):
        """peer-visible device memory: (device address, 64-byte CUDA IPC handle)"""
        ptr = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        _check(lib().lgr_ipc_alloc(self._ctx, C.c_size_t(nbytes), C.byref(ptr), handle))
        return ptr.value, bytes(handle)

    def ipc_open(self, handle):
        ptr = C.c_void_p()
        _check(lib().lgr_ipc_open(self._ctx, (C.c_ubyte * 64)(*handle), C.byref(ptr)))
        return ptr.value

    def ipc_close(self, ptr):
        _check(lib().lgr_ipc_close(self._ctx, C.c_void_p(ptr)))

    def ipc_free(self, ptr):
        _check(lib().lgr_ipc_free(self._ctx, C.c_void_p(ptr)))

    def peer_signal(self, slot_ptrs, value):
        arr = (C.c_void_p * len(slot_ptrs))(*[C.c_void_p(int(a)) for a in slot_ptrs])
        _check(lib().lgr_peer_signal(self._ctx, arr, C.c_uint32(len(slot_ptrs)), C.c_uint64(value)))

    def peer_wait(self, flags_ptr, nflags, value, err_ptr, timeout_ms=20000):
        _check(lib().lgr_peer_wait(self._ctx, C.c_void_p(int(flags_ptr)), C.c_uint32(nflags), C.c_uint64(value), C.c_uint32(timeout_ms), C.c_void_p(int(err_ptr))))

    def encode_commit(self, rows, nrows, digests, nodes=None):
        _check(lib().lgr_encode_commit(self._ctx, rows.ptr(), C.c_uint64(nrows), digests.ptr(), nodes.ptr() if nodes is not None else None))

    def encode_commit_host(self, host_rows, nrows, want_digests=False):
        """host-resident witness (numpy array or pinned torch tensor) -> (digests or None, root bytes)"""
        ptr = host_rows.ctypes.data if isinstance(host_rows, np.ndarray) else host_rows.data_ptr()
        root = np.zeros(32, np.uint8)
        dig = np.zeros((self._n, 32), np.uint8) if want_digests else None
        _check(lib().lgr_encode_commit_host(self._ctx, C.c_void_p(ptr), C.c_uint64(nrows),
                                            dig.ctypes.data_as(C.c_void_p) if want_digests else None, root.ctypes.data_as(C.c_void_p)))
        return dig, root.tobytes()

    def encode_absorb(self, sha_ctx, rows, nrows):
        """lgr_encode_absorb: encode nrows device-resident rows and absorb them, in order, into a caller-owned column-hash context"""
        _check(lib().lgr_encode_absorb(self._ctx, sha_ctx.ptr(), rows.ptr(), C.c_uint64(nrows)))

    def profile(self, enable):
        _check(lib().lgr_profile(self._ctx, C.c_int(int(enable))))

    def profile_read(self):
        em, sm = C.c_double(), C.c_double()
        el, sl = C.c_uint64(), C.c_uint64()
        _check(lib().lgr_profile_read(self._ctx, C.byref(em), C.byref(el), C.byref(sm), C.byref(sl)))
        return {"encode_ms": em.value, "encode_launches": el.value, "sha_ms": sm.value, "sha_launches": sl.value}

    def combine_code(self, tile, nrows, scalars, acc):
        """scalars: python ints or a [nrows, 8] uint32 array of canonical limbs"""
        r = np.ascontiguousarray(scalars if isinstance(scalars, np.ndarray) else ints_to_array(scalars), dtype=np.uint32)
        _check(lib().lgr_combine_code(self._ctx, tile.ptr(), C.c_uint32(nrows), r.ctypes.data_as(C.c_void_p), acc.ptr()))

    def combine_quad(self, tile_x, tile_y, tile_z, nrows, scalars, acc):
        r = np.ascontiguousarray(scalars if isinstance(scalars, np.ndarray) else ints_to_array(scalars), dtype=np.uint32)
        _check(lib().lgr_combine_quad(self._ctx, tile_x.ptr(), tile_y.ptr(), tile_z.ptr(), C.c_uint32(nrows), r.ctypes.data_as(C.c_void_p), acc.ptr()))

    def combine_quad_indexed(self, tile, x_rows, scalars, acc):
        """triples scattered over a tile encoded in emission order: triple t = codeword rows x_rows[t], +1, +2"""
        r = np.ascontiguousarray(scalars if isinstance(scalars, np.ndarray) else ints_to_array(scalars), dtype=np.uint32)
        xr = np.ascontiguousarray(x_rows, dtype=np.uint32)
        _check(lib().lgr_combine_quad_indexed(self._ctx, tile.ptr(), xr.ctypes.data_as(C.c_void_p), C.c_uint32(xr.size), r.ctypes.data_as(C.c_void_p), acc.ptr()))

    def sample_gather_rows(self, tile, nrows, out, row_stride=None):
        """out[t][s] = tile[t][idx[s]] for nrows resident codewords (sampling_init first)"""
        _check(lib().lgr_sample_gather_rows(self._ctx, tile.ptr(), C.c_uint64(self._n if row_stride is None else row_stride), C.c_uint32(nrows), out.ptr()))

    def combine_linear(self, tile_a, tile_b, nrows, acc):
        _check(lib().lgr_combine_linear(self._ctx, tile_a.ptr(), tile_b.ptr(), C.c_uint32(nrows), acc.ptr()))

    def synth(self, out, seed, row0, nrows, ncols):
        _check(lib().lgr_synth(self._ctx, out.ptr(), C.c_uint64(seed), C.c_uint64(row0), C.c_uint64(nrows), C.c_uint64(ncols)))

    def ubench(self, which):
        """chip-wide operations per second of one primitive (include/lgr_ubench.h, liblgr_ubench.so)"""
        v = C.c_double()
        _ucheck(ulib().lgru_ubench(C.c_int(self._device), C.c_int(which), C.byref(v)))
        return v.value

    def ubench_chain(self, variant, warps_per_cta=1, active_lanes=32):
        """cycles per SHA-256 compression of a lone warp (csrc/ubench.cu variants)"""
        v = C.c_double()
        _ucheck(ulib().lgru_chain(C.c_int(self._device), C.c_int(variant), C.c_int(warps_per_cta), C.c_int(active_lanes), C.byref(v)))
        return v.value

    def ubench_overlap(self):
        """(mont alone, sha alone, both) milliseconds: lgru_overlap"""
        ms = (C.c_double * 3)()
        _ucheck(ulib().lgru_overlap(C.c_int(self._device), ms))
        return {"mont_alone_ms": ms[0], "sha_alone_ms": ms[1], "both_ms": ms[2], "overlap_efficiency": (ms[0] + ms[1] - ms[2]) / min(ms[0], ms[1])}

    def dpf_mul(self, a, b):
        """a*b*2^-260 mod p through the FP64-pipe Montgomery multiplication (csrc/dpf_mont.cuh); numpy [n, 8] uint32 in / out"""
        a = np.ascontiguousarray(a, np.uint32).reshape(-1, 8); b = np.ascontiguousarray(b, np.uint32).reshape(-1, 8)
        out = np.zeros_like(a)
        _ucheck(ulib().lgru_dpf_mul(C.c_int(self._device), a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.c_uint32(a.shape[0])))
        return out


_ulib = None


def ulib():
    """liblgr_ubench.so: micro-benchmarks, outside the drop-in library"""
    global _ulib
    if _ulib is None:
        path = os.path.join(_HERE, "liblgr_ubench.so")
        if not os.path.exists(path):
            raise ImportError("liblgr_ubench.so is missing (%s); build it with `make -C %s`" % (path, _HERE))
        _ulib = C.CDLL(path)
        _ulib.lgru_last_error.restype = C.c_char_p
    return _ulib


def _ucheck(rc):
    if rc != 0:
        raise LgrError("lgr ubench error: %s" % ulib().lgru_last_error().decode())


def make_executor(l, k, n=None, device=0):
    """webgpu_init + ntt_init with the reference's default roots (src/webgpu_prover.cpp:226-237)"""
    n = 4 * k if n is None else n
    ex = Executor(device)
    ex.webgpu_init(0, "")
    ex.ntt_init(l, k, n)
    return ex

"""CPU suite: the CUDA field arithmetic under host emulation, the C ABI surface, host-side logic and
the world_size-2 (gloo) multi-rank path.  No GPU needed, no compute call into liblgr.so."""
import ctypes as C
import os
import random
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "-s"])
    return C.CDLL(os.path.join(ROOT, "tests", "cpp", "libfremu.so"))


def _run(fn, a, b, oracle, *extra):
    A, B = oracle.to_limbs(a), oracle.to_limbs(b)
    O = np.zeros_like(A)
    fn(O.ctypes.data_as(C.c_void_p), A.ctypes.data_as(C.c_void_p), B.ctypes.data_as(C.c_void_p), C.c_size_t(len(a)), *extra)
    return oracle.from_limbs(O)


def test_device_montgomery_algorithm_under_emulation(emu, oracle):
    """fr.cuh's even/odd-column CIOS (the code the kernels run) with the PTX carry primitives emulated:
    a in [0,4p), b in [0,p) -> a*b/R mod p in [0,2p); canonical variant equals the oracle's montmul"""
    rng = random.Random(5)
    edge = [0, 1, 2, P - 1, P - 2, (1 << 32) - 1, (1 << 64) - 1, P >> 1]
    a = [rng.randrange(4 * P) for _ in range(4000)] + [x for x in edge for _ in edge] + [4 * P - 1] * 3
    b = [rng.randrange(P) for _ in range(4000)] + [y for _ in edge for y in edge] + [P - 1, 0, 1]
    rinv = pow(1 << 256, -1, P)
    out = _run(emu.emu_mont_mul, a, b, oracle, C.c_int(0))
    assert all(z % P == x * y * rinv % P and z < 2 * P for x, y, z in zip(a, b, out))
    out = _run(emu.emu_mont_mul, a, b, oracle, C.c_int(1))
    assert out == [x * y * rinv % P for x, y in zip(a, b)]
    ac = [x % P for x in a]
    got = oracle.from_limbs(oracle.elt_montmul_const(oracle.to_limbs(ac[:50]), b[0]))
    assert got == [x * b[0] * rinv % P for x in ac[:50]]


def test_device_shoup_multiplication_under_emulation(emu, oracle):
    """fr_shoup_mul (twiddle multiplications: truncated high product + two low products): any x < 2^256, w < p ->
    x*w mod p in [0,2p)"""
    rng = random.Random(8)
    edge = [0, 1, 2, P - 1, P, 2 * P, 4 * P - 1, (1 << 256) - 1, (1 << 255), (1 << 224) - 1, (1 << 32) - 1]
    wedge = [0, 1, 2, P - 1, P - 2, P >> 1, (1 << 32) - 1, (1 << 224), 7]
    a = [rng.randrange(1 << 256) for _ in range(6000)] + [x for x in edge for _ in wedge]
    w = [rng.randrange(P) for _ in range(6000)] + [y for _ in edge for y in wedge]
    wq = [(y << 256) // P for y in w]
    A, W, WQ = oracle.to_limbs(a), oracle.to_limbs(w), oracle.to_limbs(wq)
    O = np.zeros_like(A)
    emu.emu_shoup_mul(O.ctypes.data_as(C.c_void_p), A.ctypes.data_as(C.c_void_p), W.ctypes.data_as(C.c_void_p), WQ.ctypes.data_as(C.c_void_p), C.c_size_t(len(a)))
    out = oracle.from_limbs(O)
    assert all(z % P == x * y % P and z < 2 * P for x, y, z in zip(a, w, out))


def test_device_lazy_reductions_under_emulation(emu, oracle):
    rng = random.Random(6)
    a = [rng.randrange(P) for _ in range(2000)] + [0, P - 1, 0, P - 1]
    b = [rng.randrange(P) for _ in range(2000)] + [0, P - 1, P - 1, 0]
    assert _run(emu.emu_binop, a, b, oracle, C.c_int(0)) == [(x + y) % P for x, y in zip(a, b)]
    assert _run(emu.emu_binop, a, b, oracle, C.c_int(1)) == [(x - y) % P for x, y in zip(a, b)]
    a2 = [rng.randrange(2 * P) for _ in range(2000)] + [2 * P - 1, 0, 2 * P - 1, 0]
    b2 = [rng.randrange(2 * P) for _ in range(2000)] + [2 * P - 1, 0, 0, 2 * P - 1]
    r = _run(emu.emu_binop, a2, b2, oracle, C.c_int(2)); assert all(z % P == (x + y) % P and z < 2 * P for x, y, z in zip(a2, b2, r))
    r = _run(emu.emu_binop, a2, b2, oracle, C.c_int(3)); assert all(z == x - y + 2 * P for x, y, z in zip(a2, b2, r))
    r = _run(emu.emu_binop, a2, b2, oracle, C.c_int(4)); assert all(z % P == (x - y) % P and z < 2 * P for x, y, z in zip(a2, b2, r))
    a4 = [rng.randrange(4 * P) for _ in range(2000)] + [4 * P - 1, 0, P, 2 * P, 3 * P, P - 1, 2 * P - 1]
    assert _run(emu.emu_binop, a4, a4, oracle, C.c_int(5)) == [x % P for x in a4]


def test_pass_schedule_index_math():
    """Python model of ntt_pass / ntt_passes_from (csrc/ntt.cuh): DIT(bit-reversed input) and DIF both
    equal the DFT for every supported tile size"""
    from oracle import pyref

    def sched(L):
        out, lo = [], 0
        while lo < L:
            left = L - lo
            t = 2 if left == 4 else (3 if left >= 3 else left)
            out.append((lo, t)); lo += t
        return out

    def ntt_pass(sm, L, lo, t, dif, tw):
        M = 1 << L; TL = M // 8 if M >= 8 else 1; E = M // TL; R = 1 << t; G = E // R; S = 1 << lo
        for tl in range(TL):
            for g in range(G):
                q = tl + g * TL; low = q & (S - 1); base = low | ((q >> lo) << (lo + t))
                x = [sm[base + j * S] for j in range(R)]
                for ss in range(t):
                    s = (t - 1 - ss) if dif else ss
                    b = lo + s; h = 1 << s
                    for jl in range(h):
                        e = low * (M >> (b + 1)) + jl * (M >> (s + 1))
                        w = tw[e] if b > 0 else 1
                        for jh in range(R >> (s + 1)):
                            j0 = jl | (jh << (s + 1)); j1 = j0 | h
                            u, v = x[j0], x[j1]
                            if dif:
                                x[j0], x[j1] = (u + v) % P, (u - v) * w % P
                            else:
                                tt = v * w % P
                                x[j0], x[j1] = (u + tt) % P, (u - tt) % P
                for j in range(R):
                    sm[base + j * S] = x[j]

    brev = lambda x, b: int(bin(x)[2:].zfill(b)[::-1], 2) if b else 0
    rng = random.Random(1)
    for L in range(1, 12):
        M = 1 << L
        w = pow(pyref.ROOT1, (1 << 28) // M, P)
        tw = [pow(w, j, P) for j in range(M // 2)]
        x = [rng.randrange(P) for _ in range(M)]
        ref = pyref.ntt(x, w)
        sm = [0] * M
        for i in range(M):
            sm[brev(i, L)] = x[i]
        for lo, t in sched(L):
            ntt_pass(sm, L, lo, t, False, tw)
        assert sm == ref
        sm = list(x)
        for lo, t in reversed(sched(L)):
            ntt_pass(sm, L, lo, t, True, tw)
        assert [sm[brev(i, L)] for i in range(M)] == ref


def test_coset_decomposition_of_the_encoder():
    """encode_kernels.cu: e[4m+r] = NTT_k(c_i * w_n^(r i)) with root w_n^4 -- equals NTT_n(c || 0)"""
    from oracle import pyref
    k = 16
    rng = random.Random(2)
    row = [rng.randrange(P) for _ in range(k)]
    wk, _, wn = pyref.omegas(k)
    c = pyref.ntt(row, wk, inverse=True)
    want = pyref.encode(row, k)
    w4 = pow(wn, 4, P)
    for r in range(4):
        tw = [c[i] * pow(wn, r * i, P) % P for i in range(k)]
        assert pyref.ntt(tw, w4) == want[r::4]


def test_c_abi_exports_every_declared_symbol(lgr):
    """liblgr.so loads on a machine without a GPU and exports exactly what include/lgr.h declares"""
    hdr = open(os.path.join(ROOT, "include", "lgr.h")).read()
    declared = set(re.findall(r"\b(lgr_[a-z0-9_]+)\s*\(", hdr))
    declared.discard("lgr_ctx")
    assert len(declared) >= 45
    lib = lgr.lib()
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert lib.lgr_version() >= 100
    assert lib.lgr_sha_ctx_bytes(C.c_uint32(1024)) == 72 * 1024 <= 1024 * 300      # fits webgpu_context::sha256_context (wgpu.hpp:63-68)
    assert lib.lgr_merkle_node_count(C.c_uint32(1024)) == 2047 and lib.lgr_merkle_node_count(C.c_uint32(1025)) == 4095


@pytest.mark.parametrize("hdr", ["lgr.h", "lgr_prover.h", "lgr_ubench.h"])
def test_public_headers_are_plain_c(hdr, tmp_path):
    """the drop-in boundary is a C ABI: every header under include/ must parse as strict C99 (what cgo / a C FFI sees)"""
    import shutil, subprocess
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    src = tmp_path / "hc.c"
    src.write_text('#include "%s"\nint main(void) { return 0; }\n' % hdr)
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only",
                        "-I", os.path.join(ROOT, "include"), str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_no_cpu_fallback_create_fails_loudly_without_gpu(lgr):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(lgr.LgrError):
        lgr.make_executor(64, 256)
    ctx = C.c_void_p()
    k = 256
    wk, w2k, wn = lgr.generate_omegas(k, 4 * k)
    rc = lgr.lib().lgr_create(C.byref(ctx), 0, 64, k, 4 * k, lgr.int_to_limbs(lgr.P), lgr.int_to_limbs(wk), lgr.int_to_limbs(w2k), lgr.int_to_limbs(wn))
    assert rc != 0 and lgr.lib().lgr_last_error()


def test_create_argument_validation(lgr):
    ctx = C.c_void_p()
    lib = lgr.lib()
    good = lgr.int_to_limbs(lgr.P)
    w = lgr.int_to_limbs(5)
    assert lib.lgr_create(C.byref(ctx), 0, 64, 250, 1000, good, w, w, w) == 1            # k not a power of two
    assert lib.lgr_create(C.byref(ctx), 0, 64, 256, 512, good, w, w, w) == 1             # n != 4k
    assert lib.lgr_create(C.byref(ctx), 0, 64, 256, 1024, lgr.int_to_limbs(lgr.P - 2), w, w, w) == 1   # wrong modulus
    assert b"modulus" in lib.lgr_last_error()


def test_product_path_never_touches_the_oracle():
    """the package, the C sources and the public header must not reference oracle/"""
    pkg = os.path.join(ROOT, "ligero-prover_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liblgo" not in text and "from oracle" not in text and "import oracle" not in text and "lgo_" not in text, os.path.join(dirpath, f)


def test_python_helpers(lgr):
    assert lgr.generate_omegas(8192, 32768)[2] == 0x1f67bc4574eaef5e630a13c710221a3e3d491e59fddabaf321e56f3ca8d91624
    vals = [0, 1, lgr.P - 1, 1 << 200]
    assert lgr.array_to_ints(lgr.ints_to_array(vals)) == vals
    assert pow(lgr.root_of_unity(12), 1 << 11, lgr.P) == lgr.P - 1
    assert lgr.root_of_unity(12) == pow(lgr.ROOT1, 1 << 16, lgr.P)


_WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.join(%(root)r, "tests"))
import conftest
lgr = conftest.load_package()
import importlib.util
spec = importlib.util.spec_from_file_location("lgr_sharding", os.path.join(%(root)r, "ligero-prover_b200", "sharding.py"))
sh = importlib.util.module_from_spec(spec); spec.loader.exec_module(sh)
from oracle import lgo
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
k, total = 16, 11
n = 4 * k
b, e = sh.shard_rows(total, world, rank)
rows = lgo.synth(3, b, e - b, k)                       # the oracle stands in for the GPU in this CPU test
dig, _, _ = lgo.encode_commit(rows, k)
local = torch.from_numpy(dig.view(np.int32).reshape(n, 8).copy())
allg = sh.gather_leaf_digests(local, world, dist)
leaves = allg.numpy().view(np.uint8).reshape(world * n, 32)
root = lgo.merkle_build(leaves)[0].tobytes().hex()
# every rank must hold the same tree; rank-major leaf order
roots = [None] * world
dist.all_gather_object(roots, root)
assert len(set(roots)) == 1
whole = lgo.synth(3, 0, total, k)
want = []
for g in range(world):
    bb, ee = sh.shard_rows(total, world, g)
    d, _, _ = lgo.encode_commit(whole[bb:ee], k)
    want.append(d)
assert np.array_equal(leaves, np.concatenate(want))
assert sh.leaf_index(1, 5, n) == n + 5
if rank == 0:
    print("ROOT", root)
dist.destroy_process_group()
"""


def test_world_size_2_gloo_sharded_commitment(tmp_path):
    """N>1 host logic (row sharding, rank-major digest all-gather, common tree) on gloo, world_size 2"""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="2")
    out = subprocess.check_output([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                                   "--master-port", "29533", str(script)], env=env, stderr=subprocess.STDOUT, timeout=300).decode()
    assert "ROOT" in out, out


_WORKER_EXACT = r"""
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
import importlib.util
spec = importlib.util.spec_from_file_location("lgr_sharding", os.path.join(%(root)r, "ligero-prover_b200", "sharding.py"))
sh = importlib.util.module_from_spec(spec); spec.loader.exec_module(sh)
from oracle import lgo
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
k, T, total = 16, 4, 4 * 5 + 3            # 6 tiles: rounds with a missing tile and a short last tile
n = 4 * k; slab = n // world

class OracleEngine:                        # the oracle stands in for the GPU in this CPU test
    def begin(self):
        self.sha = lgo.Sha(slab)
    def encode_round(self, rnd, rows, nrows):
        self.send = np.zeros((world, T, slab, 8), np.uint32)
        for r in range(nrows):
            cw = lgo.encode(rows[r], k)
            for h in range(world):
                self.send[h, r] = cw[h * slab:(h + 1) * slab]
    def exchange_and_hash(self, rnd, rows_per_rank, dist_):
        s = torch.from_numpy(self.send.view(np.int32).reshape(-1).copy()); r = torch.empty_like(s)
        dist_.all_to_all_single(r, s)
        recv = r.numpy().view(np.uint32).reshape(world, T, slab, 8)
        for h in range(world):
            for row in range(rows_per_rank[h]):
                self.sha.update(recv[h, row])
    def finish(self, dist_):
        local = torch.from_numpy(self.sha.final().view(np.int32).reshape(slab, 8).copy())
        return sh.gather_leaf_digests(local, world, dist_)

num_tiles = (total + T - 1) // T
mine = sh.tiles_of_rank(num_tiles, world, rank)
tiles = [(lgo.synth(3, t * T, min(T, total - t * T), k), min(T, total - t * T)) for t in mine]
leaves = sh.commit_exact(OracleEngine(), lambda i: tiles[i], total, T, world, rank, dist)
got = leaves.numpy().view(np.uint8).reshape(n, 32)
want, nodes, _ = lgo.encode_commit(lgo.synth(3, 0, total, k), k)
assert np.array_equal(got, want), "exact multi-rank layout differs from the single-device commitment"
if rank == 0:
    print("EXACT-ROOT", lgo.merkle_build(got)[0].tobytes().hex() == nodes[0].tobytes().hex())
dist.destroy_process_group()
"""


def test_world_size_2_gloo_exact_layout(tmp_path):
    """the bit-exact layout (round-robin tiles, all-to-all of column slabs, ordered absorption) on gloo"""
    script = tmp_path / "worker_exact.py"
    script.write_text(_WORKER_EXACT % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="2")
    out = subprocess.check_output([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                                   "--master-port", "29534", str(script)], env=env, stderr=subprocess.STDOUT, timeout=300).decode()
    assert "EXACT-ROOT True" in out, out


def test_shard_rows_partition():
    import importlib.util
    spec = importlib.util.spec_from_file_location("lgr_sharding", os.path.join(ROOT, "ligero-prover_b200", "sharding.py"))
    sh = importlib.util.module_from_spec(spec); spec.loader.exec_module(sh)
    for total in (0, 1, 7, 8, 1 << 22):
        for world in (1, 2, 3, 8):
            ranges = [sh.shard_rows(total, world, r) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            assert max(e - b for b, e in ranges) - min(e - b for b, e in ranges) <= 1


def test_wide_and_karatsuba_accumulators_under_emulation(emu, oracle):
    """fr.cuh's unreduced dot products (the tile combiners): 64 plain products, and the Karatsuba form (split at bit 127,
    three 128x128 products, recombined once per chunk) give sum a*b * 2^-288 mod p for 64-row chunks of edge and random values"""
    rng = random.Random(12)
    T, ncols = 64, 9
    edge = [0, 1, P - 1, (1 << 127) - 1, 1 << 127, (1 << 128) - 1, P >> 1, (1 << 253) + 5]
    a = [[edge[(t + c) % len(edge)] if c < 3 else rng.randrange(P) for c in range(ncols)] for t in range(T)]
    b = [[edge[(3 * t + c) % len(edge)] if c < 3 else rng.randrange(P) for c in range(ncols)] for t in range(T)]
    a[5] = [P - 1] * ncols; b[5] = [P - 1] * ncols
    A = oracle.to_limbs([x for row in a for x in row]); B = oracle.to_limbs([x for row in b for x in row])
    rinv = pow(1 << 288, -1, P)
    want = [sum(a[t][c] * b[t][c] for t in range(T)) * rinv % P for c in range(ncols)]
    for fn in (emu.emu_wide_dot, emu.emu_kara_dot):
        O = np.zeros((ncols, 8), np.uint32)
        fn(O.ctypes.data_as(C.c_void_p), A.ctypes.data_as(C.c_void_p), B.ctypes.data_as(C.c_void_p), C.c_size_t(T), C.c_size_t(ncols))
        got = oracle.from_limbs(O)
        assert all(g < 2 * P for g in got)
        assert [g % P for g in got] == want

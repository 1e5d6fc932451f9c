"""helpers shared by tests/test_refctx_cpu.py and tests/test_refctx_gpu.py: the golden vectors written by the REFERENCE's own
stage contexts / backend / interpreter (tests/golden/make_refctx_vectors.py)"""
import base64
import hashlib
import importlib.util
import json
import os
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CASES = ["i64_mul_k8192", "i64_mul3_k256", "vbn_k256", "mul64_k256", "arith32_k256", "intops_k256"]
WAT_TEXT = {"mul64": os.path.join(HERE, "golden", "mul64.wat"), "arith32": os.path.join(HERE, "golden", "arith32.wat"),
            "intops": os.path.join(HERE, "golden", "intops.wat")}          # cases whose program is a .wat file of the repo
REF_BIN_CPU = os.path.join(ROOT, "oracle", "_ref", "refctx_cpu")
REF_BIN_CUDA = os.path.join(ROOT, "oracle", "_ref", "refctx_cuda")


def _unz(s):
    return np.frombuffer(zlib.decompress(base64.b64decode(s)), np.uint32)


def load(case):
    fx = json.load(open(os.path.join(HERE, "golden", "refctx_%s.json" % case)))
    l = fx["l"]
    st = {
        "fx": fx, "l": l, "k": fx["k"], "n": fx["n"],
        "kinds": np.array(fx["kinds"], np.uint8),
        "values": _unz(fx["values_zb64"]).reshape(-1, l, 8),
        "coefs": _unz(fx["coefs_zb64"]).reshape(-1, l, 8),
        "args": _unz(fx["batch_args_zb64"]).reshape(-1, 3),
        "consts": _unz(fx["batch_consts_zb64"]).reshape(-1, 8),
        "const_sum": int.from_bytes(bytes.fromhex(fx["const_sum"]), "little"),
        "encoding_seed": bytes.fromhex(fx["encoding_seed"]),
        "instance_hash": bytes.fromhex(fx["instance_hash"]),
    }
    st["slots"] = int(st["args"][:, :2].max()) + 1 if len(st["args"]) else 0
    return st


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def compact_module():
    spec = importlib.util.spec_from_file_location("make_refctx_vectors", os.path.join(HERE, "golden", "make_refctx_vectors.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def check_envelope(env, fx, sibling_positions):
    """a parsed LigeroProofEnvelope (google.protobuf message, oracle/prover_ref.parse_envelope) against a reference-run
    fixture: root, the three test vectors, the sampled columns, the opened leaves and their sibling hashes"""
    pf = env.ligero_proof
    assert pf.merkle_tree.root.value.hex() == fx["root"]
    for name, key in (("encoded_code", "code"), ("encoded_linear", "linear"), ("encoded_quadratic", "quad")):
        assert sha(np.array(getattr(pf, name).values, np.uint32)) == fx["sha256"][key], name
    assert sha(np.array(pf.sampled_data.values, np.uint32)) == fx["sha256"]["samplings"]
    assert [int(i) for i in pf.merkle_tree.leaf_indices] == fx["sample_index"]
    # the envelope lists siblings in canonical order (proof_serializer.hpp:82-117); the reference's decommit holds them by position
    pos = sibling_positions(fx["sample_index"], fx["decommit_total"])
    assert sorted(pos) == fx["decommit_positions"]
    by_pos = dict(zip(pos, [s.value for s in pf.merkle_tree.sibling_hashes]))
    assert sha(np.frombuffer(b"".join(by_pos[p] for p in fx["decommit_positions"]), np.uint8)) == fx["sha256"]["decommit_siblings"]


# ---- folded WAT of the arithmetic-test subset -> the flat token stream tests/refctx/ref_contexts.cpp assembles ("ops:" programs)
def _sexpr(text):
    import re
    text = re.sub(r"\(;.*?;\)", " ", text, flags=re.S)
    text = re.sub(r";;[^\n]*", " ", text)
    toks = re.findall(r'"[^"]*"|[()]|[^\s()]+', text)
    pos = 0

    def parse():
        nonlocal pos
        t = toks[pos]; pos += 1
        if t != "(":
            return t
        out = []
        while toks[pos] != ")":
            out.append(parse())
        pos += 1
        return out
    return parse()


def wat_to_tokens(text):
    """'iNN.const <v>' / 'iNN.<op>' / 'call:<env function>', operands first -- what the folded text of the exported function denotes"""
    mod = _sexpr(text)
    imports = {f[3][1]: f[2].strip('"') for f in mod[1:] if f[0] == "import"}
    start = next(f[2][1] for f in mod[1:] if f[0] == "export" and f[1] == '"_start"')
    func = next(f for f in mod[1:] if f[0] == "func" and f[1] == start)
    out = []

    def lit(s):
        s = s.replace("_", "")
        v = int(s, 0)
        return v % (1 << 64)

    def emit(e):
        if e[0] in ("i64.const", "i32.const"):
            out.append("%s %d" % (e[0], lit(e[1]) % (1 << int(e[0][1:3]))))
        elif e[0] == "call":
            for a in e[2:]:
                emit(a)
            out.append("call:" + imports[e[1]])
        elif e[0][:4] in ("i32.", "i64.") or e[0] in ("drop", "nop"):
            for a in e[1:]:
                emit(a)
            out.append(e[0])
        else:
            raise ValueError("unsupported form " + str(e[0]))
    for e in func[2:]:
        if isinstance(e, list) and e[0] not in ("param", "result", "local", "type"):
            emit(e)
    return out


def run_reference_on_wat(text, k, seed_byte=7):
    """the reference's interpreter / backend / stage contexts over the CPU oracle on a program of the subset -> raw dict"""
    import subprocess
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        ops = os.path.join(tmp, "prog.ops")
        with open(ops, "w") as f:
            f.write("\n".join(wat_to_tokens(text)) + "\n")
        out = os.path.join(tmp, "out.json")
        subprocess.check_call([REF_BIN_CPU, "ops:" + ops, str(k), out, str(seed_byte)], stdout=subprocess.DEVNULL)
        return json.load(open(out))


def harness_args(case, tmpdir):
    """(program argument, k) for oracle/_ref/refctx_{cpu,cuda}: built-in programs by name, .wat programs as a token file"""
    prog, k = case.rsplit("_k", 1)
    if prog in WAT_TEXT:
        ops = os.path.join(tmpdir, prog + ".ops")
        with open(ops, "w") as f:
            f.write("\n".join(wat_to_tokens(open(WAT_TEXT[prog]).read())) + "\n")
        return "ops:" + ops, k, prog
    return prog, k, prog


# tests/i64_mul.wat of the reference (BASELINE config 4): the nine (a, b, a*b mod 2^64) cases of its single function,
# so that the program can be written out where /root/reference is absent
I64_MUL_CASES = [(1, 1, 1), (1, 0, 0), (2**64 - 1, 2**64 - 1, 1), (0x1000000000000000, 4096, 0), (0x8000000000000000, 0, 0),
                 (0x8000000000000000, 2**64 - 1, 0x8000000000000000), (0x7fffffffffffffff, 2**64 - 1, 0x8000000000000001),
                 (0x0123456789abcdef, 0xfedcba9876543210, 0x2236d88fe5618cf0), (0x7fffffffffffffff, 0x7fffffffffffffff, 1)]
WAT_HEAD = ('(module (import "env" "i64_private_const" (func $i64_private_const (param i64) (result i64)))\n'
            '(import "env" "assert_equal" (func $assert_equal (param i64 i64)))\n(func $t\n')
WAT_TAIL = ')\n(export "_start" (func $t)))\n'


WAT_HEAD32 = WAT_HEAD.replace("i64", "i32")


def binop_wat(op, cases):
    pc = lambda v: "(call $i64_private_const (i64.const %d))" % v
    return WAT_HEAD + "".join("(call $assert_equal (i64.%s %s %s) %s)\n" % (op, pc(a), pc(b), pc(c)) for a, b, c in cases) + WAT_TAIL


# ---- the reference's VERIFIER (src/webgpu_verifier.cpp:262-449 around nonbatch_verifier_context) on a proof from outside
def write_proof_file(path, root, code, linear, quad, samplings, positions, siblings, total_count, instance_hash):
    """the `key hex` lines tests/refctx/ref_contexts.cpp reads in verify: mode"""
    hx = lambda a: np.ascontiguousarray(a).tobytes().hex()
    with open(path, "w") as f:
        f.write("root %s\ninstance %s\ntotal %d\n" % (bytes(root).hex(), bytes(instance_hash).hex(), total_count))
        for name, v in (("code", code), ("linear", linear), ("quad", quad), ("samplings", samplings)):
            f.write("%s %s\n" % (name, hx(np.asarray(v, np.uint32))))
        for p, s in zip(positions, siblings):
            f.write("sibling %d %s\n" % (p, bytes(s).hex()))


def envelope_to_proof_file(path, env, sibling_positions, total_count, instance_hash):
    pf = env.ligero_proof
    leaf = [int(i) for i in pf.merkle_tree.leaf_indices]
    pos = sibling_positions(leaf, total_count)
    write_proof_file(path, pf.merkle_tree.root.value, list(pf.encoded_code.values), list(pf.encoded_linear.values), list(pf.encoded_quadratic.values),
                     list(pf.sampled_data.values), pos, [s.value for s in pf.merkle_tree.sibling_hashes], total_count, instance_hash)


def reference_verifier(binary, case, proof_path, tmpdir):
    """exit code and message of the reference's verifier run on `proof_path` for the program of `case`"""
    import subprocess
    prog, k, _ = harness_args(case, tmpdir)
    res = subprocess.run([binary, "verify:" + proof_path, prog, k], capture_output=True, text=True, timeout=900)
    return res.returncode, (res.stdout + res.stderr)[-2000:]


# ---- WebAssembly integer semantics (for expected values in generated programs) and a random program generator
UNARY_OPS = ["clz", "ctz", "popcnt", "eqz", "extend8_s", "extend16_s"]
BINARY_OPS = ["add", "sub", "mul", "and", "or", "xor", "shl", "shr_s", "shr_u", "rotl", "rotr", "eq", "ne",
              "lt_s", "lt_u", "gt_s", "gt_u", "le_s", "le_u", "ge_s", "ge_u", "div_s", "div_u", "rem_s", "rem_u"]


def wasm_op(op, w, a, b=None):
    """value of iW.op on unsigned operands a, b < 2^w (None where WebAssembly traps)"""
    M = 1 << w
    s = lambda v: v - M if v >> (w - 1) else v
    if op == "clz": return w - a.bit_length()
    if op == "ctz": return w if a == 0 else (a & -a).bit_length() - 1
    if op == "popcnt": return bin(a).count("1")
    if op == "eqz": return int(a == 0)
    if op in ("extend8_s", "extend16_s", "extend32_s"):
        n = int(op[6:-2]); v = a & ((1 << n) - 1)
        return (v - (1 << n) if v >> (n - 1) else v) % M
    if op == "add": return (a + b) % M
    if op == "sub": return (a - b) % M
    if op == "mul": return (a * b) % M
    if op == "and": return a & b
    if op == "or": return a | b
    if op == "xor": return a ^ b
    n = b % w if b is not None else 0
    if op == "shl": return (a << n) % M
    if op == "shr_u": return a >> n
    if op == "shr_s": return (s(a) >> n) % M
    if op == "rotl": return ((a << n) | (a >> (w - n))) % M if n else a
    if op == "rotr": return ((a >> n) | (a << (w - n))) % M if n else a
    if op == "eq": return int(a == b)
    if op == "ne": return int(a != b)
    if op[:2] in ("lt", "gt", "le", "ge"):
        x, y = (s(a), s(b)) if op.endswith("_s") else (a, b)
        return int({"lt": x < y, "gt": x > y, "le": x <= y, "ge": x >= y}[op[:2]])
    if op in ("div_u", "rem_u"):
        if b == 0: return None
        return a // b if op == "div_u" else a % b
    if op in ("div_s", "rem_s"):
        x, y = s(a), s(b)
        if y == 0 or (op == "div_s" and x == -(M >> 1) and y == -1): return None
        if y == -(M >> 1): return None      # (the reference's gadget compares |y| = 2^(w-1) as a SIGNED number and rejects its own witness)
        q = abs(x) // abs(y) * (1 if (x < 0) == (y < 0) else -1)
        return (q if op == "div_s" else x - q * y) % M
    raise ValueError(op)


def _rand_operand(rng, w):
    M = 1 << w
    return rng.choice([0, 1, 2, 3, M - 1, M >> 1, (M >> 1) - 1, 1 << (w // 2), 0x80, 0xff7f, rng.getrandbits(w), rng.getrandbits(w), rng.getrandbits(7), M - 1 - rng.getrandbits(5)])


def _typed(w, op, text):
    """predicates give an i32: widened where the program goes on in 64 bits, so that the module stays valid WebAssembly"""
    return "(i64.extend_i32_u %s)" % text if w == 64 and (op in ("eqz", "eq", "ne") or op[:2] in ("lt", "gt", "le", "ge")) else text


def rand_int_expr(rng, depth, w, ops=None):
    """(folded text, value) of a random expression of width w over every integer instruction: private and literal leaves,
    results that live as bit vectors (arithmetic, bitwise, shifts) and as single witnesses (counts, comparisons) mixed freely"""
    if depth == 0 or rng.random() < 0.2:
        v = _rand_operand(rng, w)
        lit = "(i%d.const %d)" % (w, v)
        return ("(call $i%d_private_const %s)" % (w, lit) if rng.random() < 0.75 else lit), v
    if ops is None and rng.random() < 0.1:                    # the value as ONE witness (env.witness_cast): a third kind of stack value
        t, v = rand_int_expr(rng, depth - 1, w, ops)
        return ("(call $cast%d %s)" % (w, t) if "private" in t else t), v
    for _ in range(100):
        op = rng.choice(ops or (UNARY_OPS + BINARY_OPS * 2))
        ta, va = rand_int_expr(rng, depth - 1, w, ops)
        if op in UNARY_OPS:
            if op == "extend16_s" and w == 64 and "private" not in ta:
                continue                                        # (the reference ZERO-extends a concrete i64 here, interpreter_impl.hpp:1208; the emitter follows it)
            return _typed(w, op, "(i%d.%s %s)" % (w, op, ta)), wasm_op(op, w, va)
        tb, vb = rand_int_expr(rng, depth - 1, w, ops)
        if op in ("div_s", "div_u", "rem_s", "rem_u") and "private" not in ta + tb:
            continue                                            # (both concrete: plain host division; nothing to check)
        v = wasm_op(op, w, va, vb)
        if v is not None:
            return _typed(w, op, "(i%d.%s %s %s)" % (w, op, ta, tb)), v
    raise RuntimeError("no valid expression found")


WAT_HEAD_BOTH = ('(module (import "env" "i32_private_const" (func $i32_private_const (param i32) (result i32)))\n'
                 '(import "env" "i64_private_const" (func $i64_private_const (param i64) (result i64)))\n'
                 '(import "env" "witness_cast_u64" (func $cast64 (param i64) (result i64)))\n'
                 '(import "env" "witness_cast_u32" (func $cast32 (param i32) (result i32)))\n'
                 '(import "env" "assert_equal" (func $assert_equal (param i64 i64)))\n(func $t\n')


def rand_int_program(rng, w, nexpr=3, depth=2, ops=None):
    exprs = [rand_int_expr(rng, rng.randrange(1, depth + 1), w, ops) for _ in range(nexpr)]
    rhs = lambda v: ("(i%d.const %d)" % (w, v)) if rng.random() < 0.5 else ("(call $i%d_private_const (i%d.const %d))" % (w, w, v))
    return WAT_HEAD_BOTH + "".join("(call $assert_equal %s %s)\n" % (t, rhs(v)) for t, v in exprs) + WAT_TAIL, exprs


# ---- a small assembler: the subset's text -> WebAssembly binary (no wabt here), for the binary front end's tests
def _uleb(v):
    out = bytearray()
    while True:
        b = v & 0x7f
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _sleb(v):
    out = bytearray()
    while True:
        b = v & 0x7f
        v >>= 7
        done = (v == 0 and not b & 0x40) or (v == -1 and b & 0x40)
        out.append(b | (0 if done else 0x80))
        if done:
            return bytes(out)


_INT_OPS = ["clz", "ctz", "popcnt", "add", "sub", "mul", "div_s", "div_u", "rem_s", "rem_u", "and", "or", "xor", "shl", "shr_s", "shr_u", "rotl", "rotr"]
_CMP_OPS = ["eqz", "eq", "ne", "lt_s", "lt_u", "gt_s", "gt_u", "le_s", "le_u", "ge_s", "ge_u"]
_OTHER_OPS = {"i32.wrap_i64": 0xA7, "i64.extend_i32_s": 0xAC, "i64.extend_i32_u": 0xAD, "i32.extend8_s": 0xC0, "i32.extend16_s": 0xC1,
              "i64.extend8_s": 0xC2, "i64.extend16_s": 0xC3, "i64.extend32_s": 0xC4}


def wat_to_wasm(text, custom_section=True):
    """binary module for a program of the subset: type, import, function, export and code sections (+ a custom section)"""
    mod = _sexpr(text)
    vt = {"i32": 0x7f, "i64": 0x7e}
    types, imports = [], []
    names = {}

    def functype(f):
        params = [vt[t] for part in f if isinstance(part, list) and part[0] == "param" for t in part[1:] if t in vt]
        results = [vt[t] for part in f if isinstance(part, list) and part[0] == "result" for t in part[1:] if t in vt]
        sig = (tuple(params), tuple(results))
        if sig not in types:
            types.append(sig)
        return types.index(sig)
    for f in mod[1:]:
        if f[0] == "import":
            names[f[3][1]] = len(imports)
            imports.append((f[1].strip('"'), f[2].strip('"'), functype(f[3])))
    start = next(f[2][1] for f in mod[1:] if f[0] == "export" and f[1] == '"_start"')
    func = next(f for f in mod[1:] if f[0] == "func" and f[1] == start)
    start_type = functype([])
    code = bytearray()

    def emit(e):
        if e[0] in ("i32.const", "i64.const"):
            w = int(e[0][1:3]); v = int(e[1].replace("_", ""), 0) % (1 << w)
            code.extend(bytes([0x41 if w == 32 else 0x42]) + _sleb(v - (1 << w) if v >> (w - 1) else v))
            return
        for a in (e[2:] if e[0] == "call" else e[1:]):
            emit(a)
        if e[0] == "call":
            code.extend(b"\x10" + _uleb(names[e[1]]))
        elif e[0] == "drop":
            code.append(0x1A)
        elif e[0] == "nop":
            code.append(0x01)
        elif e[0] in _OTHER_OPS:
            code.append(_OTHER_OPS[e[0]])
        else:
            w, op = e[0][:3], e[0][4:]
            if op in _CMP_OPS:
                code.append((0x45 if w == "i32" else 0x50) + _CMP_OPS.index(op))
            else:
                code.append((0x67 if w == "i32" else 0x79) + _INT_OPS.index(op))
    for e in func[2:]:
        if isinstance(e, list) and e[0] not in ("param", "result", "local", "type"):
            emit(e)
    code.append(0x0B)
    vec = lambda items: _uleb(len(items)) + b"".join(items)
    name = lambda s: _uleb(len(s.encode())) + s.encode()
    section = lambda sid, body: bytes([sid]) + _uleb(len(body)) + body
    out = b"\0asm\x01\0\0\0"
    out += section(1, vec([b"\x60" + vec([bytes([t]) for t in p]) + vec([bytes([t]) for t in r]) for p, r in types]))
    out += section(2, vec([name(m) + name(f) + b"\x00" + _uleb(t) for m, f, t in imports]))
    out += section(3, vec([_uleb(start_type)]))
    out += section(7, vec([name("_start") + b"\x00" + _uleb(len(imports))]))
    body = _uleb(0) + bytes(code)
    out += section(10, vec([_uleb(len(body)) + body]))
    if custom_section:
        out += section(0, name("producer") + b"tests/refctx_util.py")
    return out


def wat_to_plain(text):
    """the same module with the body of _start written as a plain instruction sequence instead of folded forms"""
    mod = _sexpr(text)
    imports = {f[3][1]: f[2].strip('"') for f in mod[1:] if f[0] == "import"}
    ids = {v: k for k, v in imports.items()}
    head = "(module\n" + "".join('(import "env" "%s" (func %s))\n' % (name, fid) for fid, name in imports.items())
    body = []
    for t in wat_to_tokens(text):
        body.append("call " + ids[t[5:]] if t.startswith("call:") else t)
    return head + "(func $plain\n" + "\n".join(body) + "\n)\n(export \"_start\" (func $plain)))\n"

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


def wat_module(text):
    """the parts of a module of the subset: imports (id -> (env name, function index)), functions (id, params, results, locals,
    body forms, local name -> index) in order, and the index of _start among the module's own functions"""
    mod = _sexpr(text)
    imports, funcs = {}, []
    for f in mod[1:]:
        if f[0] == "import":                                   # functions of wasi_snapshot_preview1 are named "wasi.<function>"
            imports[f[3][1]] = (("wasi." if f[1] == '"wasi_snapshot_preview1"' else "") + f[2].strip('"'), len(imports))
    for f in mod[1:]:
        if f[0] != "func":
            continue
        fn = {"id": f[1] if isinstance(f[1], str) else None, "params": [], "results": [], "locals": [], "names": {}, "body": []}
        for e in f[2 if fn["id"] else 1:]:
            if isinstance(e, list) and e[0] == "param" and fn.get("typed") and not fn["body"]:
                continue                                          # (func (type t) (param ..) ...): the parameters repeat the type
            if isinstance(e, list) and e[0] in ("param", "local") and not fn["body"]:
                rest = e[1:]
                if len(rest) == 2 and rest[0].startswith("$"):
                    fn["names"][rest[0]] = len(fn["params"]) + len(fn["locals"])
                    rest = rest[1:]
                fn["params" if e[0] == "param" else "locals"].extend(rest)
            elif isinstance(e, list) and e[0] == "result" and not fn["body"]:
                if not fn.get("typed"):
                    fn["results"].extend(e[1:])
            elif isinstance(e, list) and e[0] == "type" and not fn["body"]:          # (func $f (type $t) ...): the signature comes from the type
                sig = wat_tables(text)[0][e[1]]
                if not fn["params"] and not fn["results"]:
                    fn["params"], fn["results"] = list(sig[0]), list(sig[1])
                fn["typed"] = True
            else:
                fn["body"].append(e)
        funcs.append(fn)
    start_id = next(f[2][1] for f in mod[1:] if f[0] == "export" and f[1] == '"_start"')
    ids = {fn["id"]: k for k, fn in enumerate(funcs)}
    return imports, funcs, ids, ids[start_id]


def _wat_string(tok):
    """bytes of a WebAssembly text string literal"""
    out, i, body = bytearray(), 0, tok[1:-1]
    while i < len(body):
        ch = body[i]
        if ch != "\\":
            out += ch.encode(); i += 1
        elif body[i + 1] in "ntr":
            out.append({"n": 10, "t": 9, "r": 13}[body[i + 1]]); i += 2
        elif body[i + 1] in "\"'\\":
            out += body[i + 1].encode(); i += 2
        else:
            out.append(int(body[i + 1:i + 3], 16)); i += 3
    return bytes(out)


def wat_memory(text):
    """(pages, max pages or 0) or None, and the data segments [(id, active, offset, bytes)] of a module"""
    mod = _sexpr(text)
    memory, datas = None, []
    for f in mod[1:]:
        if f[0] == "memory":
            nums = [int(x, 0) for x in f[1:] if isinstance(x, str) and not x.startswith("$")]
            memory = (nums[0], nums[1] if len(nums) > 1 else 0)
        elif f[0] == "data":
            rest = f[1:]
            did = rest.pop(0) if rest and isinstance(rest[0], str) and rest[0].startswith("$") else None
            active, offset = False, 0
            if rest and isinstance(rest[0], list):
                off = rest.pop(0)
                off = off[1] if off[0] == "offset" else off
                active, offset = True, _lit(off[1]) % (1 << 32)
            datas.append((did, active, offset, b"".join(_wat_string(t) for t in rest)))
    return memory, datas


def _lit(s):
    return int(s.replace("_", ""), 0) % (1 << 64)


def wat_tables(text):
    """(named types {id: (params, results)}, table size or None, [(offset, [function ids], segment id or None)]) of a module"""
    types, table, elems = {}, None, []
    for f in _sexpr(text)[1:]:
        if f[0] == "type":                                     # by $name and by position ("0", "1", ...)
            fn = f[-1]
            sig = ([t for part in fn[1:] if part[0] == "param" for t in part[1:]], [t for part in fn[1:] if part[0] == "result" for t in part[1:]])
            types[str(sum(1 for k in types if k.isdigit()))] = sig
            if isinstance(f[1], str):
                types[f[1]] = sig
        elif f[0] == "table":
            table = int(next(x for x in f[1:] if isinstance(x, str) and x[0].isdigit()))
        elif f[0] == "elem":
            rest = f[1:]
            eid = rest.pop(0) if isinstance(rest[0], str) else None
            off = rest[0][1] if rest[0][0] == "offset" else rest[0]
            elems.append((_lit(off[1]) % (1 << 32), [x for x in rest[1:] if x != "func"], eid))
    return types, table, elems


def float_bits(lit, single):
    """bit pattern of an fNN.const literal: decimal / hexadecimal floating point, inf, nan, nan:0x<payload>"""
    import struct
    t = lit.replace("_", "")
    neg = t.startswith("-")
    t = t.lstrip("+-")
    sign = (1 << (31 if single else 63)) if neg else 0
    exp_all = 0x7F800000 if single else 0x7FF0000000000000
    if t == "inf":
        return sign | exp_all
    if t == "nan":
        return sign | exp_all | (1 << (22 if single else 51))
    if t.startswith("nan:0x"):
        return sign | exp_all | int(t[6:], 16)
    v = float.fromhex(t) if t[:2].lower() == "0x" else float(t)
    if single:
        try:
            raw = struct.pack("<f", v)
        except OverflowError:
            raw = struct.pack("<f", float("inf"))
        return sign | struct.unpack("<I", raw)[0]
    return sign | struct.unpack("<Q", struct.pack("<d", v))[0]


def wat_globals(text):
    """[(id or None, type, mutable, initial value)] of a module"""
    out = []
    for f in _sexpr(text)[1:]:
        if f[0] != "global":
            continue
        rest = f[1:]
        gid = rest.pop(0) if isinstance(rest[0], str) and rest[0].startswith("$") else None
        ty, init = rest
        mut = isinstance(ty, list)
        ty = ty[1] if mut else ty
        out.append((gid, ty, mut, _lit(init[1]) % (1 << int(ty[1:]))))
    return out


def _walk(fn, imports, ids, visit, data_ids=None, global_ids=None, types=None, elem_ids=None):
    """post-order walk of a function's body (folded forms, plain instructions, or both mixed): visit(kind, name, immediate)"""
    data_ids = data_ids or {}
    global_ids = global_ids or {}
    elem_ids = elem_ids or {}
    def local(x):
        return fn["names"][x] if x in fn["names"] else int(x)

    labels = []

    def depth(x):
        if x.startswith("$"):
            return len(labels) - 1 - max(i for i, nm in enumerate(labels) if nm == x)
        return int(x)

    def blocktype(rest):
        name = rest.pop(0) if rest and isinstance(rest[0], str) and rest[0].startswith("$") else None
        params, results = [], []
        while rest and isinstance(rest[0], list) and rest[0][0] in ("param", "result"):
            part = rest.pop(0)
            (params if part[0] == "param" else results).extend(part[1:])
        return name, params, results

    def type_use(parts):
        """(type $t) and / or (param ..) (result ..) -> (params, results)"""
        params, results = [], []
        for part in parts:
            if part[0] == "type":
                params, results = types[part[1]]
            else:
                (params if part[0] == "param" else results).extend(part[1:])
        return list(params), list(results)

    def is_label(x):
        return isinstance(x, str) and (x.startswith("$") or x[0].isdigit())

    def emit_seq(items):
        """a sequence in which folded forms and plain instructions (immediates after the instruction) may be mixed"""
        i = 0
        while i < len(items):
            e = items[i]; i += 1
            if isinstance(e, list):
                emit(e)
            elif e in ("block", "loop", "if"):
                rest = []
                while i < len(items) and ((isinstance(items[i], str) and items[i].startswith("$") and not rest) or (isinstance(items[i], list) and items[i][0] in ("param", "result"))):
                    rest.append(items[i]); i += 1
                name, params, results = blocktype(rest)
                visit("block", e, (params, results))
                labels.append(name)
            elif e in ("else", "end"):
                if i < len(items) and isinstance(items[i], str) and items[i].startswith("$"):
                    i += 1
                if e == "end":
                    labels.pop()
                visit("op", e, None)
            else:
                n = i
                if e == "br_table":
                    while n < len(items) and is_label(items[n]):
                        n += 1
                elif e == "call_indirect":
                    while n < len(items) and isinstance(items[n], list) and items[n][0] in ("type", "param", "result"):
                        n += 1
                elif e in ("table.get", "table.set", "table.size", "table.grow", "table.fill", "table.copy", "table.init", "elem.drop"):
                    while n < len(items) and is_label(items[n]) and n - i < 2:
                        n += 1
                elif e == "select":
                    while n < len(items) and isinstance(items[n], list) and items[n][0] == "result":
                        n += 1
                elif e[3:] in (".const",) or e in ("ref.null", "ref.func", "local.get", "local.set", "local.tee", "global.get", "global.set", "call", "br", "br_if", "memory.init", "data.drop"):
                    n += 1
                else:
                    while n < len(items) and isinstance(items[n], str) and items[n].partition("=")[0] in ("offset", "align"):
                        n += 1
                emit([e] + items[i:n])
                i = n

    def emit(e):
        h = e[0]
        if h in ("block", "loop"):
            rest = list(e[1:])
            name, params, results = blocktype(rest)
            visit("block", h, (params, results))
            labels.append(name)
            emit_seq(rest)
            labels.pop()
            visit("op", "end", None)
        elif h == "if":
            rest = list(e[1:])
            name, params, results = blocktype(rest)
            arms = [a for a in rest if isinstance(a, list) and a[0] in ("then", "else")]
            for a in rest:
                if a not in arms:
                    emit(a)
            visit("block", "if", (params, results))
            labels.append(name)
            emit_seq(arms[0][1:])
            if len(arms) > 1:
                visit("op", "else", None)
                emit_seq(arms[1][1:])
            labels.pop()
            visit("op", "end", None)
        elif h in ("br", "br_if"):
            for a in e[2:]:
                emit(a)
            visit("branch", h, depth(e[1]))
        elif h == "br_table":
            targets = [depth(x) for x in e[1:] if isinstance(x, str)]
            for a in e[1:]:
                if isinstance(a, list):
                    emit(a)
            visit("table", h, targets)
        elif h in ("return", "unreachable"):
            for a in e[1:]:
                emit(a)
            visit("op", h, None)
        elif h in ("i64.const", "i32.const"):
            visit("const", h, _lit(e[1]) % (1 << int(h[1:3])))
        elif h in ("f64.const", "f32.const"):
            visit("fconst", h, (float_bits(e[1], h == "f32.const"), e[1]))
        elif h in ("global.get", "global.set"):
            for a in e[2:]:
                emit(a)
            visit("global", h, global_ids[e[1]] if e[1] in global_ids else int(e[1]))
        elif h == "call":
            for a in e[2:]:
                emit(a)
            if e[1] in imports:
                visit("host", imports[e[1]][0], imports[e[1]][1])
            else:
                visit("callf", e[1], ids[e[1]])
        elif h == "call_indirect":
            use = [a for a in e[1:] if isinstance(a, list) and a[0] in ("type", "param", "result")]
            for a in e[1:]:
                if a not in use:
                    emit(a)
            visit("indirect", h, type_use(use))
        elif h in ("local.get", "local.set", "local.tee"):
            for a in e[2:]:
                emit(a)
            visit("local", h, local(e[1]))
        elif h == "ref.null":
            visit("op", h, None)
        elif h == "ref.func":
            visit("reffunc", h, ids[e[1]])
        elif h in ("table.get", "table.set", "table.size", "table.grow", "table.fill", "table.copy", "table.init", "elem.drop", "ref.is_null"):
            names = [a for a in e[1:] if isinstance(a, str)]
            for a in e[1:]:
                if isinstance(a, list):
                    emit(a)
            if h in ("table.init", "elem.drop"):
                visit("elemseg", h, elem_ids[names[-1]] if names[-1] in elem_ids else int(names[-1]))
            else:
                visit("op", h, None)
        elif h == "select":
            for a in e[1:]:
                if not (isinstance(a, list) and a[0] == "result"):
                    emit(a)
            visit("op", h, None)
        elif h[:4] in ("i32.", "i64.", "f32.", "f64.") and (h[4:8] == "load" or h[4:9] == "store"):
            offset, rest = 0, e[1:]
            while rest and isinstance(rest[0], str):
                key, _, val = rest.pop(0).partition("=")
                if key == "offset":
                    offset = int(val, 0)
            for a in rest:
                emit(a)
            visit("access", h, offset)
        elif h in ("memory.init", "data.drop"):
            for a in e[2:]:
                emit(a)
            visit("segment", h, data_ids[e[1]] if e[1] in data_ids else int(e[1]))
        elif h in ("memory.size", "memory.grow", "memory.fill", "memory.copy"):
            for a in e[1:]:
                emit(a)
            visit("op", h, None)
        elif h[:4] in ("i32.", "i64.", "f32.", "f64.") or h in ("drop", "nop"):
            for a in e[1:]:
                emit(a)
            visit("op", h, None)
        else:
            raise ValueError("unsupported form " + str(h))
    emit_seq(fn["body"])


def wat_to_tokens(text):
    """'iNN.const <v>' / 'iNN.<op>' / 'call:<env function>' / 'local.get <i>' / 'select' / 'callf <k>', operands first -- what the
    folded text denotes.  A module with several functions, parameters or locals gets 'func <params> <results> <locals>' headers
    and 'start <k>' (tests/refctx/ref_contexts.cpp: build_module)"""
    imports, funcs, ids, start = wat_module(text)
    memory, datas = wat_memory(text)
    data_ids = {d[0]: k for k, d in enumerate(datas) if d[0]}
    globals_ = wat_globals(text)
    global_ids = {g[0]: k for k, g in enumerate(globals_) if g[0]}
    out = ["global %s %d" % (ty, init) for _, ty, _, init in globals_]
    types, table, elems = wat_tables(text)
    if table is not None:
        out.append("table %d" % table)
    for off, fs, _ in elems:
        out.append("elem %d %d %s" % (off, len(fs), " ".join(str(ids[f]) for f in fs)))
    elem_ids = {eid: k for k, (_, _, eid) in enumerate(elems) if eid}
    if memory:
        out.append("memory %d %d" % memory)
    for _, active, offset, data in datas:
        out.append(("data active %d %s" % (offset, data.hex() or "-")) if active else ("data passive %s" % (data.hex() or "-")))

    def visit(kind, name, imm):
        if kind == "const":
            out.append("%s %d" % (name, imm))
        elif kind == "fconst":
            out.append("%s %d" % (name, imm[0]))
        elif kind == "global":
            out.append("%s %d" % (name, imm))
        elif kind == "indirect":
            out.append(name)
        elif kind in ("reffunc", "elemseg"):
            out.append("%s %d" % (name, imm))
        elif kind == "host":
            out.append("call:" + name)
        elif kind == "callf":
            out.append("callf %s" % imm)
        elif kind in ("local", "access", "segment", "branch"):
            out.append("%s %s" % (name, imm))
        elif kind == "block":
            out.append("%s %s %s" % (name, ",".join(imm[0]) or "-", ",".join(imm[1]) or "-"))
        elif kind == "table":
            out.append("%s %d %s" % (name, len(imm), " ".join(map(str, imm))))
        else:
            out.append(name)
    structured = len(funcs) > 1 or funcs[0]["params"] or funcs[0]["locals"]
    for fn in funcs:
        if structured:
            out.append("func %s %s %s" % tuple(",".join(fn[key]) or "-" for key in ("params", "results", "locals")))
        _walk(fn, imports, ids, visit, data_ids, global_ids, types, elem_ids)
    if structured:
        out.append("start %d" % start)
    return out


def run_reference_on_wat(text, k, seed_byte=7, args=None, private_indices=()):
    """the reference's interpreter / backend / stage contexts over the CPU oracle on a program of the subset -> raw dict.
    args: the guest's argument byte strings (for its wasi_preview1 module), private_indices: the secret ones"""
    import subprocess
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        ops = os.path.join(tmp, "prog.ops")
        with open(ops, "w") as f:
            for i, a in enumerate(args or ()):
                f.write("arg %s %s\n" % ("private" if i in private_indices else "public", bytes(a).hex() or "-"))
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


_FLOAT_ARITH = ["abs", "neg", "ceil", "floor", "trunc", "nearest", "sqrt", "add", "sub", "mul", "div", "min", "max", "copysign"]
_FLOAT_CMP = ["eq", "ne", "lt", "gt", "le", "ge"]
_FLOAT_OPS = {}
for _k, _ty in enumerate(("f32", "f64")):
    for _j, _op in enumerate(_FLOAT_CMP):
        _FLOAT_OPS["%s.%s" % (_ty, _op)] = bytes([0x5B + 6 * _k + _j])
    for _j, _op in enumerate(_FLOAT_ARITH):
        _FLOAT_OPS["%s.%s" % (_ty, _op)] = bytes([0x8B + 14 * _k + _j])
for _j, _nm in enumerate(["i32.trunc_f32_s", "i32.trunc_f32_u", "i32.trunc_f64_s", "i32.trunc_f64_u", None, None, "i64.trunc_f32_s", "i64.trunc_f32_u", "i64.trunc_f64_s",
                          "i64.trunc_f64_u", "f32.convert_i32_s", "f32.convert_i32_u", "f32.convert_i64_s", "f32.convert_i64_u", "f32.demote_f64", "f64.convert_i32_s",
                          "f64.convert_i32_u", "f64.convert_i64_s", "f64.convert_i64_u", "f64.promote_f32", "i32.reinterpret_f32", "i64.reinterpret_f64",
                          "f32.reinterpret_i32", "f64.reinterpret_i64"]):
    if _nm:
        _FLOAT_OPS[_nm] = bytes([0xA8 + _j])
for _j, _nm in enumerate(["i32.trunc_sat_f32_s", "i32.trunc_sat_f32_u", "i32.trunc_sat_f64_s", "i32.trunc_sat_f64_u", "i64.trunc_sat_f32_s", "i64.trunc_sat_f32_u",
                          "i64.trunc_sat_f64_s", "i64.trunc_sat_f64_u"]):
    _FLOAT_OPS[_nm] = bytes([0xFC, _j])


def wat_to_wasm(text, custom_section=True):
    """binary module for a program of the subset: type, import, function, export and code sections (+ a custom section)"""
    mod = _sexpr(text)
    imports, funcs, ids, start = wat_module(text)
    memory, datas = wat_memory(text)
    data_ids = {d[0]: k for k, d in enumerate(datas) if d[0]}
    globals_ = wat_globals(text)
    global_ids = {g[0]: k for k, g in enumerate(globals_) if g[0]}
    named_types, table, elems = wat_tables(text)
    access = ["i32.load", "i64.load", "f32.load", "f64.load", "i32.load8_s", "i32.load8_u", "i32.load16_s", "i32.load16_u", "i64.load8_s", "i64.load8_u", "i64.load16_s",
              "i64.load16_u", "i64.load32_s", "i64.load32_u", "i32.store", "i64.store", "f32.store", "f64.store", "i32.store8", "i32.store16", "i64.store8", "i64.store16", "i64.store32"]
    vt = {"i32": 0x7f, "i64": 0x7e, "f32": 0x7d, "f64": 0x7c, "funcref": 0x70, "externref": 0x6f}
    elem_ids = {eid: k for k, (_, _, eid) in enumerate(elems) if eid}
    table_ops = {"table.get": b"\x25\x00", "table.set": b"\x26\x00", "table.copy": b"\xfc\x0e\x00\x00", "table.grow": b"\xfc\x0f\x00", "table.size": b"\xfc\x10\x00",
                 "table.fill": b"\xfc\x11\x00", "ref.null": b"\xd0\x70", "ref.is_null": b"\xd1"}
    types, import_list = [], []

    def typeidx(params, results):
        sig = (tuple(vt[t] for t in params), tuple(vt[t] for t in results))
        if sig not in types:
            types.append(sig)
        return types.index(sig)
    for f in mod[1:]:
        if f[0] == "import":
            params = [t for part in f[3] if isinstance(part, list) and part[0] == "param" for t in part[1:] if t in vt]
            results = [t for part in f[3] if isinstance(part, list) and part[0] == "result" for t in part[1:] if t in vt]
            import_list.append((f[1].strip('"'), f[2].strip('"'), typeidx(params, results)))
    func_types = [typeidx(fn["params"], fn["results"]) for fn in funcs]
    vec = lambda items: _uleb(len(items)) + b"".join(items)
    name = lambda s: _uleb(len(s.encode())) + s.encode()
    section = lambda sid, body: bytes([sid]) + _uleb(len(body)) + body
    bodies = []
    for fn in funcs:
        code = bytearray()

        def visit(kind, nm, imm):
            if kind == "const":
                w = int(nm[1:3])
                code.extend(bytes([0x41 if w == 32 else 0x42]) + _sleb(imm - (1 << w) if imm >> (w - 1) else imm))
            elif kind == "fconst":
                code.extend(b"\x43" + imm[0].to_bytes(4, "little") if nm == "f32.const" else b"\x44" + imm[0].to_bytes(8, "little"))
            elif kind == "global":
                code.extend(bytes([0x23 if nm == "global.get" else 0x24]) + _uleb(imm))
            elif kind == "indirect":
                code.extend(b"\x11" + _uleb(typeidx(*imm)) + b"\x00")
            elif kind == "reffunc":
                code.extend(b"\xd2" + _uleb(len(import_list) + imm))
            elif kind == "elemseg":
                code.extend(b"\xfc" + (_uleb(12) + _uleb(imm) + b"\x00" if nm == "table.init" else _uleb(13) + _uleb(imm)))
            elif nm in table_ops:
                code.extend(table_ops[nm])
            elif nm in _FLOAT_OPS:
                code.extend(_FLOAT_OPS[nm])
            elif kind == "host":
                code.extend(b"\x10" + _uleb(imm))
            elif kind == "callf":
                code.extend(b"\x10" + _uleb(len(import_list) + imm))
            elif kind == "local":
                code.extend(bytes([{"local.get": 0x20, "local.set": 0x21, "local.tee": 0x22}[nm]]) + _uleb(imm))
            elif kind == "block":
                params, results = imm
                bt = b"\x40" if not params and not results else (bytes([vt[results[0]]]) if not params and len(results) == 1 else _sleb(typeidx(params, results)))
                code.extend(bytes([{"block": 0x02, "loop": 0x03, "if": 0x04}[nm]]) + bt)
            elif kind == "branch":
                code.extend(bytes([0x0C if nm == "br" else 0x0D]) + _uleb(imm))
            elif kind == "table":
                code.extend(b"\x0e" + _uleb(len(imm) - 1) + b"".join(_uleb(t) for t in imm))
            elif nm in ("else", "end", "return", "unreachable"):
                code.append({"else": 0x05, "end": 0x0B, "return": 0x0F, "unreachable": 0x00}[nm])
            elif kind == "access":
                code.extend(bytes([0x28 + access.index(nm)]) + _uleb(0) + _uleb(imm))
            elif kind == "segment":
                code.extend(b"\xfc" + (_uleb(8) + _uleb(imm) + b"\x00" if nm == "memory.init" else _uleb(9) + _uleb(imm)))
            elif nm in ("memory.size", "memory.grow"):
                code.extend(bytes([0x3F if nm == "memory.size" else 0x40, 0]))
            elif nm in ("memory.copy", "memory.fill"):
                code.extend(b"\xfc" + (_uleb(10) + b"\x00\x00" if nm == "memory.copy" else _uleb(11) + b"\x00"))
            elif nm in ("drop", "nop", "select"):
                code.append({"drop": 0x1A, "nop": 0x01, "select": 0x1B}[nm])
            elif nm in _OTHER_OPS:
                code.append(_OTHER_OPS[nm])
            else:
                w, op = nm[:3], nm[4:]
                code.append((0x45 if w == "i32" else 0x50) + _CMP_OPS.index(op) if op in _CMP_OPS else (0x67 if w == "i32" else 0x79) + _INT_OPS.index(op))
        _walk(fn, imports, ids, visit, data_ids, global_ids, named_types, elem_ids)
        code.append(0x0B)
        body = vec([_uleb(1) + bytes([vt[t]]) for t in fn["locals"]]) + bytes(code)
        bodies.append(_uleb(len(body)) + body)
    out = b"\0asm\x01\0\0\0"
    out += section(1, vec([b"\x60" + vec([bytes([t]) for t in p]) + vec([bytes([t]) for t in r]) for p, r in types]))
    out += section(2, vec([name(m) + name(f) + b"\x00" + _uleb(t) for m, f, t in import_list]))
    out += section(3, vec([_uleb(t) for t in func_types]))
    if table is not None:
        out += section(4, vec([b"\x70\x00" + _uleb(table)]))
    if memory:
        out += section(5, vec([(b"\x01" + _uleb(memory[0]) + _uleb(memory[1])) if memory[1] else (b"\x00" + _uleb(memory[0]))]))
    if globals_:
        out += section(6, vec([bytes([vt[ty], int(mut), 0x41 if ty == "i32" else 0x42]) + _sleb(init - (1 << int(ty[1:])) if init >> (int(ty[1:]) - 1) else init) + b"\x0b"
                               for _, ty, mut, init in globals_]))
    out += section(7, vec([name("_start") + b"\x00" + _uleb(len(import_list) + start)]))
    if elems:
        out += section(9, vec([b"\x00\x41" + _sleb(off) + b"\x0b" + vec([_uleb(len(import_list) + ids[f]) for f in fs]) for off, fs, _ in elems]))
    if datas:
        out += section(12, _uleb(len(datas)))
    out += section(10, vec(bodies))
    if datas:
        out += section(11, vec([(b"\x00\x41" + _sleb(off - (1 << 32) if off >> 31 else off) + b"\x0b" if active else b"\x01") + _uleb(len(data)) + data for _, active, off, data in datas]))
    if custom_section:
        out += section(0, name("producer") + b"tests/refctx_util.py")
    return out


def wat_to_plain(text):
    """the same module with every function body written as a plain instruction sequence instead of folded forms"""
    imports, funcs, ids, start = wat_module(text)
    by_index = {index: fid for fid, (_, index) in imports.items()}
    memory, datas = wat_memory(text)
    data_ids = {d[0]: k for k, d in enumerate(datas) if d[0]}
    out = ["(module"] + ['(import "%s" "%s" (func %s))' % ("wasi_snapshot_preview1" if nm.startswith("wasi.") else "env", nm.split(".")[-1], fid) for fid, (nm, _) in imports.items()]
    globals_ = wat_globals(text)
    global_ids = {g[0]: k for k, g in enumerate(globals_) if g[0]}
    for _, ty, mut, init in globals_:
        out.append("(global %s (%s.const %d))" % ("(mut %s)" % ty if mut else ty, ty, init))
    named_types, table, elems = wat_tables(text)
    if table is not None:
        out.append("(table %d funcref)" % table)
    for off, fs, _ in elems:
        out.append("(elem (i32.const %d) func %s)" % (off, " ".join(fs)))
    elem_ids = {eid: k for k, (_, _, eid) in enumerate(elems) if eid}
    if memory:
        out.append("(memory %d%s)" % (memory[0], " %d" % memory[1] if memory[1] else ""))
    for _, active, off, data in datas:
        out.append("(data %s\"%s\")" % ("(i32.const %d) " % off if active else "", "".join("\\%02x" % b for b in data)))
    for k, fn in enumerate(funcs):
        fid = fn["id"] or "$f%d" % k
        head = "(func %s" % fid + "".join(" (param %s)" % t for t in fn["params"]) + "".join(" (result %s)" % t for t in fn["results"]) + "".join(" (local %s)" % t for t in fn["locals"])
        body = []

        def visit(kind, nm, imm):
            if kind == "const":
                body.append("%s %d" % (nm, imm))
            elif kind == "fconst":
                body.append("%s %s" % (nm, imm[1]))
            elif kind == "global":
                body.append("%s %d" % (nm, imm))
            elif kind == "reffunc":
                body.append("ref.func %s" % (funcs[imm]["id"] or "$f%d" % imm))
            elif kind == "elemseg":
                body.append("%s %d" % (nm, imm))
            elif nm == "ref.null":
                body.append("ref.null func")
            elif kind == "indirect":
                body.append("call_indirect%s%s" % ("".join(" (param %s)" % t for t in imm[0]), "".join(" (result %s)" % t for t in imm[1])))
            elif kind == "host":
                body.append("call %s" % by_index.get(imm))
            elif kind == "callf":
                body.append("call %s" % nm)
            elif kind in ("local", "segment", "branch"):
                body.append("%s %s" % (nm, imm))
            elif kind == "access":
                body.append("%s offset=%s" % (nm, imm))
            elif kind == "block":
                body.append("%s%s%s" % (nm, "".join(" (param %s)" % t for t in imm[0]), "".join(" (result %s)" % t for t in imm[1])))
            elif kind == "table":
                body.append("%s %s" % (nm, " ".join(map(str, imm))))
            else:
                body.append(nm)
        _walk(fn, imports, ids, visit, data_ids, global_ids, named_types, elem_ids)
        out.append(head + "\n" + "\n".join(body) + "\n)")
    out.append('(export "_start" (func %s)))' % (funcs[start]["id"] or "$f%d" % start))
    return "\n".join(out) + "\n"


# ---- programs with locals, select and module functions on top of the integer instructions
def _helpers(w):
    """module functions of width w: (text, python model)"""
    W = "i%d" % w
    g = lambda n: "(local.get $%s)" % n
    return [
        ("(func $h0 (param $a %s) (param $b %s) (result %s) (%s.xor (%s.mul %s %s) (%s.add %s %s)))" % (W, W, W, W, W, g("a"), g("b"), W, g("a"), g("b")),
         lambda a, b: wasm_op("xor", w, wasm_op("mul", w, a, b), wasm_op("add", w, a, b))),
        ("(func $h1 (param $a %s) (param $b %s) (result %s) (local $t %s) (local.set $t (%s.popcnt %s)) (%s.add (local.get $t) (%s.clz %s)))"
         % (W, W, W, W, W, g("a"), W, W, g("b")),
         lambda a, b: wasm_op("add", w, wasm_op("popcnt", w, a), wasm_op("clz", w, b))),
        ("(func $h2 (param $a %s) (param $b %s) (result %s) (select %s (call $h0 %s %s) (%s.lt_u %s %s)))" % (W, W, W, g("a"), g("b"), g("a"), W, g("a"), g("b")),
         lambda a, b: a if wasm_op("lt_u", w, a, b) else wasm_op("xor", w, wasm_op("mul", w, b, a), wasm_op("add", w, b, a))),
    ]


def rand_struct_expr(rng, depth, w, env, helpers):
    """like rand_int_expr, plus local.get / local.tee of the locals in `env` (name -> current value, updated in evaluation
    order), select on concrete and private conditions, and calls of the helper functions"""
    r = rng.random()
    if depth > 0 and r < 0.15:
        k = rng.randrange(len(helpers))
        (ta, va), (tb, vb) = rand_struct_expr(rng, depth - 1, w, env, helpers), rand_struct_expr(rng, depth - 1, w, env, helpers)
        return "(call $h%d %s %s)" % (k, ta, tb), helpers[k][1](va, vb)
    if depth > 0 and r < 0.30:
        (ta, va), (tb, vb) = rand_struct_expr(rng, depth - 1, w, env, helpers), rand_struct_expr(rng, depth - 1, w, env, helpers)
        c = rng.randrange(2)
        form = rng.randrange(3)
        cond = ["(i32.const %d)" % (c * 5), "(call $i32_private_const (i32.const %d))" % (c * 9), None][form]
        if cond is None:
            tc, vc = rand_struct_expr(rng, depth - 1, w, env, helpers)
            cond, c = "(i%d.eqz %s)" % (w, tc), int(vc == 0)
        return "(select %s %s %s)" % (ta, tb, cond), (va if c else vb)
    if r < 0.45 and env:
        name = rng.choice(sorted(env))
        return "(local.get $%s)" % name, env[name]
    if depth > 0 and r < 0.55 and env:
        name = rng.choice(sorted(env))
        t, v = rand_struct_expr(rng, depth - 1, w, env, helpers)
        env[name] = v
        return "(local.tee $%s %s)" % (name, t), v
    if depth == 0 or r < 0.65:
        return rand_int_expr(rng, 0, w)
    for _ in range(100):
        op = rng.choice(UNARY_OPS + BINARY_OPS * 2)
        saved = dict(env)
        ta, va = rand_struct_expr(rng, depth - 1, w, env, helpers)
        if op in UNARY_OPS:
            if op == "extend16_s" and w == 64 and "private" not in ta and "local" not in ta and "call" not in ta:
                env.clear(); env.update(saved)
                continue
            return _typed(w, op, "(i%d.%s %s)" % (w, op, ta)), wasm_op(op, w, va)
        tb, vb = rand_struct_expr(rng, depth - 1, w, env, helpers)
        v = wasm_op(op, w, va, vb)
        concrete = not any(key in ta + tb for key in ("private", "local", "call"))
        if v is not None and not (op in ("div_s", "div_u", "rem_s", "rem_u") and concrete):
            return _typed(w, op, "(i%d.%s %s %s)" % (w, op, ta, tb)), v
        env.clear(); env.update(saved)
    raise RuntimeError("no valid expression found")


def rand_struct_program(rng, w, nstmt=5, depth=2):
    W = "i%d" % w
    helpers = _helpers(w)
    env = {"x": 0, "y": 0, "z": 0}
    body = []
    for _ in range(nstmt):
        t, v = rand_struct_expr(rng, rng.randrange(1, depth + 1), w, env, helpers)
        if rng.random() < 0.45:
            name = rng.choice(sorted(env))
            env[name] = v
            body.append("(local.set $%s %s)" % (name, t))
        else:
            rhs = ("(%s.const %d)" % (W, v)) if rng.random() < 0.5 else ("(call $%s_private_const (%s.const %d))" % (W, W, v))
            body.append("(call $assert_equal %s %s)" % (t, rhs))
    head = WAT_HEAD_BOTH[:WAT_HEAD_BOTH.index("(func $t")]
    return (head + "\n".join(h[0] for h in helpers) + "\n(func $t (local $x %s) (local $y %s) (local $z %s)\n" % (W, W, W) + "\n".join(body) + "\n" + WAT_TAIL)


# ---- programs over linear memory: stores of witnesses and numbers, loads of every width, fill / copy / init
def rand_memory_program(rng, nstmt=14):
    """random stores, loads (asserted against a byte-array model), memory.fill / copy / init over a 96-byte window, so that
    marked ranges are split, joined, overwritten, copied over themselves and cleared in every order"""
    mem = bytearray(65536)
    seg = bytes(rng.getrandbits(8) for _ in range(12))
    body = []
    P = lambda w, v: "(call $i%d_private_const (i%d.const %d))" % (w, w, v)
    L = lambda w, v: "(i%d.const %d)" % (w, v)
    addr = lambda a: L(32, a) if rng.random() < 0.8 else P(32, a)
    for _ in range(nstmt):
        r = rng.random()
        if r < 0.35:
            w = rng.choice([32, 64])
            nbytes = rng.choice([1, 2, w // 8] + ([4] if w == 64 else []))
            a, off = rng.randrange(0, 80), rng.choice([0, 0, 3, 8])
            v = rng.getrandbits(w)
            val = P(w, v) if rng.random() < 0.6 else L(w, v)
            if rng.random() < 0.2:
                val = "(i%d.add %s %s)" % (w, val, L(w, 1)); v = (v + 1) % (1 << w)
            name = "store" if nbytes == w // 8 else "store%d" % (8 * nbytes)
            body.append("(i%d.%s%s %s %s)" % (w, name, " offset=%d" % off if off else "", addr(a), val))
            mem[a + off:a + off + nbytes] = (v % (1 << (8 * nbytes))).to_bytes(nbytes, "little")
        elif r < 0.75:
            w = rng.choice([32, 64])
            nbytes = rng.choice([1, 2, w // 8] + ([4] if w == 64 else []))
            signed = rng.random() < 0.5
            a, off = rng.randrange(0, 80), rng.choice([0, 0, 5])
            v = int.from_bytes(mem[a + off:a + off + nbytes], "little")
            if nbytes < w // 8 and signed and v >> (8 * nbytes - 1):
                v -= 1 << (8 * nbytes)
            name = "load" if nbytes == w // 8 else "load%d_%s" % (8 * nbytes, "s" if signed else "u")
            rhs = L(w, v % (1 << w)) if rng.random() < 0.6 else P(w, v % (1 << w))
            body.append("(call $assert_equal (i%d.%s%s %s) %s)" % (w, name, " offset=%d" % off if off else "", addr(a), rhs))
        elif r < 0.83:
            d, n, val = rng.randrange(0, 88), rng.randrange(0, 9), rng.getrandbits(9)
            body.append("(memory.fill (i32.const %d) (i32.const %d) (i32.const %d))" % (d, val, n))
            mem[d:d + n] = bytes([val & 0xff]) * n
        elif r < 0.95:
            d, sA, n = rng.randrange(0, 80), rng.randrange(0, 80), rng.randrange(0, 17)
            body.append("(memory.copy (i32.const %d) (i32.const %d) (i32.const %d))" % (d, sA, n))
            mem[d:d + n] = bytes(mem[sA:sA + n])
        else:
            d, sA, n = rng.randrange(0, 88), rng.randrange(0, 8), rng.randrange(0, 5)
            body.append("(memory.init $seg (i32.const %d) (i32.const %d) (i32.const %d))" % (d, sA, n))
            mem[d:d + n] = seg[sA:sA + n]
    head = WAT_HEAD_BOTH[:WAT_HEAD_BOTH.index("(func $t")]
    return (head + '(memory 1)\n(data $seg "%s")\n(func $t\n' % "".join("\\%02x" % b for b in seg) + "\n".join(body) + "\n" + WAT_TAIL)


# ---- programs with control flow: if / else, blocks left by br / br_if / br_table, counted loops, early returns
def rand_cf_program(rng, w, nstmt=5, depth=2):
    """statements of rand_struct_program plus structured control flow around them.  Arms that are not taken and code
    skipped by a branch are generated on a scratch copy of the model, so their text is arbitrary but never runs"""
    W = "i%d" % w
    helpers = _helpers(w)
    M = 1 << w
    early = ("(func $h3 (param $a %s) (param $b %s) (result %s) (local $t %s)\n (local.set $t (%s.mul (local.get $a) (local.get $a)))\n"
             " (if (%s.gt_u (local.get $a) (local.get $b)) (then (return (local.get $t))))\n (%s.add (local.get $t) (local.get $b)))" % (W, W, W, W, W, W, W))
    env = {"x": 0, "y": 0, "z": 0}
    P32 = lambda v: "(call $i32_private_const (i32.const %d))" % v

    def cond(env):
        """(text of an i32 condition, truth)"""
        r = rng.random()
        if r < 0.3:
            c = rng.randrange(2)
            return "(i32.const %d)" % (c * 3), bool(c)
        if r < 0.6:
            c = rng.randrange(2)
            return P32(c * 7), bool(c)
        t, v = rand_struct_expr(rng, 1, w, env, helpers)
        return "(%s.eqz %s)" % (W, t), v == 0

    def simple(env):
        t, v = rand_struct_expr(rng, rng.randrange(1, depth + 1), w, env, helpers)
        if rng.random() < 0.5:
            name = rng.choice(sorted(env))
            env[name] = v
            return "(local.set $%s %s)" % (name, t)
        rhs = ("(%s.const %d)" % (W, v)) if rng.random() < 0.5 else ("(call $%s_private_const (%s.const %d))" % (W, W, v))
        return "(call $assert_equal %s %s)" % (t, rhs)

    def stmts(env, n, level):
        return " ".join(stmt(env, level) for _ in range(n))

    def stmt(env, level):
        r = rng.random()
        if level >= 2 or r < 0.4:
            return simple(env)
        if r < 0.55:                                             # if / else
            ct, c = cond(env)
            taken, other = dict(env), dict(env)
            a = stmts(taken if c else other, rng.randrange(1, 3), level + 1)
            b = stmts(other if c else taken, rng.randrange(0, 3), level + 1)
            env.clear(); env.update(taken)
            return "(if %s (then %s)%s)" % (ct, a, " (else %s)" % b if b or rng.random() < 0.5 else "")
        if r < 0.68:                                             # a block left early
            first = stmts(env, rng.randrange(0, 2), level + 1)
            ct, c = cond(env)
            scratch = dict(env)
            rest = stmts(scratch if c else env, rng.randrange(1, 3), level + 1)
            return "(block $skip %s (br_if $skip %s) %s)" % (first, ct, rest)
        if r < 0.78:                                             # a block whose result comes from a branch that drops what lies above the label
            (t1, _), (t2, _), (t3, v3) = (rand_struct_expr(rng, 1, w, env, helpers) for _ in range(3))
            name = rng.choice(sorted(env))
            env[name] = v3
            return "(local.set $%s (block $v (result %s) %s %s %s (br $v)))" % (name, W, t1, t2, t3)
        if r < 0.86:                                             # br_table over nested blocks
            k = rng.randrange(4)
            idx = "(i32.const %d)" % k if rng.random() < 0.5 else P32(k)
            # (block $a (block $b (block $c (br_table $c $b $a $c idx)) S_c) S_b) S_a : leaving $c runs S_c S_b S_a, $b runs S_b S_a, $a runs S_a
            target = ["$c", "$b", "$a", "$c"][min(k, 3)]
            run_c, run_b = target == "$c", target in ("$c", "$b")
            sc = stmts(env if run_c else dict(env), 1, level + 1)
            sb = stmts(env if run_b else dict(env), 1, level + 1)
            return "(block $a (block $b (block $c (br_table $c $b $a $c %s)) %s) %s)" % (idx, sc, sb)
        if r < 0.94:                                             # a counted loop accumulating into a local
            n = rng.randrange(1, 4)
            name = rng.choice(sorted(env))
            op = rng.choice(["add", "sub", "mul", "xor", "or", "rotl"])
            leaf_v = _rand_operand(rng, w)
            leaf = "(call $%s_private_const (%s.const %d))" % (W, W, leaf_v) if rng.random() < 0.6 else "(%s.const %d)" % (W, leaf_v)
            for _ in range(n):
                env[name] = wasm_op(op, w, env[name], leaf_v)
            return ("(local.set $i (i32.const 0)) (block $done (loop $again (br_if $done (i32.ge_u (local.get $i) (i32.const %d))) "
                    "(local.set $%s (%s.%s (local.get $%s) %s)) (local.set $i (i32.add (local.get $i) (i32.const 1))) (br $again)))" % (n, name, W, op, name, leaf))
        (ta, va), (tb, vb) = rand_struct_expr(rng, 1, w, env, helpers), rand_struct_expr(rng, 1, w, env, helpers)   # early return in a callee
        v = (va * va) % M if va > vb else ((va * va) % M + vb) % M
        return "(call $assert_equal (call $h3 %s %s) (%s.const %d))" % (ta, tb, W, v)

    body = [stmt(env, 0) for _ in range(nstmt)]
    head = WAT_HEAD_BOTH[:WAT_HEAD_BOTH.index("(func $t")]
    return (head + "\n".join(h[0] for h in helpers) + "\n" + early + "\n(func $t (local $x %s) (local $y %s) (local $z %s) (local $i i32)\n" % (W, W, W)
            + "\n".join(body) + "\n" + WAT_TAIL)


# ---- floating point and globals: numbers only in the reference; a result is observed by committing it
_F_SPECIAL = ["0.0", "-0.0", "inf", "-inf", "nan", "-nan", "nan:0x1", "1.5", "-2.25", "0x1.8p3", "1e30", "-1e-30", "3.0", "0.1", "16777217.0",
              "4294967296.0", "-2147483649.0", "9223372036854775808.0", "123456789.125", "0x1p-126", "0x1p-1022", "2.5", "3.5", "-0.5"]


def rand_float_expr(rng, depth, t, env):
    """text of a random expression of type t ('f32' / 'f64'); env: the locals per type"""
    r = rng.random()
    if depth == 0 or r < 0.2:
        r2 = rng.random()
        if r2 < 0.45:
            lit = rng.choice(_F_SPECIAL)
            if t == "f32" and lit == "0x1p-1022":
                lit = "0x1p-149"
            return "(%s.const %s)" % (t, lit)
        if r2 < 0.65:
            bits = rng.getrandbits(32 if t == "f32" else 64)
            return "(%s.reinterpret_i%s (i%s.const %d))" % (t, t[1:], t[1:], bits)
        if r2 < 0.8:
            return "(local.get $%s)" % rng.choice(env[t])
        if r2 < 0.9:
            src = rng.choice(["i32", "i64"])
            gl = {"i32": "$h", "i64": "$g"}[src]
            operand = "(global.get %s)" % gl if rng.random() < 0.4 else "(%s.const %d)" % (src, rng.choice([0, 1, -1, 7, -1234, 1 << 31, (1 << 53) + 1, rng.getrandbits(int(src[1:]))]) % (1 << int(src[1:])))
            return "(%s.convert_%s_%s %s)" % (t, src, rng.choice("su"), operand)
        return "(%s.load offset=%d (i32.const %d))" % (t, rng.choice([0, 4, 8]), rng.choice([64, 72, 80]))
    if r < 0.4:
        op = rng.choice(["abs", "neg", "ceil", "floor", "trunc", "nearest", "sqrt"])
        return "(%s.%s %s)" % (t, op, rand_float_expr(rng, depth - 1, t, env))
    if r < 0.8:
        op = rng.choice(["add", "sub", "mul", "div", "min", "max", "copysign"])
        return "(%s.%s %s %s)" % (t, op, rand_float_expr(rng, depth - 1, t, env), rand_float_expr(rng, depth - 1, t, env))
    if r < 0.88:
        other = "f64" if t == "f32" else "f32"
        return "(%s.%s %s)" % (t, "demote_f64" if t == "f32" else "promote_f32", rand_float_expr(rng, depth - 1, other, env))
    if r < 0.94:
        return "(local.tee $%s %s)" % (rng.choice(env[t]), rand_float_expr(rng, depth - 1, t, env))
    return "(select %s %s (i32.const %d))" % (rand_float_expr(rng, depth - 1, t, env), rand_float_expr(rng, depth - 1, t, env), rng.randrange(2))


def rand_float_program(rng, nstmt=10, depth=3):
    """floating-point arithmetic, conversions, loads / stores, locals, select and globals on numbers; every statement ends in
    an integer that is committed (iNN_private_const) or kept in a global / local that is committed later, so the rows the
    reference emits carry the bits of every result"""
    env = {"f32": ["a", "b"], "f64": ["c", "d"]}
    F = lambda t, d=depth: rand_float_expr(rng, rng.randrange(1, d + 1), t, env)
    body = []
    for _ in range(nstmt):
        r = rng.random()
        t = rng.choice(["f32", "f64"])
        w = t[1:]
        if r < 0.3:
            body.append("(drop (call $i%s_private_const (i%s.reinterpret_%s %s)))" % (w, w, t, F(t)))
        elif r < 0.5:
            iw = rng.choice(["32", "64"])
            body.append("(drop (call $i%s_private_const (i%s.trunc_sat_%s_%s %s)))" % (iw, iw, t, rng.choice("su"), F(t)))
        elif r < 0.62:
            body.append("(drop (call $i32_private_const (%s.%s %s %s)))" % (t, rng.choice(["eq", "ne", "lt", "gt", "le", "ge"]), F(t, 2), F(t, 2)))
        elif r < 0.7:                                            # a trapping truncation, on an operand that is in range by construction
            iw, sg = rng.choice(["32", "64"]), rng.choice("su")
            small = "(%s.div (%s.convert_i32_%s (i32.const %d)) (%s.const %s))" % (t, t, sg, rng.getrandbits(30), t, rng.choice(["1.0", "3.0", "0.7", "1e3"]))
            body.append("(drop (call $i%s_private_const (i%s.trunc_%s_%s %s)))" % (iw, iw, t, sg, small))
        elif r < 0.8:
            body.append("(local.set $%s %s)" % (rng.choice(env[t]), F(t)))
        elif r < 0.9:
            body.append("(%s.store offset=%d (i32.const %d) %s)" % (t, rng.choice([0, 4, 8]), rng.choice([64, 72, 80]), F(t)))
        elif r < 0.95:
            body.append("(global.set $g (i64.trunc_sat_%s_%s %s))" % (t, rng.choice("su"), F(t)))
        else:
            body.append("(global.set $h (i32.add (global.get $h) (i32.trunc_sat_%s_%s %s)))" % (t, rng.choice("su"), F(t)))
    body.append("(drop (call $i64_private_const (i64.add (global.get $g) (global.get $k))))")
    body.append("(drop (call $i32_private_const (global.get $h)))")
    body.append("(drop (call $i64_private_const (i64.load (i32.const 64))))")
    body.append("(drop (call $i64_private_const (i64.load (i32.const 80))))")
    for t in ("f32", "f64"):
        for name in env[t]:
            body.append("(drop (call $i%s_private_const (i%s.reinterpret_%s (local.get $%s))))" % (t[1:], t[1:], t, name))
    head = WAT_HEAD_BOTH[:WAT_HEAD_BOTH.index("(func $t")]
    return (head + "(global $g (mut i64) (i64.const 5))\n(global $h (mut i32) (i32.const -7))\n(global $k i64 (i64.const 0x123456789))\n(memory 1)\n"
            "(func $t (local $a f32) (local $b f32) (local $c f64) (local $d f64)\n" + "\n".join(body) + "\n" + WAT_TAIL)


def rand_mixed_program(rng, nmem=10, nfloat=8):
    """statements of rand_memory_program and rand_float_program shuffled into ONE function (each list keeps its own order): witness
    stores / loads / bulk memory operations between floating-point arithmetic, conversions, globals and float memory traffic.
    The float statements use their own memory window (1064 ...), so no float load meets a stored witness"""
    import re
    mem, flt = rand_memory_program(rng, nstmt=nmem), rand_float_program(rng, nstmt=nfloat)
    body = lambda text: text[text.index("\n", text.index("(func $t")) + 1:text.rindex(WAT_TAIL)].split("\n")
    fbody = [re.sub(r"\(i32\.const (64|72|80)\)\)", lambda m: "(i32.const %d))" % (int(m.group(1)) + 1000), line) if ("load" in line or "store" in line) else line for line in body(flt)]
    fbody = [re.sub(r"\(i32\.const (64|72|80)\) ", lambda m: "(i32.const %d) " % (int(m.group(1)) + 1000), line) if ".store" in line else line for line in fbody]
    mbody = body(mem)
    merged = []
    while mbody or fbody:
        src = mbody if (mbody and (not fbody or rng.random() < len(mbody) / (len(mbody) + len(fbody)))) else fbody
        merged.append(src.pop(0))
    head = WAT_HEAD_BOTH[:WAT_HEAD_BOTH.index("(func $t")]
    seg = mem[mem.index("(data $seg"):mem.index("\n", mem.index("(data $seg")) + 1]
    return (head + "(global $g (mut i64) (i64.const 5))\n(global $h (mut i32) (i32.const -7))\n(global $k i64 (i64.const 0x123456789))\n(memory 1)\n" + seg +
            "(func $t (local $a f32) (local $b f32) (local $c f64) (local $d f64)\n" + "\n".join(merged) + "\n" + WAT_TAIL)

#!/usr/bin/env python3
"""Mechanical WGSL -> C++ transliteration of the reference's compute shaders.

TEST INFRASTRUCTURE (oracle side).  Reads the reference's own shader sources where they lie
(`/root/reference/shader/{bigint,bn254fr,kernels}.wgsl.in`, `shader/sha256.wgsl`) and writes C++
headers into `oracle/_ref/` (git-ignored: reference-derived text never enters the history).  The
generated code is compiled together with `oracle/wgsl_shim.hpp` (vector / array / barrier shim) and
`oracle/wgslref.cpp` (the dispatch order of `src/webgpu/engine.cpp`) into
`oracle/_ref/libwgslref.so`, which the CPU tests use to pin `oracle/oracle.c` against the
reference's own arithmetic, pass structure and byte order.

The transliteration is purely syntactic -- no expression is evaluated, reordered or simplified:
  * `fn f(a : T, b : U) -> R {`      ->  `R f(T a, U b) {`
  * `let x : T = e;` / `let x = e;`  ->  `const T x = e;` / `const auto x = e;`
  * `var x : T;` / `var x : T = e;`  ->  `T x{};` (WGSL zero-initialises) / `T x = e;`
  * `struct S { a : T, b : U }`      ->  `struct S { T a{}; U b{}; };`  (C++20 paren-init of aggregates
                                          stands in for WGSL value constructors)
  * `alias A = B;`                   ->  `using A = B;`
  * module `const` / `var<...>`      ->  `static const` / `static` objects; storage and uniform bindings
                                          become module-scope objects the driver fills
  * attributes (`@group @binding @align @compute @workgroup_size`) are dropped; `@builtin(x)` parameters
    are recorded so that every `@compute` entry point gets a uniform `name__entry(const builtins&)` thunk
  * type syntax (`array<T,N>`, `array<T>`, `vec4<bool>`, `vec4u`, `ptr<function,T>`) is left as is: the shim
    defines C++ templates of the same names
  * `#INSTANCES` (string-replaced by the host at run time, `src/webgpu/engine.cpp:1514-1527`) becomes the
    macro `WGSL_SHA_INSTANCES`, so the header can be instantiated for several instance counts

Usage: wgsl2cpp.py <reference_root> <out_dir>
"""
import re
import sys
from pathlib import Path

TOKEN = re.compile(r"""
    (?P<ws>\s+)
  | (?P<num>0[xX][0-9a-fA-F]+[uif]?|\d+\.\d*(?:[eE][+-]?\d+)?[fh]?|\d+[uif]?)
  | (?P<id>[A-Za-z_][A-Za-z0-9_]*)
  | (?P<op><<=|>>=|<<|>>|<=|>=|==|!=|&&|\|\||\+\+|--|\+=|-=|\*=|/=|%=|&=|\|=|\^=|->)
  | (?P<attr>@[A-Za-z_][A-Za-z0-9_]*)
  | (?P<hash>\#[A-Za-z_][A-Za-z0-9_]*)
  | (?P<p>.)
""", re.X | re.S)


def strip_comments(src):
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return re.sub(r"//[^\n]*", "", src)


def tokenize(src):
    toks = []
    for m in TOKEN.finditer(src):
        kind = m.lastgroup
        if kind == "ws":
            continue
        toks.append((kind, m.group()))
    return toks


class Tr:
    """Token-stream transliterator.  `self.i` is the cursor into `self.t`."""

    def __init__(self, toks):
        self.t = toks
        self.i = 0
        self.out = []
        self.entries = []      # (name, [(param, builtin)], uses_barrier)

    # -- cursor helpers ------------------------------------------------------------------------
    def peek(self, k=0):
        j = self.i + k
        return self.t[j][1] if j < len(self.t) else None

    def next(self):
        v = self.t[self.i][1]
        self.i += 1
        return v

    def expect(self, v):
        got = self.next()
        if got != v:
            ctx = " ".join(x[1] for x in self.t[max(0, self.i - 12):self.i + 4])
            raise SyntaxError(f"expected {v!r}, got {got!r} near: {ctx}")

    def emit(self, s):
        self.out.append(s)

    # -- types ---------------------------------------------------------------------------------
    def parse_type(self):
        """identifier optionally followed by a balanced <...> list; returned verbatim (the shim provides
        templates with the WGSL names)."""
        s = self.next()
        if self.peek() == "<":
            depth = 0
            while True:
                v = self.next()
                if v == ">>":          # two closers lexed as a shift
                    depth -= 2
                    s += "> >"
                else:
                    if v == "<":
                        depth += 1
                    elif v == ">":
                        depth -= 1
                    s += v if v != "," else ", "
                if depth == 0:
                    break
        return s

    def skip_attr(self):
        """drop `@name` or `@name(args)`; return (name, args-as-text)."""
        name = self.next()[1:]
        args = ""
        if self.peek() == "(":
            depth = 0
            while True:
                v = self.next()
                if v == "(":
                    depth += 1
                elif v == ")":
                    depth -= 1
                    if depth == 0:
                        break
                else:
                    args += v
        return name, args

    # -- module scope --------------------------------------------------------------------------
    def module(self):
        pending_compute = False
        while self.i < len(self.t):
            v = self.peek()
            if v.startswith("@"):
                name, _ = self.skip_attr()
                if name == "compute":
                    pending_compute = True
                continue
            if v == ";":
                self.next()
                continue
            if v == "struct":
                self.struct()
            elif v == "alias":
                self.next()
                a = self.next()
                self.expect("=")
                b = self.parse_type()
                self.expect(";")
                self.emit(f"using {a} = {b};\n")
            elif v == "const":
                self.next()
                name = self.next()
                ty = "auto"
                if self.peek() == ":":
                    self.next()
                    ty = self.parse_type()
                self.expect("=")
                self.emit(f"static const {ty} {name} = ")
                self.copy_until_semicolon()
            elif v == "var":
                self.module_var()
            elif v == "fn":
                self.function(pending_compute)
                pending_compute = False
            else:
                raise SyntaxError(f"unexpected module-scope token {v!r}")

    def struct(self):
        self.expect("struct")
        name = self.next()
        self.expect("{")
        self.emit(f"struct {name} {{\n")
        while self.peek() != "}":
            if self.peek().startswith("@"):
                self.skip_attr()
                continue
            member = self.next()
            self.expect(":")
            ty = self.parse_type()
            if self.peek() == ",":
                self.next()
            self.emit(f"    {ty} {member}{{}};\n")
        self.expect("}")
        self.emit("};\n")

    def module_var(self):
        self.expect("var")
        space = ""
        if self.peek() == "<":
            self.next()
            while self.peek() != ">":
                space += self.next()
            self.next()
        name = self.next()
        self.expect(":")
        ty = self.parse_type()
        self.emit(f"static {ty} {name}")
        if self.peek() == "=":
            self.next()
            self.emit(" = ")
            self.copy_until_semicolon()
        else:
            self.expect(";")
            self.emit(f";   /* var<{space}> */\n")

    def copy_until_semicolon(self):
        while True:
            v = self.peek()
            if v == ";":
                self.next()
                self.emit(";\n")
                return
            self.expr_token()

    # -- functions -----------------------------------------------------------------------------
    def function(self, is_entry):
        self.expect("fn")
        name = self.next()
        self.expect("(")
        params = []
        builtins = []
        while self.peek() != ")":
            builtin = None
            while self.peek().startswith("@"):
                an, args = self.skip_attr()
                if an == "builtin":
                    builtin = args
            pname = self.next()
            self.expect(":")
            ty = self.parse_type()
            params.append(f"{ty} {pname}")
            builtins.append((pname, builtin))
            if self.peek() == ",":
                self.next()
        self.expect(")")
        ret = "void"
        if self.peek() == "->":
            self.next()
            ret = self.parse_type()
        self.emit(f"\n{ret} {name}({', '.join(params)})\n")
        start = len(self.out)
        self.block()
        body = "".join(self.out[start:])
        if is_entry:
            self.entries.append((name, builtins, "workgroupBarrier" in body))

    def block(self):
        """copy a `{ ... }` block, rewriting `let` / `var` declarations."""
        self.expect("{")
        self.emit("{\n")
        depth = 1
        while depth:
            v = self.peek()
            if v == "{":
                self.next()
                depth += 1
                self.emit("{\n")
            elif v == "}":
                self.next()
                depth -= 1
                self.emit("}\n")
            elif v in ("let", "var"):
                self.local_decl()
            elif v == ";":
                self.next()
                self.emit(";\n")
            else:
                self.expr_token()

    def local_decl(self):
        kw = self.next()
        name = self.next()
        ty = None
        if self.peek() == ":":
            self.next()
            ty = self.parse_type()
        const = "const " if kw == "let" else ""
        if self.peek() == "=":
            self.next()
            self.emit(f"{const}{ty or 'auto'} {name} = ")
        else:
            # `var x : T;` -- WGSL zero-initialises
            self.emit(f"{ty} {name}{{}}")

    def expr_token(self):
        """copy one token of an expression / statement; generic type syntax is re-assembled so that
        `array<u32, 8>(...)` keeps its brackets together, `#X` becomes a macro name."""
        v = self.peek()
        if v in ("array", "vec2", "vec3", "vec4", "ptr") and self.peek(1) == "<":
            self.emit(self.parse_type())
            return
        if v.startswith("#"):
            self.next()
            self.emit("WGSL_SHA_" + v[1:])
            return
        self.next()
        if re.fullmatch(r"\d+\.\d*(?:[eE][+-]?\d+)?", v):
            v += "f"
        self.emit(v + " ")


def transliterate(src, tag):
    tr = Tr(tokenize(strip_comments(src)))
    tr.module()
    body = "".join(tr.out)
    thunks = []
    for name, builtins, barrier in tr.entries:
        args = []
        for pname, b in builtins:
            if b is None:
                raise SyntaxError(f"entry point {name}: parameter {pname} is not a builtin")
            args.append(f"b.{b}")
        thunks.append(f"static void {name}__entry(const builtins& b) {{ {name}({', '.join(args)}); }}\n")
    table = "static const entry_point entry_points[] = {\n" + "".join(
        f'    {{"{name}", &{name}__entry, {str(barrier).lower()}}},\n' for name, _, barrier in tr.entries
    ) + "};\n"
    head = (f"// GENERATED by oracle/wgsl2cpp.py from the reference's {tag} -- do not edit, do not commit.\n"
            "// Include inside `namespace wgsl { namespace <module> { ... } }` after wgsl_shim.hpp.\n")
    return head + body + "\n" + "".join(thunks) + table


def main():
    ref = Path(sys.argv[1])
    out = Path(sys.argv[2])
    out.mkdir(parents=True, exist_ok=True)
    sh = ref / "shader"
    kernels = (sh / "kernels.wgsl.in").read_text()
    # the reference's build substitutes the two library files into kernels.wgsl.in (shader/CMake configure_file)
    kernels = kernels.replace("@LIGETRON_WGSL_BIGINT_CONTENT@", (sh / "bigint.wgsl.in").read_text())
    kernels = kernels.replace("@LIGETRON_WGSL_BN254FR_CONTENT@", (sh / "bn254fr.wgsl.in").read_text())
    (out / "wgsl_kernels.gen.inc").write_text(transliterate(kernels, "shader/{bigint,bn254fr,kernels}.wgsl.in"))
    (out / "wgsl_sha256.gen.inc").write_text(transliterate((sh / "sha256.wgsl").read_text(), "shader/sha256.wgsl"))
    print(f"wrote {out}/wgsl_kernels.gen.inc, {out}/wgsl_sha256.gen.inc")


if __name__ == "__main__":
    main()

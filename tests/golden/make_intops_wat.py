#!/usr/bin/env python3
"""Write tests/golden/intops.wat: the repo's own program over the integer instructions beyond mul / add / sub -- bitwise,
shifts and rotates (by private and by literal counts), comparisons, counts, division and remainder, sign extension and
width changes, in 32 and 64 bits, with results that live as bit vectors and as single witnesses fed into further
instructions.  Expected values come from the WebAssembly semantics in tests/refctx_util.py (wasm_op), not from a run.

    python tests/golden/make_intops_wat.py && python tests/golden/make_refctx_vectors.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from refctx_util import wasm_op  # noqa: E402

P = lambda w, v: "(call $i%d_private_const (i%d.const %s))" % (w, w, v)
L = lambda w, v: "(i%d.const %s)" % (w, v)

# (width, instruction, first operand, second operand or None, operand forms: p = private, l = literal)
CASES = [
    (32, "and", 0xf0f0a5a5, 0x0ff0ffff, "pp"), (32, "or", 0x80000001, 0x00011000, "pl"), (32, "xor", 0xdeadbeef, 0xffffffff, "pp"),
    (32, "shl", 0x80000081, 7, "pp"), (32, "shr_s", 0x80000081, 39, "pl"), (32, "shr_u", 0x80000081, 0, "pp"),
    (32, "rotl", 0x12345678, 12, "pp"), (32, "rotr", 0x12345678, 4, "pl"),
    (32, "eq", 77, 77, "pp"), (32, "ne", 77, 77, "lp"), (32, "lt_s", 0xfffffffe, 3, "pp"), (32, "ge_u", 0xfffffffe, 3, "pp"),
    (32, "clz", 0x00012000, None, "p"), (32, "ctz", 0x00012000, None, "p"), (32, "popcnt", 0xf00f0001, None, "p"), (32, "eqz", 0, None, "p"),
    (32, "div_u", 1000003, 17, "pp"), (32, "rem_s", 0xffffff85, 10, "pp"), (32, "div_s", 0xffffff85, 0xfffffff6, "pl"),
    (32, "extend8_s", 0x12345680, None, "p"), (32, "extend16_s", 0x00017fff, None, "p"),
    (64, "xor", 0x0123456789abcdef, 0xfedcba9876543210, "pp"), (64, "shr_s", 0x8000000000000000, 63, "pp"),
    (64, "gt_s", 5, 0xffffffffffffffff, "pp"), (64, "popcnt", 0xff00000000000001, None, "p"), (64, "rem_u", 0xffffffffffffffff, 1000000007, "pp"),
]


def main():
    out = [";; tests/golden/intops.wat -- written by tests/golden/make_intops_wat.py for this repo (NOT a copy of the reference's tests):",
           ";; one assertion per integer instruction family beyond mul / add / sub, then forms that feed single-witness results",
           ";; (counts, comparisons) and re-wired bit vectors (shifts, extensions) into further instructions.",
           "(module",
           '  (import "env" "i32_private_const" (func $i32_private_const (param i32) (result i32)))',
           '  (import "env" "i64_private_const" (func $i64_private_const (param i64) (result i64)))',
           '  (import "env" "assert_equal" (func $assert_equal (param i64 i64)))',
           "", "  (func $test"]
    for w, op, a, b, forms in CASES:
        lhs = [(P if f == "p" else L)(w, hex(v)) for f, v in zip(forms, (a, b))]
        v = wasm_op(op, w, a, b)
        assert v is not None
        out.append("    (call $assert_equal (i%d.%s %s) %s)" % (w, op, " ".join(lhs), P(w if not (op in ("eq", "ne", "eqz") or op[:2] in ("lt", "gt", "le", "ge")) else 32, hex(v))))
    # compositions
    x, y = 0x00f0f000, 0x0000ff00
    v = wasm_op("add", 32, wasm_op("clz", 32, x), wasm_op("popcnt", 32, y))                      # two single-witness results summed
    out.append("    (call $assert_equal (i32.add (i32.clz %s) (i32.popcnt %s)) %s)" % (P(32, hex(x)), P(32, hex(y)), L(32, v)))
    v = wasm_op("and", 32, wasm_op("shl", 32, x, 4), wasm_op("shr_u", 32, y, 4))                 # shared zero pads meet in an AND
    out.append("    (call $assert_equal (i32.and (i32.shl %s %s) (i32.shr_u %s %s)) %s)" % (P(32, hex(x)), L(32, 4), P(32, hex(y)), L(32, 4), P(32, hex(v))))
    v = wasm_op("mul", 32, wasm_op("lt_u", 32, 3, 9), 0x55)                                      # a comparison result as a factor
    out.append("    (call $assert_equal (i32.mul (i32.lt_u %s %s) %s) %s)" % (P(32, 3), P(32, 9), P(32, hex(0x55)), L(32, v)))
    v = wasm_op("sub", 32, wasm_op("extend8_s", 32, 0x80), wasm_op("rotl", 32, 0x80000000, 1))   # cloned sign bits, moved bits
    out.append("    (call $assert_equal (i32.sub (i32.extend8_s %s) (i32.rotl %s %s)) %s)" % (P(32, hex(0x80)), P(32, hex(0x80000000)), P(32, 1), P(32, hex(v))))
    out.append("    (call $assert_equal (i64.extend_i32_s %s) %s)" % (P(32, hex(0x80000001)), P(64, hex(0xffffffff80000001))))
    out.append("    (call $assert_equal (i64.extend_i32_u %s) %s)" % (P(32, hex(0x80000001)), L(64, hex(0x80000001))))
    out.append("    (call $assert_equal (i32.wrap_i64 %s) %s)" % (P(64, hex(0x1234567880000001)), P(32, hex(0x80000001))))
    out.append("    (call $assert_equal (i64.extend32_s (i64.or %s %s)) %s)" % (P(64, hex(0x80000000)), L(64, 1), P(64, hex(0xffffffff80000001))))
    out += ["  )", "", '  (export "_start" (func $test))', ")"]
    path = os.path.join(HERE, "intops.wat")
    open(path, "w").write("\n".join(out) + "\n")
    print(path, len(out), "lines")


if __name__ == "__main__":
    main()

;; tests/golden/mul64.wat -- written for this repo (NOT a copy of the reference's tests): the same import surface and
;; folded style as the reference's tests/i64_mul.wat / i64_add.wat / i64_sub.wat, with its own cases, so that the bounded
;; front end (ligero-prover_b200/host/wat_emitter.hpp) has a program to prove where /root/reference is absent.
(module
  (import "env" "i64_private_const" (func $w (param i64) (result i64)))
  (import "env" "assert_equal" (func $same (param i64 i64)))

  (func $run
    ;; wrap-around of the 64-bit product
    (call $same
      (i64.mul (call $w (i64.const 3)) (call $w (i64.const 5)))
      (call $w (i64.const 15)))
    (call $same
      (i64.mul (call $w (i64.const -2)) (call $w (i64.const -3)))
      (call $w (i64.const 6)))
    (call $same
      (i64.mul (call $w (i64.const 0xffffffffffffffff)) (call $w (i64.const 0xffffffffffffffff)))
      (call $w (i64.const 1)))
    (call $same
      (i64.mul (call $w (i64.const 0x100000000)) (call $w (i64.const 0x100000000)))
      (call $w (i64.const 0)))
    (call $same
      (i64.mul (call $w (i64.const 0xdeadbeefcafebabe)) (call $w (i64.const 0x0123456789abcdef)))
      (call $w (i64.const 0x7eb689f4ea447d62)))
    (; sums and differences, with carry and borrow ;)
    (call $same
      (i64.add (call $w (i64.const 0xffffffffffffffff)) (call $w (i64.const 2)))
      (call $w (i64.const 1)))
    (call $same
      (i64.sub (call $w (i64.const 5)) (call $w (i64.const 7)))
      (call $w (i64.const -2)))
    (call $same
      (i64.sub (i64.add (call $w (i64.const 1_000_000)) (i64.const 17)) (call $w (i64.const 17)))
      (i64.const 1000000))
  )

  (export "_start" (func $run))
)

;; tests/golden/mul64.wat -- written for this repo (NOT a copy of the reference's tests): the same import surface and
;; folded style as the reference's tests/i64_mul.wat / i64_add.wat / i64_sub.wat, with its own cases, so that the bounded
;; front end (ligero-prover_b200/host/wat_emitter.hpp) has a program to prove where /root/reference is absent.
(module
  (import "env" "i64_private_const" (func $i64_private_const (param i64) (result i64)))
  (import "env" "assert_equal" (func $assert_equal (param i64 i64)))

  (func $test
    ;; wrap-around of the 64-bit product
    (call $assert_equal (i64.mul (call $i64_private_const (i64.const 3)) (call $i64_private_const (i64.const 5))) (call $i64_private_const (i64.const 15)))
    (call $assert_equal (i64.mul (call $i64_private_const (i64.const -2)) (call $i64_private_const (i64.const -3))) (call $i64_private_const (i64.const 6)))
    (call $assert_equal (i64.mul (call $i64_private_const (i64.const 0xffffffffffffffff)) (call $i64_private_const (i64.const 0xffffffffffffffff))) (call $i64_private_const (i64.const 1)))
    (call $assert_equal (i64.mul (call $i64_private_const (i64.const 0x100000000)) (call $i64_private_const (i64.const 0x100000000))) (call $i64_private_const (i64.const 0)))
    (call $assert_equal (i64.mul (call $i64_private_const (i64.const 0xdeadbeefcafebabe)) (call $i64_private_const (i64.const 0x0123456789abcdef))) (call $i64_private_const (i64.const 0x7eb689f4ea447d62)))
    (; sums and differences, with carry and borrow ;)
    (call $assert_equal (i64.add (call $i64_private_const (i64.const 0xffffffffffffffff)) (call $i64_private_const (i64.const 2))) (call $i64_private_const (i64.const 1)))
    (call $assert_equal (i64.sub (call $i64_private_const (i64.const 5)) (call $i64_private_const (i64.const 7))) (call $i64_private_const (i64.const -2)))
    (call $assert_equal (i64.sub (i64.add (call $i64_private_const (i64.const 1_000_000)) (i64.const 17)) (call $i64_private_const (i64.const 17))) (i64.const 1000000))
  )

  (export "_start" (func $test))
)

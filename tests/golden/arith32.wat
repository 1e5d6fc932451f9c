;; tests/golden/arith32.wat -- written for this repo (NOT a copy of the reference's tests): the i32 twin of mul64.wat, same
;; import surface and folded style as the reference's tests/i32_mul.wat / i32_add.wat / i32_sub.wat, with its own cases.
(module
  (import "env" "i32_private_const" (func $w (param i32) (result i32)))
  (import "env" "assert_equal" (func $same (param i32 i32)))

  (func $run
    ;; wrap-around of the 32-bit product
    (call $same
      (i32.mul (call $w (i32.const 7)) (call $w (i32.const 6)))
      (call $w (i32.const 42)))
    (call $same
      (i32.mul (call $w (i32.const -5)) (call $w (i32.const 3)))
      (call $w (i32.const -15)))
    (call $same
      (i32.mul (call $w (i32.const 0x10000)) (call $w (i32.const 0x10000)))
      (call $w (i32.const 0)))
    (call $same
      (i32.mul (call $w (i32.const 0xdeadbeef)) (call $w (i32.const 0x01234567)))
      (call $w (i32.const 0x760b3d29)))
    ;; sums and differences, with carry and borrow, nested forms and literal operands
    (call $same
      (i32.add (call $w (i32.const 0xffffffff)) (call $w (i32.const 9)))
      (call $w (i32.const 8)))
    (call $same
      (i32.sub (call $w (i32.const 3)) (call $w (i32.const 10)))
      (call $w (i32.const -7)))
    (call $same
      (i32.sub (i32.mul (i32.add (call $w (i32.const 1000)) (i32.const 24)) (call $w (i32.const 1024))) (i32.const 48576))
      (i32.const 1000000))
  )

  (export "_start" (func $run))
)

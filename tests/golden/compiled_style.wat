;; written the way wasm2wat prints a module compiled from C (type indices, (;n;) comments, plain instructions, a shadow stack in a
;; global, WASI start-up, an indirect call, proc_exit): the spelling a maintainer is most likely to feed the front end
(module
  (type (;0;) (func (param i32 i32) (result i32)))
  (type (;1;) (func (param i64) (result i64)))
  (type (;2;) (func (param i64 i64)))
  (type (;3;) (func))
  (type (;4;) (func (param i32)))
  (import "env" "i64_private_const" (func $i64_private_const (type 1)))
  (import "env" "assert_equal" (func $assert_equal (type 2)))
  (import "wasi_snapshot_preview1" "args_sizes_get" (func $__imported_wasi_snapshot_preview1_args_sizes_get (type 0)))
  (import "wasi_snapshot_preview1" "args_get" (func $__imported_wasi_snapshot_preview1_args_get (type 0)))
  (import "wasi_snapshot_preview1" "proc_exit" (func $__imported_wasi_snapshot_preview1_proc_exit (type 4)))
  (func $__wasm_call_ctors (type 3))
  (func $_start (type 3)
    (local i32 i32 i64)
    global.get $__stack_pointer
    i32.const 16
    i32.sub
    local.tee 0
    global.set $__stack_pointer
    call $__wasm_call_ctors
    block  ;; label = @1
      local.get 0
      i32.const 12
      i32.add
      local.get 0
      i32.const 8
      i32.add
      call $__imported_wasi_snapshot_preview1_args_sizes_get
      br_if 0 (;@1;)
      i32.const 2048
      i32.const 4096
      call $__imported_wasi_snapshot_preview1_args_get
      drop
      i32.const 2052
      i32.load
      i64.load align=1
      local.set 2
      loop  ;; label = @2
        local.get 2
        local.get 2
        i64.mul
        i64.const 65535
        i64.and
        local.set 2
        local.get 1
        i32.const 1
        i32.add
        local.tee 1
        i32.const 2
        i32.lt_u
        br_if 0 (;@2;)
      end
      local.get 2
      i64.const 7
      i32.const 1
      call_indirect (type 1)
      call $assert_equal
      local.get 0
      i32.const 16
      i32.add
      global.set $__stack_pointer
      i32.const 0
      call $__imported_wasi_snapshot_preview1_proc_exit
      unreachable
    end
    unreachable)
  (func $id (type 1) (param i64) (result i64)
    local.get 0
    call $i64_private_const
    drop
    i64.const 2401)
  (table (;0;) 2 2 funcref)
  (memory (;0;) 2)
  (global $__stack_pointer (mut i32) (i32.const 66560))
  (export "memory" (memory 0))
  (export "_start" (func $_start))
  (elem (;0;) (i32.const 1) func $id)
  (data $.rodata (i32.const 1024) "hello\00"))

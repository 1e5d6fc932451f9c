// Host micro-benchmark of the .wat front end + witness machine ("Throughput"): one run of a loop of i64.mul / add / xor
// over witnesses, timed on one core.
//   g++ -std=c++17 -O2 -o /tmp/emit_bench tools/emit_bench.cpp -lcrypto && /tmp/emit_bench 5000 [seeded]
// argv[1] = loop iterations; a second argument runs the stage-2 pass (coefficients drawn from the AES stream, only they are stored);
// the second argument `c` runs a loop on numbers only instead (address arithmetic, a store, a counter): instructions per second
#include "../ligero-prover_b200/host/wat_emitter.hpp"
#include <chrono>
#include <cstdio>
using namespace ligero::cuda::host;
int main(int argc, char **argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 5000;
    const bool seeded = argc > 2 && argv[2][0] != 'c';
    char buf[4096];
    snprintf(buf, sizeof buf, R"((module (import "env" "i64_private_const" (func $pc (param i64) (result i64))) (import "env" "assert_equal" (func $eq (param i64 i64)))
(func $t (local $i i32) (local $acc i64)
 (local.set $acc (call $pc (i64.const 3)))
 (loop $l
   (local.set $acc (i64.add (i64.mul (local.get $acc) (call $pc (i64.const 6364136223846793005))) (i64.xor (local.get $acc) (i64.const 1442695040888963407))))
   (local.set $i (i32.add (local.get $i) (i32.const 1)))
   (br_if $l (i32.lt_u (local.get $i) (i32.const %d)))))
(export "_start" (func $t))))", N);
    if (argc > 2 && argv[2][0] == 'c') {                      // numbers only: what compiled code mostly executes (21 instructions per round)
        snprintf(buf, sizeof buf, R"((module (memory 1) (func (export "_start") (local $i i32) (local $acc i64)
 (loop $l
   (local.set $acc (i64.add (i64.mul (local.get $acc) (i64.const 6364136223846793005)) (i64.xor (i64.extend_i32_u (local.get $i)) (i64.const 1442695040888963407))))
   (i64.store (i32.and (i32.shl (local.get $i) (i32.const 3)) (i32.const 0xfff8)) (local.get $acc))
   (local.set $i (i32.add (local.get $i) (i32.const 1)))
   (br_if $l (i32.lt_u (local.get $i) (i32.const %d)))))))", N);
        wat_program concrete(buf);
        auto c0 = std::chrono::steady_clock::now();
        row_packer none(64);
        witness_machine cm(none, nullptr);
        wat_stats cst;
        concrete.run(cm, cst);
        const double cdt = std::chrono::duration<double>(std::chrono::steady_clock::now() - c0).count();
        printf("%d rounds of 21 instructions on numbers  %.3f s  %.1f M instructions/s\n", N, cdt, 21.0 * N / cdt / 1e6);
        return 0;
    }
    wat_program prog(buf);
    uint8_t seed[32]; for (int i = 0; i < 32; i++) seed[i] = i;
    auto t0 = std::chrono::steady_clock::now();
    row_packer pk(8000, argc <= 2, argc > 2);
    witness_machine m(pk, seeded ? seed : nullptr);
    wat_stats st;
    prog.run(m, st);
    uint32_t cs[8];
    m.finish(cs);
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("slots %lu constraints %lu  %.3f s  %.2f M slots/s\n", (unsigned long)st.quadratic_slots, (unsigned long)st.linear_constraints, dt, st.quadratic_slots / dt / 1e6);
}

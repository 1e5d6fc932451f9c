// Mutation fuzzer for the two front ends of host/wat_emitter.hpp (text and WebAssembly binary), meant to be built with
// -fsanitize=address,undefined: every mutated module must either run to the end or be rejected with std::invalid_argument --
// no out-of-bounds access, no undefined arithmetic, no leak.  Driven by tests/test_prover_cpu.py.
//   fuzz_wat <seed module (.wat or .wasm)> <iterations>
#include "../../ligero-prover_b200/host/wat_emitter.hpp"
#include <fstream>
#include <iostream>
#include <iterator>
#include <random>
using namespace ligero::cuda::host;
static int run(const std::string &data) {
    try {
        wat_program p(data);
        p.set_step_limit(20000);
        row_packer pk(64);
        witness_machine m(pk, nullptr);
        wat_stats st;
        p.run(m, st);
        m.finish(nullptr);
        return 1;
    } catch (const std::invalid_argument &e) { return 0; }
}
int main(int argc, char **argv) {
    std::ifstream in(argv[1], std::ios::binary);
    const std::string good((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    const int iters = atoi(argv[2]);
    const bool text = good.compare(0, 4, std::string("\0asm", 4)) != 0;
    std::mt19937_64 rng(12345);
    const char *frag[] = {"(", ")", "(i32.add ", "0x", ";;", "(;", "\"", "i64.clz", "(call $", " 99999999999999999999 ", "drop", "-"};
    long ok = 0, rej = 0;
    for (int it = 0; it < iters; it++) {
        std::string b = good;
        int nmut = 1 + rng() % 3;
        for (int k = 0; k < nmut && b.size() > 9; k++) {
            size_t pos = (text ? 0 : 8) + rng() % (b.size() - (text ? 0 : 8));
            switch (rng() % 3) {
            case 0: b[pos] = text ? "()$ 0x9ai.\"\n;-"[rng() % 14] : (char)(rng() & 0xff); break;
            case 1: b.erase(pos, 1 + rng() % 5); break;
            default: if (text) b.insert(pos, frag[rng() % 12]); else { std::string x; for (int j = 0, n = 1 + rng() % 3; j < n; j++) x.push_back((char)(rng() & 0xff)); b.insert(pos, x); }
            }
        }
        (run(b) ? ok : rej)++;
    }
    std::cout << argv[1] << ": accepted " << ok << " rejected " << rej << "\n";
}

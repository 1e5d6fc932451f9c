// Throughput of the PER-ROW drop-in schedule -- what a maintainer gets on day one by binding ligero::cuda_context into the
// reference's stage contexts unchanged: one row per callback,
//   stage 1 (include/zkp/nonbatch_context.hpp:445-451):  write_buffer_clear -> encode_ntt_device -> sha256_digest_update
//   stage 2 triple (:673-730,756-780): six uploads + six encodes, three check_code, three check_linear, check_quadratic
// driven through the C++ adapter exactly as the reference drives webgpu_context.  Prints one JSON object.  The stage-1 root
// is compared with lgr_encode_commit (the batched pipeline) on the same rows.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../ligero-prover_b200/host/cuda_executor.hpp"

using namespace ligero;
using big = cuda::device_uint256_t;
using buffer_t = cuda_context::buffer_type;

static const uint32_t P[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
// src/bn254.cpp:36-43 roots for k = 8192 (tests/golden/survey_vectors.json)
static void hex_limbs(const char *hex, uint32_t out[8]) {
    for (int i = 0; i < 8; i++) { unsigned v; sscanf(hex + (7 - i) * 8, "%8x", &v); out[i] = v; }
}
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char **argv) {
    const uint32_t rows = argc > 1 ? atoi(argv[1]) : 2048, triples = argc > 2 ? atoi(argv[2]) : 256;
    const uint32_t k = 8192, n = 4 * k, l = k - 192;
    big p, wk, w2k, wn;
    memcpy(p.limbs, P, 32);
    hex_limbs("10e3d295c1599ff535a1bb49f23d81aa03bd0ed25881f9ed12b179af67f67ae1", wk.limbs);
    hex_limbs("2337acd19f40bf2b2aa212849e9a0c07d626d9ca335d73a09119dbe6eaab3cac", w2k.limbs);
    hex_limbs("1f67bc4574eaef5e630a13c710221a3e3d491e59fddabaf321e56f3ca8d91624", wn.limbs);
    cuda_context exe;
    exe.webgpu_init(k, "");
    exe.ntt_init(l, k, n, p, p, wk, w2k, wn);

    // host rows: 64-bit witnesses like real programs produce, in the 2k-element scratch the stage contexts export into
    const uint32_t distinct = 64;
    std::vector<std::vector<uint64_t>> host(distinct, std::vector<uint64_t>(2 * (size_t)k * 4, 0));
    uint64_t s = 0x9e3779b97f4a7c15ull;
    for (auto &row : host) for (uint32_t i = 0; i < k; i++) { s = s * 6364136223846793005ull + 1442695040888963407ull; row[(size_t)i * 4] = s; }

    exe.sha256_init(exe.encoding_size());
    buffer_t dx = exe.make_codeword_buffer(), dy = exe.make_codeword_buffer(), dz = exe.make_codeword_buffer();
    buffer_t rx = exe.make_codeword_buffer(), ry = exe.make_codeword_buffer(), rz = exe.make_codeword_buffer();
    buffer_t sha_ctx = exe.make_device_buffer(exe.encoding_size() * sizeof(cuda_context::sha256_context));
    buffer_t sha_dig = exe.make_device_buffer(exe.encoding_size() * 32);
    auto bind_ctx = exe.bind_sha256_context(sha_ctx, sha_dig);
    auto bnx = exe.bind_ntt(dx), bny = exe.bind_ntt(dy), bnz = exe.bind_ntt(dz), bnrx = exe.bind_ntt(rx), bnry = exe.bind_ntt(ry), bnrz = exe.bind_ntt(rz);
    auto bsx = exe.bind_sha256_buffer(dx);

    auto stage1 = [&](uint32_t count) {
        exe.sha256_digest_init(bind_ctx);
        for (uint32_t r = 0; r < count; r++) {
            const std::vector<uint64_t> &limbs = host[r % distinct];
            exe.write_buffer_clear(dx, limbs.data(), limbs.size());
            exe.encode_ntt_device(bnx);
            exe.sha256_digest_update(bind_ctx, bsx);
        }
        exe.sha256_digest_final(bind_ctx);
        exe.device_synchronize();
    };
    stage1(64);                                                       // warm-up: tables, staging, clocks
    double t0 = now();
    stage1(rows);
    const double s1 = now() - t0;
    std::vector<uint8_t> dig = exe.copy_to_host<uint8_t>(sha_dig);

    // the same rows through the batched pipeline
    buffer_t all = exe.make_device_buffer((size_t)distinct * k * 32);
    for (uint32_t r = 0; r < distinct; r++) exe.write_buffer_raw(all.slice_bytes((size_t)r * k * 32, (size_t)k * 32), host[r].data(), (size_t)k * 32);
    buffer_t sha2 = exe.make_device_buffer(exe.encoding_size() * sizeof(cuda_context::sha256_context)), dig2 = exe.make_device_buffer(exe.encoding_size() * 32);
    auto bind2 = exe.bind_sha256_context(sha2, dig2);
    exe.sha256_digest_init(bind2);
    for (uint32_t r0 = 0; r0 < rows; r0 += distinct) exe.encode_absorb(sha2, all, std::min(distinct, rows - r0));
    exe.sha256_digest_final(bind2);
    const bool same = dig == exe.copy_to_host<uint8_t>(dig2);

    // stage 2, one quadratic triple per iteration
    buffer_t code = exe.make_codeword_buffer(), linear = exe.make_codeword_buffer(), quad = exe.make_codeword_buffer();
    buffer_t tmp1 = exe.make_codeword_buffer(), tmp2 = exe.make_codeword_buffer();
    auto cx = exe.bind_eltwise2(dx, code), cy = exe.bind_eltwise2(dy, code), cz = exe.bind_eltwise2(dz, code);
    auto lx = exe.bind_eltwise3(dx, rx, linear), ly = exe.bind_eltwise3(dy, ry, linear), lz = exe.bind_eltwise3(dz, rz, linear);
    auto qm = exe.bind_eltwise3(dx, dy, tmp1), qs = exe.bind_eltwise3(tmp1, dz, tmp2), qf = exe.bind_eltwise2(tmp2, quad);
    big r1(0x1234567ull), r2(0x7654321ull);
    auto stage2 = [&](uint32_t count) {
        for (uint32_t t = 0; t < count; t++) {
            const buffer_t dev[6] = {dx, rx, dy, ry, dz, rz};
            for (int j = 0; j < 6; j++) { const std::vector<uint64_t> &limbs = host[(6 * t + j) % distinct]; exe.write_buffer_clear(dev[j], limbs.data(), limbs.size()); }
            exe.encode_ntt_device(bnx); exe.encode_ntt_device(bnrx); exe.encode_ntt_device(bny); exe.encode_ntt_device(bnry);
            exe.encode_ntt_device(bnz); exe.encode_ntt_device(bnrz);
            exe.EltwiseFMAMod(cx, r1); exe.EltwiseFMAMod(cy, r1); exe.EltwiseFMAMod(cz, r1);
            exe.EltwiseFMAMod(lx); exe.EltwiseFMAMod(ly); exe.EltwiseFMAMod(lz);
            exe.EltwiseMultMod(qm); exe.EltwiseSubMod(qs); exe.EltwiseFMAMod(qf, r2);
        }
        exe.device_synchronize();
    };
    stage2(8);
    t0 = now();
    stage2(triples);
    const double s2 = now() - t0;
    uint64_t launches = 0;
    lgr_launch_count(exe.handle(), &launches);
    printf("{\"k\": %u, \"stage1_rows\": %u, \"stage1_rows_per_s\": %.1f, \"stage1_elements_per_s\": %.4e, \"stage1_us_per_row\": %.2f, "
           "\"stage1_root_equals_batched_pipeline\": %s, \"stage2_triples\": %u, \"stage2_triples_per_s\": %.1f, \"stage2_us_per_triple\": %.2f, "
           "\"h2d_bytes_per_row\": %zu}\n",
           k, rows, rows / s1, (double)rows * k / s1, s1 / rows * 1e6, same ? "true" : "false", triples, triples / s2, s2 / triples * 1e6,
           host[0].size() * 8);
    return same ? 0 : 1;
}

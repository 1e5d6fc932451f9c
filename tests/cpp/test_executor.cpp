// Drives ligero::cuda_context (ligero-prover_b200/host/cuda_executor.hpp) the way
// nonbatch_stage1_context / nonbatch_stage2_context drive webgpu_context
// (include/zkp/nonbatch_context.hpp:392-584, 586-870): one upload + encode + hash update per row,
// flush_digests, then a stage-2 pass accumulating `code` and `quad`.  Input rows and scalars come
// from a file written by the Python test; outputs go back to a file and are compared with the oracle.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../ligero-prover_b200/host/cuda_executor.hpp"

using namespace ligero;
using big = cuda::device_uint256_t;

static std::vector<uint32_t> read_all(const char *path) {
    FILE *f = fopen(path, "rb"); if (!f) { perror(path); exit(2); }
    fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<uint32_t> v(sz / 4); if (fread(v.data(), 1, sz, f) != (size_t)sz) exit(2); fclose(f); return v;
}

int main(int argc, char **argv) {
    if (argc != 4) { fprintf(stderr, "usage: %s <k> <in.bin> <out.bin>\n", argv[0]); return 2; }
    const uint32_t k = atoi(argv[1]), n = 4 * k, l = k > 192 ? k - 192 : 1;
    // file layout: p, root_k, root_2k, root_n (4 elements), nrows (1 word + 7 pad), r[nrows] scalars, rows[nrows][k]
    std::vector<uint32_t> in = read_all(argv[2]);
    big p, wk, w2k, wn;
    memcpy(p.limbs, &in[0], 32); memcpy(wk.limbs, &in[8], 32); memcpy(w2k.limbs, &in[16], 32); memcpy(wn.limbs, &in[24], 32);
    const uint32_t nrows = in[32];
    const uint32_t *scal = &in[40], *rows = &in[40 + 8 * nrows];

    cuda_context exe;
    exe.webgpu_init(k, "");
    exe.ntt_init(l, k, n, p, p, wk, w2k, wn);
    using buffer_t = cuda_context::buffer_type;

    // ---- stage 1 (nonbatch_context.hpp:412-451,555-558) ----
    exe.sha256_init(exe.encoding_size());
    std::vector<uint64_t> limbs(2 * exe.padding_size() * 4);
    buffer_t device_x = exe.make_codeword_buffer();
    buffer_t sha_ctx = exe.make_device_buffer(exe.encoding_size() * sizeof(cuda_context::sha256_context));
    buffer_t sha_dig = exe.make_device_buffer(exe.encoding_size() * 32);
    auto bind_ntt_x = exe.bind_ntt(device_x);
    auto bind_ctx = exe.bind_sha256_context(sha_ctx, sha_dig);
    auto bind_sha_x = exe.bind_sha256_buffer(device_x);
    exe.sha256_digest_init(bind_ctx);
    for (uint32_t r = 0; r < nrows; r++) {
        std::fill(limbs.begin(), limbs.end(), 0);
        memcpy(limbs.data(), rows + (size_t)r * k * 8, (size_t)k * 32);           // export_limbs
        exe.write_buffer_clear(device_x, limbs.data(), limbs.size());
        exe.encode_ntt_device(bind_ntt_x);
        exe.sha256_digest_update(bind_ctx, bind_sha_x);
    }
    exe.sha256_digest_final(bind_ctx);
    std::vector<uint8_t> digests = exe.copy_to_host<uint8_t>(sha_dig);

    // ---- stage 2: code test + quadratic test on consecutive triples (nonbatch_context.hpp:756-780) ----
    buffer_t code = exe.make_codeword_buffer(), quad = exe.make_codeword_buffer(), tmp1 = exe.make_codeword_buffer(), tmp2 = exe.make_codeword_buffer();
    buffer_t dy = exe.make_codeword_buffer(), dz = exe.make_codeword_buffer();
    auto upload = [&](buffer_t dev, uint32_t r) {
        std::fill(limbs.begin(), limbs.end(), 0);
        memcpy(limbs.data(), rows + (size_t)r * k * 8, (size_t)k * 32);
        exe.write_buffer_clear(dev, limbs.data(), limbs.size());
        exe.encode_ntt_device(exe.bind_ntt(dev));
    };
    for (uint32_t r = 0; r < nrows; r++) {
        upload(device_x, r);
        big s; memcpy(s.limbs, scal + 8 * r, 32);
        exe.EltwiseFMAMod(exe.bind_eltwise2(device_x, code), s);                    // check_code
    }
    for (uint32_t r = 0; r + 2 < nrows; r += 3) {
        upload(device_x, r); upload(dy, r + 1); upload(dz, r + 2);
        big s; memcpy(s.limbs, scal + 8 * r, 32);
        exe.EltwiseMultMod(exe.bind_eltwise3(device_x, dy, tmp1));                  // check_quadratic
        exe.EltwiseSubMod(exe.bind_eltwise3(tmp1, dz, tmp2));
        exe.EltwiseFMAMod(exe.bind_eltwise2(tmp2, quad), s);
    }
    // x == y detection used for squaring (nonbatch_context.hpp:542)
    if (!(device_x == device_x) || (device_x == dy) || !(device_x.slice(32).offset() == 32)) { fprintf(stderr, "buffer_view semantics broken\n"); return 3; }
    std::vector<uint8_t> code_h = exe.copy_to_host<uint8_t>(code), quad_h = exe.copy_to_host<uint8_t>(quad);

    FILE *f = fopen(argv[3], "wb"); if (!f) { perror(argv[3]); return 2; }
    fwrite(digests.data(), 1, digests.size(), f); fwrite(code_h.data(), 1, code_h.size(), f); fwrite(quad_h.data(), 1, quad_h.size(), f);
    fclose(f);
    printf("ok rows=%u n=%u\n", nrows, n);
    return 0;
}

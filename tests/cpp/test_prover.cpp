// Pure C++ use of the prover layer (ligero-prover_b200/host/: row_packer.hpp + matrix_prover.hpp over include/lgr.h), the way
// src/webgpu_prover.cpp would use it once the backend hands over released witnesses: witnesses -> rows -> three stages ->
// proof_data.gz.  Witness values come from a file written by the Python test (tests/test_prover_gpu.py), which compares the
// printed root / seeds and the written proof with the CPU restatement.
//   usage: test_prover <k> <l> <roots.bin> <witnesses.bin> <proof_out.gz>     roots.bin: w_k, w_2k, w_4k (src/bn254.cpp:51-64), 8 limbs each
//   witnesses.bin: records of 49 u32: kind (0 linear, 1 quadratic) then 6 x 8 limbs (x, y, z, cx, cy, cz; linear uses x, cx)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../ligero-prover_b200/host/matrix_prover.hpp"
#include "../../ligero-prover_b200/host/row_packer.hpp"

using namespace ligero::cuda::host;

static const uint32_t P[8] = {0xF0000001u, 0x43E1F593u, 0x79B97091u, 0x2833E848u, 0x8181585Du, 0xB85045B6u, 0xE131A029u, 0x30644E72u};
static void hex(const uint8_t *p, size_t n) { for (size_t i = 0; i < n; i++) printf("%02x", p[i]); }

int main(int argc, char **argv) {
    if (argc != 6) { fprintf(stderr, "usage: %s <k> <l> <roots.bin> <witnesses.bin> <proof_out.gz>\n", argv[0]); return 2; }
    const uint32_t k = atoi(argv[1]), l = atoi(argv[2]);
    uint32_t roots[24];
    { FILE *f = fopen(argv[3], "rb"); if (!f || fread(roots, 4, 24, f) != 24) { perror(argv[3]); return 2; } fclose(f); }
    std::vector<uint32_t> w;
    { FILE *f = fopen(argv[4], "rb"); if (!f) { perror(argv[4]); return 2; }
      fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET); w.resize(sz / 4);
      if (fread(w.data(), 1, sz, f) != (size_t)sz) return 2; fclose(f); }
    lgr_ctx *ctx = nullptr;
    if (lgr_create(&ctx, 0, l, k, 4 * k, P, roots, roots + 8, roots + 16)) { fprintf(stderr, "lgr_create: %s\n", lgr_last_error()); return 1; }
    try {
        row_packer pk(l);
        for (size_t off = 0; off + 49 <= w.size(); off += 49) {
            const uint32_t *r = &w[off + 1];
            if (w[off] == 0) pk.push_linear(r, r + 24);
            else pk.push_quadratic(r, r + 8, r + 16, r + 24, r + 32, r + 40);
        }
        pk.finalize();
        statement st;
        st.l = l; st.k = k;
        size_t row = 0;
        for (uint8_t kind : pk.kinds()) {
            row_event ev; ev.kind = kind ? EV_QUAD : EV_LINEAR;
            for (int j = 0; j < (kind ? 3 : 1); j++, row++) { ev.val[j] = pk.values().data() + row * l * 8; ev.coef[j] = pk.coefs().data() + row * l * 8; }
            st.events.push_back(ev);
        }
        for (int i = 0; i < 32; i++) st.encoding_seed[i] = (uint8_t)(i * 7 + 1);
        st.generated_at_seconds = 1;
        // const_sum stays 0: the test closes the linear relation itself (its last linear witness is 1 with coefficient -sum)
        matrix_prover mp(ctx);
        prove_result res = mp.prove(st);
        printf("events %zu rows %llu\n", st.events.size(), (unsigned long long)res.encoded_rows);
        printf("root "); hex(res.proof.merkle_root.data, 32); printf("\n");
        printf("stage1 "); hex(res.stage1_seed.data, 32); printf("\n");
        printf("stage2 "); hex(res.stage2_seed.data, 32); printf("\n");
        printf("valid %d %d %d\n", res.valid_code, res.valid_linear, res.valid_quad);
        FILE *f = fopen(argv[5], "wb"); if (!f) { perror(argv[5]); return 2; }
        fwrite(res.gzip.data(), 1, res.gzip.size(), f); fclose(f);
        printf("ok %zu bytes\n", res.gzip.size());
    } catch (const std::exception &e) { fprintf(stderr, "error: %s\n", e.what()); lgr_destroy(ctx); return 1; }
    lgr_destroy(ctx);
    return 0;
}

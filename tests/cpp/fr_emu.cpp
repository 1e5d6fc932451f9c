// Host build of ligero-prover_b200/csrc/fr.cuh with the PTX carry primitives emulated
// (LGR_FR_HOST_EMU): lets the CPU test-suite check the Montgomery/lazy-reduction algorithm that
// the CUDA kernels run against the oracle, bit for bit, without a GPU.
#define LGR_FR_HOST_EMU 1
#include "../../ligero-prover_b200/csrc/fr.cuh"
#include <cstddef>
using namespace lgr;
extern "C" {
// out[c] = (sum_t a[t][c] * b[t][c]) * 2^-288 mod p, in [0,2p): wide accumulation + 9-round reduction
void emu_wide_dot(uint32_t *out, const uint32_t *a, const uint32_t *b, size_t T, size_t ncols) {
    for (size_t c = 0; c < ncols; c++) {
        fr_wide w; wide_zero(w);
        for (size_t t = 0; t < T; t++) {
            fr_t x, y; for (int j = 0; j < 8; j++) { x.v[j] = a[8*(t*ncols+c)+j]; y.v[j] = b[8*(t*ncols+c)+j]; }
            wide_mad(w, x, y);
        }
        fr_t r = wide_reduce9(w);
        for (int j = 0; j < 8; j++) out[8*c+j] = r.v[j];
    }
}
// the same dot product through the Karatsuba accumulators (kara_mad: three 128x128 products per element)
void emu_kara_dot(uint32_t *out, const uint32_t *a, const uint32_t *b, size_t T, size_t ncols) {
    for (size_t c = 0; c < ncols; c++) {
        fr_kara_wide wl, wh, wm; kara_zero(wl); kara_zero(wh); kara_zero(wm);
        for (size_t t = 0; t < T; t++) {
            fr_t x, y; for (int j = 0; j < 8; j++) { x.v[j] = a[8*(t*ncols+c)+j]; y.v[j] = b[8*(t*ncols+c)+j]; }
            fr_half x0, x1, xs, y0, y1, ys;
            kara_split(x, x0, x1, xs); kara_split(y, y0, y1, ys);
            kara_mad(wl, x0, y0); kara_mad(wh, x1, y1); kara_mad(wm, xs, ys);
        }
        fr_t r = kara_reduce9(wl, wh, wm);
        for (int j = 0; j < 8; j++) out[8*c+j] = r.v[j];
    }
}
// out[i] = fr_mont_mul(a[i], b[i]) raw (in [0,2p))
void emu_mont_mul(uint32_t *out, const uint32_t *a, const uint32_t *b, size_t n, int canon) {
    for (size_t i = 0; i < n; i++) {
        fr_t x, y; for (int j = 0; j < 8; j++) { x.v[j] = a[8*i+j]; y.v[j] = b[8*i+j]; }
        fr_t r = canon ? fr_mont_mul_canon(x, y) : fr_mont_mul(x, y);
        for (int j = 0; j < 8; j++) out[8*i+j] = r.v[j];
    }
}
// out[i] = fr_shoup_mul(a[i], w[i], wq[i]) (in [0,2p)); wq = floor(w * 2^256 / p) supplied by the caller
void emu_shoup_mul(uint32_t *out, const uint32_t *a, const uint32_t *w, const uint32_t *wq, size_t n) {
    for (size_t i = 0; i < n; i++) {
        fr_t x, y, z; for (int j = 0; j < 8; j++) { x.v[j] = a[8*i+j]; y.v[j] = w[8*i+j]; z.v[j] = wq[8*i+j]; }
        fr_t r = fr_shoup_mul(x, y, z);
        for (int j = 0; j < 8; j++) out[8*i+j] = r.v[j];
    }
}
void emu_binop(uint32_t *out, const uint32_t *a, const uint32_t *b, size_t n, int op) {
    for (size_t i = 0; i < n; i++) {
        fr_t x, y, r; for (int j = 0; j < 8; j++) { x.v[j] = a[8*i+j]; y.v[j] = b[8*i+j]; }
        switch (op) {
            case 0: r = fr_add(x, y); break;
            case 1: r = fr_sub(x, y); break;
            case 2: r = fr_add_lazy(x, y); break;
            case 3: r = fr_sub_lazy4(x, y); break;
            case 4: r = fr_sub_lazy(x, y); break;
            case 5: r = fr_canon4(x); break;
            default: r = fr_add_raw(x, y); break;
        }
        for (int j = 0; j < 8; j++) out[8*i+j] = r.v[j];
    }
}
}

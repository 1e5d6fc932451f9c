// Size dispatch for the per-size translation units (encode_kernels.cu, ntt_kernels.cu are each
// compiled once per log2 size, see Makefile).
#include <cstdlib>
#include "kernels.h"

namespace lgr {

#define LGR_DECL_ENC(L) cudaError_t launch_encode_rows_##L(const fr_mem *, long long, const CodewordSink &, int, const EncodeTables &, cudaStream_t);
LGR_DECL_ENC(3) LGR_DECL_ENC(4) LGR_DECL_ENC(5) LGR_DECL_ENC(6) LGR_DECL_ENC(7) LGR_DECL_ENC(8) LGR_DECL_ENC(9) LGR_DECL_ENC(10) LGR_DECL_ENC(11)
#define LGR_DECL_NTT(L) cudaError_t launch_ntt_tile_##L(const NttTileParams &, cudaStream_t);
LGR_DECL_NTT(1) LGR_DECL_NTT(2) LGR_DECL_NTT(3) LGR_DECL_NTT(4) LGR_DECL_NTT(5) LGR_DECL_NTT(6) LGR_DECL_NTT(7) LGR_DECL_NTT(8) LGR_DECL_NTT(9) LGR_DECL_NTT(10) LGR_DECL_NTT(11)

int encode_rows_max_logk() { return 11; }
int encode_rows_min_logk() { return 3; }
int ntt_tile_max_logm() { return 11; }

cudaError_t launch_encode_rows(const fr_mem *rows_in, long long in_row_stride, const CodewordSink &sink, int R,
                               int logk, const EncodeTables &t, cudaStream_t st) {
    if (R <= 0) return cudaSuccess;
    switch (logk) {
#define LGR_CASE_ENC(L) case L: return launch_encode_rows_##L(rows_in, in_row_stride, sink, R, t, st);
        LGR_CASE_ENC(3) LGR_CASE_ENC(4) LGR_CASE_ENC(5) LGR_CASE_ENC(6) LGR_CASE_ENC(7) LGR_CASE_ENC(8) LGR_CASE_ENC(9) LGR_CASE_ENC(10) LGR_CASE_ENC(11)
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_ntt_lat(const NttTileParams &p, cudaStream_t st);

// jobs of at most this many points go to the latency-oriented kernel (lat_ntt_kernel.cu); LGR_NTT_LAT_MAX=0 disables it
static long long lat_max_points() {
    static const long long v = getenv("LGR_NTT_LAT_MAX") ? atoll(getenv("LGR_NTT_LAT_MAX")) : (1ll << 16);
    return v;
}

cudaError_t launch_ntt_tile(const NttTileParams &p, cudaStream_t st) {
    if (((long long)p.total_lanes << p.logm) <= lat_max_points()) return launch_ntt_lat(p, st);
    switch (p.logm) {
#define LGR_CASE_NTT(L) case L: return launch_ntt_tile_##L(p, st);
        LGR_CASE_NTT(1) LGR_CASE_NTT(2) LGR_CASE_NTT(3) LGR_CASE_NTT(4) LGR_CASE_NTT(5) LGR_CASE_NTT(6) LGR_CASE_NTT(7) LGR_CASE_NTT(8) LGR_CASE_NTT(9) LGR_CASE_NTT(10) LGR_CASE_NTT(11)
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace lgr

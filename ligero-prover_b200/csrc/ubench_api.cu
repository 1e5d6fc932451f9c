// C ABI of the micro-benchmarks (include/lgr_ubench.h -> liblgr_ubench.so).  Kept OUT of liblgr.so: measurement code, not product.
#include <cuda_runtime.h>
#include <string>

#include "../../include/lgr_ubench.h"
#include "kernels.h"

using namespace lgr;

static thread_local std::string g_uerr;
static int ufail(const std::string &m) { g_uerr = m; return 1; }
#define CU(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) return ufail(std::string(#expr) + ": " + cudaGetErrorString(e__)); } while (0)
#define REQUIRE(cond, msg) do { if (!(cond)) return ufail(msg); } while (0)

extern "C" {

const char *lgru_last_error(void) { return g_uerr.c_str(); }

int lgru_ubench(int device, int which, double *ops) {
    REQUIRE(ops, "null argument"); CU(cudaSetDevice(device)); cudaStream_t st = 0;
    REQUIRE(which >= 0 && which <= 8, "unknown micro-benchmark");
    uint32_t *d; CU(cudaMalloc((void **)&d, 148 * 8 * 256 * 4));
    cudaEvent_t e0, e1; CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    const int blocks = 148 * 8, threads = 256;
    const bool mulbench = (which == 1 || which == 5 || which >= 6);   // Montgomery / Shoup / FP64-pipe multiplications, 4 per iteration
    const int iters = mulbench ? 512 : ((which == 0 || which >= 3) ? 4096 : 256);
    auto run = [&](void) { return which >= 6 ? launch_ubench_dpf(which, d, iters, blocks, threads, st) : launch_ubench(which, d, iters, blocks, threads, st); };
    CU(run());           // warm-up
    CU(cudaEventRecord(e0, st));
    for (int i = 0; i < 5; i++) CU(run());
    CU(cudaEventRecord(e1, st));
    CU(cudaEventSynchronize(e1));
    float ms = 0; CU(cudaEventElapsedTime(&ms, e0, e1));
    const double per_thread = mulbench ? 4.0 * iters : ((which == 0 || which >= 3) ? 8.0 * iters : (double)iters);
    *ops = 5.0 * per_thread * blocks * threads / (ms * 1e-3);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    return 0;
}

// Montgomery multiplications per second with `warps_per_sm` resident warps and `nchain` independent
// multiplications per thread (occupancy / ILP sweep)
int lgru_mont_occ(int device, int nchain, int warps_per_sm, double *ops) {
    REQUIRE(ops, "null argument"); CU(cudaSetDevice(device)); cudaStream_t st = 0;
    REQUIRE((nchain == 1 || nchain == 2 || nchain == 4) && warps_per_sm >= 1 && warps_per_sm <= 32, "bad arguments");
    uint32_t *d; CU(cudaMalloc((void **)&d, 148 * 1024 * 4));
    cudaEvent_t e0, e1; CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    const int iters = 2048 / nchain;
    CU(launch_ubench_mont_occ(nchain, warps_per_sm, d, iters, st));
    CU(cudaEventRecord(e0, st));
    CU(launch_ubench_mont_occ(nchain, warps_per_sm, d, iters, st));
    CU(cudaEventRecord(e1, st));
    CU(cudaEventSynchronize(e1));
    float ms = 0; CU(cudaEventElapsedTime(&ms, e0, e1));
    *ops = (double)iters * nchain * 148.0 * warps_per_sm * 32 / (ms * 1e-3);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    return 0;
}

// cycles per SHA-256 compression of one warp owning a scheduler (variant 3/4/5, see ubench.cu)
int lgru_chain(int device, int variant, int warps_per_cta, int active_lanes, double *cycles) {
    REQUIRE(cycles, "null argument"); CU(cudaSetDevice(device)); cudaStream_t st = 0;
    REQUIRE(variant >= 3 && variant <= 25 && warps_per_cta >= 1 && warps_per_cta <= 4 && active_lanes >= 1 && active_lanes <= 32, "bad arguments");
    uint32_t *d; CU(cudaMalloc((void **)&d, 148 * 8 * 256 * 4));
    CU(launch_ubench_chain(variant, d, 64, warps_per_cta, active_lanes, st));
    CU(launch_ubench_chain(variant, d, 512, warps_per_cta, active_lanes, st));
    uint32_t cyc = 0;
    CU(cudaMemcpyAsync(&cyc, d + 148 * 8 * 256 - 1, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    *cycles = cyc;
    cudaFree(d);
    return 0;
}


// milliseconds of: half the warps of every SM doing Montgomery multiplications alone, the other half doing SHA-256
// compressions alone, both together.  ms[2] close to max(ms[0], ms[1]) = the two overlap; close to the sum = they contend.
int lgru_overlap(int device, double ms[3]) {
    REQUIRE(ms, "null argument");
    CU(cudaSetDevice(device));
    uint32_t *d; CU(cudaMalloc((void **)&d, 148 * 8 * 256 * 4));
    cudaEvent_t e0, e1; CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    const int blocks = 148 * 8, im = 512, is = 384;
    const int cfg[3][2] = {{im, 0}, {0, is}, {im, is}};
    CU(launch_ubench_mont_sha(d, im, is, blocks, 0));
    for (int c = 0; c < 3; c++) {
        CU(cudaEventRecord(e0, 0));
        for (int r = 0; r < 3; r++) CU(launch_ubench_mont_sha(d, cfg[c][0], cfg[c][1], blocks, 0));
        CU(cudaEventRecord(e1, 0));
        CU(cudaEventSynchronize(e1));
        float t = 0; CU(cudaEventElapsedTime(&t, e0, e1));
        ms[c] = t / 3;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    return 0;
}

// a * b * 2^-260 mod p through the FP64-pipe Montgomery multiplication (dpf_mont.cuh), n elements of 8 x u32 (device-side
// conversion both ways): lets the tests check the formulation against big-integer arithmetic
int lgru_dpf_mul(int device, const uint32_t *host_a, const uint32_t *host_b, uint32_t *host_out, uint32_t n) {
    REQUIRE(host_a && host_b && host_out && n, "null argument");
    CU(cudaSetDevice(device));
    uint32_t *d;
    CU(cudaMalloc((void **)&d, (size_t)n * 96));
    CU(cudaMemcpy(d, host_a, (size_t)n * 32, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d + (size_t)n * 8, host_b, (size_t)n * 32, cudaMemcpyHostToDevice));
    CU(launch_dpf_mul(d, d + (size_t)n * 8, d + (size_t)n * 16, (int)n, 0));
    CU(cudaMemcpy(host_out, d + (size_t)n * 16, (size_t)n * 32, cudaMemcpyDeviceToHost));
    cudaFree(d);
    return 0;
}

}  // extern "C"

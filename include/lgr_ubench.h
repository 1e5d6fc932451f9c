/*
 * lgr_ubench.h -- micro-benchmarks behind the integer rooflines DESIGN.md quotes (liblgr_ubench.so).  Measurement code: it is
 * NOT part of the drop-in library (liblgr.so) and nothing on the product path calls it; bench.py uses it for `int_roofline`.
 * All functions return 0 on success; lgru_last_error() gives the text otherwise.
 */
#ifndef LGR_UBENCH_H
#define LGR_UBENCH_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
const char *lgru_last_error(void);
/* chip-wide operations per second of one primitive, 148 x 8 CTAs of 256 threads:
 *   0 IMAD.WIDE   1 Montgomery multiplication (fr.cuh, IMAD.WIDE CIOS)   2 SHA-256 compression   3 IMAD (low word)   4 DFMA
 *   5 Shoup multiplication   6 Montgomery multiplication on the FP64 pipe (dpf_mont.cuh: 52-bit limbs, DFMA halves)
 *   7 warps alternate between 1 and 6 (do the FMA and the FP64 pipe add up?)   8 as 7 with three IMAD warps per FP64 warp */
int lgru_ubench(int device, int which, double *ops_per_sec);
/* occupancy / ILP sweep of the Montgomery multiplication */
int lgru_mont_occ(int device, int nchain, int warps_per_sm, double *ops_per_sec);
/* cycles per SHA-256 compression of a lone warp, every round formulation of csrc/ubench.cu */
int lgru_chain(int device, int variant, int warps_per_cta, int active_lanes, double *cycles);
/* ms[0]: half the warps of every SM run Montgomery multiplications, alone; ms[1]: the other half run SHA-256 compressions,
 * alone; ms[2]: both at once.  Tells whether the encoder and the column hash can hide under each other on one SM. */
int lgru_overlap(int device, double ms[3]);
/* correctness hook for 6: out[i] = a[i] * b[i] * 2^-260 mod p (canonical 8 x u32 limbs in and out) */
int lgru_dpf_mul(int device, const uint32_t *host_a, const uint32_t *host_b, uint32_t *host_out, uint32_t n);
#ifdef __cplusplus
}
#endif
#endif

// Column-wise SHA-256 leaf hashing and the Merkle tree, on the device.
//
// Replaces shader/sha256.wgsl (thread-per-instance contexts that keep every pending BYTE as a u32
// in global memory and a 64-entry message array in private memory, sha256.wgsl:23-28,65-125) and
// the host OpenSSL tree build of include/zkp/merkle_tree.hpp:343-375.
//
// Byte order facts this file relies on:
//   * sha256_update appends limb i of an element most-significant byte first
//     (sha256.wgsl:155-162)  =>  the little-endian u32 limb IS the big-endian message word.
//     An even row supplies W[0..7], the next row W[8..15]: one compression per two rows, no swaps.
//   * sha256_final stores the digest as the 8 state words in native u32 (sha256.wgsl:226-228).
//   * tree nodes are byte strings: node = SHA-256(left.bytes || right.bytes) with the standard
//     big-endian digest (OpenSSL, include/zkp/hash.hpp:181-187).  In terms of u32 loads:
//     message word = bswap32(stored word), stored parent word = bswap32(state word), uniformly for
//     every level (the leaf level's stored words are raw state words).
#include <cstdio>
#include <cstdlib>
#include "kernels.h"
#include "ntt.cuh"

namespace lgr {

// compile-time copy so the fully unrolled rounds carry K as immediates
#define LGR_K256_LIST                                                                                  \
    0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u, \
    0xd807aa98u, 0x12835b01u, 0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u, \
    0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau, \
    0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u, \
    0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u, \
    0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u, \
    0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u, \
    0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u

__device__ __forceinline__ uint32_t rotr(uint32_t x, int n) { return __funnelshift_r(x, x, n); }

// FIPS 180-4 compression; w[16] is consumed (rolling schedule), all loops unrolled.
__device__ __forceinline__ void sha256_compress(uint32_t st[8], uint32_t w[16]) {
    constexpr uint32_t K[64] = {LGR_K256_LIST};
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll
    for (int i = 0; i < 64; i++) {
        uint32_t wi;
        if (i < 16) wi = w[i];
        else {
            const uint32_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
            const uint32_t s0 = rotr(w15, 7) ^ rotr(w15, 18) ^ (w15 >> 3);
            const uint32_t s1 = rotr(w2, 17) ^ rotr(w2, 19) ^ (w2 >> 10);
            wi = w[i & 15] + s0 + w[(i + 9) & 15] + s1;
            w[i & 15] = wi;
        }
        const uint32_t t1 = h + (rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)) + ((e & f) ^ (~e & g)) + K[i] + wi;
        const uint32_t t2 = (rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
        h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}

// ---- column contexts: state[8][n] | pend[8][n] | rows_lo[n] | rows_hi[n] ----------------------
__global__ void sha_init_kernel(uint32_t *ctx, int n) {
    const uint32_t iv[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
#pragma unroll
        for (int i = 0; i < 8; i++) { ctx[(size_t)i * n + j] = iv[i]; ctx[(size_t)(8 + i) * n + j] = 0; }
        ctx[(size_t)16 * n + j] = 0; ctx[(size_t)17 * n + j] = 0;
    }
}

__device__ __forceinline__ void load_words(uint32_t *dst, const fr_mem *p) {
    fr_t x = fr_ldg(p);
#pragma unroll
    for (int i = 0; i < 8; i++) dst[i] = x.v[i];
}

// one thread per column; absorbs T rows of a row-major tile (shader/sha256.wgsl:147-177 semantics)
__global__ void __launch_bounds__(128) sha_update_kernel(uint32_t *ctx, int n, const fr_mem *__restrict__ tile, long long row_stride, int T) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint32_t st[8], w[16];
#pragma unroll
    for (int i = 0; i < 8; i++) st[i] = ctx[(size_t)i * n + j];
    const uint32_t rows_lo = ctx[(size_t)16 * n + j], rows_hi = ctx[(size_t)17 * n + j];
    const fr_mem *col = tile + j;
    int t = 0;
    if ((rows_lo & 1u) && T > 0) {
#pragma unroll
        for (int i = 0; i < 8; i++) w[i] = ctx[(size_t)(8 + i) * n + j];
        load_words(w + 8, col);
        sha256_compress(st, w);
        t = 1;
    }
    uint32_t nx[16];
    if (t + 1 < T) { load_words(nx, col + (long long)t * row_stride); load_words(nx + 8, col + (long long)(t + 1) * row_stride); }
    for (; t + 1 < T; t += 2) {
#pragma unroll
        for (int i = 0; i < 16; i++) w[i] = nx[i];
        if (t + 3 < T) { load_words(nx, col + (long long)(t + 2) * row_stride); load_words(nx + 8, col + (long long)(t + 3) * row_stride); }
        sha256_compress(st, w);
    }
    if (t < T) {
        load_words(w, col + (long long)t * row_stride);
#pragma unroll
        for (int i = 0; i < 8; i++) ctx[(size_t)(8 + i) * n + j] = w[i];
    }
#pragma unroll
    for (int i = 0; i < 8; i++) ctx[(size_t)i * n + j] = st[i];
    const uint32_t lo = rows_lo + (uint32_t)T;
    ctx[(size_t)16 * n + j] = lo;
    ctx[(size_t)17 * n + j] = rows_hi + (lo < rows_lo ? 1u : 0u);
}

// ---- narrow matrices: one dependent chain per column, few columns ------------------------------
// With n = 1024 columns there are only 32 warps of work and every column is a Merkle-Damgard chain
// over all rows, so the commit is bound by how fast ONE warp gets through a compression
// (measured on B200, ubench variants 3/4: 2660 cycles with the message schedule in line, 1938
// when K+W comes from shared memory).  sha_chain_kernel therefore splits the work inside a CTA:
//   warp 0      : the chain -- 64 rounds per block, reads K[t]+W[t] from a shared-memory ring
//   warps 1..3  : the message schedule for blocks b = h, h+3, ... (state independent), each on its
//                 own scheduler, filling the ring ahead of the chain
// Ring slots are handed over with mbarriers (full: 32 producer lanes arrive; empty: the chain
// warp's lane 0 arrives).  The CTA asks for enough shared memory that no other CTA shares the SM,
// so the chain warp never competes for issue slots.
constexpr int kChainSlots = 24;                               // 24 x 8 KiB = 192 KiB
constexpr int kChainHelpers = 3;
constexpr size_t kChainSmem = (size_t)kChainSlots * 64 * 32 * 4 + 2 * kChainSlots * 8;

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t}"
        :: "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}

// 64 rounds with pre-added K+W (kw[t*32 + lane]).
// One warp per scheduler: the ALU pipe (SHF/LOP3/IADD3) issues one warp instruction every 2 cycles and so
// does the FMA pipe (IMAD), so a round costs max(2*ALU ops, 2*FMA ops, dependent chain).  The textbook
// association  e' = d + (h + S1 + Ch + KW)  (REASSOC = false) puts SHF -> LOP3 -> IADD3 -> IADD -> IADD on
// the chain: 1938 cycles per block for a lone warp.  REASSOC sums everything that is known early
// (h + KW + d, Maj - d) ahead of time on the FMA pipe -- multiplications by an opaque 1 / -1 (`one`,
// `mone` are kernel parameters so that ptxas keeps them as IMAD) -- which leaves SHF -> LOP3 -> IADD3 on
// the chain and 12 ALU instructions per round: 1641 cycles per block (24 cycles per round is the ALU-issue
// floor).  Other splits measured with lgr_ubench_chain (tools/chain_ubench.py): every addition on the FMA
// pipe 1717 (cross-pipe hops lengthen the chain), a' through FMA adds 1677, plain C re-association 1827.
template <bool REASSOC>
__device__ __forceinline__ void sha256_rounds_kw(uint32_t st[8], const uint32_t *kw, int lane, const uint32_t one, const uint32_t mone) {
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll
    for (int i = 0; i < 64; i++) {
        const uint32_t kwi = kw[i * 32 + lane];
        const uint32_t s1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25);
        const uint32_t ch = (e & f) ^ (~e & g);
        const uint32_t s0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22);
        const uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t ne, na;
        if (REASSOC) {
            const uint32_t hkd = (h * one + kwi) * one + d;            // FMA pipe, three rounds ahead of its use
            const uint32_t pd = d * mone + mj;                         // Maj - d
            ne = hkd + s1 + ch;                                        // IADD3
            na = ne + s0 + pd;                                         // IADD3
        } else {
            const uint32_t t1 = h + s1 + ch + kwi;
            ne = d + t1;
            na = t1 + (s0 + mj);
        }
        h = g; g = f; f = e; e = ne; d = c; c = b; b = a; a = na;
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}

// virtual row v of the stream seen by this launch: the pending row (if the row count so far is
// odd) followed by the T tile rows
__device__ __forceinline__ void load_virtual_row(uint32_t *dst, int v, int p, const uint32_t *ctx, int n, int col, const fr_mem *tile, long long row_stride) {
    if (p && v == 0) {
#pragma unroll
        for (int i = 0; i < 8; i++) dst[i] = ctx[(size_t)(8 + i) * n + col];
    } else {
        load_words(dst, tile + (long long)(v - p) * row_stride + col);
    }
}

// C = chain warps per CTA.  C = 1: the CTA (1 chain + 3 schedule warps) has its SM to itself -- fastest chain, one SM per 32
// columns.  C = 4: four chains per SM, one per scheduler, each sharing its scheduler with its own three schedule warps (warp w
// runs on scheduler w mod 4): the chain slows by the schedule's ALU work (1582 -> ~2300 cycles per block) but 4096 columns
// take 32 SMs instead of 128 and leave the rest to the encoder of the next tile.  The ring is split between the chains.
template <bool REASSOC, int G, int C>
__global__ void __launch_bounds__(128 * C, 1) sha_chain_kernel(uint32_t *ctx, int n, const fr_mem *__restrict__ tile, long long row_stride, int T, const uint32_t one, const uint32_t mone) {
    constexpr int SL = kChainSlots / C;                        // ring slots per chain
    static_assert(SL % G == 0 && SL / G >= 2, "ring too shallow for this hand-over group");
    extern __shared__ __align__(16) unsigned char chain_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool is_chain = warp < C;
    const int cid = is_chain ? warp : (warp - C) % C;          // which chain of the CTA this warp serves
    const int hsel = is_chain ? 0 : (warp - C) / C;            // helper index 0..2
    uint32_t *ring = reinterpret_cast<uint32_t *>(chain_smem) + (size_t)cid * SL * 64 * 32;
    uint64_t *full = reinterpret_cast<uint64_t *>(chain_smem + (size_t)kChainSlots * 64 * 32 * 4) + cid * SL;
    uint64_t *empty = reinterpret_cast<uint64_t *>(chain_smem + (size_t)kChainSlots * 64 * 32 * 4) + kChainSlots + cid * SL;
    const int cta32 = blockIdx.x * C + cid;                   // 32-column group of this chain
    const bool live = cta32 * 32 < n;                         // n is a multiple of 32; the last CTA may hold idle chains
    const int col = cta32 * 32 + lane;
    if (threadIdx.x == 0) {
        uint64_t *f0 = reinterpret_cast<uint64_t *>(chain_smem + (size_t)kChainSlots * 64 * 32 * 4);
        for (int c = 0; c < C; c++)
            for (int s = 0; s < SL / G; s++) { mbar_init(f0 + c * SL + s, 32 * G); mbar_init(f0 + kChainSlots + c * SL + s, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (!live) return;
    const uint32_t rows_lo = ctx[(size_t)16 * n + col], rows_hi = ctx[(size_t)17 * n + col];
    const int p = (int)(rows_lo & 1u);
    const int nblk = (p + T) >> 1;
    if (is_chain) {
        uint32_t st[8];
#pragma unroll
        for (int i = 0; i < 8; i++) st[i] = ctx[(size_t)i * n + col];
        // ring slots are handed over in groups of G blocks: one barrier round trip per G compressions
        // (measured, 8192-row tiles: G = 1 3.63 ms, 2 3.48, 4 3.34, 8 3.31 per launch)
        constexpr int NG = SL / G;
        int grp = 0; uint32_t phase = 0;
        for (int b0 = 0; b0 < nblk; b0 += G) {
            mbar_wait(full + grp, phase);
            const int cnt = min(G, nblk - b0);
#pragma unroll 1
            for (int q = 0; q < cnt; q++) sha256_rounds_kw<REASSOC>(st, ring + (size_t)(grp * G + q) * 64 * 32, lane, one, mone);
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + grp);
            if (++grp == NG) { grp = 0; phase ^= 1u; }
        }
#pragma unroll
        for (int i = 0; i < 8; i++) ctx[(size_t)i * n + col] = st[i];
        if ((p + T) & 1) {                                     // an unpaired last row stays pending
            uint32_t w[8];
            load_words(w, tile + (long long)(T - 1) * row_stride + col);
#pragma unroll
            for (int i = 0; i < 8; i++) ctx[(size_t)(8 + i) * n + col] = w[i];
        }
        const uint32_t lo = rows_lo + (uint32_t)T;
        ctx[(size_t)16 * n + col] = lo;
        ctx[(size_t)17 * n + col] = rows_hi + (lo < rows_lo ? 1u : 0u);
    } else {
        constexpr uint32_t K[64] = {LGR_K256_LIST};
        uint32_t w[16], nx[16];
        if (hsel < nblk) {
            load_virtual_row(nx, 2 * hsel, p, ctx, n, col, tile, row_stride);
            load_virtual_row(nx + 8, 2 * hsel + 1, p, ctx, n, col, tile, row_stride);
        }
        for (int b = hsel; b < nblk; b += kChainHelpers) {
#pragma unroll
            for (int i = 0; i < 16; i++) w[i] = nx[i];
            const int bn = b + kChainHelpers;
            if (bn < nblk) {
                load_virtual_row(nx, 2 * bn, p, ctx, n, col, tile, row_stride);
                load_virtual_row(nx + 8, 2 * bn + 1, p, ctx, n, col, tile, row_stride);
            }
            constexpr int NG = SL / G;
            const int gi = b / G, grp = gi % NG;
            const uint32_t phase = (uint32_t)(gi / NG) & 1u;
            mbar_wait(empty + grp, phase ^ 1u);                // passes immediately the first time round
            uint32_t *dst = ring + (size_t)(grp * G + b % G) * 64 * 32 + lane;
#pragma unroll
            for (int i = 0; i < 64; i++) {
                uint32_t wi;
                if (i < 16) wi = w[i];
                else {
                    const uint32_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
                    wi = w[i & 15] + (rotr(w15, 7) ^ rotr(w15, 18) ^ (w15 >> 3)) + w[(i + 9) & 15] + (rotr(w2, 17) ^ rotr(w2, 19) ^ (w2 >> 10));
                    w[i & 15] = wi;
                }
                dst[i * 32] = wi + K[i];
            }
            mbar_arrive(full + grp);                           // every lane: releases its own stores
            if (G > 1 && b == nblk - 1)                            // a ragged last group: stand in for the missing blocks
                for (int q = nblk % G; q != 0 && q < G; q++) mbar_arrive(full + grp);
        }
    }
}

// ---- lane-split chain: 16 columns per warp, the two halves of a round on the two half-warps --------
// The compression round has two halves of identical SHAPE -- three rotations xor-ed, one bitwise
// choice, one sum:
//     e' = d + h + KW + Sigma1(e) + Ch(e,f,g)            a' = (e' - d) + Sigma0(a) + Maj(a,b,c)
// (a' = T1 + T2 with T1 = e' - d), and Maj(a,b,c) = Ch(a, b|c, b&c).  Lane L < 16 of the chain warp
// keeps (e,f,g,h) of column L and lane L+16 keeps (a,b,c,d); both run ONE instruction stream whose
// per-lane constants (rotation amounts, a mask, a sign) select the half:
//     u = q | (r & M)    v = r & (~M | q)                 (from the previous round's values: off the chain)
//     p' = [rot(p,s1) ^ rot(p,s2) ^ rot(p,s3)]  +  [(p & u) | (~p & v)]  +  [sgn*s + KW + R]
// where R is the partner lane's p' of two iterations ago, exchanged with one SHFL.BFLY per round: the
// E lanes need a_{i-3} (as d), the A lanes need e_{i+1}; running the A lanes two rounds behind the E
// lanes gives both a slack of two iterations.  8 ALU-pipe instructions per round (3 SHF, 4 LOP3, 1 IADD3) instead of
// 12, the sums known early on the FMA pipe as before; what bounds the kernel is the loop e -> shuffle -> a -> shuffle -> e,
// four rounds long with two 26-cycle hops in it (about 19 cycles per round; measured 22).
// The feed-forward at block boundaries (st += working variables) is folded into the same stream:
// every lane keeps its chaining words CV and adds them at its own boundary (E lanes at iteration 0, A
// lanes at iteration 2 of a block; multipliers mE / mA make the other half's instruction a no-op), and
// the E lanes keep a copy CA of the A-side chaining words so that d = a + CA in the first four rounds
// of a block -- which is also the new CA.  One CTA = one chain warp (16 columns) + 3 schedule warps
// (6 half-warps, each preparing every 6th block); ring slots are 4 KiB.
constexpr int kC16Slots = 48;                                 // 48 x 4 KiB = 192 KiB: the CTA has its SM to itself
constexpr int kC16Stride = 64 * 16 + 16;                       // words per ring slot: the 16-word skew puts consecutive slots on opposite
                                                              // halves of the 32 banks (two half-warps store to consecutive slots; the A lanes'
                                                              // zero words sit on the half the E lanes do not read)
constexpr size_t kC16Smem = (size_t)(kC16Slots + 1) * kC16Stride * 4 + 2 * kC16Slots * 8;

// forced instruction shapes for the lane-split rounds: left to itself nvcc re-fuses the boolean algebra into a form
// whose three LOP3 all wait for the new p, and re-associates the sums onto the dependent chain
template <int LUT> __device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d; asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(LUT)); return d;
}
__device__ __forceinline__ uint32_t mad_lo(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d; asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d;
}

template <int G, int D>
__global__ void __launch_bounds__(128, 1) sha_chain16_kernel(uint32_t *ctx, int n, const fr_mem *__restrict__ tile, long long row_stride, int T, const uint32_t one) {
    extern __shared__ __align__(16) unsigned char chain_smem[];
    constexpr int NG = kC16Slots / G;
    uint32_t *ring = reinterpret_cast<uint32_t *>(chain_smem);
    uint32_t *zeros = ring + (size_t)kC16Slots * kC16Stride;  // what the A lanes read in place of K+W
    uint64_t *full = reinterpret_cast<uint64_t *>(zeros + kC16Stride);
    uint64_t *empty = full + kC16Slots;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, half = lane >> 4;
    const int col = blockIdx.x * 16 + (lane & 15);            // n is a multiple of 16
    for (int i = threadIdx.x; i < kC16Stride; i += blockDim.x) zeros[i] = 0;
    if (threadIdx.x == 0) {
        for (int g = 0; g < NG; g++) { mbar_init(full + g, 16 * G); mbar_init(empty + g, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t rows_lo = ctx[(size_t)16 * n + col], rows_hi = ctx[(size_t)17 * n + col];
    // control flow follows the first column of the group (a uniform load: every column of a context has absorbed
    // the same number of rows, the API updates them together), so the compiler sees convergent loops
    const int pend = (int)(ctx[(size_t)16 * n + blockIdx.x * 16] & 1u);
    const int nblk = (pend + T) >> 1;
    if (warp == 0) {
        const bool isE = half == 0;
        const uint32_t s1 = isE ? 6 : 2, s2 = isE ? 11 : 13, s3 = isE ? 25 : 22;
        const uint32_t sgn = isE ? one : 0u - one, mE = isE ? one : 0u, mA = isE ? 0u : one;
        const uint32_t M = 0u - mA;                           // all-ones on the A lanes; opaque (one is a kernel parameter)
        const int base = isE ? 4 : 0;
        uint32_t p = ctx[(size_t)(base + 0) * n + col], q = ctx[(size_t)(base + 1) * n + col];
        uint32_t r = ctx[(size_t)(base + 2) * n + col], s = ctx[(size_t)(base + 3) * n + col];
        uint32_t u = 0, v = 0;                                // choice operands of the coming round: Ch(p,u,v)
        uint32_t cv0 = 0, cv1 = 0, cv2 = 0, cv3 = 0;          // own chaining words (0 until the first boundary step)
        uint32_t ca[4] = {0, 0, 0, 0};                        // E lanes: the A side's chaining words
        // delay line of received words: the E lanes start with d, c (then b, a arrive from the idle A lanes)
        uint32_t rm2 = __shfl_xor_sync(0xffffffffu, s, 16), rm1 = __shfl_xor_sync(0xffffffffu, r, 16);
        u = lop3<0xF8>(q, r, M); v = lop3<0xC4>(q, r, M);
#define LGR_C16_BOUNDARY(MX)                                                                   \
    { uint32_t t_;                                                                             \
      t_ = p; p = cv0 * (MX) + p; cv0 = t_ * (MX) + cv0;  t_ = q; q = cv1 * (MX) + q; cv1 = t_ * (MX) + cv1; \
      t_ = r; r = cv2 * (MX) + r; cv2 = t_ * (MX) + cv2;  t_ = s; s = cv3 * (MX) + s; cv3 = t_ * (MX) + cv3; \
      u = lop3<0xF8>(q, r, M); v = lop3<0xC4>(q, r, M); }
        // one iteration: E lanes round J of the current block, A lanes two rounds behind
#define LGR_C16_SHIFT(NP) { s = r; r = q; q = p; p = (NP); u = lop3<0xF8>(q, r, M); v = lop3<0xC4>(q, r, M); }
#define LGR_C16_ROUND(J, KWX, ACT_A, IDLE_SEND)                                                \
    { uint32_t d_ = rm2;                                                                       \
      if ((J) < 4) { d_ = mad_lo(ca[3 - (J)], one, rm2); ca[3 - (J)] = mad_lo(d_, mE, 0u); }   \
      const uint32_t wr_ = mad_lo(mad_lo(s, sgn, (KWX)), one, d_);                             \
      const uint32_t sg_ = lop3<0x96>(__funnelshift_r(p, p, s1), __funnelshift_r(p, p, s2), __funnelshift_r(p, p, s3)); \
      const uint32_t np_ = sg_ + lop3<0xCA>(p, u, v) + wr_;                                    \
      if (ACT_A) { const uint32_t r0_ = __shfl_xor_sync(0xffffffffu, np_, 16); LGR_C16_SHIFT(np_) rm2 = rm1; rm1 = r0_; } \
      else { const uint32_t r0_ = __shfl_xor_sync(0xffffffffu, isE ? np_ : (IDLE_SEND), 16);   \
             if (isE) LGR_C16_SHIFT(np_) rm2 = rm1; rm1 = r0_; } }
        int grp = 0; uint32_t phase = 0;
        uint32_t kq[D];                                      // K+W of the next D rounds, loaded ahead of their use
#pragma unroll
        for (int j = 0; j < D; j++) kq[j] = 0;
        for (int b0 = 0; b0 < nblk; b0 += G) {
            mbar_wait(full + grp, phase);
            const int cnt = min(G, nblk - b0);
#pragma unroll 1
            for (int qb = 0; qb < cnt; qb++) {
                const int slot = grp * G + qb;
                const uint32_t *kwp = isE ? ring + (size_t)slot * kC16Stride + lane : zeros + ((slot + 1) & 1) * 16 + (lane & 15);
                // the next block of the same group is already in the ring: its first eight K+W words are fetched during
                // the last eight rounds of this block (the A lanes keep reading zeros)
                const bool chained = qb + 1 < cnt;
                const uint32_t *kwn = (isE && chained) ? kwp + kC16Stride : zeros + (slot & 1) * 16 + (lane & 15);
                const bool a_on = (b0 + qb) > 0;              // the A lanes idle through the first two iterations of the launch
                if (qb == 0) {
#pragma unroll
                    for (int j = 0; j < D; j++) kq[j] = kwp[j * 16];
                }
                LGR_C16_BOUNDARY(mE)
                if (a_on) { LGR_C16_ROUND(0, kq[0], true, 0u) kq[0] = kwp[D * 16]; LGR_C16_ROUND(1, kq[1 % D], true, 0u) }
                else { LGR_C16_ROUND(0, kq[0], false, q) kq[0] = kwp[D * 16]; LGR_C16_ROUND(1, kq[1 % D], false, p) }
                kq[1 % D] = kwp[(1 + D) * 16];
                LGR_C16_BOUNDARY(mA)
#pragma unroll
                for (int j = 2; j < 64; j++) {
                    const uint32_t kw_ = kq[j % D];
                    kq[j % D] = (j + D < 64) ? kwp[(j + D) * 16] : kwn[(j + D - 64) * 16];
                    LGR_C16_ROUND(j, kw_, true, 0u)
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + grp);
            if (++grp == NG) { grp = 0; phase ^= 1u; }
        }
        if (nblk > 0) {                                        // the A lanes' last two rounds; the E lanes only forward
            const uint32_t *kwp = zeros + (lane & 15);
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const uint32_t wr_ = mad_lo(mad_lo(s, sgn, kwp[j * 16]), one, rm2);
                const uint32_t sg_ = lop3<0x96>(__funnelshift_r(p, p, s1), __funnelshift_r(p, p, s2), __funnelshift_r(p, p, s3));
                const uint32_t np_ = sg_ + lop3<0xCA>(p, u, v) + wr_;
                if (!isE) LGR_C16_SHIFT(np_)
                rm2 = rm1; rm1 = 0;
            }
            p += cv0; q += cv1; r += cv2; s += cv3;            // the last feed-forward
        }
#undef LGR_C16_ROUND
#undef LGR_C16_SHIFT
#undef LGR_C16_BOUNDARY
        ctx[(size_t)(base + 0) * n + col] = p; ctx[(size_t)(base + 1) * n + col] = q;
        ctx[(size_t)(base + 2) * n + col] = r; ctx[(size_t)(base + 3) * n + col] = s;
        if (isE) {
            if ((pend + T) & 1) {                              // an unpaired last row stays pending
                uint32_t w[8];
                load_words(w, tile + (long long)(T - 1) * row_stride + col);
#pragma unroll
                for (int i = 0; i < 8; i++) ctx[(size_t)(8 + i) * n + col] = w[i];
            }
            const uint32_t lo = rows_lo + (uint32_t)T;
            ctx[(size_t)16 * n + col] = lo;
            ctx[(size_t)17 * n + col] = rows_hi + (lo < rows_lo ? 1u : 0u);
        }
    } else {
        constexpr uint32_t K[64] = {LGR_K256_LIST};
        const int hv = (warp - 1) * 2 + half;                 // half-warp index
        const int kC16Helpers = (int)(blockDim.x >> 5) * 2 - 2;
        uint32_t w[16], nx[16];
        if (hv < nblk) {
            load_virtual_row(nx, 2 * hv, pend, ctx, n, col, tile, row_stride);
            load_virtual_row(nx + 8, 2 * hv + 1, pend, ctx, n, col, tile, row_stride);
        }
        for (int b = hv; b < nblk; b += kC16Helpers) {
#pragma unroll
            for (int i = 0; i < 16; i++) w[i] = nx[i];
            const int bn = b + kC16Helpers;
            if (bn < nblk) {
                load_virtual_row(nx, 2 * bn, pend, ctx, n, col, tile, row_stride);
                load_virtual_row(nx + 8, 2 * bn + 1, pend, ctx, n, col, tile, row_stride);
            }
            const int gi = b / G, grp = gi % NG;
            const uint32_t phase = (uint32_t)(gi / NG) & 1u;
            mbar_wait(empty + grp, phase ^ 1u);                // passes immediately the first time round
            uint32_t *dst = ring + (size_t)(grp * G + b % G) * kC16Stride + (lane & 15);
#pragma unroll
            for (int i = 0; i < 64; i++) {
                uint32_t wi;
                if (i < 16) wi = w[i];
                else {
                    const uint32_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
                    wi = w[i & 15] + (rotr(w15, 7) ^ rotr(w15, 18) ^ (w15 >> 3)) + w[(i + 9) & 15] + (rotr(w2, 17) ^ rotr(w2, 19) ^ (w2 >> 10));
                    w[i & 15] = wi;
                }
                dst[i * 16] = wi + K[i];
            }
            mbar_arrive(full + grp);                           // every lane: releases its own stores
            if (G > 1 && b == nblk - 1)                        // a ragged last group: stand in for the missing blocks
                for (int qq = nblk % G; qq != 0 && qq < G; qq++) mbar_arrive(full + grp);
        }
    }
}

// padding + length + digest as native state words (shader/sha256.wgsl:179-230); the context is
// left untouched so that final is idempotent.
__global__ void sha_final_kernel(const uint32_t *ctx, int n, uint32_t *digests) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint32_t st[8], w[16];
#pragma unroll
    for (int i = 0; i < 8; i++) st[i] = ctx[(size_t)i * n + j];
    const uint32_t rows_lo = ctx[(size_t)16 * n + j], rows_hi = ctx[(size_t)17 * n + j];
#pragma unroll
    for (int i = 0; i < 16; i++) w[i] = 0;
    if (rows_lo & 1u) {
#pragma unroll
        for (int i = 0; i < 8; i++) w[i] = ctx[(size_t)(8 + i) * n + j];
        w[8] = 0x80000000u;
    } else {
        w[0] = 0x80000000u;
    }
    const unsigned long long rows = ((unsigned long long)rows_hi << 32) | rows_lo;
    const unsigned long long bits = rows * 256ull;
    w[14] = (uint32_t)(bits >> 32);
    w[15] = (uint32_t)bits;
    sha256_compress(st, w);
    uint4 *o = reinterpret_cast<uint4 *>(digests + (size_t)8 * j);
    o[0] = make_uint4(st[0], st[1], st[2], st[3]);
    o[1] = make_uint4(st[4], st[5], st[6], st[7]);
}

// ---- Merkle tree ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

// parent = SHA-256(left.bytes || right.bytes): data block + constant padding block
__device__ __forceinline__ void merkle_parent(const uint32_t *left_right /*16 words*/, uint32_t *parent /*8 words*/) {
    uint32_t st[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
    uint32_t w[16];
    const uint4 *p = reinterpret_cast<const uint4 *>(left_right);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint4 v = p[i];
        w[4 * i] = bswap32(v.x); w[4 * i + 1] = bswap32(v.y); w[4 * i + 2] = bswap32(v.z); w[4 * i + 3] = bswap32(v.w);
    }
    sha256_compress(st, w);
#pragma unroll
    for (int i = 0; i < 16; i++) w[i] = 0;
    w[0] = 0x80000000u; w[15] = 512u;
    sha256_compress(st, w);
    uint4 *o = reinterpret_cast<uint4 *>(parent);
    o[0] = make_uint4(bswap32(st[0]), bswap32(st[1]), bswap32(st[2]), bswap32(st[3]));
    o[1] = make_uint4(bswap32(st[4]), bswap32(st[5]), bswap32(st[6]), bswap32(st[7]));
}

// heap layout: the level with `cnt` nodes occupies indices [cnt-1, 2cnt-1)
__global__ void merkle_level_kernel(uint32_t *nodes, int cnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cnt) return;
    const size_t parent = (size_t)(cnt - 1) + i, child = 2 * parent + 1;
    merkle_parent(nodes + 8 * child, nodes + 8 * parent);
}
// all levels with <= blockDim.x nodes in one CTA
__global__ void merkle_top_kernel(uint32_t *nodes, int cnt) {
    for (; cnt >= 1; cnt >>= 1) {
        if ((int)threadIdx.x < cnt) {
            const size_t parent = (size_t)(cnt - 1) + threadIdx.x, child = 2 * parent + 1;
            merkle_parent(nodes + 8 * child, nodes + 8 * parent);
        }
        __syncthreads();
    }
}
__global__ void merkle_place_leaves(const uint32_t *leaf, int nleaves, uint32_t *nodes, int P2) {
    const size_t total = (size_t)P2 * 8;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
        nodes[(size_t)(P2 - 1) * 8 + i] = (i < (size_t)nleaves * 8) ? leaf[i] : 0u;
}

// ---- synthetic witness (the counter-based generator of BASELINE.md section 3) -------
__device__ __forceinline__ unsigned long long splitmix(unsigned long long &s) {
    unsigned long long z = (s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31);
}
__global__ void synth_kernel(fr_mem *out, unsigned long long seed, unsigned long long row0, unsigned long long nrows, unsigned long long ncols) {
    const unsigned long long total = nrows * ncols;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < total; i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long r = row0 + i / ncols, c = i % ncols;
        unsigned long long s = seed * 0xd1342543de82ef95ull + r * 0x2545f4914f6cdd1dull + c * 0x9e3779b97f4a7c15ull + 0x632be59bd9b4e019ull;
        unsigned long long w[4];
#pragma unroll
        for (int q = 0; q < 4; q++) w[q] = splitmix(s);
        // finite_field_gmp.hpp:70-78 generate_random: 256 bits >> 2, one conditional subtract
        w[0] = (w[0] >> 2) | (w[1] << 62); w[1] = (w[1] >> 2) | (w[2] << 62); w[2] = (w[2] >> 2) | (w[3] << 62); w[3] >>= 2;
        fr_t x;
#pragma unroll
        for (int q = 0; q < 4; q++) { x.v[2 * q] = (uint32_t)w[q]; x.v[2 * q + 1] = (uint32_t)(w[q] >> 32); }
        fr_stg(out + i, fr_reduce_p(x));
    }
}

// ---- launchers --------------------------------------------------------------------------------
cudaError_t launch_sha_init(uint32_t *ctx, int n, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    sha_init_kernel<<<(n + 255) / 256, 256, 0, st>>>(ctx, n);
    return cudaGetLastError();
}
cudaError_t launch_sha_update(uint32_t *ctx, int n, const fr_mem *tile, long long row_stride, int T, cudaStream_t st) {
    if (n <= 0 || T <= 0) return cudaSuccess;
    // Lane-split chain (16 columns per warp, twice the SMs of the 32-column kernel below).  In isolation its rounds run
    // in 1277 cycles per block against 1641 (lgr_ubench_chain 19 vs 11); inside the kernel it measures 1470-1575 against
    // 1573-1586 depending on how ptxas lays the loop out, so it is used only where the extra SMs are free anyway: up to 64
    // CTAs (n <= 1024), where the encoder of the next tile still hides under the hash.  LGR_CHAIN_SPLIT=0/1 forces it.
    static const int split_env = getenv("LGR_CHAIN_SPLIT") ? atoi(getenv("LGR_CHAIN_SPLIT")) : -1;
    const bool lane_split = split_env >= 0 ? split_env != 0 : (n / 16 <= 64);
    if (lane_split && n % 16 == 0 && n / 16 <= 148 && T >= 4) {
        static const int group16 = getenv("LGR_CHAIN_GROUP") ? atoi(getenv("LGR_CHAIN_GROUP")) : 16;
        static const int helper_warps = getenv("LGR_CHAIN_HELPERS") ? atoi(getenv("LGR_CHAIN_HELPERS")) : 3;
#define LGR_C16_LAUNCH(G, D)                                                                                                            \
    {                                                                                                                                   \
        cudaError_t e = cudaFuncSetAttribute(sha_chain16_kernel<G, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kC16Smem);     \
        if (e != cudaSuccess) return e;                                                                                                 \
        sha_chain16_kernel<G, D><<<n / 16, 32 * (1 + helper_warps), kC16Smem, st>>>(ctx, n, tile, row_stride, T, 1u);                     \
        return cudaGetLastError();                                                                                                      \
    }
        // D = K+W words in flight: measured 1534 / 1511 / 1490 / 1495 / 1448 / 1448 / 1481 / 1525 / 1517 / 1578 cycles per block
        // for D = 1 / 2 / 4 / 6 / 7 / 8 / 10 / 12 / 16 / 24 (in isolation the trend is the opposite: lgr_ubench_chain 22, 24, 25)
        if (group16 == 1) LGR_C16_LAUNCH(1, 8)
        if (group16 == 4) LGR_C16_LAUNCH(4, 8)
        if (group16 == 8) LGR_C16_LAUNCH(8, 8)
        LGR_C16_LAUNCH(16, 8)                                  // hand-over group: 1 / 4 / 8 / 12 / 16 / 24 blocks -> 1511 / 1482 / 1450 / 1441 / 1431 / 1440 cycles
    }
    // widest matrix the chain kernel takes, in 32-column groups: 592 = 148 SMs x 4 chains per CTA (n <= 18944).  Measured
    // against the thread-per-column kernel (LGR_CHAIN_MAX_GROUPS=148): k = 2048 commit 1.17 -> 1.47e9 elem/s, k = 4096 commit
    // 1.30 -> 1.41e9; one rank's round of the exact layout at n/G = 8192 / 16384 columns 2.05 -> 1.53 ms / 1.69 -> 1.52 ms
    static const int max_groups = getenv("LGR_CHAIN_MAX_GROUPS") ? atoi(getenv("LGR_CHAIN_MAX_GROUPS")) : 592;
    if (n % 32 == 0 && n / 32 <= max_groups && T >= 4) {
        // narrow matrix: producer/consumer CTAs, one chain warp per SM
        static const bool textbook = getenv("LGR_CHAIN_TEXTBOOK") != nullptr;   // A/B knobs: textbook round association, blocks per hand-over
        static const int group = getenv("LGR_CHAIN_GROUP") ? atoi(getenv("LGR_CHAIN_GROUP")) : 8;
#define LGR_CHAIN_LAUNCH(RE, G, C)                                                                                                      \
    {                                                                                                                                   \
        cudaError_t e = cudaFuncSetAttribute(sha_chain_kernel<RE, G, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kChainSmem); \
        if (e != cudaSuccess) return e;                                                                                                 \
        sha_chain_kernel<RE, G, C><<<(n / 32 + (C) - 1) / (C), 128 * (C), kChainSmem, st>>>(ctx, n, tile, row_stride, T, 1u, 0xFFFFFFFFu); \
        return cudaGetLastError();                                                                                                      \
    }
        // more than LGR_CHAIN_PACK_CTAS 32-column groups (default 64, i.e. n > 2048): four chains per SM instead of one.  The
        // chain is 1.5x slower (1.42 vs 0.94 ms per 2048 rows x 4096 columns alone) but holds 32 SMs instead of 128, so the
        // encoder of the next tile runs beside it: k = 1024 commit 1.28 -> 1.48e9 elem/s; one rank's round of the exact 8-GPU
        // layout 2.17 -> 1.63 ms (tools/exact_round_sim.py)
        static const int pack_ctas = getenv("LGR_CHAIN_PACK_CTAS") ? atoi(getenv("LGR_CHAIN_PACK_CTAS")) : 64;
        if (n / 32 > pack_ctas && !textbook) LGR_CHAIN_LAUNCH(true, 2, 4)
        if (textbook) LGR_CHAIN_LAUNCH(false, 1, 1)
        if (group == 1) LGR_CHAIN_LAUNCH(true, 1, 1)
        if (group == 2) LGR_CHAIN_LAUNCH(true, 2, 1)
        if (group == 4) LGR_CHAIN_LAUNCH(true, 4, 1)
        LGR_CHAIN_LAUNCH(true, 8, 1)
    }
    // few columns: one warp per CTA so that every chain gets its own scheduler slot
    const int threads = (n <= 148 * 4 * 32) ? 32 : 128;
    sha_update_kernel<<<(n + threads - 1) / threads, threads, 0, st>>>(ctx, n, tile, row_stride, T);
    return cudaGetLastError();
}
cudaError_t launch_sha_final(const uint32_t *ctx, int n, uint32_t *digests, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    sha_final_kernel<<<(n + 127) / 128, 128, 0, st>>>(ctx, n, digests);
    return cudaGetLastError();
}
cudaError_t launch_merkle_build(const uint32_t *leaf, int nleaves, uint32_t *nodes, cudaStream_t st) {
    if (nleaves <= 0) return cudaSuccess;
    int P2 = 1; while (P2 < nleaves) P2 <<= 1;
    merkle_place_leaves<<<min(1024, (P2 * 8 + 255) / 256), 256, 0, st>>>(leaf, nleaves, nodes, P2);
    int cnt = P2 >> 1;
    for (; cnt > 256; cnt >>= 1) merkle_level_kernel<<<(cnt + 127) / 128, 128, 0, st>>>(nodes, cnt);
    if (cnt >= 1) merkle_top_kernel<<<1, 256, 0, st>>>(nodes, cnt);
    return cudaGetLastError();
}
cudaError_t launch_synth(fr_mem *out, uint64_t seed, uint64_t row0, uint64_t nrows, uint64_t ncols, cudaStream_t st) {
    if (nrows * ncols == 0) return cudaSuccess;
    const unsigned long long total = nrows * ncols;
    const int grid = (int)((total + 255) / 256 < 148ull * 16 ? (total + 255) / 256 : 148ull * 16);
    synth_kernel<<<grid, 256, 0, st>>>(out, seed, row0, nrows, ncols);
    return cudaGetLastError();
}

}  // namespace lgr

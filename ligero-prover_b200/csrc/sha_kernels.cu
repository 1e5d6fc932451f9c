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

template <bool REASSOC, int G>
__global__ void __launch_bounds__(128, 1) sha_chain_kernel(uint32_t *ctx, int n, const fr_mem *__restrict__ tile, long long row_stride, int T, const uint32_t one, const uint32_t mone) {
    extern __shared__ __align__(16) unsigned char chain_smem[];
    uint32_t *ring = reinterpret_cast<uint32_t *>(chain_smem);
    uint64_t *full = reinterpret_cast<uint64_t *>(chain_smem + (size_t)kChainSlots * 64 * 32 * 4);
    uint64_t *empty = full + kChainSlots;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int col = blockIdx.x * 32 + lane;                   // n is a multiple of 32
    if (threadIdx.x == 0) {
        for (int s = 0; s < kChainSlots / G; s++) { mbar_init(full + s, 32 * G); mbar_init(empty + s, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t rows_lo = ctx[(size_t)16 * n + col], rows_hi = ctx[(size_t)17 * n + col];
    const int p = (int)(rows_lo & 1u);
    const int nblk = (p + T) >> 1;
    if (warp == 0) {
        uint32_t st[8];
#pragma unroll
        for (int i = 0; i < 8; i++) st[i] = ctx[(size_t)i * n + col];
        // ring slots are handed over in groups of G blocks: one barrier round trip per G compressions
        // (measured, 8192-row tiles: G = 1 3.63 ms, 2 3.48, 4 3.34, 8 3.31 per launch)
        constexpr int NG = kChainSlots / G;
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
        const int hsel = warp - 1;
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
            constexpr int NG = kChainSlots / G;
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
    if (n % 32 == 0 && n / 32 <= 148 && T >= 4) {
        // narrow matrix: producer/consumer CTAs, one chain warp per SM
        static const bool textbook = getenv("LGR_CHAIN_TEXTBOOK") != nullptr;   // A/B knobs: textbook round association, blocks per hand-over
        static const int group = getenv("LGR_CHAIN_GROUP") ? atoi(getenv("LGR_CHAIN_GROUP")) : 8;
#define LGR_CHAIN_LAUNCH(RE, G)                                                                                                         \
    {                                                                                                                                   \
        cudaError_t e = cudaFuncSetAttribute(sha_chain_kernel<RE, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kChainSmem);    \
        if (e != cudaSuccess) return e;                                                                                                 \
        sha_chain_kernel<RE, G><<<n / 32, 128, kChainSmem, st>>>(ctx, n, tile, row_stride, T, 1u, 0xFFFFFFFFu);                         \
        return cudaGetLastError();                                                                                                      \
    }
        if (textbook) LGR_CHAIN_LAUNCH(false, 1)
        if (group == 1) LGR_CHAIN_LAUNCH(true, 1)
        if (group == 2) LGR_CHAIN_LAUNCH(true, 2)
        if (group == 4) LGR_CHAIN_LAUNCH(true, 4)
        LGR_CHAIN_LAUNCH(true, 8)
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

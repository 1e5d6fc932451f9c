// In-CTA NTT building blocks for BN254 Fr on sm_100a.
//
// Replaces the device side of the reference's NTT (shader/kernels.wgsl.in:57-323: one global
// radix-2 dispatch per stage + a 512-point workgroup kernel + separate bit_reverse / adjust
// dispatches, driven by src/webgpu/engine.cpp:844-882,932-968).  Here a transform (or a sub-
// transform of a four-step split) of M = 2^LOGM points lives in shared memory; M/8 threads each
// hold 8 elements (64 registers) and run up to three butterfly stages per shared-memory round
// trip (radix-8 passes), so a 2^10 transform is 4 passes instead of 10+2 global dispatches.
//
// Lazy ranges (same idea as the WGSL, kernels.wgsl.in:121-123,226-228):
//   DIF (natural in -> bit-reversed out): values stay in [0,2p)
//   DIT (bit-reversed in -> natural out): values rest in [0,4p); multiplicand may be < 4p
#pragma once
#include "fr.cuh"
#include "kernels.h"

namespace lgr {

LGR_DEV fr_t fr_unpack(const uint4 &a, const uint4 &b) {
    fr_t r; r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w; return r;
}
// shared / generic 2 x 128-bit
LGR_DEV fr_t fr_lds(const fr_mem *p) { uint4 a = p->lo, b = p->hi; return fr_unpack(a, b); }
LGR_DEV void fr_sts(fr_mem *p, const fr_t &x) {
    p->lo = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]); p->hi = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
}
// global: one 256-bit access per element (LDG.E.256 / STG.E.256 on sm_100)
LGR_DEV fr_t fr_ldg(const fr_mem *p) {
    fr_t r;
    asm volatile("ld.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p));
    return r;
}
LGR_DEV void fr_stg(fr_mem *p, const fr_t &x) {
    asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "r"(x.v[0]), "r"(x.v[1]), "r"(x.v[2]), "r"(x.v[3]), "r"(x.v[4]), "r"(x.v[5]), "r"(x.v[6]), "r"(x.v[7]) : "memory");
}
// read-only tables (twiddles): L1-cached non-coherent path
LGR_DEV fr_t fr_ldc(const fr_mem *p) {
    uint4 a = __ldg(&p->lo), b = __ldg(&p->hi); return fr_unpack(a, b);
}

// address of codeword element (row, column j) in a CodewordSink (kernels.h)
LGR_DEV fr_mem *sink_at(const CodewordSink &s, long long row, int j) {
    if (s.nslabs == 0) return s.base[0] + row * s.row_stride + j;
    return s.base[j >> s.slab_shift] + ((row << s.slab_shift) + (j & ((1 << s.slab_shift) - 1)));
}

LGR_DEV uint32_t bitrev(uint32_t x, int bits) { return bits ? (__brev(x) >> (32 - bits)) : 0u; }

// One pass over index bits [LO, LO+T) of an M-point transform.  `load(idx)` / `store(idx, x)` move the
// elements (shared memory by default; the first / last pass of a fused kernel reads or writes global
// memory directly, applying a twist or a canonicalisation on the way).
// `tl` = thread index within the lane (0 .. max(M/8,1)-1); tw[j*tws] = w^j * R for the M-point root.
template <int LOGM, int LO, int T, bool DIF, typename Load, typename Store>
LGR_DEV void ntt_pass_io(int tl, const fr_mem *__restrict__ tw, int tws, Load load, Store store) {
    constexpr int M = 1 << LOGM;
    constexpr int TL = (M >= 8) ? (M / 8) : 1;
    constexpr int E = M / TL;
    constexpr int R = 1 << T;
    constexpr int G = E / R;
    constexpr int S = 1 << LO;
    static_assert(G >= 1, "radix larger than per-thread element count");
#pragma unroll
    for (int g = 0; g < G; g++) {
        const int q = tl + g * TL;
        const int low = q & (S - 1);
        const int base = low | ((q >> LO) << (LO + T));
        fr_t x[R];
#pragma unroll
        for (int j = 0; j < R; j++) x[j] = load(base + j * S);
#pragma unroll
        for (int ss = 0; ss < T; ss++) {
            const int s = DIF ? (T - 1 - ss) : ss;
            const int b = LO + s;
            const int h = 1 << s;
#pragma unroll
            for (int jl = 0; jl < h; jl++) {
                // twiddle exponent e = low*(M>>(b+1)) + jl*(M>>(s+1)); it is 0 (w = 1, no multiplication)
                // for the whole of stage b = 0 and, in the pass that owns bit 0 (low == 0), whenever
                // jl == 0 -- both known at compile time: 3 of the 8 multiplications of that pass
                const bool mul = (b > 0) && !(LO == 0 && jl == 0);
                fr_t w;
                if (mul) {
                    const int e = low * (M >> (b + 1)) + jl * (M >> (s + 1));
                    w = fr_ldc(tw + (size_t)e * tws);
                }
#pragma unroll
                for (int jh = 0; jh < (R >> (s + 1)); jh++) {
                    const int j0 = jl | (jh << (s + 1));
                    const int j1 = j0 | h;
                    if (DIF) {
                        fr_t u = x[j0], v = x[j1];
                        x[j0] = fr_add_lazy(u, v);
                        x[j1] = mul ? fr_mont_mul(fr_sub_lazy4(u, v), w) : fr_sub_lazy(u, v);
                    } else {
                        fr_t u = fr_reduce_2p(x[j0]);
                        fr_t t = mul ? fr_mont_mul(x[j1], w) : fr_reduce_2p(x[j1]);
                        x[j0] = fr_add_raw(u, t);
                        x[j1] = fr_sub_lazy4(u, t);
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < R; j++) store(base + j * S, x[j]);
    }
}

struct smem_io {
    fr_mem *sm;
    LGR_DEV fr_t operator()(int i) const { return fr_lds(sm + i); }
    LGR_DEV void operator()(int i, const fr_t &x) const { fr_sts(sm + i, x); }
};

// Pass schedule: index bits are consumed three at a time; a tail of four bits is split 2+2 so
// that no pass is a single stage (except the 2-point transform).  DIT walks the partition upward
// (bit-reversed in -> natural out, values rest in [0,4p)), DIF walks it downward (natural in ->
// bit-reversed out, values in [0,2p)).  The first pass executed takes its input from `first`, the
// last pass executed hands its output to `last`; passes in between work in place on `sm`.
// sync() is called between passes (not after the last one).
template <int LOGM, int LO, bool DIF, typename Sync, typename First, typename Last>
LGR_DEV void ntt_passes_io(fr_mem *sm, int tl, const fr_mem *tw, int tws, Sync sync, First first, Last last) {
    if constexpr (LO < LOGM) {
        constexpr int left = LOGM - LO;
        constexpr int T = (left == 4) ? 2 : (left >= 3 ? 3 : left);
        constexpr bool low = (LO == 0), top = (LO + T == LOGM);
        smem_io io{sm};
        if constexpr (DIF) {
            ntt_passes_io<LOGM, LO + T, DIF>(sm, tl, tw, tws, sync, first, last);
            if constexpr (!top) sync();
            if constexpr (top && low) ntt_pass_io<LOGM, LO, T, true>(tl, tw, tws, first, last);
            else if constexpr (top) ntt_pass_io<LOGM, LO, T, true>(tl, tw, tws, first, io);
            else if constexpr (low) ntt_pass_io<LOGM, LO, T, true>(tl, tw, tws, io, last);
            else ntt_pass_io<LOGM, LO, T, true>(tl, tw, tws, io, io);
        } else {
            if constexpr (top && low) ntt_pass_io<LOGM, LO, T, false>(tl, tw, tws, first, last);
            else if constexpr (low) ntt_pass_io<LOGM, LO, T, false>(tl, tw, tws, first, io);
            else if constexpr (top) ntt_pass_io<LOGM, LO, T, false>(tl, tw, tws, io, last);
            else ntt_pass_io<LOGM, LO, T, false>(tl, tw, tws, io, io);
            if constexpr (!top) sync();
            ntt_passes_io<LOGM, LO + T, DIF>(sm, tl, tw, tws, sync, first, last);
        }
    }
}

// M-point DIT entirely in shared memory: sm holds the input in bit-reversed order, result in
// natural order, values in [0,4p); a sync() follows the last pass
template <int LOGM, typename Sync>
LGR_DEV void ntt_dit(fr_mem *sm, int tl, const fr_mem *tw, int tws, Sync sync) {
    ntt_passes_io<LOGM, 0, false>(sm, tl, tw, tws, sync, smem_io{sm}, smem_io{sm});
    sync();
}
// M-point DIF entirely in shared memory: natural order in (values in [0,2p)), bit-reversed out
template <int LOGM, typename Sync>
LGR_DEV void ntt_dif(fr_mem *sm, int tl, const fr_mem *tw, int tws, Sync sync) {
    ntt_passes_io<LOGM, 0, true>(sm, tl, tw, tws, sync, smem_io{sm}, smem_io{sm});
    sync();
}

}  // namespace lgr

// TEST INFRASTRUCTURE (oracle side): the C++ vocabulary the transliterated WGSL (oracle/wgsl2cpp.py) is
// compiled against.  Nothing here restates the reference's algorithms -- it only gives WGSL's built-in
// types and functions their WGSL meaning on the host:
//   * u32 / i32 / f32, vecN<T> with component-wise + - < ==, all(), select(), reverseBits(), log2()
//   * array<T,N> (zero-initialised value type with a value constructor), array<T> (runtime-sized storage
//     binding: pointer + length, arrayLength()), ptr<space,T>
//   * builtins of a compute invocation and workgroupBarrier(): the invocations of ONE workgroup run as
//     cooperative fibers (ucontext), a barrier yields to the next fiber, so every invocation reaches barrier
//     b before any passes it -- WGSL's control barrier, deterministically and on one host thread.
// WGSL shift semantics: the shift count of a u32 is taken modulo 32.  The library is built with
// -fsanitize=shift -fsanitize-undefined-trap-on-error, so a count >= 32 (where C++ and WGSL could differ)
// traps instead of silently diverging; none occurs on the tested paths.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ucontext.h>
#include <vector>

namespace wgsl {

using u32 = uint32_t;
using i32 = int32_t;
using f32 = float;

template <class T> struct vec4 {
    T v[4];
    vec4() : v{T(), T(), T(), T()} {}
    vec4(T a, T b, T c, T d) : v{a, b, c, d} {}
    T& operator[](size_t i) { return v[i]; }
    const T& operator[](size_t i) const { return v[i]; }
};
using vec4u = vec4<u32>;
template <class T> struct vec3 {
    T x, y, z;
    vec3() : x(), y(), z() {}
    vec3(T a, T b, T c) : x(a), y(b), z(c) {}
};
using vec3u = vec3<u32>;

inline vec4u operator+(const vec4u& a, const vec4u& b) { return {a[0] + b[0], a[1] + b[1], a[2] + b[2], a[3] + b[3]}; }
inline vec4u operator-(const vec4u& a, const vec4u& b) { return {a[0] - b[0], a[1] - b[1], a[2] - b[2], a[3] - b[3]}; }
inline vec4<bool> operator<(const vec4u& a, const vec4u& b) { return {a[0] < b[0], a[1] < b[1], a[2] < b[2], a[3] < b[3]}; }
inline vec4<bool> operator==(const vec4u& a, const vec4u& b) { return {a[0] == b[0], a[1] == b[1], a[2] == b[2], a[3] == b[3]}; }
inline bool all(const vec4<bool>& c) { return c[0] && c[1] && c[2] && c[3]; }

// select(f, t, cond): t if cond else f (both operands are evaluated, as in WGSL)
template <class T, class U> inline T select(T f, U t, bool cond) { return cond ? T(t) : f; }

inline u32 reverseBits(u32 x) {
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
    return (x >> 16) | (x << 16);
}
inline f32 log2(f32 x) { return std::log2(x); }

// fixed-size array: value type, zero-initialised, `array<T,N>(a, b, ...)` value constructor
template <class T, u32 N = 0> struct array {
    T v[N];
    array() : v{} {}
    template <class... A, class = std::enable_if_t<sizeof...(A) == N && (N > 0)>>
    array(const A&... a) : v{T(a)...} {}
    T& operator[](size_t i) { return v[i]; }
    const T& operator[](size_t i) const { return v[i]; }
};
// runtime-sized array = a storage-buffer binding
template <class T> struct array<T, 0> {
    T* p = nullptr;
    size_t n = 0;
    T& operator[](size_t i) const {
        if (i >= n) { std::fprintf(stderr, "wgslref: storage access %zu out of %zu\n", i, n); std::abort(); }
        return p[i];
    }
    void bind(void* ptr, size_t count) { p = static_cast<T*>(ptr); n = count; }
};
template <class T> inline u32 arrayLength(const array<T, 0>* a) { return u32(a->n); }

struct function;
struct uniform;
template <class Space, class T> using ptr = T*;

struct builtins {
    vec3u global_invocation_id, local_invocation_id, workgroup_id, num_workgroups;
};
struct entry_point {
    const char* name;
    void (*fn)(const builtins&);
    bool uses_barrier;
};

// ---- workgroup execution ----------------------------------------------------------------------------------
namespace rt {
constexpr u32 WG = 256;                       // every entry point is @workgroup_size(256)
constexpr size_t STACK = 512 << 10;
struct fiber_pool {
    ucontext_t main_ctx;
    ucontext_t fib[WG];
    std::vector<char> stacks;
    bool done[WG];
    int current = -1;
    const entry_point* ep = nullptr;
    builtins b[WG];
};
inline fiber_pool& pool() { static fiber_pool* p = new fiber_pool(); return *p; }

inline void fiber_main() {
    fiber_pool& P = pool();
    int me = P.current;
    P.ep->fn(P.b[me]);
    P.done[me] = true;
    swapcontext(&P.fib[me], &P.main_ctx);
}
inline void barrier() {
    fiber_pool& P = pool();
    if (P.current < 0) { std::fprintf(stderr, "wgslref: workgroupBarrier outside a fiber dispatch\n"); std::abort(); }
    int me = P.current;
    swapcontext(&P.fib[me], &P.main_ctx);
}

// dispatchWorkgroups(nwg, 1, 1)
inline void dispatch(const entry_point& ep, u32 nwg) {
    builtins b;
    b.num_workgroups = vec3u(nwg, 1, 1);
    if (!ep.uses_barrier) {
        for (u32 wg = 0; wg < nwg; ++wg)
            for (u32 t = 0; t < WG; ++t) {
                b.workgroup_id = vec3u(wg, 0, 0);
                b.local_invocation_id = vec3u(t, 0, 0);
                b.global_invocation_id = vec3u(wg * WG + t, 0, 0);
                ep.fn(b);
            }
        return;
    }
    fiber_pool& P = pool();
    if (P.stacks.empty()) P.stacks.resize(size_t(WG) * STACK);
    P.ep = &ep;
    for (u32 wg = 0; wg < nwg; ++wg) {
        for (u32 t = 0; t < WG; ++t) {
            b.workgroup_id = vec3u(wg, 0, 0);
            b.local_invocation_id = vec3u(t, 0, 0);
            b.global_invocation_id = vec3u(wg * WG + t, 0, 0);
            P.b[t] = b;
            P.done[t] = false;
            getcontext(&P.fib[t]);
            P.fib[t].uc_stack.ss_sp = P.stacks.data() + size_t(t) * STACK;
            P.fib[t].uc_stack.ss_size = STACK;
            P.fib[t].uc_link = &P.main_ctx;
            makecontext(&P.fib[t], fiber_main, 0);
        }
        // round-robin: each sweep runs every live invocation up to its next barrier (or to its end)
        for (bool live = true; live;) {
            live = false;
            for (u32 t = 0; t < WG; ++t) {
                if (P.done[t]) continue;
                P.current = int(t);
                swapcontext(&P.main_ctx, &P.fib[t]);
                P.current = -1;
                if (!P.done[t]) live = true;
            }
        }
    }
}
}  // namespace rt

inline void workgroupBarrier() { rt::barrier(); }

}  // namespace wgsl

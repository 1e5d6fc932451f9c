// Stand-in for <gmpxx.h> (not in this image): a value-semantic mpz_class over the real libgmp.so.10, with the
// operator set the reference's prover headers use and gmpxx's documented meanings (/ and % truncate, >> floors,
// get_ui returns the low limb of |x|).  No expression templates: every operator returns an mpz_class.
// Test infrastructure only (see tests/stubs/gmp.h).
#pragma once

#include <gmp.h>

#include <concepts>
#include <cstdint>
#include <cstdlib>
#include <iosfwd>
#include <ostream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>

// bn254_gmp has constructors from `__gmp_expr<Args...>` (finite_field_gmp.hpp:98-99); with value-returning operators
// they are never selected, the template only has to exist.
template <typename... T> struct __gmp_expr;

class mpz_class {
public:
    mpz_class() { mpz_init(z_); }
    mpz_class(const mpz_class &o) { mpz_init(z_); mpz_set(z_, o.z_); }
    mpz_class(mpz_class &&o) noexcept { mpz_init(z_); mpz_swap(z_, o.z_); }
    template <std::integral I> mpz_class(I v) { mpz_init(z_); assign(v); }
    explicit mpz_class(mpz_srcptr z) { mpz_init(z_); mpz_set(z_, z); }
    explicit mpz_class(const char *s, int base = 0) { mpz_init(z_); if (mpz_set_str(z_, s, base) != 0) { mpz_clear(z_); throw std::invalid_argument("mpz_set_str"); } }
    explicit mpz_class(const std::string &s, int base = 0) : mpz_class(s.c_str(), base) {}
    ~mpz_class() { mpz_clear(z_); }

    mpz_class &operator=(const mpz_class &o) { if (this != &o) mpz_set(z_, o.z_); return *this; }
    mpz_class &operator=(mpz_class &&o) noexcept { mpz_swap(z_, o.z_); return *this; }
    template <std::integral I> mpz_class &operator=(I v) { assign(v); return *this; }
    mpz_class &operator=(const char *s) { if (mpz_set_str(z_, s, 0) != 0) throw std::invalid_argument("mpz_set_str"); return *this; }
    mpz_class &operator=(const std::string &s) { return *this = s.c_str(); }

    mpz_ptr get_mpz_t() { return z_; }
    mpz_srcptr get_mpz_t() const { return z_; }
    unsigned long get_ui() const { return mpz_get_ui(z_); }
    long get_si() const { return mpz_get_si(z_); }
    bool fits_ulong_p() const { return mpz_fits_ulong_p(z_) != 0; }
    bool fits_slong_p() const { return mpz_fits_slong_p(z_) != 0; }
    bool fits_uint_p() const { return mpz_fits_uint_p(z_) != 0; }
    bool fits_sint_p() const { return mpz_fits_sint_p(z_) != 0; }
    int set_str(const char *s, int base) { return mpz_set_str(z_, s, base); }
    int set_str(const std::string &s, int base) { return mpz_set_str(z_, s.c_str(), base); }
    std::string get_str(int base = 10) const {
        char *p = mpz_get_str(nullptr, base, z_);
        std::string s(p);
        std::free(p);                                        // GMP's default allocator is malloc
        return s;
    }
    void swap(mpz_class &o) noexcept { mpz_swap(z_, o.z_); }

    mpz_class &operator+=(const mpz_class &o) { mpz_add(z_, z_, o.z_); return *this; }
    mpz_class &operator-=(const mpz_class &o) { mpz_sub(z_, z_, o.z_); return *this; }
    mpz_class &operator*=(const mpz_class &o) { mpz_mul(z_, z_, o.z_); return *this; }
    mpz_class &operator/=(const mpz_class &o) { mpz_tdiv_q(z_, z_, o.z_); return *this; }
    mpz_class &operator%=(const mpz_class &o) { mpz_tdiv_r(z_, z_, o.z_); return *this; }
    mpz_class &operator&=(const mpz_class &o) { mpz_and(z_, z_, o.z_); return *this; }
    mpz_class &operator|=(const mpz_class &o) { mpz_ior(z_, z_, o.z_); return *this; }
    mpz_class &operator^=(const mpz_class &o) { mpz_xor(z_, z_, o.z_); return *this; }
    mpz_class &operator<<=(mp_bitcnt_t n) { mpz_mul_2exp(z_, z_, n); return *this; }
    mpz_class &operator>>=(mp_bitcnt_t n) { mpz_fdiv_q_2exp(z_, z_, n); return *this; }
    mpz_class &operator++() { mpz_add_ui(z_, z_, 1); return *this; }
    mpz_class &operator--() { mpz_sub_ui(z_, z_, 1); return *this; }
    mpz_class operator++(int) { mpz_class t(*this); ++*this; return t; }
    mpz_class operator--(int) { mpz_class t(*this); --*this; return t; }
    mpz_class operator-() const { mpz_class r; mpz_neg(r.z_, z_); return r; }
    mpz_class operator+() const { return *this; }
    mpz_class operator~() const { mpz_class r; mpz_com(r.z_, z_); return r; }
    explicit operator bool() const { return mpz_sgn(z_) != 0; }

private:
    template <std::integral I> void assign(I v) {
        if constexpr (std::is_signed_v<I>) mpz_set_si(z_, (long)v);
        else mpz_set_ui(z_, (unsigned long)v);
    }
    mpz_t z_;
};

template <typename T> concept LgrMpzOperand = std::integral<std::remove_cvref_t<T>> || std::same_as<std::remove_cvref_t<T>, mpz_class>;
template <typename A, typename B> concept LgrMpzPair = LgrMpzOperand<A> && LgrMpzOperand<B> &&
    (std::same_as<std::remove_cvref_t<A>, mpz_class> || std::same_as<std::remove_cvref_t<B>, mpz_class>);

#define LGR_MPZ_BINOP(op, fn)                                                                     \
    template <typename A, typename B> requires LgrMpzPair<A, B>                                   \
    inline mpz_class operator op(const A &a, const B &b) {                                        \
        const mpz_class x(a), y(b);                                                               \
        mpz_class r;                                                                              \
        fn(r.get_mpz_t(), x.get_mpz_t(), y.get_mpz_t());                                          \
        return r;                                                                                 \
    }
LGR_MPZ_BINOP(+, mpz_add)
LGR_MPZ_BINOP(-, mpz_sub)
LGR_MPZ_BINOP(*, mpz_mul)
LGR_MPZ_BINOP(/, mpz_tdiv_q)
LGR_MPZ_BINOP(%, mpz_tdiv_r)
LGR_MPZ_BINOP(&, mpz_and)
LGR_MPZ_BINOP(|, mpz_ior)
LGR_MPZ_BINOP(^, mpz_xor)
#undef LGR_MPZ_BINOP

#define LGR_MPZ_CMP(op)                                                                           \
    template <typename A, typename B> requires LgrMpzPair<A, B>                                   \
    inline bool operator op(const A &a, const B &b) {                                             \
        const mpz_class x(a), y(b);                                                               \
        return mpz_cmp(x.get_mpz_t(), y.get_mpz_t()) op 0;                                        \
    }
LGR_MPZ_CMP(==)
LGR_MPZ_CMP(!=)
LGR_MPZ_CMP(<)
LGR_MPZ_CMP(<=)
LGR_MPZ_CMP(>)
LGR_MPZ_CMP(>=)
#undef LGR_MPZ_CMP

template <std::integral I> inline mpz_class operator<<(const mpz_class &a, I n) { mpz_class r; mpz_mul_2exp(r.get_mpz_t(), a.get_mpz_t(), (mp_bitcnt_t)n); return r; }
template <std::integral I> inline mpz_class operator>>(const mpz_class &a, I n) { mpz_class r; mpz_fdiv_q_2exp(r.get_mpz_t(), a.get_mpz_t(), (mp_bitcnt_t)n); return r; }

// compound assignment with built-in integers on the right
template <std::integral I> inline mpz_class &operator+=(mpz_class &a, I b) { return a += mpz_class(b); }
template <std::integral I> inline mpz_class &operator-=(mpz_class &a, I b) { return a -= mpz_class(b); }
template <std::integral I> inline mpz_class &operator*=(mpz_class &a, I b) { return a *= mpz_class(b); }
template <std::integral I> inline mpz_class &operator/=(mpz_class &a, I b) { return a /= mpz_class(b); }
template <std::integral I> inline mpz_class &operator%=(mpz_class &a, I b) { return a %= mpz_class(b); }
template <std::integral I> inline mpz_class &operator&=(mpz_class &a, I b) { return a &= mpz_class(b); }
template <std::integral I> inline mpz_class &operator|=(mpz_class &a, I b) { return a |= mpz_class(b); }

inline std::ostream &operator<<(std::ostream &os, const mpz_class &v) {
    const auto f = os.flags() & std::ios_base::basefield;
    return os << v.get_str(f == std::ios_base::hex ? 16 : f == std::ios_base::oct ? 8 : 10);
}
inline mpz_class abs(const mpz_class &a) { mpz_class r; mpz_abs(r.get_mpz_t(), a.get_mpz_t()); return r; }
inline int sgn(const mpz_class &a) { return mpz_sgn(a.get_mpz_t()); }
inline void swap(mpz_class &a, mpz_class &b) noexcept { a.swap(b); }

// Stand-in for Boost.Algorithm's hex (absent): include/host_modules/env.hpp:117 prints memory with it.
// Test infrastructure only (see tests/stubs/gmp.h).
#pragma once
#include <iterator>
namespace boost { namespace algorithm {
template <typename In, typename Out> Out hex(In first, In last, Out out) {
    static const char d[] = "0123456789ABCDEF";
    for (; first != last; ++first) { unsigned char c = (unsigned char)*first; *out++ = d[c >> 4]; *out++ = d[c & 15]; }
    return out;
}
template <typename Range, typename Out> Out hex(const Range &r, Out out) { return hex(std::begin(r), std::end(r), out); }
} }

// Stand-in for Boost.ICL (absent): the slice of icl::interval_set / icl::discrete_interval that the reference's
// memory_instance uses to track secret byte ranges (include/runtime.hpp:106-176).  Right-open integer intervals kept
// disjoint and merged in a std::map.  Test infrastructure only (see tests/stubs/gmp.h).
#pragma once
#include <algorithm>
#include <iterator>
#include <map>
#include <utility>

namespace boost { namespace icl {

template <typename T> class discrete_interval {
public:
    discrete_interval() : lo_(0), hi_(0) {}
    discrete_interval(T lo, T hi) : lo_(lo), hi_(hi) {}
    static discrete_interval right_open(T lo, T hi) { return discrete_interval(lo, hi); }
    T lower() const { return lo_; }
    T upper() const { return hi_; }
    bool empty() const { return !(lo_ < hi_); }
private:
    T lo_, hi_;
};

template <typename T> class interval_set {
    using map_t = std::map<T, T>;                       // lower -> upper, right-open
public:
    class const_iterator {
    public:
        using iterator_category = std::forward_iterator_tag;
        using value_type = discrete_interval<T>;
        using difference_type = std::ptrdiff_t;
        using pointer = const value_type *;
        using reference = const value_type &;
        const_iterator() = default;
        explicit const_iterator(typename map_t::const_iterator it) : it_(it) {}
        reference operator*() const { cur_ = value_type(it_->first, it_->second); return cur_; }
        pointer operator->() const { return &**this; }
        const_iterator &operator++() { ++it_; return *this; }
        const_iterator operator++(int) { const_iterator t(*this); ++it_; return t; }
        bool operator==(const const_iterator &o) const { return it_ == o.it_; }
        bool operator!=(const const_iterator &o) const { return it_ != o.it_; }
    private:
        typename map_t::const_iterator it_;
        mutable value_type cur_;
    };
    using iterator = const_iterator;

    const_iterator begin() const { return const_iterator(m_.begin()); }
    const_iterator end() const { return const_iterator(m_.end()); }
    bool empty() const { return m_.empty(); }
    size_t iterative_size() const { return m_.size(); }

    interval_set &operator+=(const discrete_interval<T> &iv) {
        if (iv.empty()) return *this;
        T lo = iv.lower(), hi = iv.upper();
        auto it = m_.lower_bound(lo);
        if (it != m_.begin()) { auto p = std::prev(it); if (!(p->second < lo)) it = p; }   // touching intervals merge
        while (it != m_.end() && !(hi < it->first)) {
            lo = std::min(lo, it->first); hi = std::max(hi, it->second);
            it = m_.erase(it);
        }
        m_[lo] = hi;
        return *this;
    }
    interval_set &operator+=(const interval_set &o) { for (const auto &kv : o.m_) *this += discrete_interval<T>(kv.first, kv.second); return *this; }
    interval_set &operator-=(const discrete_interval<T> &iv) {
        if (iv.empty()) return *this;
        const T lo = iv.lower(), hi = iv.upper();
        auto it = m_.lower_bound(lo);
        if (it != m_.begin()) { auto p = std::prev(it); if (lo < p->second) it = p; }
        while (it != m_.end() && it->first < hi) {
            const T a = it->first, b = it->second;
            it = m_.erase(it);
            if (a < lo) m_[a] = lo;
            if (hi < b) { m_[hi] = b; break; }
        }
        return *this;
    }
    // intervals of the set that overlap iv
    std::pair<const_iterator, const_iterator> equal_range(const discrete_interval<T> &iv) const {
        auto first = m_.lower_bound(iv.lower());
        if (first != m_.begin()) { auto p = std::prev(first); if (iv.lower() < p->second) first = p; }
        auto last = first;
        while (last != m_.end() && last->first < iv.upper()) ++last;
        if (iv.empty()) last = first;
        return {const_iterator(first), const_iterator(last)};
    }
    friend interval_set operator&(const interval_set &s, const discrete_interval<T> &iv) {
        interval_set r;
        auto range = s.equal_range(iv);
        for (auto it = range.first; it != range.second; ++it)
            r += discrete_interval<T>(std::max(it->lower(), iv.lower()), std::min(it->upper(), iv.upper()));
        return r;
    }
    friend interval_set operator&(const discrete_interval<T> &iv, const interval_set &s) { return s & iv; }

private:
    map_t m_;
};

template <typename T> bool intersects(const interval_set<T> &s, const discrete_interval<T> &iv) {
    auto r = s.equal_range(iv);
    return r.first != r.second;
}
template <typename T, typename U> bool contains(const interval_set<T> &s, const U &x) {
    return intersects(s, discrete_interval<T>((T)x, (T)x + 1));
}

} }

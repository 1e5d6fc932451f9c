// oracle_executor.hpp -- the reference's executor concept (include/wgpu.hpp:50-295) implemented on HOST memory with the CPU
// oracle (oracle/lgo.h).  TEST INFRASTRUCTURE ONLY: it lets the reference's own stage contexts
// (include/zkp/nonbatch_context.hpp) and vbn254fr module run in this GPU-less container, so that what they commit can be
// written down as golden vectors (tests/golden/refctx_*.json) for the CUDA path and its restatements to reproduce.
// It shares the buffer / binding types with the CUDA adapter (they are plain ref-counted slices), nothing else.
#pragma once

#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <map>
#include <stdexcept>
#include <vector>

#include <cuda_executor.hpp>   // cuda::buffer_view, cuda::buffer_binding, cuda::eltwise_offset, cuda::device_uint256_t
#include <lgo.h>

namespace ligero {

struct oracle_context {
    using buffer_type = cuda::buffer_view;
    using device_bignum_type = cuda::device_uint256_t;
    struct sha256_context { uint32_t data[64]; uint32_t datalen; uint32_t bitlen[2]; uint32_t state[8]; };

    oracle_context() = default;
    oracle_context(const oracle_context &) = delete;
    ~oracle_context() { for (auto &kv : sha_) lgo_sha_free(kv.second); }

    void webgpu_init(size_t, std::filesystem::path = "") {}
    template <typename Z>
    void ntt_init(uint32_t l, uint32_t k, uint32_t n, const Z &p, const Z &, const Z &wk, const Z &w2k, const Z &wn) {
        if (n != 4 * k || l > k) throw std::runtime_error("oracle_context: geometry");
        lgo_fr m, a, b, c;
        lgo_fr_modulus(&m);
        lgo_omegas(k, &a, &b, &c);
        // the oracle derives its own roots (src/bn254.cpp:51-64 restated): the caller's must be the same ones
        if (!same(device_bignum_type(p), m) || !same(device_bignum_type(wk), a) || !same(device_bignum_type(w2k), b) || !same(device_bignum_type(wn), c))
            throw std::runtime_error("oracle_context: modulus / roots differ from the oracle's");
        l_ = l; k_ = k; n_ = n;
    }
    void device_synchronize() {}
    uint32_t message_size() const { return l_; }
    uint32_t padding_size() const { return k_; }
    uint32_t encoding_size() const { return n_; }

    buffer_type make_device_buffer(size_t num_bytes) {
        void *p = std::calloc(num_bytes ? num_bytes : 1, 1);
        if (!p) throw std::bad_alloc();
        return buffer_type(std::shared_ptr<void>(p, [](void *q) { std::free(q); }), 0, num_bytes);
    }
    buffer_type make_uniform_buffer(size_t b) { return make_device_buffer(b); }
    buffer_type make_message_buffer() { return make_device_buffer((size_t)l_ * 32); }
    buffer_type make_codeword_buffer() { return make_device_buffer((size_t)n_ * 32); }
    buffer_type make_sample_buffer() { return make_device_buffer(cuda::sample_size * 32); }

    // include/ligetron/webgpu/device_context.hpp:85-98 / src/webgpu/device_context.cpp:387-421
    template <typename T> void write_buffer(buffer_type buf, const T *data, size_t len) { fits(buf, len * sizeof(T)); std::memcpy(buf.get(), data, len * sizeof(T)); }
    template <typename T> void write_buffer_clear(buffer_type buf, const T *data, size_t len) {
        fits(buf, len * sizeof(T));
        std::memset(buf.get(), 0, buf.size());
        std::memcpy(buf.get(), data, len * sizeof(T));
    }
    void clear_buffer(buffer_type buf) { std::memset(buf.get(), 0, buf.size()); }
    void copy_buffer_to_buffer(buffer_type from, buffer_type to) { fits(to, from.size()); copy_buffer_to_buffer(from, to, from.size()); }
    void copy_buffer_to_buffer(buffer_type from, buffer_type to, size_t bytes) { std::memmove(to.get(), from.get(), bytes); }
    void copy_buffer_clear(buffer_type from, buffer_type to) {
        fits(to, from.size());
        std::vector<unsigned char> tmp((unsigned char *)from.get(), (unsigned char *)from.get() + from.size());   // the two may overlap
        std::memset(to.get(), 0, to.size());
        std::memcpy(to.get(), tmp.data(), tmp.size());
    }
    template <typename T> std::vector<T> copy_to_host(buffer_type buf) {
        std::vector<T> v(buf.size() / sizeof(T));
        std::memcpy(v.data(), buf.get(), v.size() * sizeof(T));
        return v;
    }
    template <typename Z> void write_limbs(buffer_type buf, const Z &val, size_t size) {
        std::vector<device_bignum_type> host(size, device_bignum_type(val));
        write_buffer(buf, host.data(), host.size());
    }
    template <typename Z> void write_limbs(buffer_type buf, const std::vector<Z> &vals) {
        std::vector<device_bignum_type> host(vals.size());
        for (size_t i = 0; i < vals.size(); i++) host[i] = device_bignum_type(vals[i]);
        write_buffer(buf, host.data(), host.size());
    }

    cuda::buffer_binding bind_scalar(buffer_type s) { return cuda::buffer_binding({s}); }
    cuda::buffer_binding bind_eltwise2(buffer_type x, buffer_type out) { return cuda::buffer_binding({x, out}); }
    cuda::buffer_binding bind_eltwise3(buffer_type x, buffer_type y, buffer_type out) { return cuda::buffer_binding({x, y, out}); }
    cuda::buffer_binding bind_sha256_context(buffer_type c, buffer_type d) { return cuda::buffer_binding({c, d}); }
    cuda::buffer_binding bind_sha256_buffer(buffer_type in) { return cuda::buffer_binding({in}); }
    cuda::buffer_binding bind_sampling(buffer_type from, buffer_type to) { return cuda::buffer_binding({from, to}); }
    cuda::buffer_binding bind_ntt(buffer_type buf) { return cuda::buffer_binding({buf}); }

    // transforms: engine.cpp:755-968 as restated by the oracle
    void encode_ntt_device(cuda::buffer_binding b) { lgo_encode(fr(b, 0), k_); }
    void decode_ntt_device(cuda::buffer_binding b) { lgo_decode(fr(b, 0), k_); }
    void ntt_forward_k(cuda::buffer_binding b) { ntt(b, k_, 0); }
    void ntt_forward_2k(cuda::buffer_binding b) { ntt(b, 2 * k_, 0); }
    void ntt_forward_n(cuda::buffer_binding b) { ntt(b, n_, 0); }
    void ntt_inverse_k(cuda::buffer_binding b) { ntt(b, k_, 1); }
    void ntt_inverse_2k(cuda::buffer_binding b) { ntt(b, 2 * k_, 1); }
    void ntt_inverse_n(cuda::buffer_binding b) { ntt(b, n_, 1); }

    // column hashing: the context buffer only names the stream; the state lives in an lgo_sha
    void sha256_init(size_t ninst) { sha_inst_ = ninst; }
    void sha256_digest_init(cuda::buffer_binding c) { lgo_sha_init(stream(c)); }
    void sha256_digest_update(cuda::buffer_binding c, cuda::buffer_binding buf) { lgo_sha_update(stream(c), fr(buf, 0)); }
    void sha256_digest_final(cuda::buffer_binding c) { lgo_sha_final(stream(c), (uint8_t *)c.buffers()[1].get()); }

    void sampling_init(const std::vector<size_t> &idx) { sample_.assign(idx.begin(), idx.end()); }
    void sample_gather(cuda::buffer_binding b, size_t offset) {
        lgo_gather(fr(b, 1) + offset * sample_.size(), fr(b, 0), sample_.data(), sample_.size());
    }

    // element-wise kernels (kernels.wgsl.in:325-510): the loop runs over arrayLength(vector_x), dynamic offsets shift the windows
    using off = cuda::eltwise_offset;
    void EltwiseAddMod(cuda::buffer_binding b, off o = {}) { lgo_elt_add(fr(b, 2, o.z), fr(b, 0, o.x), fr(b, 1, o.y), len(b)); }
    void EltwiseSubMod(cuda::buffer_binding b, off o = {}) { lgo_elt_sub(fr(b, 2, o.z), fr(b, 0, o.x), fr(b, 1, o.y), len(b)); }
    void EltwiseMultMod(cuda::buffer_binding b, off o = {}) { lgo_elt_mul(fr(b, 2, o.z), fr(b, 0, o.x), fr(b, 1, o.y), len(b)); }
    void EltwiseDivMod(cuda::buffer_binding b, off o = {}) { lgo_elt_div(fr(b, 2, o.z), fr(b, 0, o.x), fr(b, 1, o.y), len(b)); }
    void EltwiseFMAMod(cuda::buffer_binding b, off o = {}) { lgo_elt_fma(fr(b, 2, o.z), fr(b, 0, o.x), fr(b, 1, o.y), len(b)); }
    void EltwiseAddAssignMod(cuda::buffer_binding b, off o = {}) { lgo_elt_add_assign(fr(b, 1, o.z), fr(b, 0, o.x), len(b)); }
    void EltwiseBitDecompose(cuda::buffer_binding b, size_t i, off o = {}) { lgo_elt_bit(fr(b, 1, o.z), fr(b, 0, o.x), (uint32_t)i, len(b)); }
    template <typename Z> void EltwiseAddMod(cuda::buffer_binding b, const Z &c, off o = {}) { lgo_fr s = scalar(c); lgo_elt_add_const(fr(b, 1, o.z), fr(b, 0, o.x), &s, len(b)); }
    template <typename Z> void EltwiseSubConstMod(cuda::buffer_binding b, const Z &c, off o = {}) { lgo_fr s = scalar(c); lgo_elt_sub_const(fr(b, 1, o.z), fr(b, 0, o.x), &s, len(b)); }
    template <typename Z> void EltwiseConstSubMod(cuda::buffer_binding b, const Z &c, off o = {}) { lgo_fr s = scalar(c); lgo_elt_const_sub(fr(b, 1, o.z), fr(b, 0, o.x), &s, len(b)); }
    template <typename Z> void EltwiseMultMod(cuda::buffer_binding b, const Z &c, off o = {}) { lgo_fr s = scalar(c); lgo_elt_mul_const(fr(b, 1, o.z), fr(b, 0, o.x), &s, len(b)); }
    template <typename Z> void EltwiseMontMultMod(cuda::buffer_binding b, const Z &c, off o = {}) { lgo_fr s = scalar(c); lgo_elt_montmul_const(fr(b, 1, o.z), fr(b, 0, o.x), &s, len(b)); }
    template <typename Z> void EltwiseFMAMod(cuda::buffer_binding b, const Z &c, off o = {}) { lgo_fr s = scalar(c); lgo_elt_fma_const(fr(b, 1, o.z), fr(b, 0, o.x), &s, len(b)); }

private:
    static bool same(const device_bignum_type &a, const lgo_fr &b) { return std::memcmp(a.limbs, b.v, 32) == 0; }
    static void fits(const buffer_type &b, size_t bytes) { if (bytes > b.size()) throw std::runtime_error("oracle_context: write past the end of a buffer"); }
    static lgo_fr *fr(const cuda::buffer_binding &b, size_t which, uint32_t elem_off = 0) { return (lgo_fr *)b.buffers()[which].get() + elem_off; }
    static size_t len(const cuda::buffer_binding &b) { return b.buffers()[0].size() / 32; }
    template <typename Z> static lgo_fr scalar(const Z &c) { device_bignum_type d(c); lgo_fr s; std::memcpy(s.v, d.limbs, 32); return s; }
    void ntt(const cuda::buffer_binding &b, size_t N, int inverse) {
        lgo_fr w[3];
        lgo_omegas(k_, &w[0], &w[1], &w[2]);
        lgo_ntt(fr(b, 0), N, &w[N == k_ ? 0 : N == 2 * k_ ? 1 : 2], inverse);
    }
    lgo_sha *stream(const cuda::buffer_binding &c) {
        void *key = c.buffers()[0].get();
        auto it = sha_.find(key);
        if (it == sha_.end()) it = sha_.emplace(key, lgo_sha_new(sha_inst_)).first;
        return it->second;
    }

    uint32_t l_ = 0, k_ = 0, n_ = 0;
    size_t sha_inst_ = 0;
    std::vector<uint32_t> sample_;
    std::map<void *, lgo_sha *> sha_;
};

}  // namespace ligero

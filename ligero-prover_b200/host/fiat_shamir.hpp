// Host-side Fiat-Shamir glue of the prover: the hash transcript, the SHA-256 counter
// PRG, the 192-index sampler and the AES-256-CTR field-element streams, re-stated from
//   include/zkp/hash.hpp:47-129,152-214,341-346      (what `hash << x` feeds for each argument type)
//   include/zkp/random.hpp:87-146                     (hash_random_engine)
//   include/util/portable_sample.hpp:17-33            (partial Fisher-Yates over boost::uniform_int_distribution)
//   include/util/csprng.hpp:28-110 + include/zkp/finite_field_gmp.hpp:70-78   (mpz_random_engine -> field element)
//   src/webgpu_prover.cpp:281-282,337-351             (stage seeds, sampling)
// OpenSSL EVP does the hashing / AES, as in the reference.
//
// PARITY NOTE: boost::random::uniform_int_distribution is not in this image; uniform_int()
// below restates Boost's published generate_uniform_int algorithm (boost/random/uniform_int_distribution.hpp,
// unchanged since 1.47) from memory.  Until it is checked against a real Boost header the sampled indices
// are "parity unpinned"; everything else in this file is pinned by definitions (SHA-256, AES-256-CTR).
#pragma once
#include <openssl/evp.h>

#include <algorithm>
#include <array>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>

namespace ligero::cuda::host {

struct digest {
    uint8_t data[32] = {0};
    bool operator==(const digest &o) const { return !memcmp(data, o.data, 32); }
    bool operator!=(const digest &o) const { return !(*this == o); }
};

// zkp::openssl_hash<sha2_256> (hash.hpp:152-214): streaming, flush_digest() finalises and resets
class sha256 {
public:
    sha256() : ctx_(EVP_MD_CTX_new()) { if (!ctx_) throw std::runtime_error("EVP_MD_CTX_new failed"); reset(); }
    sha256(const sha256 &) = delete;
    sha256 &operator=(const sha256 &) = delete;
    ~sha256() { EVP_MD_CTX_free(ctx_); }
    sha256 &update(const void *p, size_t len) {
        if (len && !EVP_DigestUpdate(ctx_, p, len)) throw std::runtime_error("EVP_DigestUpdate failed");
        return *this;
    }
    // the operator<< overloads of overload_hash (hash.hpp:47-99) for the argument types the prover uses
    sha256 &operator<<(const digest &d) { return update(d.data, 32); }
    sha256 &operator<<(uint64_t v) { return update(&v, 8); }                       // fundamental: raw little-endian bytes
    sha256 &operator<<(const std::string &s) { return update(s.data(), s.size()); }
    sha256 &operator<<(const std::vector<uint32_t> &v) { return update(v.data(), v.size() * 4); }   // element by element = raw bytes
    // a string LITERAL binds to the array overload (hash.hpp:61-65) and is hashed WITH its terminating NUL:
    // hash("LigetronStage1", ...) feeds 15 bytes
    template <size_t N> sha256 &operator<<(const char (&lit)[N]) { return update(lit, N); }
    digest flush_digest() {
        digest d;
        if (!EVP_DigestFinal_ex(ctx_, d.data, nullptr)) throw std::runtime_error("EVP_DigestFinal_ex failed");
        reset();
        return d;
    }

private:
    void reset() { if (!EVP_DigestInit_ex(ctx_, EVP_sha256(), nullptr)) throw std::runtime_error("EVP_DigestInit_ex failed"); }
    EVP_MD_CTX *ctx_;
};

// stage1_seed = hash("LigetronStage1", root, instance_hash)        (src/webgpu_prover.cpp:281-282)
inline digest stage1_seed(const digest &root, const digest &instance_hash) {
    sha256 h;
    h << "LigetronStage1" << root << instance_hash;
    return h.flush_digest();
}
// stage2_seed = hash("LigetronStage2", root, code, linear, quad)   (src/webgpu_prover.cpp:337-341); the three
// vectors are the n x 8 u32 limbs exactly as copied from the device
inline digest stage2_seed(const digest &root, const std::vector<uint32_t> &code, const std::vector<uint32_t> &linear,
                          const std::vector<uint32_t> &quad) {
    sha256 h;
    h << "LigetronStage2" << root << code << linear << quad;
    return h.flush_digest();
}

// zkp::hash_random_engine<sha256> (random.hpp:87-146): byte generator.  Block i = SHA-256(prefix || LE64(i))
// where prefix is empty for the first block drawn and the 32-byte seed afterwards (the seed is pushed
// into the hasher only AFTER each flush, random.hpp:134-137); bytes are handed out from index 31 down.
class hash_random_engine {
public:
    using result_type = uint8_t;
    static constexpr result_type min() { return 0; }
    static constexpr result_type max() { return 255; }
    explicit hash_random_engine(const digest &seed) : seed_(seed) {}
    result_type operator()() {
        if (offset_ < 0 || offset_ >= 32) {
            hash_ << state_++;
            buffer_ = hash_.flush_digest();
            hash_ << seed_;
            offset_ = 31;
        }
        return buffer_.data[offset_--];
    }

private:
    sha256 hash_;
    digest seed_, buffer_;
    uint64_t state_ = 0;
    int32_t offset_ = -1;
};

// boost::random::detail::generate_uniform_int(eng, min, max) for an 8-bit engine and a 64-bit result
// (see the PARITY NOTE above).  brange = 255.
template <typename Engine>
uint64_t uniform_int(Engine &eng, uint64_t min_value, uint64_t max_value) {
    const uint64_t range = max_value - min_value;
    const uint64_t brange = 255;
    if (range == 0) return min_value;
    if (brange == range) return (uint64_t)eng() + min_value;
    if (brange < range) {
        for (;;) {
            uint64_t limit;
            if (range == UINT64_MAX) {
                limit = range / (brange + 1);
                if (range % (brange + 1) == brange) ++limit;
            } else {
                limit = (range + 1) / (brange + 1);
            }
            uint64_t result = 0, mult = 1;
            while (mult <= limit) {
                result += (uint64_t)eng() * mult;
                if (mult * brange == range - mult + 1) return result;   // (sic: Boost returns without adding min_value here)
                mult *= brange + 1;
            }
            uint64_t inc = uniform_int(eng, 0, range / mult);
            if (UINT64_MAX / mult < inc) continue;
            inc *= mult;
            result += inc;
            if (result < inc) continue;
            if (result > range) continue;
            return result + min_value;
        }
    }
    // brange > range: bucket method in the engine's own unsigned type (brange == its maximum)
    uint8_t bucket = (uint8_t)(brange / (range + 1));
    if (brange % (range + 1) == range) ++bucket;
    for (;;) {
        uint8_t r = (uint8_t)eng();
        r = (uint8_t)(r / bucket);
        if (r <= range) return (uint64_t)r + min_value;
    }
}

// portable_sample(iota(n), sample_size, engine) followed by std::sort (src/webgpu_prover.cpp:343-351)
inline std::vector<uint64_t> sample_indices(const digest &stage2_seed, uint64_t n, uint64_t sample_size) {
    hash_random_engine eng(stage2_seed);
    std::vector<uint64_t> idx(n), out;
    std::iota(idx.begin(), idx.end(), 0);
    uint64_t cnt = std::min(sample_size, n);
    for (uint64_t i = 0; i < cnt; i++) {
        const uint64_t j = uniform_int(eng, i, n - 1);
        std::swap(idx[i], idx[j]);
        out.push_back(idx[i]);
    }
    std::sort(out.begin(), out.end());
    return out;
}

// mpz_random_engine (csprng.hpp:28-110) + bn254_gmp::generate_random (finite_field_gmp.hpp:70-78): the
// AES-256-CTR keystream (zero plaintext, 16 KiB refills that continue the counter), 32 bytes per draw read
// as four little-endian u64 (least significant first), >> 2, one conditional subtraction of p.
class fr_random_stream {
public:
    static constexpr size_t kBufWords = 16384 / 8;
    fr_random_stream() = default;
    fr_random_stream(const uint8_t key[32], const uint8_t iv[16]) { init(key, iv); }
    fr_random_stream(const fr_random_stream &) = delete;
    fr_random_stream &operator=(const fr_random_stream &) = delete;
    ~fr_random_stream() { if (ctx_) EVP_CIPHER_CTX_free(ctx_); }
    void init(const uint8_t key[32], const uint8_t iv[16]) {
        if (!ctx_) ctx_ = EVP_CIPHER_CTX_new();
        if (!ctx_ || 1 != EVP_EncryptInit_ex(ctx_, EVP_aes_256_ctr(), nullptr, key, iv)) throw std::runtime_error("AES-CTR init failed");
        fill();
    }
    // next element as 8 x u32 little-endian canonical limbs
    void next(uint32_t out[8]) {
        if (!ctx_) throw std::runtime_error("fr_random_stream not initialised");
        if (off_ + 4 > kBufWords) fill();
        uint64_t w[4];
        memcpy(w, buf_ + off_, 32);
        off_ += 4;
        w[0] = (w[0] >> 2) | (w[1] << 62); w[1] = (w[1] >> 2) | (w[2] << 62); w[2] = (w[2] >> 2) | (w[3] << 62); w[3] >>= 2;
        static const uint64_t P[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
        bool ge = true;
        for (int i = 3; i >= 0; i--) { if (w[i] != P[i]) { ge = w[i] > P[i]; break; } }
        if (ge) {
            unsigned __int128 br = 0;
            for (int i = 0; i < 4; i++) { unsigned __int128 d = (unsigned __int128)w[i] - P[i] - (uint64_t)br; w[i] = (uint64_t)d; br = (d >> 64) & 1; }
        }
        memcpy(out, w, 32);
    }

private:
    void fill() {
        static const std::array<uint8_t, 16384> zeros{};
        int len = 0;
        if (1 != EVP_EncryptUpdate(ctx_, reinterpret_cast<uint8_t *>(buf_), &len, zeros.data(), 16384) || len != 16384)
            throw std::runtime_error("AES-CTR keystream failed");
        off_ = 0;
    }
    EVP_CIPHER_CTX *ctx_ = nullptr;
    alignas(16) uint64_t buf_[kBufWords];
    size_t off_ = kBufWords;
};

}  // namespace ligero::cuda::host

// Proof container: the LigeroProofEnvelope protobuf of proto/ligero_proof.proto:13-60 and
// proto/common.proto:21-33, written and read on the proto3 wire directly (no protoc / libprotobuf in the
// image), then gzip (src/webgpu_prover.cpp:437-446, level 6).  Field order, packed `repeated fixed32`, and
// the canonical sibling order follow include/zkp/proof_serializer.hpp:60-224.
//
// What is comparable with the reference: the DECOMPRESSED envelope, byte for byte, for equal field values
// (libprotobuf's C++ serializer emits known fields in field-number order, omits proto3 zero scalars and
// empty repeated fields, and writes a set sub-message even when it is empty).  The gzip wrapper itself is
// not (Boost.Iostreams and zlib write different OS / XFL header bytes) and `generated_at` is wall-clock time.
#pragma once
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <exception>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "merkle_host.hpp"

namespace ligero::cuda::host {

// ---- wire primitives ---------------------------------------------------------------------------
namespace wire {
inline void varint(std::string &o, uint64_t v) {
    while (v >= 0x80) { o.push_back((char)(v | 0x80)); v >>= 7; }
    o.push_back((char)v);
}
inline void tag(std::string &o, uint32_t field, uint32_t type) { varint(o, (uint64_t)field << 3 | type); }
inline void put_varint(std::string &o, uint32_t field, uint64_t v) { if (v) { tag(o, field, 0); varint(o, v); } }   // proto3: zero is omitted
inline void put_bytes(std::string &o, uint32_t field, const void *p, size_t n, bool always = false) {
    if (!n && !always) return;
    tag(o, field, 2); varint(o, n); o.append(static_cast<const char *>(p), n);
}
inline void put_message(std::string &o, uint32_t field, const std::string &m) { put_bytes(o, field, m.data(), m.size(), true); }

struct reader {
    const uint8_t *p, *end;
    reader(const void *data, size_t n) : p(static_cast<const uint8_t *>(data)), end(p + n) {}
    bool done() const { return p >= end; }
    uint64_t varint() {
        uint64_t v = 0; int shift = 0;
        for (;;) {
            if (p >= end || shift > 63) throw std::runtime_error("proof: truncated varint");
            const uint8_t b = *p++;
            v |= (uint64_t)(b & 0x7f) << shift;
            if (!(b & 0x80)) return v;
            shift += 7;
        }
    }
    // next field: returns field number, sets type and (for length-delimited) the payload span
    uint32_t next(uint32_t &type, uint64_t &value, const uint8_t *&data, size_t &len) {
        const uint64_t t = varint();
        type = (uint32_t)(t & 7);
        data = nullptr; len = 0; value = 0;
        switch (type) {
            case 0: value = varint(); break;
            case 1: if (end - p < 8) throw std::runtime_error("proof: truncated fixed64"); memcpy(&value, p, 8); p += 8; break;
            case 2: len = (size_t)varint(); if ((size_t)(end - p) < len) throw std::runtime_error("proof: truncated field"); data = p; p += len; break;
            case 5: { if (end - p < 4) throw std::runtime_error("proof: truncated fixed32"); uint32_t v; memcpy(&v, p, 4); value = v; p += 4; break; }
            default: throw std::runtime_error("proof: unsupported wire type");
        }
        return (uint32_t)(t >> 3);
    }
};
}  // namespace wire

// ---- messages ----------------------------------------------------------------------------------
struct proof_metadata {                        // ligero.v1.ProofMetadata (set in src/webgpu_prover.cpp:410-427)
    std::string prover_version = "1.5.0";      // reference VERSION
    uint32_t proof_schema_version = 1;
    uint32_t proof_type = 1;                   // PROOF_TYPE_CLASSIC
    digest program_hash;
    int64_t generated_at_seconds = 0;
    uint32_t packing_size = 0;                 // carries k (src/webgpu_prover.cpp:418)
    uint32_t codeword_size = 0;                // n
    uint32_t sample_size = 192;
    uint32_t security_level = 128;
};

struct proof_data {                            // zkp::ProofData + metadata (proof_serializer.hpp:40-58)
    proof_metadata meta;
    digest merkle_root;
    decommitment decommit;
    std::vector<uint32_t> code, linear, quad;  // n x 8 limbs each
    std::vector<uint32_t> samplings;           // [row][sample][8]
};

inline std::string encode_hash_digest(const digest &d) { std::string m; wire::put_bytes(m, 1, d.data, 32); return m; }
inline std::string encode_fixed_u32_vector(const std::vector<uint32_t> &v) {
    std::string m;
    wire::put_bytes(m, 1, v.data(), v.size() * 4);          // packed repeated fixed32 = raw little-endian words
    return m;
}

// serialize_proof (proof_serializer.hpp:166-191) -> LigeroProofEnvelope bytes
inline std::string serialize_proof(const proof_data &pd) {
    std::string meta;
    wire::put_bytes(meta, 1, pd.meta.prover_version.data(), pd.meta.prover_version.size());
    wire::put_varint(meta, 2, pd.meta.proof_schema_version);
    wire::put_varint(meta, 3, pd.meta.proof_type);
    wire::put_message(meta, 4, encode_hash_digest(pd.meta.program_hash));
    { std::string ts; wire::put_varint(ts, 1, (uint64_t)pd.meta.generated_at_seconds); wire::put_message(meta, 5, ts); }
    wire::put_varint(meta, 6, pd.meta.packing_size);
    wire::put_varint(meta, 7, pd.meta.codeword_size);
    wire::put_varint(meta, 8, pd.meta.sample_size);
    wire::put_varint(meta, 9, pd.meta.security_level);

    std::string mt;                                         // ligero.common.v1.MerkleDecommitment
    wire::put_varint(mt, 1, 1);                             // HASH_ALGORITHM_SHA256
    wire::put_message(mt, 2, encode_hash_digest(pd.merkle_root));
    for (const digest &s : pd.decommit.siblings) wire::put_message(mt, 3, encode_hash_digest(s));
    if (!pd.decommit.known_index.empty()) {                 // packed repeated uint32
        std::string packed;
        for (uint64_t i : pd.decommit.known_index) wire::varint(packed, (uint32_t)i);
        wire::put_bytes(mt, 4, packed.data(), packed.size());
    }

    std::string proof;                                      // ligero.v1.LigeroProof
    wire::put_message(proof, 1, mt);
    wire::put_message(proof, 2, encode_fixed_u32_vector(pd.code));
    wire::put_message(proof, 3, encode_fixed_u32_vector(pd.linear));
    wire::put_message(proof, 4, encode_fixed_u32_vector(pd.quad));
    wire::put_message(proof, 5, encode_fixed_u32_vector(pd.samplings));

    std::string env;                                        // ligero.v1.LigeroProofEnvelope
    wire::put_message(env, 1, meta);
    wire::put_message(env, 2, proof);
    return env;
}

namespace detail {
inline digest parse_hash_digest(const uint8_t *d, size_t n) {
    wire::reader r(d, n); digest out;
    while (!r.done()) {
        uint32_t type; uint64_t v; const uint8_t *p; size_t len;
        if (r.next(type, v, p, len) == 1 && type == 2) { if (len != 32) throw std::runtime_error("proof: digest is not 32 bytes"); memcpy(out.data, p, 32); }
    }
    return out;
}
inline std::vector<uint32_t> parse_fixed_u32_vector(const uint8_t *d, size_t n) {
    wire::reader r(d, n); std::vector<uint32_t> out;
    while (!r.done()) {
        uint32_t type; uint64_t v; const uint8_t *p; size_t len;
        const uint32_t f = r.next(type, v, p, len);
        if (f != 1) continue;
        if (type == 2) { if (len % 4) throw std::runtime_error("proof: packed fixed32 length"); const size_t o = out.size(); out.resize(o + len / 4); memcpy(out.data() + o, p, len); }
        else if (type == 5) out.push_back((uint32_t)v);      // unpacked encoding is legal on the wire too
    }
    return out;
}
}  // namespace detail

// deserialize_proof (proof_serializer.hpp:193-224)
inline proof_data deserialize_proof(const std::string &bytes) {
    proof_data pd;
    const uint8_t *proof = nullptr; size_t proof_len = 0; bool have_proof = false;
    wire::reader env(bytes.data(), bytes.size());
    while (!env.done()) {
        uint32_t type; uint64_t v; const uint8_t *p; size_t len;
        const uint32_t f = env.next(type, v, p, len);
        if (f == 1 && type == 2) {
            wire::reader m(p, len);
            while (!m.done()) {
                uint32_t t2; uint64_t v2; const uint8_t *p2; size_t l2;
                switch (m.next(t2, v2, p2, l2)) {
                    case 1: pd.meta.prover_version.assign(reinterpret_cast<const char *>(p2), l2); break;
                    case 2: pd.meta.proof_schema_version = (uint32_t)v2; break;
                    case 3: pd.meta.proof_type = (uint32_t)v2; break;
                    case 4: pd.meta.program_hash = detail::parse_hash_digest(p2, l2); break;
                    case 5: { wire::reader ts(p2, l2); while (!ts.done()) { uint32_t t3; uint64_t v3; const uint8_t *p3; size_t l3; if (ts.next(t3, v3, p3, l3) == 1) pd.meta.generated_at_seconds = (int64_t)v3; } break; }
                    case 6: pd.meta.packing_size = (uint32_t)v2; break;
                    case 7: pd.meta.codeword_size = (uint32_t)v2; break;
                    case 8: pd.meta.sample_size = (uint32_t)v2; break;
                    case 9: pd.meta.security_level = (uint32_t)v2; break;
                    default: break;
                }
            }
        } else if (f == 2 && type == 2) { proof = p; proof_len = len; have_proof = true; }
    }
    if (!have_proof) throw std::runtime_error("Proof envelope does not contain a LigeroProof payload");
    if (pd.meta.codeword_size == 0) throw std::runtime_error("Proof metadata missing codeword_size");
    size_t leaves = 1; while (leaves < pd.meta.codeword_size) leaves <<= 1;
    pd.decommit.total_count = 2 * leaves - 1;
    wire::reader pr(proof, proof_len);
    while (!pr.done()) {
        uint32_t type; uint64_t v; const uint8_t *p; size_t len;
        const uint32_t f = pr.next(type, v, p, len);
        if (type != 2) continue;
        if (f == 1) {
            wire::reader m(p, len);
            while (!m.done()) {
                uint32_t t2; uint64_t v2; const uint8_t *p2; size_t l2;
                switch (m.next(t2, v2, p2, l2)) {
                    case 2: pd.merkle_root = detail::parse_hash_digest(p2, l2); break;
                    case 3: pd.decommit.siblings.push_back(detail::parse_hash_digest(p2, l2)); break;
                    case 4:
                        if (t2 == 2) { wire::reader ix(p2, l2); while (!ix.done()) pd.decommit.known_index.push_back(ix.varint()); }
                        else pd.decommit.known_index.push_back(v2);
                        break;
                    default: break;
                }
            }
        } else if (f == 2) pd.code = detail::parse_fixed_u32_vector(p, len);
        else if (f == 3) pd.linear = detail::parse_fixed_u32_vector(p, len);
        else if (f == 4) pd.quad = detail::parse_fixed_u32_vector(p, len);
        else if (f == 5) pd.samplings = detail::parse_fixed_u32_vector(p, len);
    }
    pd.decommit.positions = sibling_positions(pd.decommit.known_index, pd.decommit.total_count);
    if (pd.decommit.positions.size() != pd.decommit.siblings.size())
        throw std::runtime_error("Sibling hash count mismatch: expected " + std::to_string(pd.decommit.positions.size()) + ", got " +
                                 std::to_string(pd.decommit.siblings.size()));
    return pd;
}

// ---- gzip (src/webgpu_prover.cpp:437-446: level 6) ---------------------------------------------
// One gzip member, deflate level 6 as the reference asks.  Proofs are mostly sampled field elements
// (incompressible), so zlib runs at ~45 MB/s on one core; inputs above 8 MiB are therefore cut into chunks that
// are deflated in parallel as raw streams, each ended with a sync flush (byte-aligned empty stored block), and
// stitched into a single member with the CRCs combined -- the construction pigz uses; any inflater reads it.
namespace detail {
inline std::string deflate_raw_chunk(const char *data, size_t len, int level, bool last) {
    z_stream zs{};
    if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) throw std::runtime_error("deflateInit2 failed");
    std::string out(deflateBound(&zs, len) + 64, '\0');
    zs.next_in = reinterpret_cast<Bytef *>(const_cast<char *>(data)); zs.avail_in = (uInt)len;
    zs.next_out = reinterpret_cast<Bytef *>(&out[0]); zs.avail_out = (uInt)out.size();
    const int rc = deflate(&zs, last ? Z_FINISH : Z_SYNC_FLUSH);
    const bool ok = last ? rc == Z_STREAM_END : (rc == Z_OK && zs.avail_in == 0);
    out.resize(zs.total_out);
    deflateEnd(&zs);
    if (!ok) throw std::runtime_error("deflate failed");
    return out;
}
}  // namespace detail

inline std::string gzip_compress(const std::string &in, int level = 6, unsigned threads = 0) {
    const size_t chunk = (size_t)8 << 20;
    const size_t nchunks = in.empty() ? 1 : (in.size() + chunk - 1) / chunk;
    std::vector<std::string> parts(nchunks);
    std::vector<uLong> crcs(nchunks);
    auto work = [&](size_t c) {
        const size_t off = c * chunk, len = std::min(chunk, in.size() - off);
        parts[c] = detail::deflate_raw_chunk(in.data() + off, len, level, c + 1 == nchunks);
        crcs[c] = crc32(crc32(0L, Z_NULL, 0), reinterpret_cast<const Bytef *>(in.data() + off), (uInt)len);
    };
    if (threads == 0) threads = std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
    if (nchunks == 1 || threads == 1) {
        for (size_t c = 0; c < nchunks; c++) work(c);
    } else {
        std::atomic<size_t> next{0};
        std::exception_ptr err;
        std::mutex mu;
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < std::min<size_t>(threads, nchunks); t++)
            pool.emplace_back([&] {
                try { for (size_t c; (c = next.fetch_add(1)) < nchunks;) work(c); }
                catch (...) { std::lock_guard<std::mutex> g(mu); err = std::current_exception(); }
            });
        for (auto &th : pool) th.join();
        if (err) std::rethrow_exception(err);
    }
    std::string out;
    static const unsigned char hdr[10] = {0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 0, 3};      // no name, mtime 0, OS = Unix (what zlib writes)
    out.append(reinterpret_cast<const char *>(hdr), 10);
    uLong crc = crc32(0L, Z_NULL, 0);
    for (size_t c = 0; c < nchunks; c++) {
        out += parts[c];
        crc = crc32_combine(crc, crcs[c], (z_off_t)std::min(chunk, in.size() - c * chunk));
    }
    const uint32_t trailer[2] = {(uint32_t)crc, (uint32_t)(in.size() & 0xFFFFFFFFu)};
    out.append(reinterpret_cast<const char *>(trailer), 8);
    return out;
}
inline std::string gzip_decompress(const std::string &in) {
    z_stream zs{};
    if (inflateInit2(&zs, 15 + 32) != Z_OK) throw std::runtime_error("inflateInit2 failed");
    std::string out;
    std::vector<char> buf(1 << 22);
    size_t fed = 0;
    int rc = Z_OK;
    do {
        if (zs.avail_in == 0 && fed < in.size()) {
            const size_t take = std::min<size_t>(in.size() - fed, (size_t)1 << 30);
            zs.next_in = reinterpret_cast<Bytef *>(const_cast<char *>(in.data() + fed)); zs.avail_in = (uInt)take;
            fed += take;
        }
        zs.next_out = reinterpret_cast<Bytef *>(buf.data()); zs.avail_out = (uInt)buf.size();
        rc = inflate(&zs, Z_NO_FLUSH);
        if (rc != Z_OK && rc != Z_STREAM_END) { inflateEnd(&zs); throw std::runtime_error("inflate failed"); }
        out.append(buf.data(), buf.size() - zs.avail_out);
        if (rc == Z_OK && zs.avail_in == 0 && fed >= in.size() && zs.avail_out != 0) { inflateEnd(&zs); throw std::runtime_error("truncated gzip stream"); }
    } while (rc != Z_STREAM_END);
    inflateEnd(&zs);
    return out;
}

}  // namespace ligero::cuda::host

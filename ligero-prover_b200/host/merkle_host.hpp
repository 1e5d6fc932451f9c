// Host-side Merkle openings over the node array that lgr_merkle_build produces (the reference's heap
// layout: include/zkp/merkle_tree.hpp:343-375): decommit, canonical sibling order, recommit.
// Re-stated from include/zkp/merkle_tree.hpp:155-215 (decommit), :232-318 (recommit) and
// include/zkp/proof_serializer.hpp:82-117 (compute_sibling_positions).
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "fiat_shamir.hpp"

namespace ligero::cuda::host {

// Tree positions (indices into the heap array of `total_count` = 2*leaves-1 nodes) of the sibling digests
// an opening of `leaf_indices` needs, bottom-up and left-to-right within a level.  "local" indices count
// from the start of a level, exactly as merkle_tree::decommit_helper does.
inline std::vector<size_t> sibling_positions(const std::vector<uint64_t> &leaf_indices, size_t total_count) {
    std::vector<size_t> positions;
    std::unordered_set<size_t> known(leaf_indices.begin(), leaf_indices.end());
    size_t start = total_count / 2, end = total_count;
    while (start > 0) {
        std::unordered_set<size_t> upper;
        for (size_t i = start; i < end; i += 2) {
            const size_t ll = i - start, lr = ll + 1, lp = ll / 2;
            const bool kl = known.count(ll) != 0, kr = known.count(lr) != 0;
            if (kl && kr) upper.insert(lp);
            else if (kr) { positions.push_back(i); upper.insert(lp); }
            else if (kl) { positions.push_back(i + 1); upper.insert(lp); }
        }
        known.swap(upper);
        start = (start - 1) / 2;
        end = (end - 1) / 2;
    }
    return positions;
}

struct decommitment {
    size_t total_count = 0;                    // nodes in the tree
    std::vector<uint64_t> known_index;         // opened leaves (sorted by the caller, src/webgpu_prover.cpp:351)
    std::vector<size_t> positions;             // sibling positions in canonical order
    std::vector<digest> siblings;              // one digest per position
};

// merkle_tree::decommit: nodes = total_count x 32 bytes (root first, leaves last)
inline decommitment decommit(const uint8_t *nodes, size_t total_count, const std::vector<uint64_t> &known_index) {
    decommitment d;
    d.total_count = total_count;
    d.known_index = known_index;
    d.positions = sibling_positions(known_index, total_count);
    d.siblings.resize(d.positions.size());
    for (size_t i = 0; i < d.positions.size(); i++) memcpy(d.siblings[i].data, nodes + d.positions[i] * 32, 32);
    return d;
}

inline digest hash_pair(const digest &l, const digest &r) {
    sha256 h;
    h << l << r;
    return h.flush_digest();
}

// merkle_tree::recommit(vector<digest>, decommitment): leaves[i] is the digest of leaf known_index[i]
inline digest recommit(const std::vector<digest> &leaves, const decommitment &d) {
    if (leaves.size() != d.known_index.size()) throw std::invalid_argument("recommit: one digest per opened leaf expected");
    std::unordered_map<size_t, digest> saved;
    for (size_t i = 0; i < d.positions.size(); i++) saved.emplace(d.positions[i], d.siblings[i]);
    std::vector<digest> buffer(d.total_count / 2 + 1);
    for (size_t i = 0; i < leaves.size(); i++) buffer[d.known_index[i]] = leaves[i];
    std::unordered_set<size_t> known(d.known_index.begin(), d.known_index.end());
    size_t start = d.total_count / 2, end = d.total_count;
    while (start > 0) {
        std::unordered_set<size_t> upper;
        for (size_t i = start; i < end; i += 2) {
            const size_t ll = i - start, lr = ll + 1, lp = ll / 2;
            const bool kl = known.count(ll) != 0, kr = known.count(lr) != 0;
            if (kl && kr) buffer[lp] = hash_pair(buffer[ll], buffer[lr]);
            else if (kr) buffer[lp] = hash_pair(saved.at(i), buffer[lr]);
            else if (kl) buffer[lp] = hash_pair(buffer[ll], saved.at(i + 1));
            else continue;
            upper.insert(lp);
        }
        known.swap(upper);
        start = (start - 1) / 2;
        end = (end - 1) / 2;
    }
    return buffer[0];
}

}  // namespace ligero::cuda::host

// Packing of committed witnesses into rows: the last step of the reference's backend before the hot path
// (the seam to the interpreter).  Re-stated from witness_manager::commit_release_witness,
// process_reset_linear_row / process_reset_quadratic_rows and finalize
// (include/zkp/backend/witness_manager.hpp:117-186,188-269,497-503):
//   * a released linear witness (value, linear-test coefficient) is appended to the open linear row; a released
//     quadratic slot appends x, y, z (each with its coefficient) to the three open quadratic rows;
//   * a row is emitted LAZILY -- when a witness arrives and the open row already holds l of them -- so the order
//     of row events (which fixes the order of rows in every column hash) depends on when the
//     (l+1)-th witness of each kind shows up;
//   * finalize emits the partly filled linear row, then the partly filled triple (zero-filled to l); masks are
//     appended by the prover (matrix_prover.hpp).
// Output is the row-event list matrix_prover / lgrp_prove take.  What stays outside is the producer of the
// witnesses themselves (interpreter + ligetron_backend expression logic).
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstring>
#include <vector>

namespace ligero::cuda::host {

class row_packer {
public:
    // keep_values / keep_coefs: a prover pass that needs only one of the two (stage 1 commits the values, stage 2 re-runs the
    // program for the coefficients) does not store the other; the row events and counts are the same either way
    explicit row_packer(uint32_t l, bool keep_values = true, bool keep_coefs = true) : l_(l), keep_values_(keep_values), keep_coefs_(keep_coefs) {
        const size_t row = (size_t)l * 8;                                  // the open rows: written in place, zero beyond what was pushed
        if (keep_values) { lin_val_.assign(row, 0); for (auto &v : quad_val_) v.assign(row, 0); }
        if (keep_coefs) { lin_coef_.assign(row, 0); for (auto &v : quad_coef_) v.assign(row, 0); }
    }

    // commit_status::linear_ready
    void push_linear(const uint32_t value[8], const uint32_t coef[8]) {
        if (lin_count_ >= l_) flush_linear();
        if (keep_values_) memcpy(lin_val_.data() + (size_t)lin_count_ * 8, value, 32);
        if (keep_coefs_) memcpy(lin_coef_.data() + (size_t)lin_count_ * 8, coef, 32);
        lin_count_++;
    }
    // commit_status::quadratic_ready: one slot = (x, y, z) with x*y = z, plus their linear-test coefficients
    void push_quadratic(const uint32_t x[8], const uint32_t y[8], const uint32_t z[8], const uint32_t cx[8], const uint32_t cy[8], const uint32_t cz[8]) {
        if (quad_count_ >= l_) flush_quadratic();
        const size_t at = (size_t)quad_count_ * 8;
        if (keep_values_) { memcpy(quad_val_[0].data() + at, x, 32); memcpy(quad_val_[1].data() + at, y, 32); memcpy(quad_val_[2].data() + at, z, 32); }
        if (keep_coefs_) { memcpy(quad_coef_[0].data() + at, cx, 32); memcpy(quad_coef_[1].data() + at, cy, 32); memcpy(quad_coef_[2].data() + at, cz, 32); }
        quad_count_++;
    }
    // witness_manager::finalize (the mask rows are the prover's business)
    void finalize() { flush_linear(); flush_quadratic(); }

    uint32_t l() const { return l_; }
    const std::vector<uint8_t> &kinds() const { return kinds_; }          // per event: 0 linear row, 1 quadratic triple
    const std::vector<uint32_t> &values() const { return values_; }       // encoded rows in emission order: [rows][l][8]
    const std::vector<uint32_t> &coefs() const { return coefs_; }
    std::vector<uint32_t> take_coefs() { return std::move(coefs_); }
    void reserve_rows(size_t rows) {                                      // a second pass over the same program knows how many rows come
        if (keep_values_) values_.reserve(rows * (size_t)l_ * 8);
        if (keep_coefs_) coefs_.reserve(rows * (size_t)l_ * 8);
    }
    size_t rows() const { return (keep_values_ ? values_.size() : coefs_.size()) / ((size_t)l_ * 8); }
    uint64_t linear_count() const { return linear_total_; }               // "Num Linear constraints" / "Num quadratic constraints"
    uint64_t quadratic_count() const { return quadratic_total_; }

private:
    // the open row joins the output, zero-filled beyond its `count` witnesses (push_back_zeros(row_size - data_size)), and is cleared
    void emit(std::vector<uint32_t> &val, std::vector<uint32_t> &coef, uint32_t count) {
        if (keep_values_) { values_.insert(values_.end(), val.begin(), val.end()); std::fill_n(val.begin(), (size_t)count * 8, 0u); }
        if (keep_coefs_) { coefs_.insert(coefs_.end(), coef.begin(), coef.end()); std::fill_n(coef.begin(), (size_t)count * 8, 0u); }
    }
    void flush_linear() {                                                  // process_reset_linear_row
        if (!lin_count_) return;
        linear_total_ += lin_count_;
        kinds_.push_back(0);
        emit(lin_val_, lin_coef_, lin_count_);
        lin_count_ = 0;
    }
    void flush_quadratic() {                                               // process_reset_quadratic_rows
        if (!quad_count_) return;
        quadratic_total_ += quad_count_;
        kinds_.push_back(1);
        for (int i = 0; i < 3; i++) emit(quad_val_[i], quad_coef_[i], quad_count_);
        quad_count_ = 0;
    }

    uint32_t l_;
    bool keep_values_, keep_coefs_;
    uint32_t lin_count_ = 0, quad_count_ = 0;
    uint64_t linear_total_ = 0, quadratic_total_ = 0;
    std::vector<uint32_t> lin_val_, lin_coef_;
    std::array<std::vector<uint32_t>, 3> quad_val_, quad_coef_;
    std::vector<uint8_t> kinds_;
    std::vector<uint32_t> values_, coefs_;
};

}  // namespace ligero::cuda::host

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
#include <array>
#include <cstdint>
#include <cstring>
#include <vector>

namespace ligero::cuda::host {

class row_packer {
public:
    explicit row_packer(uint32_t l) : l_(l) {}

    // commit_status::linear_ready
    void push_linear(const uint32_t value[8], const uint32_t coef[8]) {
        if (lin_count_ >= l_) flush_linear();
        append(lin_val_, value); append(lin_coef_, coef);
        lin_count_++;
    }
    // commit_status::quadratic_ready: one slot = (x, y, z) with x*y = z, plus their linear-test coefficients
    void push_quadratic(const uint32_t x[8], const uint32_t y[8], const uint32_t z[8], const uint32_t cx[8], const uint32_t cy[8], const uint32_t cz[8]) {
        if (quad_count_ >= l_) flush_quadratic();
        append(quad_val_[0], x); append(quad_val_[1], y); append(quad_val_[2], z);
        append(quad_coef_[0], cx); append(quad_coef_[1], cy); append(quad_coef_[2], cz);
        quad_count_++;
    }
    // witness_manager::finalize (the mask rows are the prover's business)
    void finalize() { flush_linear(); flush_quadratic(); }

    uint32_t l() const { return l_; }
    const std::vector<uint8_t> &kinds() const { return kinds_; }          // per event: 0 linear row, 1 quadratic triple
    const std::vector<uint32_t> &values() const { return values_; }       // encoded rows in emission order: [rows][l][8]
    const std::vector<uint32_t> &coefs() const { return coefs_; }
    size_t rows() const { return values_.size() / ((size_t)l_ * 8); }
    uint64_t linear_count() const { return linear_total_; }               // "Num Linear constraints" / "Num quadratic constraints"
    uint64_t quadratic_count() const { return quadratic_total_; }

private:
    static void append(std::vector<uint32_t> &v, const uint32_t x[8]) { v.insert(v.end(), x, x + 8); }
    void emit(std::vector<uint32_t> &val, std::vector<uint32_t> &coef) {
        val.resize((size_t)l_ * 8, 0); coef.resize((size_t)l_ * 8, 0);    // push_back_zeros(row_size - data_size)
        values_.insert(values_.end(), val.begin(), val.end());
        coefs_.insert(coefs_.end(), coef.begin(), coef.end());
        val.clear(); coef.clear();
    }
    void flush_linear() {                                                  // process_reset_linear_row
        if (!lin_count_) return;
        linear_total_ += lin_count_;
        kinds_.push_back(0);
        emit(lin_val_, lin_coef_);
        lin_count_ = 0;
    }
    void flush_quadratic() {                                               // process_reset_quadratic_rows
        if (!quad_count_) return;
        quadratic_total_ += quad_count_;
        kinds_.push_back(1);
        for (int i = 0; i < 3; i++) emit(quad_val_[i], quad_coef_[i]);
        quad_count_ = 0;
    }

    uint32_t l_;
    uint32_t lin_count_ = 0, quad_count_ = 0;
    uint64_t linear_total_ = 0, quadratic_total_ = 0;
    std::vector<uint32_t> lin_val_, lin_coef_;
    std::array<std::vector<uint32_t>, 3> quad_val_, quad_coef_;
    std::vector<uint8_t> kinds_;
    std::vector<uint32_t> values_, coefs_;
};

}  // namespace ligero::cuda::host

// gsb_sparse_matrix.hpp -- drop-in replacement for the reference's header-only container (included by sparse-matrix.h)
//   labs/lab3/src/OpenCVHW1/sparse-matrix.h                       ("v1")
//   labs/lab8/src/OpenCVHW1/sparse-matrix.h == project/src/PhotoMontage/sparse-matrix.h   ("v2")
// Same class name, template parameters, public members and default arguments; the method bodies
// call the C ABI of libgsb200.so (include/gsb200.h), which runs them on a B200.
//
//   bulk work  -> device:  initializeFromVector, initialize(r,c,list), initializeFromTriplets,
//                          initializeFromEigenRowMajor, gaussSeidel, applyToVector, conjugateGradient*,
//                          manhattonDist/dotProd/veclen2/vecadd/vecsub/vecmul
//   single-element access/modify (at, coeff, insert*) -> host copy of the five layout arrays, exactly
//   the reference's data structure; an insert marks the device copy stale and the next solver
//   call re-uploads it (gsb_matrix_upload).
//
// Behavioural notes:
//   * gaussSeidel sweeps in red-black / multicolour order instead of lexicographic order: the
//     iterates differ, the fixed point and the stop rule do not.
//   * insertNoneZero keeps rows sorted and insertZero moves sizeof(Index) bytes per column index;
//     the reference's versions of both corrupt the row for some inputs (v2 :198, :207-221).
//   * initialize(r,c) sizes row_num_nze_ so that initializeFromTriplets works (upstream it crashes).
//   * errors: the reference asserts only under _DEBUG.  Here a failing device call throws
//     std::runtime_error carrying gsb_last_error(); there is no CPU fallback.
//
// T must be int or double (the reference's two instantiations); IndexType must be int.
#pragma once

#include <algorithm>
#include <cstring>
#include <initializer_list>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "gsb200.h"

#ifdef USE_NAME_SPACE
namespace USE_NAME_SPACE {
#endif

namespace gsb_detail {
inline void check(int status, const char *where) {
    if (status != GSB_OK)
        throw std::runtime_error(std::string(where) + ": " + gsb_last_error() + " (status " + std::to_string(status) + ")");
}
}  // namespace gsb_detail

// ---- free vector helpers (v2 :45-105) ------------------------------------------------------------
inline double manhattonDist(const std::vector<double> &a, const std::vector<double> &b) {
    double out = 0;
    gsb_detail::check(gsb_l1_dist(a.data(), b.data(), (int64_t)a.size(), &out), "manhattonDist");
    return out;
}
inline double dotProd(const std::vector<double> &a, const std::vector<double> &b) {
    double out = 0;
    gsb_detail::check(gsb_dot(a.data(), b.data(), (int64_t)a.size(), &out), "dotProd");
    return out;
}
inline double veclen2(const std::vector<double> &a) { return dotProd(a, a); }
inline void vecadd(const std::vector<double> &a, const std::vector<double> &b, double scale_b, std::vector<double> &out) {
    gsb_detail::check(gsb_axpy(a.data(), b.data(), scale_b, (int64_t)a.size(), out.data()), "vecadd");
}
inline void vecsub(const std::vector<double> &a, const std::vector<double> &b, std::vector<double> &out) {
    vecadd(a, b, -1.0, out);
}
inline void vecmul(const std::vector<double> &a, const std::vector<double> &b, std::vector<double> &out) {
    gsb_detail::check(gsb_vecmul(a.data(), b.data(), (int64_t)a.size(), out.data()), "vecmul");
}
inline void vecadd(const std::vector<double> &src, const double inc, std::vector<double> &out) {
    std::transform(src.begin(), src.end(), out.begin(), [inc](double a) { return a + inc; });
}
inline void vecmul(const std::vector<double> &src, const double scale, std::vector<double> &out) {
    std::transform(src.begin(), src.end(), out.begin(), [scale](double a) { return a * scale; });
}

template <typename T, typename IndexType = int>
class SparseMatrix {
    static_assert(std::is_same<T, int>::value || std::is_same<T, double>::value,
                  "gs-b200 SparseMatrix: T must be int or double (the reference's instantiations)");
    static_assert(std::is_same<IndexType, int>::value, "gs-b200 SparseMatrix: IndexType must be int");
    using Vector = std::vector<T>;

public:
    using Index = IndexType;
    using IndexVector = std::vector<Index>;

    SparseMatrix() = default;
    SparseMatrix(SparseMatrix &&o) noexcept { *this = std::move(o); }
    SparseMatrix &operator=(SparseMatrix &&o) noexcept {
        if (this != &o) {
            release();
            values_ = std::move(o.values_);
            col_offset_ = std::move(o.col_offset_);
            row_begin_ = std::move(o.row_begin_);
            row_num_nze_ = std::move(o.row_num_nze_);
            row_space_left_ = std::move(o.row_space_left_);
            n_rows_ = o.n_rows_;
            n_cols_ = o.n_cols_;
            handle_ = o.handle_;
            device_stale_ = o.device_stale_;
            o.handle_ = nullptr;
        }
        return *this;
    }
    SparseMatrix(const SparseMatrix &) = delete;
    SparseMatrix &operator=(const SparseMatrix &) = delete;
    ~SparseMatrix() { release(); }

    Index cols() const { return n_cols_; }
    Index rows() const { return n_rows_; }

    // v2 :162-173
    T at(Index row, Index col) const {
        if (!row_num_nze_[row]) return T(0);
        Index idx = getNearestIndex(row, col);
        if (col_offset_[idx] == col) return values_[idx];
        return T(0);
    }
    T coeff(Index row, Index col) const { return at(row, col); }  // v2 :176

    // v2 :183-201
    void insertZero(Index row, Index col) {
        if (row_num_nze_[row] == 0) return;
        Index idx = getNearestIndex(row, col);
        if (col_offset_[idx] == col) {
            Index end = row_begin_[row] + row_num_nze_[row];
            std::memmove(values_.data() + idx, values_.data() + idx + 1, sizeof(T) * (size_t)(end - idx - 1));
            std::memmove(col_offset_.data() + idx, col_offset_.data() + idx + 1, sizeof(Index) * (size_t)(end - idx - 1));
            --row_num_nze_[row];
            ++row_space_left_[row];
            device_stale_ = true;
        }
    }

    // v2 :203-237
    void insertNoneZero(T &&val, Index row, Index col) {
        Index idx = row_begin_[row];
        if (row_num_nze_[row]) {
            idx = getNearestIndex(row, col);
            if (col_offset_[idx] == col) {
                values_[idx] = std::forward<T>(val);
                device_stale_ = true;
                return;
            }
            if (col_offset_[idx] < col) ++idx;  // col lies beyond every live column of the row
        }
        Index end = row_begin_[row] + row_num_nze_[row];
        if (row_space_left_[row]) {
            --row_space_left_[row];
            std::memmove(values_.data() + idx + 1, values_.data() + idx, sizeof(T) * (size_t)(end - idx));
            std::memmove(col_offset_.data() + idx + 1, col_offset_.data() + idx, sizeof(Index) * (size_t)(end - idx));
            values_[idx] = std::forward<T>(val);
            col_offset_[idx] = col;
        } else {
            values_.insert(values_.begin() + idx, val);
            col_offset_.insert(col_offset_.begin() + idx, col);
            auto sz = static_cast<Index>(row_begin_.size());
            for (Index i = row + 1; i < sz; ++i) ++row_begin_[i];
        }
        ++row_num_nze_[row];
        device_stale_ = true;
    }

    void insert(const T &val, Index row, Index col) { insert(T(val), row, col); }
    void insert(T &&val, Index row, Index col) {
        return val == T(0) ? insertZero(row, col) : insertNoneZero(std::forward<T>(val), row, col);
    }

    struct Triplet {
        Index row;
        Index col;
        T val;
    };

    // v2 :256-263: the insert() loop, done as one device sort (last duplicate wins, zeros dropped)
    void initializeFromTriplets(Triplet *a, Index cnt) {
        std::vector<Index> r((size_t)cnt), c((size_t)cnt);
        Vector v((size_t)cnt);
        for (Index i = 0; i < cnt; ++i) {
            r[i] = a[i].row;
            c[i] = a[i].col;
            v[i] = a[i].val;
        }
        ensure_handle();
        gsb_detail::check(gsb_matrix_assemble_coo(handle_, r.data(), c.data(), v.data(), cnt, n_rows_, n_cols_),
                          "initializeFromTriplets");
        pull();
    }

    // v2 :265-319
    void initializeFromVector(const IndexVector &rows, IndexVector &&cols, Vector &&vals) {
        IndexVector c = std::forward<IndexVector>(cols);
        Vector v = std::forward<Vector>(vals);
        ensure_handle();
        gsb_detail::check(gsb_matrix_assemble_sorted_coo(handle_, rows.data(), c.data(), v.data(), (int64_t)rows.size()),
                          "initializeFromVector");
        pull();
    }

    // v2 :321-330
    void initialize(int row, int col) {
        n_rows_ = row;
        n_cols_ = col;
        values_.clear();
        col_offset_.clear();
        row_begin_.assign((size_t)n_rows_, 0);
        row_num_nze_.assign((size_t)n_rows_, 0);
        row_space_left_.assign((size_t)n_rows_, 0);
        device_stale_ = true;
    }

    // v2 :332-347
    void initialize(int row, int col, std::initializer_list<T> x) {
        std::vector<T> v(x);
        std::vector<Index> rows(x.size()), cols(x.size());
        int cnt = 0;
        for (int i = 0; i < row; ++i)
            for (int j = 0; j < col; ++j) {
                rows[cnt] = i;
                cols[cnt] = j;
                ++cnt;
            }
        initializeFromVector(rows, std::move(cols), std::move(v));
    }

    // v2 :350-380 (v1 :275-305 takes b by value; both call sites compile against this signature)
    std::vector<double> gaussSeidel(const std::vector<double> &b, double epsilon = 1e-6, int max_iteration = 1000) {
        push();
        std::vector<double> x(b.size(), 1.0);
        gsb_detail::check(gsb_gauss_seidel(handle_, b.data(), 1, epsilon, max_iteration, nullptr, x.data(), &last_stats),
                          "gaussSeidel");
        return x;
    }
    // EXTENSION: up to 4 right-hand sides (colour channels) sharing one pass over the matrix per sweep.
    // b = nrhs vectors of rows() doubles, one after another; same layout for the result.
    std::vector<double> gaussSeidelMulti(const std::vector<double> &b, int nrhs, double epsilon = 1e-6,
                                         int max_iteration = 1000, const gsb_gs_options *opts = nullptr) {
        push();
        std::vector<double> x(b.size(), 1.0);
        gsb_detail::check(gsb_gauss_seidel(handle_, b.data(), nrhs, epsilon, max_iteration, opts, x.data(), &last_stats),
                          "gaussSeidelMulti");
        return x;
    }

    // v2 :382-393
    void applyToVector(const std::vector<double> &in, std::vector<double> &out) {
        push();
        gsb_detail::check(gsb_spmv(handle_, in.data(), out.data()), "applyToVector");
    }

    // v2 :396-434
    std::vector<double> conjugateGradient(const std::vector<double> &b, double epsilon = 1e-16, int max_iteration = 1000,
                                          const std::vector<double> &initialize = std::vector<double>()) {
        push();
        std::vector<double> x(b.size(), 0.0);
        gsb_detail::check(gsb_conjugate_gradient(handle_, b.data(), epsilon, max_iteration,
                                                 initialize.size() ? initialize.data() : nullptr, x.data(), nullptr),
                          "conjugateGradient");
        return x;
    }
    // v2 :436-468 (same recurrence as conjugateGradient without the initial guess)
    std::vector<double> conjugateGradientPaper(const std::vector<double> &b, double epsilon = 1e-16,
                                               int max_iteration = 1000) {
        return conjugateGradient(b, epsilon, max_iteration);
    }
    // v2 :494-535
    std::vector<double> conjugateGradientEigen(const std::vector<double> &b, double epsilon = 1e-16,
                                               int max_iteration = 180) {
        push();
        std::vector<double> x(b.size(), 0.0);
        gsb_detail::check(gsb_conjugate_gradient_jacobi(handle_, b.data(), epsilon, max_iteration, x.data(), nullptr),
                          "conjugateGradientEigen");
        return x;
    }
    // v2 :472-491
    std::vector<T> extractDiagnolColInv() {
        std::vector<T> res((size_t)cols(), T(1));
        for (Index i = 0; i < n_rows_; ++i) {
            Index idx = row_begin_[i];
            for (Index j = 0; j < row_num_nze_[i]; ++j, ++idx)
                if (col_offset_[idx] == i) {
                    if (values_[idx] != 0) res[i] = T(1) / values_[idx];
                    break;
                }
        }
        return res;
    }

    // v2 :537-620
    void initializeFromEigenRowMajor(const T *values, Index n_values, const Index *row_offset, Index n_row_offset,
                                     const Index *col_offset, Index n_col_offset, const Index *non_zeros,
                                     Index n_non_zeros) {
        ensure_handle();
        gsb_detail::check(gsb_matrix_import_csr(handle_, values, n_values, row_offset, n_row_offset, col_offset,
                                                n_col_offset, non_zeros, n_non_zeros),
                          "initializeFromEigenRowMajor");
        pull();
    }

    // diagnostics of the last gaussSeidel call (sweeps, last_eps, colours, device ms)
    gsb_gs_stats last_stats{};
    gsb_matrix *device_handle() {
        push();
        return handle_;
    }

private:
    // v2 :627-645
    inline Index getNearestIndex(Index row, Index col) const {
        Index idx = row_begin_[row];
        Index end = row_begin_[row] + row_num_nze_[row] - 1;
        if (col_offset_[idx] == col) return idx;
        while (end > idx) {
            Index mid = (end + idx) / 2;
            if (col_offset_[mid] < col)
                idx = mid + 1;
            else
                end = mid;
        }
        return idx;
    }

    void ensure_handle() {
        if (!handle_)
            gsb_detail::check(gsb_matrix_create(&handle_, std::is_same<T, int>::value ? GSB_I32 : GSB_F64),
                              "SparseMatrix (gsb_matrix_create)");
    }
    void release() {
        if (handle_) gsb_matrix_destroy(handle_);
        handle_ = nullptr;
    }
    // device -> host copy of the five arrays after a bulk build
    void pull() {
        int64_t store = 0, nnz = 0;
        int nr = 0, nc = 0;
        gsb_detail::check(gsb_matrix_shape(handle_, &store, &nr, &nc, &nnz), "gsb_matrix_shape");
        n_rows_ = nr;
        n_cols_ = nc;
        values_.resize((size_t)store);
        col_offset_.resize((size_t)store);
        row_begin_.resize((size_t)nr);
        row_num_nze_.resize((size_t)nr);
        row_space_left_.resize((size_t)nr);
        gsb_detail::check(gsb_matrix_download(handle_, values_.data(), col_offset_.data(), row_begin_.data(),
                                              row_num_nze_.data(), row_space_left_.data()),
                          "gsb_matrix_download");
        device_stale_ = false;
    }
    // host -> device after insert() edits
    void push() {
        ensure_handle();
        if (!device_stale_) return;
        gsb_detail::check(gsb_matrix_upload(handle_, values_.data(), col_offset_.data(), (int64_t)values_.size(),
                                            row_begin_.data(), row_num_nze_.data(), row_space_left_.data(), n_rows_,
                                            n_cols_),
                          "gsb_matrix_upload");
        device_stale_ = false;
    }

    Vector values_;
    IndexVector col_offset_;
    IndexVector row_begin_;
    IndexVector row_num_nze_;  // number of non zero elements in row
    IndexVector row_space_left_;
    Index n_rows_ = 0;
    Index n_cols_ = 0;
    gsb_matrix *handle_ = nullptr;
    bool device_stale_ = false;
};

#ifdef USE_NAME_SPACE
}
#endif

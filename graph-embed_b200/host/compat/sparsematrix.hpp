// graph-embed_b200 :: linalgcpp compatibility surface (CSR container only).
//
// The reference depends on github.com/gelever/linalgcpp (find_package at
// /root/reference/CMakeLists.txt:26, un-vendored, unpinned).  The ForceAtlas hot
// path uses exactly six accessors of linalgcpp::SparseMatrix<double> plus
// Transpose() (/root/reference/include/forceatlas.hpp:112-116, 342-346;
// /root/reference/src/embed.cpp:605, 684).  This header provides that surface so
// that (a) callers without linalgcpp can use the drop-in `partition::embed`, and
// (b) the unmodified reference sources can be compiled as the parity oracle
// (oracle/Makefile).  If the real linalgcpp is on the include path, use it
// instead: the drop-in only touches the accessors listed above.
#ifndef GE_B200_COMPAT_SPARSEMATRIX_HPP
#define GE_B200_COMPAT_SPARSEMATRIX_HPP

#include <algorithm>
#include <cassert>
#include <chrono>
#include <cmath>
#include <functional>
#include <iomanip>
#include <iostream>
#include <limits>
#include <map>
#include <numeric>
#include <random>
#include <tuple>
#include <utility>
#include <vector>

namespace linalgcpp {

template <typename T = double>
class SparseMatrix {
 public:
  SparseMatrix() : rows_(0), cols_(0), indptr_(1, 0) {}

  SparseMatrix(std::vector<int> indptr, std::vector<int> indices, std::vector<T> data,
               int rows, int cols)
      : rows_(rows), cols_(cols), indptr_(std::move(indptr)),
        indices_(std::move(indices)), data_(std::move(data)) {
    assert(static_cast<int>(indptr_.size()) == rows_ + 1);
    assert(indices_.size() == data_.size());
  }

  // Square diagonal matrix from its diagonal.
  explicit SparseMatrix(std::vector<T> diag)
      : rows_(static_cast<int>(diag.size())), cols_(static_cast<int>(diag.size())),
        indptr_(diag.size() + 1), indices_(diag.size()), data_(std::move(diag)) {
    std::iota(indptr_.begin(), indptr_.end(), 0);
    std::iota(indices_.begin(), indices_.end(), 0);
  }

  int Rows() const { return rows_; }
  int Cols() const { return cols_; }
  int nnz() const { return static_cast<int>(data_.size()); }

  const std::vector<int>& GetIndptr() const { return indptr_; }
  const std::vector<int>& GetIndices() const { return indices_; }
  const std::vector<T>& GetData() const { return data_; }
  std::vector<int>& GetIndptr() { return indptr_; }
  std::vector<int>& GetIndices() { return indices_; }
  std::vector<T>& GetData() { return data_; }

  // Counting-sort transpose: rows of the result have ascending column ids.
  SparseMatrix<T> Transpose() const {
    std::vector<int> tptr(cols_ + 1, 0);
    for (int c : indices_) tptr[c + 1]++;
    for (int c = 0; c < cols_; ++c) tptr[c + 1] += tptr[c];
    std::vector<int> fill(tptr.begin(), tptr.end() - 1);
    std::vector<int> tind(indices_.size());
    std::vector<T> tdat(data_.size());
    for (int r = 0; r < rows_; ++r) {
      for (int k = indptr_[r]; k < indptr_[r + 1]; ++k) {
        int dst = fill[indices_[k]]++;
        tind[dst] = r;
        tdat[dst] = data_[k];
      }
    }
    return SparseMatrix<T>(std::move(tptr), std::move(tind), std::move(tdat), cols_, rows_);
  }

  // Row-wise Gustavson product; output rows sorted by column, duplicates summed.
  SparseMatrix<T> Mult(const SparseMatrix<T>& rhs) const {
    assert(cols_ == rhs.rows_);
    std::vector<int> optr(rows_ + 1, 0), oind;
    std::vector<T> odat;
    std::vector<int> marker(rhs.cols_, -1);
    std::vector<int> touched;
    std::vector<T> acc(rhs.cols_, T(0));
    for (int r = 0; r < rows_; ++r) {
      touched.clear();
      for (int k = indptr_[r]; k < indptr_[r + 1]; ++k) {
        const int mid = indices_[k];
        const T lhs_val = data_[k];
        for (int k2 = rhs.indptr_[mid]; k2 < rhs.indptr_[mid + 1]; ++k2) {
          const int c = rhs.indices_[k2];
          if (marker[c] != r) {
            marker[c] = r;
            acc[c] = T(0);
            touched.push_back(c);
          }
          acc[c] += lhs_val * rhs.data_[k2];
        }
      }
      std::sort(touched.begin(), touched.end());
      for (int c : touched) {
        oind.push_back(c);
        odat.push_back(acc[c]);
      }
      optr[r + 1] = static_cast<int>(oind.size());
    }
    return SparseMatrix<T>(std::move(optr), std::move(oind), std::move(odat), rows_, rhs.cols_);
  }

  void ScaleRows(const std::vector<T>& v) {
    for (int r = 0; r < rows_; ++r)
      for (int k = indptr_[r]; k < indptr_[r + 1]; ++k) data_[k] *= v[r];
  }
  void ScaleCols(const std::vector<T>& v) {
    for (size_t k = 0; k < data_.size(); ++k) data_[k] *= v[indices_[k]];
  }

 private:
  int rows_, cols_;
  std::vector<int> indptr_, indices_;
  std::vector<T> data_;
};

// Coordinate-format builder; duplicates are summed, rows come out column-sorted.
template <typename T = double>
class CooMatrix {
 public:
  CooMatrix() : rows_(0), cols_(0) {}
  CooMatrix(int rows, int cols) : rows_(rows), cols_(cols) {}
  void Add(int i, int j, T val) { entries_[std::make_pair(i, j)] += val; }
  SparseMatrix<T> ToSparse() const {
    std::vector<int> ptr(rows_ + 1, 0), ind;
    std::vector<T> dat;
    ind.reserve(entries_.size());
    dat.reserve(entries_.size());
    for (const auto& e : entries_) {
      ptr[e.first.first + 1]++;
      ind.push_back(e.first.second);
      dat.push_back(e.second);
    }
    for (int r = 0; r < rows_; ++r) ptr[r + 1] += ptr[r];
    return SparseMatrix<T>(std::move(ptr), std::move(ind), std::move(dat), rows_, cols_);
  }

 private:
  int rows_, cols_;
  std::map<std::pair<int, int>, T> entries_;
};

// Wall-clock stopwatch: operator[](i) = seconds between click i and click i+1.
class Timer {
 public:
  enum class Start { True, False };
  explicit Timer(Start start = Start::False) {
    if (start == Start::True) Click();
  }
  void Click() { marks_.push_back(std::chrono::steady_clock::now()); }
  double operator[](int i) const {
    return std::chrono::duration<double>(marks_[i + 1] - marks_[i]).count();
  }

 private:
  std::vector<std::chrono::steady_clock::time_point> marks_;
};

}  // namespace linalgcpp

#endif  // GE_B200_COMPAT_SPARSEMATRIX_HPP

// TEST INFRASTRUCTURE ONLY -- facade for the Eigen 3 classes the reference calls (Eigen is a third-party,
// header-only dependency that is absent from /root/reference and from this image; makefile:11 takes the system
// package, version unpinned).  It exists so that the reference's OWN dynrec.cpp, ftsolver.cpp, periodic.cpp and
// player.cpp compile unmodified into oracle/_ref (oracle/Makefile, target `ref`).
//
// Everything is evaluated eagerly on one dense column-major type (MatrixXd == VectorXd == Eigen::Dense); sparse
// matrices are dense underneath.  The factorisations are the ones of oracle/orc_linalg.hpp:
//   SparseQR<SpMat,COLAMDOrdering<int>>   dense Householder QR (column-pivoted when the matrix has zero columns that
//                                         are not trailing or turns out rank deficient); ftsolver.cpp:110-129,351-353
//   FullPivLU<MatrixXd>                   restatement of Eigen's algorithm incl. its default threshold; :210-217
//   ColPivHouseholderQR / colPivHouseholderQr()   restatement of Eigen's algorithm; :227, player.cpp:537
// The quantities the reference takes from the QRs (the unique solution of a square nonsingular system, an orthonormal
// basis of a null space whose span is all that matters downstream, a least-squares solution of a full-column-rank
// system) do not depend on the factorisation used; this back end is cross-checked against LAPACK in
// tests/test_oracle_lapack.py.  Eigen behaviours the reference leans on are kept: FullPivLU::kernel() of an
// invertible matrix is ONE zero column, the comma initialiser asserts when over- or under-filled, insert() returns
// a reference to a zero-initialised coefficient.
#ifndef ORACLE_SHIM_EIGEN_HPP
#define ORACLE_SHIM_EIGEN_HPP

// Standard headers the real Eigen/Core pulls in; the reference relies on getting std::stringstream, std::copy
// and friends through it (player.cpp:171,416, cpc.cpp:391 include no <sstream> / <algorithm> of their own).
#include <algorithm>
#include <cassert>
#include <climits>
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iostream>
#include <limits>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#if __cplusplus >= 201103L
#define ORACLE_SHIM_DTOR_MAY_ASSERT noexcept(false)
#else
#define ORACLE_SHIM_DTOR_MAY_ASSERT
#endif

#include "../../orc_linalg.hpp"

namespace Eigen {

const int Dynamic = -1;

class Dense;
class BlockRef;

class ArrayProxy {
 public:
  explicit ArrayProxy(const Dense& m) : m_(m) {}
  Dense square() const;
  Dense abs() const;

 private:
  const Dense& m_;
};

class CommaInit {
 public:
  CommaInit(Dense& m, const Dense& first);
  CommaInit(Dense& m, double first);
  ~CommaInit() ORACLE_SHIM_DTOR_MAY_ASSERT;
  CommaInit& operator,(const Dense& b);
  CommaInit& operator,(double v);

 private:
  void place(const Dense& b);
  Dense& m_;
  int row_, col_, block_rows_;
};

class MapRef {  // VectorXd::Map(ptr, n) = expression
 public:
  MapRef(double* p, int n) : p_(p), n_(n) {}
  MapRef& operator=(const Dense& v);
  operator Dense() const;

 private:
  double* p_;
  int n_;
};

class Dense {
 public:
  Dense() : r_(0), c_(0) {}
  explicit Dense(int n) : r_(n), c_(1), d_((size_t)n, 0.0) {}
  Dense(int r, int c) : r_(r), c_(c), d_((size_t)r * c, 0.0) {}
  Dense(const BlockRef& b);  // NOLINT: implicit, like evaluating an Eigen block expression

  int rows() const { return r_; }
  int cols() const { return c_; }
  int size() const { return r_ * c_; }
  double& operator()(int i) { return d_[(size_t)i]; }
  double operator()(int i) const { return d_[(size_t)i]; }
  double& operator()(int i, int j) { return d_[(size_t)j * r_ + i]; }
  double operator()(int i, int j) const { return d_[(size_t)j * r_ + i]; }
  double& operator[](int i) { return d_[(size_t)i]; }
  double operator[](int i) const { return d_[(size_t)i]; }
  double* data() { return d_.empty() ? 0 : &d_[0]; }
  const double* data() const { return d_.empty() ? 0 : &d_[0]; }

  void resize(int n) { r_ = n; c_ = 1; d_.assign((size_t)n, 0.0); }
  void resize(int r, int c) { r_ = r; c_ = c; d_.assign((size_t)r * c, 0.0); }
  void conservativeResize(int n) { conservativeResize(n, 1); }
  void conservativeResize(int r, int c) {
    Dense t(r, c);
    for (int j = 0; j < c && j < c_; j++)
      for (int i = 0; i < r && i < r_; i++) t(i, j) = (*this)(i, j);
    *this = t;
  }
  Dense& setZero() { for (size_t i = 0; i < d_.size(); i++) d_[i] = 0; return *this; }
  Dense& setOnes() { for (size_t i = 0; i < d_.size(); i++) d_[i] = 1; return *this; }
  Dense& setConstant(double v) { for (size_t i = 0; i < d_.size(); i++) d_[i] = v; return *this; }
  static Dense Constant(int n, double v) { Dense m(n); m.setConstant(v); return m; }
  static Dense Constant(int r, int c, double v) { Dense m(r, c); m.setConstant(v); return m; }
  static Dense Zero(int n) { return Dense(n); }
  static Dense Zero(int r, int c) { return Dense(r, c); }
  static Dense Identity(int r, int c) { Dense m(r, c); for (int i = 0; i < r && i < c; i++) m(i, i) = 1; return m; }
  static MapRef Map(double* p, int n) { return MapRef(p, n); }

  Dense transpose() const {
    Dense t(c_, r_);
    for (int j = 0; j < c_; j++) for (int i = 0; i < r_; i++) t(j, i) = (*this)(i, j);
    return t;
  }
  double norm() const { double s = 0; for (size_t i = 0; i < d_.size(); i++) s += d_[i] * d_[i]; return std::sqrt(s); }
  double squaredNorm() const { double s = 0; for (size_t i = 0; i < d_.size(); i++) s += d_[i] * d_[i]; return s; }
  template <int P>
  double lpNorm() const {  // P = 1 (periodic.cpp:231) or 2
    double s = 0;
    for (size_t i = 0; i < d_.size(); i++) s += (P == 1) ? std::fabs(d_[i]) : d_[i] * d_[i];
    return P == 1 ? s : std::sqrt(s);
  }
  double sum() const { double s = 0; for (size_t i = 0; i < d_.size(); i++) s += d_[i]; return s; }
  double dot(const Dense& o) const { double s = 0; for (size_t i = 0; i < d_.size(); i++) s += d_[i] * o.d_[i]; return s; }
  ArrayProxy array() const { return ArrayProxy(*this); }
  Dense cwiseAbs() const { Dense t(*this); for (size_t i = 0; i < d_.size(); i++) t.d_[i] = std::fabs(d_[i]); return t; }
  double minCoeff() const { double m = d_[0]; for (size_t i = 1; i < d_.size(); i++) if (d_[i] < m) m = d_[i]; return m; }
  double maxCoeff() const { double m = d_[0]; for (size_t i = 1; i < d_.size(); i++) if (d_[i] > m) m = d_[i]; return m; }
  Dense inverse() const;
  double determinant() const;
  Dense eigenvalues() const;  // symmetric matrices only (cpc.cpp:207 applies it to b b^T-like Gram matrices)
  class SparseShim sparseView(double = 0, double = 0) const;

  BlockRef block(int i, int j, int r, int c);
  const BlockRef block(int i, int j, int r, int c) const;
  BlockRef col(int j);
  const BlockRef col(int j) const;
  BlockRef row(int i);
  const BlockRef row(int i) const;
  BlockRef segment(int i, int n);
  const BlockRef segment(int i, int n) const;
  BlockRef head(int n);
  const BlockRef head(int n) const;
  BlockRef tail(int n);
  const BlockRef tail(int n) const;
  BlockRef topRows(int n);
  BlockRef bottomRows(int n);
  BlockRef leftCols(int n);
  BlockRef rightCols(int n);
  const BlockRef rightCols(int n) const;

  Dense& operator+=(const Dense& o) { same(o); for (size_t i = 0; i < d_.size(); i++) d_[i] += o.d_[i]; return *this; }
  Dense& operator-=(const Dense& o) { same(o); for (size_t i = 0; i < d_.size(); i++) d_[i] -= o.d_[i]; return *this; }
  Dense& operator*=(double s) { for (size_t i = 0; i < d_.size(); i++) d_[i] *= s; return *this; }
  Dense& operator/=(double s) { for (size_t i = 0; i < d_.size(); i++) d_[i] /= s; return *this; }
  Dense operator-() const { Dense t(*this); for (size_t i = 0; i < d_.size(); i++) t.d_[i] = -d_[i]; return t; }
  CommaInit operator<<(const Dense& first) { return CommaInit(*this, first); }
  CommaInit operator<<(double first) { return CommaInit(*this, first); }
  class ColPivHouseholderQRShim colPivHouseholderQr() const;

  orc::Mat to_orc() const {
    orc::Mat m(r_, c_);
    m.d = d_;
    return m;
  }
  static Dense from_orc(const orc::Mat& m) {
    Dense t(m.r, m.c);
    t.d_ = m.d;
    return t;
  }
  static Dense from_vec(const orc::Vec& v) {
    Dense t((int)v.size());
    t.d_ = v;
    return t;
  }
  const std::vector<double>& vec() const { return d_; }

 private:
  void same(const Dense& o) const { assert(r_ == o.r_ && c_ == o.c_); (void)o; }
  int r_, c_;
  std::vector<double> d_;
};

typedef Dense MatrixXd;
typedef Dense VectorXd;
typedef Dense RowVectorXd;

class BlockRef {  // writable view of a rectangular part of a Dense
 public:
  BlockRef(Dense& m, int i0, int j0, int r, int c) : m_(m), i0_(i0), j0_(j0), r_(r), c_(c) {
    assert(i0 >= 0 && j0 >= 0 && r >= 0 && c >= 0 && i0 + r <= m.rows() && j0 + c <= m.cols());
  }
  int rows() const { return r_; }
  int cols() const { return c_; }
  int size() const { return r_ * c_; }
  double& operator()(int i, int j) { return m_(i0_ + i, j0_ + j); }
  double operator()(int i, int j) const { return m_(i0_ + i, j0_ + j); }
  double& operator()(int i) { return r_ == 1 ? m_(i0_, j0_ + i) : m_(i0_ + i, j0_); }
  double operator()(int i) const { return r_ == 1 ? m_(i0_, j0_ + i) : m_(i0_ + i, j0_); }
  BlockRef& operator=(const Dense& v) {
    if (v.rows() == r_ && v.cols() == c_) {
      for (int j = 0; j < c_; j++) for (int i = 0; i < r_; i++) (*this)(i, j) = v(i, j);
    } else {  // a vector assigned to a row (B.row(i) = decomp.solve(...), player.cpp:541)
      assert(v.size() == size() && (v.rows() == 1 || v.cols() == 1) && (r_ == 1 || c_ == 1));
      for (int i = 0; i < size(); i++) (*this)(i) = v(i);
    }
    return *this;
  }
  BlockRef& operator=(const BlockRef& v) { return *this = Dense(v); }
  BlockRef& operator*=(double s) { for (int j = 0; j < c_; j++) for (int i = 0; i < r_; i++) (*this)(i, j) *= s; return *this; }
  BlockRef& operator+=(const Dense& v) { for (int j = 0; j < c_; j++) for (int i = 0; i < r_; i++) (*this)(i, j) += v(i, j); return *this; }
  BlockRef& operator-=(const Dense& v) { for (int j = 0; j < c_; j++) for (int i = 0; i < r_; i++) (*this)(i, j) -= v(i, j); return *this; }
  BlockRef& setZero() { return fill(0.0); }
  BlockRef& setOnes() { return fill(1.0); }
  BlockRef& setConstant(double v) { return fill(v); }
  Dense transpose() const { return Dense(*this).transpose(); }
  double norm() const { return Dense(*this).norm(); }
  BlockRef col(int j) { return BlockRef(m_, i0_, j0_ + j, r_, 1); }
  BlockRef row(int i) { return BlockRef(m_, i0_ + i, j0_, 1, c_); }

 private:
  BlockRef& fill(double v) { for (int j = 0; j < c_; j++) for (int i = 0; i < r_; i++) (*this)(i, j) = v; return *this; }
  Dense& m_;
  int i0_, j0_, r_, c_;
};

inline Dense::Dense(const BlockRef& b) : r_(b.rows()), c_(b.cols()), d_((size_t)b.rows() * b.cols(), 0.0) {
  for (int j = 0; j < c_; j++) for (int i = 0; i < r_; i++) (*this)(i, j) = b(i, j);
}
inline BlockRef Dense::block(int i, int j, int r, int c) { return BlockRef(*this, i, j, r, c); }
inline const BlockRef Dense::block(int i, int j, int r, int c) const { return BlockRef(const_cast<Dense&>(*this), i, j, r, c); }
inline BlockRef Dense::col(int j) { return block(0, j, r_, 1); }
inline const BlockRef Dense::col(int j) const { return block(0, j, r_, 1); }
inline BlockRef Dense::row(int i) { return block(i, 0, 1, c_); }
inline const BlockRef Dense::row(int i) const { return block(i, 0, 1, c_); }
inline BlockRef Dense::segment(int i, int n) { return c_ == 1 ? block(i, 0, n, 1) : block(0, i, 1, n); }
inline const BlockRef Dense::segment(int i, int n) const { return c_ == 1 ? block(i, 0, n, 1) : block(0, i, 1, n); }
inline BlockRef Dense::head(int n) { return segment(0, n); }
inline const BlockRef Dense::head(int n) const { return segment(0, n); }
inline BlockRef Dense::tail(int n) { return segment(size() - n, n); }
inline const BlockRef Dense::tail(int n) const { return segment(size() - n, n); }
inline BlockRef Dense::topRows(int n) { return block(0, 0, n, c_); }
inline BlockRef Dense::bottomRows(int n) { return block(r_ - n, 0, n, c_); }
inline BlockRef Dense::leftCols(int n) { return block(0, 0, r_, n); }
inline BlockRef Dense::rightCols(int n) { return block(0, c_ - n, r_, n); }
inline const BlockRef Dense::rightCols(int n) const { return block(0, c_ - n, r_, n); }

inline MapRef& MapRef::operator=(const Dense& v) {
  assert(v.size() == n_);
  for (int i = 0; i < n_; i++) p_[i] = v(i);
  return *this;
}
inline MapRef::operator Dense() const {
  Dense t(n_);
  for (int i = 0; i < n_; i++) t(i) = p_[i];
  return t;
}

template <class T>
class Map : public Dense {  // Map<VectorXd>(ptr, n) used as a read-only source (periodic.cpp:371, cpc.cpp:68)
 public:
  Map(const double* p, int n) : Dense(n) { for (int i = 0; i < n; i++) (*this)(i) = p[i]; }
  Map(const double* p, int r, int c) : Dense(r, c) { for (int i = 0; i < r * c; i++) (*this)(i) = p[i]; }
};

inline Dense ArrayProxy::square() const { Dense t(m_); for (int i = 0; i < t.size(); i++) t(i) = m_(i) * m_(i); return t; }
inline Dense ArrayProxy::abs() const { return m_.cwiseAbs(); }

inline Dense operator+(const Dense& a, const Dense& b) { Dense t(a); t += b; return t; }
inline Dense operator-(const Dense& a, const Dense& b) { Dense t(a); t -= b; return t; }
inline Dense operator*(const Dense& a, double s) { Dense t(a); t *= s; return t; }
inline Dense operator*(double s, const Dense& a) { Dense t(a); t *= s; return t; }
inline Dense operator/(const Dense& a, double s) { Dense t(a); t /= s; return t; }
inline Dense operator*(const Dense& a, const Dense& b) {
  assert(a.cols() == b.rows());
  return Dense::from_orc(orc::matmul(a.to_orc(), b.to_orc()));
}
inline std::ostream& operator<<(std::ostream& os, const Dense& m) {
  for (int i = 0; i < m.rows(); i++) {
    for (int j = 0; j < m.cols(); j++) os << (j ? " " : "") << m(i, j);
    if (i + 1 < m.rows()) os << "\n";
  }
  return os;
}
// a 1x1 product printed or used as a scalar (cout << x.transpose()*D*x, ftsolver.cpp:71-72) goes through operator<<.

inline CommaInit::CommaInit(Dense& m, const Dense& first) : m_(m), row_(0), col_(0), block_rows_(0) { place(first); }
inline CommaInit::CommaInit(Dense& m, double first) : m_(m), row_(0), col_(0), block_rows_(0) { *this, first; }
inline void CommaInit::place(const Dense& b) {
  if (b.cols() == 0) { if (block_rows_ == 0) block_rows_ = b.rows(); return; }
  if (col_ == m_.cols()) { row_ += block_rows_; col_ = 0; block_rows_ = 0; }
  if (block_rows_ == 0) block_rows_ = b.rows();
  assert(b.rows() == block_rows_ && "comma initializer: block height mismatch");
  assert(row_ + b.rows() <= m_.rows() && col_ + b.cols() <= m_.cols() && "Too many coefficients passed to comma initializer");
  for (int j = 0; j < b.cols(); j++) for (int i = 0; i < b.rows(); i++) m_(row_ + i, col_ + j) = b(i, j);
  col_ += b.cols();
}
inline CommaInit& CommaInit::operator,(const Dense& b) { place(b); return *this; }
inline CommaInit& CommaInit::operator,(double v) { Dense b(1, 1); b(0, 0) = v; place(b); return *this; }
inline CommaInit::~CommaInit() ORACLE_SHIM_DTOR_MAY_ASSERT {
  assert((m_.size() == 0 || (row_ + block_rows_ == m_.rows() && col_ == m_.cols())) &&
         "Too few coefficients passed to comma initializer");
}

template <class T, int N>
class DiagonalMatrix {
 public:
  DiagonalMatrix() {}
  explicit DiagonalMatrix(const Dense& d) : d_(d) {}
  const Dense& diagonal() const { return d_; }

 private:
  Dense d_;
};
template <class T, int N>
inline Dense operator*(const DiagonalMatrix<T, N>& D, const Dense& m) {
  assert(D.diagonal().size() == m.rows());
  Dense t(m);
  for (int j = 0; j < m.cols(); j++) for (int i = 0; i < m.rows(); i++) t(i, j) = D.diagonal()(i) * m(i, j);
  return t;
}
template <class T, int N>
inline Dense operator*(const Dense& m, const DiagonalMatrix<T, N>& D) {
  assert(D.diagonal().size() == m.cols());
  Dense t(m);
  for (int j = 0; j < m.cols(); j++) for (int i = 0; i < m.rows(); i++) t(i, j) = m(i, j) * D.diagonal()(j);
  return t;
}

// ------------------------------------------------------------------ dense factorisations
template <class M>
class FullPivLU {
 public:
  explicit FullPivLU(const Dense& A) : lu_(A.to_orc()) {}
  int rank() const { return lu_.rank(); }
  double threshold() const { return lu_.threshold(); }
  FullPivLU& setThreshold(double t) { lu_.setThreshold(t); return *this; }
  Dense solve(const Dense& b) const { return Dense::from_vec(lu_.solve(b.vec())); }
  Dense kernel() const {
    orc::Mat K = lu_.kernel();
    if (K.c == 0) return Dense(K.r, 1);  // Eigen: the kernel of an invertible matrix is returned as one zero column
    return Dense::from_orc(K);
  }
  Dense image(const Dense& original) const {
    orc::Mat I = lu_.image(original.to_orc());
    if (I.c == 0) return Dense(I.r, 1);
    return Dense::from_orc(I);
  }
  bool isInvertible() const { return lu_.rank() == lu_.lu.r && lu_.lu.r == lu_.lu.c; }

 private:
  orc::FullPivLU lu_;
};

class ColPivHouseholderQRShim {
 public:
  ColPivHouseholderQRShim() {}
  explicit ColPivHouseholderQRShim(const Dense& A) : a_(A) {}
  ColPivHouseholderQRShim& compute(const Dense& A) { a_ = A; return *this; }
  Dense solve(const Dense& b) const {
    Dense x(a_.cols(), b.cols());
    for (int j = 0; j < b.cols(); j++) {
      orc::Vec col(b.rows());
      for (int i = 0; i < b.rows(); i++) col[i] = b(i, j);
      orc::Vec xs = orc::colpiv_qr_solve(a_.to_orc(), col);
      for (int i = 0; i < a_.cols(); i++) x(i, j) = xs[i];
    }
    return x;
  }

 private:
  Dense a_;
};
template <class M>
class ColPivHouseholderQR : public ColPivHouseholderQRShim {
 public:
  ColPivHouseholderQR() {}
  explicit ColPivHouseholderQR(const Dense& A) : ColPivHouseholderQRShim(A) {}
};
inline ColPivHouseholderQRShim Dense::colPivHouseholderQr() const { return ColPivHouseholderQRShim(*this); }

inline Dense Dense::inverse() const {
  assert(r_ == c_);
  orc::FullPivLU lu(to_orc());
  Dense inv(r_, c_);
  for (int j = 0; j < c_; j++) {
    orc::Vec e(r_, 0.0);
    e[j] = 1;
    orc::Vec x = lu.solve(e);
    for (int i = 0; i < r_; i++) inv(i, j) = x[i];
  }
  return inv;
}
inline double Dense::determinant() const {
  assert(r_ == c_);
  orc::FullPivLU lu(to_orc());
  double det = 1;
  for (int i = 0; i < r_; i++) det *= lu.lu(i, i);
  int swaps = 0;  // parity of the two permutations
  std::vector<int> p = lu.p, q = lu.q;
  for (int i = 0; i < r_; i++) { while (p[i] != i) { std::swap(p[i], p[p[i]]); swaps++; } }
  for (int i = 0; i < c_; i++) { while (q[i] != i) { std::swap(q[i], q[q[i]]); swaps++; } }
  return (swaps & 1) ? -det : det;
}
inline Dense Dense::eigenvalues() const {  // cyclic Jacobi, symmetric input
  assert(r_ == c_);
  Dense a(*this);
  const int n = r_;
  for (int sweep = 0; sweep < 100; sweep++) {
    double off = 0;
    for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) off += a(i, j) * a(i, j);
    if (off < 1e-300) break;
    for (int p = 0; p < n; p++)
      for (int q = p + 1; q < n; q++) {
        if (a(p, q) == 0) continue;
        double th = (a(q, q) - a(p, p)) / (2 * a(p, q));
        double t = (th >= 0 ? 1.0 : -1.0) / (std::fabs(th) + std::sqrt(th * th + 1));
        double c = 1 / std::sqrt(t * t + 1), s = t * c;
        for (int k = 0; k < n; k++) { double x = a(k, p), y = a(k, q); a(k, p) = c * x - s * y; a(k, q) = s * x + c * y; }
        for (int k = 0; k < n; k++) { double x = a(p, k), y = a(q, k); a(p, k) = c * x - s * y; a(q, k) = s * x + c * y; }
      }
  }
  Dense ev(n);
  for (int i = 0; i < n; i++) ev(i) = a(i, i);
  return ev;
}

// ------------------------------------------------------------------ "sparse" matrices (dense underneath)
class SparseBlockRef {
 public:
  SparseBlockRef(Dense& m, int i, int j, int r, int c) : b_(m, i, j, r, c) {}
  SparseBlockRef& operator*=(double s) { b_ *= s; return *this; }

 private:
  BlockRef b_;
};

class SparseShim {
 public:
  SparseShim() {}
  SparseShim(int r, int c) : m_(r, c) {}
  explicit SparseShim(const Dense& m) : m_(m) {}
  int rows() const { return m_.rows(); }
  int cols() const { return m_.cols(); }
  void resize(int r, int c) { m_.resize(r, c); }
  void conservativeResize(int r, int c) { m_.conservativeResize(r, c); }
  double& insert(int i, int j) {
    assert(i >= 0 && i < m_.rows() && j >= 0 && j < m_.cols());
    return m_(i, j);
  }
  double& coeffRef(int i, int j) { return m_(i, j); }
  double coeff(int i, int j) const { return m_(i, j); }
  void makeCompressed() {}
  void setZero() { m_.setZero(); }
  int nonZeros() const { int n = 0; for (int i = 0; i < m_.size(); i++) n += (m_(i) != 0); return n; }
  SparseShim transpose() const { return SparseShim(m_.transpose()); }
  SparseBlockRef block(int i, int j, int r, int c) { return SparseBlockRef(m_, i, j, r, c); }
  const Dense& dense() const { return m_; }

 private:
  Dense m_;
};
template <class T>
class SparseMatrix : public SparseShim {
 public:
  SparseMatrix() {}
  SparseMatrix(int r, int c) : SparseShim(r, c) {}
  SparseMatrix(const SparseShim& s) : SparseShim(s) {}  // NOLINT
};
inline SparseShim Dense::sparseView(double, double) const { return SparseShim(*this); }
inline std::ostream& operator<<(std::ostream& os, const SparseShim& s) { return os << s.dense(); }
inline Dense operator*(const SparseShim& a, const Dense& b) { return a.dense() * b; }

template <class I>
class COLAMDOrdering {};

class SparseQRShim;
class SparseQRMatrixQ {
 public:
  explicit SparseQRMatrixQ(const SparseQRShim& s) : s_(s) {}
  const SparseQRShim& s_;
};

class SparseQRShim {
 public:
  SparseQRShim() : qr_(0), rows_(0), cols_(0) {}
  ~SparseQRShim() { delete qr_; }
  void compute(const SparseShim& A) {
    delete qr_;
    const Dense& a = A.dense();
    rows_ = a.rows();
    cols_ = a.cols();
    // zero columns (the padding of ftsolver.cpp:138, the torso columns of :349): if they all trail, plain
    // Householder QR treats them correctly (its reflectors for them are the identity); otherwise pivot.
    bool seen_zero = false, need_pivot = false;
    double maxn = 0;
    for (int j = 0; j < cols_; j++) {
      double s = 0;
      for (int i = 0; i < rows_; i++) s += a(i, j) * a(i, j);
      if (s > maxn) maxn = s;
      if (s == 0) seen_zero = true; else if (seen_zero) need_pivot = true;
    }
    qr_ = new orc::HouseholderQR(a.to_orc(), need_pivot);
    if (!need_pivot) {
      int nz = 0;  // number of leading non-zero columns
      bool deficient = false;
      for (int j = 0; j < cols_ && j < rows_; j++) {
        double s = 0;
        for (int i = 0; i < rows_; i++) s += a(i, j) * a(i, j);
        if (s == 0) break;
        nz++;
        if (std::fabs(qr_->qr(j, j)) <= 1e-12 * std::sqrt(maxn)) deficient = true;
      }
      if (deficient) {
        delete qr_;
        qr_ = new orc::HouseholderQR(a.to_orc(), true);
      } else {
        qr_->rank_ = nz;
      }
    }
  }
  int rank() const { return qr_->rank(); }
  Dense solve(const Dense& f) const { return Dense::from_vec(qr_->solve(f.vec())); }
  SparseQRMatrixQ matrixQ() const { return SparseQRMatrixQ(*this); }
  Dense q_times(const Dense& c) const {
    orc::Vec v = c.vec();
    qr_->apply_q(v);
    return Dense::from_vec(v);
  }

 private:
  SparseQRShim(const SparseQRShim&);
  SparseQRShim& operator=(const SparseQRShim&);
  orc::HouseholderQR* qr_;
  int rows_, cols_;
};
inline Dense operator*(const SparseQRMatrixQ& Q, const Dense& c) { return Q.s_.q_times(c); }

template <class M, class O>
class SparseQR : public SparseQRShim {};

}  // namespace Eigen
#endif

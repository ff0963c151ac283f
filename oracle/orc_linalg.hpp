// TEST INFRASTRUCTURE ONLY -- part of the CPU oracle (see oracle/README.md).
// Nothing under hslabs_b200/ may include, link or call this file.
//
// Small dense linear algebra standing in for the Eigen 3 classes the
// reference calls on the gait-evaluation path (Eigen is a third-party
// dependency that is absent from /root/reference and from this image):
//   SparseQR<SpMat,COLAMDOrdering<int>>  ftsolver.h:25, ftsolver.cpp:110-129,351-353
//   FullPivLU<MatrixXd>                  ftsolver.cpp:210-217
//   colPivHouseholderQr().solve          ftsolver.cpp:227
// The sparse QR is replaced by a dense Householder QR (the quantities taken
// from it -- the unique solution of a square nonsingular system and an
// orthonormal null-space basis -- do not depend on the factorisation used).
// FullPivLU / ColPivHouseholderQR restate Eigen's published algorithms
// including their default rank thresholds (epsilon * diagonalSize), because
// the reference's rank decisions go through them.
#pragma once
#include "orc_real.hpp"
#include <cmath>
#include <cstddef>
#include <limits>
#include <vector>

namespace orc {

typedef std::vector<real> Vec;

struct Mat {  // column-major dense matrix
  int r, c;
  std::vector<real> d;
  Mat() : r(0), c(0) {}
  Mat(int r_, int c_) : r(r_), c(c_), d((size_t)r_ * c_, 0.0) {}
  real& operator()(int i, int j) { return d[(size_t)j * r + i]; }
  real operator()(int i, int j) const { return d[(size_t)j * r + i]; }
};

inline Mat transpose(const Mat& A) {
  Mat T(A.c, A.r);
  for (int j = 0; j < A.c; j++)
    for (int i = 0; i < A.r; i++) T(j, i) = A(i, j);
  return T;
}
inline Mat matmul(const Mat& A, const Mat& B) {
  Mat C(A.r, B.c);
  for (int j = 0; j < B.c; j++)
    for (int k = 0; k < A.c; k++) {
      real b = B(k, j);
      if (b == 0) continue;
      for (int i = 0; i < A.r; i++) C(i, j) += A(i, k) * b;
    }
  return C;
}
inline Vec matvec(const Mat& A, const Vec& x) {
  Vec y(A.r, 0.0);
  for (int k = 0; k < A.c; k++)
    for (int i = 0; i < A.r; i++) y[i] += A(i, k) * x[k];
  return y;
}
inline real norm2(const Vec& v) {
  real s = 0;
  for (size_t i = 0; i < v.size(); i++) s += v[i] * v[i];
  return orc::m_sqrt(s);
}

// Householder QR of an m x n matrix (m >= n), reflectors kept in place.
// Stands in for SparseQR: solve() for square nonsingular systems and
// least-squares, q_times_unit() for columns of Q (null-space basis).
class HouseholderQR {
 public:
  Mat qr;    // R in the upper triangle, essential reflector parts below
  Vec beta;  // reflector coefficients
  std::vector<int> perm;  // column permutation (identity unless pivoted)
  std::vector<std::vector<int> > vrows;  // per reflector: the rows below the diagonal in which v is non-zero
  int rank_;
  explicit HouseholderQR(const Mat& A, bool pivot = false) : qr(A), beta(A.c, 0.0), perm(A.c) {
    const int m = qr.r, n = qr.c;
    for (int j = 0; j < n; j++) perm[j] = j;
    Vec cn(n, 0.0);
    real maxcn = 0;
    for (int j = 0; j < n; j++) {
      for (int i = 0; i < m; i++) cn[j] += qr(i, j) * qr(i, j);
      if (cn[j] > maxcn) maxcn = cn[j];
    }
    rank_ = n < m ? n : m;
    bool rank_set = false;
    const int steps = n < m ? n : m;
    std::vector<int> nzr;
    nzr.reserve(m);
    vrows.assign(steps, std::vector<int>());
    for (int k = 0; k < steps; k++) {
      if (pivot) {
        int best = k;
        real bn = -1;
        for (int j = k; j < n; j++) {
          real s = 0;
          for (int i = k; i < m; i++) s += qr(i, j) * qr(i, j);
          if (s > bn) { bn = s; best = j; }
        }
        if (!rank_set && bn <= maxcn * 1e-26) { rank_ = k; rank_set = true; }
        if (best != k) {
          for (int i = 0; i < m; i++) std::swap(qr(i, k), qr(i, best));
          std::swap(perm[k], perm[best]);
        }
      }
      real s = 0;
      for (int i = k + 1; i < m; i++) s += qr(i, k) * qr(i, k);
      real a0 = qr(k, k);
      if (s == 0) { beta[k] = 0; continue; }
      real nrm = orc::m_sqrt(a0 * a0 + s);
      real alpha = (a0 >= 0) ? -nrm : nrm;
      real v0 = a0 - alpha;
      // The matrices of this path are sparse (block rows of a few entries), and so are most reflectors: only the rows in
      // which v is non-zero take part.  Skipping the others drops terms that are exactly +-0, so every result keeps its bits.
      nzr.clear();
      for (int i = k + 1; i < m; i++) {
        if (qr(i, k) != 0) { qr(i, k) /= v0; nzr.push_back(i); }  // v = [1; essential]
      }
      beta[k] = -v0 / alpha;                            // H = I - beta v v^T
      qr(k, k) = alpha;
      vrows[k] = nzr;
      const size_t nn = nzr.size();
      for (int j = k + 1; j < n; j++) {
        real w = qr(k, j);
        for (size_t t = 0; t < nn; t++) w += qr(nzr[t], k) * qr(nzr[t], j);
        if (w == 0) continue;
        w *= beta[k];
        qr(k, j) -= w;
        for (size_t t = 0; t < nn; t++) qr(nzr[t], j) -= w * qr(nzr[t], k);
      }
    }
  }
  int rank() const { return rank_; }
  void apply_qt(Vec& b) const {  // b <- Q^T b
    const int m = qr.r, steps = qr.c < qr.r ? qr.c : qr.r;
    (void)m;
    for (int k = 0; k < steps; k++) {
      if (beta[k] == 0) continue;
      const std::vector<int>& nz = vrows[k];   // the other rows contribute exact zeros
      real w = b[k];
      for (size_t t = 0; t < nz.size(); t++) w += qr(nz[t], k) * b[nz[t]];
      if (w == 0) continue;
      w *= beta[k];
      b[k] -= w;
      for (size_t t = 0; t < nz.size(); t++) b[nz[t]] -= w * qr(nz[t], k);
    }
  }
  void apply_q(Vec& b) const {  // b <- Q b
    const int m = qr.r, steps = qr.c < qr.r ? qr.c : qr.r;
    (void)m;
    for (int k = steps - 1; k >= 0; k--) {
      if (beta[k] == 0) continue;
      const std::vector<int>& nz = vrows[k];
      real w = b[k];
      for (size_t t = 0; t < nz.size(); t++) w += qr(nz[t], k) * b[nz[t]];
      if (w == 0) continue;
      w *= beta[k];
      b[k] -= w;
      for (size_t t = 0; t < nz.size(); t++) b[nz[t]] -= w * qr(nz[t], k);
    }
  }
  // Basic least-squares solution (free variables = 0), like SparseQR::solve.
  Vec solve(const Vec& f) const {
    Vec b(f);
    apply_qt(b);
    const int n = qr.c, rk = rank_;
    Vec y(n, 0.0);
    for (int i = rk - 1; i >= 0; i--) {
      real s = b[i];
      for (int j = i + 1; j < rk; j++) s -= qr(i, j) * y[j];
      y[i] = s / qr(i, i);
    }
    Vec x(n, 0.0);
    for (int j = 0; j < n; j++) x[perm[j]] = y[j];
    return x;
  }
  Vec q_times_unit(int j) const {  // Q e_j
    Vec e(qr.r, 0.0);
    e[j] = 1;
    apply_q(e);
    return e;
  }
};

// Restatement of Eigen::FullPivLU (complete pivoting, rank by threshold).
class FullPivLU {
 public:
  Mat lu;
  std::vector<int> p, q;  // row / column permutations: (P A Q)(i,j) = A(p[i], q[j])
  int nonzero_pivots;
  real maxpivot, thresh;
  explicit FullPivLU(const Mat& A) : lu(A), p(A.r), q(A.c) {
    const int rows = lu.r, cols = lu.c, size = rows < cols ? rows : cols;
    for (int i = 0; i < rows; i++) p[i] = i;
    for (int j = 0; j < cols; j++) q[j] = j;
    nonzero_pivots = size;
    maxpivot = 0;
    thresh = std::numeric_limits<double>::epsilon() * size;
    for (int k = 0; k < size; k++) {
      int br = k, bc = k;
      real big = -1;
      for (int j = k; j < cols; j++)  // column-major visit order, first maximum wins
        for (int i = k; i < rows; i++) {
          real a = orc::m_fabs(lu(i, j));
          if (a > big) { big = a; br = i; bc = j; }
        }
      if (big == 0) { nonzero_pivots = k; break; }
      if (big > maxpivot) maxpivot = big;
      if (br != k) { for (int j = 0; j < cols; j++) std::swap(lu(k, j), lu(br, j)); std::swap(p[k], p[br]); }
      if (bc != k) { for (int i = 0; i < rows; i++) std::swap(lu(i, k), lu(i, bc)); std::swap(q[k], q[bc]); }
      for (int i = k + 1; i < rows; i++) lu(i, k) /= lu(k, k);
      for (int j = k + 1; j < cols; j++) {
        real u = lu(k, j);
        if (u == 0) continue;
        for (int i = k + 1; i < rows; i++) lu(i, j) -= lu(i, k) * u;
      }
    }
  }
  real threshold() const { return thresh; }
  void setThreshold(real t) { thresh = t; }
  int rank() const {
    real pt = orc::m_fabs(maxpivot) * thresh;
    int r = 0;
    for (int i = 0; i < nonzero_pivots; i++) r += (orc::m_fabs(lu(i, i)) > pt);
    return r;
  }
  // Square systems only (that is all the path uses).
  Vec solve(const Vec& b) const {
    const int n = lu.r, rk = rank();
    Vec c(n);
    for (int i = 0; i < n; i++) c[i] = b[p[i]];
    for (int j = 0; j < n; j++)  // unit lower
      for (int i = j + 1; i < n; i++) c[i] -= lu(i, j) * c[j];
    for (int i = rk - 1; i >= 0; i--) {
      real s = c[i];
      for (int j = i + 1; j < rk; j++) s -= lu(i, j) * c[j];
      c[i] = s / lu(i, i);
    }
    Vec x(lu.c, 0.0);
    for (int i = 0; i < rk; i++) x[q[i]] = c[i];
    return x;
  }
  std::vector<int> pivots() const {
    real pt = orc::m_fabs(maxpivot) * thresh;
    std::vector<int> pv;
    for (int i = 0; i < nonzero_pivots; i++)
      if (orc::m_fabs(lu(i, i)) > pt) pv.push_back(i);
    return pv;
  }
  Mat kernel() const {
    const int cols = lu.c;
    std::vector<int> pv = pivots();
    const int rk = (int)pv.size(), dimker = cols - rk;
    Mat K(cols, dimker);
    if (dimker == 0) return K;
    Mat m(rk, cols);
    for (int i = 0; i < rk; i++)
      for (int j = i; j < cols; j++) m(i, j) = lu(pv[i], j);
    for (int i = 0; i < rk; i++)
      if (pv[i] != i) for (int r = 0; r < rk; r++) std::swap(m(r, i), m(r, pv[i]));
    for (int k = 0; k < dimker; k++)  // upper-triangular solve on the trailing block
      for (int i = rk - 1; i >= 0; i--) {
        real s = m(i, rk + k);
        for (int j = i + 1; j < rk; j++) s -= m(i, j) * m(j, rk + k);
        m(i, rk + k) = s / m(i, i);
      }
    for (int i = rk - 1; i >= 0; i--)
      if (pv[i] != i) for (int r = 0; r < rk; r++) std::swap(m(r, i), m(r, pv[i]));
    for (int i = 0; i < rk; i++)
      for (int k = 0; k < dimker; k++) K(q[i], k) = -m(i, rk + k);
    for (int k = 0; k < dimker; k++) K(q[rk + k], k) = 1;
    return K;
  }
  Mat image(const Mat& original) const {
    std::vector<int> pv = pivots();
    Mat I(original.r, (int)pv.size());
    for (size_t i = 0; i < pv.size(); i++)
      for (int r = 0; r < original.r; r++) I(r, (int)i) = original(r, q[pv[i]]);
    return I;
  }
};

// Restatement of Eigen::ColPivHouseholderQR::solve for square systems.
inline Vec colpiv_qr_solve(const Mat& A, const Vec& b) {
  const int rows = A.r, cols = A.c, size = rows < cols ? rows : cols;
  Mat qr(A);
  std::vector<int> perm(cols);
  Vec cn(cols), hb(size, 0.0);
  for (int j = 0; j < cols; j++) {
    perm[j] = j;
    real s = 0;
    for (int i = 0; i < rows; i++) s += qr(i, j) * qr(i, j);
    cn[j] = orc::m_sqrt(s);
  }
  real mx = 0;
  for (int j = 0; j < cols; j++) if (cn[j] > mx) mx = cn[j];
  const real eps = std::numeric_limits<double>::epsilon();
  real helper = (mx * eps / rows) * (mx * eps / rows);
  int nonzero = size;
  for (int k = 0; k < size; k++) {
    int best = k;
    real bn = -1;
    for (int j = k; j < cols; j++) {  // recomputed norms (Eigen down-dates; same pivots away from ties)
      real s = 0;
      for (int i = k; i < rows; i++) s += qr(i, j) * qr(i, j);
      if (s > bn) { bn = s; best = j; }
    }
    if (nonzero == size && bn < helper * (rows - k)) nonzero = k;
    if (best != k) { for (int i = 0; i < rows; i++) std::swap(qr(i, k), qr(i, best)); std::swap(perm[k], perm[best]); }
    real s = 0;
    for (int i = k + 1; i < rows; i++) s += qr(i, k) * qr(i, k);
    real a0 = qr(k, k);
    if (s == 0) { hb[k] = 0; continue; }
    real nrm = orc::m_sqrt(a0 * a0 + s), alpha = (a0 >= 0) ? -nrm : nrm, v0 = a0 - alpha;
    for (int i = k + 1; i < rows; i++) qr(i, k) /= v0;
    hb[k] = -v0 / alpha;
    qr(k, k) = alpha;
    for (int j = k + 1; j < cols; j++) {
      real w = qr(k, j);
      for (int i = k + 1; i < rows; i++) w += qr(i, k) * qr(i, j);
      w *= hb[k];
      qr(k, j) -= w;
      for (int i = k + 1; i < rows; i++) qr(i, j) -= w * qr(i, k);
    }
  }
  Vec c(b);
  for (int k = 0; k < nonzero; k++) {
    if (hb[k] == 0) continue;
    real w = c[k];
    for (int i = k + 1; i < rows; i++) w += qr(i, k) * c[i];
    w *= hb[k];
    c[k] -= w;
    for (int i = k + 1; i < rows; i++) c[i] -= w * qr(i, k);
  }
  for (int i = nonzero - 1; i >= 0; i--) {
    real s = c[i];
    for (int j = i + 1; j < nonzero; j++) s -= qr(i, j) * c[j];
    c[i] = s / qr(i, i);
  }
  Vec x(cols, 0.0);
  for (int i = 0; i < nonzero; i++) x[perm[i]] = c[i];
  return x;
}

}  // namespace orc

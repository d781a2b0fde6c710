// TEST INFRASTRUCTURE ONLY — tiny dense linear algebra for the CPU oracle.
// Stands in for the Eigen/Sophus calls on the reference's host path
// (Core/Utils/RGBDOdometry.cpp:309-367,407-417,522-534,566-575; Core/Utils/OdometryProvider.h:34-96;
//  Core/ElasticFusion.cpp:371-374). Eigen/Sophus are vendored in the reference tree only (third-party/Eigen
// 3.3.90 @8e47906, third-party/Sophus @26c2002) and do not exist on the GPU box, so the published
// algorithms are restated: LDLT with diagonal pivoting (Eigen::LDLT), cofactor 3x3 inverse, Gauss-Jordan
// 4x4 / NxN inverse, Rodrigues, orthogonal polar factor (== JacobiSVD U*V^T), SE3 logarithm.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstring>
#include <limits>

namespace la {

static inline void mul3(const double* a, const double* b, double* c) {
  double r[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r[i * 3 + j] = a[i * 3 + 0] * b[0 * 3 + j] + a[i * 3 + 1] * b[1 * 3 + j] + a[i * 3 + 2] * b[2 * 3 + j];
  memcpy(c, r, sizeof(r));
}
static inline void mulv3(const double* a, const double* v, double* o) {
  double r[3];
  for (int i = 0; i < 3; ++i) r[i] = a[i * 3 + 0] * v[0] + a[i * 3 + 1] * v[1] + a[i * 3 + 2] * v[2];
  memcpy(o, r, sizeof(r));
}
static inline void mul4(const double* a, const double* b, double* c) {
  double r[16];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = 0;
      for (int k = 0; k < 4; ++k) s += a[i * 4 + k] * b[k * 4 + j];
      r[i * 4 + j] = s;
    }
  memcpy(c, r, sizeof(r));
}

template <typename T>
static inline void inv3_t(const T* m, T* o) {
  T c00 = m[4] * m[8] - m[5] * m[7];
  T c01 = m[5] * m[6] - m[3] * m[8];
  T c02 = m[3] * m[7] - m[4] * m[6];
  T det = m[0] * c00 + m[1] * c01 + m[2] * c02;
  T id = T(1) / det;
  T r[9];
  r[0] = c00 * id;
  r[1] = (m[2] * m[7] - m[1] * m[8]) * id;
  r[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  r[3] = c01 * id;
  r[4] = (m[0] * m[8] - m[2] * m[6]) * id;
  r[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  r[6] = c02 * id;
  r[7] = (m[1] * m[6] - m[0] * m[7]) * id;
  r[8] = (m[0] * m[4] - m[1] * m[3]) * id;
  memcpy(o, r, sizeof(r));
}
static inline void inv3(const double* m, double* o) { inv3_t<double>(m, o); }
static inline void inv3f(const float* m, float* o) { inv3_t<float>(m, o); }

// general NxN inverse, Gauss-Jordan with partial pivoting (N <= 6)
static inline bool inv_n(const double* m, double* o, int n) {
  double a[6][12];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      a[i][j] = m[i * n + j];
      a[i][n + j] = (i == j) ? 1.0 : 0.0;
    }
  for (int c = 0; c < n; ++c) {
    int p = c;
    for (int r = c + 1; r < n; ++r)
      if (std::fabs(a[r][c]) > std::fabs(a[p][c])) p = r;
    if (p != c)
      for (int j = 0; j < 2 * n; ++j) {
        double t = a[c][j];
        a[c][j] = a[p][j];
        a[p][j] = t;
      }
    double d = a[c][c];
    for (int j = 0; j < 2 * n; ++j) a[c][j] /= d;
    for (int r = 0; r < n; ++r)
      if (r != c) {
        double f = a[r][c];
        if (f != 0)
          for (int j = 0; j < 2 * n; ++j) a[r][j] -= f * a[c][j];
      }
  }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) o[i * n + j] = a[i][n + j];
  return true;
}
static inline void inv4(const double* m, double* o) { inv_n(m, o, 4); }

// Eigen::LDLT-style solve: symmetric A (row-major, lower triangle used), diagonal pivoting,
// pivots with |D_i| <= 1/highest() are treated as zero in the solve (Eigen 3.3 LDLT::_solve_impl).
template <typename T, int N>
static inline void ldlt_solve(const T* A_, const T* b_, T* x) {
  T A[N][N];
  int perm[N];
  for (int i = 0; i < N; ++i) {
    perm[i] = i;
    for (int j = 0; j < N; ++j) A[i][j] = A_[i * N + j];
  }
  T bb[N];
  for (int i = 0; i < N; ++i) bb[i] = b_[i];
  for (int k = 0; k < N; ++k) {
    // largest remaining diagonal
    int p = k;
    T big = std::fabs(A[k][k]);
    for (int i = k + 1; i < N; ++i)
      if (std::fabs(A[i][i]) > big) {
        big = std::fabs(A[i][i]);
        p = i;
      }
    if (p != k) {
      // symmetric row/column swap of the full matrix + rhs
      for (int j = 0; j < N; ++j) {
        T t = A[k][j];
        A[k][j] = A[p][j];
        A[p][j] = t;
      }
      for (int i = 0; i < N; ++i) {
        T t = A[i][k];
        A[i][k] = A[i][p];
        A[i][p] = t;
      }
      T t = bb[k];
      bb[k] = bb[p];
      bb[p] = t;
      int ti = perm[k];
      perm[k] = perm[p];
      perm[p] = ti;
    }
    T d = A[k][k];
    if (d == T(0)) continue;  // leaves column k of L as (unit, zeros below are whatever: masked by zero pivot)
    for (int i = k + 1; i < N; ++i) A[i][k] /= d;
    for (int i = k + 1; i < N; ++i)
      for (int j = k + 1; j <= i; ++j) {
        A[i][j] -= A[i][k] * d * A[j][k];
        A[j][i] = A[i][j];
      }
  }
  // forward: L y = P b
  T y[N];
  for (int i = 0; i < N; ++i) {
    T s = bb[i];
    for (int j = 0; j < i; ++j) s -= A[i][j] * y[j];
    y[i] = s;
  }
  const T tol = T(1) / std::numeric_limits<T>::max();
  for (int i = 0; i < N; ++i) y[i] = (std::fabs(A[i][i]) > tol) ? y[i] / A[i][i] : T(0);
  // backward: L^T z = y
  T z[N];
  for (int i = N - 1; i >= 0; --i) {
    T s = y[i];
    for (int j = i + 1; j < N; ++j) s -= A[j][i] * z[j];
    z[i] = s;
  }
  for (int i = 0; i < N; ++i) x[perm[i]] = z[i];
}
static inline void solve_sym6(const double* A, const double* b, double* x) { ldlt_solve<double, 6>(A, b, x); }
static inline void solve_sym3f(const float* A, const float* b, float* x) { ldlt_solve<float, 3>(A, b, x); }

// OdometryProvider::rodrigues, OdometryProvider.h:34-71
static inline void rodrigues(const double* src, double* dst) {
  const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  memcpy(dst, I, sizeof(I));
  double rx = src[0], ry = src[1], rz = src[2];
  double theta = std::sqrt(rx * rx + ry * ry + rz * rz);
  if (theta >= DBL_EPSILON) {
    double c = std::cos(theta), s = std::sin(theta), c1 = 1. - c;
    double itheta = theta ? 1. / theta : 0.;
    rx *= itheta;
    ry *= itheta;
    rz *= itheta;
    double rrt[9] = {rx * rx, rx * ry, rx * rz, rx * ry, ry * ry, ry * rz, rx * rz, ry * rz, rz * rz};
    double r_x[9] = {0, -rz, ry, rz, 0, -rx, -ry, rx, 0};
    for (int k = 0; k < 9; ++k) dst[k] = c * I[k] + c1 * rrt[k] + s * r_x[k];
  }
}

// Orthogonal polar factor of a near-rotation matrix: Newton iteration X <- (X + X^-T)/2.
// Equals U*V^T of the SVD (RGBDOdometry.cpp:566-570) for det > 0.
static inline void polar_orthogonal(const double* m, double* o) {
  double X[9];
  memcpy(X, m, sizeof(X));
  for (int it = 0; it < 30; ++it) {
    double Xi[9];
    inv3(X, Xi);
    double N[9], delta = 0;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        N[r * 3 + c] = 0.5 * (X[r * 3 + c] + Xi[c * 3 + r]);
        delta += std::fabs(N[r * 3 + c] - X[r * 3 + c]);
      }
    memcpy(X, N, sizeof(N));
    if (delta < 1e-17) break;
  }
  memcpy(o, X, sizeof(X));
}

// rigid inverse of a row-major 4x4 pose (Sophus SE3::inverse)
static inline void se3_inverse(const double* T, double* o) {
  double r[16] = {0};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r[i * 4 + j] = T[j * 4 + i];
  for (int i = 0; i < 3; ++i) r[i * 4 + 3] = -(r[i * 4 + 0] * T[3] + r[i * 4 + 1] * T[7] + r[i * 4 + 2] * T[11]);
  r[15] = 1;
  memcpy(o, r, sizeof(r));
}

// |log(T)| of an SE3 (Sophus SE3::log = [V^-1 t, omega]); used only for the velocity weighting
// (ElasticFusion.cpp:371-383), which saturates at 0.01.
static inline double se3_log_norm(const double* T) {
  double R[9] = {T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]};
  double t[3] = {T[3], T[7], T[11]};
  double tr = R[0] + R[4] + R[8];
  double cs = (tr - 1.0) * 0.5;
  if (cs > 1) cs = 1;
  if (cs < -1) cs = -1;
  double ax[3] = {R[7] - R[5], R[2] - R[6], R[3] - R[1]};  // 2 sin(theta) * axis
  double sn = 0.5 * std::sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
  double theta = std::atan2(sn, cs);
  double w[3];
  if (theta < 1e-10) {
    w[0] = 0.5 * ax[0];
    w[1] = 0.5 * ax[1];
    w[2] = 0.5 * ax[2];
  } else {
    double k = theta / (2.0 * sn);
    w[0] = k * ax[0];
    w[1] = k * ax[1];
    w[2] = k * ax[2];
  }
  // V^-1 = I - 0.5*W + (1/theta^2)(1 - (theta sin)/(2(1-cos))) W^2
  double W[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
  double W2[9];
  mul3(W, W, W2);
  double coef;
  if (theta < 1e-5)
    coef = 1.0 / 12.0;
  else
    coef = (1.0 - (theta * std::sin(theta)) / (2.0 * (1.0 - std::cos(theta)))) / (theta * theta);
  double u[3];
  for (int i = 0; i < 3; ++i) {
    u[i] = 0;
    for (int j = 0; j < 3; ++j) {
      double Vinv = (i == j ? 1.0 : 0.0) - 0.5 * W[i * 3 + j] + coef * W2[i * 3 + j];
      u[i] += Vinv * t[j];
    }
  }
  return std::sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2] + w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
}

}  // namespace la

// TEST INFRASTRUCTURE ONLY -- stand-in for the sliver of OpenCV (absent here) that
// vtkOpenSURF3D/fasthessian.cxx uses: cv::Matx small fixed matrices and cv::SVD::compute on a 4 x 4
// double matrix (FastHessian::interpolateStep, fasthessian.cxx:614-661).  The SVD is a one-sided
// (Hestenes) Jacobi iteration in double -- the published algorithm OpenCV's own JacobiSVD follows --
// written from the algorithm, not from OpenCV's sources.  OpenCV's exact bits (which depend on its
// version and on whether it was built against LAPACK) cannot be reproduced, so the sub-voxel
// interpolation offsets are pinned to a TOLERANCE (see tests/test_surf_oracle.py), everything else
// in the producer bit for bit.
#ifndef ORACLE_OPENCV_SHIM_HPP
#define ORACLE_OPENCV_SHIM_HPP

#include <cmath>
#include <cstdlib>
#include <utility>

namespace cv {

class Mat;

template <typename T, int m, int n>
class Matx {
 public:
  enum { rows = m, cols = n };
  T val[m * n];
  Matx() { for (int i = 0; i < m * n; i++) val[i] = T(0); }
  static Matx zeros() { return Matx(); }
  T& operator()(int i, int j) { return val[i * n + j]; }
  const T& operator()(int i, int j) const { return val[i * n + j]; }
  T& operator()(int i) { return val[i]; }
  const T& operator()(int i) const { return val[i]; }
  Matx<T, n, m> t() const {
    Matx<T, n, m> r;
    for (int i = 0; i < m; i++) for (int j = 0; j < n; j++) r(j, i) = (*this)(i, j);
    return r;
  }
  Matx operator-() const { Matx r; for (int i = 0; i < m * n; i++) r.val[i] = -val[i]; return r; }
};

template <typename T, int m, int k, int n>
Matx<T, m, n> operator*(const Matx<T, m, k>& a, const Matx<T, k, n>& b) {
  Matx<T, m, n> r;
  for (int i = 0; i < m; i++)
    for (int j = 0; j < n; j++) {
      T s = 0;
      for (int q = 0; q < k; q++) s += a(i, q) * b(q, j);
      r(i, j) = s;
    }
  return r;
}

typedef Matx<double, 4, 1> Matx41d;
typedef Matx<double, 4, 4> Matx44d;

class SVD {
 public:
  // a = u * diag(w) * vt, singular values descending (OpenCV's convention)
  template <int n>
  static void compute(const Matx<double, n, n>& a, Matx<double, n, 1>& w, Matx<double, n, n>& u, Matx<double, n, n>& vt) {
    double g[n][n], v[n][n], s2[n];  // g: rows = columns of a, rotated until mutually orthogonal
    for (int i = 0; i < n; i++) {
      s2[i] = 0;
      for (int k = 0; k < n; k++) { g[i][k] = a(k, i); s2[i] += g[i][k] * g[i][k]; v[i][k] = (i == k); }
    }
    const double eps = 2.220446049250313e-16 * 10;
    for (int sweep = 0; sweep < 60; sweep++) {
      bool rotated = false;
      for (int i = 0; i < n - 1; i++)
        for (int j = i + 1; j < n; j++) {
          double p = 0;
          for (int k = 0; k < n; k++) p += g[i][k] * g[j][k];
          if (std::fabs(p) <= eps * std::sqrt(s2[i] * s2[j])) continue;
          p *= 2;
          double beta = s2[i] - s2[j], gamma = std::hypot(p, beta), c, s;
          if (beta < 0) { s = std::sqrt((gamma - beta) * 0.5 / gamma); c = p / (gamma * s * 2); }
          else { c = std::sqrt((gamma + beta) / (gamma * 2)); s = p / (gamma * c * 2); }
          double ni = 0, nj = 0;
          for (int k = 0; k < n; k++) {
            double t0 = c * g[i][k] + s * g[j][k], t1 = c * g[j][k] - s * g[i][k];
            g[i][k] = t0; g[j][k] = t1; ni += t0 * t0; nj += t1 * t1;
            double u0 = c * v[i][k] + s * v[j][k], u1 = c * v[j][k] - s * v[i][k];
            v[i][k] = u0; v[j][k] = u1;
          }
          s2[i] = ni; s2[j] = nj;
          rotated = true;
        }
      if (!rotated) break;
    }
    int order[n];
    double sv[n];
    for (int i = 0; i < n; i++) {
      double s = 0;
      for (int k = 0; k < n; k++) s += g[i][k] * g[i][k];
      sv[i] = std::sqrt(s); order[i] = i;
    }
    for (int i = 0; i < n - 1; i++) {
      int best = i;
      for (int k = i + 1; k < n; k++) if (sv[order[best]] < sv[order[k]]) best = k;
      std::swap(order[i], order[best]);
    }
    for (int i = 0; i < n; i++) {
      int o = order[i];
      w(i) = sv[o];
      for (int k = 0; k < n; k++) {
        u(k, i) = sv[o] > 0 ? g[o][k] / sv[o] : 0.0;
        vt(i, k) = v[o][k];
      }
    }
  }
  // FastHessian::FittingQuadric (never called by the producer) instantiates other shapes
  template <class A, class B, class C, class D>
  static void compute(const A&, B&, C&, D&) { std::abort(); }
};

}  // namespace cv

#endif

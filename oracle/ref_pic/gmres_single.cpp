// gmres_single.cpp -- OURS (test infrastructure): restarted GMRES(m) for one rank behind the reference's linear_solver_wrapper call
// (see linear_solver_wrapper_c.h).  x0 = Sol_I on entry; stops when |r| <= Tol |r0| or after nMaxIter matrix-vector products.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>
#include "linear_solver_wrapper_c.h"

void (*linear_solver_matvec_c)(double *VecIn, double *VecOut, int n) = nullptr;
int ref_gmres_last_iterations = 0;
double ref_gmres_last_relative_residual = 0.0;

static double dot(const double *a, const double *b, int n) {
  double s = 0.0;
  for (int i = 0; i < n; i++) s += a[i] * b[i];
  return s;
}

void linear_solver_wrapper(const char *, double *Tol, int *nMaxIter, int *nVar, int *, int *nI, int *nJ, int *nK, int *nBlock, int *, double *rhs,
                           double *x, double *, double *, int *) {
  const int n = (*nVar) * (*nI) * (*nJ) * (*nK) * (*nBlock);
  const int maxIter = *nMaxIter, m = maxIter < 100 ? maxIter : 100;
  std::vector<double> V((size_t)(m + 1) * n), H((size_t)(m + 1) * m, 0.0), cs(m), sn(m), g(m + 1), w(n), y(m);
  int iters = 0;
  double r0norm = -1.0, rel = 1.0;
  while (iters < maxIter) {
    linear_solver_matvec_c(x, w.data(), n);
    double *v0 = V.data();
    for (int i = 0; i < n; i++) v0[i] = rhs[i] - w[i];
    const double beta = std::sqrt(dot(v0, v0, n));
    if (r0norm < 0.0) r0norm = beta;
    rel = r0norm > 0.0 ? beta / r0norm : 0.0;
    if (beta == 0.0 || rel <= *Tol) break;
    for (int i = 0; i < n; i++) v0[i] /= beta;
    std::fill(g.begin(), g.end(), 0.0);
    g[0] = beta;
    int j = 0;
    for (; j < m && iters < maxIter; j++) {
      double *vj = V.data() + (size_t)j * n, *vn = V.data() + (size_t)(j + 1) * n;
      linear_solver_matvec_c(vj, vn, n);
      iters++;
      for (int i = 0; i <= j; i++) {
        const double h = dot(vn, V.data() + (size_t)i * n, n);
        H[(size_t)i * m + j] = h;
        const double *vi = V.data() + (size_t)i * n;
        for (int q = 0; q < n; q++) vn[q] -= h * vi[q];
      }
      const double hn = std::sqrt(dot(vn, vn, n));
      H[(size_t)(j + 1) * m + j] = hn;
      if (hn > 0.0)
        for (int q = 0; q < n; q++) vn[q] /= hn;
      for (int i = 0; i < j; i++) {
        const double t = cs[i] * H[(size_t)i * m + j] + sn[i] * H[(size_t)(i + 1) * m + j];
        H[(size_t)(i + 1) * m + j] = -sn[i] * H[(size_t)i * m + j] + cs[i] * H[(size_t)(i + 1) * m + j];
        H[(size_t)i * m + j] = t;
      }
      const double a = H[(size_t)j * m + j], b = H[(size_t)(j + 1) * m + j], d = std::sqrt(a * a + b * b);
      cs[j] = a / d, sn[j] = b / d;
      H[(size_t)j * m + j] = d, H[(size_t)(j + 1) * m + j] = 0.0;
      g[j + 1] = -sn[j] * g[j];
      g[j] = cs[j] * g[j];
      rel = std::fabs(g[j + 1]) / r0norm;
      if (rel <= *Tol) {
        j++;
        break;
      }
    }
    for (int i = j - 1; i >= 0; i--) {
      double s = g[i];
      for (int q = i + 1; q < j; q++) s -= H[(size_t)i * m + q] * y[q];
      y[i] = s / H[(size_t)i * m + i];
    }
    for (int i = 0; i < j; i++) {
      const double *vi = V.data() + (size_t)i * n;
      for (int q = 0; q < n; q++) x[q] += y[i] * vi[q];
    }
    if (rel <= *Tol) break;
  }
  ref_gmres_last_iterations = iters;
  ref_gmres_last_relative_residual = rel;
}

#include "gaussian_process.h"
#include <cmath>
#include <limits>

namespace hvd {

bool Cholesky(const Mat& a, Mat* l) {
  size_t n = a.size();
  l->assign(n, Vec(n, 0.0));
  for (size_t i = 0; i < n; ++i) {
    for (size_t j = 0; j <= i; ++j) {
      double s = a[i][j];
      for (size_t k = 0; k < j; ++k) s -= (*l)[i][k] * (*l)[j][k];
      if (i == j) {
        if (s <= 0) return false;
        (*l)[i][i] = std::sqrt(s);
      } else {
        (*l)[i][j] = s / (*l)[j][j];
      }
    }
  }
  return true;
}

Vec CholeskySolve(const Mat& l, const Vec& b) {
  size_t n = b.size();
  Vec y(n), x(n);
  for (size_t i = 0; i < n; ++i) {
    double s = b[i];
    for (size_t k = 0; k < i; ++k) s -= l[i][k] * y[k];
    y[i] = s / l[i][i];
  }
  for (size_t ii = n; ii-- > 0;) {
    double s = y[ii];
    for (size_t k = ii + 1; k < n; ++k) s -= l[k][ii] * x[k];
    x[ii] = s / l[ii][ii];
  }
  return x;
}

double GaussianProcessRegressor::Kernel(const Vec& a, const Vec& b, double length, double var) const {
  double d2 = 0;
  for (size_t i = 0; i < a.size(); ++i) d2 += (a[i] - b[i]) * (a[i] - b[i]);
  return var * std::exp(-0.5 * d2 / (length * length));
}

double GaussianProcessRegressor::LogMarginalLikelihood(const Mat& x, const Vec& y, double length, double var, Mat* l,
                                                       Vec* alpha) const {
  size_t n = x.size();
  Mat k(n, Vec(n));
  for (size_t i = 0; i < n; ++i)
    for (size_t j = 0; j < n; ++j) k[i][j] = Kernel(x[i], x[j], length, var) + (i == j ? noise_ * noise_ + 1e-10 : 0.0);
  if (!Cholesky(k, l)) return -std::numeric_limits<double>::infinity();
  *alpha = CholeskySolve(*l, y);
  double lml = 0;
  for (size_t i = 0; i < n; ++i) lml += -0.5 * y[i] * (*alpha)[i] - std::log((*l)[i][i]);
  lml -= 0.5 * n * std::log(2 * M_PI);
  return lml;
}

void GaussianProcessRegressor::Fit(const Mat& x, const Vec& y) {
  x_ = x;
  size_t n = y.size();
  y_mean_ = 0;
  for (double v : y) y_mean_ += v;
  y_mean_ /= (double)n;
  double var = 0;
  for (double v : y) var += (v - y_mean_) * (v - y_mean_);
  y_std_ = n > 1 ? std::sqrt(var / (double)n) : 1.0;
  if (y_std_ < 1e-12) y_std_ = 1.0;
  Vec yn(n);
  for (size_t i = 0; i < n; ++i) yn[i] = (y[i] - y_mean_) / y_std_;
  double best = -std::numeric_limits<double>::infinity();
  for (double length = 0.05; length <= 4.0; length *= 1.5) {
    for (double v : {0.25, 1.0, 4.0}) {
      Mat l; Vec alpha;
      double lml = LogMarginalLikelihood(x, yn, length, v, &l, &alpha);
      if (lml > best) { best = lml; length_ = length; var_ = v; l_ = std::move(l); alpha_ = std::move(alpha); }
    }
  }
}

void GaussianProcessRegressor::Predict(const Vec& x, double* mu, double* sigma) const {
  size_t n = x_.size();
  if (n == 0) { *mu = 0; *sigma = 1; return; }
  Vec ks(n);
  for (size_t i = 0; i < n; ++i) ks[i] = Kernel(x, x_[i], length_, var_);
  double m = 0;
  for (size_t i = 0; i < n; ++i) m += ks[i] * alpha_[i];
  Vec v = CholeskySolve(l_, ks);
  double s2 = var_;
  for (size_t i = 0; i < n; ++i) s2 -= ks[i] * v[i];
  if (s2 < 1e-12) s2 = 1e-12;
  *mu = m * y_std_ + y_mean_;
  *sigma = std::sqrt(s2) * y_std_;
}

}  // namespace hvd

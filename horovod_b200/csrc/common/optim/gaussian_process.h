// Gaussian-process regression (RBF kernel) on a handful of samples, with a
// self-contained dense Cholesky solver.
// Parity: horovod/common/optim/gaussian_process.{h,cc}; the reference uses
// Eigen + LBFGS++ (unavailable offline) to fit kernel hyper-parameters by
// L-BFGS — here the length scale / signal variance are fitted by a log-spaced
// grid search over the log marginal likelihood (documented deviation).
#pragma once
#include <vector>

namespace hvd {

using Vec = std::vector<double>;
using Mat = std::vector<Vec>;

// Cholesky factor L (lower) of SPD matrix a; returns false if not SPD.
bool Cholesky(const Mat& a, Mat* l);
Vec CholeskySolve(const Mat& l, const Vec& b);  // solves (L L^T) x = b

class GaussianProcessRegressor {
 public:
  explicit GaussianProcessRegressor(double noise = 0.8) : noise_(noise) {}
  void Fit(const Mat& x, const Vec& y);
  // predictive mean and standard deviation at x
  void Predict(const Vec& x, double* mu, double* sigma) const;
  double length_scale() const { return length_; }

 private:
  double Kernel(const Vec& a, const Vec& b, double length, double var) const;
  double LogMarginalLikelihood(const Mat& x, const Vec& y, double length, double var, Mat* l, Vec* alpha) const;
  double noise_;
  double length_ = 1.0, var_ = 1.0;
  double y_mean_ = 0.0, y_std_ = 1.0;
  Mat x_;
  Mat l_;
  Vec alpha_;
};

}  // namespace hvd

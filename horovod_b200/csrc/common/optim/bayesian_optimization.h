// Bayesian optimisation (GP surrogate + expected improvement) over a box.
// Parity: horovod/common/optim/bayesian_optimization.{h,cc}.
#pragma once
#include <random>
#include <utility>
#include <vector>
#include "gaussian_process.h"

namespace hvd {

class BayesianOptimization {
 public:
  // bounds: per-dimension [lo, hi]; alpha: GP noise; xi: exploration margin
  BayesianOptimization(std::vector<std::pair<double, double>> bounds, double alpha, double xi = 0.01);
  void AddSample(const Vec& x, double y);
  Vec NextSample(bool normalize_known = true);
  void Clear();
  size_t num_samples() const { return ys_.size(); }

 private:
  Vec Normalize(const Vec& x) const;
  Vec Denormalize(const Vec& u) const;
  double ExpectedImprovement(const Vec& u, double best) const;
  std::vector<std::pair<double, double>> bounds_;
  double xi_;
  GaussianProcessRegressor gp_;
  Mat xs_;  // normalised to [0,1]^d
  Vec ys_;
  std::mt19937 rng_{1234};
};

}  // namespace hvd

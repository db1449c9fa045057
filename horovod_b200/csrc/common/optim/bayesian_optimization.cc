#include "bayesian_optimization.h"
#include <algorithm>
#include <cmath>

namespace hvd {

namespace {
double NormPdf(double z) { return std::exp(-0.5 * z * z) / std::sqrt(2 * M_PI); }
double NormCdf(double z) { return 0.5 * std::erfc(-z / std::sqrt(2.0)); }
}  // namespace

BayesianOptimization::BayesianOptimization(std::vector<std::pair<double, double>> bounds, double alpha, double xi)
    : bounds_(std::move(bounds)), xi_(xi), gp_(alpha) {}

Vec BayesianOptimization::Normalize(const Vec& x) const {
  Vec u(x.size());
  for (size_t i = 0; i < x.size(); ++i) u[i] = (x[i] - bounds_[i].first) / (bounds_[i].second - bounds_[i].first);
  return u;
}
Vec BayesianOptimization::Denormalize(const Vec& u) const {
  Vec x(u.size());
  for (size_t i = 0; i < u.size(); ++i) x[i] = bounds_[i].first + u[i] * (bounds_[i].second - bounds_[i].first);
  return x;
}

void BayesianOptimization::AddSample(const Vec& x, double y) { xs_.push_back(Normalize(x)); ys_.push_back(y); }
void BayesianOptimization::Clear() { xs_.clear(); ys_.clear(); }

double BayesianOptimization::ExpectedImprovement(const Vec& u, double best) const {
  double mu, sigma;
  gp_.Predict(u, &mu, &sigma);
  if (sigma < 1e-12) return 0.0;
  double imp = mu - best - xi_;
  double z = imp / sigma;
  return imp * NormCdf(z) + sigma * NormPdf(z);
}

Vec BayesianOptimization::NextSample(bool) {
  const size_t d = bounds_.size();
  std::uniform_real_distribution<double> uni(0.0, 1.0);
  if (ys_.empty()) {
    Vec u(d);
    for (auto& v : u) v = uni(rng_);
    return Denormalize(u);
  }
  gp_.Fit(xs_, ys_);
  double best = *std::max_element(ys_.begin(), ys_.end());
  Vec best_u(d, 0.5);
  double best_ei = -1;
  // 25 random restarts, each followed by a shrinking random local search
  for (int restart = 0; restart < 25; ++restart) {
    Vec u(d);
    for (auto& v : u) v = uni(rng_);
    double ei = ExpectedImprovement(u, best);
    double step = 0.25;
    for (int it = 0; it < 40; ++it) {
      Vec c = u;
      for (auto& v : c) v = std::min(1.0, std::max(0.0, v + (uni(rng_) - 0.5) * 2 * step));
      double e = ExpectedImprovement(c, best);
      if (e > ei) { ei = e; u = c; } else { step *= 0.9; }
    }
    if (ei > best_ei) { best_ei = ei; best_u = u; }
  }
  return Denormalize(best_u);
}

}  // namespace hvd

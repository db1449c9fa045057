// Autotuner for the runtime knobs.  Scores a configuration by negotiated
// bytes per microsecond (median of several samples of `steps_per_sample`
// training steps after warm-up) and searches:
//   * (fusion threshold MiB, cycle time ms) jointly with Bayesian optimisation
//   * categorical knobs by coordinate sweep: response cache on/off and — new
//     for the NVLink backend — the one-shot/two-shot crossover size, the NVLS
//     minimum size and the number of communication CTAs (replacing the
//     reference's hierarchical/torus categorical knobs, which are meaningless on
//     one NVSwitch box).
// Only the coordinator tunes; Controller::SynchronizeParameters broadcasts the
// POD TunableParams.  Values fixed through the environment are not touched.
// Parity: horovod/common/parameter_manager.{h,cc}.
#pragma once
#include <chrono>
#include <fstream>
#include <memory>
#include <string>
#include <unordered_set>
#include <vector>
#include "controller.h"
#include "optim/bayesian_optimization.h"

namespace hvd {

class ParameterManager {
 public:
  ParameterManager();
  void ConfigureFromEnv();
  void Initialize(int rank, const std::string& log_file);
  void SetAutoTuning(bool active) { active_ = active; params_.active = active; }
  bool IsAutoTuning() const { return active_; }

  // Called by the coordinator after each executed response. Returns true when
  // the parameters changed and must be re-broadcast.
  bool Update(const std::vector<std::string>& tensor_names, int64_t bytes);

  const TunableParams& params() const { return params_; }
  void SetParams(const TunableParams& p) { params_ = p; active_ = p.active != 0; }
  void SetFusionThresholdBytes(int64_t b, bool fixed = false) { params_.fusion_threshold_bytes = b; fixed_fusion_ |= fixed; }
  void SetCycleTimeMs(double ms, bool fixed = false) { params_.cycle_time_ms = ms; fixed_cycle_ |= fixed; }
  void SetCacheEnabled(bool e, bool fixed = false) { params_.cache_enabled = e; fixed_cache_ |= fixed; }
  void SetOneshotMaxBytes(int64_t b, bool fixed = false) { params_.oneshot_max_bytes = b; fixed_oneshot_ |= fixed; }
  void SetNvlsMinBytes(int64_t b, bool fixed = false) { params_.nvls_min_bytes = b; fixed_nvls_ |= fixed; }
  void SetCommCtas(int32_t c, bool fixed = false) { params_.comm_ctas = c; fixed_ctas_ |= fixed; }
  void Reset();

 private:
  enum class Phase { WARMUP, BAYES, CATEGORICAL, DONE };
  void FinishSample(double score);
  void NextCandidate();
  void ApplyBest();
  void LogRow(double score);

  TunableParams params_, best_params_;
  double best_score_ = -1;
  bool active_ = false;
  bool fixed_fusion_ = false, fixed_cycle_ = false, fixed_cache_ = false, fixed_oneshot_ = false,
       fixed_nvls_ = false, fixed_ctas_ = false;
  int rank_ = 0;
  int warmups_ = 3, steps_per_sample_ = 10, max_bayes_samples_ = 20, samples_per_point_ = 3;
  double gp_noise_ = 0.8;

  Phase phase_ = Phase::WARMUP;
  int warmup_left_ = 3;
  std::unordered_set<std::string> seen_this_step_;
  int steps_ = 0;
  int64_t bytes_ = 0;
  std::chrono::steady_clock::time_point sample_start_;
  std::vector<double> point_scores_;
  int bayes_samples_ = 0;
  std::unique_ptr<BayesianOptimization> bayes_;
  // categorical sweep state
  int cat_index_ = 0, cat_value_ = 0;
  std::ofstream log_;
};

}  // namespace hvd

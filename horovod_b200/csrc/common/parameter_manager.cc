#include "parameter_manager.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include "env.h"
#include "logging.h"

namespace hvd {

namespace {
const int64_t kOneshotChoices[] = {64 << 10, 256 << 10, 512 << 10, 1 << 20, 2 << 20};
const int64_t kNvlsChoices[] = {256 << 10, 1 << 20, 4 << 20, 1ll << 40};
const int32_t kCtaChoices[] = {32, 64, 128};
constexpr int kNumCategorical = 4;  // cache, oneshot, nvls, ctas
}  // namespace

ParameterManager::ParameterManager() { Reset(); }

void ParameterManager::ConfigureFromEnv() {
  warmups_ = (int)EnvInt(HOROVOD_AUTOTUNE_WARMUP_SAMPLES, 3);
  steps_per_sample_ = (int)EnvInt(HOROVOD_AUTOTUNE_STEPS_PER_SAMPLE, 10);
  max_bayes_samples_ = (int)EnvInt(HOROVOD_AUTOTUNE_BAYES_OPT_MAX_SAMPLES, 20);
  gp_noise_ = EnvDouble(HOROVOD_AUTOTUNE_GAUSSIAN_PROCESS_NOISE, 0.8);
  Reset();
}

void ParameterManager::Reset() {
  phase_ = Phase::WARMUP;
  warmup_left_ = warmups_;
  steps_ = 0; bytes_ = 0; bayes_samples_ = 0; cat_index_ = 0; cat_value_ = 0;
  best_score_ = -1;
  point_scores_.clear(); seen_this_step_.clear();
  // x0 = fusion threshold in MiB, x1 = cycle time in ms
  bayes_.reset(new BayesianOptimization({{1.0, 256.0}, {0.02, 5.0}}, gp_noise_));
  sample_start_ = std::chrono::steady_clock::now();
}

void ParameterManager::Initialize(int rank, const std::string& log_file) {
  rank_ = rank;
  if (rank == 0 && !log_file.empty()) {
    log_.open(log_file, std::ios::out | std::ios::trunc);
    if (log_.good()) log_ << "fusion_threshold_mb,cycle_time_ms,cache_enabled,oneshot_max_kb,nvls_min_kb,comm_ctas,score_bytes_per_us" << std::endl;
  }
}

void ParameterManager::LogRow(double score) {
  if (!log_.is_open()) return;
  log_ << params_.fusion_threshold_bytes / 1048576.0 << "," << params_.cycle_time_ms << "," << (int)params_.cache_enabled << ","
       << params_.oneshot_max_bytes / 1024 << "," << params_.nvls_min_bytes / 1024 << "," << params_.comm_ctas << "," << score
       << std::endl;
}

bool ParameterManager::Update(const std::vector<std::string>& names, int64_t bytes) {
  if (!active_ || phase_ == Phase::DONE) return false;
  bool new_step = false;
  for (auto& n : names) {
    if (seen_this_step_.count(n)) { new_step = true; break; }
  }
  if (new_step) { seen_this_step_.clear(); ++steps_; }
  for (auto& n : names) seen_this_step_.insert(n);
  bytes_ += bytes;
  if (steps_ < steps_per_sample_) return false;
  auto now = std::chrono::steady_clock::now();
  double us = std::chrono::duration<double, std::micro>(now - sample_start_).count();
  double score = us > 0 ? (double)bytes_ / us : 0.0;
  steps_ = 0; bytes_ = 0; sample_start_ = now;
  TunableParams before = params_;
  FinishSample(score);
  return memcmp(&before, &params_, sizeof before) != 0;
}

void ParameterManager::FinishSample(double score) {
  if (phase_ == Phase::WARMUP) {
    if (--warmup_left_ <= 0) { phase_ = Phase::BAYES; }
    return;
  }
  point_scores_.push_back(score);
  if ((int)point_scores_.size() < samples_per_point_) return;
  std::sort(point_scores_.begin(), point_scores_.end());
  double med = point_scores_[point_scores_.size() / 2];
  point_scores_.clear();
  LogRow(med);
  if (med > best_score_) { best_score_ = med; best_params_ = params_; }
  if (phase_ == Phase::BAYES) {
    bayes_->AddSample({params_.fusion_threshold_bytes / 1048576.0, params_.cycle_time_ms}, med);
    ++bayes_samples_;
  }
  NextCandidate();
}

void ParameterManager::NextCandidate() {
  if (phase_ == Phase::BAYES) {
    if (bayes_samples_ < max_bayes_samples_ && !(fixed_fusion_ && fixed_cycle_)) {
      Vec x = bayes_->NextSample();
      if (!fixed_fusion_) params_.fusion_threshold_bytes = (int64_t)(std::round(x[0]) * 1048576.0);
      if (!fixed_cycle_) params_.cycle_time_ms = x[1];
      return;
    }
    phase_ = Phase::CATEGORICAL;
    cat_index_ = 0; cat_value_ = -1;
    params_ = best_params_;
  }
  if (phase_ == Phase::CATEGORICAL) {
    params_ = best_params_;  // coordinate sweep always restarts from the incumbent
    while (cat_index_ < kNumCategorical) {
      ++cat_value_;
      bool fixed = cat_index_ == 0 ? fixed_cache_ : cat_index_ == 1 ? fixed_oneshot_ : cat_index_ == 2 ? fixed_nvls_ : fixed_ctas_;
      int nchoices = cat_index_ == 0 ? 2 : cat_index_ == 1 ? 5 : cat_index_ == 2 ? 4 : 3;
      if (fixed || cat_value_ >= nchoices) { ++cat_index_; cat_value_ = -1; continue; }
      if (cat_index_ == 0) { if ((bool)cat_value_ == best_params_.cache_enabled) continue; params_.cache_enabled = (bool)cat_value_; }
      if (cat_index_ == 1) { if (kOneshotChoices[cat_value_] == best_params_.oneshot_max_bytes) continue; params_.oneshot_max_bytes = kOneshotChoices[cat_value_]; }
      if (cat_index_ == 2) { if (kNvlsChoices[cat_value_] == best_params_.nvls_min_bytes) continue; params_.nvls_min_bytes = kNvlsChoices[cat_value_]; }
      if (cat_index_ == 3) { if (kCtaChoices[cat_value_] == best_params_.comm_ctas) continue; params_.comm_ctas = kCtaChoices[cat_value_]; }
      return;
    }
    ApplyBest();
  }
}

void ParameterManager::ApplyBest() {
  params_ = best_params_;
  phase_ = Phase::DONE;
  active_ = false;
  params_.active = 0;
  LOG(INFO) << "autotune finished: fusion " << params_.fusion_threshold_bytes / 1048576.0 << " MiB, cycle "
            << params_.cycle_time_ms << " ms, cache " << (int)params_.cache_enabled << ", oneshot<= "
            << params_.oneshot_max_bytes << " B, nvls>= " << params_.nvls_min_bytes << " B, ctas " << params_.comm_ctas
            << ", score " << best_score_ << " B/us";
  if (log_.is_open()) log_.flush();
}

}  // namespace hvd

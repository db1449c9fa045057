#include "logging.h"
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>

namespace hvd {
namespace {
std::atomic<int> g_level{-1};
std::atomic<int> g_rank{-1};
std::atomic<bool> g_hide_time{false};

LogLevel ParseLevel(const char* s) {
  if (!s) return LogLevel::WARNING;
  std::string v(s);
  for (auto& c : v) c = (char)tolower(c);
  if (v == "trace") return LogLevel::TRACE;
  if (v == "debug") return LogLevel::DEBUG;
  if (v == "info") return LogLevel::INFO;
  if (v == "warning" || v == "warn") return LogLevel::WARNING;
  if (v == "error") return LogLevel::ERROR;
  if (v == "fatal") return LogLevel::FATAL;
  return LogLevel::WARNING;
}
}  // namespace

void ResetLogLevelFromEnv() {
  g_level = (int)ParseLevel(getenv("HOROVOD_LOG_LEVEL"));
  const char* h = getenv("HOROVOD_LOG_HIDE_TIME");
  g_hide_time = h && atoi(h) > 0;
}
LogLevel MinLogLevel() {
  if (g_level.load() < 0) ResetLogLevelFromEnv();
  return (LogLevel)g_level.load();
}
void SetLogRank(int rank) { g_rank = rank; }

LogMessage::LogMessage(const char* file, int line, LogLevel level, int rank)
    : file_(file), line_(line), level_(level), rank_(rank) {}

LogMessage::~LogMessage() {
  static const char* names = "TDIWEF";
  const char* base = strrchr(file_, '/');
  base = base ? base + 1 : file_;
  char tbuf[64] = "";
  if (!g_hide_time) {
    auto now = std::chrono::system_clock::now();
    std::time_t t = std::chrono::system_clock::to_time_t(now);
    auto us = std::chrono::duration_cast<std::chrono::microseconds>(now.time_since_epoch()).count() % 1000000;
    struct tm tmv; localtime_r(&t, &tmv);
    char d[32]; strftime(d, sizeof d, "%Y-%m-%d %H:%M:%S", &tmv);
    snprintf(tbuf, sizeof tbuf, "%s.%06ld: ", d, (long)us);
  }
  int r = rank_ >= 0 ? rank_ : g_rank.load();
  if (r >= 0) fprintf(stderr, "[%s%c %s:%d][%d] %s\n", tbuf, names[(int)level_], base, line_, r, str().c_str());
  else fprintf(stderr, "[%s%c %s:%d] %s\n", tbuf, names[(int)level_], base, line_, str().c_str());
  if (level_ == LogLevel::FATAL) abort();
}
}  // namespace hvd

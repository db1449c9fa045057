// Stream-style logging: LOG(LEVEL) << ...; level from HOROVOD_LOG_LEVEL
// (trace|debug|info|warning|error|fatal, default warning), timestamps hidden
// with HOROVOD_LOG_HIDE_TIME.  Parity: horovod/common/logging.{h,cc}.
#pragma once
#include <sstream>
#include <string>

namespace hvd {
enum class LogLevel { TRACE = 0, DEBUG = 1, INFO = 2, WARNING = 3, ERROR = 4, FATAL = 5 };

class LogMessage : public std::basic_ostringstream<char> {
 public:
  LogMessage(const char* file, int line, LogLevel level, int rank = -1);
  ~LogMessage();
 private:
  const char* file_; int line_; LogLevel level_; int rank_;
};
LogLevel MinLogLevel();
void ResetLogLevelFromEnv();
void SetLogRank(int rank);

#define HVD_LOG_TRACE ::hvd::LogLevel::TRACE
#define HVD_LOG_DEBUG ::hvd::LogLevel::DEBUG
#define HVD_LOG_INFO ::hvd::LogLevel::INFO
#define HVD_LOG_WARNING ::hvd::LogLevel::WARNING
#define HVD_LOG_ERROR ::hvd::LogLevel::ERROR
#define HVD_LOG_FATAL ::hvd::LogLevel::FATAL
#define LOG(level) \
  if (HVD_LOG_##level >= ::hvd::MinLogLevel()) ::hvd::LogMessage(__FILE__, __LINE__, HVD_LOG_##level)
}  // namespace hvd

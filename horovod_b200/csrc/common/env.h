// Typed getenv helpers.  Parity: horovod/common/utils/env_parser.{h,cc}.
#pragma once
#include <cstdint>
#include <string>
namespace hvd {
bool EnvIsSet(const char* name);
int64_t EnvInt(const char* name, int64_t dflt);
double EnvDouble(const char* name, double dflt);
bool EnvBool(const char* name, bool dflt);
std::string EnvStr(const char* name, const std::string& dflt = "");
}  // namespace hvd

#include "env.h"
#include <cstdlib>
#include <cstring>
namespace hvd {
bool EnvIsSet(const char* n) { const char* v = getenv(n); return v && *v; }
int64_t EnvInt(const char* n, int64_t d) { const char* v = getenv(n); return (v && *v) ? strtoll(v, nullptr, 10) : d; }
double EnvDouble(const char* n, double d) { const char* v = getenv(n); return (v && *v) ? strtod(v, nullptr) : d; }
bool EnvBool(const char* n, bool d) {
  const char* v = getenv(n);
  if (!v || !*v) return d;
  return !(strcmp(v, "0") == 0 || strcasecmp(v, "false") == 0 || strcasecmp(v, "no") == 0 || strcasecmp(v, "off") == 0);
}
std::string EnvStr(const char* n, const std::string& d) { const char* v = getenv(n); return (v && *v) ? std::string(v) : d; }
}  // namespace hvd

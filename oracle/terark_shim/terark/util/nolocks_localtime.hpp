#pragma once
#include "../stdtypes.hpp"
#include <time.h>
namespace terark {
inline struct tm* nolocks_localtime_r(const time_t* t, struct tm* r) { return localtime_r(t, r); }
inline const char* StrDateTimeNow() {
  static thread_local char buf[64];
  time_t t = time(nullptr); struct tm r; localtime_r(&t, &r);
  strftime(buf, sizeof buf, "%F %T", &r);
  return buf;
}
}

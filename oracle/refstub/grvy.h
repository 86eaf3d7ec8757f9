/* Test infrastructure only: minimal stand-in for libGRVY's logging (reference uses
 * grvy_printf(GRVY_ERROR, ...), src/logger.hpp:37).  Nothing here is shipped. */
#pragma once
#include <cstdarg>
#include <cstdio>
#define GRVY_ERROR 0
#define GRVY_WARN 1
#define GRVY_INFO 2
#define GRVY_DEBUG 3
#define GINFO GRVY_INFO
#define GERROR GRVY_ERROR
#define GWARN GRVY_WARN
#define GDEBUG GRVY_DEBUG
static inline int grvy_printf(int, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  int r = vprintf(fmt, ap);
  va_end(ap);
  return r;
}

// TEST INFRASTRUCTURE -- error channel for the emulation library (the product's lives in jt_api.cu).
#include <cstdarg>
#include <cstdio>
#include <string>
static thread_local std::string g_error;
int jt_set_error(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_error = buf;
  return code;
}
extern "C" const char* jt_last_error(void) { return g_error.c_str(); }

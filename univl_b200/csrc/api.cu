// univl_b200 — C-ABI plumbing shared by every kernel file: error strings and version.
#include "common.cuh"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

namespace univl {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("UNIVL_PDL");
    v = (e && e[0] == '1') ? 1 : 0;  // measured neutral on the FT-Align step (graph replay): opt-in
  }
  return v == 1;
}

}  // namespace univl

extern "C" const char* univl_last_error_string() { return univl::g_err; }
extern "C" int univl_abi_version() { return 1; }

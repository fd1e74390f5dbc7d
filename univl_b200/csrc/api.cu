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

// SMs a concurrently running collective kernel (NCCL) occupies while the kernels being enqueued run: persistent grids are
// sized to the SMs that are actually free, otherwise the CTAs that find no SM run as a second wave and double the time
// of every persistent kernel under the collective.  Read at enqueue time, so it is baked into captured graphs.
static int g_reserved_sms = 0;

int usable_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0, n = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    sms = n;
  }
  const int n = sms - g_reserved_sms;
  return n < 8 ? 8 : n;
}

// CTA pairs need both SMs of a TPC: every reserved SM may sit in a different TPC
int usable_sm_pairs() {
  const int all = (usable_sms() + g_reserved_sms) / 2;
  const int n = all - g_reserved_sms;
  return n < 4 ? 4 : n;
}

}  // namespace univl

extern "C" int univl_set_reserved_sms(int n) {
  if (n < 0) return univl::set_error(UNIVL_ERR_ARG, "univl_set_reserved_sms: n = %d", n);
  univl::g_reserved_sms = n;
  return UNIVL_OK;
}
extern "C" const char* univl_last_error_string() { return univl::g_err; }
extern "C" int univl_abi_version() { return 1; }

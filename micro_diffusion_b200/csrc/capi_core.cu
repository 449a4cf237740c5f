// Error bookkeeping and version query of the C ABI.
#include <string.h>

#include "common.cuh"

namespace md {
static thread_local char g_err[512] = "";
int md_set_error(int code, const char* msg) {
  strncpy(g_err, msg ? msg : "", sizeof(g_err) - 1);
  g_err[sizeof(g_err) - 1] = 0;
  return code;
}
}  // namespace md

extern "C" const char* md_last_error(void) { return md::g_err; }
extern "C" int md_abi_version(void) { return 4; }

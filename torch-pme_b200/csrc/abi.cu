// ABI bookkeeping: version and per-thread error message.
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "../../include/torchpme_b200.h"

namespace tpme {
static thread_local char g_last_error[512] = "";
void set_last_error(const char* where, const char* what) {
  snprintf(g_last_error, sizeof(g_last_error), "%s: %s", what ? what : "error", where ? where : "");
}
bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("TPME_PDL"); return !(e != nullptr && e[0] == '0'); }();
  return on;
}
bool pdl_everywhere() {
  static const bool on = [] { const char* e = getenv("TPME_PDL"); return e != nullptr && strcmp(e, "all") == 0; }();
  return on;
}
}  // namespace tpme

extern "C" int tpme_abi_version(void) { return TPME_ABI_VERSION; }
extern "C" const char* tpme_last_error(void) { return tpme::g_last_error; }

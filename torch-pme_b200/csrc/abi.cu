// ABI bookkeeping: version and per-thread error message.
#include <cstring>

#include "common.cuh"
#include "../../include/torchpme_b200.h"

namespace tpme {
static thread_local char g_last_error[512] = "";
void set_last_error(const char* where, const char* what) {
  snprintf(g_last_error, sizeof(g_last_error), "%s: %s", what ? what : "error", where ? where : "");
}
}  // namespace tpme

extern "C" int tpme_abi_version(void) { return TPME_ABI_VERSION; }
extern "C" const char* tpme_last_error(void) { return tpme::g_last_error; }

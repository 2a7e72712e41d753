// Thread-local error string + ABI version for libscvae_b200.
#include <stdarg.h>

#include "common.cuh"

namespace scvae {
static thread_local char g_error[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}
}  // namespace scvae

extern "C" int scvae_abi_version(void) { return SCVAE_B200_ABI_VERSION; }
extern "C" const char *scvae_last_error(void) { return scvae::g_error; }
